"""Debug aid: a perturbed batch stepped by the engine (spec rounds on) against the CPU oracle; prints the worlds that differ."""
import os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
import fizz2d_b200 as fz
from helpers import oracle_from_spec, orc, bits
from conftest import scene_file
name, seed, n, steps, mc, chunk = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
bw = fz.BatchedWorld.from_in(scene_file(name), n_worlds=n, perturb=True, seed=seed, max_calls=mc)
bw.set_tuning("hybrid", chunk)
bw.step(steps // 2); bw.step(steps - steps // 2)
com, st, sd, calls = bw.com(), bw.status(), bw.steps_done(), bw.calls()
spec = bw.groups[0].spec
worlds = [oracle_from_spec(spec, w) for w in range(n)]
orc.run_batch(worlds, steps, os.cpu_count() or 1)
bad = 0
for w, ow in enumerate(worlds):
    ok = ow.steps_done() == sd[w] and ow.status() == st[w] and np.array_equal(bits(ow.com()), bits(com[w])) and ow.total_calls() == calls[w]
    if not ok:
        bad += 1
        if bad <= 8:
            print("world", w, "steps", sd[w], ow.steps_done(), "status", st[w], ow.status(), "calls", calls[w], ow.total_calls(),
                  "com equal", np.array_equal(bits(ow.com()), bits(com[w])))
print("mismatching worlds:", bad, "of", n)
