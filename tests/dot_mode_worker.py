"""Worker of tests/test_dot_mode.py::test_gpu_other_dot_modes: runs in a process whose FIZZ_DOT_MODE / FZO_DOT_MODE select
the library and the oracle of ONE numpy dot/norm contract and checks the CUDA engine against that oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fizz2d_b200 as fz  # noqa: E402
from helpers import bits, oracle_from_spec, orc  # noqa: E402

mode = int(os.environ["FIZZ_DOT_MODE"])
assert fz.lib.load().fizz_dot_mode() == mode and orc.lib().fzo_dot_mode() == mode
for scene, n, seed, steps in (("RAMPWORLD_2TRIANGLES_5SQUARES", 256, 20260003, 300), ("CORRIDOR_2SQUARES", 128, 5, 300)):
    for tune in (("hybrid", 64), ("thread", 1)):
        bw = fz.BatchedWorld.from_in(fz.scenes.scene_path(scene), n_worlds=n, perturb=True, seed=seed, max_calls=256)
        bw.set_tuning(*tune)
        bw.step(steps)
        com, heat, st, sd, calls = bw.com(), bw.heat(), bw.status(), bw.steps_done(), bw.calls()
        spec = bw.groups[0].spec
        worlds = [oracle_from_spec(spec, w) for w in range(n)]
        orc.run_batch(worlds, steps, os.cpu_count() or 1)
        for w, ow in enumerate(worlds):
            assert ow.steps_done() == sd[w] and ow.status() == st[w] and ow.total_calls() == calls[w], (mode, scene, tune, w)
            assert np.array_equal(bits(ow.com()), bits(com[w])) and bits(ow.heat()) == bits(heat[w]), (mode, scene, tune, w)
print("dot mode %d OK" % mode)
