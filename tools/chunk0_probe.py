"""Who bounds the first chunk?  Heaviest worlds of the benchmark batch after 256 steps, then each of them stepped alone
from t = 0: time, calls, share of calls made in the single-contact fast path."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
from fizz2d_b200.batch import GroupSpec
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 256
path = fz.scenes.scene_path("RAMPWORLD_2TRIANGLES_5SQUARES")
bw = fz.BatchedWorld.from_in(path, n_worlds=65536, perturb=True, seed=20260003)
torch.cuda.synchronize(); t0 = time.perf_counter(); bw.step(steps); torch.cuda.synchronize()
print("batch: %d steps in %.1f ms" % (steps, (time.perf_counter() - t0) * 1e3), bw.counters()[0])
c = bw.calls(); st = bw.status()
o = np.argsort(-c)
print("top calls", [(int(i), int(c[i]), int(st[i])) for i in o[:10]])
sc = fz.load_in(path)
d = fz.perturbation_deltas(20260003, 65536, 7)
for wi in [int(i) for i in o[:6]]:
    one = fz.BatchedWorld([GroupSpec.from_scene(sc, 1, np.ascontiguousarray(d[[wi]]))])
    one.step(1); one.counters()   # warm-up launch
    one = fz.BatchedWorld([GroupSpec.from_scene(sc, 1, np.ascontiguousarray(d[[wi]]))])
    torch.cuda.synchronize(); t0 = time.perf_counter(); one.step(steps); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    k = one.counters()[0]
    print("world %5d alone: %.1f ms  calls %d  fast %d (%.0f%%)  status %d  us/call %.3f" % (
        wi, dt, k["contact_calls"], k["fast_path_calls"], 100.0 * k["fast_path_calls"] / max(1, k["contact_calls"]),
        int(one.status()[0]), dt * 1e3 / max(1, k["contact_calls"])))
