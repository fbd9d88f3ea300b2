"""Each of the heaviest worlds of the benchmark batch stepped ALONE for the benchmark's 10 000 timesteps: time, calls,
share of calls in the fast path -- which chain really bounds a pass?"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
from fizz2d_b200.batch import GroupSpec
heavy = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "63739,12860,58248,44270,36291,36227,47547,29797").split(",")]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
sc = fz.load_in(fz.scenes.scene_path("RAMPWORLD_2TRIANGLES_5SQUARES"))
d = fz.perturbation_deltas(20260003, 65536 * 8, sc.n_free)
for wi in heavy:
    bw = fz.BatchedWorld([GroupSpec.from_scene(sc, 1, np.ascontiguousarray(d[[wi]]))])
    bw.step(1); bw.upload(); bw.counters(); torch.cuda.synchronize()
    t0 = time.perf_counter(); bw.step(steps); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    k = bw.counters()[0]
    print("world %5d: %.1f ms  calls %d  fast %.1f%%  general calls %d  us/call %.3f" % (
        wi, dt * 1e3, k["contact_calls"], 100.0 * k["fast_path_calls"] / max(1, k["contact_calls"]),
        k["contact_calls"] - k["fast_path_calls"], dt * 1e6 / max(1, k["contact_calls"])))
    bw.close()
