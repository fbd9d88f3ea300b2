"""Latency probe: step a handful of hand-picked (trapped) worlds alone and report microseconds per
check_collisions()-equivalent -- the sequential depth that bounds a whole batch pass."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import fizz2d_b200 as fz
from fizz2d_b200.batch import GroupSpec

worlds = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "103,532,1555,1188").split(",")]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
sc = fz.load_in(fz.scenes.scene_path("RAMPWORLD_2TRIANGLES_5SQUARES"))
d = fz.perturbation_deltas(20260003, 65536, 7)[worlds]
for mode in ("hybrid", "thread"):
    for sel in ([0], list(range(len(worlds)))):
        bw = fz.BatchedWorld([GroupSpec.from_scene(sc, len(sel), np.ascontiguousarray(d[sel]))])
        bw.set_tuning(mode, 256)
        bw.step(10)
        bw.upload()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bw.step(steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        calls = bw.calls()
        print("%-7s worlds=%s steps=%d time=%.1f ms calls(max)=%d  us/call=%.3f" % (
            mode, [worlds[i] for i in sel], steps, dt * 1e3, calls.max(), dt * 1e6 / calls.max()))
