"""ms per 256-step call over the whole benchmark batch (where does a pass spend its time?)."""
import sys, time
import torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
bw = fz.BatchedWorld.from_in(fz.scenes.scene_path("RAMPWORLD_2TRIANGLES_5SQUARES"), n_worlds=65536, perturb=True, seed=20260003)
snap = bw.snapshot()
for rep in range(2):
    bw.restore(snap); torch.cuda.synchronize()
    ts = []
    t00 = time.perf_counter()
    for c in range(40):
        t0 = time.perf_counter(); bw.step(256 if c < 39 else 10000 - 39 * 256); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    tot = (time.perf_counter() - t00) * 1e3
print("total %.1f ms; per chunk:" % tot, " ".join("%.1f" % t for t in ts))
