"""Throughput phase of the contact kernel: first chunk of the benchmark batch (every world inside the scene)."""
import sys, time
import torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
bw = fz.BatchedWorld.from_in(fz.scenes.scene_path("RAMPWORLD_2TRIANGLES_5SQUARES"), n_worlds=n, perturb=True, seed=20260003)
snap = bw.snapshot()
for rep in range(3):
    bw.restore(snap)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bw.step(256)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    bw.step(256)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("chunk1 %.1f ms  chunk2 %.1f ms  regs %s" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, bw.kernel_info()[0]["contact"]))
