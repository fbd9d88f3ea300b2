"""The throughput phase alone: the benchmark batch, first N timesteps (every world inside the scene), twice (the second
launch of the contact kernel is the one to profile: `ncu -k regex:contact_kernel -s 1 -c 1`)."""
import sys, time
import torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 96
mode = sys.argv[2] if len(sys.argv) > 2 else "hybrid"
bw = fz.BatchedWorld.from_in(fz.scenes.scene_path("RAMPWORLD_2TRIANGLES_5SQUARES"), n_worlds=65536, perturb=True, seed=20260003)
bw.set_tuning(mode, steps)
snap = bw.snapshot()
for rep in range(2):
    bw.restore(snap); torch.cuda.synchronize()
    bw.counters(reset=True)
    t0 = time.perf_counter(); bw.step(steps); torch.cuda.synchronize()
    print("first %d timesteps: %.2f ms  %s" % (steps, (time.perf_counter() - t0) * 1e3, bw.counters()[0]), flush=True)
