"""ms per 10 000-step pass of the benchmark batch for several base chunk lengths (the schedule doubles every second
chunk up to 16x the base)."""
import sys, time
import torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
bw = fz.BatchedWorld.from_in(fz.scenes.scene_path("RAMPWORLD_2TRIANGLES_5SQUARES"), n_worlds=65536, perturb=True, seed=20260003)
snap = bw.snapshot()
for chunk in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "256,128,64,32,16,8").split(",")]:
    bw.set_tuning("hybrid", chunk)
    out = []
    for rep in range(3):
        bw.restore(snap); torch.cuda.synchronize()
        t0 = time.perf_counter(); bw.step(10000); torch.cuda.synchronize(); out.append((time.perf_counter() - t0) * 1e3)
    print("chunk %4d: %s ms" % (chunk, " ".join("%.1f" % t for t in out)), bw.counters()[0])
