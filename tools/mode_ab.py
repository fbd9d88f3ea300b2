"""Same-box A/B of execution modes on the benchmark batch: ms per 10 000-step pass."""
import sys, time
import torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
modes = (sys.argv[1] if len(sys.argv) > 1 else "hybrid_sync:256,hybrid:256,hybrid:64").split(",")
bw = fz.BatchedWorld.from_in(fz.scenes.scene_path("RAMPWORLD_2TRIANGLES_5SQUARES"), n_worlds=65536, perturb=True, seed=20260003)
snap = bw.snapshot()
ref = None
for rnd in range(2):
    for m in modes:
        mode, chunk = m.split(":")
        bw.set_tuning(mode, int(chunk))
        out = []
        for rep in range(3):
            bw.restore(snap); torch.cuda.synchronize()
            t0 = time.perf_counter(); bw.step(10000); torch.cuda.synchronize(); out.append((time.perf_counter() - t0) * 1e3)
        sig = (bw.com().tobytes(), bw.calls().tobytes(), bw.status().tobytes())
        ref = ref or sig
        print("%-18s %s ms   same result: %s" % (m, " ".join("%.1f" % t for t in out), sig == ref))
