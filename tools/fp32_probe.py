"""FP32 mode against the FP64 path: per-world divergence horizon (first timestep whose check_collisions() count differs:
a contact decision went the other way), position error before it, energy drift.  Input for the bound in DESIGN.md and
tests/test_gpu_fp32.py."""
import sys
import numpy as np
sys.path.insert(0, ".")
import fizz2d_b200 as fz

def run(scene, seed, n, steps, mode, mc=256):
    bw = fz.BatchedWorld.from_in(fz.scenes.scene_path(scene), n_worlds=n, perturb=True, seed=seed, max_calls=mc)
    bw.set_tuning(mode, 256)
    calls, com, en = [], [], [bw.energy()]
    for t in range(steps):
        bw.step(1)
        calls.append(bw.calls()); com.append(bw.com()); en.append(bw.energy())
    return np.array(calls), np.array(com), np.array(en), bw.status()

for scene, seed, n, steps in (("BOX_1SQUAREFRICTION", 20260002, 512, 300), ("RAMPWORLD_2TRIANGLES_5SQUARES", 20260003, 256, 300)):
    c64, x64, e64, s64 = run(scene, seed, n, steps, "thread")
    c32, x32, e32, s32 = run(scene, seed, n, steps, "thread_fp32")
    diff = (c64 != c32)                                   # [steps, n]
    horizon = np.where(diff.any(axis=0), diff.argmax(axis=0), steps)
    print(scene, "worlds", n, "steps", steps)
    print("  divergence horizon (timesteps): min %d  p10 %d  median %d  p90 %d; never diverged: %d worlds" % (
        horizon.min(), np.percentile(horizon, 10), np.median(horizon), np.percentile(horizon, 90), (horizon == steps).sum()))
    err = []
    for w in range(n):
        h = horizon[w]
        if h > 0:
            d = np.abs(x32[:h, w, :, 0:2] - x64[:h, w, :, 0:2]).max()
            err.append(d / max(1.0, np.abs(x64[:h, w, :, 0:2]).max()))
    err = np.array(err)
    print("  max |pos32 - pos64| / max|pos| before the horizon: median %.2e  p99 %.2e  max %.2e" % (np.median(err), np.percentile(err, 99), err.max()))
    ok = (s32 == 0) & (e32[0, :, 0] > 0)
    print("  status fp32:", np.unique(s32, return_counts=True), "fp64:", np.unique(s64, return_counts=True))
    if not ok.any():
        continue
    drift = np.abs(e32[-1, ok, 0] - e32[0, ok, 0]) / e32[0, ok, 0]
    print("  FP32 total-energy drift over the run (uncapped worlds): median %.2e  max %.2e; capped fp32 %d fp64 %d" % (
        np.median(drift), drift.max(), (s32 != 0).sum(), (s64 != 0).sum()))
