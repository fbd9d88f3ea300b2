"""Throughput of the other BASELINE configs in both modes (not bench lines; context for DESIGN.md)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz

CHUNKS = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [256]
MODES = sys.argv[3].split(",") if len(sys.argv) > 3 else ["hybrid"]
def run(name, make, steps):
    for mode, chunk in [(m, c) for m in MODES for c in CHUNKS]:
        bw = make()
        bw.set_tuning(mode, chunk)
        snap = bw.snapshot()
        bw.step(steps); torch.cuda.synchronize()
        bw.restore(snap); torch.cuda.synchronize()
        bw.counters(reset=True); bw.profile(reset=True); bw.set_profiling(True)
        t0 = time.perf_counter(); bw.step(steps); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        prof = bw.profile(reset=True); cnt = bw.counters(reset=True); bw.set_profiling(False)
        print("   ", [(dict((k, round(v[0], 1)) for k, v in p.items() if v[1]), c) for p, c in zip(prof, cnt)])
        sd = bw.steps_done(); st = bw.status()
        print("%-34s %-7s chunk=%d worlds=%d steps=%d  %.1f ms  %.3g world-steps/s  capped=%d calls/ws=%.2f" % (
            name, mode, chunk, bw.n_worlds, steps, dt * 1e3, sd.sum() / dt, int((st != 0).sum()), bw.calls().sum() / sd.sum()))
        bw.close()

P = fz.scenes.scene_path
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("2", "all"):
    run("config2 BOX_1SQUAREFRICTION x4096", lambda: fz.BatchedWorld.from_in(P("BOX_1SQUAREFRICTION"), 4096, perturb=True, seed=20260002), 1000)
if which in ("4", "all"):
    run("config4 mixed x262144", lambda: fz.BatchedWorld.from_mixed([P("2CHAMBERS_3BALLS"), P("STACKEDCHAMBERS_1SQUARE")], 262144, seed=20260004), 1000)
if which in ("3", "all"):
    run("config3 RAMPWORLD x65536", lambda: fz.BatchedWorld.from_in(P("RAMPWORLD_2TRIANGLES_5SQUARES"), 65536, perturb=True, seed=20260003), 1000)
if which in ("5", "all"):
    pop = fz.ga.random_convex_polygons(20260005, 131072)
    run("config5 GA x131072 (1/8 of 1M)", lambda: fz.BatchedWorld.from_ga(P("PENTAGONVOID_1SQUARE"), 131072, population=pop), 2000)
