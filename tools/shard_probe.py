"""Every shard of an N-GPU weak-scaling run of a BASELINE config, one after the other on ONE GPU: pass time, the heaviest
worlds (check_collisions() calls), fast-path share -- which shard's chain bounds the N-GPU pass, and why."""
import sys, time, types
import numpy as np, torch
sys.path.insert(0, ".")
import fizz2d_b200 as fz
import bench
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
config = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ranks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else list(range(G))
cfg = bench.CONFIGS[config]
if cfg["kind"] == "ga":            # the population law is sequential: draw the whole population once, not once per shard
    from fizz2d_b200 import ga
    _pop = ga.random_convex_polygons(cfg["seed"], cfg["worlds"] * G)
    ga.random_convex_polygons = lambda seed, n: _pop
args = types.SimpleNamespace(config=config, worlds=cfg["worlds"], max_calls=4096)
for r in ranks:
    specs, ids = bench.build_specs(args, r, G)
    bw = fz.BatchedWorld(specs, ids)
    bw.set_tuning(cfg["mode"], cfg["chunk"])
    if cfg.get("spec") and cfg["mode"] == "hybrid":
        bw.set_spec(**cfg["spec"])
    snap = bw.snapshot()
    import ctypes as C
    gstat = [torch.zeros(g.spec.n_worlds, dtype=torch.int64, device="cuda") for g in bw.groups]
    for g, t in zip(bw.groups, gstat):
        fz.lib.check(fz.lib.load().fizz_set_world_stats(g.handle, C.c_void_p(t.data_ptr())))
    ts = []
    for rep in range(2):
        bw.restore(snap); bw.counters(reset=True)
        for t in gstat:
            t.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); bw.step(cfg["timesteps"]); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    c, st, sd = bw.calls(), bw.status(), bw.steps_done()
    o = np.argsort(-c)[:5]
    k = bw.counters()
    print("rank %d/%d: pass %.1f ms | heaviest worlds (local idx, calls, steps done, status): %s | fast %.3f of %d contact calls" % (
        r, G, ts[-1], [(int(i), int(c[i]), int(sd[i]), int(st[i])) for i in o],
        sum(x["fast_path_calls"] for x in k) / max(1, sum(x["contact_calls"] for x in k)), sum(x["contact_calls"] for x in k)), flush=True)
    for gi, g in enumerate(bw.groups):
        gc = gstat[gi].cpu().numpy()
        og = np.argsort(-gc)[:3]
        print("   group %d (%s vertices): most general-path calls (local idx, general calls, all calls, steps, status):" % (gi, g.spec.nverts_free),
              [(int(g.world_ids[i]), int(gc[i]), int(c[g.world_ids[i]]), int(sd[g.world_ids[i]]), int(st[g.world_ids[i]])) for i in og])
    bw.close()
    del bw
