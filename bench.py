#!/usr/bin/env python
"""bench.py -- world-steps/sec of the batched Fizz-2D stepping engine on B200 (BASELINE.json metric).

Headline workload (config.workload): BASELINE config 3 -- RAMPWORLD_2TRIANGLES_5SQUARES.in x 65536 perturbed worlds
(seed PCG64(20260003), SURVEY 8d law) x 10000 timesteps per GPU (weak scaling: every rank steps its own 65536
worlds of an N x 65536 population; no inter-GPU traffic while stepping, one NCCL all_gather of per-world
[E,K,P,H,fitness,status] per pass).  One bench "step" = one pass of the hot path: reset to the initial conditions,
advance every world by all timesteps.  `value` = world-steps completed / device time with inputs resident in HBM;
`e2e` = the same through the public API with HOST (pinned) buffers: H2D of initial conditions and constants, stepping,
D2H of the per-world energy/status/step counters, all inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|3|4|5]
                  [--worlds W] [--timesteps T] [--no-extra] [--sweep-worlds W1,W2,..]

`--config` selects another BASELINE config as the workload of the line (2: BOX_1SQUAREFRICTION x 4096 x 1000,
4: 2CHAMBERS+STACKEDCHAMBERS mixed x 262144 x 1000, 5: GA polygons in PENTAGONVOID x 131072/GPU x 2000); the default run
(config 3, N = 1) also measures those three briefly and reports them under `extra.configs`.

`--impl reference` times the reference path on the host cores: the reference's OWN Python `World.update()`
(`cpu_baseline.kind` "reference": unmodified sources, shipped to the box under the git-ignored oracle/_ref/pyref/ by
`make -C oracle ref`) on a bounded sample of the workload, all cores, and -- once, next to it -- the scalar C port
(oracle/fizz_oracle.c).  That arm never imports the product package and never loads its .so.
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "world_steps_per_sec"
UNIT = "world-steps/s"

CONFIGS = {
    2: dict(kind="scene", scenes=["BOX_1SQUAREFRICTION"], worlds=4096, timesteps=1000, seed=20260002,
            mode="hybrid", chunk=256, label="BASELINE config 2"),
    # (spec: tuning of the time-parallel spec rounds, DESIGN.md 5d, like `chunk` a schedule knob that cannot change a
    #  result -- worlds take part from a cost of 100 units per timestep (library default 48): measured 258 -> 237 ms per pass)
    3: dict(kind="scene", scenes=["RAMPWORLD_2TRIANGLES_5SQUARES"], worlds=65536, timesteps=10000, seed=20260003,
            mode="hybrid", chunk=96, spec=dict(min_cost=100), label="BASELINE config 3"),
    4: dict(kind="mixed", scenes=["2CHAMBERS_3BALLS", "STACKEDCHAMBERS_1SQUARE"], worlds=262144, timesteps=1000,
            seed=20260004, mode="hybrid3", chunk=32, label="BASELINE config 4"),
    5: dict(kind="ga", scenes=["PENTAGONVOID_1SQUARE"], worlds=131072, timesteps=2000, seed=20260005,
            mode="hybrid", chunk=256, label="BASELINE config 5 (1/8 of 1M per GPU)"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--worlds", type=int, default=None, help="worlds per GPU (default: the config's)")
    ap.add_argument("--timesteps", type=int, default=None)
    ap.add_argument("--chunk", type=int, default=None, help="timesteps per chunk")
    ap.add_argument("--mode", default=None, choices=["hybrid", "hybrid_sync", "hybrid3", "thread"])
    ap.add_argument("--max-calls", type=int, default=4096)
    ap.add_argument("--cpu-sample-worlds", type=int, default=16384, help="worlds of the C-port sample")
    ap.add_argument("--pyref-worlds", type=int, default=64, help="worlds of the Python-reference sample")
    ap.add_argument("--pyref-timesteps", type=int, default=600,
                    help="timesteps of the Python-reference sample (the first ~300 of config 3 are the contact-heavy "
                         "phase, ~4x slower per step in the reference than the quiescent rest)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip extra.configs (configs 2/4/5 beside the headline)")
    ap.add_argument("--sweep-worlds", default="", help="comma list of worlds/GPU: extra.worlds_sweep (value vs batch)")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    for k in ("worlds", "timesteps", "chunk", "mode"):
        if getattr(a, k) is None:
            setattr(a, k, cfg[k])
    return a


def workload_config(args, n_gpus):
    cfg = CONFIGS[args.config]
    d = {"workload": "%s: %s x %d perturbed worlds/GPU x %d timesteps (seed PCG64(%d), %s; watchdog max_calls=%d)"
                     % (cfg["label"], " + ".join(s + ".in" for s in cfg["scenes"]), args.worlds, args.timesteps, cfg["seed"],
                        "random convex polygons, fizz2d_b200.ga law" if cfg["kind"] == "ga" else "dpos 5 px, dvel 30 px/s per body",
                        args.max_calls),
         "baseline_config": args.config, "scene": " + ".join(s + ".in" for s in cfg["scenes"]),
         "worlds_per_gpu": args.worlds, "timesteps": args.timesteps, "max_calls": args.max_calls,
         "sharding": "world w of the %d-world population lives on rank w %% %d; no collective while stepping"
                     % (args.worlds * n_gpus, n_gpus),
         "l2": "L2 flushed (512 MiB write) before every pass"}
    return d


# ----------------------------------------------------------------------------------------------------------
# CPU arms (test infrastructure under oracle/; the only places bench.py executes it)
# ----------------------------------------------------------------------------------------------------------
def _cpu_arm():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_arm
    return cpu_arm


def port_from_specs(specs, n_sample, timesteps, threads):
    """C port on the first worlds of every group (proportional sample), fed with the product loader's constants: the
    op-counting leg of the GPU arm (ops per world-step for the roofline) and the `port` CPU figure."""
    ca = _cpu_arm()
    import oracle as orc
    total = sum(s.n_worlds for s in specs)
    ws = []
    for s in specs:
        n = max(1, min(s.n_worlds, int(round(n_sample * s.n_worlds / total))))
        p = s.params
        prm = orc.make_params(s.scene.params, s.height, p.max_calls, (p.gx, p.gy))
        for w in range(n):
            ws.append(orc.OracleWorld(s.nverts_free, s.nverts_fixed,
                                      np.concatenate([s.pos[:, w].reshape(-1, 2), s.fixed_com]),
                                      s.vel[:, w].reshape(-1, 2), np.concatenate([s.radius[:, w], s.fixed_radius]),
                                      s.mass[:, w], s.off[0::2, w], s.off[1::2, w], s.fixed_xy, prm))
    r = ca.port_run(None, None, None, None, timesteps, threads, None, worlds=ws)
    return r


def pyref_sample(args, n_total):
    """The unmodified Python reference on the first `--pyref-worlds` worlds x `--pyref-timesteps` timesteps of the
    workload, one process per host core.  None when the reference sources are not on this machine, or for workloads
    whose population is not a perturbed .in scene."""
    ca = _cpu_arm()
    cfg = CONFIGS[args.config]
    if cfg["kind"] != "scene" or not ca.pyref_available():
        return None
    procs = os.cpu_count() or 1
    n = min(args.pyref_worlds, args.worlds)
    pool = ca.PyRefPool(cfg["scenes"][0], cfg["seed"], n_total, list(range(n)), args.max_calls, procs)
    return pool, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])      # the checker's own C file; no product build
    ca = _cpu_arm()
    cfg = CONFIGS[args.config]
    threads = os.cpu_count() or 1
    n_total = args.worlds * max(1, args.gpus)
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, args.gpus)}
    port = None
    if cfg["kind"] == "scene":
        # the C port, once, on the FULL per-GPU batch of the workload (informational, beside the reference)
        t0 = time.perf_counter()
        ws = ca.port_worlds(cfg["scenes"][0], cfg["seed"], n_total, list(range(args.worlds)), args.max_calls, threads)
        setup = time.perf_counter() - t0
        r = ca.port_run(None, None, None, None, args.timesteps, threads, None, worlds=ws)
        port = {"value": r["world_steps"] / r["seconds"], "unit": UNIT, "cores": threads, "kind": "port",
                "sample": "all %d worlds of rank 0's batch x %d timesteps, oracle/fizz_oracle.c on %d pthreads, one run "
                          "of %.1f s (+ %.1f s building the worlds, untimed)" % (r["n"], args.timesteps, threads,
                                                                                 r["seconds"], setup),
                "capped_worlds": r["capped"]}
        del ws
    ps = pyref_sample(args, n_total)
    if ps is not None:
        pool, n = ps
        for _ in range(args.warmup):
            pool.run(args.pyref_timesteps)
        tot_steps, tot_s, cpu_s = 0, 0.0, 0.0
        for _ in range(args.steps):
            r = pool.run(args.pyref_timesteps)
            tot_steps += r["world_steps"]
            tot_s += r["seconds"]
            cpu_s += r["cpu_seconds"]
        pool.close()
        v = tot_steps / tot_s
        sample = ("per step: the first %d of %d worlds x the first %d of %d timesteps with the reference's own "
                  "World.update() (src/physics.py:76-259, unmodified, via oracle/ref_harness.py), one world per task over "
                  "%d processes; %.0f world-steps/s per core while busy"
                  % (n, args.worlds, args.pyref_timesteps, args.timesteps, r["procs"], tot_steps / cpu_s))
        line.update({"value": v, "ms_per_step": 1e3 * tot_s / args.steps,
                     "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["procs"], "kind": "reference", "sample": sample}})
        if port:
            line["port"] = port
    elif port:
        line.update({"value": port["value"], "ms_per_step": None, "cpu_baseline": port})
    else:
        # mixed / GA populations: the generators live in the product's host code (pure numpy, no native library)
        import fizz2d_b200.lib as fl
        assert fl._lib is None
        specs, _ = build_specs(args, 0, max(1, args.gpus))
        for _ in range(args.warmup):
            port_from_specs(specs, max(64, args.cpu_sample_worlds // 8), args.timesteps, threads)
        tot_steps, tot_s = 0, 0.0
        for _ in range(args.steps):
            r = port_from_specs(specs, args.cpu_sample_worlds, args.timesteps, threads)
            tot_steps += r["world_steps"]
            tot_s += r["seconds"]
        v = tot_steps / tot_s
        line.update({"value": v, "ms_per_step": 1e3 * tot_s / args.steps,
                     "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                      "sample": "first %d of %d worlds (proportionally per group) x %d timesteps, "
                                                "oracle/fizz_oracle.c on %d pthreads" % (r["n"], args.worlds,
                                                                                         args.timesteps, threads)}})
        assert fl._lib is None, "the reference arm must not load the product library"
    line["e2e"] = {"value": line["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------
def build_specs(args, rank, world_size, config=None, worlds=None):
    """GroupSpecs (+ global-order ids inside this rank's shard) of this rank's worlds of BASELINE config `config`."""
    import fizz2d_b200 as fz
    from fizz2d_b200.batch import GroupSpec
    from fizz2d_b200 import distributed as fd
    cfg = CONFIGS[config if config is not None else args.config]
    W = worlds if worlds is not None else (args.worlds if config is None else cfg["worlds"])
    total = W * world_size
    idx = fd.shard_indices(total, rank, world_size)          # interleaved: world w lives on rank w % G
    mc = args.max_calls
    if cfg["kind"] == "scene":
        sc = fz.load_in(fz.scenes.scene_path(cfg["scenes"][0]))
        deltas = fz.perturbation_deltas(cfg["seed"], total, sc.n_free)[idx]
        return [GroupSpec.from_scene(sc, W, np.ascontiguousarray(deltas), max_calls=mc)], None
    if cfg["kind"] == "mixed":
        scenes = [fz.load_in(fz.scenes.scene_path(s)) for s in cfg["scenes"]]
        fmax = max(sc.n_free for sc in scenes)
        deltas = fz.perturbation_deltas(cfg["seed"], total, fmax)
        specs, ids = [], []
        for k, sc in enumerate(scenes):
            loc = np.nonzero(idx % len(scenes) == k)[0]      # world w runs scene w % n_scenes (global index)
            if len(loc) == 0:
                continue
            d = np.ascontiguousarray(deltas[idx[loc]][:, :sc.n_free, :])
            specs.append(GroupSpec.from_scene(sc, len(loc), d, max_calls=mc))
            ids.append(loc)
        return specs, ids
    # GA population: one random convex polygon per world in the fixed geometry of the scene
    from fizz2d_b200 import ga, scene as _scene
    sc = fz.load_in(fz.scenes.scene_path(cfg["scenes"][0]))
    nverts, pts, vel = ga.random_convex_polygons(cfg["seed"], total)
    base = _scene.Scene(sc.width, sc.height, dict(sc.params), [], sc.fixed, sc.name)
    specs, ids = [], []
    nv_loc = nverts[idx]
    for n in sorted(set(int(x) for x in nv_loc)):
        loc = np.nonzero(nv_loc == n)[0]
        P = np.stack([pts[w] for w in idx[loc]])
        specs.append(GroupSpec(base, [np.ascontiguousarray(P)], vel[idx[loc]][:, None, :].copy(),
                               np.ones((len(loc), 1)), max_calls=mc))
        ids.append(loc)
    return specs, ids


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def source_id():
    """Hash of the CUDA sources + build flags: ties a profiles/*.json capture to the build it was taken from."""
    import glob
    h = hashlib.sha256()
    for p in sorted(glob.glob(os.path.join(ROOT, "fizz2d_b200", "csrc", "*"))) + [os.path.join(ROOT, "include", "fizz_b200.h"),
                                                                                os.path.join(ROOT, "fizz2d_b200", "build.py")]:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def measured_traffic():
    """DRAM bytes per launch of the kernels, from the newest ncu --set full capture of THIS build (profiles/r02_traffic.json,
    written by tools/ncu_traffic.py from the raw page; the command is recorded inside).  A capture of another build is
    not a measurement of the benchmarked binary: then `traffic` is null."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)
    except Exception:
        return {}, "no profiles/r02_traffic.json"
    if t.get("source_id") != source_id():
        return {}, "profiles/r02_traffic.json was captured from another build (source_id %s, this build %s)" % (
            t.get("source_id"), source_id())
    return t.get("kernels", {}), "profiles/r02_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, %s)" % t.get("command", "")


class Bench:
    """One resident batch on this rank's GPU and the measurement procedure of the contract."""

    def __init__(self, args, specs, ids, local_rank, world_size, rank, mode, chunk, timesteps, config_id):
        import torch
        self.config_id = config_id
        import fizz2d_b200 as fz
        self.torch, self.fz = torch, fz
        self.args, self.world_size, self.rank, self.dev_index = args, world_size, rank, local_rank
        self.dev = torch.device("cuda", local_rank)
        self.timesteps, self.mode, self.chunk = timesteps, mode, chunk
        # pinned host copies of the per-pass inputs (e2e uploads come from these)
        self.h2d_bytes = 0
        for spec in specs:
            for name in ("pos", "vel", "off", "thr", "mass"):
                t = torch.from_numpy(np.ascontiguousarray(getattr(spec, name))).pin_memory()
                setattr(spec, name, t.numpy())
                setattr(spec, "_pin_" + name, t)
                self.h2d_bytes += t.numel() * 8
        self.specs = specs
        self.bw = fz.BatchedWorld(specs, ids, device=local_rank)
        self.bw.set_tuning(mode, chunk)
        self.spec_tuning = dict(CONFIGS[config_id].get("spec") or {}) if mode == "hybrid" else {}
        if self.spec_tuning:
            self.bw.set_spec(**self.spec_tuning)
        self.W = self.bw.n_worlds
        self.d2h_bytes = 8 * self.W * (4 + 1 + 1)
        self.state0 = self.bw.snapshot()

    def close(self):
        self.bw.close()

    def gather_results(self):
        # the one collective of the path: end-of-run gather of [E,K,P,H,fitness,status] per world (NCCL)
        if self.world_size > 1:
            from fizz2d_b200 import distributed as fd
            torch, fz, bw = self.torch, self.fz, self.bw
            cur = torch.cuda.current_stream().cuda_stream
            parts = []
            for g in bw.groups:
                fz.lib.check(fz.lib.load().fizz_energy(g.handle, cur))
                parts.append(fd.pack_results(g.section("energy", 4), g.section("heat", 1)[0],      # fitness := heat
                                             g.section("status", 1, torch.int64)[0]))
            if len(parts) == 1:
                local = parts[0]
            else:
                local = torch.empty((fd.RESULT_ROWS, self.W), dtype=torch.float64, device=self.dev)
                for g, p in zip(bw.groups, parts):
                    local[:, torch.from_numpy(np.asarray(g.world_ids)).to(self.dev)] = p
            return fd.gather_results(local, self.W * self.world_size), local

    def pass_resident(self):
        self.bw.restore(self.state0)
        self.bw.step(self.timesteps)
        self.gather_results()

    def pass_e2e(self):
        self.bw.upload(sync=False)
        self.bw.step(self.timesteps)
        self.gather_results()
        return self.bw.download(energy=True, status=True, steps=True)

    def barrier(self):
        if self.world_size > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def check_gather(self):
        """The gather puts world w of the global batch at column w on every rank: this rank's columns are its shard."""
        from fizz2d_b200 import distributed as fd
        torch = self.torch
        full, mine = self.gather_results()
        idx = torch.from_numpy(fd.shard_indices(self.W * self.world_size, self.rank, self.world_size)).to(self.dev)
        assert full.shape == (fd.RESULT_ROWS, self.W * self.world_size)
        assert torch.equal(full[:, idx].view(torch.int64), mine.view(torch.int64)), "NCCL gather misplaced this rank's worlds"

    def time_gather(self, reps=5):
        if self.world_size == 1:
            return None
        torch = self.torch
        ms = []
        for _ in range(reps):
            self.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            self.gather_results()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    def measure(self, steps, warmup, flush, sample_clocks=True):
        """W warm-up passes, then K timed resident passes and K timed end-to-end passes.  Returns a dict of raw figures
        (this rank's); the caller reduces over ranks."""
        torch, bw = self.torch, self.bw
        for _ in range(warmup):
            flush.fill_(1)
            self.pass_resident()
        self.barrier()
        if self.world_size > 1:
            self.check_gather()
            self.barrier()
        sampler = ClockSampler(self.dev_index) if sample_clocks else None
        if sampler:
            sampler.start()
        launches0 = sum(i["launches"] for i in bw.kernel_info())
        bw.counters(reset=True)
        bw.spec_stats(reset=True)
        bw.profile(reset=True)
        bw.set_profiling(True)
        ev = []
        self.barrier()
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            self.pass_resident()
            b.record()
            ev.append((a, b))
        self.barrier()
        if sampler:
            sampler.stop_flag = True
        launches = sum(i["launches"] for i in bw.kernel_info()) - launches0
        profs = bw.profile(reset=True)
        cnts = bw.counters(reset=True)
        spec_stats = bw.spec_stats(reset=True)
        bw.set_profiling(False)
        t_ms = sum(a.elapsed_time(b) for a, b in ev)
        steps_done, status, calls = bw.steps_done(), bw.status(), bw.calls()
        # ---- end to end through the public API with host buffers ----------------------------------------------------
        self.pass_e2e()
        self.barrier()
        e2e_ms = []
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a.record()
            res = self.pass_e2e()
            b.record()
            torch.cuda.synchronize()
            e2e_ms.append(max(a.elapsed_time(b), 1e3 * (time.perf_counter() - t0)))
        e2e_steps = int(sum(r["steps"].sum() for r in res))
        self.barrier()
        return dict(t_ms=t_ms, e2e_ms=sum(e2e_ms), world_steps=int(steps_done.sum()), e2e_world_steps=e2e_steps,
                    launches=int(launches), profs=profs, cnts=cnts, spec_stats=spec_stats, status=status, calls=calls,
                    clocks=sampler.summary() if sampler else None)


def reduce_ranks(b, m):
    """max over ranks of the times, sum over ranks of the world-steps (device-timed, as the contract asks)."""
    torch = b.torch
    t = torch.tensor([m["t_ms"], m["e2e_ms"]], dtype=torch.float64, device=b.dev)
    ws = torch.tensor([m["world_steps"], m["e2e_world_steps"]], dtype=torch.float64, device=b.dev)
    if b.world_size > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ws, op=dist.ReduceOp.SUM)
    return t.tolist(), ws.tolist()


def kernel_tables(b, m, steps, ops_per_ws, peak_tflops):
    """Per-kernel time/occupancy table and FP64 rooflines (algorithmic ops / device time / measured DFMA peak)."""
    bw = b.bw
    infos = bw.kernel_info()
    prof = {}
    for p in m["profs"]:
        for k, (ms, n) in p.items():
            a = prof.setdefault(k, [0.0, 0])
            a[0] += ms
            a[1] += n
    cnt = {}
    for c in m["cnts"]:
        for k, v in c.items():
            cnt[k] = cnt.get(k, 0) + int(v)
    note = None
    hybrid = b.mode in ("hybrid", "hybrid_sync", "hybrid3")
    if b.mode == "hybrid" and prof.get("contact", [0, 0])[1] and len(bw.groups) == 1:
        # the library's spans are taken on the caller's stream; the runner launches of the contact kernel (trapped worlds,
        # library-owned streams) overlap them and end with the pass.  Contact kernels are busy whenever the ballistic
        # kernel is not: that is the time the roofline of the contact kernel is taken over.
        note = ("contact ms = pass - ballistic (the spec rounds and the runner launches of the contact kernel run on library-"
                "owned streams beside the chunk pipeline; the caller-stream spans alone were %.1f ms per pass)"
                % (prof["contact"][0] / steps))
        prof["contact"][0] = max(m["t_ms"] - prof["ballistic"][0], prof["contact"][0])
    info = infos[0]
    kernels = {k: {"ms_per_pass": prof[k][0] / steps, "launches_per_pass": prof[k][1] // steps,
                   "regs": info[k]["regs"], "smem_bytes": info[k]["smem"], "block": info[k]["threads"],
                   "blocks_per_sm": info[k]["blocks_per_sm"]} for k in prof if prof[k][1]}
    if len(bw.groups) > 1:
        for k in kernels:
            kernels[k]["note"] = "summed over %d groups that run concurrently on their own streams" % len(bw.groups)
    if note:
        kernels["contact"]["note"] = note
    if "contact" in kernels and b.mode == "hybrid":
        t = info["trapped"]
        kernels["contact"]["instances"] = {
            "contact_kernel<false> (throughput phase: chunks whose work list is >= 16384 worlds long)":
                {"regs": info["contact"]["regs"], "blocks_per_sm": info["contact"]["blocks_per_sm"]},
            "contact_kernel<true> (all later chunks, the spec rounds, runner launches)":
                {"regs": t["regs"], "blocks_per_sm": t["blocks_per_sm"]}}
    dom = max(prof, key=lambda k: prof[k][0]) if prof else None
    out = {"kernels": kernels, "dominant_kernel": dom, "work": {k: v // steps for k, v in cnt.items()}}
    sr = {}
    for c in m.get("spec_stats", []):
        for k, v in c.items():
            sr[k] = sr.get(k, 0) + int(v)
    if sr.get("rounds"):
        # time-parallel stepping of the expensive worlds: per pass
        out["work"]["spec_rounds"] = {k: v // steps for k, v in sr.items()}
    if ops_per_ws is None or not prof:
        return out
    # quiescent closed form per group: 12F + 2V + 8P ops per world-step, weighted by the ballistic steps
    q_ops, b_ops = [], 0.0
    for s, c in zip(b.specs, m["cnts"]):
        F_, X_, V_ = s.n_free, s.n_fixed, sum(s.nverts_free)
        q = 12 * F_ + 2 * V_ + 8 * (F_ * (F_ - 1) // 2 + F_ * X_)
        q_ops.append(q)
        b_ops += q * c["ballistic_world_steps"]
    total_ops = ops_per_ws * m["world_steps"] * steps
    traffic, traffic_src = measured_traffic()
    rk = {}
    if hybrid:
        for name, ops in (("ballistic", b_ops), ("contact", max(0.0, total_ops - b_ops))):
            sec = prof.get(name, [0, 0])[0] * 1e-3
            if sec > 0:
                rk[name] = {"bound": "fp64", "achieved": ops / sec / 1e12, "peak": peak_tflops, "unit": "TFLOP/s",
                            "frac": ops / sec / 1e12 / peak_tflops, "ms_per_pass": prof[name][0] / steps,
                            "traffic": traffic.get("%s@config%d" % (name, b.config_id))}
    else:
        sec = sum(prof[k][0] for k in prof) * 1e-3
        rk[dom] = {"bound": "fp64", "achieved": total_ops / sec / 1e12, "peak": peak_tflops, "unit": "TFLOP/s",
                   "frac": total_ops / sec / 1e12 / peak_tflops, "ms_per_pass": sec * 1e3 / steps, "traffic": None}
    out["roofline_by_kernel"] = rk
    whole = total_ops / (m["t_ms"] * 1e-3) / 1e12
    out["roofline_whole_pass"] = {"bound": "fp64", "achieved": whole, "peak": peak_tflops, "unit": "TFLOP/s",
                                  "frac": whole / peak_tflops}
    if dom in rk:
        out["roofline"] = dict(rk[dom], kernel=dom + "_kernel", ops_per_world_step=ops_per_ws,
                               quiescent_ops_per_world_step=q_ops[0] if len(q_ops) == 1 else q_ops,
                               traffic_source=traffic_src,
                               peak_source="fizz_fp64_peak(): DFMA microbenchmark on this GPU at bench start "
                                           "(MEASURED_PEAKS.json has no FP64 figure)")
    return out


def hbm_view(bytes_per_chunk, launches, total_ms):
    """The HBM side of the roofline, for completeness: the state slab is read and written once per chunk of timesteps."""
    peak, src = 6538.3, "fallback of B200_PROFILING.md"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    chunks = max(1, launches // 2)
    ach = bytes_per_chunk * chunks / (total_ms * 1e-3) / 1e9
    return {"bound": "hbm", "algorithmic_bytes_per_chunk": bytes_per_chunk, "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "peak_source": src,
            "note": "pos+vel+heat read and written once per chunk; HBM is four orders of magnitude from binding"}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build()
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    import fizz2d_b200 as fz
    dev = torch.device("cuda", local_rank)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    peak_tflops, _ = fz.fp64_peak(local_rank, 4096)
    threads = os.cpu_count() or 1

    specs, ids = build_specs(args, rank, world_size)
    b = Bench(args, specs, ids, local_rank, world_size, rank, args.mode, args.chunk, args.timesteps, args.config)
    m = b.measure(args.steps, args.warmup, flush)
    gather_ms = b.time_gather()
    (t_ms_max, e2e_ms_max), (ws_total, e2e_ws_total) = reduce_ranks(b, m)

    line = None
    if rank == 0:
        value = ws_total * args.steps / (t_ms_max * 1e-3)
        e2e_value = e2e_ws_total * args.steps / (e2e_ms_max * 1e-3)
        status, calls = m["status"], m["calls"]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak",
                "scaling_note": "weak scaling: every GPU steps its own %d worlds of an N x %d population.  Read as STRONG "
                                "scaling (the same %d worlds over N GPUs) config 3 gains little: the chains of its trapped "
                                "worlds, time-parallel as they are stepped, bound the pass (DESIGN.md sections 5d, 6)"
                                % (args.worlds, args.worlds, args.worlds),
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, world_size),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": b.h2d_bytes, "d2h_bytes_per_step": b.d2h_bytes,
                        "ms_per_step": e2e_ms_max / args.steps},
                "gpu_launches": m["launches"],
                "clocks": m["clocks"],
                "worlds": {"running": int((status == 0).sum()), "capped": int((status == 1).sum()),
                           "other": int((status > 1).sum()), "world_steps_per_pass": m["world_steps"],
                           "check_collisions_per_world_step": float(calls.sum()) / max(1, m["world_steps"])},
                "mode": args.mode, "chunk_steps": args.chunk, "spec_rounds_tuning": b.spec_tuning or "library defaults",
                "source_id": source_id()}
        if gather_ms is not None:
            line["gather"] = {"ms": gather_ms, "bytes_per_rank": 48 * b.W, "collective": "all_gather_into_tensor (NCCL)",
                              "note": "median of 5, timed alone after the passes; it is also inside every timed pass"}
        # ---- algorithmic work (op-counting C port on a sample of the same worlds) and CPU baselines -----------------
        ops_per_ws = None
        if not args.no_cpu_baseline:
            n = min(args.cpu_sample_worlds, b.W) if world_size == 1 else min(2048, b.W)
            r = port_from_specs(specs, n, args.timesteps, threads)
            ops_per_ws = r["ops"] / r["world_steps"]
            port = {"value": r["world_steps"] / r["seconds"], "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": "first %d of %d worlds x %d timesteps, C restatement of the reference "
                              "(oracle/fizz_oracle.c), %.1f s" % (r["n"], b.W, args.timesteps, r["seconds"])}
            line["worlds"]["sample_capped_fraction"] = r["capped"] / r["n"]
            if world_size == 1:
                ps = pyref_sample(args, b.W)
                if ps is not None:
                    pool, npy = ps
                    pool.run(min(20, args.pyref_timesteps))          # process start-up, imports
                    pr = pool.run(args.pyref_timesteps)
                    pool.close()
                    line["cpu_baseline"] = {
                        "value": pr["world_steps"] / pr["seconds"], "unit": UNIT, "cores": pr["procs"], "kind": "reference",
                        "sample": "first %d of %d worlds x first %d of %d timesteps, the reference's own World.update() "
                                  "(src/physics.py:76-259, unmodified, oracle/ref_harness.py), one world per task over %d "
                                  "processes, %.1f s; %.0f world-steps/s per busy core"
                                  % (npy, b.W, args.pyref_timesteps, args.timesteps, pr["procs"], pr["seconds"],
                                     pr["world_steps"] / pr["cpu_seconds"]),
                        "port": port}
                else:
                    line["cpu_baseline"] = port
        line.update(kernel_tables(b, m, args.steps, ops_per_ws, peak_tflops))
        if "roofline" in line:
            s0 = specs[0]
            line["roofline"]["note"] = ("the dominant kernel is latency-bound: every chunk of a pass ends when its slowest "
                                        "chain of resolve passes ends -- trapped worlds, stepped time-parallel by the spec "
                                        "rounds, and the worlds no predictor fits (DESIGN.md, 'critical path')")
            line["roofline"]["hbm"] = hbm_view(sum(8 * s.n_worlds * 2 * (4 * s.n_free + 1) for s in specs),
                                               sum(v["launches_per_pass"] for v in line["kernels"].values()) * args.steps,
                                               m["t_ms"])
    b.close()
    del b

    # ---- extra: the other BASELINE configs, briefly, on this GPU (N = 1 default run only) -------------------------
    extra = {}
    if world_size == 1 and args.config == 3 and not args.no_extra:
        cfgs = {}
        for c in (2, 4, 5):
            cfg = CONFIGS[c]
            sp, idc = build_specs(args, 0, 1, config=c)
            bb = Bench(args, sp, idc, local_rank, 1, 0, cfg["mode"], cfg["chunk"], cfg["timesteps"], c)
            mm = bb.measure(3, 3, flush, sample_clocks=False)
            ops = None
            if not args.no_cpu_baseline:
                r = port_from_specs(sp, 2048, cfg["timesteps"], threads)
                ops = r["ops"] / r["world_steps"]
            kt = kernel_tables(bb, mm, 3, ops, peak_tflops)
            st = mm["status"]
            cfgs["config%d" % c] = {
                "workload": "%s: %s x %d worlds x %d timesteps, mode %s, chunk %d" % (
                    cfg["label"], " + ".join(cfg["scenes"]), bb.W, cfg["timesteps"], cfg["mode"], cfg["chunk"]),
                "value": mm["world_steps"] * 3 / (mm["t_ms"] * 1e-3), "ms_per_pass": mm["t_ms"] / 3,
                "e2e": {"value": mm["e2e_world_steps"] * 3 / (mm["e2e_ms"] * 1e-3), "h2d_bytes_per_step": bb.h2d_bytes,
                        "d2h_bytes_per_step": bb.d2h_bytes},
                "unit": UNIT, "steps": 3, "warmup": 3, "groups": len(sp),
                "capped_worlds": int((st == 1).sum()), "ops_per_world_step": ops,
                "kernels": kt["kernels"], "roofline_by_kernel": kt.get("roofline_by_kernel"),
                "roofline_whole_pass": kt.get("roofline_whole_pass"), "work": kt["work"]}
            bb.close()
            del bb
        extra["configs"] = cfgs
    if world_size == 1 and args.sweep_worlds:
        sweep = []
        for W in [int(x) for x in args.sweep_worlds.split(",")]:
            sp, idc = build_specs(args, 0, 1, worlds=W)
            bb = Bench(args, sp, idc, local_rank, 1, 0, args.mode, args.chunk, args.timesteps, -1)
            mm = bb.measure(2, 1, flush, sample_clocks=False)
            ops = None
            if not args.no_cpu_baseline:
                r = port_from_specs(sp, 4096, args.timesteps, threads)
                ops = r["ops"] / r["world_steps"]
            kt = kernel_tables(bb, mm, 2, ops, peak_tflops)
            sweep.append({"worlds": W, "value": mm["world_steps"] * 2 / (mm["t_ms"] * 1e-3), "ms_per_pass": mm["t_ms"] / 2,
                          "roofline_whole_pass_frac": (kt.get("roofline_whole_pass") or {}).get("frac"),
                          "kernels_ms": {k: v["ms_per_pass"] for k, v in kt["kernels"].items()}})
            bb.close()
            del bb
        extra["worlds_sweep"] = sweep
    if rank == 0:
        if extra:
            line["extra"] = extra
        print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
