#!/usr/bin/env python
"""bench.py -- world-steps/sec of the batched Fizz-2D stepping engine on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE config 3 -- RAMPWORLD_2TRIANGLES_5SQUARES.in x 65536 perturbed worlds
(seed PCG64(20260003), SURVEY 8d law) x 10000 timesteps per GPU (weak scaling: every rank steps its own 65536
worlds; no inter-GPU traffic while stepping, one NCCL all_gather of per-world [E,K,P,H,fitness,status] per pass).
One bench "step" = one pass of the hot path: reset to the initial conditions, advance every world by all
timesteps.  `value` = world-steps completed / device time with inputs resident in HBM; `e2e` = the same through
the public API with HOST (pinned) buffers: H2D of initial conditions and constants, stepping, D2H of the
per-world energy/status/step counters, all inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--worlds W] [--timesteps T]

`--impl reference` times the CPU restatement of the reference (oracle/fizz_oracle.c, kind "port": the reference
itself is Python and cannot travel to the GPU box) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENE = "RAMPWORLD_2TRIANGLES_5SQUARES"
SEED = 20260003
METRIC = "world_steps_per_sec"
UNIT = "world-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--worlds", type=int, default=65536, help="worlds per GPU")
    ap.add_argument("--timesteps", type=int, default=10000)
    ap.add_argument("--chunk", type=int, default=96, help="timesteps per chunk (hybrid mode)")
    ap.add_argument("--mode", default="hybrid", choices=["hybrid", "hybrid_sync", "thread"])
    ap.add_argument("--max-calls", type=int, default=4096)
    ap.add_argument("--cpu-sample-worlds", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {"workload": "BASELINE config 3: %s.in x %d perturbed worlds/GPU x %d timesteps (seed PCG64(%d), "
                        "dpos 5 px, dvel 30 px/s per body; watchdog max_calls=%d)"
                        % (SCENE, args.worlds, args.timesteps, SEED, args.max_calls),
            "scene": SCENE + ".in", "worlds_per_gpu": args.worlds, "timesteps": args.timesteps,
            "free_bodies": 7, "fixed_bodies": 5, "max_calls": args.max_calls,
            "sharding": "worlds sharded over %d GPU(s), no collective while stepping" % n_gpus,
            "l2": "L2 flushed (512 MiB write) before every pass"}


# ----------------------------------------------------------------------------------------------------------
# CPU arm (oracle port) -- the only place bench.py executes oracle/
# ----------------------------------------------------------------------------------------------------------
def cpu_run(spec, worlds, timesteps, threads, max_calls):
    """Times the CPU oracle on `worlds` (indices into spec) x timesteps. Returns (world_steps, seconds, ops)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    prm = orc.make_params(spec.scene.params, spec.height, max_calls)
    ws = [orc.OracleWorld(spec.nverts_free, spec.nverts_fixed,
                          np.concatenate([spec.pos[:, w].reshape(-1, 2), spec.fixed_com]),
                          spec.vel[:, w].reshape(-1, 2), np.concatenate([spec.radius[:, w], spec.fixed_radius]),
                          spec.mass[:, w], spec.off[0::2, w], spec.off[1::2, w], spec.fixed_xy, prm)
          for w in worlds]
    t0 = time.perf_counter()
    orc.run_batch(ws, timesteps, threads)
    dt = time.perf_counter() - t0
    steps = sum(w.steps_done() for w in ws)
    ops = sum(w.total_ops() for w in ws)
    calls = sum(w.total_calls() for w in ws)
    capped = sum(w.status() != 0 for w in ws)
    return dict(world_steps=steps, seconds=dt, ops=ops, calls=calls, capped=capped, n=len(ws))


def build_spec(args, rank, world_size):
    import fizz2d_b200 as fz
    from fizz2d_b200.batch import GroupSpec
    sc = fz.load_in(fz.scenes.scene_path(SCENE))
    from fizz2d_b200 import distributed as fd
    total = args.worlds * world_size
    idx = fd.shard_indices(total, rank, world_size)          # interleaved: world w lives on rank w % G
    deltas = fz.perturbation_deltas(SEED, total, sc.n_free)[idx]
    return GroupSpec.from_scene(sc, args.worlds, np.ascontiguousarray(deltas), max_calls=args.max_calls)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as ge
    ge.build()
    spec = build_spec(args, 0, 1)
    threads = os.cpu_count() or 1
    n = min(args.cpu_sample_worlds, args.worlds)
    sel = list(range(n))
    for _ in range(args.warmup):
        cpu_run(spec, sel[:max(64, n // 8)], args.timesteps, threads, args.max_calls)
    tot_steps, tot_s = 0, 0.0
    for _ in range(args.steps):
        r = cpu_run(spec, sel, args.timesteps, threads, args.max_calls)
        tot_steps += r["world_steps"]
        tot_s += r["seconds"]
    v = tot_steps / tot_s
    sample = "first %d of %d worlds x %d timesteps per step (pthreads over %d host threads)" % (
        n, args.worlds, args.timesteps, threads)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


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

    spec = build_spec(args, rank, world_size)
    # pinned host copies of the per-pass inputs (e2e uploads come from these)
    for name in ("pos", "vel", "off", "thr", "mass"):
        t = torch.from_numpy(np.ascontiguousarray(getattr(spec, name))).pin_memory()
        setattr(spec, name, t.numpy())
        setattr(spec, "_pin_" + name, t)
    bw = fz.BatchedWorld([spec], device=local_rank)
    bw.set_tuning(args.mode, args.chunk)
    g = bw.groups[0]
    L = g.layout
    W = args.worlds
    state0 = bw.snapshot()                           # device-resident initial state for the kernel-only arm
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    peak_tflops, _ = fz.fp64_peak(local_rank, 4096)
    h2d_bytes = 8 * W * (spec.pos.shape[0] + spec.vel.shape[0] + spec.off.shape[0] + spec.thr.shape[0] + spec.mass.shape[0])
    d2h_bytes = 8 * W * (4 + 1 + 1)
    from fizz2d_b200 import distributed as fd

    def stepping():
        bw.step(args.timesteps)

    def gather_results():
        # the one collective of the path: end-of-run gather of [E,K,P,H,fitness,status] per world (NCCL)
        if world_size > 1:
            fz.lib.check(fz.lib.load().fizz_energy(g.handle, torch.cuda.current_stream().cuda_stream))
            local = fd.pack_results(g.section("energy", 4), g.section("heat", 1)[0],      # fitness := heat
                                    g.section("status", 1, torch.int64)[0])
            return fd.gather_results(local, W * world_size)

    def pass_resident():
        bw.restore(state0)
        stepping()
        gather_results()

    def pass_e2e():
        g.upload(torch.cuda.current_stream().cuda_stream)
        stepping()
        gather_results()
        r = bw.download(energy=True, status=True, steps=True)
        return r[0]

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.fill_(1)
        pass_resident()
    barrier()
    if world_size > 1:
        # the gather puts world w of the global batch at column w on every rank: this rank's columns are its shard
        full = gather_results()
        fz.lib.check(fz.lib.load().fizz_energy(g.handle, torch.cuda.current_stream().cuda_stream))
        mine = fd.pack_results(g.section("energy", 4), g.section("heat", 1)[0], g.section("status", 1, torch.int64)[0])
        idx = torch.from_numpy(fd.shard_indices(W * world_size, rank, world_size)).to(dev)
        assert full.shape == (fd.RESULT_ROWS, W * world_size)
        assert torch.equal(full[:, idx].view(torch.int64), mine.view(torch.int64)), "NCCL gather misplaced this rank's worlds"
        barrier()

    # ---- timed: resident --------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = bw.kernel_info()[0]["launches"]
    bw.counters(reset=True)
    bw.profile(reset=True)
    bw.set_profiling(True)
    pass_ms = []
    barrier()
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pass_resident()
        b.record()
        pass_ms.append((a, b))
    barrier()
    sampler.stop_flag = True
    launches = bw.kernel_info()[0]["launches"] - launches0
    prof = bw.profile(reset=True)[0]
    cnt = bw.counters(reset=True)[0]
    bw.set_profiling(False)
    t_ms = sum(a.elapsed_time(b) for a, b in pass_ms)
    runner_note = None
    if args.mode == "hybrid" and prof.get("contact", (0, 0))[1]:
        # the library's spans are taken on the caller's stream; the runner launches of the contact kernel (trapped worlds,
        # library-owned streams) overlap them and end with the pass.  Contact kernels are busy whenever the ballistic
        # kernel is not: that is the time the roofline of the contact kernel is taken over.
        runner_note = ("contact ms = pass - ballistic (runner launches of contact_kernel<true> overlap the chunk "
                       "pipeline; the caller-stream spans alone were %.1f ms per pass)" % (prof["contact"][0] / args.steps))
        prof["contact"] = (max(t_ms - prof["ballistic"][0], prof["contact"][0]), prof["contact"][1])
    steps_done = bw.steps_done()
    status = bw.status()
    calls = bw.calls()
    world_steps = int(steps_done.sum())

    # ---- timed: end to end through the public API with host buffers ----------------------------------------------
    pass_e2e()
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.record()
        res = pass_e2e()
        b.record()
        torch.cuda.synchronize()
        e2e_ms.append(max(a.elapsed_time(b), 1e3 * (time.perf_counter() - t0)))
    e2e_steps = int(res["steps"].sum())
    barrier()

    t = torch.tensor([t_ms, sum(e2e_ms)], dtype=torch.float64, device=dev)
    ws = torch.tensor([world_steps, e2e_steps], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ws, op=dist.ReduceOp.SUM)
    t_ms_max, e2e_ms_max = t.tolist()
    ws_total, e2e_ws_total = ws.tolist()

    if rank == 0:
        value = ws_total * args.steps / (t_ms_max * 1e-3)
        e2e_value = e2e_ws_total * args.steps / (e2e_ms_max * 1e-3)
        info = bw.kernel_info()[0]
        dom = max(prof, key=lambda k: prof[k][0])     # dominant kernel by device time
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, world_size),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": e2e_ms_max / args.steps},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "worlds": {"running": int((status == 0).sum()), "capped": int((status == 1).sum()),
                           "other": int((status > 1).sum()), "world_steps_per_pass": world_steps,
                           "check_collisions_per_world_step": float(calls.sum()) / max(1, world_steps)},
                "mode": args.mode, "chunk_steps": args.chunk,
                "kernels": {k: {"ms_per_pass": prof[k][0] / args.steps, "launches_per_pass": prof[k][1] // args.steps,
                                "regs": info[k]["regs"], "smem_bytes": info[k]["smem"], "block": info[k]["threads"],
                                "blocks_per_sm": info[k]["blocks_per_sm"]} for k in prof if prof[k][1]},
                "dominant_kernel": dom}
        if runner_note:
            line["kernels"]["contact"]["note"] = runner_note
        if "contact" in line["kernels"] and args.mode == "hybrid":
            # mode hybrid: contact_kernel<false> (the figures above) only runs the chunk that starts at step 0; every
            # later chunk and every runner launch -- the latency-bound tail -- runs the second instance of the template
            t = info["trapped"]
            line["kernels"]["contact"]["instances"] = {
                "contact_kernel<false> (first chunk)": {"regs": info["contact"]["regs"], "blocks_per_sm": info["contact"]["blocks_per_sm"]},
                "contact_kernel<true> (all later chunks, runner launches)": {"regs": t["regs"], "blocks_per_sm": t["blocks_per_sm"]}}
        # ---- CPU baseline + algorithmic work (bounded sample of the same workload) ----------------------------
        ops_per_ws = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n = min(args.cpu_sample_worlds, W)
            r = cpu_run(spec, list(range(n)), args.timesteps, threads, args.max_calls)
            ops_per_ws = r["ops"] / r["world_steps"]
            line["cpu_baseline"] = {"value": r["world_steps"] / r["seconds"], "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "first %d of %d worlds x %d timesteps, C restatement of the reference "
                                              "(oracle/fizz_oracle.c), %.1f s" % (n, W, args.timesteps, r["seconds"]),
                                    "python_reference_note": "the Python reference itself ran at 384 world-steps/s/core "
                                                             "on this scene in the build container (BASELINE.md)"}
            line["worlds"]["sample_capped_fraction"] = r["capped"] / r["n"]
        # ---- rooflines: the FP64 pipe binds (state stays on chip: HBM sees 2*S bytes per world per launch) -------
        # ALGORITHMIC ops follow SURVEY 8d: a quiescent world-step is 12F + 2V + 8P ops (closed form), the mean over
        # all world-steps is measured by the op-counting CPU restatement on the sample above.
        F_, X_, V_ = spec.n_free, spec.n_fixed, sum(spec.nverts_free)
        P_ = F_ * (F_ - 1) // 2 + F_ * X_
        q_ops = 12 * F_ + 2 * V_ + 8 * P_
        line["work"] = {k: int(v) // args.steps for k, v in cnt.items()}
        if ops_per_ws is not None and args.mode in ("hybrid", "hybrid_sync"):
            total_ops = ops_per_ws * world_steps * args.steps
            b_ops = q_ops * cnt["ballistic_world_steps"]
            rk = {}
            for name, ops in (("ballistic", b_ops), ("contact", max(0.0, total_ops - b_ops))):
                sec = prof[name][0] * 1e-3
                if sec > 0:
                    rk[name] = {"bound": "fp64", "achieved": ops / sec / 1e12, "peak": peak_tflops, "unit": "TFLOP/s",
                                "frac": ops / sec / 1e12 / peak_tflops, "ms_per_pass": prof[name][0] / args.steps}
            if "ballistic" in rk and W == 65536:
                # per launch; from profiles/r01_ballistic_raw.csv (ncu --set full, dram__bytes_read.sum + write.sum)
                rk["ballistic"]["traffic"] = 19.94e6
                rk["ballistic"]["algorithmic_bytes_per_launch"] = 8 * W * (5 * F_)   # pos, vel, thr rows read once
            line["roofline_by_kernel"] = rk
            # DRAM bytes of one launch of the dominant kernel from profiles/r01_final_kernels_raw.csv (ncu --set full,
            # steady-state contact launch): 49.5 MB read + 0.3 MB written -- the state slab once; irrelevant next to HBM
            dom_traffic = 49.8e6 if (dom == "contact" and W == 65536) else None
            line["roofline"] = dict(rk[dom], traffic=dom_traffic, kernel=dom + "_kernel", ops_per_world_step=ops_per_ws,
                                    quiescent_ops_per_world_step=q_ops,
                                    peak_source="fizz_fp64_peak(): DFMA microbenchmark on this GPU at bench start "
                                                "(MEASURED_PEAKS.json has no FP64 figure)",
                                    note="the dominant kernel is latency-bound: a pass ends when the world with the "
                                         "longest sequential chain of resolve passes ends (DESIGN.md, 'critical path')",
                                    hbm=hbm_view(8 * W * 2 * (4 * F_ + 1), prof["contact"][1] + prof["ballistic"][1],
                                                 t_ms))
        elif ops_per_ws is not None:
            sec = sum(prof[k][0] for k in prof) * 1e-3
            ach = ops_per_ws * world_steps * args.steps / sec / 1e12
            line["roofline"] = {"bound": "fp64", "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s",
                                "frac": ach / peak_tflops, "traffic": None, "kernel": "step_kernel",
                                "ops_per_world_step": ops_per_ws}
        print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


def hbm_view(bytes_per_chunk, launches, total_ms):
    """The HBM side of the roofline, for completeness: the state slab is read and written once per chunk of timesteps."""
    peak, src = 6538.3, "fallback of B200_PROFILING.md"
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    chunks = max(1, launches // 2)
    ach = bytes_per_chunk * chunks / (total_ms * 1e-3) / 1e9
    return {"bound": "hbm", "algorithmic_bytes_per_chunk": bytes_per_chunk, "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "peak_source": src,
            "note": "pos+vel+heat read and written once per chunk; HBM is four orders of magnitude from binding"}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
