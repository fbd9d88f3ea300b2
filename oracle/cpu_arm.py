"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- the CPU arms of bench.py (`--impl reference` and the `cpu_baseline` leg).

Nothing here imports the product package or loads its .so: scene texts come from the committed fixtures
(tests/golden/<SCENE>.npz `scene_text`), the perturbation law is restated below, and worlds are built with oracle.py's
own parser and init-time math.

Two implementations of the reference path can be timed on the host cores:

* kind "reference": the reference's OWN Python (`World.update()`, src/physics.py:76-259), unmodified, imported through
  `ref_harness` from /root/reference/src where that exists (the build container) or from `oracle/_ref/pyref/` (a verbatim
  copy written there by `make -C oracle ref`; `oracle/_ref/` is git-ignored and travels to the GPU box with the
  snapshot).  One world per task, one process per host core, `OMP_NUM_THREADS=1`.
* kind "port": `oracle/fizz_oracle.c`, the scalar C restatement pinned against the reference, pthreads over all cores.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def scene_text(name: str) -> str:
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as g:
        return str(g["scene_text"])


def perturbation_deltas(seed, n_worlds, n_free, dpos=5.0, dvel=30.0):
    """SURVEY 8d perturbation law, restated: world 0 unperturbed; worlds >= 1 consume, in world then body order, four
    uniforms from Generator(PCG64(seed)): dx, dy ~ U(-dpos, dpos), dvx, dvy ~ U(-dvel, dvel)."""
    d = np.zeros((n_worlds, n_free, 4))
    if n_worlds > 1:
        u = np.random.Generator(np.random.PCG64(seed)).random((n_worlds - 1, n_free, 4))
        lo = np.array([-dpos, -dpos, -dvel, -dvel])
        d[1:] = lo + (-lo - lo) * u
    return d


# ---------------------------------------------------------------------------------------------------------------
# kind "port"
# ---------------------------------------------------------------------------------------------------------------
def _port_constants(args):
    """Init-time constants of a slice of worlds (scalar, oracle.body_constants): runs in a pool worker."""
    text, deltas = args
    import oracle as orc
    sc = orc.parse_scene(text)
    eps = sc["params"].get("EPSILON", orc.DEFAULT_PARAMS["EPSILON"])
    out = []
    for d in deltas:
        com, vel, radius, mass, offx, offy = [], [], [], [], [], []
        for b, body in enumerate(sc["free"]):
            pts = [[p[0] + d[b][0], p[1] + d[b][1]] for p in body["points"]]
            c = orc.body_constants(pts, body["density"], eps)
            com.append(c["com"]); radius.append(c["radius"]); mass.append(c["mass"])
            vel.append([body["vel"][0] + d[b][2], body["vel"][1] + d[b][3]])
            offx += c["offx"]; offy += c["offy"]
        out.append((np.array(com), np.array(vel), np.array(radius), np.array(mass), np.array(offx), np.array(offy)))
    return out


def port_worlds(scene: str, seed: int, n_total: int, world_ids, max_calls: int, procs: int):
    """OracleWorld objects for `world_ids` of the n_total-world population of `scene` (own parser + init math)."""
    import multiprocessing as mp
    import oracle as orc
    text = scene_text(scene)
    sc = orc.parse_scene(text)
    deltas = perturbation_deltas(seed, n_total, len(sc["free"]))[np.asarray(world_ids)]
    fx = [orc.body_constants(pts, None) for pts in sc["fixed"]]
    fcom = np.array([c["com"] for c in fx]).reshape(-1, 2)
    frad = np.array([c["radius"] for c in fx])
    fixedxy = np.array([p for pts in sc["fixed"] for p in pts], dtype=float).reshape(-1, 2)
    nvf = [len(b["points"]) for b in sc["free"]]
    nvx = [len(p) for p in sc["fixed"]]
    prm = orc.make_params(sc["params"], sc["height"], max_calls)
    procs = max(1, min(procs, len(deltas) // 64 or 1))
    chunks = [(text, deltas[i::procs].tolist()) for i in range(procs)]
    if procs > 1:
        with mp.get_context("fork").Pool(procs) as pool:
            parts = pool.map(_port_constants, chunks)
    else:
        parts = [_port_constants(chunks[0])]
    consts = [None] * len(deltas)
    for i, part in enumerate(parts):
        consts[i::procs] = part
    worlds = [orc.OracleWorld(nvf, nvx, np.concatenate([c[0], fcom]), c[1], np.concatenate([c[2], frad]), c[3], c[4], c[5],
                              fixedxy, prm, width=sc["width"], height=sc["height"]) for c in consts]
    return worlds


def port_run(scene, seed, n_total, world_ids, timesteps, threads, max_calls, worlds=None):
    """Times the C port on the given worlds.  Returns dict(world_steps, seconds, ops, calls, capped, n)."""
    import oracle as orc
    ws = worlds if worlds is not None else port_worlds(scene, seed, n_total, world_ids, max_calls, threads)
    t0 = time.perf_counter()
    orc.run_batch(ws, timesteps, threads)
    dt = time.perf_counter() - t0
    return dict(world_steps=sum(w.steps_done() for w in ws), seconds=dt, ops=sum(w.total_ops() for w in ws),
                calls=sum(w.total_calls() for w in ws), capped=sum(w.status() != 0 for w in ws), n=len(ws))


# ---------------------------------------------------------------------------------------------------------------
# kind "reference": the unmodified Python reference
# ---------------------------------------------------------------------------------------------------------------
def pyref_available() -> bool:
    import ref_harness as rh
    return rh.available()


def _pyref_world(args):
    text, timesteps, max_calls = args
    os.environ["OMP_NUM_THREADS"] = "1"
    import ref_harness as rh
    w = rh.RefWorld(text=text, max_calls=max_calls)
    t0 = time.perf_counter()
    n = 0
    for _ in range(timesteps):
        if not w.update():
            break
        n += 1
    return n, time.perf_counter() - t0, sum(w.calls_per_step)


class PyRefPool:
    """Persistent pool of one process per host core; each task steps ONE world of the population with the reference's
    own World.update() (timed region = the update() loop of every world; wall time of the whole sample is reported)."""

    def __init__(self, scene, seed, n_total, world_ids, max_calls, procs):
        import multiprocessing as mp
        import ref_harness as rh
        os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
        text = scene_text(scene)
        import oracle as orc
        n_free = len(orc.parse_scene(text)["free"])
        deltas = perturbation_deltas(seed, n_total, n_free)[np.asarray(world_ids)]
        import tempfile
        with tempfile.NamedTemporaryFile("w", suffix=".in", delete=False) as f:
            f.write(text)
            path = f.name
        try:
            self.texts = [rh.perturbed_scene_text(path, d) for d in deltas]
        finally:
            os.unlink(path)
        self.max_calls = max_calls
        self.procs = max(1, min(procs, len(self.texts)))
        self.pool = mp.get_context("fork").Pool(self.procs)

    def run(self, timesteps):
        t0 = time.perf_counter()
        res = self.pool.map(_pyref_world, [(t, timesteps, self.max_calls) for t in self.texts], chunksize=1)
        dt = time.perf_counter() - t0
        return dict(world_steps=sum(r[0] for r in res), seconds=dt, cpu_seconds=sum(r[1] for r in res),
                    calls=sum(r[2] for r in res), n=len(res), procs=self.procs,
                    capped=sum(r[0] < timesteps for r in res))

    def close(self):
        self.pool.close()
        self.pool.join()
