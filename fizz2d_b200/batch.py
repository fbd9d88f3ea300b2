"""Batched entry point: thousands to millions of independent Fizz-2D worlds advanced in lockstep on one B200.

`BatchedWorld` is the batched counterpart of the reference's `physics.World` (physics.py:30-301):
`.step(n)` = n x `World.update()` for every world, `.energy()` = `World.energy()`, `.dump_world(i)` =
`str(world)`.  World state lives in one device slab per *group* (worlds sharing a scene topology), owned by a
torch tensor; all compute goes through the C-ABI of libfizz_b200.so (include/fizz_b200.h).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import lib as _lib
from . import scene as _scene


class GroupSpec:
    """Host-side description of one group: topology + fixed geometry + per-world SoA initial conditions."""

    def __init__(self, scene: _scene.Scene, points: Sequence[np.ndarray], velocity: np.ndarray,
                 density: np.ndarray, gravity=(0.0, 0.0), max_calls=4096, check_valid=True, friction=None,
                 ff_restitution=None):
        """points[b]: (W, n_b, 2) vertices of free body b per world; velocity: (W, F, 2); density: (W, F).
        friction / ff_restitution: the default-off physics extensions (see scene.resolve_params)."""
        self.scene = scene
        self.width, self.height = scene.width, scene.height
        F = len(points)
        W = points[0].shape[0]
        self.n_worlds, self.n_free, self.n_fixed = W, F, scene.n_fixed
        self.nverts_free = [p.shape[1] for p in points]
        self.nverts_fixed = [len(p) for p in scene.fixed]
        eps = scene.params.get("EPSILON", _scene.DEFAULT_PARAMS["EPSILON"])
        V = sum(self.nverts_free)
        self.pos = np.empty((2 * F, W))
        self.vel = np.empty((2 * F, W))
        self.off = np.empty((2 * V, W))
        self.radius = np.empty((F, W))
        self.mass = np.empty((F, W))
        valid = np.ones(W, dtype=bool)
        v0 = 0
        for b, P in enumerate(points):
            c = _scene.polygon_constants(P, density[:, b], eps)
            self.pos[2 * b], self.pos[2 * b + 1] = c["com"][:, 0], c["com"][:, 1]
            self.vel[2 * b], self.vel[2 * b + 1] = velocity[:, b, 0], velocity[:, b, 1]
            n = P.shape[1]
            for v in range(n):
                self.off[2 * (v0 + v)] = c["offx"][:, v]
                self.off[2 * (v0 + v) + 1] = c["offy"][:, v]
            v0 += n
            self.radius[b] = c["radius"]
            self.mass[b] = c["mass"]
            valid &= c["valid"]
        self.valid = valid
        if check_valid and not valid.all():
            raise ValueError("%d world(s) hold a degenerate polygon (NaN/zero inertia or a vertex on the centroid): "
                             "the reference turns those into NaN on step 1 (SURVEY Q13)" % int((~valid).sum()))
        self.thr = _scene.sqrt_threshold(self.radius)
        # fixed bodies (FixedPolygon -> Polygon.__init__ -> Obj.__init__: centroid and radius the same way)
        self.fixed_xy = np.concatenate(scene.fixed, axis=0) if scene.fixed else np.zeros((0, 2))
        self.fixed_com = np.zeros((self.n_fixed, 2))
        self.fixed_radius = np.zeros(self.n_fixed)
        for j, pts in enumerate(scene.fixed):
            c = _scene.polygon_constants(pts[None], None)
            self.fixed_com[j] = c["com"][0]
            self.fixed_radius[j] = c["radius"][0]
        self.fixed_thr = _scene.sqrt_threshold(self.fixed_radius)
        self.params = _scene.resolve_params(scene.params, gravity, scene.height, max_calls, friction, ff_restitution)

    @classmethod
    def from_scene(cls, scene: _scene.Scene, n_worlds=1, deltas: Optional[np.ndarray] = None, **kw):
        """Replicate a scene n_worlds times; deltas (W,F,4) = per-body (dx,dy,dvx,dvy) (scene.perturbation_deltas)."""
        F = scene.n_free
        if deltas is None:
            deltas = np.zeros((n_worlds, F, 4))
        assert deltas.shape == (n_worlds, F, 4)
        points, vel, dens = [], np.empty((n_worlds, F, 2)), np.empty((n_worlds, F))
        for b, body in enumerate(scene.free):
            P = body.points[None, :, :] + deltas[:, b, None, 0:2]
            points.append(np.ascontiguousarray(P))
            vel[:, b, :] = body.velocity[None, :] + deltas[:, b, 2:4]
            dens[:, b] = body.density
        return cls(scene, points, vel, dens, **kw)


class _Group:
    def __init__(self, spec: GroupSpec, device: int, world_ids: np.ndarray):
        import torch
        self.spec = spec
        self.world_ids = world_ids
        L = _lib.load()
        nverts = np.ascontiguousarray(spec.nverts_free + spec.nverts_fixed, dtype=np.int32)
        self._keep = (nverts, np.ascontiguousarray(spec.fixed_xy), np.ascontiguousarray(spec.fixed_com),
                      np.ascontiguousarray(spec.fixed_thr))
        self.desc = _lib.FizzGroupDesc(
            n_free=spec.n_free, n_fixed=spec.n_fixed, nverts=nverts.ctypes.data_as(_lib.ip),
            fixed_xy=_lib.as_dp(self._keep[1]), fixed_com=_lib.as_dp(self._keep[2]),
            fixed_thr=_lib.as_dp(self._keep[3]), params=spec.params)
        self.layout = _lib.FizzLayout()
        _lib.check(L.fizz_layout(C.byref(self.desc), spec.n_worlds, C.byref(self.layout)))
        if not torch.cuda.is_available():
            raise _lib.FizzError("no CUDA device: the batched engine has no CPU fallback")
        self.device = device
        self.slab = torch.zeros(self.layout.total_words, dtype=torch.float64, device="cuda:%d" % device)
        h = C.c_void_p()
        _lib.check(L.fizz_group_create(C.byref(self.desc), spec.n_worlds, device, C.c_void_p(self.slab.data_ptr()),
                                       self.layout.total_words, C.byref(h)))
        self.handle = h
        self.trace = None

    def close(self):
        if self.handle:
            _lib.load().fizz_group_destroy(self.handle)
            self.handle = None

    def upload(self, stream=0):
        s = self.spec
        arrs = [np.ascontiguousarray(a) for a in (s.pos, s.vel, s.off, s.thr, s.mass)]
        _lib.check(_lib.load().fizz_upload(self.handle, *[_lib.as_dp(a) for a in arrs], C.c_void_p(stream)))
        self._held = arrs

    def section(self, name, rows, dtype=None):
        """Zero-copy torch view [rows, W] of a slab section."""
        import torch
        L = self.layout
        o = getattr(L, name)
        t = self.slab[o:o + rows * L.stride].view(rows, L.stride)[:, :L.n_worlds]
        if dtype is not None:
            t = self.slab[o:o + rows * L.stride].view(dtype).view(rows, L.stride)[:, :L.n_worlds]
        return t

    def enable_trace(self, capacity):
        import torch
        self.trace = (torch.zeros(capacity * 6, dtype=torch.int64, device=self.slab.device),
                      torch.zeros(1, dtype=torch.int64, device=self.slab.device), capacity)
        _lib.check(_lib.load().fizz_set_trace(self.handle, C.c_void_p(self.trace[0].data_ptr()), capacity,
                                              C.c_void_p(self.trace[1].data_ptr())))

    def read_trace(self):
        import torch
        torch.cuda.synchronize(self.slab.device)
        n = int(self.trace[1].item())
        if n > self.trace[2]:
            raise _lib.FizzError("contact trace overflow: %d > %d" % (n, self.trace[2]))
        raw = self.trace[0][:n * 6].cpu().numpy().tobytes()
        recs = (_lib.FizzContact * n).from_buffer_copy(raw)
        out = [(int(self.world_ids[r.world]), r.step, r.call, r.i, r.j, r.nx, r.ny, r.mag) for r in recs]
        out.sort(key=lambda r: r[:5])
        return out


class BatchedWorld:
    """W independent worlds.  World index order is the order given at construction; groups are an internal
    partition by scene topology (mixed batches keep one resident scene descriptor per group)."""

    def __init__(self, specs: Sequence[GroupSpec], world_ids: Optional[Sequence[np.ndarray]] = None, device: int = 0):
        import torch
        self.device = device
        torch.cuda.set_device(device)
        if world_ids is None:
            world_ids, o = [], 0
            for s in specs:
                world_ids.append(np.arange(o, o + s.n_worlds))
                o += s.n_worlds
        self.groups: List[_Group] = [_Group(s, device, np.asarray(ids)) for s, ids in zip(specs, world_ids)]
        self.n_worlds = sum(s.n_worlds for s in specs)
        self._where = None
        self.streams = [torch.cuda.Stream(device=device) for _ in self.groups] if len(self.groups) > 1 else [None]
        self.steps_requested = 0
        self.upload()

    # -- construction ------------------------------------------------------------------------------------
    @classmethod
    def from_in(cls, path, n_worlds=1, perturb=False, seed=0, dpos=5.0, dvel=30.0, gravity=(0.0, 0.0),
                max_calls=4096, device=0, friction=None, ff_restitution=None):
        sc = _scene.load_in(path)
        deltas = _scene.perturbation_deltas(seed, n_worlds, sc.n_free, dpos, dvel) if perturb else None
        return cls([GroupSpec.from_scene(sc, n_worlds, deltas, gravity=gravity, max_calls=max_calls, friction=friction,
                                         ff_restitution=ff_restitution)], device=device)

    @classmethod
    def from_text(cls, text, n_worlds=1, deltas=None, gravity=(0.0, 0.0), max_calls=4096, device=0, friction=None,
                  ff_restitution=None):
        sc = _scene.parse_in(text)
        return cls([GroupSpec.from_scene(sc, n_worlds, deltas, gravity=gravity, max_calls=max_calls, friction=friction,
                                         ff_restitution=ff_restitution)], device=device)

    @classmethod
    def from_mixed(cls, paths, n_worlds, seed=0, perturb=True, dpos=5.0, dvel=30.0, max_calls=4096, device=0):
        """Mixed batch: world w runs scene paths[w % len(paths)].  One perturbation row of
        max(F) bodies is drawn per world; a scene with fewer bodies uses the first rows.  One group (one resident
        scene descriptor, own stream) per scene; world order is preserved by `world_ids`."""
        scenes = [_scene.load_in(p) for p in paths]
        fmax = max(sc.n_free for sc in scenes)
        deltas = _scene.perturbation_deltas(seed, n_worlds, fmax, dpos, dvel) if perturb else np.zeros((n_worlds, fmax, 4))
        specs, ids = [], []
        for k, sc in enumerate(scenes):
            idx = np.arange(k, n_worlds, len(scenes))
            if len(idx) == 0:
                continue
            d = np.ascontiguousarray(deltas[idx][:, :sc.n_free, :])
            specs.append(GroupSpec.from_scene(sc, len(idx), d, max_calls=max_calls))
            ids.append(idx)
        return cls(specs, ids, device=device)

    @classmethod
    def from_ga(cls, scene_path, n_worlds, seed=0, max_calls=4096, device=0, population=None):
        """GA population: the scene's fixed geometry, its free bodies replaced by ONE random
        convex polygon per world (fizz2d_b200.ga).  Worlds are grouped by vertex count (one kernel topology each)."""
        from . import ga as _ga
        sc = _scene.load_in(scene_path)
        nverts, pts, vel = population if population is not None else _ga.random_convex_polygons(seed, n_worlds)
        base = _scene.Scene(sc.width, sc.height, dict(sc.params), [], sc.fixed, sc.name)
        specs, ids = [], []
        for n in sorted(set(int(x) for x in nverts)):
            idx = np.nonzero(nverts == n)[0]
            P = np.stack([pts[w] for w in idx])
            specs.append(GroupSpec(base, [np.ascontiguousarray(P)], vel[idx][:, None, :].copy(),
                                   np.ones((len(idx), 1)), max_calls=max_calls))
            ids.append(idx)
        return cls(specs, ids, device=device)

    def close(self):
        for g in self.groups:
            g.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- stepping ----------------------------------------------------------------------------------------
    def _stream_ptr(self, k):
        st = self.streams[k] if k < len(self.streams) else None
        if st is None:
            import torch
            return torch.cuda.current_stream(self.device).cuda_stream
        return st.cuda_stream

    def _fork(self):
        """Mixed batches step their groups concurrently on streams of their own: every group stream first waits for
        what the caller's (torch's current) stream has queued so far ..."""
        if self.streams[0] is not None:
            import torch
            cur = torch.cuda.current_stream(self.device)
            for st in self.streams:
                st.wait_stream(cur)

    def _join(self):
        """... and the caller's stream then waits for the group streams: from outside, a BatchedWorld call is ordered
        on the current stream like any other torch operation (events recorded there bracket it)."""
        if self.streams[0] is not None:
            import torch
            cur = torch.cuda.current_stream(self.device)
            for st in self.streams:
                cur.wait_stream(st)

    def upload(self, sync=True):
        """(Re)load the initial conditions: host -> device, resets heat/status/step counters."""
        self._fork()
        for k, g in enumerate(self.groups):
            g.upload(self._stream_ptr(k))
        self._join()
        self.steps_requested = 0
        if sync:
            self.synchronize()

    def step(self, n_steps=1):
        """n_steps x World.update() for every world; one kernel launch per group (groups run concurrently)."""
        L = _lib.load()
        self._fork()
        for k, g in enumerate(self.groups):
            _lib.check(L.fizz_step(g.handle, int(n_steps), C.c_void_p(self._stream_ptr(k))))
        self._join()
        self.steps_requested += int(n_steps)

    def step_logged(self, n_steps):
        """n_steps x update() with the energy row of EVERY world recorded before each step (the batched energy.txt).
        Returns a float64 array (n_steps, W, 4): total, kinetic, potential, heat."""
        import torch
        L = _lib.load()
        logs = []
        self._fork()
        for k, g in enumerate(self.groups):
            t = torch.empty((int(n_steps), 4, g.layout.stride), dtype=torch.float64, device=g.slab.device)
            _lib.check(L.fizz_step_logged(g.handle, int(n_steps), C.c_void_p(t.data_ptr()), C.c_void_p(self._stream_ptr(k))))
            logs.append(t)
        self._join()
        self.steps_requested += int(n_steps)
        self.synchronize()
        out = np.empty((int(n_steps), self.n_worlds, 4))
        for g, t in zip(self.groups, logs):
            out[:, g.world_ids, :] = t[:, :, :g.spec.n_worlds].permute(0, 2, 1).cpu().numpy()
        return out

    def synchronize(self):
        import torch
        torch.cuda.synchronize(self.device)

    # -- readback ----------------------------------------------------------------------------------------
    def _gather(self, fn, shape_tail, dtype=np.float64):
        out = np.empty((self.n_worlds,) + shape_tail, dtype=dtype)
        for g in self.groups:
            out[g.world_ids] = fn(g)
        return out

    def _download(self, g, k, **want):
        s = g.spec
        W, F, V = s.n_worlds, s.n_free, sum(s.nverts_free)
        bufs = dict(pos=np.empty((2 * F, W)) if want.get("pos") else None,
                    vel=np.empty((2 * F, W)) if want.get("vel") else None,
                    heat=np.empty(W) if want.get("heat") else None,
                    energy=np.empty((4, W)) if want.get("energy") else None,
                    verts=np.empty((2 * V, W)) if want.get("verts") else None)
        ibufs = dict(status=np.empty(W, dtype=np.int64) if want.get("status") else None,
                     steps=np.empty(W, dtype=np.int64) if want.get("steps") else None,
                     calls=np.empty(W, dtype=np.int64) if want.get("calls") else None)

        def p(a, t):
            return a.ctypes.data_as(t) if a is not None else None
        _lib.check(_lib.load().fizz_download(
            g.handle, p(bufs["pos"], _lib.dp), p(bufs["vel"], _lib.dp), p(bufs["heat"], _lib.dp),
            p(bufs["energy"], _lib.dp), p(bufs["verts"], _lib.dp), p(ibufs["status"], _lib.lp),
            p(ibufs["steps"], _lib.lp), p(ibufs["calls"], _lib.lp), C.c_void_p(self._stream_ptr(k))))
        bufs.update(ibufs)
        return bufs

    def download(self, **want):
        self._fork()
        res = [self._download(g, k, **want) for k, g in enumerate(self.groups)]
        self._join()
        self.synchronize()
        return res

    def energy(self):
        """(W, 4): total, kinetic, potential, heat -- World.energy() of every world (physics.py:66-74)."""
        res = self.download(energy=True)
        out = np.empty((self.n_worlds, 4))
        for g, r in zip(self.groups, res):
            out[g.world_ids] = r["energy"].T
        return out

    def heat(self):
        res = self.download(heat=True)
        out = np.empty(self.n_worlds)
        for g, r in zip(self.groups, res):
            out[g.world_ids] = r["heat"]
        return out

    def status(self):
        res = self.download(status=True)
        out = np.empty(self.n_worlds, dtype=np.int64)
        for g, r in zip(self.groups, res):
            out[g.world_ids] = r["status"]
        return out

    def steps_done(self):
        res = self.download(steps=True)
        out = np.empty(self.n_worlds, dtype=np.int64)
        for g, r in zip(self.groups, res):
            out[g.world_ids] = r["steps"]
        return out

    def calls(self):
        res = self.download(calls=True)
        out = np.empty(self.n_worlds, dtype=np.int64)
        for g, r in zip(self.groups, res):
            out[g.world_ids] = r["calls"]
        return out

    def com(self):
        """List (one entry per group) of (W_g, F_g, 4) arrays: com.pos.x, com.pos.y, com.vel.x, com.vel.y.
        For a single-group batch returns the array itself."""
        res = self.download(pos=True, vel=True)
        outs = []
        for g, r in zip(self.groups, res):
            F = g.spec.n_free
            a = np.empty((g.spec.n_worlds, F, 4))
            a[:, :, 0] = r["pos"][0::2].T
            a[:, :, 1] = r["pos"][1::2].T
            a[:, :, 2] = r["vel"][0::2].T
            a[:, :, 3] = r["vel"][1::2].T
            outs.append(a)
        return outs[0] if len(outs) == 1 else outs

    def vertices(self):
        """List per group of (W_g, V_g, 2) free-vertex positions (point.pos) after the last step()."""
        res = self.download(verts=True)
        outs = []
        for g, r in zip(self.groups, res):
            V = sum(g.spec.nverts_free)
            a = np.empty((g.spec.n_worlds, V, 2))
            a[:, :, 0] = r["verts"][0::2].T
            a[:, :, 1] = r["verts"][1::2].T
            outs.append(a)
        return outs[0] if len(outs) == 1 else outs

    def dump_world(self, i, verts=None):
        """str(world) of world i (physics.py:291-301, :713-722, :600-606): int() truncation, clamp to the frame."""
        gi, k = self.locate(i)
        g = self.groups[gi]
        if verts is None:
            verts = self.vertices()
            verts = verts if len(self.groups) == 1 else verts[gi]
        return format_world(g.spec, verts[k])

    def render(self, worlds=None):
        """Frames of the given worlds (global indices; default: all) as they stand after the last step(): what the
        reference pipeline would draw from `str(world)` (physics.py:291-301 -> draw.c), rasterised on the GPU straight
        from the resident state -- vertices never leave the device.  Returns a torch.uint8 tensor [n, H, W] per group
        that holds any of the worlds (a single tensor when there is one group), rows in the order given."""
        import torch
        from . import raster
        worlds = list(range(self.n_worlds)) if worlds is None else [int(w) for w in worlds]
        per_group = {}
        for w in worlds:
            gi, k = self.locate(w)
            per_group.setdefault(gi, []).append(k)
        outs = []
        for gi in sorted(per_group):
            g = self.groups[gi]
            sp = g.spec
            # everything below (the library's kernels AND torch's indexing) runs on torch's current stream: one order
            cur = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.load().fizz_vertices(g.handle, C.c_void_p(cur)))
            V = sum(sp.nverts_free)
            idx = torch.tensor(per_group[gi], dtype=torch.int64, device=g.slab.device)
            v = g.section("verts", 2 * V)[:, idx]                       # [2V, n]: rows x0, y0, x1, y1, ...
            verts = v.t().reshape(len(per_group[gi]), V, 2)              # [n, V, 2]
            fixed = np.array([(max(min(int(p[0]), sp.width - 1), 0), max(min(int(p[1]), sp.height - 1), 0))
                              for pts in sp.scene.fixed for p in pts], dtype=np.int32).reshape(-1, 2)
            pts = raster.points_from_vertices(sp.width, sp.height, verts, fixed, device=g.device, stream=cur)
            nverts = list(sp.nverts_free) + [len(pts_) for pts_ in sp.scene.fixed]
            outs.append(raster.render_points(sp.width, sp.height, nverts, pts, device=g.device, stream=cur))
        return outs[0] if len(self.groups) == 1 else outs

    def locate(self, i):
        """(group index, index inside the group) of global world i."""
        if self._where is None:
            self._where = {}
            for gi, g in enumerate(self.groups):
                for k, wid in enumerate(g.world_ids):
                    self._where[int(wid)] = (gi, k)
        return self._where[int(i)]

    # -- contact trace (tests) ---------------------------------------------------------------------------
    def enable_trace(self, capacity=1 << 20):
        for g in self.groups:
            g.enable_trace(capacity)

    def contacts(self):
        out = []
        for g in self.groups:
            out += g.read_trace()
        out.sort(key=lambda r: r[:5])
        return out

    def set_tuning(self, mode="hybrid", chunk_steps=256):
        """Execution strategy (results never depend on it): "hybrid" = ballistic thread-per-world kernel +
        warp-per-world contact kernel per chunk of timesteps, trapped worlds on their own "runner" launches that do not
        stop at chunk boundaries; "hybrid_sync" = the same without runners; "hybrid3" = plus a thread-per-world
        separation prover tier (busy scenes, use small chunks); "thread" = one thread-per-world kernel;
        "thread_fp32" = that kernel in single precision, the one mode whose results DO differ (optional FP32 mode)."""
        m = {"hybrid": _lib.MODE_HYBRID, "thread": _lib.MODE_THREAD, "hybrid3": _lib.MODE_HYBRID3,
             "hybrid_sync": _lib.MODE_HYBRID_SYNC, "thread_fp32": _lib.MODE_THREAD_FP32}[mode]
        for g in self.groups:
            _lib.check(_lib.load().fizz_set_tuning(g.handle, m, int(chunk_steps)))

    def snapshot(self):
        """Device-resident copy of the dynamic state of every group (replay without a host round trip)."""
        import torch
        snaps = []
        self._fork()
        for k, g in enumerate(self.groups):
            t = torch.empty(g.layout.state_words, dtype=torch.float64, device=g.slab.device)
            _lib.check(_lib.load().fizz_snapshot(g.handle, C.c_void_p(t.data_ptr()), C.c_void_p(self._stream_ptr(k))))
            snaps.append(t)
        self._join()
        return (snaps, self.steps_requested)

    def restore(self, snap):
        snaps, steps = snap
        self._fork()
        for k, (g, t) in enumerate(zip(self.groups, snaps)):
            _lib.check(_lib.load().fizz_restore(g.handle, C.c_void_p(t.data_ptr()), int(steps),
                                                C.c_void_p(self._stream_ptr(k))))
        self._join()
        self.steps_requested = steps

    def set_profiling(self, enabled=True):
        for g in self.groups:
            _lib.check(_lib.load().fizz_set_profiling(g.handle, 1 if enabled else 0))

    def profile(self, reset=True):
        """Per group: dict kernel -> (milliseconds, launches), device-timed with CUDA events inside fizz_step."""
        out = []
        for g in self.groups:
            ms = (C.c_double * 3)()
            n = (C.c_int64 * 3)()
            _lib.check(_lib.load().fizz_profile_read(g.handle, ms, n, 1 if reset else 0))
            out.append({"thread": (ms[0], n[0]), "ballistic": (ms[1], n[1]), "contact": (ms[2], n[2])})
        return out

    def counters(self, reset=True):
        """Per group: device-side work counters since the last reset (hybrid mode)."""
        out = []
        for k, g in enumerate(self.groups):
            a = (C.c_int64 * 4)()
            _lib.check(_lib.load().fizz_counters(g.handle, a, 1 if reset else 0, C.c_void_p(self._stream_ptr(k))))
            out.append(dict(ballistic_world_steps=a[0], contact_world_steps=a[1], contact_calls=a[2],
                            fast_path_calls=a[3]))
        return out

    def set_spec(self, enabled=-1, window0=0, slot_cost=0, min_cost=0):
        """Spec rounds (time-parallel stepping of expensive worlds in mode "hybrid"; include/fizz_b200.h: fizz_set_spec).
        Results never depend on these settings."""
        for g in self.groups:
            _lib.check(_lib.load().fizz_set_spec(g.handle, int(enabled), int(window0), int(slot_cost), int(min_cost)))
        return self

    def spec_stats(self, reset=True):
        """Per group: what the spec rounds did since the last reset."""
        out = []
        for k, g in enumerate(self.groups):
            a = (C.c_int64 * 4)()
            _lib.check(_lib.load().fizz_spec_stats(g.handle, a, 1 if reset else 0, C.c_void_p(self._stream_ptr(k))))
            out.append(dict(timesteps_accepted=a[0], timesteps_evaluated=a[1], rounds=a[2], rounds_with_rejection=a[3]))
        return out

    def kernel_info(self):
        info = []
        for g in self.groups:
            d = {}
            for name, which in (("thread", _lib.KERNEL_THREAD), ("ballistic", _lib.KERNEL_BALLISTIC),
                                ("prover", _lib.KERNEL_PROVER), ("contact", _lib.KERNEL_CONTACT),
                                ("trapped", _lib.KERNEL_TRAPPED)):
                r, s, t, b = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
                n = C.c_int64()
                _lib.check(_lib.load().fizz_kernel_info(g.handle, which, C.byref(r), C.byref(s), C.byref(t), C.byref(b),
                                                        C.byref(n)))
                d[name] = dict(regs=r.value, smem=s.value, threads=t.value, blocks_per_sm=b.value)
                d["launches"] = n.value
            info.append(d)
        return info


def _fmt_point(x, y, w, h):
    # Point.__str__ (physics.py:600-606)
    bx = max(min(int(x), w - 1), 0)
    by = max(min(int(y), h - 1), 0)
    return "point," + str(bx) + "," + str(by)


def format_world(spec: GroupSpec, verts: np.ndarray) -> str:
    """World.__str__ for one world given its (V,2) free-vertex array."""
    out = str(spec.width) + "," + str(spec.height) + "\n"
    o = 0
    for n in spec.nverts_free:
        out += "polygon\nsides," + str(n) + "\n" + \
            "\n".join(_fmt_point(verts[o + k, 0], verts[o + k, 1], spec.width, spec.height) for k in range(n)) + "\n"
        o += n
    for pts in spec.scene.fixed:
        out += "polygon\nsides," + str(len(pts)) + "\n" + \
            "\n".join(_fmt_point(p[0], p[1], spec.width, spec.height) for p in pts) + "\n"
    return out


def fp64_peak(device=0, iters=4096):
    """Measured FP64 FMA throughput of the device in TFLOP/s (roofline denominator)."""
    t, ms = C.c_double(), C.c_double()
    _lib.check(_lib.load().fizz_fp64_peak(device, iters, C.byref(t), C.byref(ms)))
    return t.value, ms.value
