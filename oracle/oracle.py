"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

* `parse_scene` / `init_constants`: literal scalar numpy restatement of the reference's scene loader and
  init-time math (`main.read_input` main.py:23-83, `Obj.__init__` physics.py:305-348, `Polygon.__init__`
  physics.py:610-644, `moment_of_inertia` :664-711, `triangle_area` :843-847,
  `triangle_moment_of_inertia` :849-868).  Written independently of the product loader
  (`fizz2d_b200/scene.py`) so the two can be checked against each other and against tests/golden/.
* `OracleWorld`: ctypes binding of `oracle/libfizz_oracle.so` (fizz_oracle.c), one world per object.

Only tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline / `--impl reference` legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
from numpy.linalg import norm

HERE = os.path.dirname(os.path.abspath(__file__))
# numpy's 2-vector dot/norm contract the oracle is built for: $FZO_DOT_MODE (0, 1 = default, 2)
DOT_MODE = int(os.environ.get("FZO_DOT_MODE", "1") or 1)
LIB_PATH = os.path.join(HERE, "libfizz_oracle.so" if DOT_MODE == 1 else "libfizz_oracle_dot%d.so" % DOT_MODE)

DEFAULT_PARAMS = dict(ELASTICITY=1.0, EPSILON=1e-12, TIME_TOL=1e-10, COLLISION_TOL=.1, VIBRATE_TOL=1e-5)
TIME_DISC = .05


def build(force=False):
    if force or not os.path.isfile(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "fizz_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []) + ([] if DOT_MODE == 1 else ["dotmodes"]))
    return LIB_PATH


# ------------------------------------------------------------------------------------------------------
# scene loading (main.py:23-83) -- scalar, literal
# ------------------------------------------------------------------------------------------------------
def parse_scene(text):
    """Returns dict(width,height,params{NAME:float},free[list of dict(points,density,vel)],fixed[list of points])."""
    lines = text.split("\n")
    w, h = [int(x) for x in lines[0].strip().split(",")]
    params = {}
    free, fixed = [], []
    shape = ""
    vel = None
    density = np.inf
    for raw in lines[1:]:
        line = raw.strip()
        if not line or line[0] == "#":
            continue
        words = line.split()
        if words[0] == "PARAM":
            params[words[1]] = float(words[2])
            continue
        if shape == "circle":
            raise NotImplementedError("circle: Ball.__init__ raises NameError in the reference (physics.py:735)")
        elif shape in ("polygon", "fixedpolygon"):
            data = line.split(",")
            if data[0] == "density":
                density = float(data[1])
            elif data[0] == "sides":
                npoints = int(data[1])
                points = []
            elif data[0] == "velocity":
                vel = [float(data[1]), float(data[2])]
            elif data[0] == "point":
                npoints -= 1
                points.append([float(_) for _ in data[1:]])
                if npoints == 0:
                    if shape == "polygon":
                        if density == np.inf:
                            density = 1
                        free.append(dict(points=points, density=density, vel=list(vel)))
                    else:
                        fixed.append(points)
                    shape = ""
        else:
            vel = [0.0, 0.0]
            shape = line
            density = np.inf
    return dict(width=w, height=h, params=params, free=free, fixed=fixed)


def _triangle_area(pA, pB, pC):
    return .5 * abs(pA[0]*pB[1] + pB[0]*pC[1] + pC[0]*pA[1] - pA[0]*pC[1] - pC[0]*pB[1] - pB[0]*pA[1])


def _triangle_moi(p0, p1, p2, eps):
    if abs(p2[0] - p0[0]) < eps:
        p1, p2, p0 = p0, p1, p2
    with np.errstate(all="ignore"):
        slope = (p2[1] - p0[1]) / (p2[0] - p0[0])
        b = norm(p2 - p0)
        xmin = (slope**2 * p0[0] + slope * p1[1] - slope * p0[1] + p1[0]) / (slope**2 + 1)
        pmin = np.array([xmin, slope*(xmin - p0[0]) + p0[1]])
        a = norm(p0 - pmin)
        h = norm(p1 - pmin)
        return ((b**3)*h - (b**2)*h*a + b*h*(a**2) + b*(h**3)) / 36


def body_constants(points, density, eps=1e-12):
    """Init-time constants of one polygon, op for op as the reference derives them."""
    pts = [np.array(p, dtype=float) for p in points]
    com = sum(pts) / len(pts)                                  # physics.py:314
    radius = max([norm(p - com) for p in pts])                 # :322
    out = dict(com=com, radius=float(radius), n=len(pts))
    if density is not None:
        if len(pts) == 3:                                      # :671-677
            area = _triangle_area(*pts)
            moi = _triangle_moi(*pts, eps)
        else:                                                  # :679-707
            area = 0.0
            moi = 0.0
            wm = 0.0
            for p1, p2 in zip(pts, pts[1:] + [pts[0]]):
                it = _triangle_moi(com, p1, p2, eps)
                c3 = (com + p1 + p2) / 3
                dsq = norm(c3 - com)
                a = _triangle_area(com, p1, p2)
                area += a
                moi += it
                wm += a * dsq
            with np.errstate(all="ignore"):
                wm *= (1 / area)                               # Obj.mass is still 1 here (:611,:706)
            moi += wm
        out["area"] = float(area)
        out["mass"] = float(area * density)                    # :624
        out["moi"] = float(moi)
    offx, offy = [], []
    with np.errstate(all="ignore"):
        for p in pts:                                          # :631-635, :358-359
            r = p - com
            phi = np.arccos(r[0] / norm(r))
            if r[1] > 0:
                phi = 2*np.pi - phi
            offx.append(float(norm(r) * np.cos(phi + 0.0)))
            offy.append(float(norm(r) * np.sin(phi + 0.0)))
    out["offx"] = offx
    out["offy"] = offy
    return out


class _Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("elasticity", "epsilon", "time_tol", "collision_tol", "vibrate_tol",
                                          "time_disc", "gx", "gy", "one_plus_e", "one_minus_e2", "height",
                                          "friction_mu", "ff_restitution")] + \
               [("max_calls", C.c_int)]


class _Contact(C.Structure):
    _fields_ = [("step", C.c_int), ("call", C.c_int), ("i", C.c_int), ("j", C.c_int),
                ("nx", C.c_double), ("ny", C.c_double), ("mag", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.fzo_create.restype = C.c_void_p
        L.fzo_create.argtypes = [C.c_int, C.c_int, ip, dp, dp, dp, dp, dp, dp, dp, C.POINTER(_Params)]
        L.fzo_destroy.argtypes = [C.c_void_p]
        L.fzo_update.argtypes = [C.c_void_p]; L.fzo_update.restype = C.c_int
        L.fzo_run.argtypes = [C.c_void_p, C.c_long]; L.fzo_run.restype = C.c_long
        L.fzo_energy.argtypes = [C.c_void_p, dp]
        L.fzo_com.argtypes = [C.c_void_p, dp]
        L.fzo_verts.argtypes = [C.c_void_p, dp]
        L.fzo_heat.argtypes = [C.c_void_p]; L.fzo_heat.restype = C.c_double
        L.fzo_steps_done.argtypes = [C.c_void_p]; L.fzo_steps_done.restype = C.c_long
        L.fzo_status.argtypes = [C.c_void_p]; L.fzo_status.restype = C.c_int
        L.fzo_calls_last_step.argtypes = [C.c_void_p]; L.fzo_calls_last_step.restype = C.c_long
        L.fzo_total_calls.argtypes = [C.c_void_p]; L.fzo_total_calls.restype = C.c_long
        L.fzo_total_ops.argtypes = [C.c_void_p]; L.fzo_total_ops.restype = C.c_double
        L.fzo_dot_mode.restype = C.c_int
        L.fzo_trace.argtypes = [C.c_void_p, C.POINTER(_Contact), C.c_long]
        L.fzo_contact_count.argtypes = [C.c_void_p]; L.fzo_contact_count.restype = C.c_long
        L.fzo_run_batch.argtypes = [C.POINTER(C.c_void_p), C.c_long, C.c_long, C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_params(params=None, height=0, max_calls=4096, gravity=(0.0, 0.0), friction=None, ff_restitution=None):
    """friction / ff_restitution: the default-off extensions; None = what `PARAM FRICTION` / `PARAM FF_RESTITUTION`
    lines of the scene say, else off (0 / -1)."""
    p = dict(DEFAULT_PARAMS)
    p.update(params or {})
    e = p["ELASTICITY"]
    mu = float(p.get("FRICTION", 0.0) if friction is None else friction)
    ffr = float(p.get("FF_RESTITUTION", -1.0) if ff_restitution is None else ff_restitution)
    return _Params(friction_mu=mu, ff_restitution=ffr, elasticity=e, epsilon=p["EPSILON"], time_tol=p["TIME_TOL"], collision_tol=p["COLLISION_TOL"],
                   vibrate_tol=p["VIBRATE_TOL"], time_disc=TIME_DISC, gx=float(gravity[0]), gy=float(gravity[1]),
                   one_plus_e=(1 + e), one_minus_e2=(1 - e**2), height=float(height), max_calls=int(max_calls))


class OracleWorld:
    """One world stepped by the C oracle.  Build from scene text (own parser + init math) or from explicit
    constants (e.g. the golden fixtures' init constants, or the product loader's arrays)."""

    def __init__(self, nverts_free, nverts_fixed, com, vel, radius, mass, offx, offy, fixedxy, prm,
                 width=0, height=0, trace_cap=0):
        L = lib()
        self.nfree, self.nfixed = len(nverts_free), len(nverts_fixed)
        self.nverts_free = list(nverts_free)
        self.nverts_fixed = list(nverts_fixed)
        self.width, self.height = width, height
        self.fixedxy = np.ascontiguousarray(fixedxy, dtype=np.float64).reshape(-1, 2)
        nv = np.ascontiguousarray(list(nverts_free) + list(nverts_fixed), dtype=np.int32)
        arrs = [np.ascontiguousarray(a, dtype=np.float64).ravel() for a in (com, vel, radius, mass, offx, offy)]
        self._prm = prm
        self._h = L.fzo_create(self.nfree, self.nfixed, nv.ctypes.data_as(C.POINTER(C.c_int)),
                               *[_dp(a) for a in arrs], _dp(self.fixedxy), C.byref(prm))
        if not self._h:
            raise ValueError("fzo_create failed (too many vertices?)")
        self._trace = None
        if trace_cap:
            self._trace = (_Contact * trace_cap)()
            L.fzo_trace(self._h, self._trace, trace_cap)

    @classmethod
    def from_text(cls, text, max_calls=4096, gravity=(0.0, 0.0), trace_cap=0):
        sc = parse_scene(text)
        eps = sc["params"].get("EPSILON", DEFAULT_PARAMS["EPSILON"])
        fr = [body_constants(b["points"], b["density"], eps) for b in sc["free"]]
        fx = [body_constants(pts, None) for pts in sc["fixed"]]
        com = [c["com"] for c in fr] + [c["com"] for c in fx]
        vel = [b["vel"] for b in sc["free"]]
        radius = [c["radius"] for c in fr] + [c["radius"] for c in fx]
        mass = [c["mass"] for c in fr]
        offx = [x for c in fr for x in c["offx"]]
        offy = [x for c in fr for x in c["offy"]]
        fixedxy = [p for pts in sc["fixed"] for p in pts]
        prm = make_params(sc["params"], sc["height"], max_calls, gravity)
        w = cls([c["n"] for c in fr], [c["n"] for c in fx], com, vel, radius, mass, offx, offy,
                fixedxy if fixedxy else np.zeros((0, 2)), prm, sc["width"], sc["height"], trace_cap)
        w.moi = [c["moi"] for c in fr]
        return w

    @classmethod
    def from_file(cls, path, **kw):
        with open(path) as f:
            return cls.from_text(f.read(), **kw)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fzo_destroy(self._h)
            self._h = None

    def update(self):
        return lib().fzo_update(self._h) == 0

    def run(self, n):
        return lib().fzo_run(self._h, n)

    def energy(self):
        out = np.zeros(4)
        lib().fzo_energy(self._h, _dp(out))
        return out

    def com(self):
        out = np.zeros((self.nfree, 4))
        lib().fzo_com(self._h, _dp(out))
        return out

    def verts(self):
        out = np.zeros((sum(self.nverts_free), 2))
        lib().fzo_verts(self._h, _dp(out))
        return out

    def heat(self):
        return lib().fzo_heat(self._h)

    def steps_done(self):
        return lib().fzo_steps_done(self._h)

    def status(self):
        return lib().fzo_status(self._h)

    def calls_last_step(self):
        return lib().fzo_calls_last_step(self._h)

    def total_calls(self):
        return lib().fzo_total_calls(self._h)

    def total_ops(self):
        return lib().fzo_total_ops(self._h)

    def contacts(self):
        n = lib().fzo_contact_count(self._h)
        if n > (len(self._trace) if self._trace else 0):
            raise OverflowError("oracle contact trace overflow: %d records, capacity %d"
                                % (n, len(self._trace) if self._trace else 0))
        return [(c.step, c.call, c.i, c.j, c.nx, c.ny, c.mag) for c in self._trace[:n]]

    def dump(self):
        """World.__str__ (physics.py:291-301, :713-722, :600-606)."""
        out = "%d,%d\n" % (self.width, self.height)
        v = self.verts()
        o = 0
        for n in self.nverts_free:
            out += "polygon\nsides,%d\n" % n + "\n".join(_pt(v[o + k], self.width, self.height) for k in range(n)) + "\n"
            o += n
        o = 0
        for n in self.nverts_fixed:
            out += "polygon\nsides,%d\n" % n + "\n".join(_pt(self.fixedxy[o + k], self.width, self.height) for k in range(n)) + "\n"
            o += n
        return out


def _pt(p, w, h):
    x = max(min(int(p[0]), w - 1), 0)
    y = max(min(int(p[1]), h - 1), 0)
    return "point,%d,%d" % (x, y)


def run_batch(worlds, nsteps, nthreads):
    arr = (C.c_void_p * len(worlds))(*[w._h for w in worlds])
    lib().fzo_run_batch(arr, len(worlds), nsteps, nthreads)
