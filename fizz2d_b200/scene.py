"""Host-side scene loader: the `.in` grammar of the reference (`main.read_input`, main.py:23-83) and the
init-time math of `Obj.__init__` / `Polygon.__init__` (physics.py:305-348, :610-644, :664-711, :843-868),
vectorised over worlds.

Everything here runs once, on the host, in numpy -- exactly the arithmetic the reference performs at load time,
so the per-vertex polar offsets |r|cos(phi), |r|sin(phi), radii and masses the device receives are the constants
the reference would recompute every step.  numpy's 2-vector `norm` is sqrt(fma(y,y,x*x)) (SURVEY
Q14); the fused form is provided by host helpers in libfizz_b200.so (no GPU needed for those).
"""
from __future__ import annotations

import dataclasses
import math
import warnings
from typing import List, Optional

import numpy as np

from . import lib as _lib

# defaults of the reference (physics.py:14-28)
DEFAULT_PARAMS = dict(ELASTICITY=1.0, EPSILON=1e-12, TIME_TOL=1e-10, COLLISION_TOL=.1, VIBRATE_TOL=1e-5)
TIME_DISC = .05
# names `PARAM` may set that the hot path reads at call time; others have no effect there
_LIVE_PARAMS = set(DEFAULT_PARAMS)
_INERT_PARAMS = {"TIME_DISC", "GRAVITY", "GLOBALROTSPEED", "DENSITY", "CONTACT_TOL", "ANGLE_TOL"}
# default-off extensions of this engine (csrc/fizz_ext.cuh); the reference's parser would setattr() them without effect
_EXT_PARAMS = {"FRICTION", "FF_RESTITUTION"}


@dataclasses.dataclass
class FreePolygon:
    points: np.ndarray          # (n, 2) float64, as written in the file
    density: float
    velocity: np.ndarray        # (2,)


@dataclasses.dataclass
class Scene:
    """Parsed `.in` file: what main.read_input() turns into World/Polygon/FixedPolygon objects."""
    width: int
    height: int
    params: dict
    free: List[FreePolygon]
    fixed: List[np.ndarray]     # each (n, 2)
    name: str = ""

    @property
    def n_free(self):
        return len(self.free)

    @property
    def n_fixed(self):
        return len(self.fixed)

    def topology(self):
        return (tuple(len(b.points) for b in self.free), tuple(len(p) for p in self.fixed))


def parse_in(text: str, name: str = "") -> Scene:
    """Same grammar, same quirks as main.read_input (main.py:23-83): first line "W,H"; blank / '#' lines
    skipped; `PARAM NAME VALUE`; a shape header line starts a shape and resets velocity to (0,0) and density to
    inf; inside a polygon `density,D` `sides,N` `velocity,vx,vy` `point,x,y`; other keys (e.g. `mass,5`) are
    silently ignored; a polygon without density gets 1 (main.py:72-74)."""
    lines = text.split("\n")
    w, h = [int(x) for x in lines[0].strip().split(",")]
    params = {}
    free, fixed = [], []
    shape = ""
    vel = np.array([0.0, 0.0])
    density = math.inf
    npoints = 0
    points = []
    for raw in lines[1:]:
        line = raw.strip()
        if not line or line[0] == "#":
            continue
        words = line.split()
        if words[0] == "PARAM":
            name_, val = words[1], float(words[2])
            if name_ not in _LIVE_PARAMS and name_ not in _EXT_PARAMS:
                if name_ in _INERT_PARAMS:
                    warnings.warn("PARAM %s has no effect on the stepping path (bound at definition time in the "
                                  "reference, physics.py:32-33)" % name_)
                else:
                    warnings.warn("PARAM %s is not read by the stepping path; ignored" % name_)
            params[name_] = val
            continue
        if shape == "circle":
            raise NotImplementedError("circle shapes are unusable in the reference (Ball.__init__ raises "
                                      "NameError, physics.py:735)")
        elif shape in ("polygon", "fixedpolygon"):
            data = line.split(",")
            if data[0] == "density":
                density = float(data[1])
            elif data[0] == "sides":
                npoints = int(data[1])
                points = []
            elif data[0] == "velocity":
                vel = np.array([float(data[1]), float(data[2])])
            elif data[0] == "point":
                npoints -= 1
                points.append([float(_) for _ in data[1:]])
                if npoints == 0:
                    if shape == "polygon":
                        if density == math.inf:
                            warnings.warn("density never set for polygon, set to 1")
                            density = 1
                        free.append(FreePolygon(np.array(points, dtype=float), float(density), vel.copy()))
                    else:
                        fixed.append(np.array(points, dtype=float))
                    shape = ""
        else:
            vel = np.array([0.0, 0.0])
            shape = line
            density = math.inf
    return Scene(w, h, params, free, fixed, name)


def load_in(path: str) -> Scene:
    with open(path) as f:
        return parse_in(f.read(), name=path)


# ----------------------------------------------------------------------------------------------------------
# numpy 2-vector contract helpers (host code inside libfizz_b200.so)
# ----------------------------------------------------------------------------------------------------------
def norm2(x, y):
    """np.linalg.norm([x, y]) elementwise == sqrt(fma(y, y, x*x))."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(x)
    _lib.load().fizz_host_norm2(x.size, _lib.as_dp(x), _lib.as_dp(y), _lib.as_dp(out))
    return out


def sqrt_threshold(r):
    """T(r) = largest double s with sqrt(s) <= r, so `norm(d) > r` (physics.py:278) == `fma(dy,dy,dx*dx) > T(r)`."""
    r = np.ascontiguousarray(r, dtype=np.float64)
    out = np.empty_like(r)
    _lib.load().fizz_host_sqrt_threshold(r.size, _lib.as_dp(r), _lib.as_dp(out))
    return out


# ----------------------------------------------------------------------------------------------------------
# init-time constants, vectorised over worlds
# ----------------------------------------------------------------------------------------------------------
def _triangle_area(pA, pB, pC):
    # physics.py:843-847, left to right
    return .5 * np.abs(pA[..., 0]*pB[..., 1] + pB[..., 0]*pC[..., 1] + pC[..., 0]*pA[..., 1]
                       - pA[..., 0]*pC[..., 1] - pC[..., 0]*pB[..., 1] - pB[..., 0]*pA[..., 1])


def _triangle_moi(p0, p1, p2, eps):
    # physics.py:849-868 (only used to detect NaN/zero inertia)
    cyc = np.abs(p2[..., 0] - p0[..., 0]) < eps
    c = cyc[..., None]
    q0 = np.where(c, p2, p0)
    q1 = np.where(c, p0, p1)
    q2 = np.where(c, p1, p2)
    with np.errstate(all="ignore"):
        slope = (q2[..., 1] - q0[..., 1]) / (q2[..., 0] - q0[..., 0])
        b = norm2(q2[..., 0] - q0[..., 0], q2[..., 1] - q0[..., 1])
        xmin = (slope**2 * q0[..., 0] + slope * q1[..., 1] - slope * q0[..., 1] + q1[..., 0]) / (slope**2 + 1)
        ymin = slope*(xmin - q0[..., 0]) + q0[..., 1]
        a = norm2(q0[..., 0] - xmin, q0[..., 1] - ymin)
        h = norm2(q1[..., 0] - xmin, q1[..., 1] - ymin)
        return ((b**3)*h - (b**2)*h*a + b*h*(a**2) + b*(h**3)) / 36


def polygon_constants(P: np.ndarray, density: Optional[np.ndarray], eps: float = 1e-12):
    """Init-time constants of one polygon slot for W worlds.  P: (W, n, 2) vertex coordinates.
    Returns dict(com (W,2), radius (W,), offx/offy (W,n) and, for free polygons, area, mass, moi, valid)."""
    P = np.ascontiguousarray(P, dtype=np.float64)
    W, n, _ = P.shape
    com = 0 + P[:, 0, :]                      # builtin sum() starts from int 0 (physics.py:314)
    for k in range(1, n):
        com = com + P[:, k, :]
    com = com / n
    r = P - com[:, None, :]                   # point.radius (physics.py:632)
    rn = norm2(r[..., 0], r[..., 1])          # norm(point.radius)
    radius = rn.max(axis=1)                   # physics.py:322
    out = dict(com=com, radius=radius)
    with np.errstate(all="ignore"):
        phi = np.arccos(r[..., 0] / rn)       # physics.py:633
        phi = np.where(r[..., 1] > 0, 2*np.pi - phi, phi)      # :634-635
        out["offx"] = rn * np.cos(phi + 0.0)  # physics.py:358  (rotate(angle=0.0))
        out["offy"] = rn * np.sin(phi + 0.0)  # physics.py:359
    if density is not None:
        if n == 3:                            # physics.py:671-677
            area = _triangle_area(P[:, 0], P[:, 1], P[:, 2])
            moi = _triangle_moi(P[:, 0], P[:, 1], P[:, 2], eps)
        else:                                 # physics.py:679-707 (fan about the centroid, wrap-around edge included)
            area = np.zeros(W)
            moi = np.zeros(W)
            wm = np.zeros(W)
            for k in range(n):
                p1, p2 = P[:, k], P[:, (k + 1) % n]
                it = _triangle_moi(com, p1, p2, eps)
                c3 = (com + p1 + p2) / 3
                dsq = norm2(c3[:, 0] - com[:, 0], c3[:, 1] - com[:, 1])
                a = _triangle_area(com, p1, p2)
                area = area + a
                moi = moi + it
                wm = wm + a * dsq
            with np.errstate(all="ignore"):
                moi = moi + wm * (1 / area)
        out["area"] = area
        out["mass"] = area * density          # physics.py:624
        out["moi"] = moi
        # move_angular divides 0 by the inertia (physics.py:545): NaN or 0 poisons every vertex
        out["valid"] = np.isfinite(moi) & (moi != 0) & np.isfinite(out["offx"]).all(axis=1) & \
            np.isfinite(out["offy"]).all(axis=1) & np.isfinite(out["mass"]) & (out["mass"] > 0)
    return out


def perturbation_deltas(seed: int, n_worlds: int, n_free: int, dpos: float = 5.0, dvel: float = 30.0):
    """SURVEY 8d perturbation law.  World 0 is the unperturbed scene.  For worlds w >= 1, in world order and,
    inside a world, in body (file) order, four uniforms are consumed from Generator(PCG64(seed)):
    dx, dy ~ U(-dpos, dpos) added to every vertex of the body, dvx, dvy ~ U(-dvel, dvel) added to its velocity.
    Returns (n_worlds, n_free, 4)."""
    d = np.zeros((n_worlds, n_free, 4))
    if n_worlds > 1:
        rng = np.random.Generator(np.random.PCG64(seed))
        u = rng.random((n_worlds - 1, n_free, 4))
        lo = np.array([-dpos, -dpos, -dvel, -dvel])
        hi = np.array([dpos, dpos, dvel, dvel])
        d[1:] = lo + (hi - lo) * u
    return d


def resolve_params(scene_params: dict, gravity=(0.0, 0.0), height=0, max_calls=4096, friction=None, ff_restitution=None):
    """FizzParams with the derived fields evaluated as the Python float expressions of the reference.
    friction / ff_restitution: the default-off extensions (Coulomb coefficient; restitution between two free bodies);
    None = what the scene's `PARAM FRICTION` / `PARAM FF_RESTITUTION` lines say, else off."""
    p = dict(DEFAULT_PARAMS)
    p.update({k: v for k, v in scene_params.items() if k in _LIVE_PARAMS})
    e = p["ELASTICITY"]
    gx, gy = float(gravity[0]), float(gravity[1])
    mu = float(scene_params.get("FRICTION", 0.0) if friction is None else friction)
    ffr = float(scene_params.get("FF_RESTITUTION", -1.0) if ff_restitution is None else ff_restitution)
    if not mu >= 0.0:
        raise ValueError("friction must be >= 0")
    if ffr > 1.0:
        raise ValueError("ff_restitution must be in [0, 1] (or negative: off)")
    return _lib.FizzParams(friction_mu=mu, ff_restitution=ffr, 
        elasticity=e, epsilon=p["EPSILON"], time_tol=p["TIME_TOL"], collision_tol=p["COLLISION_TOL"],
        vibrate_tol=p["VIBRATE_TOL"], time_disc=TIME_DISC, gx=gx, gy=gy,
        one_plus_e=(1 + e),                    # physics.py:188
        one_minus_e2=(1 - e**2),               # physics.py:191
        acc_x=(0 + 1 * gx) / 1,                # physics.py:60,516  sum([1*global_accel]) / com.mass
        acc_y=(0 + 1 * gy) / 1,
        height=float(height), max_calls=int(max_calls), reserved=0)
