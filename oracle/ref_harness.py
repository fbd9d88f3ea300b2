"""TEST INFRASTRUCTURE ONLY -- harness that imports the UNMODIFIED reference (read-only at
/root/reference) through a small shim and records full-precision traces of `World.update()`.

It exists to (1) pin the C restatement in `oracle/fizz_oracle.c`, (2) generate the golden fixtures
committed under `tests/golden/` (see `oracle/make_golden.py`) and (3) time the reference itself as bench.py's CPU
baseline (`oracle/cpu_arm.py`).  It runs where `/root/reference` exists (the build container) or, on the GPU box,
from the verbatim copy `make -C oracle ref` leaves under the git-ignored `oracle/_ref/pyref/`; no `-m gpu` test and
no smoke() imports it.

Shim (no reference file is edited or copied):
  * `matplotlib` is stubbed (main.py:19 -> plot.py:1 imports it; absent here),
  * `physics.verbosity` / `physics.energylog` module globals are injected (physics.py:40,81,94 read them,
    main.py:91,93 normally sets them),
  * `World.init_obj` (physics.py:56-61) is replaced by an equivalent that skips the
    `if self.global_damping_force:` truth test on an empty ndarray (a hard error on numpy >= 2.2),
  * `side_contact` (physics.py:871-925) is replaced by a no-op: its result is discarded at
    physics.py:148-153 and it only emits numpy deprecation warnings.
The watchdog counts the state-relevant `check_collisions` calls of one `update()` (call sites
physics.py:124, :137, :234) and raises `Capped` instead of making call number `max_calls + 1` from the
resolve loop -- the same rule the CUDA engine and the C oracle apply.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("FIZZ_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")
SCENES = os.path.join(REF_ROOT, "simulations")
if not os.path.isfile(os.path.join(REF_SRC, "physics.py")):
    # the GPU box has no /root/reference: `make -C oracle ref` left a verbatim copy of the reference's Python sources
    # under oracle/_ref/pyref/ (git-ignored, travels with the gpurun snapshot) for bench.py's CPU-baseline leg
    REF_SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "pyref")

# line numbers of the check_collisions() call sites inside World.update (physics.py)
_SITE_TRIAL, _SITE_FALLBACK, _SITE_RESOLVE = 124, 137, 234
_SITE_DIAG = (243, 253)


class Capped(Exception):
    """Raised by the watchdog when the resolve loop wants more than max_calls check_collisions."""


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "physics.py"))


def load_physics():
    """(Re)import the reference's physics/main modules with the shim applied. PARAM values leak across
    read_input() calls inside one module instance (main.py:39-41), hence the reload per scene."""
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        for name in ("plot", "legend", "show"):
            setattr(plt, name, lambda *a, **k: None)
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    for mod in ("physics", "main", "plot", "config"):
        sys.modules.pop(mod, None)
    physics = importlib.import_module("physics")
    physics.verbosity = 0
    physics.energylog = 0

    def init_obj(self, obj):
        obj.forces.append(obj.mass * self.global_accel)
        self.objs.append(obj)

    physics.World.init_obj = init_obj
    physics.side_contact = lambda a, b: (0, None, None, None)
    main = importlib.import_module("main")
    return physics, main


class RefWorld:
    """One reference world with tracing + watchdog."""

    def __init__(self, path=None, max_calls=4096, text=None, gravity=None):
        self.physics, self.main = load_physics()
        self.max_calls = max_calls
        if gravity is not None:
            # gravity extension oracle: World() binds GRAVITY as a default argument at def
            # time, so patch both the default used by read_input's World(width,height) and p_energy's global
            g = np.array(gravity, dtype=float)
            self.physics.GRAVITY = g
            self.physics.World.__init__.__defaults__ = (set(), g, self.physics.TIME_DISC, [])
        if text is not None:
            import tempfile
            with tempfile.NamedTemporaryFile("w", suffix=".in", delete=False) as f:
                f.write(text)
                path = f.name
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    self.world = self.main.read_input(path)
            finally:
                os.unlink(path)
        else:
            with contextlib.redirect_stdout(io.StringIO()):
                self.world = self.main.read_input(path)
        self.step_index = 0
        self.capped_at = -1
        self.contacts = []   # (step, call#, site, i, j, nx, ny, mag)
        self.calls_per_step = []
        self._ncalls = 0
        self._ncalls_all = 0
        w = self.world
        orig = w.check_collisions
        harness = self

        def traced():
            site = sys._getframe(1).f_lineno
            if site not in _SITE_DIAG:
                if site == _SITE_RESOLVE and harness._ncalls >= harness.max_calls:
                    raise Capped()
                harness._ncalls += 1
            harness._ncalls_all += 1
            res = orig()
            objects = w.objs + w.fixed_objs
            for (a, b, n, mag) in res:
                harness.contacts.append((harness.step_index, harness._ncalls_all, site,
                                         objects.index(a), objects.index(b),
                                         float(n[0]), float(n[1]), float(mag)))
            return res

        w.check_collisions = traced

    # -- state accessors -------------------------------------------------------------------------
    def com(self):
        return np.array([[o.com.pos[0], o.com.pos[1], o.com.vel[0], o.com.vel[1]] for o in self.world.objs])

    def verts(self):
        return [np.array([p.pos for p in o.points], dtype=float) for o in self.world.objs]

    def energy(self):
        return np.array(self.world.energy(), dtype=float)

    def heat(self):
        return float(self.world.heat)

    def init_constants(self):
        """Init-time constants exactly as the reference derives them (physics.py:314,322,624,632-635,358-359)."""
        out = []
        for o in self.world.objs:
            offx = [float(np.linalg.norm(p.radius) * np.cos(p.rotpos)) for p in o.points]
            offy = [float(np.linalg.norm(p.radius) * np.sin(p.rotpos)) for p in o.points]
            out.append(dict(com=[float(o.com.pos[0]), float(o.com.pos[1])],
                            vel=[float(o.com.vel[0]), float(o.com.vel[1])],
                            radius=float(o.radius), mass=float(o.mass), offx=offx, offy=offy))
        fixed = []
        for o in self.world.fixed_objs:
            fixed.append(dict(com=[float(o.com.pos[0]), float(o.com.pos[1])], radius=float(o.radius),
                              pts=[[float(p.pos[0]), float(p.pos[1])] for p in o.points]))
        return out, fixed

    # -- stepping --------------------------------------------------------------------------------
    def update(self):
        """One World.update() (physics.py:76-259). Returns False if the world is (or becomes) capped."""
        if self.capped_at >= 0:
            return False
        self._ncalls = 0
        self._ncalls_all = 0
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                self.world.update()
        except Capped:
            self.capped_at = self.step_index
            self.calls_per_step.append(self._ncalls)
            return False
        self.calls_per_step.append(self._ncalls)
        self.step_index += 1
        return True

    def dump(self):
        return str(self.world)


def perturbed_scene_text(path, deltas):
    """Rewrite a .in scene with per-free-body offsets deltas[b] = (dx, dy, dvx, dvy): dx,dy are added to every
    vertex of body b (hence to its centroid), dvx,dvy to its `velocity` (SURVEY 8d perturbation law; the draws
    themselves come from fizz2d_b200.perturbation_deltas).  Floats are written with repr() so that the
    reference parser's float() reproduces x+dx bit for bit."""
    lines = open(path).read().split("\n")
    out = []
    shape = ""
    cur = None
    body = 0
    for line in lines:
        s = line.strip()
        if not s or s[0] == "#" or s.split()[0] == "PARAM":
            out.append(line)
            continue
        if shape == "polygon":
            data = s.split(",")
            if data[0] == "sides":
                cur["left"] = int(data[1])
                out.append(line)
            elif data[0] == "velocity":
                cur["vel_line"] = len(out)
                cur["vel"] = [float(data[1]), float(data[2])]
                out.append(line)
            elif data[0] == "point":
                cur["pts"].append((len(out), float(data[1]), float(data[2])))
                out.append(line)
                cur["left"] -= 1
                if cur["left"] == 0:
                    dx, dy, dvx, dvy = [float(x) for x in deltas[body]]
                    for (li, x, y) in cur["pts"]:
                        out[li] = "point,%r,%r" % (x + dx, y + dy)
                    vx, vy = cur.get("vel", [0.0, 0.0])
                    vline = "velocity,%r,%r" % (vx + dvx, vy + dvy)
                    if "vel_line" in cur:
                        out[cur["vel_line"]] = vline
                    else:
                        out.insert(cur["pts"][0][0], vline)
                    body += 1
                    shape = ""
            else:
                out.append(line)
        elif shape == "fixedpolygon":
            data = s.split(",")
            out.append(line)
            if data[0] == "sides":
                cur = {"left": int(data[1])}
            elif data[0] == "point":
                cur["left"] -= 1
                if cur["left"] == 0:
                    shape = ""
        else:
            shape = s
            cur = {"pts": []}
            out.append(line)
    return "\n".join(out)
