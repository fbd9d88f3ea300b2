"""Drop-in faces of the reference's main.py-level API, backed by the batched CUDA engine.

`World` offers what main.py uses of `physics.World` (physics.py:30-301): `update()`, `str(world)`, `energy()`,
`.heat`, `.state`, `.width`, `.height`, `.log`, `.objs`, `.fixed_objs`; `read_input(fname)` mirrors `main.read_input`
(main.py:23-83).  The reference's object API works too -- `World(width, height, global_accel=...)`, then
`Point(world, pos=, speed=)`, `Polygon(world, density=, points=[Point...], speed=)`, `FixedPolygon(world, points=)`
(physics.py:32, :472, :610, :753), the way `main.read_input` builds a scene: the bodies are collected on the host and the
device batch is built on first use.  A `World` may carry a whole batch (n_worlds > 1, perturbed copies of the scene); the
text outputs then show world `view` (0 by default) so the reference's file formats stay byte-identical for that world.
"""
from __future__ import annotations

import numpy as np

from . import scene as _scene
from .batch import BatchedWorld, GroupSpec

SIMULATION_DIR = './simulations/'  # config.py:1


class Point:
    """physics.Point (physics.py:472-498) as far as scene construction and read-back need it: `.pos`, `.vel`."""

    def __init__(self, world=None, mass=1, pos=(0.0, 0.0), speed=(0.0, 0.0), I=0.0, rotvel=0.0):
        self.name = "point"
        self.world, self.mass = world, mass
        self.pos = np.array(pos, dtype=float)
        self.vel = np.array(speed, dtype=float)


class _Body:
    def __init__(self, world, points, speed, density, fixed):
        self.world, self.density, self.is_fixed = world, density, fixed
        self.name = "fixedpolygon" if fixed else "polygon"
        self.points = list(points)
        self.num_edges = len(self.points)
        self._speed = np.array(speed, dtype=float)
        self._index = None          # position in world.objs / world.fixed_objs once registered

    def _xy(self):
        return np.array([p.pos for p in self.points], dtype=float)

    # read-back: what a caller of the reference would read off the objects after update() --------------------------
    @property
    def pos(self):
        """obj.pos: the snapshot finish_update() writes (physics.py:453-464) = com.pos at a step boundary."""
        if self.is_fixed or self.world is None or self.world._batch is None:
            return self._xy().mean(axis=0)
        return self.world._com()[self._index, 0:2].copy()

    @property
    def vel(self):
        if self.is_fixed or self.world is None or self.world._batch is None:
            return self._speed.copy()
        return self.world._com()[self._index, 2:4].copy()

    @property
    def mass(self):
        if self.is_fixed:
            return np.inf
        self.world._build()
        return float(self.world._batch.groups[0].spec.mass[self._index, self.world.view])


class Polygon(_Body):
    """physics.Polygon(world, density, mass, points, speed, rotation_speed) (physics.py:610).  `rotation_speed` must
    be 0: the reference's rotation machinery is inert from `.in` files and not part of the stepping path."""

    def __init__(self, world=None, density=1, mass=1, points=(), speed=(0.0, 0.0), rotation_speed=0.0):
        if rotation_speed != 0.0:
            raise NotImplementedError("rotation_speed != 0: outside the stepping path (SURVEY 8a: rotation is inert)")
        if density == np.inf:
            raise NotImplementedError("a Polygon of infinite density: use FixedPolygon (main.py:77)")
        super().__init__(world, points, speed, density, fixed=False)
        if world is not None:
            world.init_obj(self)


class FixedPolygon(_Body):
    """physics.FixedPolygon(world, points) (physics.py:752-760)."""

    def __init__(self, world=None, points=(), speed=(0.0, 0.0), rotation_speed=0.0):
        super().__init__(world, points, (0.0, 0.0), np.inf, fixed=True)
        if world is not None:
            world.init_fixed_obj(self)


class World:
    """physics.World (physics.py:30-301).  Two ways in, as in the reference:
    `World(width, height[, objs, global_accel, time_disc])` + Polygon/FixedPolygon objects registering themselves, or
    `World(scene, ...)` / `read_input(fname)` from a parsed `.in` file."""

    def __init__(self, width, height=None, objs=(), global_accel=(0.0, 0.0), time_disc=_scene.TIME_DISC, gamma=(),
                 n_worlds=1, perturb=False, seed=0, view=0, energylog=0, gravity=None, max_calls=4096, device=0,
                 params=None):
        if len(gamma) > 0:
            raise NotImplementedError("damping (gamma): unreachable from .in files, not part of the stepping path")
        if time_disc != _scene.TIME_DISC:
            raise NotImplementedError("time_disc other than the reference's 0.05")
        self.objs, self.fixed_objs = [], []
        self.time_disc = time_disc
        self.state = 0                                   # physics.py:39
        self.view = view
        self._n_worlds, self._perturb, self._seed = n_worlds, perturb, seed
        self._max_calls, self._device = max_calls, device
        self._batch = None
        self._com_cache = None
        self.params = dict(params or {})                 # PARAM lines / setattr(physics, NAME, value) (main.py:39-41)
        self.global_accel = np.array(gravity if gravity is not None else global_accel, dtype=float)
        if isinstance(width, _scene.Scene):              # from a parsed .in file
            sc = width
            self.scene = sc
            self.width, self.height = sc.width, sc.height
            self.params = dict(sc.params)
            for body in sc.free:
                Polygon(self, density=body.density, points=[Point(self, pos=p) for p in body.points], speed=body.velocity)
            for pts in sc.fixed:
                FixedPolygon(self, points=[Point(self, pos=p) for p in pts])
        else:
            self.scene = None
            self.width, self.height = width, height
            for o in objs:
                self.init_obj(o)
        self.energylog = energylog
        self.log = open(SIMULATION_DIR + 'energy.txt', 'a+') if energylog else None   # physics.py:40-41

    # -- construction (physics.py:56-64) -----------------------------------------------------------------------------
    def init_obj(self, obj):
        if self._batch is not None:
            raise RuntimeError("the world is resident on the GPU already: add bodies before the first update()")
        obj.world, obj._index = self, len(self.objs)
        self.objs.append(obj)

    def init_fixed_obj(self, obj):
        if self._batch is not None:
            raise RuntimeError("the world is resident on the GPU already: add bodies before the first update()")
        obj.world, obj._index = self, len(self.fixed_objs)
        self.fixed_objs.append(obj)

    def _build(self):
        if self._batch is None:
            sc = _scene.Scene(self.width, self.height, dict(self.params),
                              [_scene.FreePolygon(o._xy(), float(o.density), o._speed.copy()) for o in self.objs],
                              [o._xy() for o in self.fixed_objs], getattr(self.scene, "name", ""))
            deltas = _scene.perturbation_deltas(self._seed, self._n_worlds, sc.n_free) if self._perturb else None
            spec = GroupSpec.from_scene(sc, self._n_worlds, deltas, gravity=tuple(self.global_accel),
                                        max_calls=self._max_calls)
            self._batch = BatchedWorld([spec], device=self._device)
        return self._batch

    @property
    def batch(self):
        return self._build()

    def _com(self):
        if self._com_cache is None:
            self._com_cache = self.batch.com()[self.view]
        return self._com_cache

    # -- the reference's methods -------------------------------------------------------------------------------------
    def energy(self):
        """World.energy() of the viewed world (physics.py:66-74)."""
        return self.batch.energy()[self.view]

    @property
    def heat(self):
        return float(self.batch.heat()[self.view])

    def update(self, n=1):
        """World.update() (physics.py:76-259) for every world of the batch, n times."""
        self._com_cache = None
        if self.log is not None:
            # one row BEFORE every update() (physics.py:81-82): n rows for n steps
            rows = self.batch.step_logged(n)[:, self.view, :]
            for r in rows:
                self.log.write(','.join([str(x) for x in r]) + '\n')
        else:
            self.batch.step(n)
        for _ in range(n):
            self.state += self.time_disc                                              # physics.py:259

    def check_collisions(self):
        raise NotImplementedError("check_collisions() runs inside update() on the device; use BatchedWorld.enable_trace() "
                                  "to record the tuples it returns")

    def __str__(self):
        return self.batch.dump_world(self.view)                                       # physics.py:291-301


def read_input(fname, energylog=0, **kw):
    """main.read_input (main.py:23-83): parse a .in scene and build the world(s)."""
    return World(_scene.load_in(fname), energylog=energylog, **kw)
