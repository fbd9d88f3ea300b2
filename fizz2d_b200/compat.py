"""Drop-in faces of the reference's main.py-level API, backed by the batched CUDA engine.

`World` offers what main.py uses of `physics.World` (physics.py:30-301): `update()`, `str(world)`, `energy()`,
`.heat`, `.state`, `.width`, `.height`, `.log`; `read_input(fname)` mirrors `main.read_input` (main.py:23-83).
A `World` may carry a whole batch (n_worlds > 1, perturbed copies of the scene); the text outputs then show
world `view` (0 by default) so the reference's file formats stay byte-identical for that world.
"""
from __future__ import annotations

import os

import numpy as np

from . import scene as _scene
from .batch import BatchedWorld, GroupSpec

SIMULATION_DIR = './simulations/'  # config.py:1


class World:
    def __init__(self, scene: _scene.Scene, n_worlds=1, perturb=False, seed=0, view=0, energylog=0, gravity=(0.0, 0.0),
                 max_calls=4096, device=0):
        self.scene = scene
        self.width, self.height = scene.width, scene.height
        self.time_disc = _scene.TIME_DISC
        self.state = 0                                   # physics.py:39
        self.view = view
        deltas = _scene.perturbation_deltas(seed, n_worlds, scene.n_free) if perturb else None
        self.batch = BatchedWorld([GroupSpec.from_scene(scene, n_worlds, deltas, gravity=gravity, max_calls=max_calls)],
                                  device=device)
        self.energylog = energylog
        self.log = open(SIMULATION_DIR + 'energy.txt', 'a+') if energylog else None   # physics.py:40-41

    def energy(self):
        """World.energy() of the viewed world (physics.py:66-74)."""
        return self.batch.energy()[self.view]

    @property
    def heat(self):
        return float(self.batch.heat()[self.view])

    def update(self, n=1):
        """World.update() (physics.py:76-259) for every world of the batch."""
        if self.log is not None:
            # one row BEFORE every update() (physics.py:81-82): n rows for n steps
            rows = self.batch.step_logged(n)[:, self.view, :]
            for r in rows:
                self.log.write(','.join([str(x) for x in r]) + '\n')
        else:
            self.batch.step(n)
        for _ in range(n):
            self.state += self.time_disc                                              # physics.py:259

    def __str__(self):
        return self.batch.dump_world(self.view)                                       # physics.py:291-301


def read_input(fname, energylog=0, **kw):
    """main.read_input (main.py:23-83): parse a .in scene and build the world(s)."""
    return World(_scene.load_in(fname), energylog=energylog, **kw)
