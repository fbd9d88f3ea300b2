"""Scene inputs.  The reference's `simulations/*.in` files are DATA the path is defined on ("correctness is checked
against the reference run on the repo's own .in scenes"); they travel inside the committed fixtures
(`tests/golden/<SCENE>.npz`, field `scene_text`, written by oracle/make_golden.py from the reference tree) and are
materialised as `simulations/<SCENE>.in` on demand, so tests, smoke() and bench.py run where /root/reference does
not exist."""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
SIM_DIR = os.path.join(ROOT, "simulations")
NAMES = ["BOX_1SQUARE", "BOX_1SQUAREFRICTION", "RAMPWORLD_2TRIANGLES_5SQUARES", "2CHAMBERS_3BALLS",
         "STACKEDCHAMBERS_1SQUARE", "PENTAGONVOID_1SQUARE", "CORRIDOR_2SQUARES", "PLANE_2SQUARES",
         "RAMP_1SQUARE_1RECTANGLE"]


def scene_text(name: str) -> str:
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as g:
        return str(g["scene_text"])


def scene_path(name: str) -> str:
    """Path of simulations/<name>.in, written from the fixture if it is not there yet."""
    path = os.path.join(SIM_DIR, name + ".in")
    if not os.path.isfile(path):
        os.makedirs(SIM_DIR, exist_ok=True)
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "w") as f:
            f.write(scene_text(name))
        os.replace(tmp, path)
    return path


def ensure_scenes():
    return [scene_path(n) for n in NAMES]
