"""Scene inputs.  The reference's `simulations/*.in` files are DATA the path is defined on ("correctness is checked
against the reference run on the repo's own .in scenes"); they ship as package data (`data/scenes.json`, verbatim
texts keyed by scene name) and are materialised as `simulations/<SCENE>.in` on demand, so the CLI, tests, smoke() and
bench.py run where /root/reference does not exist.  Callers with their own scenes pass a path or text instead
(`BatchedWorld.from_in`, `.from_text`)."""
from __future__ import annotations

import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DATA = os.path.join(HERE, "data", "scenes.json")
SIM_DIR = os.path.join(ROOT, "simulations")
NAMES = ["BOX_1SQUARE", "BOX_1SQUAREFRICTION", "RAMPWORLD_2TRIANGLES_5SQUARES", "2CHAMBERS_3BALLS",
         "STACKEDCHAMBERS_1SQUARE", "PENTAGONVOID_1SQUARE", "CORRIDOR_2SQUARES", "PLANE_2SQUARES",
         "RAMP_1SQUARE_1RECTANGLE"]
_texts = None


def scene_text(name: str) -> str:
    global _texts
    if _texts is None:
        with open(DATA) as f:
            _texts = json.load(f)["scenes"]
    return _texts[name]


def scene_path(name: str) -> str:
    """Path of simulations/<name>.in, written from the package data if it is not there yet."""
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
