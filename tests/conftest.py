import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENES_DIR = os.path.join(ROOT, "simulations")

SCENES = ["BOX_1SQUARE", "BOX_1SQUAREFRICTION", "RAMPWORLD_2TRIANGLES_5SQUARES", "2CHAMBERS_3BALLS",
          "STACKEDCHAMBERS_1SQUARE", "PENTAGONVOID_1SQUARE", "CORRIDOR_2SQUARES", "PLANE_2SQUARES",
          "RAMP_1SQUARE_1RECTANGLE"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the product .so and the oracle .so exist (both compile without a GPU)."""
    import __graft_entry__ as ge
    ge.build()


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def scene_file(name):
    from fizz2d_b200 import scenes
    return scenes.scene_path(name)
