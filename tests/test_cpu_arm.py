"""CPU: the product-free CPU arms of bench.py (oracle/cpu_arm.py) build the same worlds as the product loader, and the
Python reference shipped under oracle/_ref/pyref/ is the reference."""
import filecmp
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, scene_file
from helpers import bits, oracle_from_spec

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cpu_arm  # noqa: E402


def test_perturbation_law_restated():
    import fizz2d_b200 as fz
    a = fz.perturbation_deltas(20260003, 300, 7)
    b = cpu_arm.perturbation_deltas(20260003, 300, 7)
    assert np.array_equal(bits(a), bits(b))


def test_port_worlds_equal_product_loader():
    """Same worlds, bit for bit, from oracle.py's own parser + init math and from the product loader."""
    import fizz2d_b200 as fz
    from fizz2d_b200.batch import GroupSpec
    sc = fz.load_in(scene_file("RAMPWORLD_2TRIANGLES_5SQUARES"))
    ids = [0, 1, 7, 130, 255]
    d = fz.perturbation_deltas(20260003, 256, sc.n_free)[ids]
    spec = GroupSpec.from_scene(sc, len(ids), np.ascontiguousarray(d))
    mine = cpu_arm.port_worlds("RAMPWORLD_2TRIANGLES_5SQUARES", 20260003, 256, ids, 4096, 2)
    for k, ow in enumerate(mine):
        ref = oracle_from_spec(spec, k)
        ow.run(150)
        ref.run(150)
        assert np.array_equal(bits(ow.com()), bits(ref.com())) and ow.total_calls() == ref.total_calls()


def test_shipped_python_reference_is_the_reference():
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        pytest.skip("no /root/reference here")
    shipped = os.path.join(ROOT, "oracle", "_ref", "pyref")
    for f in ("physics.py", "main.py", "config.py", "plot.py"):
        assert filecmp.cmp(os.path.join(ref, f), os.path.join(shipped, f), shallow=False), f


def test_pyref_pool_counts_steps():
    if not cpu_arm.pyref_available():
        pytest.skip("reference sources not on this machine")
    pool = cpu_arm.PyRefPool("BOX_1SQUAREFRICTION", 20260002, 64, [0, 1, 2, 3], 4096, 2)
    r = pool.run(40)
    pool.close()
    assert r["world_steps"] == 160 and r["n"] == 4 and r["seconds"] > 0
