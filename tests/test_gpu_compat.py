"""GPU: the drop-in CLI driver writes the reference's files byte for byte (SURVEY 8c anchors are sha256 over the
concatenated plane_%03d.txt bodies and over the energy.txt rows)."""
import hashlib
import json
import os

import pytest

from conftest import GOLDEN, scene_file

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["BOX_1SQUARE", "RAMPWORLD_2TRIANGLES_5SQUARES"])
def test_cli_outputs_match_reference_bytes(name, tmp_path, monkeypatch):
    from fizz2d_b200 import main as fmain
    with open(os.path.join(GOLDEN, "anchors.json")) as f:
        anchors = json.load(f)[name]
    monkeypatch.chdir(tmp_path)
    os.mkdir("simulations")
    fmain.main(["main.py", scene_file(name), "1000", "0", "0", "1"])
    hp = hashlib.sha256()
    for t in range(1000):
        with open("simulations/plane_%03d.txt" % t, "rb") as f:
            hp.update(f.read())
    assert hp.hexdigest() == anchors["planes"]
    with open("simulations/energy.txt", "rb") as f:
        assert hashlib.sha256(f.read()).hexdigest() == anchors["energy"]


def test_cli_batch_view(tmp_path, monkeypatch):
    """--worlds N runs perturbed copies in lockstep; world 0 (the view) is still the reference run."""
    from fizz2d_b200 import main as fmain
    with open(os.path.join(GOLDEN, "anchors.json")) as f:
        anchors = json.load(f)["BOX_1SQUAREFRICTION"]
    monkeypatch.chdir(tmp_path)
    os.mkdir("simulations")
    plane = fmain.main(["main.py", scene_file("BOX_1SQUAREFRICTION"), "1000", "--worlds", "512", "--seed", "20260002"])
    hp = hashlib.sha256()
    for t in range(1000):
        with open("simulations/plane_%03d.txt" % t, "rb") as f:
            hp.update(f.read())
    assert hp.hexdigest() == anchors["planes"]
    assert plane.batch.n_worlds == 512 and abs(plane.state - 50.0) < 1e-6


def test_object_api_builds_the_same_world():
    """The reference's object API (physics.World / Point / Polygon / FixedPolygon, the calls main.read_input makes,
    main.py:28,68,75,77): a world assembled that way steps exactly like the one read from the .in file."""
    import numpy as np
    from fizz2d_b200 import compat, parse_in
    from conftest import golden
    from helpers import bits
    g = golden("2CHAMBERS_3BALLS")
    sc = parse_in(str(g["scene_text"]))
    w = compat.World(sc.width, sc.height, params=sc.params)
    for body in sc.free:
        pts = [compat.Point(w, pos=p, speed=body.velocity) for p in body.points]
        compat.Polygon(w, points=pts, density=body.density, speed=body.velocity)
    for fx in sc.fixed:
        compat.FixedPolygon(w, [compat.Point(w, pos=p) for p in fx])
    assert len(w.objs) == 3 and len(w.fixed_objs) == 5 and w.objs[0].name == "polygon" and w.fixed_objs[0].is_fixed
    for t in range(60):
        assert np.array_equal(bits(w.energy()), bits(g["energy_before"][t]))
        w.update()
    assert np.array_equal(bits([o.pos for o in w.objs]), bits(g["com"][59][:, 0:2]))
    assert np.array_equal(bits([o.vel for o in w.objs]), bits(g["com"][59][:, 2:4]))
    assert w.objs[0].mass == 2500.0 and w.fixed_objs[0].mass == np.inf
    w2 = compat.World(500, 300, global_accel=np.array([0.0, -300.0]))       # gravity through the constructor (physics.py:32)
    compat.Polygon(w2, density=0.1, points=[compat.Point(w2, pos=p) for p in ((100., 100.), (150., 100.), (150., 150.), (100., 150.))])
    w2.update(3)
    assert w2.objs[0].vel[1] < 0 and abs(w2.state - 0.15) < 1e-12


def test_simulate_sh_drop_in(tmp_path):
    """simulate.sh with the reference's arguments (INPUT NSTEPS VERBOSITY PNG ENERGYLOG): the text outputs are the
    reference's bytes, the frames come from the GPU rasteriser (no ./draw binary here) and equal the draw.c restatement."""
    import shutil
    import subprocess
    import sys
    import numpy as np
    from conftest import ROOT
    from fizz2d_b200 import raster
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import raster_oracle as ro
    with open(os.path.join(GOLDEN, "anchors.json")) as f:
        anchors = json.load(f)["BOX_1SQUARE"]
    shutil.copy(os.path.join(ROOT, "simulate.sh"), tmp_path / "simulate.sh")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    subprocess.check_call(["bash", str(tmp_path / "simulate.sh"), scene_file("BOX_1SQUARE"), "1000", "0", "1", "1"],
                          env=env, stdout=subprocess.DEVNULL)
    hp = hashlib.sha256()
    for t in range(1000):
        with open(tmp_path / "simulations" / ("plane_%03d.txt" % t), "rb") as f:
            hp.update(f.read())
    assert hp.hexdigest() == anchors["planes"]
    with open(tmp_path / "simulations" / "energy.txt", "rb") as f:
        assert hashlib.sha256(f.read()).hexdigest() == anchors["energy"]
    for t in (0, 7, 500, 998):
        with open(tmp_path / "simulations" / ("plane_%03d.png" % t), "rb") as f:
            frame = raster.read_png_gray(f.read())
        with open(tmp_path / "simulations" / ("plane_%03d.txt" % t)) as f:
            assert np.array_equal(frame, ro.frame_from_text(f.read())), t
