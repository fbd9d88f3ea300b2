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
