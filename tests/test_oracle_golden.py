"""CPU: the oracle (oracle/fizz_oracle.c + oracle/oracle.py) against fixtures produced by the reference itself."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, SCENES, golden
from helpers import bits, contact_rows, oracle_from_golden, orc


@pytest.mark.parametrize("name", SCENES + ["gravity_BOX_1SQUARE", "gravity_PLANE_2SQUARES"])
def test_oracle_matches_reference_trace(name):
    g = golden(name)
    grav = tuple(g["gravity"])
    # (a) own parser + own init-time math
    ow = orc.OracleWorld.from_text(str(g["scene_text"]), gravity=grav, trace_cap=1 << 16)
    # (b) constants taken from the fixture
    ow2 = oracle_from_golden(g)
    n = int(g["steps_done"])
    for t in range(n):
        assert np.array_equal(bits(ow.energy()), bits(g["energy_before"][t])), (name, t)
        assert ow.update() and ow2.update()
        assert np.array_equal(bits(ow.com()), bits(g["com"][t])), (name, t)
        assert np.array_equal(bits(ow2.com()), bits(g["com"][t])), (name, t)
        assert np.array_equal(bits(ow.verts()), bits(g["verts"][t])), (name, t)
        assert bits(ow.heat()) == bits(g["heat"][t]), (name, t)
        assert ow.calls_last_step() == g["calls"][t], (name, t)
    got = ow.contacts()
    want = contact_rows(g["contacts_idx"], g["contacts_val"])
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a[:4] == b[:4]
        assert np.array_equal(bits(a[4:]), bits(b[4:]))


def test_oracle_init_constants_match_reference():
    for name in SCENES:
        g = golden(name)
        sc = orc.parse_scene(str(g["scene_text"]))
        offx, offy = [], []
        for b, body in enumerate(sc["free"]):
            c = orc.body_constants(body["points"], body["density"])
            assert np.array_equal(bits(c["com"]), bits(g["init_com"][b]))
            assert bits(c["radius"]) == bits(g["init_radius"][b])
            assert bits(c["mass"]) == bits(g["init_mass"][b])
            offx += c["offx"]
            offy += c["offy"]
        assert np.array_equal(bits(offx), bits(g["init_offx"]))
        assert np.array_equal(bits(offy), bits(g["init_offy"]))


@pytest.mark.parametrize("name", ["BOX_1SQUARE", "BOX_1SQUAREFRICTION", "RAMPWORLD_2TRIANGLES_5SQUARES",
                                  "2CHAMBERS_3BALLS", "STACKEDCHAMBERS_1SQUARE", "PENTAGONVOID_1SQUARE"])
def test_oracle_1000_step_anchors(name):
    """SURVEY 8c known-answer table: sha256 of the 1000 text dumps, the 1000 energy rows, the contact pairs."""
    with open(os.path.join(GOLDEN, "anchors.json")) as f:
        anchors = json.load(f)[name]
    g = golden(name)
    ow = orc.OracleWorld.from_text(str(g["scene_text"]), trace_cap=1 << 16)
    hp, he = hashlib.sha256(), hashlib.sha256()
    for t in range(1000):
        he.update((','.join(str(x) for x in ow.energy()) + '\n').encode())
        assert ow.update()
        hp.update(ow.dump().encode())
    assert hp.hexdigest() == anchors["planes"]
    assert he.hexdigest() == anchors["energy"]
    hc = hashlib.sha256('\n'.join('%d:%d:%d-%d' % c[:4] for c in ow.contacts()).encode()).hexdigest()
    assert hc == anchors["contacts_relevant"]
    assert ow.total_calls() == anchors["calls_relevant"]
    assert np.array_equal(bits(ow.energy()), bits(anchors["final_energy"]))
    assert np.array_equal(bits(ow.com()), bits(anchors["final_com"]))


@pytest.mark.parametrize("tag", ["config2_boxfriction", "config3_rampworld", "config4_mixed", "critical_rampworld"])
def test_oracle_perturbed_worlds(tag):
    """Perturbed worlds (SURVEY 8d law), including worlds the watchdog caps: final state, checkpoints, per-step
    call counts and the full contact trace equal the reference's."""
    g = golden("perturbed_" + tag)
    mc = int(g["max_calls"])
    for k, wi in enumerate(g["worlds"]):
        scene = str(g["scenes"][k])
        text = str(golden(scene)["scene_text"])
        sc = orc.parse_scene(text)
        d = g["deltas"][k]
        for b, body in enumerate(sc["free"]):
            body["points"] = [[p[0] + d[b, 0], p[1] + d[b, 1]] for p in body["points"]]
            body["vel"] = [body["vel"][0] + d[b, 2], body["vel"][1] + d[b, 3]]
        fr = [orc.body_constants(b["points"], b["density"]) for b in sc["free"]]
        fx = [orc.body_constants(p, None) for p in sc["fixed"]]
        pre = "w%d_" % wi
        assert np.array_equal(bits([c["offx"] for c in fr]), bits(g[pre + "init_offx"].reshape(len(fr), -1))) \
            if len({c["n"] for c in fr}) == 1 else True
        ow = orc.OracleWorld([c["n"] for c in fr], [c["n"] for c in fx], [c["com"] for c in fr] + [c["com"] for c in fx],
                             [b["vel"] for b in sc["free"]], [c["radius"] for c in fr] + [c["radius"] for c in fx],
                             [c["mass"] for c in fr], [x for c in fr for x in c["offx"]], [x for c in fr for x in c["offy"]],
                             [p for pts in sc["fixed"] for p in pts], orc.make_params(sc["params"], sc["height"], mc),
                             width=sc["width"], height=sc["height"], trace_cap=1 << 17)
        calls = []
        for t in range(int(g["nsteps"])):
            ok = ow.update()
            calls.append(ow.calls_last_step())
            if not ok:
                break
            if (pre + "com_at_%d" % (t + 1)) in g.files:
                assert np.array_equal(bits(ow.com()), bits(g[pre + "com_at_%d" % (t + 1)])), (wi, t)
        assert ow.steps_done() == int(g[pre + "steps_done"])
        assert (ow.status() == 1) == (int(g[pre + "capped_at"]) >= 0)
        assert np.array_equal(bits(ow.com()), bits(g[pre + "final_com"])), wi
        assert np.array_equal(bits(ow.verts()), bits(g[pre + "final_verts"])), wi
        assert bits(ow.heat()) == bits(g[pre + "final_heat"])
        assert calls == list(g[pre + "calls"]), wi
        got = ow.contacts()
        want = contact_rows(g[pre + "contacts_idx"], g[pre + "contacts_val"])
        assert [c[:4] for c in got] == [c[:4] for c in want], wi
        assert np.array_equal(bits([c[4:] for c in got]), bits([c[4:] for c in want])), wi


def test_oracle_ga_population():
    """GA worlds: the oracle (own parser + init math on the same text the reference got)
    equals the reference, and the committed population is what the generator produces."""
    import fizz2d_b200 as fz
    g = golden("ga_config5_pentagonvoid")
    sc = fz.parse_in(str(g["scene_text"]))
    nverts, pts, vel = fz.ga.random_convex_polygons(int(g["seed"]), int(g["worlds"].max()) + 1)
    for wi in g["worlds"]:
        pre = "w%d_" % wi
        assert np.array_equal(bits(pts[wi]), bits(g[pre + "points"])) and np.array_equal(bits(vel[wi]), bits(g[pre + "vel"]))
        text = fz.ga.fixed_scene_text(sc) + fz.ga.polygon_in_text(g[pre + "points"], g[pre + "vel"])
        ow = orc.OracleWorld.from_text(text, max_calls=int(g["max_calls"]), trace_cap=1 << 16)
        assert np.array_equal(bits(ow.com()[0, :2]), bits(g[pre + "init_com"][0]))
        calls = []
        for t in range(int(g["nsteps"])):
            ok = ow.update()
            calls.append(ow.calls_last_step())
            if not ok:
                break
            key = pre + "com_at_%d" % (t + 1)
            if key in g.files:
                assert np.array_equal(bits(ow.com()), bits(g[key])), (wi, t)
        assert ow.steps_done() == int(g[pre + "steps_done"])
        assert (ow.status() == 1) == (int(g[pre + "capped_at"]) >= 0)
        assert np.array_equal(bits(ow.com()), bits(g[pre + "final_com"])), wi
        assert np.array_equal(bits(ow.verts()), bits(g[pre + "final_verts"])), wi
        assert bits(ow.heat()) == bits(g[pre + "final_heat"])
        assert calls == list(g[pre + "calls"]), wi
        got, want = ow.contacts(), contact_rows(g[pre + "contacts_idx"], g[pre + "contacts_val"])
        assert [c[:4] for c in got] == [c[:4] for c in want], wi
        assert np.array_equal(bits([c[4:] for c in got]), bits([c[4:] for c in want])), wi


@pytest.mark.parametrize("seed", [1003, 1007, 1012, 1018])
def test_oracle_vs_live_reference_on_fuzz_scenes(seed):
    """Where the unmodified reference can be imported (the build container), the C oracle is also checked against it
    on scenes from the SAME random family the GPU fuzz test uses (tests/test_gpu_fuzz.py) -- polygons with up to 12
    vertices, random PARAMs, gravity -- which the fixtures do not contain: every step, bit for bit."""
    import sys
    from helpers import ROOT
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_harness as rh
    if not rh.available():
        pytest.skip("/root/reference is not here")
    from test_gpu_fuzz import random_scene
    text, gravity = random_scene(seed)
    ref = rh.RefWorld(text=text, max_calls=256, gravity=gravity if gravity != (0.0, 0.0) else None)
    ow = orc.OracleWorld.from_text(text, gravity=gravity, max_calls=256)
    for t in range(45):
        alive = ref.update()
        assert ow.update() == alive, (seed, t)
        if not alive:
            break
        assert np.array_equal(bits(ow.com()), bits(ref.com())), (seed, t)
        assert np.array_equal(bits(ow.verts()), bits(np.concatenate(ref.verts()))), (seed, t)
        assert bits(ow.heat()) == bits(ref.heat()), (seed, t)
