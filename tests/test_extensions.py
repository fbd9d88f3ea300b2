"""The default-off physics extensions: Coulomb friction and restitution between free bodies.
The reference has neither, so there is nothing to be bit-compatible with: oracle/fizz_oracle.c is their definition (CPU
tests below check its physics), and the CUDA kernels are checked against it bit for bit (GPU tests)."""
import os

import numpy as np
import pytest

from conftest import golden, scene_file
from helpers import bits, oracle_from_spec, orc


def _oracle(name, extra="", **kw):
    text = str(golden(name)["scene_text"]) + extra
    return orc.OracleWorld.from_text(text, **kw)


@pytest.mark.parametrize("name", ["BOX_1SQUARE", "BOX_1SQUAREFRICTION", "2CHAMBERS_3BALLS", "CORRIDOR_2SQUARES", "PLANE_2SQUARES"])
def test_extensions_off_is_the_reference(name):
    """friction 0 / ff_restitution < 0 written out explicitly: the reference's bits, every step."""
    g = golden(name)
    ow = _oracle(name, "\nPARAM FRICTION 0.0\nPARAM FF_RESTITUTION -1.0\n")
    for t in range(int(g["steps_done"])):
        ow.update()
        assert np.array_equal(bits(ow.com()), bits(g["com"][t])) and bits(ow.heat()) == bits(g["heat"][t]), (name, t)


@pytest.mark.parametrize("name,extra", [("RAMPWORLD_2TRIANGLES_5SQUARES", "\nPARAM FRICTION 0.3\n"),
                                        ("CORRIDOR_2SQUARES", "\nPARAM FRICTION 0.5\nPARAM FF_RESTITUTION 0.6\n"),
                                        ("CORRIDOR_2SQUARES", "\nPARAM FF_RESTITUTION 0.0\n"),
                                        ("2CHAMBERS_3BALLS", "\nPARAM FRICTION 1.5\nPARAM FF_RESTITUTION 0.9\nPARAM ELASTICITY 0.7\n")])
def test_energy_is_booked_never_created(name, extra):
    """K + P + H is constant to rounding with the extensions on (what friction and restitution take from K goes to H),
    H never decreases, and the extensions do change the trajectory."""
    ow, ref = _oracle(name, extra, max_calls=512), _oracle(name, max_calls=512)
    e0 = ow.energy()
    h_prev, differs = 0.0, False
    for t in range(400):
        if not ow.update():
            break
        ref.update()
        e = ow.energy()
        assert abs(e[0] - e0[0]) <= 1e-9 * max(1.0, abs(e0[0])), (name, t, e, e0)
        assert e[3] >= h_prev - 1e-9 * max(1.0, abs(e0[0])), (name, t)
        h_prev = e[3]
        differs = differs or not np.array_equal(bits(ow.com()), bits(ref.com()))
    assert differs, "the extension had no effect on %s" % name


def test_coulomb_clamp_and_stick():
    """One contact against a fixed floor-like wall: the tangential velocity drops by mu (1 + e) |v.n| while sliding and
    to exactly zero when that would overshoot (stick)."""
    text = "400,300\nfixedpolygon\nsides,4\npoint,0,0\npoint,0,20\npoint,400,20\npoint,400,0\n" \
           "polygon\ndensity,1\nsides,4\nvelocity,%r,-100.0\npoint,100,40\npoint,120,40\npoint,120,60\npoint,100,60\n" \
           "PARAM ELASTICITY 0.5\nPARAM FRICTION %r\n"
    for vx, mu, expect in ((50.0, 0.2, 50.0 - 0.2 * 1.5 * 100.0), (10.0, 0.2, 0.0), (-50.0, 0.1, -50.0 + 0.1 * 1.5 * 100.0)):
        ow = orc.OracleWorld.from_text(text % (vx, mu), trace_cap=4096)
        for t in range(40):
            ow.update()
            if ow.contacts():
                break
        c = ow.contacts()
        assert c and abs(abs(c[0][5]) - 1.0) < 1e-12, c[:1]      # the floor's normal is vertical
        # the first resolve pass reflects v_y and applies the friction impulse once (later passes see v.n of the other
        # sign and act again: compare right after the step in which the first contact was resolved only if one pass)
        vx_after = ow.com()[0, 2]
        assert abs(vx_after) <= abs(vx) + 1e-9
        if expect == 0.0:
            assert abs(vx_after) < 1e-9
        else:
            assert abs(vx_after) <= abs(expect) + 1e-6


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("mode", [("hybrid", 37), ("hybrid3", 16), ("thread", 1)], ids=["hybrid", "hybrid3", "thread"])
@pytest.mark.parametrize("seed", list(range(16)))
def test_gpu_extensions_fuzz_vs_oracle(seed, mode):
    """Random scenes (the family of tests/test_gpu_fuzz.py) with random friction / free-free restitution: the CUDA engine
    equals the extension oracle bit for bit -- state, heat, call counts, energies."""
    import fizz2d_b200 as fz
    from fizz2d_b200.batch import GroupSpec
    from test_gpu_fuzz import random_scene
    rng = np.random.default_rng(77 + seed)
    text, gravity = random_scene(3000 + seed)
    mu = float(rng.choice([0.0, 0.1, 0.5, 2.0]))
    ffr = float(rng.choice([-1.0, 0.0, 0.4, 1.0]))
    if mu == 0.0 and ffr < 0.0:
        mu = 0.3
    sc = fz.parse_in(text)
    n, steps, mc = 48, 160, [32, 64, 256][seed % 3]
    deltas = fz.perturbation_deltas(seed, n, sc.n_free, dpos=3.0, dvel=40.0)
    spec = GroupSpec.from_scene(sc, n, deltas, gravity=gravity, max_calls=mc, friction=mu, ff_restitution=ffr)
    bw = fz.BatchedWorld([spec])
    bw.set_tuning(*mode)
    e0 = bw.energy()
    bw.step(steps // 2)
    bw.step(steps - steps // 2)
    com, verts, heat, st, sd, calls, en = bw.com(), bw.vertices(), bw.heat(), bw.status(), bw.steps_done(), bw.calls(), bw.energy()
    worlds = [oracle_from_spec(spec, w, gravity=gravity) for w in range(n)]
    orc.run_batch(worlds, steps, os.cpu_count() or 1)
    for w, ow in enumerate(worlds):
        tag = (seed, w, mu, ffr, sc.topology())
        assert ow.steps_done() == sd[w] and ow.status() == st[w], tag
        assert np.array_equal(bits(ow.com()), bits(com[w])), tag
        assert bits(ow.heat()) == bits(heat[w]) and ow.total_calls() == calls[w], tag
        if st[w] == 0:
            assert np.array_equal(bits(ow.verts()), bits(verts[w])), tag
            assert np.array_equal(bits(ow.energy()), bits(en[w])), tag
            if gravity == (0.0, 0.0):
                assert abs(en[w][0] - e0[w][0]) <= 1e-9 * max(1.0, abs(e0[w][0])), tag     # never created, never lost


@pytest.mark.gpu
def test_gpu_extensions_from_param_lines_and_trapped_worlds():
    """`PARAM FRICTION` / `PARAM FF_RESTITUTION` lines switch the extensions on from a .in file; with friction on the
    register-resident single-contact path is not taken (it implements the reference's response only) and trapped worlds
    still equal the oracle."""
    import fizz2d_b200 as fz
    from fizz2d_b200.batch import GroupSpec
    text = open(scene_file("RAMPWORLD_2TRIANGLES_5SQUARES")).read() + "\nPARAM FRICTION 0.25\nPARAM FF_RESTITUTION 0.5\n"
    sc = fz.parse_in(text)
    heavy = [63739, 12860, 5657, 58248, 1, 2, 3, 4]
    d = np.ascontiguousarray(fz.perturbation_deltas(20260003, 65536, sc.n_free)[heavy])
    spec = GroupSpec.from_scene(sc, len(heavy), d, max_calls=1024)
    assert spec.params.friction_mu == 0.25 and spec.params.ff_restitution == 0.5
    for mode in (("hybrid", 64), ("thread", 1)):
        bw = fz.BatchedWorld([spec])
        bw.set_tuning(*mode)
        bw.step(300)
        com, heat, st, sd, calls = bw.com(), bw.heat(), bw.status(), bw.steps_done(), bw.calls()
        worlds = [oracle_from_spec(spec, w) for w in range(len(heavy))]
        orc.run_batch(worlds, 300, os.cpu_count() or 1)
        for w, ow in enumerate(worlds):
            assert ow.steps_done() == sd[w] and ow.status() == st[w] and ow.total_calls() == calls[w], (mode, heavy[w])
            assert np.array_equal(bits(ow.com()), bits(com[w])) and bits(ow.heat()) == bits(heat[w]), (mode, heavy[w])
        assert (heat > 0).any()
