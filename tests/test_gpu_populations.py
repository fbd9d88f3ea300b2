"""GPU: mixed batches (BASELINE config 4) and GA populations (config 5) -- several groups, one resident scene
descriptor each, world order preserved -- against reference fixtures and the CPU oracle."""
import os

import numpy as np
import pytest

import fizz2d_b200 as fz
from conftest import golden, scene_file
from helpers import bits, contact_rows, oracle_from_spec, orc

pytestmark = pytest.mark.gpu
MODES = [("hybrid", 64), ("hybrid3", 32), ("thread", 1)]


@pytest.mark.parametrize("mode", MODES, ids=["hybrid", "hybrid3", "thread"])
def test_mixed_batch_matches_reference(mode):
    """Worlds 0..12 of config 4 (even: 2CHAMBERS_3BALLS, odd: STACKEDCHAMBERS_1SQUARE, seed 20260004)."""
    g = golden("perturbed_config4_mixed")
    n = int(g["worlds"].max()) + 1
    bw = fz.BatchedWorld.from_mixed([scene_file("2CHAMBERS_3BALLS"), scene_file("STACKEDCHAMBERS_1SQUARE")], n,
                                    seed=20260004, max_calls=int(g["max_calls"]))
    bw.set_tuning(*mode)
    assert len(bw.groups) == 2
    bw.step(int(g["nsteps"]))
    com, heat, st, sd, en = bw.com(), bw.heat(), bw.status(), bw.steps_done(), bw.energy()
    for wi in g["worlds"]:
        gi, k = bw.locate(int(wi))
        pre = "w%d_" % wi
        assert np.array_equal(bits(com[gi][k]), bits(g[pre + "final_com"])), wi
        assert bits(heat[wi]) == bits(g[pre + "final_heat"])
        assert sd[wi] == int(g[pre + "steps_done"])
        if st[wi] == 0:
            assert np.array_equal(bits(en[wi]), bits(g[pre + "final_energy"]))


@pytest.mark.parametrize("mode", MODES, ids=["hybrid", "hybrid3", "thread"])
def test_ga_population_matches_reference(mode):
    g = golden("ga_config5_pentagonvoid")
    n = int(g["worlds"].max()) + 1
    bw = fz.BatchedWorld.from_ga(scene_file("PENTAGONVOID_1SQUARE"), n, seed=int(g["seed"]), max_calls=int(g["max_calls"]))
    bw.set_tuning(*mode)
    bw.enable_trace(1 << 18)
    assert len(bw.groups) >= 5          # one group per vertex count
    done = 0
    for c in [int(x) for x in g["checkpoints"]]:
        bw.step(c - done)
        done = c
        com = bw.com()
        for wi in g["worlds"]:
            key = "w%d_com_at_%d" % (wi, c)
            if key in g.files:
                gi, k = bw.locate(int(wi))
                assert np.array_equal(bits(com[gi][k]), bits(g[key])), (wi, c)
    com, verts, heat, st, sd = bw.com(), bw.vertices(), bw.heat(), bw.status(), bw.steps_done()
    cons = bw.contacts()
    for wi in g["worlds"]:
        gi, k = bw.locate(int(wi))
        pre = "w%d_" % wi
        assert sd[wi] == int(g[pre + "steps_done"]) and (st[wi] == fz.ST_CAPPED) == (int(g[pre + "capped_at"]) >= 0)
        assert np.array_equal(bits(com[gi][k]), bits(g[pre + "final_com"])), wi
        assert np.array_equal(bits(verts[gi][k]), bits(g[pre + "final_verts"])), wi
        assert bits(heat[wi]) == bits(g[pre + "final_heat"])
        got = [c[1:] for c in cons if c[0] == wi]
        want = contact_rows(g[pre + "contacts_idx"], g[pre + "contacts_val"])
        assert [c[:4] for c in got] == [c[:4] for c in want], wi
        assert np.array_equal(bits([c[4:] for c in got]), bits([c[4:] for c in want])), wi


def test_ga_population_vs_oracle_and_dump():
    """4096 GA candidates x 400 steps against the oracle; world dumps keep the reference format."""
    n, steps = 4096, 400
    bw = fz.BatchedWorld.from_ga(scene_file("PENTAGONVOID_1SQUARE"), n, seed=77, max_calls=256)
    bw.step(steps)
    com, heat, st, sd, calls = bw.com(), bw.heat(), bw.status(), bw.steps_done(), bw.calls()
    worlds, where = [], []
    for gi, grp in enumerate(bw.groups):
        for k in range(grp.spec.n_worlds):
            worlds.append(oracle_from_spec(grp.spec, k))
            where.append((gi, k, int(grp.world_ids[k])))
    orc.run_batch(worlds, steps, os.cpu_count() or 1)
    for ow, (gi, k, wi) in zip(worlds, where):
        assert ow.steps_done() == sd[wi] and ow.status() == st[wi], wi
        assert np.array_equal(bits(ow.com()), bits(com[gi][k])), wi
        assert bits(ow.heat()) == bits(heat[wi]) and ow.total_calls() == calls[wi], wi
    text = bw.dump_world(5)
    assert text.startswith("200,400\npolygon\nsides,") and text.count("polygon") == 3
    # fitness := dissipated energy is non-negative and zero for worlds that never touched anything
    assert (heat >= 0).all()


def test_ga_generation_loop_improves_fitness():
    """Two GA generations on the GPU: selection keeps the best candidates bit for bit, offspring are valid shapes."""
    path = scene_file("PENTAGONVOID_1SQUARE")
    pop = fz.ga.random_convex_polygons(5, 1024)
    best = []
    for gen in range(2):
        bw = fz.BatchedWorld.from_ga(path, 1024, population=pop, max_calls=128)
        bw.step(300)
        fit = np.where(bw.status() == 0, bw.calls().astype(float), -1.0)   # any deterministic score will do here
        best.append(fit.max())
        new = fz.ga.next_generation(pop, fit, 5, gen)
        top = int(np.argmax(fit))
        assert np.array_equal(bits(new[1][0]), bits(pop[1][top]))          # elitism: the best survives unchanged
        assert all(3 <= len(p) <= 12 for p in new[1])
        pop = new
        bw.close()
    assert best[1] >= best[0]
