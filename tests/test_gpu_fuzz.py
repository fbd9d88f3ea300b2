"""GPU: randomly generated scenes against the CPU oracle -- topologies the shipped scenes do not reach: up to 8 free
polygons with 3..12 vertices (16/32-lane SAT slots, free-free contacts between different shapes), no fixed bodies,
PARAM overrides, gravity, tiny watchdog budgets.  Both execution modes, bit-exact."""
import os

import numpy as np
import pytest

import fizz2d_b200 as fz
from fizz2d_b200.batch import GroupSpec
from helpers import bits, oracle_from_spec, orc

pytestmark = pytest.mark.gpu


def _convex(rng, cx, cy, r, n):
    while True:
        a = np.sort(rng.uniform(0, 2 * np.pi, n))
        gaps = np.diff(np.concatenate([a, [a[0] + 2 * np.pi]]))
        if gaps.min() > 0.25 and gaps.max() < np.pi * 0.95:
            return [(cx + r * np.cos(t), cy + r * np.sin(t)) for t in a]


def random_scene(seed):
    rng = np.random.default_rng(seed)
    W, H = 600, 400
    F = int(rng.integers(1, 9))
    X = int(rng.integers(0, 7))
    out = "%d,%d\n" % (W, H)
    walls = [[(0, 0), (0, H), (20, H), (20, 0)], [(W - 20, 0), (W - 20, H), (W, H), (W, 0)],
             [(21, 0), (21, 20), (W - 21, 20), (W - 21, 0)], [(21, H - 20), (21, H), (W - 21, H), (W - 21, H - 20)]]
    fixed = walls[:min(X, 4)]
    for _ in range(max(0, X - 4)):
        fixed.append(_convex(rng, rng.uniform(150, 450), rng.uniform(120, 280), rng.uniform(20, 45), int(rng.integers(3, 9))))
    for pts in fixed:
        out += "fixedpolygon\nsides,%d\n" % len(pts) + "".join("point,%r,%r\n" % (float(x), float(y)) for x, y in pts)
    for _ in range(F):
        n = int(rng.integers(3, 13))
        pts = _convex(rng, rng.uniform(60, 540), rng.uniform(60, 340), rng.uniform(8, 22), n)
        out += "polygon\ndensity,%r\nsides,%d\nvelocity,%r,%r\n" % (float(rng.uniform(0.5, 4)), n,
                                                                  float(rng.uniform(-150, 150)), float(rng.uniform(-150, 150)))
        out += "".join("point,%r,%r\n" % (float(x), float(y)) for x, y in pts)
    if rng.random() < 0.7:
        out += "PARAM ELASTICITY %r\n" % float(rng.choice([0.5, 0.8, 0.9, 1.0]))
    if rng.random() < 0.3:
        out += "PARAM COLLISION_TOL %r\n" % float(rng.choice([0.05, 0.5]))
    if rng.random() < 0.3:
        out += "PARAM EPSILON %r\n" % float(rng.choice([1e-9, 1e-6]))
    if rng.random() < 0.2:
        out += "PARAM VIBRATE_TOL %r\n" % float(rng.choice([1e-3, 1e-8]))
    if rng.random() < 0.2:
        out += "PARAM TIME_TOL %r\n" % float(rng.choice([1e-6, 1e-12]))
    gravity = (0.0, 0.0) if rng.random() < 0.6 else (float(rng.uniform(-20, 20)), float(rng.uniform(-300, 300)))
    return out, gravity


@pytest.mark.parametrize("mode", [("hybrid", 37), ("hybrid3", 16), ("thread", 1)], ids=["hybrid", "hybrid3", "thread"])
@pytest.mark.parametrize("seed", list(range(24)))
def test_random_scene_vs_oracle(seed, mode):
    text, gravity = random_scene(1000 + seed)
    sc = fz.parse_in(text)
    n, steps, mc = 48, 160, [32, 64, 256][seed % 3]
    deltas = fz.perturbation_deltas(seed, n, sc.n_free, dpos=3.0, dvel=40.0)
    spec = GroupSpec.from_scene(sc, n, deltas, gravity=gravity, max_calls=mc)
    bw = fz.BatchedWorld([spec])
    bw.set_tuning(*mode)
    bw.step(steps // 2)
    bw.step(steps - steps // 2)
    com, verts, heat, st, sd, calls, en = bw.com(), bw.vertices(), bw.heat(), bw.status(), bw.steps_done(), bw.calls(), bw.energy()
    worlds = [oracle_from_spec(spec, w, gravity=gravity) for w in range(n)]
    orc.run_batch(worlds, steps, os.cpu_count() or 1)
    for w, ow in enumerate(worlds):
        tag = (seed, w, sc.topology())
        assert ow.steps_done() == sd[w] and ow.status() == st[w], tag
        assert np.array_equal(bits(ow.com()), bits(com[w])), tag
        assert bits(ow.heat()) == bits(heat[w]) and ow.total_calls() == calls[w], tag
        if st[w] == 0:
            assert np.array_equal(bits(ow.verts()), bits(verts[w])), tag
            assert np.array_equal(bits(ow.energy()), bits(en[w])), tag


def test_zero_steps_and_tiny_batches():
    bw = fz.BatchedWorld.from_in(fz.scenes.scene_path("BOX_1SQUARE"), n_worlds=1)
    bw.step(0)
    assert bw.steps_done()[0] == 0
    for n in (1, 31, 33, 129):
        b2 = fz.BatchedWorld.from_in(fz.scenes.scene_path("CORRIDOR_2SQUARES"), n_worlds=n, perturb=True, seed=3)
        b2.step(50)
        assert (b2.steps_done()[b2.status() == 0] == 50).all() and len(b2.heat()) == n
