"""CPU: the product's host-side loader (fizz2d_b200/scene.py, batch.GroupSpec) against the reference-derived
fixtures and against the oracle's independent scalar restatement of the same init-time math."""
import numpy as np
import pytest

import fizz2d_b200 as fz
from fizz2d_b200 import scene as fscene
from fizz2d_b200.batch import GroupSpec
from conftest import SCENES, golden, scene_file
from helpers import bits, orc


@pytest.mark.parametrize("name", SCENES)
def test_parse_and_constants_match_reference(name):
    g = golden(name)
    sc = fz.parse_in(str(g["scene_text"]))
    assert [len(b.points) for b in sc.free] == list(g["nverts_free"])
    assert [len(p) for p in sc.fixed] == list(g["nverts_fixed"])
    spec = GroupSpec.from_scene(sc, 1)
    F = sc.n_free
    assert np.array_equal(bits(spec.pos[:, 0].reshape(F, 2)), bits(g["init_com"]))
    assert np.array_equal(bits(spec.vel[:, 0].reshape(F, 2)), bits(g["init_vel"]))
    assert np.array_equal(bits(spec.radius[:, 0]), bits(g["init_radius"]))
    assert np.array_equal(bits(spec.mass[:, 0]), bits(g["init_mass"]))
    assert np.array_equal(bits(spec.off[0::2, 0]), bits(g["init_offx"]))
    assert np.array_equal(bits(spec.off[1::2, 0]), bits(g["init_offy"]))
    assert np.array_equal(bits(spec.fixed_com), bits(g["fixed_com"]))
    assert np.array_equal(bits(spec.fixed_radius), bits(g["fixed_radius"]))
    assert np.array_equal(bits(spec.fixed_xy), bits(g["fixed_xy"]))


def test_scene_file_equals_fixture_text():
    for name in SCENES:
        with open(scene_file(name)) as f:
            assert f.read() == str(golden(name)["scene_text"])


@pytest.mark.parametrize("tag,seed,nw,fmax", [("config2_boxfriction", 20260002, 4096, 1),
                                             ("config3_rampworld", 20260003, 65536, 7),
                                             ("config4_mixed", 20260004, 262144, 3)])
def test_perturbed_constants_match_reference(tag, seed, nw, fmax):
    """Vectorised init math over perturbed worlds == what the reference derived world by world."""
    g = golden("perturbed_" + tag)
    deltas = fz.perturbation_deltas(seed, nw, fmax)
    assert np.array_equal(bits(deltas[g["worlds"]]), bits(g["deltas"]))
    for scene in sorted(set(str(s) for s in g["scenes"])):
        sel = [k for k, s in enumerate(g["scenes"]) if str(s) == scene]
        sc = fz.parse_in(str(golden(scene)["scene_text"]))
        d = g["deltas"][sel][:, :sc.n_free, :]
        spec = GroupSpec.from_scene(sc, len(sel), np.ascontiguousarray(d))
        for kk, k in enumerate(sel):
            pre = "w%d_" % int(g["worlds"][k])
            assert np.array_equal(bits(spec.pos[:, kk].reshape(-1, 2)), bits(g[pre + "init_com"]))
            assert np.array_equal(bits(spec.vel[:, kk].reshape(-1, 2)), bits(g[pre + "init_vel"]))
            assert np.array_equal(bits(spec.radius[:, kk]), bits(g[pre + "init_radius"]))
            assert np.array_equal(bits(spec.mass[:, kk]), bits(g[pre + "init_mass"]))
            assert np.array_equal(bits(spec.off[0::2, kk]), bits(g[pre + "init_offx"]))
            assert np.array_equal(bits(spec.off[1::2, kk]), bits(g[pre + "init_offy"]))


def test_world0_is_unperturbed_and_law_is_stable():
    d = fz.perturbation_deltas(20260002, 8, 1)
    assert not d[0].any()
    assert (np.abs(d[1:, :, :2]) <= 5).all() and (np.abs(d[1:, :, 2:]) <= 30).all()
    # prefix property: the first worlds do not depend on the batch size
    assert np.array_equal(d, fz.perturbation_deltas(20260002, 4096, 1)[:8])
    rng = np.random.Generator(np.random.PCG64(20260002))
    assert d[1, 0, 0] == -5 + 10 * rng.random()


def test_constants_agree_with_oracle_scalar_restatement():
    rng = np.random.default_rng(7)
    for n in (3, 4, 5, 8, 12):
        ang = np.sort(rng.uniform(0, 2 * np.pi, (50, n)), axis=1)
        P = np.stack([100 + 20 * np.cos(ang), 200 + 20 * np.sin(ang)], axis=-1)
        c = fz.polygon_constants(P, np.full(50, 2.5))
        for w in range(50):
            o = orc.body_constants(P[w].tolist(), 2.5)
            assert np.array_equal(bits(c["com"][w]), bits(o["com"]))
            assert bits(c["radius"][w]) == bits(o["radius"])
            assert bits(c["mass"][w]) == bits(o["mass"])
            assert np.array_equal(bits(c["offx"][w]), bits(o["offx"]))
            assert np.array_equal(bits(c["offy"][w]), bits(o["offy"]))
            assert np.isclose(c["moi"][w], o["moi"], rtol=1e-12, equal_nan=True)


def test_degenerate_polygons_are_rejected():
    sc = fz.parse_in("100,100\npolygon\ndensity,1\nsides,4\npoint,10,10\npoint,10,10\npoint,20,20\npoint,30,30\n")
    with pytest.raises(ValueError):
        GroupSpec.from_scene(sc, 1)


def test_sqrt_threshold_is_exact():
    """`sqrt(s) > r` (physics.py:277-278) == `s > T(r)` for every s, in particular right at the boundary."""
    rng = np.random.default_rng(3)
    r = np.concatenate([rng.uniform(1e-3, 1e3, 2000), [7.0710678118654755, 34.66666666666667, 1.0, 0.0]])
    T = fscene.sqrt_threshold(r)
    for k in range(-3, 4):
        s = T.copy()
        for _ in range(abs(k)):
            s = np.nextafter(s, np.inf if k > 0 else -np.inf)
        s = np.maximum(s, 0.0)
        assert np.array_equal(np.sqrt(s) > r, s > T)
    s = rng.uniform(0, 1e6, r.size)
    assert np.array_equal(np.sqrt(s) > r, s > T)


def test_host_norm_matches_numpy_linalg_norm():
    """np.linalg.norm([x,y]) == sqrt(fma(y,y,x*x)) on this host's numpy/BLAS."""
    rng = np.random.default_rng(11)
    x, y = rng.normal(size=5000) * 100, rng.normal(size=5000) * 100
    got = fscene.norm2(x, y)
    want = np.array([np.linalg.norm(np.array([a, b])) for a, b in zip(x, y)])
    if not np.array_equal(got, want):
        pytest.skip("this host's BLAS does not fuse the 2-vector dot the way the golden host did (SURVEY Q14)")


def test_param_parsing_and_derived_fields():
    sc = fz.parse_in("10,10\nPARAM ELASTICITY .9\nPARAM EPSILON 1e-9\n")
    p = fscene.resolve_params(sc.params, height=10)
    assert p.elasticity == .9 and p.epsilon == 1e-9
    assert p.one_plus_e == 1 + .9 and p.one_minus_e2 == 1 - .9**2
    assert p.time_disc == .05 and p.acc_x == 0.0 and p.acc_y == 0.0
    with pytest.warns(UserWarning):
        fz.parse_in("10,10\nPARAM TIME_DISC .01\n")
