"""GPU: the CUDA engine (through the C-ABI) against the reference-derived fixtures and against the CPU oracle.
Bit-exact everywhere: FP64 positions/velocities/heat/energy are compared as 64-bit patterns, contact pairs as
index tuples (tolerance stated by north_star: 1e-9 relative; achieved: 0 ulp)."""
import hashlib
import json
import os

import numpy as np
import pytest

import fizz2d_b200 as fz
from fizz2d_b200.batch import GroupSpec
from conftest import GOLDEN, SCENES, golden, scene_file
from helpers import bits, contact_rows, oracle_from_spec, orc

pytestmark = pytest.mark.gpu

# execution strategies (results must never depend on them): hybrid = ballistic thread-per-world kernel + warp-per-
# world contact kernel per chunk of timesteps, trapped worlds on runner launches that skip the chunk boundaries (calls of
# >= 128 timesteps); hybrid_sync = without runners; thread = one thread-per-world kernel with the full semantics
MODES = [("hybrid", 256), ("hybrid3", 7), ("thread", 1)]


def tune(bw, mode):
    bw.set_tuning(mode[0], mode[1])
    return bw


@pytest.mark.parametrize("mode", [MODES[0], MODES[2]], ids=["hybrid", "thread"])
@pytest.mark.parametrize("name", SCENES + ["gravity_BOX_1SQUARE", "gravity_PLANE_2SQUARES"])
def test_single_world_per_step_trace(name, mode):
    """One launch per timestep: com, vertices, heat, energy rows and the contact trace of every step."""
    g = golden(name)
    bw = tune(fz.BatchedWorld.from_text(str(g["scene_text"]), gravity=tuple(g["gravity"])), mode)
    bw.enable_trace(1 << 16)
    n = int(g["steps_done"])
    for t in range(n):
        assert np.array_equal(bits(bw.energy()[0]), bits(g["energy_before"][t])), (name, t)
        bw.step(1)
        assert np.array_equal(bits(bw.com()[0]), bits(g["com"][t])), (name, t)
        assert np.array_equal(bits(bw.vertices()[0]), bits(g["verts"][t])), (name, t)
        assert bits(bw.heat()[0]) == bits(g["heat"][t]), (name, t)
    assert bw.calls()[0] == int(g["calls"].sum())
    got = [c[1:] for c in bw.contacts()]
    want = contact_rows(g["contacts_idx"], g["contacts_val"])
    assert [c[:4] for c in got] == [c[:4] for c in want]
    assert np.array_equal(bits([c[4:] for c in got]), bits([c[4:] for c in want]))


@pytest.mark.parametrize("mode", MODES, ids=["hybrid256", "hybrid3-7", "thread"])
@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("chunks", [(320,), (1, 7, 100, 212)])
def test_fused_steps_equal_stepwise(name, chunks, mode):
    """K fused steps per call (state resident on chip) end in exactly the per-step state."""
    g = golden(name)
    bw = tune(fz.BatchedWorld.from_text(str(g["scene_text"])), mode)
    done = 0
    for c in chunks:
        bw.step(c)
        done += c
        assert np.array_equal(bits(bw.com()[0]), bits(g["com"][done - 1]))
        assert np.array_equal(bits(bw.vertices()[0]), bits(g["verts"][done - 1]))
        assert bits(bw.heat()[0]) == bits(g["heat"][done - 1])
    assert bw.steps_done()[0] == done and bw.status()[0] == 0


@pytest.mark.parametrize("name", ["BOX_1SQUARE", "BOX_1SQUAREFRICTION", "RAMPWORLD_2TRIANGLES_5SQUARES",
                                  "2CHAMBERS_3BALLS", "STACKEDCHAMBERS_1SQUARE", "PENTAGONVOID_1SQUARE"])
def test_1000_step_anchors_through_text_formats(name):
    """SURVEY 8c table through the drop-in formats: plane_%03d.txt bodies and energy.txt rows, byte for byte."""
    with open(os.path.join(GOLDEN, "anchors.json")) as f:
        anchors = json.load(f)[name]
    bw = fz.BatchedWorld.from_in(scene_file(name))
    bw.enable_trace(1 << 16)
    hp, he = hashlib.sha256(), hashlib.sha256()
    for t in range(1000):
        he.update((','.join(str(x) for x in bw.energy()[0]) + '\n').encode())
        bw.step(1)
        hp.update(bw.dump_world(0).encode())
    assert hp.hexdigest() == anchors["planes"]
    assert he.hexdigest() == anchors["energy"]
    hc = hashlib.sha256('\n'.join('%d:%d:%d-%d' % c[1:5] for c in bw.contacts()).encode()).hexdigest()
    assert hc == anchors["contacts_relevant"]
    assert bw.calls()[0] == anchors["calls_relevant"]
    assert np.array_equal(bits(bw.energy()[0]), bits(anchors["final_energy"]))
    assert np.array_equal(bits(bw.com()[0]), bits(anchors["final_com"]))


@pytest.mark.parametrize("mode", MODES, ids=["hybrid256", "hybrid3-7", "thread"])
@pytest.mark.parametrize("tag,seed,nw,fmax", [("config2_boxfriction", 20260002, 4096, 1),
                                             ("config3_rampworld", 20260003, 65536, 7),
                                             ("config4_mixed", 20260004, 262144, 3),
                                             ("critical_rampworld", 20260003, 65536, 7)])
def test_perturbed_worlds_match_reference(tag, seed, nw, fmax, mode):
    """Worlds of the BASELINE configs (same seeds, same law) that the reference itself was run on, including
    worlds the watchdog caps."""
    g = golden("perturbed_" + tag)
    mc, nsteps = int(g["max_calls"]), int(g["nsteps"])
    cps = [int(c) for c in g["checkpoints"]]
    for scene in sorted(set(str(s) for s in g["scenes"])):
        sel = [k for k, s in enumerate(g["scenes"]) if str(s) == scene]
        sc = fz.parse_in(str(golden(scene)["scene_text"]))
        d = np.ascontiguousarray(g["deltas"][sel][:, :sc.n_free, :])
        bw = tune(fz.BatchedWorld([GroupSpec.from_scene(sc, len(sel), d, max_calls=mc)]), mode)
        bw.enable_trace(1 << 20)
        done = 0
        for c in cps:
            bw.step(c - done)
            done = c
            com = bw.com()
            for kk, k in enumerate(sel):
                key = "w%d_com_at_%d" % (int(g["worlds"][k]), c)
                if key in g.files:
                    assert np.array_equal(bits(com[kk]), bits(g[key])), (scene, k, c)
        assert done == nsteps
        com, verts, heat, st, sd, en = bw.com(), bw.vertices(), bw.heat(), bw.status(), bw.steps_done(), bw.energy()
        cons = bw.contacts()
        for kk, k in enumerate(sel):
            pre = "w%d_" % int(g["worlds"][k])
            assert sd[kk] == int(g[pre + "steps_done"]), (scene, k)
            assert (st[kk] == fz.ST_CAPPED) == (int(g[pre + "capped_at"]) >= 0)
            assert np.array_equal(bits(com[kk]), bits(g[pre + "final_com"])), (scene, k)
            assert np.array_equal(bits(verts[kk]), bits(g[pre + "final_verts"])), (scene, k)
            assert bits(heat[kk]) == bits(g[pre + "final_heat"])
            if st[kk] == 0:
                assert np.array_equal(bits(en[kk]), bits(g[pre + "final_energy"]))
            got = [c[1:] for c in cons if c[0] == kk]
            want = contact_rows(g[pre + "contacts_idx"], g[pre + "contacts_val"])
            assert [c[:4] for c in got] == [c[:4] for c in want], (scene, k)
            assert np.array_equal(bits([c[4:] for c in got]), bits([c[4:] for c in want]))


@pytest.mark.parametrize("mode", [("hybrid", 256), ("hybrid", 16), ("hybrid3", 32), ("hybrid_sync", 64)],
                         ids=["hybrid256", "hybrid16", "hybrid3-32", "hybrid_sync64"])
def test_critical_worlds_match_reference_without_trace(mode):
    """The benchmark regime against the REFERENCE ITSELF (fixture perturbed_critical_rampworld: the two longest chains of
    BASELINE config 3, max_calls = 4096, 130 timesteps, ~330 check_collisions() per timestep once trapped), with the
    contact trace OFF: the speculative halving search, its all-trials-fail proof and the register-resident trapped path
    are what runs here -- state, vertices, heat and the per-world call total must still be the reference's bits."""
    g = golden("perturbed_critical_rampworld")
    sc = fz.parse_in(str(golden("RAMPWORLD_2TRIANGLES_5SQUARES")["scene_text"]))
    d = np.ascontiguousarray(g["deltas"][:, :sc.n_free, :])
    bw = tune(fz.BatchedWorld([GroupSpec.from_scene(sc, len(d), d, max_calls=int(g["max_calls"]))]), mode)
    done = 0
    for c in [int(c) for c in g["checkpoints"]]:
        bw.step(c - done)
        done = c
        com = bw.com()
        for k, wi in enumerate(g["worlds"]):
            assert np.array_equal(bits(com[k]), bits(g["w%d_com_at_%d" % (int(wi), c)])), (wi, c)
    com, verts, heat, calls, k4 = bw.com(), bw.vertices(), bw.heat(), bw.calls(), bw.counters()[0]
    assert k4["fast_path_calls"] > 0.7 * k4["contact_calls"]
    for k, wi in enumerate(g["worlds"]):
        pre = "w%d_" % int(wi)
        assert np.array_equal(bits(com[k]), bits(g[pre + "final_com"])) and bits(heat[k]) == bits(g[pre + "final_heat"])
        assert np.array_equal(bits(verts[k]), bits(g[pre + "final_verts"]))
        assert calls[k] == int(g[pre + "calls"].sum()), wi


@pytest.mark.parametrize("mode", [("hybrid", 50), ("hybrid3", 32), ("thread", 1)], ids=["hybrid50", "hybrid3-32", "thread"])
@pytest.mark.parametrize("name,seed,n,steps,mc", [("BOX_1SQUAREFRICTION", 20260002, 4096, 400, 4096),
                                                   ("RAMPWORLD_2TRIANGLES_5SQUARES", 20260003, 1024, 400, 256),
                                                   ("2CHAMBERS_3BALLS", 20260004, 1024, 300, 256),
                                                   ("CORRIDOR_2SQUARES", 5, 777, 300, 512)])
@pytest.mark.parametrize("trace", [True, False], ids=["trace", "notrace"])
def test_batch_vs_oracle(name, seed, n, steps, mc, mode, trace):
    """Whole batches against the CPU oracle fed with the same constants: state, status, step counters, call
    counts and contact pairs of every world.  Run twice: the contact trace switches the speculative halving search and
    the register-resident single-contact path OFF (they never see every trial's tuples), so the trace-less run is the
    one that pins the paths production and bench runs take."""
    bw = tune(fz.BatchedWorld.from_in(scene_file(name), n_worlds=n, perturb=True, seed=seed, max_calls=mc), mode)
    if trace:
        bw.enable_trace(1 << 22)
    bw.step(steps // 2)
    bw.step(steps - steps // 2)
    com, verts, heat, st, sd, calls = bw.com(), bw.vertices(), bw.heat(), bw.status(), bw.steps_done(), bw.calls()
    cons = bw.contacts() if trace else []
    by_world = {}
    for c in cons:
        by_world.setdefault(c[0], []).append(c[1:])
    spec = bw.groups[0].spec
    traced = set(range(0, n, max(1, n // 48)))      # full contact traces for a spread-out subset (memory)
    worlds = [oracle_from_spec(spec, w, trace_cap=(1 << 17) if w in traced else 0) for w in range(n)]
    orc.run_batch(worlds, steps, os.cpu_count() or 1)
    ncapped = 0
    for w, ow in enumerate(worlds):
        assert ow.steps_done() == sd[w] and ow.status() == st[w], (name, w)
        assert np.array_equal(bits(ow.com()), bits(com[w])), (name, w)
        assert bits(ow.heat()) == bits(heat[w]), (name, w)
        assert ow.total_calls() == calls[w], (name, w)
        if st[w] == 0:
            assert np.array_equal(bits(ow.verts()), bits(verts[w])), (name, w)
        if trace and w in traced:
            got = by_world.get(w, [])
            want = ow.contacts()
            assert [c[:4] for c in got] == [c[:4] for c in want], (name, w)
            assert np.array_equal(bits([c[4:] for c in got]), bits([c[4:] for c in want])), (name, w)
        ncapped += st[w] != 0
    print(name, "worlds", n, "capped", ncapped, "contacts", len(cons))


def test_energy_bookkeeping_at_full_size():
    """BASELINE config 2 at full size (4096 worlds x 1000 steps): total energy is conserved to rounding because
    heat is booked as exactly the kinetic energy removed (physics.py:191-195); uncapped worlds only."""
    bw = fz.BatchedWorld.from_in(scene_file("BOX_1SQUAREFRICTION"), n_worlds=4096, perturb=True, seed=20260002)
    e0 = bw.energy()
    bw.step(1000)
    e1, st = bw.energy(), bw.status()
    ok = st == 0
    assert ok.sum() > 3000
    rel = np.abs(e1[ok, 0] - e0[ok, 0]) / e0[ok, 0]
    assert rel.max() < 1e-12
    assert (e1[ok, 3] >= 0).all() and (bw.steps_done()[ok] == 1000).all()
    # world 0 is the unperturbed scene: the SURVEY 8c anchor row
    assert e1[0].tolist() == [3124999.9999999963, 126.51176882138402, 0.0, 3124873.488231175]


@pytest.mark.parametrize("mode", [MODES[0], MODES[2]], ids=["hybrid", "thread"])
def test_relaunch_is_idempotent_for_frozen_worlds(mode):
    bw = tune(fz.BatchedWorld.from_in(scene_file("RAMPWORLD_2TRIANGLES_5SQUARES"), n_worlds=256, perturb=True,
                                      seed=20260003, max_calls=64), mode)
    bw.step(400)
    st, com, sd = bw.status(), bw.com(), bw.steps_done()
    assert (st == fz.ST_CAPPED).any()
    bw.step(50)
    st2, com2, sd2 = bw.status(), bw.com(), bw.steps_done()
    frozen = st != 0
    assert np.array_equal(st[frozen], st2[frozen]) and np.array_equal(sd[frozen], sd2[frozen])
    assert np.array_equal(bits(com[frozen]), bits(com2[frozen]))
    assert (sd2[~frozen & (st2 == 0)] == 450).all()


@pytest.mark.parametrize("mode", [("hybrid", 256), ("hybrid3", 32), ("hybrid", 5), ("hybrid_sync", 64)],
                         ids=["hybrid256", "hybrid3-32", "hybrid5", "hybrid_sync64"])
def test_trapped_worlds_vs_oracle(mode):
    """The worlds that bound a pass of the benchmark batch -- a square wedged between fixed bodies, ~300 resolve passes per
    timestep, almost all of them in the single-contact fast path -- plus their neighbours, bit for bit against the
    oracle; the uncapped default watchdog and a tight one that trips inside the resolve loops."""
    heavy = [63739, 12860, 5657, 58248, 44270, 36291, 36227, 47547]
    sel = sorted(set(heavy + [w + d for w in heavy[:4] for d in (-1, 1)]))
    sc = fz.load_in(scene_file("RAMPWORLD_2TRIANGLES_5SQUARES"))
    d = np.ascontiguousarray(fz.perturbation_deltas(20260003, 65536, sc.n_free)[sel])
    for mc, steps in ((4096, 400), (300, 400), (97, 300)):
        spec = GroupSpec.from_scene(sc, len(sel), d, max_calls=mc)
        bw = tune(fz.BatchedWorld([spec]), mode)
        bw.step(steps // 3)
        bw.step(steps - steps // 3)
        com, verts, heat, st, sd, calls = bw.com(), bw.vertices(), bw.heat(), bw.status(), bw.steps_done(), bw.calls()
        k = bw.counters()[0]
        if mc == 4096:
            assert calls.max() > 250 * steps * 0.5 and k["fast_path_calls"] > 0.8 * k["contact_calls"], k
        else:
            assert (st == fz.ST_CAPPED).any()
        worlds = [oracle_from_spec(spec, w) for w in range(len(sel))]
        orc.run_batch(worlds, steps, os.cpu_count() or 1)
        for w, ow in enumerate(worlds):
            tag = (mc, sel[w])
            assert ow.steps_done() == sd[w] and ow.status() == st[w] and ow.total_calls() == calls[w], tag
            assert np.array_equal(bits(ow.com()), bits(com[w])) and bits(ow.heat()) == bits(heat[w]), tag
            if st[w] == 0:
                assert np.array_equal(bits(ow.verts()), bits(verts[w])), tag


@pytest.mark.parametrize("mode", [("hybrid", 96), ("hybrid", 7), ("hybrid3", 32), ("hybrid_sync", 64)],
                         ids=["hybrid96", "hybrid7", "hybrid3-32", "hybrid_sync64"])
def test_two_trapped_bodies_vs_oracle(mode):
    """Worlds of the 8-GPU population of BASELINE config 3 (524288 worlds, seed 20260003) that bound its slowest shards: TWO
    squares trapped at once, each inside a fixed body (every check_collisions() of the resolve loop returns two tuples until
    one is out) -- the contact kernel runs the two resolve loops one after the other in registers and validates the
    decomposition on the trajectories' bounding boxes; two bodies inside the SAME fixed body (the validation fails: general
    path); a free-free pair stuck together.  Bit for bit against the oracle, default and tight watchdog."""
    worlds = [235581, 309402, 305392, 161573, 502740, 404382, 248613, 63739]
    sc = fz.load_in(scene_file("RAMPWORLD_2TRIANGLES_5SQUARES"))
    d = np.ascontiguousarray(fz.perturbation_deltas(20260003, 524288, sc.n_free)[worlds])
    for mc, steps in ((4096, 1500), (100, 600)):
        spec = GroupSpec.from_scene(sc, len(worlds), d, max_calls=mc)
        bw = tune(fz.BatchedWorld([spec]), mode)
        bw.step(steps // 3)
        bw.step(steps - steps // 3)
        com, verts, heat, st, sd, calls = bw.com(), bw.vertices(), bw.heat(), bw.status(), bw.steps_done(), bw.calls()
        ows = [oracle_from_spec(spec, w) for w in range(len(worlds))]
        orc.run_batch(ows, steps, os.cpu_count() or 1)
        for w, ow in enumerate(ows):
            tag = (mode, mc, worlds[w])
            assert ow.steps_done() == sd[w] and ow.status() == st[w] and ow.total_calls() == calls[w], tag
            assert np.array_equal(bits(ow.com()), bits(com[w])) and bits(ow.heat()) == bits(heat[w]), tag
            if st[w] == 0:
                assert np.array_equal(bits(ow.verts()), bits(verts[w])), tag
        if mc == 4096:
            assert calls[0] > 100 * steps // 2      # the two-body world is busy from early on


def test_runner_mode_call_patterns():
    """Runner launches only exist inside fizz_step() calls of >= 128 timesteps and end with them: any split of a run
    into calls -- long ones that use runners, short ones that do not, a mix -- ends in the same bits as one call and as
    the pipeline without runners."""
    def run(mode, chunk, calls):
        bw = tune(fz.BatchedWorld.from_in(scene_file("RAMPWORLD_2TRIANGLES_5SQUARES"), n_worlds=4096, perturb=True,
                                          seed=20260003, max_calls=512), (mode, chunk))
        for c in calls:
            bw.step(c)
        return bw.com(), bw.heat(), bw.status(), bw.steps_done(), bw.calls(), bw.vertices()
    ref = run("hybrid_sync", 256, (1500,))
    for mode, chunk, calls in (("hybrid", 64, (1500,)), ("hybrid", 32, (127, 128, 129, 1116)), ("hybrid", 256, (700, 1, 799)),
                               ("hybrid", 17, (300, 300, 300, 300, 300))):
        got = run(mode, chunk, calls)
        for x, y in zip(ref, got):
            assert np.array_equal(bits(x) if x.dtype == np.float64 else x, bits(y) if y.dtype == np.float64 else y), (mode, chunk, calls)


def test_runner_stress_many_worlds_tiny_chunks():
    """Many runner-owned worlds and a chunk boundary every 4 timesteps: a work-list kernel of the chunk pipeline must
    never list a world a runner launch is storing (a world listed twice would be stepped past the target); every running
    world ends exactly at the target, in the bits of the pipeline without runners."""
    heavy = [63739, 12860, 5657, 58248, 44270, 36291, 36227, 47547, 29797]
    sc = fz.load_in(scene_file("RAMPWORLD_2TRIANGLES_5SQUARES"))
    d = np.ascontiguousarray(np.tile(fz.perturbation_deltas(20260003, 65536, sc.n_free)[heavy], (57, 1, 1)))
    outs = []
    for mode in (("hybrid_sync", 64), ("hybrid", 4), ("hybrid", 3)):
        bw = tune(fz.BatchedWorld([GroupSpec.from_scene(sc, len(d), d, max_calls=1024)]), mode)
        for c in (400, 131, 260):
            bw.step(c)
        st, sd = bw.status(), bw.steps_done()
        assert (sd[st == 0] == 791).all(), (mode, np.unique(sd[st == 0]))
        outs.append((bw.com(), bw.heat(), st, sd, bw.calls()))
    for other in outs[1:]:
        for x, y in zip(outs[0], other):
            assert np.array_equal(bits(x) if x.dtype == np.float64 else x, bits(y) if y.dtype == np.float64 else y)


def test_runner_mode_mixed_groups():
    """Several groups step on their own streams, each with its own runner streams: same bits as without runners."""
    paths = [scene_file("2CHAMBERS_3BALLS"), scene_file("STACKEDCHAMBERS_1SQUARE"), scene_file("RAMPWORLD_2TRIANGLES_5SQUARES")]
    outs = []
    for mode in (("hybrid_sync", 256), ("hybrid", 32)):
        bw = tune(fz.BatchedWorld.from_mixed(paths, 6000, seed=20260004, max_calls=384), mode)
        bw.step(450)
        bw.step(200)
        outs.append((bw.heat(), bw.status(), bw.steps_done(), bw.calls(), np.concatenate([c.ravel() for c in bw.com()])))
    for x, y in zip(*outs):
        assert np.array_equal(bits(x) if x.dtype == np.float64 else x, bits(y) if y.dtype == np.float64 else y)


def test_upload_resets_state():
    bw = fz.BatchedWorld.from_in(scene_file("BOX_1SQUARE"), n_worlds=100, perturb=True, seed=1)
    bw.step(200)
    a = bw.com()
    bw.upload()
    assert (bw.steps_done() == 0).all() and (bw.heat() == 0).all()
    bw.step(200)
    assert np.array_equal(bits(a), bits(bw.com()))


def test_snapshot_restore_replays_bit_exactly():
    bw = fz.BatchedWorld.from_in(scene_file("2CHAMBERS_3BALLS"), n_worlds=300, perturb=True, seed=9, max_calls=256)
    bw.step(40)
    snap = bw.snapshot()
    bw.step(260)
    a, ha, sa = bw.com(), bw.heat(), bw.steps_done()
    bw.restore(snap)
    assert (bw.steps_done()[bw.status() == 0] == 40).all()
    bw.step(260)
    assert np.array_equal(bits(a), bits(bw.com())) and np.array_equal(bits(ha), bits(bw.heat()))
    assert np.array_equal(sa, bw.steps_done())


def test_hybrid_equals_thread_mode_large_batch():
    """The two independent kernel families agree bit for bit on 8192 perturbed RAMPWORLD worlds x 600 steps."""
    outs = []
    for mode in (("hybrid", 128), ("hybrid_sync", 256), ("hybrid3", 32), ("thread", 1)):
        bw = tune(fz.BatchedWorld.from_in(scene_file("RAMPWORLD_2TRIANGLES_5SQUARES"), n_worlds=8192, perturb=True,
                                          seed=20260003, max_calls=128), mode)
        bw.step(600)
        outs.append((bw.com(), bw.heat(), bw.status(), bw.steps_done(), bw.calls(), bw.vertices()))
    for other in outs[1:]:
        for x, y in zip(outs[0], other):
            assert np.array_equal(bits(x) if x.dtype == np.float64 else x, bits(y) if y.dtype == np.float64 else y)


def test_batched_energy_log_equals_reference_rows():
    """step_logged: the per-step energy rows of every world; world 0 reproduces the reference's energy.txt."""
    g = golden("BOX_1SQUAREFRICTION")
    bw = fz.BatchedWorld.from_in(scene_file("BOX_1SQUAREFRICTION"), n_worlds=200, perturb=True, seed=20260002)
    n = int(g["steps_done"])
    log = bw.step_logged(n)
    assert log.shape == (n, 200, 4)
    assert np.array_equal(bits(log[:, 0, :]), bits(g["energy_before"]))
    assert np.array_equal(bits(bw.com()[0]), bits(g["com"][n - 1]))
    # conservation for every uncapped world, every step
    ok = bw.status() == 0
    rel = np.abs(log[:, ok, 0] - log[0, ok, 0]) / log[0, ok, 0]
    assert rel.max() < 1e-12
