"""GPU: the spec rounds (time-parallel stepping of expensive worlds, csrc/fizz_spec.cuh) never change a result.

A round evaluates K pieces of one world's future side by side from PREDICTED start states and accepts a piece only if
the state it started from equals, bit for bit, the state the piece before it ended in; everything here checks that the
worlds they step end exactly where the plain chunk pipeline, the thread-per-world kernel and the CPU oracle put them --
whatever the window, the slot length, the threshold, or the hints left behind by an earlier run."""
import os

import numpy as np
import pytest

import fizz2d_b200 as fz
from conftest import scene_file
from helpers import bits, oracle_from_spec, orc

pytestmark = pytest.mark.gpu

SCENE, SEED = "RAMPWORLD_2TRIANGLES_5SQUARES", 20260003   # BASELINE config 3: ~1 % of its worlds stay trapped


def run(n, steps, chunk, mode="hybrid", spec=None, calls=(), max_calls=4096):
    bw = fz.BatchedWorld.from_in(scene_file(SCENE), n_worlds=n, perturb=True, seed=SEED, max_calls=max_calls)
    bw.set_tuning(mode, chunk)
    if spec is not None:
        bw.set_spec(**spec)
    for c in (calls or (steps,)):
        bw.step(c)
    out = (bw.com(), bw.heat(), bw.status(), bw.steps_done(), bw.calls(), bw.vertices())
    return out, bw.spec_stats()[0], bw


def same(a, b):
    for x, y in zip(a, b):
        if x.dtype.kind == "f":
            assert np.array_equal(bits(x), bits(y))
        else:
            assert np.array_equal(x, y)


@pytest.mark.parametrize("spec", [dict(enabled=1), dict(enabled=1, window0=4, slot_cost=50, min_cost=8),
                                  dict(enabled=1, window0=512, slot_cost=4000, min_cost=64)],
                         ids=["default", "short-slots", "long-slots"])
def test_spec_rounds_equal_the_plain_pipeline(spec):
    """2048 worlds x 900 timesteps: spec rounds on (three very different shapes) == spec rounds off == no runners."""
    ref, st0, _ = run(2048, 900, 96, spec=dict(enabled=0))
    assert st0["rounds"] == 0
    sync, _, _ = run(2048, 900, 96, mode="hybrid_sync")
    same(ref, sync)
    got, st, _ = run(2048, 900, 96, spec=spec)
    same(ref, got)
    if spec.get("min_cost", 0) < 64:                                  # (cost is measured: a high bar may select no world)
        assert st["rounds"] > 0 and st["timesteps_accepted"] > 0      # (the rounds did step worlds)
    assert st["timesteps_evaluated"] >= st["timesteps_accepted"]


def test_spec_rounds_vs_oracle_and_thread_kernel():
    """512 worlds x 700 timesteps in three calls: every world against the CPU oracle and the thread-per-world kernel."""
    got, st, bw = run(512, 700, 64, calls=(200, 372, 128))
    assert st["rounds"] > 0
    thread, _, _ = run(512, 700, 1, mode="thread")
    same(got, thread)
    spec = bw.groups[0].spec
    worlds = [oracle_from_spec(spec, w) for w in range(512)]
    orc.run_batch(worlds, 700, os.cpu_count() or 1)
    com, heat, status, steps, calls, verts = got
    for w, ow in enumerate(worlds):
        assert ow.steps_done() == steps[w] and ow.status() == status[w] and ow.total_calls() == calls[w], w
        assert np.array_equal(bits(ow.com()), bits(com[w])) and bits(ow.heat()) == bits(heat[w]), w
        if status[w] == 0:
            assert np.array_equal(bits(ow.verts()), bits(verts[w])), w


def test_stale_hints_do_not_matter():
    """Predictor classes and windows survive calls and restore(): a replay from a snapshot, with the hints of the END of
    the first run, ends in the same bits."""
    bw = fz.BatchedWorld.from_in(scene_file(SCENE), n_worlds=1024, perturb=True, seed=SEED, max_calls=4096)
    bw.set_tuning("hybrid", 96)
    bw.step(150)
    snap = bw.snapshot()
    bw.step(650)
    a = (bw.com(), bw.heat(), bw.status(), bw.steps_done(), bw.calls())
    bw.restore(snap)
    bw.step(650)
    same(a, (bw.com(), bw.heat(), bw.status(), bw.steps_done(), bw.calls()))


def test_watchdog_inside_the_rounds():
    """A small watchdog stops many trapped worlds mid-round: status, the frozen state and the step count are the plain
    pipeline's."""
    ref, _, _ = run(1024, 600, 64, spec=dict(enabled=0), max_calls=64)
    got, st, _ = run(1024, 600, 64, spec=dict(enabled=1, min_cost=8), max_calls=64)
    same(ref, got)
    assert (got[2] != 0).sum() > 50 and st["rounds"] > 0


def test_set_spec_rejects_bad_arguments():
    bw = fz.BatchedWorld.from_in(scene_file(SCENE), n_worlds=8)
    with pytest.raises(fz.FizzError):
        bw.set_spec(window0=100000)
    bw.set_spec(enabled=0).set_spec(enabled=1, window0=16)
