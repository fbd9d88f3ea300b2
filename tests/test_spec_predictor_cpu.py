"""CPU: the observation the spec rounds (csrc/fizz_spec.cuh) are built on, checked with the C oracle.

The state a trapped world starts its next timestep from is predictable from the current one, coordinate by coordinate,
by one of three moves: unchanged / Point.move over TIME_DISC / Point.move over the remainder TIME_DISC - dt_k* of a
timestep that fell back (physics.py:133-137, :248-251).  The GPU engine never trusts this -- a predicted start state is
accepted only if it equals the real one bit for bit -- so nothing here is needed for correctness; the test pins the model
the predictor uses (and that it does NOT fit a pair of free bodies stuck together, which is why such worlds go to a runner
launch instead)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cpu_arm  # noqa: E402
import oracle as orc  # noqa: E402

SCENE, SEED, N = "RAMPWORLD_2TRIANGLES_5SQUARES", 20260003, 65536   # BASELINE config 3
DT = orc.TIME_DISC


def remainder_dt():
    d, tol = DT, orc.DEFAULT_PARAMS["TIME_TOL"]
    while d > tol:                       # :106 dt /= 2 until the :99 loop ends on TIME_TOL
        d = d / 2
    return DT - d


def move(p, v, h):                       # Point.move without gravity (:502-529): pos += (vel + acc * h/2) * h, acc = 0
    return p + (v + 0.0 * (.5 * h)) * h


def explained_fraction(world_id, warm, steps):
    w = cpu_arm.port_worlds(SCENE, SEED, N, [world_id], 4096, 1)[0]
    w.run(warm)
    rdt = remainder_dt()
    prev, ok, calls = np.array(w.com()), 0, []
    for _ in range(steps):
        assert w.update()
        cur = np.array(w.com())
        good = np.array_equal(cur[:, 2:], prev[:, 2:])            # velocities unchanged
        for b in range(cur.shape[0]):
            for c in range(2):
                a, v, n = prev[b, c], prev[b, 2 + c], cur[b, c]
                good = good and (n == a or n == move(a, v, DT) or n == move(a, v, rdt))   # bit-exact compares
        ok += good
        calls.append(w.calls_last_step())
        prev = cur
    return ok / steps, calls


def test_trapped_worlds_are_predictable_between_events():
    # world 63739: the longest chain of the benchmark (329 check_collisions() per timestep, an event every ~94 timesteps)
    frac, calls = explained_fraction(63739, 300, 400)
    assert min(calls) >= 329 and frac >= 0.97
    # lighter trapped worlds of the population
    for wid in (103, 152, 532):
        frac, calls = explained_fraction(wid, 400, 300)
        assert min(calls) >= 30 and frac >= 0.95, wid


def test_a_stuck_free_free_pair_is_not():
    # world 27167: two free bodies stuck together far from the scene, their velocities alternate between two values
    frac, calls = explained_fraction(27167, 3000, 100)
    assert min(calls) >= 30 and frac == 0.0
