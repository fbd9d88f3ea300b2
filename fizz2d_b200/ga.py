"""GA population front end: one random convex polygon per world dropped into a fixed scene.

The reference only *mentions* genetic algorithms (README.md:1); it defines no population, no fitness.  The law below
is this project's definition; it is deterministic, independent of the population size (worlds are drawn in blocks of
4096 from `PCG64(SeedSequence([seed, block]))`) and rejects the shapes the reference cannot simulate:

  per world:  n ~ U{3..12};  n angles ~ U(0, 2pi), sorted, all gaps (wrap-around included) >= min_gap;
              radius r ~ U(6, 14);  vertices centre + r (cos a, sin a)  (counter-clockwise);
              density 1;  velocity (U(-20, 20), +-U(40, 80)) (the reference has no gravity from .in files, so
              "dropped" is an initial velocity);  redraw while the reference's init-time math gives a NaN/zero
              moment of inertia or a vertex on the centroid.
Fitness (reported by the multi-GPU gather): world.heat, the dissipated energy.
"""
from __future__ import annotations

import numpy as np

from . import scene as _scene

BLOCK = 4096


def _draw_block(rng, m, n_lo, n_hi, min_gap):
    n = rng.integers(n_lo, n_hi + 1, size=m)
    ang = rng.random((m, n_hi)) * (2 * np.pi)
    r = 6.0 + 8.0 * rng.random(m)
    vx = -20.0 + 40.0 * rng.random(m)
    vy = (40.0 + 40.0 * rng.random(m)) * np.where(rng.random(m) < 0.5, 1.0, -1.0)
    return n, ang, r, vx, vy


def random_convex_polygons(seed: int, n_worlds: int, center=(112.5, 290.0), n_lo=3, n_hi=12, min_gap=0.2, eps=1e-12):
    """Returns (nverts (W,), points list of (n,2) arrays, velocity (W,2))."""
    nverts = np.zeros(n_worlds, dtype=np.int64)
    pts = [None] * n_worlds
    vel = np.zeros((n_worlds, 2))
    col = np.arange(n_hi)[None, :]
    for blk in range((n_worlds + BLOCK - 1) // BLOCK):
        lo = blk * BLOCK
        m_total = min(BLOCK, n_worlds - lo)
        rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([seed, blk])))
        pending = np.arange(BLOCK)          # always draw for a full block: worlds do not depend on n_worlds
        while len(pending):
            n, ang, r, vx, vy = _draw_block(rng, len(pending), n_lo, n_hi, min_gap)
            a = np.sort(np.where(col < n[:, None], ang, np.inf), axis=1)         # first n entries sorted, rest inf
            first = a[:, 0]
            last = a[np.arange(len(n)), n - 1]
            with np.errstate(invalid="ignore"):
                gaps = np.diff(a, axis=1)
            gaps = np.where(col[:, 1:] < n[:, None], gaps, np.inf)
            ok = (np.minimum(gaps.min(axis=1), first + 2 * np.pi - last) >= min_gap)
            for nv in range(n_lo, n_hi + 1):
                sel = np.nonzero(ok & (n == nv))[0]
                if len(sel) == 0:
                    continue
                av = a[sel, :nv]
                P = np.stack([center[0] + r[sel, None] * np.cos(av), center[1] + r[sel, None] * np.sin(av)], axis=2)
                valid = _scene.polygon_constants(P, np.ones(len(sel)), eps)["valid"]
                ok[sel[~valid]] = False
                for kk, k in enumerate(sel):
                    slot = pending[k]
                    if valid[kk] and slot < m_total:
                        w = lo + slot
                        nverts[w] = nv
                        pts[w] = P[kk]
                        vel[w] = (vx[k], vy[k])
            pending = pending[~ok]
    return nverts, pts, vel


def polygon_in_text(points, velocity, density=1.0):
    """`.in` stanza of one free polygon (repr floats: float() reproduces them bit for bit)."""
    lines = ["polygon", "density,%r" % float(density), "sides,%d" % len(points),
             "velocity,%r,%r" % (float(velocity[0]), float(velocity[1]))]
    lines += ["point,%r,%r" % (float(p[0]), float(p[1])) for p in points]
    return "\n".join(lines) + "\n"


def fixed_scene_text(scene: _scene.Scene):
    """`.in` text of the fixed geometry + PARAMs of a scene (its free bodies dropped)."""
    out = "%d,%d\n" % (scene.width, scene.height)
    for pts in scene.fixed:
        out += "fixedpolygon\nsides,%d\n" % len(pts) + "".join("point,%r,%r\n" % (float(p[0]), float(p[1])) for p in pts)
    for k, v in scene.params.items():
        out += "PARAM %s %r\n" % (k, v)
    return out


def next_generation(population, fitness, seed, generation, survivors=0.25, sigma=0.6, center=(112.5, 290.0), eps=1e-12):
    """One GA generation step on the host: keep the `survivors` fraction with the highest fitness,
    refill the population with mutated copies (every vertex jittered by N(0, sigma) px, velocity by N(0, 3 sigma)),
    re-sorted by polar angle about the centroid so the polygon stays simple; candidates that are not strictly
    convex or that the reference cannot simulate are redrawn.  Deterministic in (seed, generation).
    population = (nverts, points, velocity) as returned by random_convex_polygons; returns the same triple."""
    nverts, pts, vel = population
    W = len(pts)
    order = np.argsort(-np.asarray(fitness), kind="stable")
    keep = order[:max(1, int(round(W * survivors)))]
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([seed, 1000003, generation])))
    new_n = np.empty(W, dtype=np.int64)
    new_pts = [None] * W
    new_vel = np.empty((W, 2))
    for slot in range(W):
        if slot < len(keep):
            w = keep[slot]
            new_n[slot], new_pts[slot], new_vel[slot] = nverts[w], pts[w], vel[w]
            continue
        parent = keep[rng.integers(0, len(keep))]
        for _ in range(64):
            P = pts[parent] + rng.normal(0.0, sigma, pts[parent].shape)
            c = P.mean(axis=0)
            P = P[np.argsort(np.arctan2(P[:, 1] - c[1], P[:, 0] - c[0]))]
            e = np.roll(P, -1, axis=0) - P
            cross = e[:, 0] * np.roll(e, -1, axis=0)[:, 1] - e[:, 1] * np.roll(e, -1, axis=0)[:, 0]
            if (cross > 1e-6).all() and bool(_scene.polygon_constants(P[None], np.ones(1), eps)["valid"][0]):
                break
        else:
            P = pts[parent]
        new_n[slot], new_pts[slot] = len(P), P
        new_vel[slot] = vel[parent] + rng.normal(0.0, 3 * sigma, 2)
    return new_n, new_pts, new_vel
