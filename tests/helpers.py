"""Shared helpers for the parity tests (test infrastructure; may use oracle/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402


def oracle_from_spec(spec, w, max_calls=None, trace_cap=0, gravity=None):
    """CPU-oracle world fed with the PRODUCT loader's constants for world w of a GroupSpec."""
    p = spec.params
    prm = orc.make_params(spec.scene.params, spec.height, max_calls if max_calls is not None else p.max_calls,
                          gravity if gravity is not None else (p.gx, p.gy), friction=p.friction_mu,
                          ff_restitution=p.ff_restitution)
    return orc.OracleWorld(spec.nverts_free, spec.nverts_fixed,
                           np.concatenate([spec.pos[:, w].reshape(-1, 2), spec.fixed_com]),
                           spec.vel[:, w].reshape(-1, 2), np.concatenate([spec.radius[:, w], spec.fixed_radius]),
                           spec.mass[:, w], spec.off[0::2, w], spec.off[1::2, w], spec.fixed_xy, prm,
                           width=spec.width, height=spec.height, trace_cap=trace_cap)


def oracle_from_golden(g, max_calls=4096, trace_cap=0):
    """CPU-oracle world fed with the init constants the REFERENCE derived (stored in the fixture)."""
    text = str(g["scene_text"])
    sc = orc.parse_scene(text)
    prm = orc.make_params(sc["params"], sc["height"], max_calls, tuple(g["gravity"]))
    return orc.OracleWorld(list(g["nverts_free"]), list(g["nverts_fixed"]),
                           np.concatenate([g["init_com"], g["fixed_com"]]), g["init_vel"],
                           np.concatenate([g["init_radius"], g["fixed_radius"]]), g["init_mass"],
                           g["init_offx"], g["init_offy"], g["fixed_xy"], prm,
                           width=sc["width"], height=sc["height"], trace_cap=trace_cap)


def contact_rows(idx, val):
    return [(int(a[0]), int(a[1]), int(a[2]), int(a[3]), float(v[0]), float(v[1]), float(v[2]))
            for a, v in zip(idx, val)]


def bits(a):
    """View float64 as int64 so that comparisons are bit-exact (NaN == NaN, -0.0 != 0.0)."""
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)
