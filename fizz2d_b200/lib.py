"""ctypes binding of the C-ABI in include/fizz_b200.h (libfizz_b200.so, built in-tree by build.py).

There is no fallback: if the shared library is missing or a CUDA call fails, the error propagates.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FIZZ_B200_LIB") or os.path.join(HERE, "libfizz_b200.so")  # override: A/B of builds

MAX_FREE, MAX_FIXED, MAX_POLY_VERTS, MAX_FIXED_VERTS, MAX_HITS = 8, 16, 16, 96, 32
ST_RUNNING, ST_CAPPED, ST_NONFINITE, ST_HIT_OVERFLOW = 0, 1, 2, 3
MODE_THREAD, MODE_HYBRID, MODE_HYBRID3, MODE_HYBRID_SYNC, MODE_THREAD_FP32 = 0, 1, 2, 3, 4
KERNEL_THREAD, KERNEL_BALLISTIC, KERNEL_CONTACT, KERNEL_PROVER, KERNEL_TRAPPED = 0, 1, 2, 3, 4

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
lp = C.POINTER(C.c_int64)


class FizzParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "elasticity", "epsilon", "time_tol", "collision_tol", "vibrate_tol", "time_disc", "gx", "gy",
        "one_plus_e", "one_minus_e2", "acc_x", "acc_y", "height", "friction_mu", "ff_restitution")] + \
        [("max_calls", C.c_int32), ("reserved", C.c_int32)]


class FizzGroupDesc(C.Structure):
    _fields_ = [("n_free", C.c_int32), ("n_fixed", C.c_int32), ("nverts", ip), ("fixed_xy", dp),
                ("fixed_com", dp), ("fixed_thr", dp), ("params", FizzParams)]


class FizzLayout(C.Structure):
    _fields_ = [("n_worlds", C.c_int64), ("stride", C.c_int64),
                ("n_free", C.c_int32), ("n_fixed", C.c_int32), ("n_free_verts", C.c_int32), ("reserved", C.c_int32)] + \
               [(n, C.c_int64) for n in ("pos", "vel", "heat", "status", "steps", "calls", "vflag", "tier", "state_words",
                                         "off", "thr", "mass", "verts", "energy", "sched", "scene", "total_words")]


class FizzContact(C.Structure):
    _fields_ = [("world", C.c_int64), ("step", C.c_int32), ("call", C.c_int32), ("i", C.c_int32), ("j", C.c_int32),
                ("nx", C.c_double), ("ny", C.c_double), ("mag", C.c_double)]


EXPORTS = {
    # name: (restype, argtypes)
    "fizz_abi_version": (C.c_int, []),
    "fizz_dot_mode": (C.c_int, []),
    "fizz_last_error": (C.c_char_p, []),
    "fizz_host_norm2": (None, [C.c_int64, dp, dp, dp]),
    "fizz_host_dot2": (None, [C.c_int64, dp, dp, dp, dp, dp]),
    "fizz_host_sqrt_threshold": (None, [C.c_int64, dp, dp]),
    "fizz_layout": (C.c_int, [C.POINTER(FizzGroupDesc), C.c_int64, C.POINTER(FizzLayout)]),
    "fizz_group_create": (C.c_int, [C.POINTER(FizzGroupDesc), C.c_int64, C.c_int, C.c_void_p, C.c_int64,
                                    C.POINTER(C.c_void_p)]),
    "fizz_group_destroy": (C.c_int, [C.c_void_p]),
    "fizz_upload": (C.c_int, [C.c_void_p, dp, dp, dp, dp, dp, C.c_void_p]),
    "fizz_step": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "fizz_energy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fizz_step_logged": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "fizz_download": (C.c_int, [C.c_void_p, dp, dp, dp, dp, dp, lp, lp, lp, C.c_void_p]),
    "fizz_set_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "fizz_set_world_stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fizz_kernel_info": (C.c_int, [C.c_void_p, C.c_int32, ip, ip, ip, ip, lp]),
    "fizz_set_tuning": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64]),
    "fizz_snapshot": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fizz_restore": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "fizz_vertices": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fizz_counters": (C.c_int, [C.c_void_p, lp, C.c_int32, C.c_void_p]),
    "fizz_set_spec": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "fizz_spec_stats": (C.c_int, [C.c_void_p, lp, C.c_int32, C.c_void_p]),
    "fizz_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "fizz_profile_read": (C.c_int, [C.c_void_p, dp, lp, C.c_int32]),
    "fizz_fp64_peak": (C.c_int, [C.c_int, C.c_int32, dp, dp]),
    "fizz_raster_frames": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_int64, C.c_int32, ip, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "fizz_points_from_vertices": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                                            C.c_void_p, C.c_void_p, C.c_void_p]),
}
RASTER_MAX_POLYGONS = 64

_lib = None
_dot_mode = None


class FizzError(RuntimeError):
    pass


def numpy_dot_mode(trials=2000, seed=12345):
    """Which 2-vector dot/norm contract does THIS host's numpy/BLAS follow?  1: fma(a1,b1,a0*b0) -- the golden
    host's --, 2: fma(a0,b0,a1*b1), 0: unfused.  Exact integer arithmetic on the operands decides (no libm, no GPU)."""
    import numpy as np
    from fractions import Fraction
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((trials, 4)) * np.exp(rng.uniform(-3, 3, (trials, 4)))
    alive = {0, 1, 2}

    def rn(fr):                                       # Fraction -> nearest double (ties to even): float() does that
        return float(fr)
    for a0, a1, b0, b1 in v:
        got = float(np.dot(np.array([a0, a1]), np.array([b0, b1])))
        A0, A1, B0, B1 = (Fraction(float(x)) for x in (a0, a1, b0, b1))
        want = {1: rn(A1 * B1 + Fraction(rn(A0 * B0))), 2: rn(A0 * B0 + Fraction(rn(A1 * B1))),
                0: rn(Fraction(rn(A0 * B0)) + Fraction(rn(A1 * B1)))}
        alive = {m for m in alive if want[m] == got}
        if not alive:
            raise FizzError("this host's numpy 2-vector dot follows none of the three known contracts (SURVEY Q14)")
    return 1 if 1 in alive else min(alive)


def dot_mode():
    """The contract the loaded library reproduces: $FIZZ_DOT_MODE if set, else what numpy does on this host."""
    global _dot_mode
    if _dot_mode is None:
        env = os.environ.get("FIZZ_DOT_MODE")
        _dot_mode = int(env) if env not in (None, "") else numpy_dot_mode()
    return _dot_mode


def load():
    """Load libfizz_b200.so (raises if it has not been built: run `python -m fizz2d_b200.build` or
    `__graft_entry__.build()`).  On a host whose numpy fuses the 2-vector dot differently from the golden host, the
    library of the matching FIZZ_DOT_MODE is loaded instead (built in-tree on first use: needs nvcc)."""
    global _lib
    if _lib is None:
        path = LIB_PATH
        mode = dot_mode()
        if mode != 1 and not os.environ.get("FIZZ_B200_LIB"):
            from . import build as _build
            path = _build.build_extension(dot_mode=mode)
        if not os.path.isfile(path):
            raise FizzError("CUDA extension %s is missing; build it with __graft_entry__.build()" % path)
        L = C.CDLL(path)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.fizz_abi_version() != 2:
            raise FizzError("ABI version mismatch")
        if L.fizz_dot_mode() != mode and not os.environ.get("FIZZ_B200_LIB"):
            raise FizzError("library %s was built for dot mode %d, this host needs %d" % (path, L.fizz_dot_mode(), mode))
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise FizzError("fizz_b200 error %d: %s" % (rc, load().fizz_last_error().decode()))


def as_dp(a):
    return a.ctypes.data_as(dp)
