"""CPU: the C-ABI shared library loads, exports every symbol include/fizz_b200.h declares, does its host-side
arithmetic, and fails loudly (no fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from fizz2d_b200 import lib as flib
from fizz2d_b200 import scene as fscene
from conftest import ROOT


def _declared():
    with open(os.path.join(ROOT, "include", "fizz_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fizz_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 14
    L = C.CDLL(flib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "missing export " + n
        assert n in flib.EXPORTS, "python binding missing for " + n
    assert sorted(flib.EXPORTS) == names


def test_abi_version_and_struct_sizes():
    L = flib.load()
    assert L.fizz_abi_version() == 2
    assert C.sizeof(flib.FizzContact) == 48
    assert C.sizeof(flib.FizzParams) == 15 * 8 + 8


def _desc(F=2, X=1, nv=(4, 3, 4)):
    nverts = np.array(nv, dtype=np.int32)
    fx = np.zeros((sum(nv[F:]), 2))
    fc = np.zeros((X, 2))
    ft = np.ones(X)
    d = flib.FizzGroupDesc(n_free=F, n_fixed=X, nverts=nverts.ctypes.data_as(flib.ip), fixed_xy=flib.as_dp(fx),
                           fixed_com=flib.as_dp(fc), fixed_thr=flib.as_dp(ft),
                           params=fscene.resolve_params({}, height=10))
    return d, (nverts, fx, fc, ft)


def test_layout_arithmetic():
    L = flib.load()
    d, keep = _desc()
    lay = flib.FizzLayout()
    assert L.fizz_layout(C.byref(d), 1000, C.byref(lay)) == 0
    assert lay.n_worlds == 1000 and lay.stride == 1024 and lay.n_free_verts == 7
    assert lay.pos == 0 and lay.vel == 4 * 1024 and lay.heat == 8 * 1024
    assert lay.state_words == (8 + 6) * 1024
    assert lay.off == lay.state_words and lay.thr == lay.off + 14 * 1024
    assert lay.sched == lay.energy + 4 * 1024 and lay.scene == lay.sched + 2 * 1024 + 16
    assert lay.total_words >= lay.scene + 4 * 6


def test_bad_arguments_fail_loudly():
    L = flib.load()
    lay = flib.FizzLayout()
    d, keep = _desc(F=2, X=1, nv=(4, 2, 4))     # a 2-gon
    assert L.fizz_layout(C.byref(d), 10, C.byref(lay)) == -1
    assert b"vertices" in L.fizz_last_error()
    d, keep = _desc()
    d.n_free = 9
    assert L.fizz_layout(C.byref(d), 10, C.byref(lay)) == -1
    d, keep = _desc()
    assert L.fizz_layout(C.byref(d), 0, C.byref(lay)) == -1
    assert L.fizz_step(None, 1, None) == -1
    with pytest.raises(flib.FizzError):
        flib.check(L.fizz_upload(None, None, None, None, None, None, None))


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = flib.load()
    d, keep = _desc()
    lay = flib.FizzLayout()
    assert L.fizz_layout(C.byref(d), 10, C.byref(lay)) == 0
    h = C.c_void_p()
    buf = np.zeros(lay.total_words)
    rc = L.fizz_group_create(C.byref(d), 10, 0, C.c_void_p(buf.ctypes.data), lay.total_words, C.byref(h))
    assert rc == -2 and not h.value
    import fizz2d_b200 as fz
    with pytest.raises(Exception):
        fz.BatchedWorld.from_text("10,10\npolygon\ndensity,1\nsides,3\npoint,1,1\npoint,5,1\npoint,3,4\n")
    t, ms = C.c_double(), C.c_double()
    assert L.fizz_fp64_peak(0, 16, C.byref(t), C.byref(ms)) == -2


def test_host_helpers():
    L = flib.load()
    a0, a1, b0, b1 = (np.array([1.5, 3.0]), np.array([2.0, 4.0]), np.array([3.0, 0.1]), np.array([0.5, 0.3]))
    out = np.empty(2)
    L.fizz_host_dot2(2, flib.as_dp(a0), flib.as_dp(a1), flib.as_dp(b0), flib.as_dp(b1), flib.as_dp(out))
    import math
    assert out[0] == 5.5 and out[1] == math.fma(4.0, 0.3, 3.0 * 0.1) if hasattr(math, "fma") else True
    assert fscene.norm2(np.array([3.0]), np.array([4.0]))[0] == 5.0
