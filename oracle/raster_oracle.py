"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/raster_oracle.c (CPU restatement of the reference's draw.c)
plus an independent parser of the `plane_%03d.txt` format (draw.c:189-221 reads it with fscanf).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libfizz_raster_oracle.so")
        if not os.path.isfile(so):
            subprocess.check_call(["make", "-C", HERE, "libfizz_raster_oracle.so"], stdout=subprocess.DEVNULL)
        _LIB = C.CDLL(so)
        _LIB.fzr_frame.restype = C.c_int
        _LIB.fzr_frame.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _LIB


def parse_plane(text):
    """-> (W, H, nverts[int32], xs[int32], ys[int32]) of one plane text (polygons only, file order)."""
    lines = text.split("\n")
    W, H = (int(v) for v in lines[0].split(","))
    nverts, xs, ys = [], [], []
    i = 1
    while i < len(lines) and lines[i] == "polygon":
        n = int(lines[i + 1].split(",")[1])
        for k in range(n):
            _, x, y = lines[i + 2 + k].split(",")
            xs.append(int(x))
            ys.append(int(y))
        nverts.append(n)
        i += 2 + n
    return W, H, np.array(nverts, dtype=np.int32), np.array(xs, dtype=np.int32), np.array(ys, dtype=np.int32)


def frame(W, H, nverts, xs, ys):
    nverts = np.ascontiguousarray(nverts, dtype=np.int32)
    xs = np.ascontiguousarray(xs, dtype=np.int32)
    ys = np.ascontiguousarray(ys, dtype=np.int32)
    out = np.empty((H, W), dtype=np.uint8)
    rc = lib().fzr_frame(W, H, len(nverts), nverts.ctypes.data, xs.ctypes.data, ys.ctypes.data, out.ctypes.data)
    if rc:
        raise MemoryError("fzr_frame")
    return out


def frame_from_text(text):
    return frame(*parse_plane(text))
