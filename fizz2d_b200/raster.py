"""Batched rasteriser: the host side of `fizz_raster_frames` (include/fizz_b200.h) -- a drop-in for the reference's
only native program, `./draw START END` (/root/reference/src/draw.c), and a renderer for worlds that are resident on
the GPU (BatchedWorld.render).  Pixels come from the CUDA kernel only; there is no CPU fallback.

Reference behaviour mirrored here (file:line = /root/reference/src/...):
  draw.c:229-247     main(): for every index in [START, END) read ./simulations/plane_%03d.txt, write plane_%03d.png
  draw.c:189-221     spec_to_image(): "W,H" header, then "polygon / sides,N / point,x,y * N" blocks; reading stops at the
                     first line that is neither "polygon" nor "circle"
  png_util.c:238-262 write_gray_png(): intensity 0 / 255 in all three channels of an 8-bit RGB PNG
"""
from __future__ import annotations

import ctypes as C
import struct
import sys
import zlib

import numpy as np

from . import lib as _lib

SIMULATION_DIR = "./simulations/"


def parse_plane(text):
    """-> (W, H, nverts tuple, points int32 [n_points, 2]) of one plane_%03d.txt (draw.c:189-221)."""
    lines = text.split("\n")
    W, H = (int(v) for v in lines[0].split(","))
    nverts, pts = [], []
    i = 1
    while i < len(lines) and lines[i] == "polygon":
        n = int(lines[i + 1].split(",")[1])
        for k in range(n):
            _, x, y = lines[i + 2 + k].split(",")
            pts.append((int(x), int(y)))
        nverts.append(n)
        i += 2 + n
    if i < len(lines) and lines[i] == "circle":
        raise ValueError("circles are not supported (the reference's physics.py never writes one)")
    return W, H, tuple(nverts), np.array(pts, dtype=np.int32).reshape(-1, 2)


def render_points(width, height, nverts, points, device=0, stream=0):
    """points: int32 [n_frames, n_points, 2] (numpy, or a torch tensor already on `device`) -> torch.uint8
    [n_frames, height, width] on the device (255 = FILLED)."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.FizzError("no CUDA device: the rasteriser has no CPU fallback")
    nv = np.ascontiguousarray(nverts, dtype=np.int32)
    if len(nv) > _lib.RASTER_MAX_POLYGONS:
        raise ValueError("at most %d polygons per frame" % _lib.RASTER_MAX_POLYGONS)
    dev = torch.device("cuda:%d" % device)
    if isinstance(points, np.ndarray):
        points = torch.from_numpy(np.ascontiguousarray(points, dtype=np.int32)).to(dev)
    points = points.to(device=dev, dtype=torch.int32).contiguous()
    if points.dim() != 3 or points.shape[1] != int(nv.sum()) or points.shape[2] != 2:
        raise ValueError("points must be [n_frames, sum(nverts), 2]")
    n = points.shape[0]
    out = torch.empty((n, height, width), dtype=torch.uint8, device=dev)
    _lib.check(_lib.load().fizz_raster_frames(device, width, height, n, len(nv), nv.ctypes.data_as(_lib.ip),
                                              C.c_void_p(points.data_ptr()), C.c_void_p(out.data_ptr()),
                                              C.c_void_p(stream)))
    return out


def points_from_vertices(width, height, verts, fixed_points, device=0, stream=0):
    """Point.__str__ on the device: verts = torch float64 [n_frames, n_free_points, 2] on `device`; fixed_points = int32
    [n_fixed_points, 2] (numpy) -> torch.int32 [n_frames, n_free_points + n_fixed_points, 2]."""
    import torch
    dev = torch.device("cuda:%d" % device)
    verts = verts.to(device=dev, dtype=torch.float64).contiguous()
    fp = torch.from_numpy(np.ascontiguousarray(fixed_points, dtype=np.int32).reshape(-1, 2)).to(dev)
    n, nf, nx = verts.shape[0], verts.shape[1], fp.shape[0]
    out = torch.empty((n, nf + nx, 2), dtype=torch.int32, device=dev)
    _lib.check(_lib.load().fizz_points_from_vertices(device, width, height, n, nf, C.c_void_p(verts.data_ptr()), nx,
                                                     C.c_void_p(fp.data_ptr() if nx else 0), C.c_void_p(out.data_ptr()),
                                                     C.c_void_p(stream)))
    return out


def render_texts(texts, device=0):
    """Frames of a list of plane texts, as numpy uint8 [H, W] arrays in input order.  Texts that share canvas and
    topology are rasterised in one launch."""
    import torch
    parsed = [parse_plane(t) for t in texts]
    groups = {}
    for i, (W, H, nv, pts) in enumerate(parsed):
        groups.setdefault((W, H, nv), []).append(i)
    out = [None] * len(texts)
    for (W, H, nv), idx in groups.items():
        pts = np.stack([parsed[i][3] for i in idx]) if sum(nv) else np.zeros((len(idx), 0, 2), dtype=np.int32)
        fr = render_points(W, H, nv, pts, device=device)
        torch.cuda.synchronize(fr.device)
        fr = fr.cpu().numpy()
        for k, i in enumerate(idx):
            out[i] = fr[k]
    return out


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)


def png_bytes(gray):
    """8-bit RGB PNG with the intensity in all three channels (png_util.c:238-262 -> write_png with alpha == NULL)."""
    g = np.ascontiguousarray(gray, dtype=np.uint8)
    h, w = g.shape
    rows = np.empty((h, 1 + 3 * w), dtype=np.uint8)
    rows[:, 0] = 0                                  # filter type 0 on every scan line
    rows[:, 1:] = np.repeat(g, 3, axis=1)
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
            _chunk(b"IDAT", zlib.compress(rows.tobytes(), 6)) + _chunk(b"IEND", b""))


def write_png(path, gray):
    with open(path, "wb") as f:
        f.write(png_bytes(gray))


def read_png_gray(data):
    """Decoder for the PNGs this module writes (8-bit RGB, no interlace, filter 0): -> uint8 [H, W] of channel 0.
    Used by the tests and by tools; not a general PNG reader."""
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    o, idat, w, h = 8, b"", 0, 0
    while o < len(data):
        n, tag = struct.unpack(">I", data[o:o + 4])[0], data[o + 4:o + 8]
        body = data[o + 8:o + 8 + n]
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert (depth, ctype) == (8, 2)
        elif tag == b"IDAT":
            idat += body
        o += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1::3].copy()


def main(argv=None):
    """`python -m fizz2d_b200.raster START END`: the reference's `./draw START END` (draw.c:229-247)."""
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        sys.exit("usage: python -m fizz2d_b200.raster START END")
    start, end = int(argv[0]) & 0xffff, int(argv[1]) & 0xffff       # uint16_t counters, draw.c:230-231
    texts = []
    for ind in range(start, end):
        with open(SIMULATION_DIR + "plane_%03d.txt" % ind) as f:
            texts.append(f.read())
    for ind, fr in zip(range(start, end), render_texts(texts)):
        write_png(SIMULATION_DIR + "plane_%03d.png" % ind, fr)
    print("processed images %03d through %03d" % (start, end))


if __name__ == "__main__":
    main()
