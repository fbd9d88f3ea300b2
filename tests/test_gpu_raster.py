"""GPU: the batched rasteriser (fizz_raster_frames through the C-ABI) against frames drawn by the UNMODIFIED reference
draw.c (tests/golden/frames_*.npz) and against the CPU restatement (oracle/raster_oracle.c) on seeded random input;
bit-exact (every pixel)."""
import os

import numpy as np
import pytest

import fizz2d_b200 as fz
from fizz2d_b200 import raster
from conftest import scene_file
from helpers import ROOT  # noqa: F401
import raster_oracle as ro
from test_raster_oracle import golden_frames

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["frames_scenes.npz", "frames_synthetic.npz"])
def test_frames_equal_reference(name):
    items = list(golden_frames(name))
    got = raster.render_texts([t for _, t, _ in items])
    for (i, _, want), g in zip(items, got):
        assert g.shape == want.shape and np.array_equal(g, want), (name, i, int((g != want).sum()))


def test_random_batches_vs_oracle():
    """Batches that share a topology (the shape of a time series or of a population): 64 frames per launch, random
    polygons that move from frame to frame, several canvas sizes including non-multiples of 4 and of 32."""
    rng = np.random.default_rng(7)
    for W, H, nverts in ((500, 300, (4, 4, 3, 5)), (97, 61, (3, 3, 8)), (33, 17, (4,)), (1010, 710, (4, 6, 3)),
                         (64, 64, (5, 5, 5, 5, 5, 5, 5, 5))):
        n = 64 if W * H < 400000 else 12
        pts = np.empty((n, sum(nverts), 2), dtype=np.int32)
        pts[:, :, 0] = rng.integers(0, W, size=pts.shape[:2])
        pts[:, :, 1] = rng.integers(0, H, size=pts.shape[:2])
        o = 0
        for nv in nverts[::2]:            # every other polygon convex: sorted by angle around its centroid
            p = pts[:, o:o + nv].astype(np.float64)
            c = p.mean(axis=1, keepdims=True)
            order = np.argsort(np.arctan2(p[:, :, 1] - c[:, :, 1], p[:, :, 0] - c[:, :, 0]), axis=1)
            pts[:, o:o + nv] = np.take_along_axis(pts[:, o:o + nv], order[:, :, None], axis=1)
            o += nv
        got = raster.render_points(W, H, nverts, pts).cpu().numpy()
        for f in range(n):
            want = ro.frame(W, H, nverts, pts[f, :, 0], pts[f, :, 1])
            assert np.array_equal(got[f], want), (W, H, f, int((got[f] != want).sum()))


def test_render_from_resident_state_equals_text_pipeline():
    """BatchedWorld.render (device-side int() + clamp + raster) == rasterising the text dump of the same worlds, which
    is byte-identical to the reference's plane_%03d.txt (test_gpu_compat)."""
    bw = fz.BatchedWorld.from_in(scene_file("RAMPWORLD_2TRIANGLES_5SQUARES"), n_worlds=40, perturb=True, seed=11)
    for steps in (0, 90, 200):
        if steps:
            bw.step(steps)
        sel = [0, 3, 17, 39]
        fr = bw.render(sel).cpu().numpy()
        verts = bw.vertices()
        for k, w in enumerate(sel):
            want = ro.frame_from_text(bw.dump_world(w, verts))
            assert np.array_equal(fr[k], want), (steps, w)
    assert fr.shape == (4, 710, 1010) and 0 < (fr > 0).mean() < 1


def test_png_round_trip_and_draw_cli(tmp_path, monkeypatch):
    """`python -m fizz2d_b200.raster START END` reads ./simulations/plane_%03d.txt and writes plane_%03d.png like the
    reference's ./draw; the PNG is 8-bit RGB with the frame in every channel."""
    items = list(golden_frames("frames_scenes.npz"))[:3]
    monkeypatch.chdir(tmp_path)
    os.mkdir("simulations")
    for i, (_, text, _) in enumerate(items):
        with open("simulations/plane_%03d.txt" % i, "w") as f:
            f.write(text)
    raster.main(["0", "3"])
    for i, (_, _, want) in enumerate(items):
        data = open("simulations/plane_%03d.png" % i, "rb").read()
        assert np.array_equal(raster.read_png_gray(data), want)


def test_raster_rejects_what_it_cannot_draw():
    with pytest.raises(fz.FizzError):
        raster.render_points(4096, 4096, (3,), np.zeros((1, 3, 2), dtype=np.int32))     # bitmaps exceed shared memory
    with pytest.raises(ValueError):
        raster.parse_plane("10,10\ncircle\n5,5,3\n")
    assert raster.render_points(16, 8, (), np.zeros((2, 0, 2), dtype=np.int32)).sum().item() == 0   # empty frames
