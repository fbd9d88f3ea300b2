"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/frames_*.npz with the UNMODIFIED reference rasteriser
(/root/reference/src/draw.c, built by `make -C oracle ref` as oracle/_ref/draw_ref with libpng stubbed out: see
oracle/ref_stubs/).  Re-run in the build container with:  python oracle/make_golden_frames.py

  tests/golden/frames_scenes.npz     frames of the reference's own scenes: the `plane_%03d.txt` texts come from the
                                     unmodified physics.py (str(World), through oracle/ref_harness.py), the pixels from
                                     draw_ref run on those files -- i.e. the reference pipeline of simulate.sh end to end
  tests/golden/frames_synthetic.npz  hand-made and random plane texts that exercise the rasteriser's corner cases
                                     (axis-aligned edges that drop their last pixel, seeds outside concave polygons,
                                     borders with gaps that let the fill leak, polygons on the canvas edge, degenerate
                                     polygons, overlapping polygons, odd canvas sizes)
Every frame is stored bit-packed (np.packbits along x) next to the text it was drawn from.
"""
import os
import resource
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DRAW_REF = os.path.join(HERE, "_ref", "draw_ref")


def read_pgm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P5"
        w, h = (int(x) for x in f.readline().split())
        assert f.readline().strip() == b"255"
        return np.frombuffer(f.read(w * h), dtype=np.uint8).reshape(h, w).copy()


def draw_ref(texts):
    """Run the reference binary the way simulate.sh would: ./simulations/plane_%03d.txt -> plane_%03d.png (here: PGM)."""
    frames = []
    with tempfile.TemporaryDirectory() as d:
        os.mkdir(os.path.join(d, "simulations"))
        for lo in range(0, len(texts), 900):           # (the reference formats the index with %03d)
            chunk = texts[lo:lo + 900]
            for i, t in enumerate(chunk):
                with open(os.path.join(d, "simulations", "plane_%03d.txt" % i), "w") as f:
                    f.write(t)
            # fill_shape recurses once per filled pixel: give it the stack it needs
            subprocess.run([DRAW_REF, "0", str(len(chunk))], cwd=d, check=True, stdout=subprocess.DEVNULL,
                           preexec_fn=lambda: resource.setrlimit(resource.RLIMIT_STACK,
                                                                 (resource.RLIM_INFINITY, resource.RLIM_INFINITY)))
            for i in range(len(chunk)):
                frames.append(read_pgm(os.path.join(d, "simulations", "plane_%03d.png" % i)))
    return frames


def pack(texts, frames):
    assert len(texts) == len(frames)
    dims = np.array([f.shape for f in frames], dtype=np.int32)              # (H, W) per frame
    assert all(set(np.unique(f)) <= {0, 255} for f in frames)
    bits = [np.packbits(f > 0, axis=1).ravel() for f in frames]
    offs = np.cumsum([0] + [len(b) for b in bits]).astype(np.int64)
    return dict(texts=np.array(texts), dims=dims, offsets=offs, bits=np.concatenate(bits))


def scene_texts():
    texts, tags = [], []
    for name, steps in (("BOX_1SQUARE", (0, 1, 30, 120)), ("RAMPWORLD_2TRIANGLES_5SQUARES", (0, 40, 150, 300)),
                        ("2CHAMBERS_3BALLS", (0, 60, 200)), ("PENTAGONVOID_1SQUARE", (0, 100)),
                        ("PLANE_2SQUARES", (0, 90)), ("CORRIDOR_2SQUARES", (0, 75)),
                        ("RAMP_1SQUARE_1RECTANGLE", (0, 110)), ("STACKEDCHAMBERS_1SQUARE", (0, 140))):
        w = rh.RefWorld(os.path.join(rh.SCENES, name + ".in"))
        t = 0
        for s in steps:
            while t < s:
                w.update()
                t += 1
            texts.append(w.dump())
            tags.append("%s@%d" % (name, s))
    return texts, tags


def plane_text(W, H, polys):
    out = "%d,%d\n" % (W, H)
    for pts in polys:
        out += "polygon\nsides,%d\n" % len(pts) + "".join("point,%d,%d\n" % (x, y) for x, y in pts)
    return out


def synthetic_texts():
    T = []
    # axis-aligned rectangle: the special cases of draw_line leave out each edge's last pixel (the next edge draws it)
    T.append(plane_text(40, 30, [[(5, 5), (20, 5), (20, 15), (5, 15)]]))
    # the same, traversed the other way round
    T.append(plane_text(40, 30, [[(5, 15), (20, 15), (20, 5), (5, 5)]]))
    # L-shape whose vertex mean lies OUTSIDE the polygon: the fill floods the background instead
    T.append(plane_text(64, 48, [[(4, 4), (60, 4), (60, 10), (10, 10), (10, 44), (4, 44)]]))
    # a second polygon inside an already flooded background: its own fill finds nothing EMPTY
    T.append(plane_text(64, 48, [[(4, 4), (60, 4), (60, 10), (10, 10), (10, 44), (4, 44)],
                                 [(30, 20), (50, 20), (50, 40), (30, 40)]]))
    # the order reversed: the square is filled first and stays visible inside the flooded background
    T.append(plane_text(64, 48, [[(30, 20), (50, 20), (50, 40), (30, 40)],
                                 [(4, 4), (60, 4), (60, 10), (10, 10), (10, 44), (4, 44)]]))
    # degenerate polygons: a point, a horizontal segment, a vertical segment, a diagonal there-and-back
    T.append(plane_text(33, 17, [[(7, 7), (7, 7), (7, 7)], [(10, 3), (20, 3), (15, 3)], [(25, 2), (25, 12), (25, 6)],
                                 [(2, 10), (8, 16), (2, 10)]]))
    # polygons on the canvas edge / covering the whole canvas
    T.append(plane_text(50, 20, [[(0, 0), (49, 0), (49, 19), (0, 19)]]))
    T.append(plane_text(50, 20, [[(0, 0), (10, 0), (0, 10)], [(49, 19), (40, 19), (49, 12)], [(49, 0), (49, 5), (44, 0)]]))
    # thin sliver: the border pixels leave no EMPTY pixel at the seed
    T.append(plane_text(80, 20, [[(2, 2), (70, 4), (70, 5)]]))
    # overlapping polygons, steep and shallow edges
    T.append(plane_text(97, 61, [[(10, 10), (60, 12), (55, 50), (12, 45)], [(40, 5), (90, 30), (45, 58)],
                                 [(70, 40), (95, 40), (95, 59), (70, 59)]]))
    # star (concave, seed inside)
    T.append(plane_text(120, 120, [[(60, 5), (72, 45), (115, 45), (80, 70), (95, 112), (60, 86), (25, 112), (40, 70),
                                    (5, 45), (48, 45)]]))
    # a ring of two squares: the inner one is drawn first, the outer one's seed is inside the inner one (already filled)
    T.append(plane_text(60, 60, [[(20, 20), (40, 20), (40, 40), (20, 40)], [(5, 5), (55, 5), (55, 55), (5, 55)]]))
    # 1-pixel canvas rows/columns and a wide canvas (more than 16 words per row)
    T.append(plane_text(1, 1, [[(0, 0), (0, 0), (0, 0)]]))
    T.append(plane_text(700, 9, [[(3, 1), (690, 2), (650, 7), (20, 6)]]))
    T.append(plane_text(9, 400, [[(1, 3), (7, 20), (6, 390), (2, 380)]]))
    # random polygons (not necessarily simple), several per frame, odd canvas sizes
    rng = np.random.default_rng(20260401)
    for k in range(36):
        W = int(rng.choice([31, 32, 33, 64, 100, 257, 500, 640]))
        H = int(rng.choice([17, 32, 75, 300, 480]))
        polys = []
        for _ in range(int(rng.integers(1, 7))):
            n = int(rng.integers(3, 9))
            if rng.random() < 0.5:      # convex-ish: points on an ellipse, sorted by angle
                cx, cy = rng.uniform(0, W), rng.uniform(0, H)
                rx, ry = rng.uniform(2, W / 2), rng.uniform(2, H / 2)
                ang = np.sort(rng.uniform(0, 2 * np.pi, n))
                pts = [(int(np.clip(cx + rx * np.cos(a), 0, W - 1)), int(np.clip(cy + ry * np.sin(a), 0, H - 1))) for a in ang]
            else:                       # arbitrary
                pts = [(int(rng.integers(0, W)), int(rng.integers(0, H))) for _ in range(n)]
            polys.append(pts)
        T.append(plane_text(W, H, polys))
    return T


if __name__ == "__main__":
    if not (rh.available() and os.path.isfile(DRAW_REF)):
        sys.exit("needs /root/reference and oracle/_ref/draw_ref (make -C oracle ref)")
    texts, tags = scene_texts()
    frames = draw_ref(texts)
    np.savez_compressed(os.path.join(OUT, "frames_scenes.npz"), tags=np.array(tags), **pack(texts, frames))
    print("frames_scenes.npz:", len(frames), "frames", sorted({f.shape for f in frames}))
    texts = synthetic_texts()
    frames = draw_ref(texts)
    np.savez_compressed(os.path.join(OUT, "frames_synthetic.npz"), **pack(texts, frames))
    print("frames_synthetic.npz:", len(frames), "frames; filled fraction %.3f .. %.3f" % (
        min((f > 0).mean() for f in frames), max((f > 0).mean() for f in frames)))
    for p in ("frames_scenes.npz", "frames_synthetic.npz"):
        print(p, os.path.getsize(os.path.join(OUT, p)), "bytes")
