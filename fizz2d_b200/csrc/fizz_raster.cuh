// fizz_raster.cuh -- batched rasteriser: one frame per CTA, the canvas as a BITMAP in shared
// memory, bytes only touch HBM once, on the way out.
//
// Reference semantics (file:line = /root/reference/src/draw.c unless noted):
//   spec_to_image :189-221   polygons in file order, every one filled
//   draw_polygon  :116-157   border edge by edge, then flood fill from the truncated float mean of the vertices
//   draw_line     :87-114    Bresenham; the two axis-aligned special cases leave out the segment's last pixel
//   fill_shape    :159-187   4-connected flood fill of the EMPTY component that contains the seed (nothing if the seed
//                            is FILLED or off the canvas); earlier polygons act as walls
//   write_gray_png png_util.c:238-262   FILLED (1.0) -> 255, EMPTY (0.0) -> 0
//   Point.__str__ physics.py:600-606    int() truncation of a vertex, clamped to [0,W-1] x [0,H-1]  (points_kernel)
// Circles (:41-85) are not drawn: physics.py never writes one (Ball.__init__ raises, physics.py:735).
//
// The flood fill is the only part that is not embarrassingly parallel.  A row of the canvas is RW <= 32 words, one per
// lane of a warp; the horizontal part of the fill -- "every run of free pixels that contains a reached pixel" -- is a
// multiword addition per direction (free + reached: the carry runs through a run of 1s and stops at the first wall)
// whose carries between words come from a two-ballot carry-lookahead; the vertical part is a scan-line sweep down and a
// sweep up by that warp, repeated only while a sweep opens something behind itself (concave regions).  Cost: about two
// row visits per row of the region, each a few hundred cycles, instead of one visit per pixel.
#pragma once

#include "fizz_b200.h"

#include <cuda_runtime.h>

namespace fizz {

constexpr int RASTER_BLOCK = 256;

struct RasterParams {
    int W, H, RW;                  // canvas; words per row = ceil(W/32)
    int n_poly, n_points;          // polygons per frame; points per frame = sum(nverts)
    int nverts[FIZZ_RASTER_MAX_POLYGONS];
    const int *points;             // [n_frames][n_points][2]  (x, y)
    unsigned char *gray;           // [n_frames][H][W]
};

__device__ __forceinline__ void raster_px(unsigned *wall, int W, int H, int RW, int x, int y) {  // draw_pixel :32-36
    if (x >= 0 && x < W && y >= 0 && y < H) atomicOr(&wall[y * RW + (x >> 5)], 1u << (x & 31));
}

__device__ void raster_line(unsigned *wall, int W, int H, int RW, int x0, int y0, int x1, int y1) {  // draw_line :87-114
    if (x0 == x1 && y0 == y1) { raster_px(wall, W, H, RW, x0, y0); return; }
    const int dx = abs(x1 - x0), sx = x0 < x1 ? 1 : -1;
    const int dy = -abs(y1 - y0), sy = y0 < y1 ? 1 : -1;
    int err = dx + dy;
    if (dx == 0) { for (int i = 0; i < -dy; i++) raster_px(wall, W, H, RW, x0, y0 + sy * i); return; }
    if (dy == 0) { for (int i = 0; i < dx; i++) raster_px(wall, W, H, RW, x0 + sx * i, y0); return; }
    for (;;) {
        raster_px(wall, W, H, RW, x0, y0);
        if (x0 == x1 && y0 == y1) break;
        const int e2 = 2 * err;
        if (e2 >= dy) { err += dy; x0 += sx; }
        if (e2 <= dx) { err += dx; y0 += sy; }
    }
}

// Horizontal part of the fill for one row held by a warp, lane = word of the row (lanes >= RW hold fr = sd = 0):
// returns the free runs that contain a seed bit.  The multiword addition fr + sd lets the carry run through a run of free
// pixels and stop at the first wall; the carries between words come from a carry-lookahead over two ballots
// (generate = the word's own sum overflowed, propagate = the sum is all ones), once per direction.
__device__ __forceinline__ unsigned raster_hfill(unsigned fr, unsigned sd) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // towards higher x
    unsigned s = fr + sd;
    unsigned G = __ballot_sync(FULL, s < fr), P = __ballot_sync(FULL, s == 0xffffffffu);
    unsigned U = G | P;
    unsigned cin = ((U + G) ^ U ^ G);                       // bit j: carry INTO word j
    s += (cin >> lane) & 1u;
    unsigned r = ((s ^ fr) & fr) | sd;
    // towards lower x: the same on the bit-reversed words, carries travelling to lower lanes
    const unsigned frb = __brev(fr), sdb = __brev(r);
    s = frb + sdb;
    G = __brev(__ballot_sync(FULL, s < frb));
    P = __brev(__ballot_sync(FULL, s == 0xffffffffu));
    U = G | P;
    cin = ((U + G) ^ U ^ G);                                // bit 31-j: carry INTO word j (from word j+1)
    s += (cin >> (31 - lane)) & 1u;
    return __brev(((s ^ frb) & frb) | sdb);
}

// Flood fill (fill_shape :159-187) by ONE warp: scan-line sweeps.  A sweep walks the rows in one direction, hands each
// row the reached pixels of the row it came from and fills horizontally; sweeps alternate until neither direction
// reaches anything new.  A convex region is done after one sweep down and one sweep up from the seed row.
__device__ void raster_flood(const unsigned *wall, unsigned *reach, int H, int RW, unsigned last_mask, int sx, int sy) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool mine = lane < RW;
    const unsigned vmask = (lane == RW - 1) ? last_mask : 0xffffffffu;
    if (lane == (sx >> 5)) reach[sy * RW + lane] = 1u << (sx & 31);
    __syncwarp();
    int lo = sy, hi = sy;              // rows that hold reached pixels
    bool again_down = true, again_up = true;
    for (int round = 0; again_down || again_up; round++) {
        const bool down = (round & 1) == 0;
        if (down ? !again_down : !again_up) continue;
        if (down) again_down = false; else again_up = false;
        // the sweep starts at the first reached row on its side and ends when it walks off the canvas or a row beyond
        // the reached band takes nothing over
        int y = down ? lo : hi;
        unsigned prev = 0u;            // reached pixels of the row the sweep comes from
        for (; y >= 0 && y < H; y += down ? 1 : -1) {
            const unsigned fr = mine ? (~wall[y * RW + lane] & vmask) : 0u;
            const unsigned old = mine ? reach[y * RW + lane] : 0u;
            const unsigned sd = (old | prev) & fr;
            const unsigned r = raster_hfill(fr, sd);
            const unsigned grew = __ballot_sync(FULL, r != old);
            if (grew) {
                if (mine) reach[y * RW + lane] = r;
                lo = min(lo, y); hi = max(hi, y);
                // pixels this row gained may open the row the sweep came from: the other direction has to run again
                const int yb = down ? y - 1 : y + 1;
                if (yb >= 0 && yb < H) {
                    const unsigned back = mine ? (r & ~old & ~wall[yb * RW + lane] & ~reach[yb * RW + lane]) : 0u;
                    if (__ballot_sync(FULL, back != 0u)) { if (down) again_up = true; else again_down = true; }
                }
            } else if (!__ballot_sync(FULL, r != 0u) && (down ? y > hi : y < lo)) {
                break;                 // an empty row beyond the band: nothing travels further
            }
            prev = r;
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(RASTER_BLOCK) raster_kernel(const RasterParams rp) {
    extern __shared__ unsigned smem_r[];
    const int W = rp.W, H = rp.H, RW = rp.RW, tid = threadIdx.x;
    unsigned *wall = smem_r, *reach = smem_r + (size_t)H * RW;
    __shared__ int seed[2];
    const int *pts = rp.points + (size_t)blockIdx.x * rp.n_points * 2;
    const unsigned last_mask = (W & 31) ? ((1u << (W & 31)) - 1u) : 0xffffffffu;

    for (int i = tid; i < H * RW; i += RASTER_BLOCK) wall[i] = 0u;
    __syncthreads();
    int o = 0;
    for (int p = 0; p < rp.n_poly; p++) {
        const int n = rp.nverts[p];
        // ---- border (:120-131): one edge per thread -------------------------------------------------------------
        for (int i = 1 + tid; i <= n; i += RASTER_BLOCK) {
            const int a = o + i - 1, b = o + (i % n);
            raster_line(wall, W, H, RW, pts[2 * a], pts[2 * a + 1], pts[2 * b], pts[2 * b + 1]);
        }
        // ---- seed (:133-155): float accumulation in vertex order, float division, truncation -------------------
        if (tid == 0) {
            float cx = 0.f, cy = 0.f;
            for (int i = 0; i < n; i++) {
                int vx = pts[2 * (o + i)], vy = pts[2 * (o + i) + 1];
                if (vx < 0) vx = -1;
                if (vx >= W) vx = W;
                if (vy < 0) vy = -1;
                if (vy >= H) vy = H;
                cx = cx + (float)vx;
                cy = cy + (float)vy;
            }
            seed[0] = (int)(cx / (float)n);
            seed[1] = (int)(cy / (float)n);
        }
        for (int i = tid; i < H * RW; i += RASTER_BLOCK) reach[i] = 0u;
        __syncthreads();
        // ---- flood fill (:159-187) ----------------------------------------------------------------------------
        const int sx = seed[0], sy = seed[1];
        const bool start = sx >= 0 && sx < W && sy >= 0 && sy < H && !((wall[sy * RW + (sx >> 5)] >> (sx & 31)) & 1u);
        if (start) {  // (block-uniform)
            if (tid < 32) raster_flood(wall, reach, H, RW, last_mask, sx, sy);
            __syncthreads();
            for (int i = tid; i < H * RW; i += RASTER_BLOCK) wall[i] |= reach[i];
        }
        __syncthreads();
        o += n;
    }
    // ---- bitmap -> bytes (png_util.c:249): 4 pixels per 32-bit store where the frame allows it ------------------
    unsigned char *out = rp.gray + (size_t)blockIdx.x * W * H;
    const long long npx = (long long)W * H;
    if (((npx & 3) == 0) && ((reinterpret_cast<unsigned long long>(rp.gray) & 3ull) == 0)) {
        unsigned *out4 = reinterpret_cast<unsigned *>(out);
        for (long long q = tid; q < npx / 4; q += RASTER_BLOCK) {
            unsigned v = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const long long i = 4 * q + k;
                const int y = (int)(i / W), x = (int)(i - (long long)y * W);
                if ((wall[y * RW + (x >> 5)] >> (x & 31)) & 1u) v |= 0xffu << (8 * k);
            }
            out4[q] = v;
        }
    } else {
        for (long long i = tid; i < npx; i += RASTER_BLOCK) {
            const int y = (int)(i / W), x = (int)(i - (long long)y * W);
            out[i] = ((wall[y * RW + (x >> 5)] >> (x & 31)) & 1u) ? 255 : 0;
        }
    }
}

// Point.__str__ (physics.py:600-606) for a batch: free vertices as doubles [n_frames][n_free][2] -> int() -> clamp, with
// the (already integer) fixed points appended: one point list per frame in the order of World.__str__ (:291-301).
__global__ void points_kernel(const double *verts, const int *fixed_pts, int *points, long long n_frames, int n_free,
                              int n_fixed, int W, int H) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = n_free + n_fixed;
    if (i >= n_frames * per) return;
    const long long f = i / per;
    const int k = (int)(i - f * per);
    int x, y;
    if (k < n_free) {
        const double vx = verts[(f * n_free + k) * 2], vy = verts[(f * n_free + k) * 2 + 1];
        // int(): truncation toward zero; a non-finite coordinate (Python raises there) lands on the frame edge
        const long long ix = (vx != vx) ? 0 : (vx >= 9e18 ? (long long)9e18 : (vx <= -9e18 ? -(long long)9e18 : (long long)vx));
        const long long iy = (vy != vy) ? 0 : (vy >= 9e18 ? (long long)9e18 : (vy <= -9e18 ? -(long long)9e18 : (long long)vy));
        x = (int)max(min(ix, (long long)W - 1), 0ll);
        y = (int)max(min(iy, (long long)H - 1), 0ll);
    } else {
        x = fixed_pts[2 * (k - n_free)];
        y = fixed_pts[2 * (k - n_free) + 1];
    }
    points[2 * i] = x;
    points[2 * i + 1] = y;
}

}  // namespace fizz
