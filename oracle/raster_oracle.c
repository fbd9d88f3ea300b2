/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's rasteriser.
 *
 * Follows /root/reference/src/draw.c: spec_to_image :189-221 (polygons in file order, every polygon filled),
 * draw_polygon :116-157 (border, then flood fill from the truncated mean of the vertices), draw_line :87-114
 * (Bresenham with the two axis-aligned special cases that leave out the last pixel), fill_shape :159-187
 * (4-connected flood fill of the EMPTY component of the seed; restated with an explicit stack -- the set of pixels a
 * flood fill reaches does not depend on the visiting order, and the reference's recursion overflows the C stack on
 * large regions), write_gray_png png_util.c:238-262 (FILLED = 1.0 -> 255, EMPTY = 0.0 -> 0).
 * Circles (draw_circle :41-85) are not restated: physics.py never writes one (Ball.__init__ raises, physics.py:735).
 * Pinned against the unmodified draw.c (oracle/_ref/draw_ref) by tests/test_raster_oracle.py and the fixtures in
 * tests/golden/frames_*.npz. */
#include <stdlib.h>
#include <string.h>

static void px(unsigned char *img, int W, int H, int x, int y) { /* draw_pixel :32-36 */
    if (x >= 0 && x < W && y >= 0 && y < H) img[(size_t)y * W + x] = 255;
}

static void line(unsigned char *img, int W, int H, int x0, int y0, int x1, int y1) { /* draw_line :87-114 */
    if (x0 == x1 && y0 == y1) { px(img, W, H, x0, y0); return; }
    int dx = abs(x1 - x0), sx = x0 < x1 ? 1 : -1;
    int dy = -abs(y1 - y0), sy = y0 < y1 ? 1 : -1;
    int err = dx + dy, e2;
    if (dx == 0) { for (int i = 0; i < abs(dy); i++) px(img, W, H, x0, y0 + sy * i); return; }
    if (dy == 0) { for (int i = 0; i < abs(dx); i++) px(img, W, H, x0 + sx * i, y0); return; }
    for (;;) {
        px(img, W, H, x0, y0);
        if (x0 == x1 && y0 == y1) break;
        e2 = 2 * err;
        if (e2 >= dy) { err += dy; x0 += sx; }
        if (e2 <= dx) { err += dx; y0 += sy; }
    }
}

static void flood(unsigned char *img, int W, int H, int x, int y, int *stack) { /* fill_shape :159-187 */
    if (x < 0 || x >= W || y < 0 || y >= H || img[(size_t)y * W + x]) return;
    size_t top = 0;
    img[(size_t)y * W + x] = 255;
    stack[top++] = y * W + x;
    while (top) {
        const int p = stack[--top], cx = p % W, cy = p / W;
        const int nx[4] = {cx + 1, cx - 1, cx, cx}, ny[4] = {cy, cy, cy + 1, cy - 1};
        for (int k = 0; k < 4; k++)
            if (nx[k] >= 0 && nx[k] < W && ny[k] >= 0 && ny[k] < H && !img[(size_t)ny[k] * W + nx[k]]) {
                img[(size_t)ny[k] * W + nx[k]] = 255;
                stack[top++] = ny[k] * W + nx[k];
            }
    }
}

/* One frame.  nverts[p] vertices per polygon, xs/ys concatenated in file order; out: H rows of W bytes (0 / 255).
 * Returns 0, or -1 when out of memory. */
int fzr_frame(int W, int H, int npoly, const int *nverts, const int *xs, const int *ys, unsigned char *out) {
    int *stack = (int *)malloc(sizeof(int) * (size_t)W * (size_t)H + sizeof(int));
    if (!stack) return -1;
    memset(out, 0, (size_t)W * (size_t)H);
    int o = 0;
    for (int p = 0; p < npoly; p++) {
        const int n = nverts[p];
        for (int i = 1; i <= n; i++)                                             /* draw_polygon :120-131 */
            line(out, W, H, xs[o + i - 1], ys[o + i - 1], xs[o + i % n], ys[o + i % n]);
        float cx = 0, cy = 0;                                                    /* :133-153 */
        for (int i = 0; i < n; i++) {
            int vx = xs[o + i], vy = ys[o + i];
            if (vx < 0) vx = -1;
            if (vx >= W) vx = W;
            if (vy < 0) vy = -1;
            if (vy >= H) vy = H;
            cx = cx + vx;
            cy = cy + vy;
        }
        flood(out, W, H, (int)(cx / n), (int)(cy / n), stack);                   /* :154-156 */
        o += n;
    }
    free(stack);
    return 0;
}
