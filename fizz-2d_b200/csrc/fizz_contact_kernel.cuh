// fizz_contact_kernel.cuh -- warp-per-world kernel with the FULL update() semantics ("hybrid" mode, stage 2).
//
// One WARP per world, persistent: warps pull world indices from a device-side work list (worlds the ballistic
// kernel stopped at a non-empty broad phase, plus step 0 of every world).  The world's state (com, snapshot,
// polar offsets, vertices, contact list) lives in the warp's slice of shared memory for the whole launch; the
// fixed scene is staged once per CTA.  Each check_collisions() (physics.py:262-289) is executed cooperatively:
//   * broad phase: lanes stride over the pair table, __ballot compacts the survivors in reference pair order;
//   * narrow phase: lanes map to (pair, axis) items of polypoly_collision (:765-840) -- one SAT axis per lane,
//     several pairs per 32-lane round -- followed by segmented __shfl reductions ("any separating axis",
//     first strict minimum overlap in axis order) and a __ballot-compacted contact list;
//   * resolution (:142-233) walks the list in order; vertices are pushed lane-parallel, the centre of mass by
//     lane 0.
// The sequential depth of a contact step (30 halving trials + fallback + N resolve passes) is
// what bounds a pass over a batch; this kernel exists to make each of those rounds short.
#pragma once

#include "fizz_common.cuh"

namespace fizz {

constexpr int CONTACT_WARPS = 4;  // warps per CTA
constexpr int CONTACT_BLOCK = CONTACT_WARPS * 32;
constexpr int SP_CAP = 160;       // survivor list capacity (>= MAX_PAIRS)

struct ContactSmem {
    // CTA-shared
    double *sfix, *fcx, *fcy, *fthr;
    int *nv, *vs;
    unsigned char *vbody, *pair_i, *pair_j;
    // per-warp
    double *px, *py, *vx, *vy, *sx, *sy, *thr, *mass, *npx, *npy, *off, *vert, *hnx, *hny, *hmag;
    unsigned char *hi, *hj, *spi, *spj;
};

__host__ __device__ inline size_t contact_smem_bytes(int V, int FV) {
    size_t cta_d = (size_t)FV * 6 + 3 * FIZZ_MAX_FIXED;
    size_t warp_d = 80 + 4 * (size_t)V + 96;
    size_t bytes = 8 * (cta_d + CONTACT_WARPS * warp_d);
    bytes += 4 * 64;                                   // nv, vs
    bytes += ((size_t)V + 15) / 16 * 16 + 2 * SP_CAP;  // vbody, pair tables
    bytes += CONTACT_WARPS * (64 + 2 * SP_CAP);        // hi, hj, spi, spj
    return (bytes + 15) / 16 * 16;
}

__device__ __forceinline__ ContactSmem contact_carve(double *base, int V, int FV, int warp) {
    ContactSmem s;
    double *d = base;
    s.sfix = d; d += FV * 6;
    s.fcx = d; d += FIZZ_MAX_FIXED;
    s.fcy = d; d += FIZZ_MAX_FIXED;
    s.fthr = d; d += FIZZ_MAX_FIXED;
    const int warp_d = 80 + 4 * V + 96;
    double *wd = d + (size_t)warp * warp_d;
    d += (size_t)CONTACT_WARPS * warp_d;
    s.px = wd; s.py = wd + 8; s.vx = wd + 16; s.vy = wd + 24; s.sx = wd + 32; s.sy = wd + 40;
    s.thr = wd + 48; s.mass = wd + 56; s.npx = wd + 64; s.npy = wd + 72;
    s.off = wd + 80; s.vert = s.off + 2 * V;
    s.hnx = s.vert + 2 * V; s.hny = s.hnx + 32; s.hmag = s.hny + 32;
    int *ip = reinterpret_cast<int *>(d);
    s.nv = ip; s.vs = ip + 32;
    unsigned char *b = reinterpret_cast<unsigned char *>(ip + 64);
    s.vbody = b; b += (V + 15) / 16 * 16;
    s.pair_i = b; b += SP_CAP;
    s.pair_j = b; b += SP_CAP;
    unsigned char *wb = b + (size_t)warp * (64 + 2 * SP_CAP);
    s.hi = wb; s.hj = wb + 32; s.spi = wb + 64; s.spj = wb + 64 + SP_CAP;
    return s;
}

// One SAT axis of polypoly_collision (:787-830) for pair (ii free, jj free-or-fixed): axis index k in
// [0, n1+n2) in reference order (poly1's edges, then poly2's).  Returns the overlap on this axis (NaN-safe)
// and whether the axis separates the polygons.
__device__ __forceinline__ void sat_axis(const ContactSmem &s, int F, int ii, int jj, int k, double &nx, double &ny,
                                         double &cur, bool &sep) {
    const int n1 = s.nv[ii], v1 = s.vs[ii];
    const int n2 = s.nv[jj], v2 = s.vs[jj];
    const bool fx2 = jj >= F;
    const double *vert = s.vert;
    double lo2 = 0.0, hi2 = 0.0;
    bool own2 = false;
    if (k < n1) {
        const int k2 = (k + 1 == n1) ? 0 : k + 1;
        const double ax = vert[2 * (v1 + k2) + 1] - vert[2 * (v1 + k) + 1];  // :775 (p2.y - p1.y, p1.x - p2.x)
        const double ay = vert[2 * (v1 + k)] - vert[2 * (v1 + k2)];
        const double nrm = sqrt(fma(ay, ay, ax * ax));                       // :790
        nx = ax / nrm; ny = ay / nrm;
    } else if (fx2) {
        const double *g = s.sfix + (v2 + (k - n1)) * 6;                      // precomputed, bit-identical to :778,:790
        nx = g[2]; ny = g[3]; lo2 = g[4]; hi2 = g[5];
        own2 = true;
    } else {
        const int e = k - n1;
        const int e2 = (e + 1 == n2) ? 0 : e + 1;
        const double ax = vert[2 * (v2 + e2) + 1] - vert[2 * (v2 + e) + 1];  // :778
        const double ay = vert[2 * (v2 + e)] - vert[2 * (v2 + e2)];
        const double nrm = sqrt(fma(ay, ay, ax * ax));
        nx = ax / nrm; ny = ay / nrm;
    }
    double lo1, hi1;
    lo1 = hi1 = dot2(nx, ny, vert[2 * v1], vert[2 * v1 + 1]);                // :795
    for (int v = 1; v < n1; v++) {                                           // :799-804
        const double c = dot2(nx, ny, vert[2 * (v1 + v)], vert[2 * (v1 + v) + 1]);
        if (c < lo1) lo1 = c;
        if (c > hi1) hi1 = c;
    }
    if (!own2) {
        if (fx2) {
            const double *g = s.sfix + v2 * 6;
            lo2 = hi2 = dot2(nx, ny, g[0], g[1]);                            // :796
            for (int v = 1; v < n2; v++) {                                   // :806-811
                const double c = dot2(nx, ny, g[v * 6], g[v * 6 + 1]);
                if (c < lo2) lo2 = c;
                if (c > hi2) hi2 = c;
            }
        } else {
            lo2 = hi2 = dot2(nx, ny, vert[2 * v2], vert[2 * v2 + 1]);
            for (int v = 1; v < n2; v++) {
                const double c = dot2(nx, ny, vert[2 * (v2 + v)], vert[2 * (v2 + v) + 1]);
                if (c < lo2) lo2 = c;
                if (c > hi2) hi2 = c;
            }
        }
    }
    if (lo1 > lo2) {                                                         // :815-817
        double t = lo1; lo1 = lo2; lo2 = t;
        t = hi1; hi1 = hi2; hi2 = t;
    }
    sep = hi1 < lo2;                                                         // :820
    cur = hi1 - lo2;                                                         // :824
}

__global__ void __launch_bounds__(CONTACT_BLOCK, 4)
contact_kernel(const __grid_constant__ KParams kp, const long long *__restrict__ list,
               const unsigned long long *__restrict__ count_ptr, unsigned long long *cursor) {
    extern __shared__ double smem_c[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int F = kp.F, X = kp.X, V = kp.V;
    const ContactSmem s = contact_carve(smem_c, V, kp.FV, warp);
    const int P = F * (F - 1) / 2 + F * X;

    // ---- stage the scene once per CTA ----------------------------------------------------------------------
    for (int i = tid; i < kp.FV * 6; i += CONTACT_BLOCK) s.sfix[i] = kp.fixed_geom[i];
    if (tid < FIZZ_MAX_FIXED) { s.fcx[tid] = kp.fcx[tid]; s.fcy[tid] = kp.fcy[tid]; s.fthr[tid] = kp.fthr[tid]; }
    if (tid < F + X) { s.nv[tid] = kp.nv[tid]; s.vs[tid] = kp.vs[tid]; }
    if (tid == 0) {
        int p = 0;
        for (int i = 0; i < F; i++)
            for (int j = i + 1; j < F + X; j++) { s.pair_i[p] = (unsigned char)i; s.pair_j[p] = (unsigned char)j; p++; }
        for (int b = 0; b < F; b++)
            for (int v = 0; v < kp.nv[b]; v++) s.vbody[kp.vs[b] + v] = (unsigned char)b;
    }
    __syncthreads();

    const long long Wp = kp.Wp;
    const unsigned long long count = *count_ptr;

    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(cursor, 1ULL);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= count) break;
        const long long w = list[idx];

        // ---- load the world into the warp's shared-memory slice ------------------------------------------------
        if (lane < F) {
            const double x = kp.pos[(2 * lane) * Wp + w], y = kp.pos[(2 * lane + 1) * Wp + w];
            s.px[lane] = x; s.py[lane] = y; s.sx[lane] = x; s.sy[lane] = y;  // finish_update ended the last step
            s.vx[lane] = kp.vel[(2 * lane) * Wp + w];
            s.vy[lane] = kp.vel[(2 * lane + 1) * Wp + w];
            s.thr[lane] = kp.thr[lane * Wp + w];
            s.mass[lane] = kp.mass[lane * Wp + w];
        }
        for (int r = lane; r < 2 * V; r += 32) s.off[r] = kp.off[(size_t)r * Wp + w];
        double heat = kp.heat[w];
        long long step = kp.steps[w];
        long long status = 0;
        long long ncalls_total = 0;
        __syncwarp();

        bool resolve = false, verts_explicit = false;
        int k = 0, ncalls = 0, ns = 0;
        double dt = kp.time_disc;

        while (true) {
            const bool acc_on = step > 0;  // acceleration of every com at the start of this update() (:494)
            const double a0x = acc_on ? kp.accx : 0.0, a0y = acc_on ? kp.accy : 0.0;
            bool need_mask = resolve;
            double tvx = 0.0, tvy = 0.0;   // lane b < F: velocity of body b after this trial's move
            if (!resolve) {
                if (lane < F) {
                    double ox, oy;
                    FIZZ_MOVE_POINT(s.px[lane], s.py[lane], s.vx[lane], s.vy[lane], dt, a0x, a0y, kp.accx, kp.accy, ox, oy,
                                    tvx, tvy);
                    s.npx[lane] = ox; s.npy[lane] = oy;
                    if (k == 0 && step == 0) { s.sx[lane] = ox; s.sy[lane] = oy; }  // obj.pos aliases com.pos (Q4)
                }
                if (k == 0) need_mask = true;
                __syncwarp();
            } else if (ncalls >= kp.max_calls) {
                status = FIZZ_ST_CAPPED;  // watchdog
                break;
            }

            // ---- broad phase (:272-281); survivors compacted in pair order ---------------------------------------
            if (need_mask) {
                ns = 0;
                for (int p0 = 0; p0 < P; p0 += 32) {
                    const int p = p0 + lane;
                    bool surv = false;
                    int i = 0, j = 0;
                    if (p < P) {
                        i = s.pair_i[p]; j = s.pair_j[p];
                        double ox, oy, ot;
                        if (j < F) { ox = s.sx[j]; oy = s.sy[j]; ot = s.thr[j]; }
                        else { ox = s.fcx[j - F]; oy = s.fcy[j - F]; ot = s.fthr[j - F]; }
                        const double dx = s.sx[i] - ox, dy = s.sy[i] - oy;
                        const double d2 = fma(dy, dy, dx * dx);
                        surv = !(d2 > s.thr[i] && d2 > ot);
                    }
                    const unsigned b = __ballot_sync(FULL, surv);
                    if (surv) {
                        const int slot = ns + __popc(b & lt_mask);
                        s.spi[slot] = (unsigned char)i; s.spj[slot] = (unsigned char)j;
                    }
                    ns += __popc(b);
                }
                __syncwarp();
            }
            ncalls++;
            ncalls_total++;

            int nh = 0;
            double maxov = 0.0;  // :125-129
            if (ns) {
                if (!resolve) {
                    // pre_update rebuilt every vertex from the moved com (:358-359)
                    for (int v = lane; v < V; v += 32) {
                        const int b = s.vbody[v];
                        s.vert[2 * v] = s.off[2 * v] + s.npx[b];
                        s.vert[2 * v + 1] = s.off[2 * v + 1] + s.npy[b];
                    }
                    __syncwarp();
                }
                // ---- narrow phase: (pair, axis) items, several pairs per 32-lane round --------------------------
                int p0 = 0;
                while (p0 < ns) {
                    int acc = 0, p = p0, myp = -1, myk = 0, seg0 = 0, segn = 1;
                    while (p < ns) {
                        const int na = s.nv[s.spi[p]] + s.nv[s.spj[p]];
                        if (acc + na > 32) break;
                        if (lane >= acc && lane < acc + na) { myp = p; myk = lane - acc; seg0 = acc; segn = na; }
                        acc += na;
                        p++;
                    }
                    double nx = 0.0, ny = 0.0, cur = INFINITY;
                    bool sep = false;
                    int ii = 0, jj = 0;
                    if (myp >= 0) {
                        ii = s.spi[myp]; jj = s.spj[myp];
                        sat_axis(s, F, ii, jj, myk, nx, ny, cur, sep);
                    }
                    const unsigned sepb = __ballot_sync(FULL, sep);
                    const unsigned segmask = (segn >= 32 ? FULL : ((1u << segn) - 1u)) << seg0;
                    const bool pair_sep = (sepb & segmask) != 0;
                    // first strict minimum in axis order (:827-830) = lexicographic min of (overlap, axis)
                    const bool valid = (myp >= 0) && (cur < INFINITY);
                    double bov = valid ? cur : INFINITY;
                    int bk = valid ? myk : 99;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const double oov = __shfl_down_sync(FULL, bov, o);
                        const int ok = __shfl_down_sync(FULL, bk, o);
                        if (myp >= 0 && lane + o < seg0 + segn) {
                            if (oov < bov || (oov == bov && ok < bk)) { bov = oov; bk = ok; }
                        }
                    }
                    const int src = (seg0 + (bk == 99 ? 0 : bk)) & 31;
                    const double fxn = __shfl_sync(FULL, nx, src), fyn = __shfl_sync(FULL, ny, src);
                    const bool hit = (myp >= 0) && (lane == seg0) && !pair_sep && (bk != 99);  // :286-288
                    const unsigned hitb = __ballot_sync(FULL, hit);
                    double am = -1.0;
                    if (hit) {
                        const int slot = nh + __popc(hitb & lt_mask);
                        const double hx = fxn * -1.0, hy = fyn * -1.0;  // :838-839 always negated (Q7)
                        if (slot < FIZZ_MAX_HITS) {
                            s.hi[slot] = (unsigned char)ii; s.hj[slot] = (unsigned char)jj;
                            s.hnx[slot] = hx; s.hny[slot] = hy; s.hmag[slot] = bov;
                        }
                        am = fabs(bov);
                        if (kp.trace) trace_hit(kp, w, step, ncalls, ii, jj, hx, hy, bov);
                    }
                    nh += __popc(hitb);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double other = __shfl_xor_sync(FULL, am, o);
                        if (other > am) am = other;
                    }
                    if (am > maxov) maxov = am;
                    p0 = p;
                }
                __syncwarp();
                if (nh > FIZZ_MAX_HITS) { status = FIZZ_ST_HIT_OVERFLOW; break; }
            }

            if (!resolve) {
                if (maxov > kp.coll_tol && dt > kp.time_tol) {  // :99
                    dt = dt / 2;                                // :106
                    k++;
                    continue;
                }
                if (dt < kp.time_tol) {
                    // :133-137 fallback: com reverts, vertices keep the last trial's positions (Q5) -> same list again
                    ncalls++;
                    ncalls_total++;
                    if (kp.trace && lane == 0)
                        for (int h = 0; h < nh; h++)
                            trace_hit(kp, w, step, ncalls, s.hi[h], s.hj[h], s.hnx[h], s.hny[h], s.hmag[h]);
                } else if (lane < F) {
                    s.px[lane] = s.npx[lane]; s.py[lane] = s.npy[lane]; s.vx[lane] = tvx; s.vy[lane] = tvy;
                }
                if (nh) verts_explicit = true;
                __syncwarp();
            }

            if (nh) {
                // ---- one resolve pass over the list computed above (:142-233, Q11) -----------------------------
                for (int h = 0; h < nh; h++) {
                    const int ii = s.hi[h], jj = s.hj[h];
                    const double ux = s.hnx[h], uy = s.hny[h];
                    const double sc = s.hmag[h] + kp.eps;      // :158
                    const double nvx = ux * sc, nvy = uy * sc;
                    const int n1 = s.nv[ii], v1 = s.vs[ii];
                    if (jj >= F) {                             // :171-198 free body vs fixed body
                        if (lane < n1) { s.vert[2 * (v1 + lane)] += nvx; s.vert[2 * (v1 + lane) + 1] += nvy; }  // :174
                        const bool gate = sqrt(fma(nvy, nvy, nvx * nvx)) > kp.vib_tol;  // :180
                        double cpx = s.px[ii], cpy = s.py[ii], cvx = s.vx[ii], cvy = s.vy[ii];
                        const double m1 = s.mass[ii];
                        if (gate) { cpx += nvx; cpy += nvy; }                           // :181
                        const double vdotN = dot2(cvx, cvy, ux, uy);                    // :183
                        const double cc = kp.ope * vdotN;                               // :188
                        cvx -= cc * ux;
                        cvy -= cc * uy;
                        heat += .5 * m1 * kp.ome2 * (vdotN * vdotN);                    // :191,:195
                        __syncwarp();
                        if (lane == 0) {
                            s.px[ii] = cpx; s.py[ii] = cpy; s.vx[ii] = cvx; s.vy[ii] = cvy;
                            s.sx[ii] = cpx; s.sy[ii] = cpy;                             // :197
                        }
                    } else {                                   // :201-223 two free bodies
                        const int n2 = s.nv[jj], v2 = s.vs[jj];
                        const double m1 = s.mass[ii], m2 = s.mass[jj];
                        const double mtotal = m1 + m2;
                        const double mprop1 = m1 / mtotal;
                        const double mprop2 = 1 - mprop1;
                        double c1x = s.px[ii], c1y = s.py[ii], v1x = s.vx[ii], v1y = s.vy[ii];
                        double c2x = s.px[jj], c2y = s.py[jj], v2x = s.vx[jj], v2y = s.vy[jj];
                        const double vn1 = dot2(v1x, v1y, ux, uy);                      // :207
                        const double vn2 = dot2(v2x, v2y, ux, uy);                      // :208
                        const double v1u = (mprop1 - mprop2) * vn1 + 2 * mprop2 * vn2;  // :209
                        const double v2u = 2 * mprop1 * vn1 + (mprop2 - mprop1) * vn2;  // :218
                        const double d1 = v1u - vn1, d2 = v2u - vn2;                    // :210,:219
                        const double p1x = mprop1 * nvx, p1y = mprop1 * nvy;
                        const double p2x = mprop2 * nvx, p2y = mprop2 * nvy;
                        c1x += p1x; c1y += p1y; v1x += d1 * ux; v1y += d1 * uy;         // :206,:210
                        c2x -= p2x; c2y -= p2y; v2x += d2 * ux; v2y += d2 * uy;         // :217,:219
                        if (lane < n1) { s.vert[2 * (v1 + lane)] += p1x; s.vert[2 * (v1 + lane) + 1] += p1y; }  // :212
                        if (lane < n2) { s.vert[2 * (v2 + lane)] -= p2x; s.vert[2 * (v2 + lane) + 1] -= p2y; }  // :221
                        __syncwarp();
                        if (lane == 0) {
                            s.px[ii] = c1x; s.py[ii] = c1y; s.vx[ii] = v1x; s.vy[ii] = v1y; s.sx[ii] = c1x; s.sy[ii] = c1y;
                            s.px[jj] = c2x; s.py[jj] = c2y; s.vx[jj] = v2x; s.vy[jj] = v2y; s.sx[jj] = c2x; s.sy[jj] = c2y;
                        }
                    }
                    __syncwarp();
                }
                resolve = true;  // next round: check_collisions() at :234
                continue;
            }

            // ---- end of update(): :237-259 -------------------------------------------------------------------
            if (k > 0) {                                       // :248  dt != time_disc
                const double rdt = kp.time_disc - dt;
                const bool fell_back = dt < kp.time_tol;
                const double b0x = fell_back ? a0x : kp.accx, b0y = fell_back ? a0y : kp.accy;
                if (lane < F) {
                    double ox, oy, ovx, ovy;
                    FIZZ_MOVE_POINT(s.px[lane], s.py[lane], s.vx[lane], s.vy[lane], rdt, b0x, b0y, kp.accx, kp.accy, ox, oy,
                                    ovx, ovy);
                    s.px[lane] = ox; s.py[lane] = oy; s.vx[lane] = ovx; s.vy[lane] = ovy;
                }
                verts_explicit = false;                        // pre_update rebuilt the vertices from the com
            }
            if (lane < F) { s.sx[lane] = s.px[lane]; s.sy[lane] = s.py[lane]; }  // finish_update (:237-239, :251)
            __syncwarp();
            step++;                                            // :259
            resolve = false; k = 0; dt = kp.time_disc; ncalls = 0;
            if (step >= kp.target_steps) break;
            verts_explicit = false;
        }

        // ---- write the world back ----------------------------------------------------------------------------
        __syncwarp();
        bool finite = true;
        if (lane < F) {
            const double x = s.px[lane], y = s.py[lane];
            kp.pos[(2 * lane) * Wp + w] = x;
            kp.pos[(2 * lane + 1) * Wp + w] = y;
            kp.vel[(2 * lane) * Wp + w] = s.vx[lane];
            kp.vel[(2 * lane + 1) * Wp + w] = s.vy[lane];
            finite = (fabs(x) < INFINITY) && (fabs(y) < INFINITY);
        }
        const bool all_finite = __all_sync(FULL, finite);
        if (status == 0 && !all_finite) status = FIZZ_ST_NONFINITE;
        if (verts_explicit)
            for (int r = lane; r < 2 * V; r += 32) kp.verts[(size_t)r * Wp + w] = s.vert[r];
        if (lane == 0) {
            kp.heat[w] = heat;
            kp.status[w] = status;
            kp.steps[w] = step;
            kp.calls[w] += ncalls_total;
            kp.vflag[w] = verts_explicit ? 1 : 0;
        }
        __syncwarp();
    }
}

// Work list for the contact kernel: every running world that has not reached the chunk target.
__global__ void classify_kernel(const long long *status, const long long *steps, long long W, long long target,
                                long long *list, unsigned long long *count) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool lag = (w < W) && status[w] == 0 && steps[w] < target;
    const unsigned b = __ballot_sync(0xffffffffu, lag);
    if (b == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lag) list[base + __popc(b & ((1u << lane) - 1u))] = w;
}

// Derived vertices for worlds whose point.pos are not explicit state: |r|cos(phi) + com.x, |r|sin(phi) + com.y
// (physics.py:358-359).  Runs on demand (vertex download), never in the stepping loop.
__global__ void derive_verts_kernel(const double *pos, const double *off, const long long *vflag, double *verts,
                                    long long W, long long Wp, int V, const unsigned char *vbody) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W || vflag[w] != 0) return;
    for (int v = 0; v < V; v++) {
        const int b = vbody[v];
        verts[(size_t)(2 * v) * Wp + w] = off[(size_t)(2 * v) * Wp + w] + pos[(size_t)(2 * b) * Wp + w];
        verts[(size_t)(2 * v + 1) * Wp + w] = off[(size_t)(2 * v + 1) * Wp + w] + pos[(size_t)(2 * b + 1) * Wp + w];
    }
}

}  // namespace fizz
