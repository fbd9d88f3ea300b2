// fizz_engine.cu -- B200 (sm_100a) batched stepping engine for Fizz-2D's World.update() hot path.
//
// One THREAD per world, F (free bodies) a template parameter so that centre-of-mass position, velocity and
// the broad-phase snapshot of every body live in registers for all n_steps of a launch.  Vertices (only needed
// on contact steps) are staged in shared memory, column tid of a [vertex][xy][BLOCK] array (bank-conflict free);
// the fixed scene (vertices, unit edge normals, own projection bounds) is staged in shared memory once per CTA.
// A per-lane state machine makes ONE check_collisions()-equivalent per loop iteration, so lanes of a warp that
// sit in different timesteps / halving trials / resolve passes still execute the same instruction stream.
//
// Semantics follow the reference op for op (file:line = /root/reference/src/physics.py):
//   World.update :76-259, check_collisions :262-289, polypoly_collision :765-840, Obj.pre_update :384-427,
//   Point.move :502-529, Obj.reverse_update :430-451, Obj.finish_update :453-464, World.energy :66-74.
// numpy's 2-vector dot/norm fuse the second product: dot = fma(a1,b1,a0*b0).  Compiled with
// -fmad=false so nothing else is contracted; IEEE div/sqrt are the CUDA defaults for double.
#include "fizz_b200.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

namespace {

constexpr int BLOCK = 128;
constexpr int TAB = FIZZ_MAX_FREE + FIZZ_MAX_FIXED;  // 24 bodies max

struct KParams {
    double *pos, *vel, *heat;
    long long *status, *steps, *calls;
    const double *off, *thr, *mass;
    double *verts;
    const double *fixed_geom;  // [FV][6]: x, y, unit normal of edge k->k+1 (nx, ny), own projection (lo, hi)
    FizzContact *trace;
    unsigned long long *trace_count;
    long long trace_cap;
    long long W, Wp, target_steps;
    int F, X, V, FV;
    int nv[TAB], vs[TAB];
    double fcx[FIZZ_MAX_FIXED], fcy[FIZZ_MAX_FIXED], fthr[FIZZ_MAX_FIXED];
    double coll_tol, time_tol, eps, vib_tol, time_disc, ope, ome2, accx, accy;
    int max_calls;
};

__device__ __forceinline__ double dot2(double a0, double a1, double b0, double b1) { return fma(a1, b1, a0 * b0); }

__device__ __forceinline__ void trace_hit(const KParams &kp, long long w, long long step, int call, int i, int j,
                                          double nx, double ny, double mag) {
    unsigned long long idx = atomicAdd(kp.trace_count, 1ULL);
    if ((long long)idx < kp.trace_cap) {
        FizzContact c;
        c.world = w; c.step = (int)step; c.call = call; c.i = i; c.j = j; c.nx = nx; c.ny = ny; c.mag = mag;
        kp.trace[idx] = c;
    }
}

#define SV(v, c) sv[(((v) * 2 + (c)) * BLOCK) + tid]

// polypoly_collision (physics.py:765-840) for the lane's pair (ii free, jj free-or-fixed).
// Free vertices come from the lane's shared-memory column, fixed ones (plus their precomputed unit normals and
// own projection bounds, which are bit-identical to what :790-811 recompute) from the staged fixed scene.
__device__ __forceinline__ bool sat_pair(const double *sv, const double *sfix, const int *stab, int tid, int F,
                                         int ii, int jj, double &onx, double &ony, double &omag) {
    const int n1 = stab[ii], v1 = stab[32 + ii];
    const int n2 = stab[jj], v2 = stab[32 + jj];
    const bool fx2 = jj >= F;
    double overlap = INFINITY, fxn = 0.0, fyn = 0.0;
    bool have = false;
    for (int k = 0; k < n1 + n2; k++) {
        double nx, ny, lo2, hi2;
        bool own2 = false;  // axis comes from a fixed poly2 edge: its own bounds are precomputed
        if (k < n1) {
            const int k2 = (k + 1 == n1) ? 0 : k + 1;
            const double ax = SV(v1 + k2, 1) - SV(v1 + k, 1);  // :775  (p2.y - p1.y, p1.x - p2.x)
            const double ay = SV(v1 + k, 0) - SV(v1 + k2, 0);
            const double nrm = sqrt(fma(ay, ay, ax * ax));     // :790
            nx = ax / nrm; ny = ay / nrm;
        } else if (fx2) {
            const double *g = sfix + (v2 + (k - n1)) * 6;
            nx = g[2]; ny = g[3]; lo2 = g[4]; hi2 = g[5];
            own2 = true;
        } else {
            const int e = k - n1;
            const int e2 = (e + 1 == n2) ? 0 : e + 1;
            const double ax = SV(v2 + e2, 1) - SV(v2 + e, 1);  // :778
            const double ay = SV(v2 + e, 0) - SV(v2 + e2, 0);
            const double nrm = sqrt(fma(ay, ay, ax * ax));
            nx = ax / nrm; ny = ay / nrm;
        }
        double lo1, hi1;
        lo1 = hi1 = dot2(nx, ny, SV(v1, 0), SV(v1, 1));        // :795
        for (int v = 1; v < n1; v++) {                         // :799-804 (v = 0 compares equal: no change)
            const double cur = dot2(nx, ny, SV(v1 + v, 0), SV(v1 + v, 1));
            if (cur < lo1) lo1 = cur;
            if (cur > hi1) hi1 = cur;
        }
        if (!own2) {
            if (fx2) {
                const double *g = sfix + v2 * 6;
                lo2 = hi2 = dot2(nx, ny, g[0], g[1]);          // :796
                for (int v = 1; v < n2; v++) {                 // :806-811
                    const double cur = dot2(nx, ny, g[v * 6], g[v * 6 + 1]);
                    if (cur < lo2) lo2 = cur;
                    if (cur > hi2) hi2 = cur;
                }
            } else {
                lo2 = hi2 = dot2(nx, ny, SV(v2, 0), SV(v2, 1));
                for (int v = 1; v < n2; v++) {
                    const double cur = dot2(nx, ny, SV(v2 + v, 0), SV(v2 + v, 1));
                    if (cur < lo2) lo2 = cur;
                    if (cur > hi2) hi2 = cur;
                }
            }
        }
        if (lo1 > lo2) {                                       // :815-817
            double t = lo1; lo1 = lo2; lo2 = t;
            t = hi1; hi1 = hi2; hi2 = t;
        }
        if (hi1 < lo2) return false;                           // :820-821
        const double cur = hi1 - lo2;                          // :824
        if (cur < overlap) { fxn = nx; fyn = ny; overlap = cur; have = true; }  // :827-830
    }
    if (!have) return false;  // all-NaN axes (reference raises)
    onx = fxn * -1.0;         // :838-839 always negated
    ony = fyn * -1.0;
    omag = overlap;
    return true;
}

template <int F>
__global__ void __launch_bounds__(BLOCK, 3) step_kernel(const __grid_constant__ KParams kp) {
    extern __shared__ double smem[];
    double *sfix = smem;
    double *sv = smem + kp.FV * 6;
    int *stab = reinterpret_cast<int *>(sv + (size_t)kp.V * 2 * BLOCK);
    const int tid = threadIdx.x;
    for (int i = tid; i < kp.FV * 6; i += BLOCK) sfix[i] = kp.fixed_geom[i];
    if (tid < F + kp.X) {
        stab[tid] = kp.nv[tid];
        stab[32 + tid] = kp.vs[tid];
    }
    __syncthreads();
    const long long w = (long long)blockIdx.x * BLOCK + tid;
    if (w >= kp.W) return;
    long long status = kp.status[w];
    long long step = kp.steps[w];
    if (status != 0 || step >= kp.target_steps) return;
    const long long Wp = kp.Wp;
    const int X = kp.X;

    // ---- registers: com.pos, com.vel, snapshot obj.pos, broad-phase thresholds -------------------------------
    double px[F], py[F], vx[F], vy[F], sx[F], sy[F], thr[F];
#pragma unroll
    for (int b = 0; b < F; b++) {
        px[b] = kp.pos[(2 * b) * Wp + w];
        py[b] = kp.pos[(2 * b + 1) * Wp + w];
        vx[b] = kp.vel[(2 * b) * Wp + w];
        vy[b] = kp.vel[(2 * b + 1) * Wp + w];
        sx[b] = px[b]; sy[b] = py[b];  // finish_update ran at the end of the previous update() (:237-239)
        thr[b] = kp.thr[b * Wp + w];
    }
    double heat = kp.heat[w];
    long long ncalls_total = 0;

    // per-lane state machine
    bool resolve = false;      // false: step-halving trial (:99-129); true: resolve loop (:140-234)
    bool verts_explicit = false;
    int k = 0;                 // halving trial index; dt = time_disc / 2^k
    double dt = kp.time_disc;  // :89
    int ncalls = 0;            // state-relevant check_collisions calls of the current update()
    unsigned mask[F];          // broad-phase survivors: bit (j-i-1) of mask[i] = pair (i, j), j over objs+fixed_objs
#pragma unroll
    for (int b = 0; b < F; b++) mask[b] = 0;

    // contact tuples of the current check_collisions (local memory; touched on contact steps only)
    unsigned char hi_i[FIZZ_MAX_HITS], hi_j[FIZZ_MAX_HITS];
    double hnx[FIZZ_MAX_HITS], hny[FIZZ_MAX_HITS], hmag[FIZZ_MAX_HITS];

    // Point.move (physics.py:502-529) of body b over `h`, starting acceleration (a0x,a0y)
#define FIZZ_MOVE(b, h, a0x, a0y, ox, oy, ovx, ovy)                                                   \
    {                                                                                                 \
        const double halfdt_ = .5 * (h);                                                              \
        const double vax_ = vx[b] + (a0x) * halfdt_, vay_ = vy[b] + (a0y) * halfdt_;                  \
        ox = px[b] + vax_ * (h);                                                                      \
        oy = py[b] + vay_ * (h);                                                                      \
        ovx = vax_ + kp.accx * halfdt_;                                                               \
        ovy = vay_ + kp.accy * halfdt_;                                                               \
    }

    while (true) {
        // acceleration held by every com at the start of this update(): 0 before the first move ever (:494)
        const bool acc_on = step > 0;
        const double a0x = acc_on ? kp.accx : 0.0, a0y = acc_on ? kp.accy : 0.0;
        bool need_mask = resolve;
        if (!resolve) {
            if (k == 0) {
                need_mask = true;
                if (step == 0) {
                    // obj.pos aliases com.pos until the first reverse/finish_update (:614): the first
                    // trial of the first step sees the moved position, and it then stays frozen for the step.
#pragma unroll
                    for (int b = 0; b < F; b++) {
                        double tvx, tvy;
                        FIZZ_MOVE(b, dt, a0x, a0y, sx[b], sy[b], tvx, tvy);
                        (void)tvx; (void)tvy;
                    }
                }
            }
        } else if (ncalls >= kp.max_calls) {
            status = FIZZ_ST_CAPPED;  // watchdog: the resolve loop wants call max_calls+1
            break;
        }

        // ---- broad phase (:272-281): dist > max(r_i, r_j)  <=>  s > T(r_i) && s > T(r_j) ---------------------
        if (need_mask) {
#pragma unroll
            for (int i = 0; i < F; i++) {
                unsigned m = 0;
#pragma unroll
                for (int j = i + 1; j < F; j++) {
                    const double dx = sx[i] - sx[j], dy = sy[i] - sy[j];
                    const double s = fma(dy, dy, dx * dx);
                    if (!(s > thr[i] && s > thr[j])) m |= 1u << (j - i - 1);
                }
                mask[i] = m;
            }
            for (int j = 0; j < X; j++) {
                const double cx = kp.fcx[j], cy = kp.fcy[j], ft = kp.fthr[j];
#pragma unroll
                for (int i = 0; i < F; i++) {
                    const double dx = sx[i] - cx, dy = sy[i] - cy;
                    const double s = fma(dy, dy, dx * dx);
                    if (!(s > thr[i] && s > ft)) mask[i] |= 1u << (F - i - 1 + j);
                }
            }
        }
        ncalls++;
        ncalls_total++;

        unsigned any = 0;
#pragma unroll
        for (int b = 0; b < F; b++) any |= mask[b];

        int nh = 0;
        double maxov = 0.0;  // :125-129
        if (any) {
            if (!resolve) {
                // pre_update rebuilt every vertex from the moved com (:358-359): |r|cos(phi) + com.x
#pragma unroll
                for (int b = 0; b < F; b++) {
                    double nxp, nyp, tvx, tvy;
                    FIZZ_MOVE(b, dt, a0x, a0y, nxp, nyp, tvx, tvy);
                    (void)tvx; (void)tvy;
                    const int n = kp.nv[b], v0 = kp.vs[b];
                    for (int v = 0; v < n; v++) {
                        SV(v0 + v, 0) = kp.off[(size_t)((v0 + v) * 2) * Wp + w] + nxp;
                        SV(v0 + v, 1) = kp.off[(size_t)((v0 + v) * 2 + 1) * Wp + w] + nyp;
                    }
                }
            }
            unsigned m2[F];
#pragma unroll
            for (int b = 0; b < F; b++) m2[b] = mask[b];
            while (true) {
                int ii = -1, jj = 0;
#pragma unroll
                for (int b = 0; b < F; b++) {
                    if (ii < 0 && m2[b]) {
                        const int bit = __ffs(m2[b]) - 1;
                        m2[b] &= m2[b] - 1;
                        ii = b;
                        jj = b + 1 + bit;
                    }
                }
                if (ii < 0) break;
                double nx, ny, mag;
                if (sat_pair(sv, sfix, stab, tid, F, ii, jj, nx, ny, mag)) {  // :286-288
                    if (nh < FIZZ_MAX_HITS) {
                        hi_i[nh] = (unsigned char)ii; hi_j[nh] = (unsigned char)jj;
                        hnx[nh] = nx; hny[nh] = ny; hmag[nh] = mag;
                    }
                    nh++;
                    const double am = fabs(mag);
                    if (nh == 1 || am > maxov) maxov = am;
                    if (kp.trace) trace_hit(kp, w, step, ncalls, ii, jj, nx, ny, mag);
                }
            }
            if (nh > FIZZ_MAX_HITS) { status = FIZZ_ST_HIT_OVERFLOW; break; }
        }

        if (!resolve) {
            // ---- :99 loop condition ---------------------------------------------------------------------------
            if (maxov > kp.coll_tol && dt > kp.time_tol) {
                dt = dt / 2;  // :106 (reverse_update: every trial restarts from the start-of-step com)
                k++;
                continue;
            }
            if (dt < kp.time_tol) {
                // :133-137 fallback: com goes back to the start of the step, vertices keep the last trial's
                // positions, so check_collisions() returns the very same list again.
                ncalls++;
                ncalls_total++;
                if (kp.trace)
                    for (int h = 0; h < nh; h++) trace_hit(kp, w, step, ncalls, hi_i[h], hi_j[h], hnx[h], hny[h], hmag[h]);
            } else {
#pragma unroll
                for (int b = 0; b < F; b++) {
                    double nxp, nyp, nvx, nvy;
                    FIZZ_MOVE(b, dt, a0x, a0y, nxp, nyp, nvx, nvy);
                    px[b] = nxp; py[b] = nyp; vx[b] = nvx; vy[b] = nvy;
                }
            }
            if (nh) verts_explicit = true;
        }

        if (nh) {
            // ---- one resolve pass over the list computed above (:142-233) -------------------------
            for (int h = 0; h < nh; h++) {
                const int ii = hi_i[h], jj = hi_j[h];
                const double ux = hnx[h], uy = hny[h];
                const double s = hmag[h] + kp.eps;         // :158
                const double nvx = ux * s, nvy = uy * s;
                const int n1 = stab[ii], v1 = stab[32 + ii];
                if (jj >= F) {                             // :171-198 free body vs fixed body
                    for (int v = 0; v < n1; v++) { SV(v1 + v, 0) += nvx; SV(v1 + v, 1) += nvy; }  // :174
                    const bool gate = sqrt(fma(nvy, nvy, nvx * nvx)) > kp.vib_tol;               // :180
                    const double m1 = kp.mass[ii * Wp + w];
#pragma unroll
                    for (int b = 0; b < F; b++) {
                        if (b == ii) {
                            if (gate) { px[b] += nvx; py[b] += nvy; }                            // :181
                            const double vdotN = dot2(vx[b], vy[b], ux, uy);                     // :183
                            const double cc = kp.ope * vdotN;                                    // :188
                            vx[b] -= cc * ux;
                            vy[b] -= cc * uy;
                            heat += .5 * m1 * kp.ome2 * (vdotN * vdotN);                         // :191,:195
                            sx[b] = px[b]; sy[b] = py[b];                                        // :197
                        }
                    }
                } else {                                   // :201-223 two free bodies
                    const int n2 = stab[jj], v2 = stab[32 + jj];
                    const double m1 = kp.mass[ii * Wp + w], m2 = kp.mass[jj * Wp + w];
                    const double mtotal = m1 + m2;
                    const double mprop1 = m1 / mtotal;
                    const double mprop2 = 1 - mprop1;
                    double v1x = 0, v1y = 0, v2x = 0, v2y = 0;
#pragma unroll
                    for (int b = 0; b < F; b++) {
                        if (b == ii) { v1x = vx[b]; v1y = vy[b]; }
                        if (b == jj) { v2x = vx[b]; v2y = vy[b]; }
                    }
                    const double vn1 = dot2(v1x, v1y, ux, uy);                                   // :207
                    const double vn2 = dot2(v2x, v2y, ux, uy);                                   // :208
                    const double v1u = (mprop1 - mprop2) * vn1 + 2 * mprop2 * vn2;               // :209
                    const double v2u = 2 * mprop1 * vn1 + (mprop2 - mprop1) * vn2;               // :218
                    const double d1 = v1u - vn1, d2 = v2u - vn2;                                 // :210,:219
                    const double p1x = mprop1 * nvx, p1y = mprop1 * nvy;
                    const double p2x = mprop2 * nvx, p2y = mprop2 * nvy;
#pragma unroll
                    for (int b = 0; b < F; b++) {
                        if (b == ii) {
                            px[b] += p1x; py[b] += p1y;                                          // :206
                            vx[b] += d1 * ux; vy[b] += d1 * uy;                                  // :210
                            sx[b] = px[b]; sy[b] = py[b];                                        // :215
                        }
                        if (b == jj) {
                            px[b] -= p2x; py[b] -= p2y;                                          // :217
                            vx[b] += d2 * ux; vy[b] += d2 * uy;                                  // :219
                            sx[b] = px[b]; sy[b] = py[b];                                        // :223
                        }
                    }
                    for (int v = 0; v < n1; v++) { SV(v1 + v, 0) += p1x; SV(v1 + v, 1) += p1y; } // :212
                    for (int v = 0; v < n2; v++) { SV(v2 + v, 0) -= p2x; SV(v2 + v, 1) -= p2y; } // :221
                }
            }
            resolve = true;  // next iteration: check_collisions() at :234
            continue;
        }

        // ---- end of update(): :237-259 ---------------------------------------------------------------------
        if (k > 0) {                                       // :248  dt != time_disc
            const double rdt = kp.time_disc - dt;
            // acceleration now held by the com: installed by the accepted trial's move, or still the
            // start-of-step one when the fallback reverted it
            const bool fell_back = dt < kp.time_tol;
            const double b0x = fell_back ? a0x : kp.accx, b0y = fell_back ? a0y : kp.accy;
#pragma unroll
            for (int b = 0; b < F; b++) {
                double nxp, nyp, nvx, nvy;
                FIZZ_MOVE(b, rdt, b0x, b0y, nxp, nyp, nvx, nvy);
                px[b] = nxp; py[b] = nyp; vx[b] = nvx; vy[b] = nvy;
            }
            verts_explicit = false;                        // pre_update rebuilt the vertices from the com
        }
#pragma unroll
        for (int b = 0; b < F; b++) { sx[b] = px[b]; sy[b] = py[b]; }  // finish_update (:237-239, :251)
        step++;                                            // :259
        resolve = false; k = 0; dt = kp.time_disc; ncalls = 0;
        if (step >= kp.target_steps) break;
        verts_explicit = false;  // the next pre_update rebuilds every vertex
    }
#undef FIZZ_MOVE

    // ---- write back -------------------------------------------------------------------------------------------
    bool finite = true;
#pragma unroll
    for (int b = 0; b < F; b++) {
        kp.pos[(2 * b) * Wp + w] = px[b];
        kp.pos[(2 * b + 1) * Wp + w] = py[b];
        kp.vel[(2 * b) * Wp + w] = vx[b];
        kp.vel[(2 * b + 1) * Wp + w] = vy[b];
        finite = finite && (fabs(px[b]) < INFINITY) && (fabs(py[b]) < INFINITY);
    }
    if (status == 0 && !finite) status = FIZZ_ST_NONFINITE;
    kp.heat[w] = heat;
    kp.status[w] = status;
    kp.steps[w] = step;
    kp.calls[w] += ncalls_total;
    if (kp.verts) {
#pragma unroll
        for (int b = 0; b < F; b++) {
            const int n = kp.nv[b], v0 = kp.vs[b];
            for (int v = 0; v < n; v++) {
                const size_t r0 = (size_t)((v0 + v) * 2) * Wp + w, r1 = (size_t)((v0 + v) * 2 + 1) * Wp + w;
                if (verts_explicit) {
                    kp.verts[r0] = SV(v0 + v, 0);
                    kp.verts[r1] = SV(v0 + v, 1);
                } else {
                    kp.verts[r0] = kp.off[r0] + px[b];
                    kp.verts[r1] = kp.off[r1] + py[b];
                }
            }
        }
    }
}

// World.energy (physics.py:66-74): K = sum (.5*m)*norm(v)**2 ; P = sum (GRAVITY[1]*m)*(height - y) ; H = world.heat
__global__ void energy_kernel(const double *pos, const double *vel, const double *heat, const double *mass,
                              double *energy, long long W, long long Wp, int F, double gy, double height) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    double K = 0.0, P = 0.0;
    for (int b = 0; b < F; b++) {
        const double m = mass[b * Wp + w];
        const double vx = vel[(2 * b) * Wp + w], vy = vel[(2 * b + 1) * Wp + w];
        const double nv = sqrt(fma(vy, vy, vx * vx));
        K += (.5 * m) * (nv * nv);
        P += (gy * m) * (height - pos[(2 * b + 1) * Wp + w]);
    }
    const double H = heat[w];
    energy[0 * Wp + w] = ((0.0 + K) + P) + H;
    energy[1 * Wp + w] = K;
    energy[2 * Wp + w] = P;
    energy[3 * Wp + w] = H;
}

// FP64 FMA throughput probe: 8 independent chains per thread, fully unrolled.
__global__ void fp64_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char *what) {
    return fail(FIZZ_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                            \
    do {                                                    \
        cudaError_t e_ = (call);                            \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

typedef void (*step_fn)(const KParams);
template <int F>
void launch_step(const KParams &kp, dim3 grid, size_t smem, cudaStream_t st) {
    step_kernel<F><<<grid, BLOCK, smem, st>>>(kp);
}
template <int F>
cudaError_t prep_step(size_t smem, cudaFuncAttributes *attr, int *blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(step_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(attr, step_kernel<F>);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, step_kernel<F>, BLOCK, smem);
}

}  // namespace

struct FizzGroup {
    FizzLayout lay;
    FizzParams prm;
    KParams kp;
    int device;
    double *slab;
    size_t smem;
    bool uploaded;
    long long target;
    long long launches;
    cudaFuncAttributes attr;
    int blocks_per_sm;
    // staging for uploads of padded rows
};

extern "C" {

int fizz_abi_version(void) { return FIZZ_ABI_VERSION; }
const char *fizz_last_error(void) { return g_err.c_str(); }

void fizz_host_norm2(int64_t n, const double *x, const double *y, double *out) {
    for (int64_t i = 0; i < n; i++) out[i] = std::sqrt(std::fma(y[i], y[i], x[i] * x[i]));
}
void fizz_host_dot2(int64_t n, const double *a0, const double *a1, const double *b0, const double *b1, double *out) {
    for (int64_t i = 0; i < n; i++) out[i] = std::fma(a1[i], b1[i], a0[i] * b0[i]);
}
void fizz_host_sqrt_threshold(int64_t n, const double *r, double *out) {
    for (int64_t i = 0; i < n; i++) {
        const double ri = r[i];
        if (!(ri >= 0.0) || std::isinf(ri)) { out[i] = ri; continue; }  // NaN / negative / inf: pass through
        double s = ri * ri;
        while (std::sqrt(s) > ri) s = std::nextafter(s, 0.0);
        for (;;) {
            const double up = std::nextafter(s, INFINITY);
            if (std::isinf(up) || std::sqrt(up) > ri) break;
            s = up;
        }
        out[i] = s;
    }
}

int fizz_layout(const FizzGroupDesc *d, int64_t n_worlds, FizzLayout *out) {
    if (!d || !out) return fail(FIZZ_E_ARG, "fizz_layout: null argument");
    if (d->n_free < 1 || d->n_free > FIZZ_MAX_FREE) return fail(FIZZ_E_ARG, "fizz_layout: n_free must be 1..8");
    if (d->n_fixed < 0 || d->n_fixed > FIZZ_MAX_FIXED) return fail(FIZZ_E_ARG, "fizz_layout: n_fixed must be 0..16");
    if (n_worlds < 1) return fail(FIZZ_E_ARG, "fizz_layout: n_worlds must be >= 1");
    int V = 0, FV = 0;
    for (int b = 0; b < d->n_free + d->n_fixed; b++) {
        const int n = d->nverts[b];
        if (n < 3 || n > FIZZ_MAX_POLY_VERTS) return fail(FIZZ_E_ARG, "fizz_layout: polygons need 3..16 vertices");
        if (b < d->n_free) V += n; else FV += n;
    }
    if (FV > FIZZ_MAX_FIXED_VERTS) return fail(FIZZ_E_ARG, "fizz_layout: too many fixed vertices (max 96)");
    FizzLayout L;
    std::memset(&L, 0, sizeof(L));
    const int64_t Wp = round_up(n_worlds, BLOCK);
    const int F = d->n_free;
    L.n_worlds = n_worlds; L.stride = Wp; L.n_free = F; L.n_fixed = d->n_fixed; L.n_free_verts = V;
    int64_t o = 0;
    L.pos = o; o += 2 * F * Wp;
    L.vel = o; o += 2 * F * Wp;
    L.heat = o; o += Wp;
    L.status = o; o += Wp;
    L.steps = o; o += Wp;
    L.calls = o; o += Wp;
    L.state_words = o;
    L.off = o; o += 2 * (int64_t)V * Wp;
    L.thr = o; o += F * Wp;
    L.mass = o; o += F * Wp;
    L.verts = o; o += 2 * (int64_t)V * Wp;
    L.energy = o; o += 4 * Wp;
    o += round_up((int64_t)FV * 6, 16);  // fixed scene constants (library-written)
    L.total_words = o;
    *out = L;
    return FIZZ_OK;
}

int fizz_group_create(const FizzGroupDesc *d, int64_t n_worlds, int device, void *dev_slab, int64_t slab_words,
                      FizzGroup **out) {
    if (!out) return fail(FIZZ_E_ARG, "fizz_group_create: null out");
    *out = nullptr;
    FizzLayout L;
    int rc = fizz_layout(d, n_worlds, &L);
    if (rc) return rc;
    if (!dev_slab || slab_words < L.total_words) return fail(FIZZ_E_ARG, "fizz_group_create: slab too small");
    if (d->params.max_calls < 32) return fail(FIZZ_E_ARG, "fizz_group_create: max_calls must be >= 32");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(FIZZ_E_CUDA, "fizz_group_create: no such CUDA device");
    CU(cudaSetDevice(device));

    FizzGroup *g = new FizzGroup();
    std::memset(&g->kp, 0, sizeof(g->kp));
    g->lay = L; g->prm = d->params; g->device = device; g->slab = (double *)dev_slab;
    g->uploaded = false; g->target = 0; g->launches = 0;
    KParams &kp = g->kp;
    double *s = g->slab;
    kp.pos = s + L.pos; kp.vel = s + L.vel; kp.heat = s + L.heat;
    kp.status = (long long *)(s + L.status); kp.steps = (long long *)(s + L.steps); kp.calls = (long long *)(s + L.calls);
    kp.off = s + L.off; kp.thr = s + L.thr; kp.mass = s + L.mass; kp.verts = s + L.verts;
    double *fixed_dev = s + L.energy + 4 * L.stride;
    kp.fixed_geom = fixed_dev;
    kp.W = L.n_worlds; kp.Wp = L.stride;
    kp.F = L.n_free; kp.X = L.n_fixed; kp.V = L.n_free_verts;
    const int F = kp.F, X = kp.X;
    int vfree = 0, vfix = 0;
    for (int b = 0; b < F + X; b++) {
        kp.nv[b] = d->nverts[b];
        if (b < F) { kp.vs[b] = vfree; vfree += d->nverts[b]; }
        else { kp.vs[b] = vfix; vfix += d->nverts[b]; }
    }
    kp.FV = vfix;
    // fixed-scene constants, computed exactly as polypoly_collision would (:775-811) every call
    double *fg = new double[(size_t)(vfix > 0 ? vfix : 1) * 6];
    for (int j = 0; j < X; j++) {
        const int n = kp.nv[F + j], v0 = kp.vs[F + j];
        const double *xy = d->fixed_xy + 2 * v0;
        kp.fcx[j] = d->fixed_com[2 * j]; kp.fcy[j] = d->fixed_com[2 * j + 1]; kp.fthr[j] = d->fixed_thr[j];
        for (int k = 0; k < n; k++) {
            const int k2 = (k + 1 == n) ? 0 : k + 1;
            const double ax = xy[2 * k2 + 1] - xy[2 * k + 1];
            const double ay = xy[2 * k] - xy[2 * k2];
            const double nrm = std::sqrt(std::fma(ay, ay, ax * ax));
            const double nx = ax / nrm, ny = ay / nrm;
            double lo, hi;
            lo = hi = std::fma(ny, xy[1], nx * xy[0]);
            for (int v = 1; v < n; v++) {
                const double cur = std::fma(ny, xy[2 * v + 1], nx * xy[2 * v]);
                if (cur < lo) lo = cur;
                if (cur > hi) hi = cur;
            }
            double *o6 = fg + (size_t)(v0 + k) * 6;
            o6[0] = xy[2 * k]; o6[1] = xy[2 * k + 1]; o6[2] = nx; o6[3] = ny; o6[4] = lo; o6[5] = hi;
        }
    }
    cudaError_t e = cudaMemcpy(fixed_dev, fg, sizeof(double) * (size_t)vfix * 6, cudaMemcpyHostToDevice);
    delete[] fg;
    if (e != cudaSuccess) { delete g; return cuda_fail(e, "cudaMemcpy(fixed scene)"); }
    const FizzParams &p = d->params;
    kp.coll_tol = p.collision_tol; kp.time_tol = p.time_tol; kp.eps = p.epsilon; kp.vib_tol = p.vibrate_tol;
    kp.time_disc = p.time_disc; kp.ope = p.one_plus_e; kp.ome2 = p.one_minus_e2; kp.accx = p.acc_x; kp.accy = p.acc_y;
    kp.max_calls = p.max_calls;
    g->smem = sizeof(double) * ((size_t)kp.FV * 6 + (size_t)kp.V * 2 * BLOCK) + sizeof(int) * 64;
    switch (F) {
        case 1: e = prep_step<1>(g->smem, &g->attr, &g->blocks_per_sm); break;
        case 2: e = prep_step<2>(g->smem, &g->attr, &g->blocks_per_sm); break;
        case 3: e = prep_step<3>(g->smem, &g->attr, &g->blocks_per_sm); break;
        case 4: e = prep_step<4>(g->smem, &g->attr, &g->blocks_per_sm); break;
        case 5: e = prep_step<5>(g->smem, &g->attr, &g->blocks_per_sm); break;
        case 6: e = prep_step<6>(g->smem, &g->attr, &g->blocks_per_sm); break;
        case 7: e = prep_step<7>(g->smem, &g->attr, &g->blocks_per_sm); break;
        default: e = prep_step<8>(g->smem, &g->attr, &g->blocks_per_sm); break;
    }
    if (e != cudaSuccess) { delete g; return cuda_fail(e, "step kernel setup (shared memory too large?)"); }
    *out = g;
    return FIZZ_OK;
}

int fizz_group_destroy(FizzGroup *g) {
    delete g;
    return FIZZ_OK;
}

static int copy_rows(double *dst, const double *src, int64_t rows, int64_t W, int64_t Wp, cudaMemcpyKind kind,
                     cudaStream_t st) {
    if (rows == 0) return FIZZ_OK;
    if (kind == cudaMemcpyHostToDevice)
        CU(cudaMemcpy2DAsync(dst, sizeof(double) * Wp, src, sizeof(double) * W, sizeof(double) * W, rows, kind, st));
    else
        CU(cudaMemcpy2DAsync(dst, sizeof(double) * W, src, sizeof(double) * Wp, sizeof(double) * W, rows, kind, st));
    return FIZZ_OK;
}

int fizz_upload(FizzGroup *g, const double *pos, const double *vel, const double *off, const double *thr,
                const double *mass, void *stream) {
    if (!g || !pos || !vel || !off || !thr || !mass) return fail(FIZZ_E_ARG, "fizz_upload: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(g->device));
    const FizzLayout &L = g->lay;
    const int64_t W = L.n_worlds, Wp = L.stride;
    const int F = L.n_free, V = L.n_free_verts;
    double *s = g->slab;
    // heat/status/steps/calls := 0 (also clears the padding columns of pos/vel so they stay finite)
    CU(cudaMemsetAsync(s + L.pos, 0, sizeof(double) * L.state_words, st));
    int rc;
    if ((rc = copy_rows(s + L.pos, pos, 2 * F, W, Wp, cudaMemcpyHostToDevice, st))) return rc;
    if ((rc = copy_rows(s + L.vel, vel, 2 * F, W, Wp, cudaMemcpyHostToDevice, st))) return rc;
    if ((rc = copy_rows(s + L.off, off, 2 * V, W, Wp, cudaMemcpyHostToDevice, st))) return rc;
    if ((rc = copy_rows(s + L.thr, thr, F, W, Wp, cudaMemcpyHostToDevice, st))) return rc;
    if ((rc = copy_rows(s + L.mass, mass, F, W, Wp, cudaMemcpyHostToDevice, st))) return rc;
    g->uploaded = true;
    g->target = 0;
    return FIZZ_OK;
}

int fizz_step(FizzGroup *g, int64_t n_steps, void *stream) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_step: null group");
    if (!g->uploaded) return fail(FIZZ_E_STATE, "fizz_step: call fizz_upload first");
    if (n_steps < 0) return fail(FIZZ_E_ARG, "fizz_step: n_steps < 0");
    if (n_steps == 0) return FIZZ_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(g->device));
    g->target += n_steps;
    KParams kp = g->kp;
    kp.target_steps = g->target;
    dim3 grid((unsigned)(g->lay.stride / BLOCK));
    switch (kp.F) {
        case 1: launch_step<1>(kp, grid, g->smem, st); break;
        case 2: launch_step<2>(kp, grid, g->smem, st); break;
        case 3: launch_step<3>(kp, grid, g->smem, st); break;
        case 4: launch_step<4>(kp, grid, g->smem, st); break;
        case 5: launch_step<5>(kp, grid, g->smem, st); break;
        case 6: launch_step<6>(kp, grid, g->smem, st); break;
        case 7: launch_step<7>(kp, grid, g->smem, st); break;
        default: launch_step<8>(kp, grid, g->smem, st); break;
    }
    CU(cudaGetLastError());
    g->launches++;
    return FIZZ_OK;
}

int fizz_energy(FizzGroup *g, void *stream) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_energy: null group");
    if (!g->uploaded) return fail(FIZZ_E_STATE, "fizz_energy: call fizz_upload first");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(g->device));
    const FizzLayout &L = g->lay;
    double *s = g->slab;
    const int threads = 256;
    const unsigned blocks = (unsigned)((L.n_worlds + threads - 1) / threads);
    energy_kernel<<<blocks, threads, 0, st>>>(s + L.pos, s + L.vel, s + L.heat, s + L.mass, s + L.energy, L.n_worlds,
                                              L.stride, L.n_free, g->prm.gy, g->prm.height);
    CU(cudaGetLastError());
    g->launches++;
    return FIZZ_OK;
}

int fizz_download(FizzGroup *g, double *pos, double *vel, double *heat, double *energy, double *verts,
                  int64_t *status, int64_t *steps, int64_t *calls, void *stream) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_download: null group");
    if (!g->uploaded) return fail(FIZZ_E_STATE, "fizz_download: call fizz_upload first");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(g->device));
    const FizzLayout &L = g->lay;
    const int64_t W = L.n_worlds, Wp = L.stride;
    const int F = L.n_free, V = L.n_free_verts;
    double *s = g->slab;
    int rc;
    if (energy && (rc = fizz_energy(g, stream))) return rc;
    if (pos && (rc = copy_rows(pos, s + L.pos, 2 * F, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (vel && (rc = copy_rows(vel, s + L.vel, 2 * F, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (heat && (rc = copy_rows(heat, s + L.heat, 1, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (energy && (rc = copy_rows(energy, s + L.energy, 4, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (verts && (rc = copy_rows(verts, s + L.verts, 2 * V, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (status && (rc = copy_rows((double *)status, s + L.status, 1, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (steps && (rc = copy_rows((double *)steps, s + L.steps, 1, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    if (calls && (rc = copy_rows((double *)calls, s + L.calls, 1, W, Wp, cudaMemcpyDeviceToHost, st))) return rc;
    return FIZZ_OK;
}

int fizz_set_trace(FizzGroup *g, void *dev_records, int64_t capacity, void *dev_counter) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_set_trace: null group");
    if (dev_records && (!dev_counter || capacity < 1)) return fail(FIZZ_E_ARG, "fizz_set_trace: need counter and capacity");
    g->kp.trace = (FizzContact *)dev_records;
    g->kp.trace_count = (unsigned long long *)dev_counter;
    g->kp.trace_cap = dev_records ? capacity : 0;
    return FIZZ_OK;
}

int fizz_kernel_info(FizzGroup *g, int32_t *regs, int32_t *smem, int32_t *threads, int32_t *blocks_per_sm,
                     int64_t *launches) {
    if (!g) return fail(FIZZ_E_ARG, "fizz_kernel_info: null group");
    if (regs) *regs = g->attr.numRegs;
    if (smem) *smem = (int32_t)g->smem;
    if (threads) *threads = BLOCK;
    if (blocks_per_sm) *blocks_per_sm = g->blocks_per_sm;
    if (launches) *launches = g->launches;
    return FIZZ_OK;
}

int fizz_fp64_peak(int device, int32_t iters, double *tflops, double *ms_out) {
    if (!tflops || iters < 1) return fail(FIZZ_E_ARG, "fizz_fp64_peak: bad argument");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8;
    double *out = nullptr;
    CU(cudaMalloc(&out, sizeof(double) * (size_t)threads * blocks));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    fp64_peak_kernel<<<blocks, threads>>>(out, iters / 8 + 1, 1.0000001, 1e-9);  // warm-up
    CU(cudaEventRecord(e0));
    fp64_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    CU(cudaGetLastError());
    const double flops = 2.0 * 64.0 * (double)iters * threads * blocks;
    *tflops = flops / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return FIZZ_OK;
}

}  // extern "C"
