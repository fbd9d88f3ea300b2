// fizz_thread_kernel.cuh -- thread-per-world kernel with the FULL update() semantics ("thread" mode).
//
// One THREAD per world, F (free bodies) a template parameter so that centre-of-mass position, velocity and
// the broad-phase snapshot of every body live in registers for all steps of a launch.  Vertices (only needed
// on contact steps) are staged in shared memory, column tid of a [vertex][xy][BLOCK] array (bank-conflict free);
// the fixed scene (vertices, unit edge normals, own projection bounds) is staged in shared memory once per CTA.
// A per-lane state machine makes ONE check_collisions()-equivalent per loop iteration, so lanes of a warp that
// sit in different timesteps / halving trials / resolve passes still execute the same instruction stream.
// Best for tiny worlds (F = 1) and as the independent cross-check of the hybrid path.
#pragma once

#include "fizz_common.cuh"
#include "fizz_ext.cuh"

namespace fizz {

constexpr int BLOCK = 128;

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
                            if (kp.mu > 0.0) ext_friction_fixed(kp.mu, kp.ope, m1, ux, uy, vdotN, vx[b], vy[b], heat);
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
                    double v1u = (mprop1 - mprop2) * vn1 + 2 * mprop2 * vn2;                     // :209
                    double v2u = 2 * mprop1 * vn1 + (mprop2 - mprop1) * vn2;                     // :218
                    if (kp.ffr >= 0.0) ext_restitution_free(kp.ffr, m1, mprop1, mprop2, vn1, vn2, v1u, v2u, heat);
                    const double d1 = v1u - vn1, d2 = v2u - vn2;                                 // :210,:219
                    v1x += d1 * ux; v1y += d1 * uy;                                              // :210
                    v2x += d2 * ux; v2y += d2 * uy;                                              // :219
                    if (kp.mu > 0.0)
                        ext_friction_free(kp.mu, kp.ffr >= 0.0 ? kp.ffr : 1.0, m1, m2, mprop2, ux, uy, vn1, vn2, v1x, v1y, v2x,
                                          v2y, heat);
                    const double p1x = mprop1 * nvx, p1y = mprop1 * nvy;
                    const double p2x = mprop2 * nvx, p2y = mprop2 * nvy;
#pragma unroll
                    for (int b = 0; b < F; b++) {
                        if (b == ii) {
                            px[b] += p1x; py[b] += p1y;                                          // :206
                            vx[b] = v1x; vy[b] = v1y;
                            sx[b] = px[b]; sy[b] = py[b];                                        // :215
                        }
                        if (b == jj) {
                            px[b] -= p2x; py[b] -= p2y;                                          // :217
                            vx[b] = v2x; vy[b] = v2y;
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
    kp.vflag[w] = verts_explicit ? 1 : 0;
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


#undef SV

}  // namespace fizz
