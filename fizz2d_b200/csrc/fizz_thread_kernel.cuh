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
template <typename T>
__device__ __forceinline__ bool sat_pair(const T *sv, const T *sfix, const int *stab, int tid, int F,
                                         int ii, int jj, T &onx, T &ony, T &omag) {
    const int n1 = stab[ii], v1 = stab[32 + ii];
    const int n2 = stab[jj], v2 = stab[32 + jj];
    const bool fx2 = jj >= F;
    T overlap = INFINITY, fxn = 0, fyn = 0;
    bool have = false;
    for (int k = 0; k < n1 + n2; k++) {
        T nx, ny, lo2, hi2;
        bool own2 = false;  // axis comes from a fixed poly2 edge: its own bounds are precomputed
        if (k < n1) {
            const int k2 = (k + 1 == n1) ? 0 : k + 1;
            const T ax = SV(v1 + k2, 1) - SV(v1 + k, 1);  // :775  (p2.y - p1.y, p1.x - p2.x)
            const T ay = SV(v1 + k, 0) - SV(v1 + k2, 0);
            const T nrm = sqrt(sq2(ax, ay));     // :790
            nx = ax / nrm; ny = ay / nrm;
        } else if (fx2) {
            const T *g = sfix + (v2 + (k - n1)) * 6;
            nx = g[2]; ny = g[3]; lo2 = g[4]; hi2 = g[5];
            own2 = true;
        } else {
            const int e = k - n1;
            const int e2 = (e + 1 == n2) ? 0 : e + 1;
            const T ax = SV(v2 + e2, 1) - SV(v2 + e, 1);  // :778
            const T ay = SV(v2 + e, 0) - SV(v2 + e2, 0);
            const T nrm = sqrt(sq2(ax, ay));
            nx = ax / nrm; ny = ay / nrm;
        }
        T lo1, hi1;
        lo1 = hi1 = dot2(nx, ny, SV(v1, 0), SV(v1, 1));        // :795
        for (int v = 1; v < n1; v++) {                         // :799-804 (v = 0 compares equal: no change)
            const T cur = dot2(nx, ny, SV(v1 + v, 0), SV(v1 + v, 1));
            if (cur < lo1) lo1 = cur;
            if (cur > hi1) hi1 = cur;
        }
        if (!own2) {
            if (fx2) {
                const T *g = sfix + v2 * 6;
                lo2 = hi2 = dot2(nx, ny, g[0], g[1]);          // :796
                for (int v = 1; v < n2; v++) {                 // :806-811
                    const T cur = dot2(nx, ny, g[v * 6], g[v * 6 + 1]);
                    if (cur < lo2) lo2 = cur;
                    if (cur > hi2) hi2 = cur;
                }
            } else {
                lo2 = hi2 = dot2(nx, ny, SV(v2, 0), SV(v2, 1));
                for (int v = 1; v < n2; v++) {
                    const T cur = dot2(nx, ny, SV(v2 + v, 0), SV(v2 + v, 1));
                    if (cur < lo2) lo2 = cur;
                    if (cur > hi2) hi2 = cur;
                }
            }
        }
        if (lo1 > lo2) {                                       // :815-817
            T t = lo1; lo1 = lo2; lo2 = t;
            t = hi1; hi1 = hi2; hi2 = t;
        }
        if (hi1 < lo2) return false;                           // :820-821
        const T cur = hi1 - lo2;                          // :824
        if (cur < overlap) { fxn = nx; fyn = ny; overlap = cur; have = true; }  // :827-830
    }
    if (!have) return false;  // all-NaN axes (reference raises)
    onx = fxn * T(-1.0);         // :838-839 always negated
    ony = fyn * T(-1.0);
    omag = overlap;
    return true;
}

// T = T: the reference's arithmetic, bit for bit.  T = float (FIZZ_MODE_THREAD_FP32): the same operation sequence in
// single precision -- state is converted on load and store (the slab stays FP64), every constant is rounded to float
// once; nothing is bit-compatible with anything, see DESIGN.md for the measured bound.
template <int F, typename T>
__global__ void __launch_bounds__(BLOCK, 3) step_kernel(const __grid_constant__ KParams kp) {
    extern __shared__ double smem[];
    T *sfix = reinterpret_cast<T *>(smem);
    T *sv = sfix + kp.FV * 6;
    int *stab = reinterpret_cast<int *>(sv + (size_t)kp.V * 2 * BLOCK);
    // the parameter block, in the working precision
    const T c_accx = (T)kp.accx, c_accy = (T)kp.accy, c_time_disc = (T)kp.time_disc, c_time_tol = (T)kp.time_tol;
    // (FP32: EPSILON, the spacing a resolve pass adds to its push (:158), is 1e-12 by default -- far below a float's
    //  resolution at scene coordinates (ulp(1000) = 6e-5), where the push would be absorbed and the resolve loop never end:
    //  the single-precision mode raises it to 1e-3 px)
    const T c_eps = sizeof(T) == 4 ? (T)fmax(kp.eps, 1e-3) : (T)kp.eps;
    const T c_coll_tol = (T)kp.coll_tol, c_vib_tol = (T)kp.vib_tol, c_ope = (T)kp.ope, c_ome2 = (T)kp.ome2;
    const T c_mu = (T)kp.mu, c_ffr = (T)kp.ffr;
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
    T px[F], py[F], vx[F], vy[F], sx[F], sy[F], thr[F];
#pragma unroll
    for (int b = 0; b < F; b++) {
        px[b] = kp.pos[(2 * b) * Wp + w];
        py[b] = kp.pos[(2 * b + 1) * Wp + w];
        vx[b] = kp.vel[(2 * b) * Wp + w];
        vy[b] = kp.vel[(2 * b + 1) * Wp + w];
        sx[b] = px[b]; sy[b] = py[b];  // finish_update ran at the end of the previous update() (:237-239)
        thr[b] = kp.thr[b * Wp + w];
    }
    T heat = kp.heat[w];
    long long ncalls_total = 0;

    // per-lane state machine
    bool resolve = false;      // false: step-halving trial (:99-129); true: resolve loop (:140-234)
    bool verts_explicit = false;
    int k = 0;                 // halving trial index; dt = time_disc / 2^k
    T dt = c_time_disc;  // :89
    int ncalls = 0;            // state-relevant check_collisions calls of the current update()
    unsigned mask[F];          // broad-phase survivors: bit (j-i-1) of mask[i] = pair (i, j), j over objs+fixed_objs
#pragma unroll
    for (int b = 0; b < F; b++) mask[b] = 0;

    // contact tuples of the current check_collisions (local memory; touched on contact steps only)
    unsigned char hi_i[FIZZ_MAX_HITS], hi_j[FIZZ_MAX_HITS];
    T hnx[FIZZ_MAX_HITS], hny[FIZZ_MAX_HITS], hmag[FIZZ_MAX_HITS];

    // Point.move (physics.py:502-529) of body b over `h`, starting acceleration (a0x,a0y)
#define FIZZ_MOVE(b, h, a0x, a0y, ox, oy, ovx, ovy)                                                   \
    {                                                                                                 \
        const T halfdt_ = T(.5) * (h);                                                             \
        const T vax_ = vx[b] + (a0x) * halfdt_, vay_ = vy[b] + (a0y) * halfdt_;                  \
        ox = px[b] + vax_ * (h);                                                                      \
        oy = py[b] + vay_ * (h);                                                                      \
        ovx = vax_ + c_accx * halfdt_;                                                               \
        ovy = vay_ + c_accy * halfdt_;                                                               \
    }

    while (true) {
        // acceleration held by every com at the start of this update(): 0 before the first move ever (:494)
        const bool acc_on = step > 0;
        const T a0x = acc_on ? c_accx : T(0), a0y = acc_on ? c_accy : T(0);
        bool need_mask = resolve;
        if (!resolve) {
            if (k == 0) {
                need_mask = true;
                if (step == 0) {
                    // obj.pos aliases com.pos until the first reverse/finish_update (:614): the first
                    // trial of the first step sees the moved position, and it then stays frozen for the step.
#pragma unroll
                    for (int b = 0; b < F; b++) {
                        T tvx, tvy;
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
                    const T dx = sx[i] - sx[j], dy = sy[i] - sy[j];
                    const T s = sq2(dx, dy);
                    if (!(s > thr[i] && s > thr[j])) m |= 1u << (j - i - 1);
                }
                mask[i] = m;
            }
            for (int j = 0; j < X; j++) {
                const T cx = kp.fcx[j], cy = kp.fcy[j], ft = kp.fthr[j];
#pragma unroll
                for (int i = 0; i < F; i++) {
                    const T dx = sx[i] - cx, dy = sy[i] - cy;
                    const T s = sq2(dx, dy);
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
        T maxov = 0;  // :125-129
        if (any) {
            if (!resolve) {
                // pre_update rebuilt every vertex from the moved com (:358-359): |r|cos(phi) + com.x
#pragma unroll
                for (int b = 0; b < F; b++) {
                    T nxp, nyp, tvx, tvy;
                    FIZZ_MOVE(b, dt, a0x, a0y, nxp, nyp, tvx, tvy);
                    (void)tvx; (void)tvy;
                    const int n = kp.nv[b], v0 = kp.vs[b];
                    for (int v = 0; v < n; v++) {
                        SV(v0 + v, 0) = (T)kp.off[(size_t)((v0 + v) * 2) * Wp + w] + nxp;
                        SV(v0 + v, 1) = (T)kp.off[(size_t)((v0 + v) * 2 + 1) * Wp + w] + nyp;
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
                T nx, ny, mag;
                if (sat_pair(sv, sfix, stab, tid, F, ii, jj, nx, ny, mag)) {  // :286-288
                    if (nh < FIZZ_MAX_HITS) {
                        hi_i[nh] = (unsigned char)ii; hi_j[nh] = (unsigned char)jj;
                        hnx[nh] = nx; hny[nh] = ny; hmag[nh] = mag;
                    }
                    nh++;
                    const T am = fabs(mag);
                    if (nh == 1 || am > maxov) maxov = am;
                    if (kp.trace) trace_hit(kp, w, step, ncalls, ii, jj, nx, ny, mag);
                }
            }
            if (nh > FIZZ_MAX_HITS) { status = FIZZ_ST_HIT_OVERFLOW; break; }
        }

        if (!resolve) {
            // ---- :99 loop condition ---------------------------------------------------------------------------
            if (maxov > c_coll_tol && dt > c_time_tol) {
                dt = dt / 2;  // :106 (reverse_update: every trial restarts from the start-of-step com)
                k++;
                continue;
            }
            if (dt < c_time_tol) {
                // :133-137 fallback: com goes back to the start of the step, vertices keep the last trial's
                // positions, so check_collisions() returns the very same list again.
                ncalls++;
                ncalls_total++;
                if (kp.trace)
                    for (int h = 0; h < nh; h++) trace_hit(kp, w, step, ncalls, hi_i[h], hi_j[h], hnx[h], hny[h], hmag[h]);
            } else {
#pragma unroll
                for (int b = 0; b < F; b++) {
                    T nxp, nyp, nvx, nvy;
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
                const T ux = hnx[h], uy = hny[h];
                const T s = hmag[h] + c_eps;         // :158
                const T nvx = ux * s, nvy = uy * s;
                const int n1 = stab[ii], v1 = stab[32 + ii];
                if (jj >= F) {                             // :171-198 free body vs fixed body
                    for (int v = 0; v < n1; v++) { SV(v1 + v, 0) += nvx; SV(v1 + v, 1) += nvy; }  // :174
                    const bool gate = sqrt(sq2(nvx, nvy)) > c_vib_tol;               // :180
                    const T m1 = kp.mass[ii * Wp + w];
#pragma unroll
                    for (int b = 0; b < F; b++) {
                        if (b == ii) {
                            if (gate) { px[b] += nvx; py[b] += nvy; }                            // :181
                            const T vdotN = dot2(vx[b], vy[b], ux, uy);                     // :183
                            const T cc = c_ope * vdotN;                                    // :188
                            vx[b] -= cc * ux;
                            vy[b] -= cc * uy;
                            heat += T(.5) * m1 * c_ome2 * (vdotN * vdotN);                         // :191,:195
                            if (c_mu > T(0)) ext_friction_fixed(c_mu, c_ope, m1, ux, uy, vdotN, vx[b], vy[b], heat);
                            sx[b] = px[b]; sy[b] = py[b];                                        // :197
                        }
                    }
                } else {                                   // :201-223 two free bodies
                    const int n2 = stab[jj], v2 = stab[32 + jj];
                    const T m1 = kp.mass[ii * Wp + w], m2 = kp.mass[jj * Wp + w];
                    const T mtotal = m1 + m2;
                    const T mprop1 = m1 / mtotal;
                    const T mprop2 = 1 - mprop1;
                    T v1x = 0, v1y = 0, v2x = 0, v2y = 0;
#pragma unroll
                    for (int b = 0; b < F; b++) {
                        if (b == ii) { v1x = vx[b]; v1y = vy[b]; }
                        if (b == jj) { v2x = vx[b]; v2y = vy[b]; }
                    }
                    const T vn1 = dot2(v1x, v1y, ux, uy);                                   // :207
                    const T vn2 = dot2(v2x, v2y, ux, uy);                                   // :208
                    T v1u = (mprop1 - mprop2) * vn1 + 2 * mprop2 * vn2;                     // :209
                    T v2u = 2 * mprop1 * vn1 + (mprop2 - mprop1) * vn2;                     // :218
                    if (c_ffr >= T(0)) ext_restitution_free(c_ffr, m1, mprop1, mprop2, vn1, vn2, v1u, v2u, heat);
                    const T d1 = v1u - vn1, d2 = v2u - vn2;                                 // :210,:219
                    v1x += d1 * ux; v1y += d1 * uy;                                              // :210
                    v2x += d2 * ux; v2y += d2 * uy;                                              // :219
                    if (c_mu > T(0))
                        ext_friction_free(c_mu, c_ffr >= T(0) ? c_ffr : T(1), m1, m2, mprop2, ux, uy, vn1, vn2, v1x, v1y, v2x,
                                          v2y, heat);
                    const T p1x = mprop1 * nvx, p1y = mprop1 * nvy;
                    const T p2x = mprop2 * nvx, p2y = mprop2 * nvy;
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
            const T rdt = c_time_disc - dt;
            // acceleration now held by the com: installed by the accepted trial's move, or still the
            // start-of-step one when the fallback reverted it
            const bool fell_back = dt < c_time_tol;
            const T b0x = fell_back ? a0x : c_accx, b0y = fell_back ? a0y : c_accy;
#pragma unroll
            for (int b = 0; b < F; b++) {
                T nxp, nyp, nvx, nvy;
                FIZZ_MOVE(b, rdt, b0x, b0y, nxp, nyp, nvx, nvy);
                px[b] = nxp; py[b] = nyp; vx[b] = nvx; vy[b] = nvy;
            }
            verts_explicit = false;                        // pre_update rebuilt the vertices from the com
        }
#pragma unroll
        for (int b = 0; b < F; b++) { sx[b] = px[b]; sy[b] = py[b]; }  // finish_update (:237-239, :251)
        step++;                                            // :259
        resolve = false; k = 0; dt = c_time_disc; ncalls = 0;
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
                    kp.verts[r0] = (T)kp.off[r0] + px[b];
                    kp.verts[r1] = (T)kp.off[r1] + py[b];
                }
            }
        }
    }
}


#undef SV

}  // namespace fizz
