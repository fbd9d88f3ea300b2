// fizz_ballistic_kernel.cuh -- thread-per-world kernel for CONTACT-FREE timesteps ("hybrid" mode, stage 1).
//
// A timestep whose broad phase (physics.py:272-281) lets no pair through is: one check_collisions() that returns
// [], i.e. max overlap 0 on the first halving trial (:99-129), then finish_update (:237-239).  State-wise that is
// a velocity-Verlet move of every centre of mass (Point.move :502-529) and nothing else -- vertices are derived
// state.  This kernel advances each world through such steps with com position/velocity and the
// per-body thresholds in registers (F is a template parameter: everything is unrolled, fixed-body constants come
// straight from the kernel-parameter constant bank).  When the broad phase is not empty, every candidate pair must be
// PROVEN contact-free (prove_separated below: exact SAT on the fixed partner's precomputed axes, a rigorous margin
// test on free axes); the world STOPS at the first step where that fails and the warp-per-world contact kernel picks
// it up from there (fizz_contact_kernel.cuh).
// The broad phase of step t reads the snapshot obj.pos written at the end of step t-1, which at a
// step boundary equals com.pos, so no separate snapshot is kept here.  Step 0 is never taken here: before the
// first finish_update obj.pos aliases com.pos and the broad phase sees the MOVED position (SURVEY Q4 exception).
#pragma once

#include "fizz_common.cuh"

namespace fizz {

constexpr int BALLISTIC_BLOCK = 128;

// Shared memory of the ballistic kernel: fixed scene [FV*6], moved centres of mass [2F][BLOCK] (lane column),
// optionally the world's polar offsets [2V][BLOCK] (when they fit in 32 KB), body tables.
__host__ __device__ inline bool ballistic_caches_offsets(int V) { return (size_t)V * 2 * BALLISTIC_BLOCK * 8 <= 32 * 1024; }
__host__ __device__ inline size_t ballistic_smem_bytes(int F, int V, int FV) {
    size_t d = (size_t)FV * 6 + (size_t)2 * F * BALLISTIC_BLOCK + (ballistic_caches_offsets(V) ? (size_t)2 * V * BALLISTIC_BLOCK : 0);
    return d * 8 + 64 * sizeof(int);
}

// "Provably no contact" for the candidate pair (ii free, jj free-or-fixed) with the vertices the first halving trial
// would see (|r|cos(phi) + moved com, :358-359).  Two kinds of proof, both imply that polypoly_collision (:765-840)
// returns ([], None) for the pair:
//   P1 (exact)  an edge normal of the FIXED partner separates: its unit normal and own projection bounds are the
//               precomputed bit-identical constants, the free body's projections fma(ny,y,nx*x) are computed exactly
//               as :799-804 does, and :815-821 is applied literally;
//   P2 (margin) an edge normal a of a FREE body separates the unnormalised projections a.v by more than
//               1e-9 * M * (|ax|+|ay|), M >= |x|+|y| of every vertex involved.  The reference normalises a first
//               (:790) and projects with rounding; both effects are bounded by 16 * 2^-53 * M per unit of |a|, seven
//               orders of magnitude below the margin, so its test `hi < lo` on that axis is true as well.
// If neither proof works the caller hands the world to the contact kernel, which runs the exact narrow phase.
__device__ __noinline__ bool prove_separated(const double *sfix, const double *snp, const int *stab,
                                                const double *offp, long long ostride, long long obase, int tid, int F,
                                                int ii, int jj, double Mi, double Mj) {
    const int n1 = stab[ii], v1 = stab[32 + ii];
    const int n2 = stab[jj], v2 = stab[32 + jj];
    FIZZ_ASSERT(ii >= 0 && ii < F && jj > ii && jj < TAB && n1 >= 3 && n2 >= 3);
    const double c1x = snp[(2 * ii) * BALLISTIC_BLOCK + tid], c1y = snp[(2 * ii + 1) * BALLISTIC_BLOCK + tid];
#define FZ_V1X(v) (offp[(long long)(2 * (v1 + (v))) * ostride + obase] + c1x)
#define FZ_V1Y(v) (offp[(long long)(2 * (v1 + (v)) + 1) * ostride + obase] + c1y)
    const double margin = 1e-9 * (Mi + Mj);
    if (jj >= F) {
        for (int k = 0; k < n2; k++) {  // P1: the fixed partner's axes
            const double *g = sfix + (v2 + k) * 6;
            const double nx = g[2], ny = g[3], lo2 = g[4], hi2 = g[5];
            double lo1, hi1;
            lo1 = hi1 = dot2(nx, ny, FZ_V1X(0), FZ_V1Y(0));
            for (int v = 1; v < n1; v++) {
                const double c = dot2(nx, ny, FZ_V1X(v), FZ_V1Y(v));
                if (c < lo1) lo1 = c;
                if (c > hi1) hi1 = c;
            }
            if ((lo1 > lo2) ? (hi2 < lo1) : (hi1 < lo2)) return true;   // :815-821
        }
        for (int k = 0; k < n1; k++) {  // P2: the free body's axes against the fixed vertices
            const int k2 = (k + 1 == n1) ? 0 : k + 1;
            const double ax = FZ_V1Y(k2) - FZ_V1Y(k), ay = FZ_V1X(k) - FZ_V1X(k2);
            double loA = INFINITY, hiA = -INFINITY, loB = INFINITY, hiB = -INFINITY;
            for (int v = 0; v < n1; v++) {
                const double c = fma(ay, FZ_V1Y(v), ax * FZ_V1X(v));
                loA = fmin(loA, c); hiA = fmax(hiA, c);
            }
            for (int v = 0; v < n2; v++) {
                const double c = fma(ay, sfix[(v2 + v) * 6 + 1], ax * sfix[(v2 + v) * 6]);
                loB = fmin(loB, c); hiB = fmax(hiB, c);
            }
            if (fmax(loB - hiA, loA - hiB) > margin * (fabs(ax) + fabs(ay))) return true;
        }
        return false;
    }
    const double c2x = snp[(2 * jj) * BALLISTIC_BLOCK + tid], c2y = snp[(2 * jj + 1) * BALLISTIC_BLOCK + tid];
#define FZ_V2X(v) (offp[(long long)(2 * (v2 + (v))) * ostride + obase] + c2x)
#define FZ_V2Y(v) (offp[(long long)(2 * (v2 + (v)) + 1) * ostride + obase] + c2y)
    for (int k = 0; k < n1 + n2; k++) {  // P2 on both free bodies' axes
        double ax, ay;
        if (k < n1) {
            const int k2 = (k + 1 == n1) ? 0 : k + 1;
            ax = FZ_V1Y(k2) - FZ_V1Y(k); ay = FZ_V1X(k) - FZ_V1X(k2);
        } else {
            const int e = k - n1, e2 = (e + 1 == n2) ? 0 : e + 1;
            ax = FZ_V2Y(e2) - FZ_V2Y(e); ay = FZ_V2X(e) - FZ_V2X(e2);
        }
        double loA = INFINITY, hiA = -INFINITY, loB = INFINITY, hiB = -INFINITY;
        for (int v = 0; v < n1; v++) {
            const double c = fma(ay, FZ_V1Y(v), ax * FZ_V1X(v));
            loA = fmin(loA, c); hiA = fmax(hiA, c);
        }
        for (int v = 0; v < n2; v++) {
            const double c = fma(ay, FZ_V2Y(v), ax * FZ_V2X(v));
            loB = fmin(loB, c); hiB = fmax(hiB, c);
        }
        if (fmax(loB - hiA, loA - hiB) > margin * (fabs(ax) + fabs(ay))) return true;
    }
    return false;
#undef FZ_V1X
#undef FZ_V1Y
#undef FZ_V2X
#undef FZ_V2Y
}

// PROVER = false: the lean variant (registers only, no shared memory): stops at the first step with a candidate pair.
// PROVER = true : candidates are put through prove_separated; stops at the first step where a contact is possible.
// Scheduling hint `tier` (performance only): 0 = quiescent (lean variant), 1 = candidates but no contact (prover
// variant), 2 = contact phase (contact kernel).  Each kernel demotes / promotes the worlds it touches.
template <int F, bool PROVER>
__global__ void __launch_bounds__(BALLISTIC_BLOCK)
ballistic_kernel(const __grid_constant__ KParams kp, const long long *__restrict__ list,
                 const unsigned long long *__restrict__ count_ptr) {
    extern __shared__ double smem_b[];
    const int tid = threadIdx.x;
    const int X = kp.X, V = kp.V;
    const bool cache_off = PROVER && ballistic_caches_offsets(V);
    double *sfix = smem_b;
    double *snp = smem_b + kp.FV * 6;
    double *soff = snp + 2 * F * BALLISTIC_BLOCK;
    int *stab = reinterpret_cast<int *>(soff + (cache_off ? 2 * V * BALLISTIC_BLOCK : 0));
    if (PROVER) {
        for (int i = tid; i < kp.FV * 6; i += BALLISTIC_BLOCK) sfix[i] = kp.fixed_geom[i];
        if (tid < F + X) { stab[tid] = kp.nv[tid]; stab[32 + tid] = kp.vs[tid]; }
        __syncthreads();
    }
    // threads map to the compacted live list (ascending world index inside a warp: rows stay mostly coalesced)
    const long long t = (long long)blockIdx.x * BALLISTIC_BLOCK + tid;
    if (t >= (long long)*count_ptr) return;
    const long long w = list[t];
    FIZZ_ASSERT(w >= 0 && w < kp.W);
    const long long step0 = kp.steps[w];
    const long long Wp = kp.Wp;
    double px[F], py[F], vx[F], vy[F], thr[F];
#pragma unroll
    for (int b = 0; b < F; b++) {
        px[b] = kp.pos[(2 * b) * Wp + w];
        py[b] = kp.pos[(2 * b + 1) * Wp + w];
        vx[b] = kp.vel[(2 * b) * Wp + w];
        vy[b] = kp.vel[(2 * b + 1) * Wp + w];
        thr[b] = kp.thr[b * Wp + w];
    }
    if (cache_off)
        for (int r = 0; r < 2 * V; r++) soff[r * BALLISTIC_BLOCK + tid] = kp.off[(size_t)r * Wp + w];
    const double *offp = cache_off ? soff : kp.off;
    const long long ostride = cache_off ? BALLISTIC_BLOCK : Wp, obase = cache_off ? tid : w;
    const double dt = kp.time_disc;
    long long step = step0;
    // CONSERVATIVE threshold test on the integer pipe.  s = fma(dy,dy,dx*dx) is >= +0.0 (or NaN), and for
    // non-negative doubles the bit patterns order like integers, so  hi32(s) > hi32(T)  implies  s > T.  A pair is
    // counted as "skipped by the broad phase" only when that holds for both thresholds; anything else -- a pair
    // that passes, or one within 2^-20 of its threshold -- becomes a CANDIDATE that must be proven contact-free
    // (prove_separated) or the world goes to the contact kernel, which repeats the step with the exact tests.
    // The FP64 pipe is left with the 4 arithmetic ops per pair.  A NaN/inf coordinate (whose pairs the reference does
    // NOT skip: `nan > r` is False, physics.py:278) is caught by the exponent test.
    int thr_h[F];
#pragma unroll
    for (int b = 0; b < F; b++) thr_h[b] = __double2hiint(thr[b]);
    bool handed_over = false;  // stopped because a later tier has to look at the step
    int empty_streak = 0;      // consecutive steps without any candidate pair
    if constexpr (!PROVER) {
        // ---- lean variant: integer-pipe threshold tests folded into one predicate, in-place move ---------------------
        // Non-finite coordinates (whose pairs the reference does NOT skip: `nan > r` is False, physics.py:278) cannot
        // appear inside this launch when every |pos|, |vel dt| and |acc dt^2| starts below 2^993: n steps move a coordinate
        // by at most n |v dt| + n^2 |a dt^2|, and no launch makes 2^20 steps.  Anything larger (or non-finite already)
        // goes to the contact kernel, which tests every step; so the loop below carries no exponent test.
        {
            unsigned big = 0;
#pragma unroll
            for (int b = 0; b < F; b++) {
                big |= (((unsigned)__double2hiint(px[b]) & 0x7ff00000u) >= 0x7e000000u) | (((unsigned)__double2hiint(py[b]) & 0x7ff00000u) >= 0x7e000000u);
                big |= (((unsigned)__double2hiint(vx[b]) & 0x7ff00000u) >= 0x7e000000u) | (((unsigned)__double2hiint(vy[b]) & 0x7ff00000u) >= 0x7e000000u);
            }
            big |= (((unsigned)__double2hiint(kp.accx) & 0x7ff00000u) >= 0x7e000000u) | (((unsigned)__double2hiint(kp.accy) & 0x7ff00000u) >= 0x7e000000u);
            if (big) handed_over = true;
        }
        // One test per free body for ALL its pairs with fixed bodies.  With (hb, hh) the centre and half extents of the box
        // around the fixed centres and R >= max(r_i, r_j) for every fixed j:  |x_i - hb_x| > hh_x + R  implies
        // |x_i - c_j,x| > R for every j, hence norm(d) > max(r_i, r_j): the reference skips each of those pairs (:277-279),
        // and so does the exact test on fma(dy, dy, dx dx) (a margin of 2^-17 on R covers every rounding on the way, and
        // the one-unit gap in the high word that the tests below ask for).  Same for y.  Compared on the integer pipe.
        int qx_h[F], qy_h[F];
        {
            const double grow = 1.0 + 1.0 / 131072.0;
#pragma unroll
            for (int b = 0; b < F; b++) {
                const double r = sqrt(thr[b]);                          // radius of body b, to an ulp
                const double R = (r > kp.frmax) ? r : kp.frmax;         // (NaN radius: the comparisons below fail, no reject)
                qx_h[b] = __double2hiint((kp.hhx + R) * grow + 1e-6) + 1;   // (+1e-6 px: the hull's own rounding)
                qy_h[b] = __double2hiint((kp.hhy + R) * grow + 1e-6) + 1;
                if (!(R == R) || !(R < 1e300)) { qx_h[b] = 0x7fffffff; qy_h[b] = 0x7fffffff; }
            }
        }
        // the broad phase of one timestep: does any pair pass (conservatively)?
        auto candidates = [&]() -> bool {
            bool any = false;
            bool near_fixed = false;
#pragma unroll
            for (int i = 0; i < F; i++) {
                const int ex = __double2hiint(px[i] - kp.hbx) & 0x7fffffff, ey = __double2hiint(py[i] - kp.hby) & 0x7fffffff;
                near_fixed |= !((ex > qx_h[i]) | (ey > qy_h[i]));
            }
#pragma unroll
            for (int i = 0; i < F; i++) {
#pragma unroll
                for (int j = i + 1; j < F; j++) {
                    const double dx = px[i] - px[j], dy = py[i] - py[j];
                    const int sh = __double2hiint(sq2(dx, dy));
                    any |= !(sh > max(thr_h[i], thr_h[j]));
                }
            }
            if (near_fixed) {
                for (int j = 0; j < X; j++) {
                    const double cx = kp.fcx[j], cy = kp.fcy[j];
                    const int fh = __double2hiint(kp.fthr[j]);
#pragma unroll
                    for (int i = 0; i < F; i++) {
                        const double dx = px[i] - cx, dy = py[i] - cy;
                        const int sh = __double2hiint(sq2(dx, dy));
                        any |= !(sh > max(thr_h[i], fh));
                    }
                }
            }
            return any;
        };
        auto full_move = [&]() {
#pragma unroll
            for (int b = 0; b < F; b++) {
                double nx_, ny_, nvx_, nvy_;
                FIZZ_MOVE_POINT(px[b], py[b], vx[b], vy[b], dt, kp.accx, kp.accy, kp.accx, kp.accy, nx_, ny_, nvx_, nvy_);
                px[b] = nx_; py[b] = ny_; vx[b] = nvx_; vy[b] = nvy_;
            }
        };
        // Without gravity Point.move (:502-529) is v + 0.0, x + v dt, v + 0.0: after ONE complete move a velocity is no
        // longer -0.0 (the only value v + 0.0 changes), so every further move of the launch is x += v dt, bit for bit.
        const bool nograv = (kp.accx == 0.0) && (kp.accy == 0.0);
        if (!handed_over && step < kp.target_steps) {
            if (candidates()) handed_over = true;
            else { full_move(); step++; }
        }
        if (!handed_over) {
            if (nograv) {
                while (step < kp.target_steps) {
                    if (candidates()) { handed_over = true; break; }  // candidates: the next tier looks at this step
#pragma unroll
                    for (int b = 0; b < F; b++) { px[b] = px[b] + vx[b] * dt; py[b] = py[b] + vy[b] * dt; }
                    step++;
                }
            } else {
                while (step < kp.target_steps) {
                    if (candidates()) { handed_over = true; break; }
                    full_move();
                    step++;
                }
            }
        }
    } else {
        while (step < kp.target_steps) {
            unsigned cand[F];  // bit (j-i-1) of cand[i]: pair (i, j), j over objs+fixed_objs
    #pragma unroll
            for (int i = 0; i < F; i++) {
                unsigned m = 0;
    #pragma unroll
                for (int j = i + 1; j < F; j++) {
                    const double dx = px[i] - px[j], dy = py[i] - py[j];
                    const int sh = __double2hiint(sq2(dx, dy));
                    m |= (sh > max(thr_h[i], thr_h[j])) ? 0u : (1u << (j - i - 1));
                }
                cand[i] = m;
            }
            for (int j = 0; j < X; j++) {
                const double cx = kp.fcx[j], cy = kp.fcy[j];
                const int fh = __double2hiint(kp.fthr[j]);
    #pragma unroll
                for (int i = 0; i < F; i++) {
                    const double dx = px[i] - cx, dy = py[i] - cy;
                    const int sh = __double2hiint(sq2(dx, dy));
                    cand[i] |= (sh > max(thr_h[i], fh)) ? 0u : (1u << (F - i - 1 + j));
                }
            }
            // non-finite coordinates: exponent field all ones (integer pipe); such a world goes to the contact kernel
            unsigned expo = 0, any = 0;
    #pragma unroll
            for (int b = 0; b < F; b++) {
                const unsigned hx = (unsigned)__double2hiint(px[b]) & 0x7ff00000u, hy = (unsigned)__double2hiint(py[b]) & 0x7ff00000u;
                expo |= (hx == 0x7ff00000u) | (hy == 0x7ff00000u);
                any |= cand[b];
            }
            if (expo) { handed_over = true; break; }
            // the first halving trial's move (every com holds the installed acceleration once step 0 is over)
            double npx[F], npy[F], nvx[F], nvy[F];
    #pragma unroll
            for (int b = 0; b < F; b++)
                FIZZ_MOVE_POINT(px[b], py[b], vx[b], vy[b], dt, kp.accx, kp.accy, kp.accx, kp.accy, npx[b], npy[b], nvx[b], nvy[b]);
            empty_streak = any ? 0 : empty_streak + 1;
            if (any) {
                // candidates: every one of them must be proven contact-free with the moved vertices, else hand over
    #pragma unroll
                for (int b = 0; b < F; b++) {
                    snp[(2 * b) * BALLISTIC_BLOCK + tid] = npx[b];
                    snp[(2 * b + 1) * BALLISTIC_BLOCK + tid] = npy[b];
                }
                bool unproven = false;
                while (true) {
                    int ii = -1, jj = 0;
                    double Mi = 0.0;
    #pragma unroll
                    for (int b = 0; b < F; b++) {
                        if (ii < 0 && cand[b]) {
                            const int bit = __ffs(cand[b]) - 1;
                            cand[b] &= cand[b] - 1;
                            ii = b; jj = b + 1 + bit;
                            Mi = fabs(npx[b]) + fabs(npy[b]) + 2.0 * (1.0 + 1.0000001 * thr[b]);  // radius <= 1 + radius^2
                        }
                    }
                    if (ii < 0) break;
                    double Mj = 0.0;
                    if (jj >= F) {
                        Mj = fabs(kp.fcx[jj - F]) + fabs(kp.fcy[jj - F]) + 2.0 * (1.0 + 1.0000001 * kp.fthr[jj - F]);
                    } else {
    #pragma unroll
                        for (int b = 0; b < F; b++)
                            if (b == jj) Mj = fabs(npx[b]) + fabs(npy[b]) + 2.0 * (1.0 + 1.0000001 * thr[b]);
                    }
                    if (!prove_separated(sfix, snp, stab, offp, ostride, obase, tid, F, ii, jj, Mi, Mj)) { unproven = true; break; }
                }
                if (unproven) { handed_over = true; break; }  // a contact is possible: the contact kernel takes over here
            }
    #pragma unroll
            for (int b = 0; b < F; b++) { px[b] = npx[b]; py[b] = npy[b]; vx[b] = nvx[b]; vy[b] = nvy[b]; }
            step++;
        }
    }
    if (step != step0) {
#pragma unroll
        for (int b = 0; b < F; b++) {
            kp.pos[(2 * b) * Wp + w] = px[b];
            kp.pos[(2 * b + 1) * Wp + w] = py[b];
            kp.vel[(2 * b) * Wp + w] = vx[b];
            kp.vel[(2 * b + 1) * Wp + w] = vy[b];
        }
        kp.steps[w] = step;
        kp.calls[w] += step - step0;  // one state-relevant check_collisions() per contact-free step
        kp.vflag[w] = 0;              // vertices are derived: off + com
        atomicAdd(kp.counters + 0, (unsigned long long)(step - step0));  // compiler warp-aggregates (REDUX)
    }
    if (PROVER) kp.tier[w] = handed_over ? 2 : (empty_streak >= 8 ? 0 : 1);
    else if (handed_over) kp.tier[w] = 1;
}

// Live list of a ballistic launch: running worlds past step 0, short of the target, with a hint <= max_tier.
__global__ void live_kernel(const long long *status, const long long *steps, const long long *tier, long long W,
                            long long target, long long max_tier, long long *list, unsigned long long *count) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = (w < W) && status[w] == 0 && (tier[w] & 255) <= max_tier && steps[w] > 0 && steps[w] < target;
    const unsigned b = __ballot_sync(0xffffffffu, live);
    if (b == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) list[base + __popc(b & ((1u << lane) - 1u))] = w;
}

}  // namespace fizz
