// fizz_ballistic_kernel.cuh -- thread-per-world kernel for QUIESCENT timesteps ("hybrid" mode, stage 1).
//
// A timestep whose broad phase (physics.py:272-281) lets no pair through is: one check_collisions() that returns
// [], i.e. max overlap 0 on the first halving trial (:99-129), then finish_update (:237-239).  State-wise that is
// a velocity-Verlet move of every centre of mass (Point.move :502-529) and nothing else -- vertices are derived
// state.  This kernel advances each world through such steps with com position/velocity and the
// per-body thresholds in registers (F is a template parameter: everything is unrolled, fixed-body constants come
// straight from the kernel-parameter constant bank), and STOPS a world at the first step whose broad phase is
// not empty; the warp-per-world contact kernel picks those worlds up from there (fizz_contact_kernel.cuh).
// The broad phase of step t reads the snapshot obj.pos written at the end of step t-1, which at a
// step boundary equals com.pos, so no separate snapshot is kept here.  Step 0 is never taken here: before the
// first finish_update obj.pos aliases com.pos and the broad phase sees the MOVED position (SURVEY Q4 exception).
#pragma once

#include "fizz_common.cuh"

namespace fizz {

constexpr int BALLISTIC_BLOCK = 128;

template <int F>
__global__ void __launch_bounds__(BALLISTIC_BLOCK)
ballistic_kernel(const __grid_constant__ KParams kp, const long long *__restrict__ list,
                 const unsigned long long *__restrict__ count_ptr) {
    // threads map to the compacted live list (ascending world index inside a warp: rows stay mostly coalesced)
    const long long t = (long long)blockIdx.x * BALLISTIC_BLOCK + threadIdx.x;
    if (t >= (long long)*count_ptr) return;
    const long long w = list[t];
    const long long step0 = kp.steps[w];
    const long long Wp = kp.Wp;
    const int X = kp.X;
    double px[F], py[F], vx[F], vy[F], thr[F];
#pragma unroll
    for (int b = 0; b < F; b++) {
        px[b] = kp.pos[(2 * b) * Wp + w];
        py[b] = kp.pos[(2 * b + 1) * Wp + w];
        vx[b] = kp.vel[(2 * b) * Wp + w];
        vy[b] = kp.vel[(2 * b + 1) * Wp + w];
        thr[b] = kp.thr[b * Wp + w];
    }
    const double dt = kp.time_disc;
    long long step = step0;
    // CONSERVATIVE threshold test on the integer pipe.  s = fma(dy,dy,dx*dx) is >= +0.0 (or NaN), and for
    // non-negative doubles the bit patterns order like integers, so  hi32(s) > hi32(T)  implies  s > T.  A pair is
    // counted as "skipped by the broad phase" only when that holds for both thresholds; anything else -- a pair
    // that passes, or one within 2^-20 of its threshold -- ends the ballistic run and the contact kernel repeats the
    // step with the exact test.  The FP64 pipe is left with the 4 arithmetic ops per pair.  A NaN/inf coordinate
    // (whose pairs the reference does NOT skip: `nan > r` is False, physics.py:278) is caught by the exponent test.
    int thr_h[F];
#pragma unroll
    for (int b = 0; b < F; b++) thr_h[b] = __double2hiint(thr[b]);
    while (step < kp.target_steps) {
        bool any = false;
#pragma unroll
        for (int i = 0; i < F; i++) {
#pragma unroll
            for (int j = i + 1; j < F; j++) {
                const double dx = px[i] - px[j], dy = py[i] - py[j];
                const int sh = __double2hiint(fma(dy, dy, dx * dx));
                any |= !(sh > max(thr_h[i], thr_h[j]));
            }
        }
        for (int j = 0; j < X; j++) {
            const double cx = kp.fcx[j], cy = kp.fcy[j];
            const int fh = __double2hiint(kp.fthr[j]);
#pragma unroll
            for (int i = 0; i < F; i++) {
                const double dx = px[i] - cx, dy = py[i] - cy;
                const int sh = __double2hiint(fma(dy, dy, dx * dx));
                any |= !(sh > max(thr_h[i], fh));
            }
        }
        // non-finite coordinates: exponent field all ones (integer pipe); such a world goes to the contact kernel
        unsigned expo = 0;
#pragma unroll
        for (int b = 0; b < F; b++) {
            const unsigned hx = (unsigned)__double2hiint(px[b]) & 0x7ff00000u, hy = (unsigned)__double2hiint(py[b]) & 0x7ff00000u;
            expo |= (hx == 0x7ff00000u) | (hy == 0x7ff00000u);
        }
        any |= expo != 0;
        if (any) break;  // a pair reaches the narrow phase: hand the world to the contact kernel at this step
#pragma unroll
        for (int b = 0; b < F; b++) {
            double nx_, ny_, nvx_, nvy_;
            // every com holds the installed acceleration once step 0 is over
            FIZZ_MOVE_POINT(px[b], py[b], vx[b], vy[b], dt, kp.accx, kp.accy, kp.accx, kp.accy, nx_, ny_, nvx_, nvy_);
            px[b] = nx_; py[b] = ny_; vx[b] = nvx_; vy[b] = nvy_;
        }
        step++;
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
        kp.calls[w] += step - step0;  // one state-relevant check_collisions() per quiescent step
        kp.vflag[w] = 0;              // vertices are derived: off + com
        atomicAdd(kp.counters + 0, (unsigned long long)(step - step0));  // compiler warp-aggregates (REDUX)
    } else {
        kp.tier[w] = 1;  // no progress: the world is in a contact phase; the contact kernel clears the hint
    }
}

// Live list of the ballistic kernel: running worlds past step 0, short of the target, not hinted "contact phase".
__global__ void live_kernel(const long long *status, const long long *steps, const long long *tier, long long W,
                            long long target, long long *list, unsigned long long *count) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = (w < W) && status[w] == 0 && tier[w] == 0 && steps[w] > 0 && steps[w] < target;
    const unsigned b = __ballot_sync(0xffffffffu, live);
    if (b == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) list[base + __popc(b & ((1u << lane) - 1u))] = w;
}

}  // namespace fizz
