// fizz_common.cuh -- shared declarations of the sm_100a stepping kernels.
//
// Reference semantics (file:line = /root/reference/src/physics.py): World.update :76-259, check_collisions
// :262-289, polypoly_collision :765-840, Obj.pre_update :384-427, Point.move :502-529, Obj.reverse_update
// :430-451, Obj.finish_update :453-464.  numpy's 2-vector dot/norm fuse the second product:
// dot = fma(a1,b1,a0*b0), norm = sqrt(fma(a1,a1,a0*a0)).  Everything is compiled with -fmad=false so that no other
// a*b+c is contracted; IEEE-754 division and square root are the CUDA defaults for double.
#pragma once

#include "fizz_b200.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>

namespace fizz {

constexpr int TAB = FIZZ_MAX_FREE + FIZZ_MAX_FIXED;  // 24 bodies max
constexpr int MAX_PAIRS = FIZZ_MAX_FREE * (FIZZ_MAX_FREE - 1) / 2 + FIZZ_MAX_FREE * FIZZ_MAX_FIXED;  // 156

struct KParams {
    // slab sections (row r of a section starts at ptr + r*Wp; element w of a row is world w)
    double *pos, *vel, *heat;
    long long *status, *steps, *calls, *vflag, *tier;
    const double *off, *thr, *mass;
    double *verts;
    const double *fixed_geom;  // [FV][6]: x, y, unit normal of edge k->k+1 (nx, ny), own projection (lo, hi)
    unsigned long long *counters;  // [0] world-steps advanced by the ballistic kernel, [1] by the contact kernel,
                                   // [2] check_collisions() made by the contact kernel, [3] of those in the fast path,
                                   // [4..7] spec rounds: timesteps accepted, evaluated, rounds, rounds with a rejection
    long long *general_calls;      // optional, per world: check_collisions() made OUTSIDE the single-contact fast path
    FizzContact *trace;
    unsigned long long *trace_count;
    long long trace_cap;
    long long W, Wp, target_steps;
    int F, X, V, FV;
    int swl;                   // log2 of the narrow phase's slot width: 8, 16 or 32 lanes >= the largest n1 + n2 of the scene
    int nv[TAB], vs[TAB];      // vertices per body / first vertex (free: into the world's rows, fixed: into fixed_geom)
    double fcx[FIZZ_MAX_FIXED], fcy[FIZZ_MAX_FIXED], fthr[FIZZ_MAX_FIXED];
    // hull of the fixed bodies' centres (centre, half extents) and their largest radius: one conservative test per free
    // body rejects all its pairs with fixed bodies at once (fizz_ballistic_kernel.cuh)
    double hbx, hby, hhx, hhy, frmax;
    double coll_tol, time_tol, eps, vib_tol, time_disc, ope, ome2, accx, accy;
    double mu, ffr;            // extensions (fizz_ext.cuh): Coulomb coefficient (0 = off), free-free restitution (< 0 = off)
    double vib_thr;            // T(vib_tol): norm(nv) > VIBRATE_TOL  <=>  fma(nvy,nvy,nvx*nvx) > vib_thr  (:180)
    int max_calls;
};

// ---------------------------------------------------------------------------------------------------------------
// Time-parallel stepping of expensive worlds ("spec rounds"; fizz_spec.cuh, fizz_engine.cu: fizz_step).
// A world is a sequential chain (update() t+1 needs the state update() t left), but the state a trapped world starts its
// next timestep from is almost always PREDICTABLE without running the timestep: the trapped coordinate comes back to
// the same value, everything else moved ballistically.  So K warps evaluate update() t, t+L, t+2L, ... of one world
// side by side (a ROUND), warp k starting from the state predicted for timestep t + kL, and a commit step accepts the
// longest prefix of slots whose START state equals, bit for bit, the END state of the slot before (slot 0 starts from the
// real state).  An accepted slot ran the real update() on the real state, so its result is the reference's; a
// mispredicted slot is thrown away and its timesteps are evaluated again in the next round.  Results never depend on the
// predictor, the window, or any other hint below.
//
// Slot record (doubles / 64-bit words), one per (list entry, slot), SPEC_K slots per list entry:
//   [0, 4F)   start state: x, y, vx, vy of every free body      [4F, 8F) end state
//   [8F + 0]  flags: bit 0 "ran", bit 1 "clean" (status 0, heat bit-identical to the world's, vertices derived)
//   [8F + 1]  timesteps advanced   [8F + 2] check_collisions() made   [8F + 3] ... of those in the fast path
//   [8F + 4]  check_collisions() of the last completed timestep        [8F + 5] heat   [8F + 6] status   [8F + 7] vflag
//   [8F + 8]  cost per timestep of the item (see the tier word)
// Control record per WORLD (hints and statistics only; survives chunks and calls; a stale one costs a misprediction):
//   [0] 2 = the classes below are a usable predictor   [1] [2] predictor class of every position coordinate (one byte
//   each: 0 unchanged, 1 Point.move over TIME_DISC, 2 Point.move over the remainder of a timestep that fell back, 3
//   unknown)   [3] consecutive rounds whose last accepted slot no move explains   [4] items of the current round that
//   are still out (the warp that finishes the last one commits)   statistics: [5] timesteps accepted  [6] timesteps
//   evaluated  [7] rounds that accepted fewer slots than ran  [8] rounds without a predictor  [9] rounds   [10] 16 x the
//   running estimate of the timesteps a round gets accepted (spec_shape)
struct SpecParams {
    double *slots;
    long long *ctrl;
    int S;                      // window (timesteps per round) of a world without history
    int D;                      // cost per item the slot length aims at: L ~ D / cost of a timestep (spec_shape)
    int budget_cost;            // a world without a usable predictor: slot 0 alone, until it has cost this much
    int budget_steps;           // ... or made this many timesteps, per round
    int enabled;
    // work queue of the rounds: tickets.  state[0] list length, [1] head (next ticket to consume), [2] tail (next ticket to
    // produce), [3] worlds in flight.  Entry of ticket t, at queue[t % qcap]: (t + 1) << 25 | list entry << 5 | slot.
    unsigned long long *queue, *state;
    unsigned long long qcap;
    // launches that only run when the list *gate points at is long / short (contact_kernel: runner bits 1, 2)
    const unsigned long long *gate;
    int abort_x4;               // a predicted slot is given up at abort_x4 / 4 times its expected cost
    int nopred;                 // (debugging: never use a predictor)
    int round;                  // (debug statistics only)
    unsigned long long *dbg;    // optional [chunks][4]: longest item (cycles), sum, items, items given up
};
constexpr int SPEC_K = 32;          // slots per list entry (stride of the item index; a world uses K <= SPEC_K of them)
constexpr int SPEC_CTRL_WORDS = 12;
__host__ __device__ __forceinline__ int spec_slot_words(int F) { return 8 * F + 10; }
// The tier word of a world: bits 0..7 scheduling tier, 8..31 check_collisions() of its last completed timestep, 32..55
// COST of that timestep = the clock cycles the warp spent on it / 256 (about one call of the register-resident path).
// Slot shape of a world for the spec rounds:
// The WINDOW of a world (timesteps a round tries to cover) follows how far its predictions have been holding: a16 is 16 x
// a running estimate of the timesteps a round gets accepted (control record [10]; 0 = no history: S).  Events -- the
// trapped body reaching the other wall of its cage: one coordinate jumps, one velocity flips -- end a prediction; a world
// with an event every 8 timesteps gets short windows of single-timestep slots, one with an event every 90 long ones.
// L = timesteps per slot: D / cost (items of about D cost units), but at least window / SPEC_K; K = slots.
__host__ __device__ __forceinline__ void spec_shape(long long tier_word, long long a16, int S, int D, int &K, int &L) {
    long long cost = (tier_word >> 32) & 0xffffff;
    if (cost < 1) cost = 1;
    if (a16 <= 0) a16 = 16LL * S;
    long long window = (3 * a16 + 31) / 32;                 // 1.5 x the estimate
    window = window < 4 ? 4 : (window > 16 * SPEC_K ? 16 * SPEC_K : window);
    const long long lmin = (window + SPEC_K - 1) / SPEC_K;
    long long l = D / cost;
    l = l < lmin ? lmin : (l > 16 ? 16 : l);
    L = (int)l;
    K = (int)((window + l - 1) / l);
    if (K > SPEC_K) K = SPEC_K;
    if (K < 2) K = 2;
}

// Debug builds (-DFIZZ_DEBUG): index checks that trap instead of corrupting memory (compute-sanitizer is closed on
// the development pool).  Release builds compile them out.
#ifdef FIZZ_DEBUG
#define FIZZ_ASSERT(cond)                                                                  \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            printf("FIZZ_ASSERT failed: %s  (%s:%d)\n", #cond, __FILE__, __LINE__);        \
            __trap();                                                                      \
        }                                                                                  \
    } while (0)
#else
#define FIZZ_ASSERT(cond) ((void)0)
#endif

// np.dot([a0, a1], [b0, b1]) and np.linalg.norm([x, y])**2 as the HOST's numpy/BLAS evaluates them.  The
// contract is host-dependent (OpenBLAS picks its kernel per CPU), hence a build-time switch, -DFIZZ_DOT_MODE=:
//   1 (default; measured on the golden host: 100 % of 100 000 trials)  fma(a1, b1, rn(a0 b0))
//   2  the other fused order                                            fma(a0, b0, rn(a1 b1))
//   0  no fusion                                                        rn(rn(a0 b0) + rn(a1 b1))
// fizz2d_b200/lib.py probes numpy at load time and picks (building it on demand) the library of the matching mode; the
// oracle has the same switch (oracle/Makefile: FZO_DOT_MODE).  Everything else in the kernels is mode-independent.
#ifndef FIZZ_DOT_MODE
#define FIZZ_DOT_MODE 1
#endif
template <typename T>
__host__ __device__ __forceinline__ T dot2(T a0, T a1, T b0, T b1) {
#if FIZZ_DOT_MODE == 1
    return fma(a1, b1, a0 * b0);
#elif FIZZ_DOT_MODE == 2
    return fma(a0, b0, a1 * b1);
#else
    return a0 * b0 + a1 * b1;   // (-fmad=false: never contracted)
#endif
}
// dot([x, y], [x, y]): the argument of the sqrt in np.linalg.norm
template <typename T>
__host__ __device__ __forceinline__ T sq2(T x, T y) { return dot2(x, y, x, y); }

__device__ __forceinline__ void trace_hit(const KParams &kp, long long w, long long step, int call, int i, int j,
                                          double nx, double ny, double mag) {
    unsigned long long idx = atomicAdd(kp.trace_count, 1ULL);
    if ((long long)idx < kp.trace_cap) {
        FizzContact c;
        c.world = w; c.step = (int)step; c.call = call; c.i = i; c.j = j; c.nx = nx; c.ny = ny; c.mag = mag;
        kp.trace[idx] = c;
    }
}

// Point.move (physics.py:502-529) of one centre of mass over `h`, starting acceleration (a0x, a0y); the move
// installs the acceleration (accx, accy) = sum(forces)/mass (:516).
//   halfdt = .5*dt ; uv = acc*halfdt ; va = vel + uv ; pos += va*dt ; acc = new ; vel = va + acc*halfdt
#define FIZZ_MOVE_POINT(PX, PY, VX, VY, h, a0x, a0y, accx, accy, ox, oy, ovx, ovy)   \
    {                                                                                \
        const double halfdt_ = .5 * (h);                                             \
        const double vax_ = (VX) + (a0x) * halfdt_, vay_ = (VY) + (a0y) * halfdt_;   \
        ox = (PX) + vax_ * (h);                                                      \
        oy = (PY) + vay_ * (h);                                                      \
        ovx = vax_ + (accx) * halfdt_;                                               \
        ovy = vay_ + (accy) * halfdt_;                                               \
    }


// One coordinate of Point.move: the operation sequence of FIZZ_MOVE_POINT (used by the predictor of the spec rounds).
__host__ __device__ __forceinline__ void move_coord(double &p, double &v, double h, double a0, double acc) {
    const double halfdt = .5 * h;
    const double va = v + a0 * halfdt;
    p = p + va * h;
    v = va + acc * halfdt;
}
// One step of the predictor for one coordinate (class byte: see SpecParams).  Classes 0 and 3 leave it alone.
__device__ __forceinline__ void spec_predict(double &p, double &v, int cls, double dt, double rdt, double acc) {
    if (cls == 1) move_coord(p, v, dt, acc, acc);
    else if (cls == 2) move_coord(p, v, rdt, acc, acc);
}

// The trial that ends the halving search of a timestep in which EVERY trial fails: dt / 2 / 2 ... (:106) is dt * 2^-t
// exactly as long as nothing gets subnormal; the first t >= 1 with dt_t <= TIME_TOL (:99).  Warp-wide; kstar < 0: none.
__device__ __forceinline__ void halving_floor(const KParams &kp, int lane, int &kstar, double &dtk) {
    double dtt = kp.time_disc;
    if (dtt > 1e-280) dtt = dtt * __longlong_as_double((long long)(1023 - lane) << 52);
    else for (int t = 0; t < lane; t++) dtt = dtt / 2;
    const unsigned doneb = __ballot_sync(0xffffffffu, lane >= 1 && !(dtt > kp.time_tol));
    kstar = __ffs(doneb) - 1;
    dtk = __shfl_sync(0xffffffffu, dtt, kstar < 0 ? 0 : kstar);
}

}  // namespace fizz
