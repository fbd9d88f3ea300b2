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
                                   // [2] check_collisions() made by the contact kernel, [3] of those in the fast path
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

}  // namespace fizz
