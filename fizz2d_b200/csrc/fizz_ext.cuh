// fizz_ext.cuh -- default-off physics extensions of the resolution step.
//
// The reference has NO friction and ignores ELASTICITY between two free bodies (physics.py:201-223 is a perfectly
// elastic exchange); notes/collisions.tex:16-35 derives the restitution it meant to have.  These two
// functions are this project's DEFINITION of both (there is nothing to be bit-compatible with; oracle/fizz_oracle.c
// restates them operation for operation and the CUDA kernels are checked against it).  With mu == 0 and
// ff_restitution < 0 -- the defaults -- neither is called and every result is the reference's, bit for bit.
//
//   Coulomb friction, coefficient mu >= 0: after the normal response of a contact with unit normal u, the tangential
//   velocity (relative velocity for two free bodies) along t = (-u_y, u_x) is reduced by an impulse of at most mu times
//   the normal impulse; it sticks (tangential velocity -> 0) when less is enough.  The kinetic energy it removes is
//   booked as heat, like the reference books the normal loss (physics.py:191-195): K + P + H stays constant.
//   Restitution c_r in [0, 1] between two free bodies (notes/collisions.tex:24-35):
//     v1' = (m1 v1 + m2 v2 + c_r m2 (v2 - v1)) / (m1 + m2),  v2' likewise; loss (1 - c_r^2) m_red (v1 - v2)^2 / 2 -> heat.
// Arithmetic: plain IEEE operations in the order written (no contraction), np.dot-style fma only in dot2().
#pragma once

#include "fizz_common.cuh"

namespace fizz {

// Free body (mass m1, velocity v after the reference's normal update :188) against a FIXED body.
// ope = (1 + ELASTICITY), vdotN = v.u before the update (:183): the normal impulse per unit mass is ope * |vdotN|.
template <typename T>
__device__ __forceinline__ void ext_friction_fixed(T mu, T ope, T m1, T ux, T uy, T vdotN, T &vx, T &vy, T &heat) {
    const T tx = -uy, ty = ux;
    const T vt = dot2(vx, vy, tx, ty);
    const T lim = mu * (ope * fabs(vdotN));
    T dvt = -vt;                                   // stick
    if (fabs(vt) > lim) dvt = (vt > T(0)) ? -lim : lim;  // slide
    vx += dvt * tx;
    vy += dvt * ty;
    const T vt1 = vt + dvt;
    heat += (T(.5) * m1) * (vt * vt - vt1 * vt1);
}

// Normal velocities after a free-free contact with restitution c_r (replaces :209 and :218).
template <typename T>
__device__ __forceinline__ void ext_restitution_free(T cr, T m1, T mprop1, T mprop2, T vn1, T vn2, T &v1u, T &v2u, T &heat) {
    const T mean = mprop1 * vn1 + mprop2 * vn2;
    v1u = mean + (mprop2 * cr) * (vn2 - vn1);
    v2u = mean + (mprop1 * cr) * (vn1 - vn2);
    const T dvn = vn1 - vn2;
    heat += ((T(.5) * (m1 * mprop2)) * (T(1) - cr * cr)) * (dvn * dvn);
}

// Coulomb friction between two free bodies, applied to their velocities after the normal exchange.
// vn1, vn2 = normal velocities BEFORE the exchange; cr_eff = the restitution that exchange used (1 for the reference's).
template <typename T>
__device__ __forceinline__ void ext_friction_free(T mu, T cr_eff, T m1, T m2, T mprop2, T ux, T uy, T vn1, T vn2, T &v1x,
                                                  T &v1y, T &v2x, T &v2y, T &heat) {
    const T tx = -uy, ty = ux;
    const T vt1 = dot2(v1x, v1y, tx, ty), vt2 = dot2(v2x, v2y, tx, ty);
    const T rel = vt1 - vt2;
    const T mred = m1 * mprop2;                    // m1 m2 / (m1 + m2)
    const T jmax = mu * (((T(1) + cr_eff) * fabs(vn1 - vn2)) * mred);
    T jt = mred * fabs(rel);                       // stick
    if (jt > jmax) jt = jmax;                           // slide
    const T sj = (rel > T(0)) ? -jt : jt;           // tangential impulse on body 1
    const T dv1 = sj / m1, dv2 = -(sj / m2);
    v1x += dv1 * tx; v1y += dv1 * ty;
    v2x += dv2 * tx; v2y += dv2 * ty;
    const T w1 = vt1 + dv1, w2 = vt2 + dv2;
    heat += (T(.5) * m1) * (vt1 * vt1 - w1 * w1) + (T(.5) * m2) * (vt2 * vt2 - w2 * w2);
}

}  // namespace fizz
