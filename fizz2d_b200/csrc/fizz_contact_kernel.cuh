// fizz_contact_kernel.cuh -- warp-per-world kernel with the FULL update() semantics ("hybrid" mode, stage 2).
//
// One WARP per world, persistent: warps pull world indices from a device-side work list (worlds the ballistic
// kernel stopped at a non-empty broad phase, plus step 0 of every world).  The world's state (com, snapshot,
// polar offsets, vertices, contact list) lives in the warp's slice of shared memory for the whole launch; the
// fixed scene is staged once per CTA.  A pass over a batch is bounded by the SEQUENTIAL depth of its worst world
// (30 halving trials + fallback + up to ~300 resolve passes per timestep for a trapped body), so
// everything here is organised around the latency of one check_collisions() (physics.py:262-289):
//   * broad phase: lanes stride over a packed pair table, __ballot compacts the survivors in reference order;
//   * narrow phase: lanes map to (pair, axis) items of polypoly_collision (:765-840) in fixed-width slots, two
//     items per lane for ILP, butterfly __shfl reductions inside a slot ("any separating axis", first strict
//     minimum overlap in axis order), __ballot-compacted contact list;
//   * the unit normal of a free polygon's edge, sqrt + 2 divisions (:790), is memoised on the exact edge vector:
//     it is a pure function of (ax, ay) and >98% of evaluations repeat the previous bit pattern;
//   * step halving (:99-129) is searched speculatively: once trial 0 fails, lane t evaluates trial t on its own
//     (every trial restarts from the start-of-step state, so they are independent) and a __ballot picks the
//     first trial that ends the loop;
//   * resolution (:142-233) walks the list in order; vertices are pushed lane-parallel, the centre of mass by
//     lane 0;
//   * a resolve loop that keeps returning ONE tuple (free body, fixed body) -- the trapped-body pattern that bounds a
//     pass -- continues in fast_resolve<N>: state in registers, one branch per pass (see there).
#pragma once

#include <type_traits>

#include "fizz_common.cuh"
#include "fizz_ext.cuh"
#include "fizz_spec.cuh"

namespace fizz {

constexpr int CONTACT_WARPS = 4;  // warps per CTA
constexpr int CONTACT_BLOCK = CONTACT_WARPS * 32;
#ifdef FIZZ_RUNNER_DEBUG
__device__ unsigned long long g_dbg_cycles[65536];   // per world: cycles in ordinary (non-spec) items from timestep 2688 on
#endif
constexpr int SP_CAP = 160;       // survivor list capacity (>= MAX_PAIRS)
#define FIZZ_SPEC_TIMEOUT_LOG2 32
constexpr unsigned long long BULK_LIST = 16384;  // work lists this long are a throughput phase: the compact instance takes them
constexpr long long FIZZ_ST_SPEC_ABORT = 100;    // internal: a spec item gave up (never leaves its slot record)

// packed pair descriptor: ii[0,4) jj[4,9) n1[9,14) n2[14,19) v1[19,27) v2[27,35)
typedef unsigned long long pairmeta_t;
__host__ __device__ __forceinline__ pairmeta_t pack_pair(int ii, int jj, int n1, int n2, int v1, int v2) {
    return (pairmeta_t)ii | ((pairmeta_t)jj << 4) | ((pairmeta_t)n1 << 9) | ((pairmeta_t)n2 << 14) |
           ((pairmeta_t)v1 << 19) | ((pairmeta_t)v2 << 27);
}
#define PM_II(m) ((int)((m) & 15))
#define PM_JJ(m) ((int)(((m) >> 4) & 31))
#define PM_N1(m) ((int)(((m) >> 9) & 31))
#define PM_N2(m) ((int)(((m) >> 14) & 31))
#define PM_V1(m) ((int)(((m) >> 19) & 255))
#define PM_V2(m) ((int)(((m) >> 27) & 255))

struct ContactSmem {
    // CTA-shared
    double *sfix, *fcx, *fcy, *fthr;
    pairmeta_t *pair_meta;
    unsigned char *vbody;
    // per-warp
    double *px, *py, *vx, *vy, *sx, *sy, *thr, *mass, *npx, *npy, *off, *vert, *hnx, *hny, *hmag;
    double *max_, *may_, *mnx, *mny;  // edge-normal memo per free vertex (edge v -> v+1)
    pairmeta_t *sp;                   // surviving pairs of the last broad phase
    unsigned char *hi, *hj;
};

__host__ __device__ inline size_t contact_smem_bytes(int V, int FV) {
    size_t cta_d = (size_t)FV * 6 + 3 * FIZZ_MAX_FIXED + SP_CAP;
    size_t warp_d = 80 + 8 * (size_t)V + 96 + SP_CAP;
    size_t bytes = 8 * (cta_d + CONTACT_WARPS * warp_d);
    bytes += ((size_t)V + 15) / 16 * 16;  // vbody
    bytes += CONTACT_WARPS * 64;          // hi, hj
    return (bytes + 15) / 16 * 16;
}

__device__ __forceinline__ ContactSmem contact_carve(double *base, int V, int FV, int warp) {
    ContactSmem s;
    double *d = base;
    s.sfix = d; d += FV * 6;
    s.fcx = d; d += FIZZ_MAX_FIXED;
    s.fcy = d; d += FIZZ_MAX_FIXED;
    s.fthr = d; d += FIZZ_MAX_FIXED;
    s.pair_meta = reinterpret_cast<pairmeta_t *>(d); d += SP_CAP;
    const int warp_d = 80 + 8 * V + 96 + SP_CAP;
    double *wd = d + (size_t)warp * warp_d;
    d += (size_t)CONTACT_WARPS * warp_d;
    s.px = wd; s.py = wd + 8; s.vx = wd + 16; s.vy = wd + 24; s.sx = wd + 32; s.sy = wd + 40;
    s.thr = wd + 48; s.mass = wd + 56; s.npx = wd + 64; s.npy = wd + 72;
    s.off = wd + 80; s.vert = s.off + 2 * V;
    s.max_ = s.vert + 2 * V; s.may_ = s.max_ + V; s.mnx = s.may_ + V; s.mny = s.mnx + V;
    s.hnx = s.mny + V; s.hny = s.hnx + 32; s.hmag = s.hny + 32;
    s.sp = reinterpret_cast<pairmeta_t *>(s.hmag + 32);
    unsigned char *b = reinterpret_cast<unsigned char *>(d);
    s.vbody = b; b += (V + 15) / 16 * 16;
    unsigned char *wb = b + (size_t)warp * 64;
    s.hi = wb; s.hj = wb + 32;
    return s;
}

// (ax, ay) / norm((ax, ay)) (:790): sqrt + two divisions, ~150 instructions with their slow paths -- ONE copy, out of line.
__device__ __noinline__ double2 unit_normal(double ax, double ay) {
    const double nrm = sqrt(sq2(ax, ay));
    return make_double2(ax / nrm, ay / nrm);
}

// Unit normal of the edge e -> e+1 of a free polygon whose vertices are (vx[e], vy[e]) (:775,:778,:790).
// MEMO: read-write memo on the exact edge vector; pass false from code where several lanes may evaluate the same
// edge with different vertex sets (speculative trials): then the memo is only read.
template <bool MEMO_WRITE>
__device__ __forceinline__ void edge_normal(const ContactSmem &s, int m, double x0, double y0, double x1, double y1,
                                            double &nx, double &ny) {
    const double ax = y1 - y0;   // (p2.y - p1.y, p1.x - p2.x)
    const double ay = x0 - x1;
    // bitwise key: -0.0 and +0.0 give normals that differ in the sign of zero
    if (__double_as_longlong(ax) == __double_as_longlong(s.max_[m]) &&
        __double_as_longlong(ay) == __double_as_longlong(s.may_[m])) {
        nx = s.mnx[m]; ny = s.mny[m];
    } else {
        const double2 un = unit_normal(ax, ay);         // :790
        nx = un.x; ny = un.y;
        if (MEMO_WRITE) { s.max_[m] = ax; s.may_[m] = ay; s.mnx[m] = nx; s.mny[m] = ny; }
    }
}

#define FIZZ_SAT_NAME(n) n
#define FIZZ_SAT_UNROLL
#include "fizz_sat.inc"
#undef FIZZ_SAT_NAME
#undef FIZZ_SAT_UNROLL
#define FIZZ_SAT_NAME(n) n##_compact
#define FIZZ_SAT_UNROLL _Pragma("unroll 1")
#include "fizz_sat.inc"
#undef FIZZ_SAT_NAME
#undef FIZZ_SAT_UNROLL

// ---------------------------------------------------------------------------------------------------------------
// Single-contact fast path: the resolve loop AND, while the pattern holds, whole timesteps, in registers.
//
// The worlds that bound a pass are a free polygon b trapped against FIXED polygons: every check_collisions() of
// the resolve loop (:140-234) returns exactly one tuple (b, fixed j) for hundreds of passes per timestep, and the next
// timestep starts the same way: trial 0 collides deeper than COLLISION_TOL (:99), every halving trial does too (proved,
// not evaluated: see the contact kernel), the loop ends on TIME_TOL, the fallback (:133-137) reverts the centre of mass
// and the resolve loop starts again.  Between two passes only b changes, so (1) every pair without b keeps its result
// ("no contact"), (2) only the pairs (b, *) need the broad phase again, and as long as the same fixed partners (and no
// free body) pass it, (3) the SAT work is a fixed set of (partner, axis) items.  trapped_run keeps b's vertices, the
// lane's axis constants, its edge-normal memo and the partner's projection bounds in REGISTERS and runs
// check_collisions() + resolution back to back; when the check that ends the resolve loop finds nothing it finishes
// update() (:237-259) and starts the next one -- broad phase of all pairs on the snapshots, trial 0, the halving proof,
// the last trial, the fallback -- without going back to the general path.  It returns
//   TRAP_MIDSTEP   with the state written back exactly as it is right before the check_collisions() call at :234, as
//                  soon as anything else happens inside a resolve loop (other candidate set, zero or several tuples,
//                  watchdog): the general path continues that loop;
//   TRAP_EMPTY     like TRAP_MIDSTEP, but the :234 check has been made (and counted) and found nothing: the general path
//                  finishes update() (only for timesteps that did not take the fallback);
//   TRAP_BOUNDARY  at a timestep boundary (target reached, or the next timestep does not start with the pattern: its
//                  prologue is evaluated without side effects and discarded, the general path runs that timestep).
// Arithmetic is the same, operation for operation, as the general path.
// ---------------------------------------------------------------------------------------------------------------
enum { TRAP_MIDSTEP = 0, TRAP_EMPTY = 1, TRAP_BOUNDARY = 2 };

// Lanes map to (candidate partner, SAT axis) items in slots of 8 lanes -- or 16 in scenes whose largest n1 + n2 exceeds 8
// (kp.swl; e.g. a triangle against PENTAGONVOID's hexagon, the GA population of BASELINE config 5).
// b has 3 or 4 vertices, kept in four register pairs: a triangle's first vertex (in the lane's rotation) sits there twice,
// which changes neither a minimum nor a maximum of projections -- one instruction stream for both shapes (the kernel is
// bound by instruction fetch wherever it is not bound by this chain).  MULTI = false: the variant of the throughput
// instance, which leaves update()'s end and the next timestep to the general path (smaller footprint).
// TRACK = true (with MULTI = false): additionally reports the bounding box of b's centre of mass over the loop -- for the
// caller that runs the resolve loops of TWO trapped bodies one after the other (see the contact kernel).
struct TrapBox { double xlo, xhi, ylo, yhi; };
template <bool MULTI, bool TRACK = false>
__device__ __forceinline__ int trapped_run(const ContactSmem &s, const KParams &kp, int F, int X, int lane, int b,
                                           double &heat, int &ncalls, long long &ncalls_total, long long &step, int &k,
                                           double &dt, int &steps_here, int &last_calls, const int kstar, const double dtk,
                                           const long long target, const long long stop_at, const long long abort_at,
                                           TrapBox *box = nullptr) {
    // (spec rounds: stop_at / abort_at are values of ncalls_total -- every call made here costs one unit -- at which the
    //  item ends at the next timestep boundary / gives the resolve loop back to the general path, which aborts the item)
    constexpr int NM = 4;
    const unsigned FULL = 0xffffffffu;
    const int n1 = kp.nv[b];
    const int v1 = kp.vs[b];
    FIZZ_ASSERT(b >= 0 && b < F && (n1 == 3 || n1 == 4));
    double cpx = s.px[b], cpy = s.py[b], cvx = s.vx[b], cvy = s.vy[b];
    [[maybe_unused]] double bxlo = cpx, bxhi = cpx, bylo = cpy, byhi = cpy;
    const double m1 = s.mass[b], thr_b = s.thr[b];
    // this lane's broad-phase partner: body `lane` of objs+fixed_objs.  (a-b)^2 == (b-a)^2 exactly, so the order of
    // the subtraction at :277 does not matter.
    const bool partner = (lane < F + X) && (lane != b);
    double ox = 0.0, oy = 0.0, ot = 0.0;
    if (partner) {
        if (lane < F) { ox = s.sx[lane]; oy = s.sy[lane]; ot = s.thr[lane]; }
        else { ox = s.fcx[lane - F]; oy = s.fcy[lane - F]; ot = s.fthr[lane - F]; }
    }
    bool surv = false;
    {
        const double dx = cpx - ox, dy = cpy - oy;
        const double d2 = sq2(dx, dy);
        surv = partner && !(d2 > thr_b && d2 > ot);
    }
    const unsigned mask0 = __ballot_sync(FULL, surv);
    const int ncand = __popc(mask0);
    // slots of 8 lanes (up to 4 candidate partners) or, in scenes with a larger n1 + n2, of 16 lanes (up to 2)
    const int swl = kp.swl, sw = 1 << swl;
    if (ncand == 0 || ncand > (32 >> swl) || (mask0 & ((1u << F) - 1u)) != 0) return TRAP_MIDSTEP;  // only fixed partners, one 32-lane set

    const int slot = lane >> swl, kk = lane & (sw - 1);
    int jj = -1;
    {
        unsigned mm = mask0;
        for (int q = 0; q < slot; q++) mm &= mm - 1;
        if (slot < ncand) jj = __ffs(mm) - 1;
    }
    int n2 = 0, v2 = 0;
    if (jj >= 0) { n2 = kp.nv[jj]; v2 = kp.vs[jj]; }
    const bool active = (jj >= 0) && (kk < n1 + n2);
    const bool freeaxis = active && (kk < n1);
    const bool active_slot = (jj >= 0);
    const unsigned slot_bits = ((1u << sw) - 1u) << (sw * slot);
    // b's vertices, all of them in every lane, ROTATED so that a lane that owns edge k of b finds the edge's end points
    // in registers 0 and 1 (min/max over the same set of projections: only the sign of a zero could depend on order).
    const int rot = freeaxis ? kk : 0;
    double vtx[NM], vty[NM];
#pragma unroll
    for (int v = 0; v < NM; v++) {
        int src = rot + v;
        if (src >= n1) src -= n1;      // (v = 3 of a triangle: the vertex of v = 0 again)
        vtx[v] = s.vert[2 * (v1 + src)];
        vty[v] = s.vert[2 * (v1 + src) + 1];
    }
    // axis registers of this lane: unit normal, the partner's projection bounds on it and -- for an edge of b -- the
    // exact edge vector the normal was computed from (memo key; NaN = not computed yet).
    // Lanes without an item carry bounds that decide their result without a flag: (+inf, +inf) in a slot without a
    // candidate ("separated": hi1 < +inf), (-inf, +inf) in the spare lanes of a candidate slot ("overlap +inf", which never
    // wins a minimum against a finite overlap).
    double anx = 0.0, any_ = 0.0, alo2 = active_slot ? -INFINITY : INFINITY, ahi2 = INFINITY;
    double keyx = NAN, keyy = NAN;  // (keys only used by edge lanes)
    if (active && !freeaxis) {
        const double *g = s.sfix + (v2 + (kk - n1)) * 6;
        anx = g[2]; any_ = g[3]; alo2 = g[4]; ahi2 = g[5];
    }
    int passes = 0;
    auto capped_budget = [&]() -> int {
        const long long wd = kp.max_calls - ncalls, left = abort_at - ncalls_total;
        return (int)(left < wd ? (left < 0 ? 0 : left) : wd);
    };
    int budget = capped_budget();  // watchdog: the general path raises FIZZ_ST_CAPPED
    // (a "travel budget" that skips the broad phase while b cannot have reached a partner's threshold was measured and
    //  dropped: trapped bodies jump 20-100 px per pass, the budget is exhausted almost every pass)
    //
    // Shape of the loop.  One warp runs this chain alone, so dependent-issue LATENCY sets its speed, and control flow is
    // the most expensive instruction class there is (tools/micro/latency.cu on B200: a taken branch ~22 cycles, a
    // BSSY/BRA/BSYNC guarded block ~47, against 8 for a DFMA and 18 for a redux).  The hot loop therefore has exactly
    // one branch, its back edge: check() evaluates a whole check_collisions() without side effects and says whether
    // its result is the plain case (same candidates, one colliding partner, memoised normals valid, finite overlap,
    // watchdog not due); the loop body applies that result and checks again.  Everything else -- a stale edge-normal
    // memo (<2% of the passes) or the hand-over to the general path -- happens outside the hot loop.
    // d2 > thr_b && d2 > ot  <=>  d2 > max(thr_b, ot)  (:277, finite radii); lanes without a partner never change
    const double tmax = partner ? ((thr_b > ot) ? thr_b : ot) : INFINITY;
    const bool surv0 = partner ? surv : true;
    // loop constants in registers (an opaque zero keeps them from being re-read from the constant bank every pass)
    const long long opaque0 = kp.W >> 62;  // W < 2^62
    const double eps_r = __longlong_as_double(__double_as_longlong(kp.eps) ^ opaque0);
    const double vthr_r = __longlong_as_double(__double_as_longlong(kp.vib_thr) ^ opaque0);
    const double ope_r = __longlong_as_double(__double_as_longlong(kp.ope) ^ opaque0);
    const double hm_r = __longlong_as_double(__double_as_longlong(.5 * m1 * kp.ome2) ^ opaque0);
    const double gate_min_r = __longlong_as_double(__double_as_longlong(2.0 * fabs(kp.vib_tol) + 1e-300) ^ opaque0);
    double wmag = 0.0;                                    // check(): |overlap| on the winning axis
    bool okv = false;
    unsigned liveb = 0u;  // lanes of the slots whose pair collides, as of the last check()

    // (the latency of one warp is also why the statements below are ordered the way they are: the chain
    //  push -> vertices -> projections -> ballot -> three redux -> shuffle comes first, and the rest of a pass --
    //  velocity, heat, centre of mass, broad phase, memo check -- is issued while the warp-wide reductions are in flight)
    // check() leaves the negated winning axis and the push it asks for here; the next check() completes that resolution
    double pux = 0.0, puy = 0.0, pnvx = 0.0, pnvy = 0.0;

    // b's projections on this lane's axis (:795-804)
    auto project = [&](double &lo1, double &hi1) {
        const double c0 = dot2(anx, any_, vtx[0], vty[0]), c1 = dot2(anx, any_, vtx[1], vty[1]);   // as a two-level tree
        const double c2 = dot2(anx, any_, vtx[2], vty[2]), c3 = dot2(anx, any_, vtx[3], vty[3]);
        const double la = (c1 < c0) ? c1 : c0, ha = (c1 > c0) ? c1 : c0;
        const double lb = (c3 < c2) ? c3 : c2, hb = (c3 > c2) ? c3 : c2;
        lo1 = (lb < la) ? lb : la;
        hi1 = (hb > ha) ? hb : ha;
    };

    double plo1 = 0.0, phi1 = 0.0;   // b's projection bounds on this lane's axis as the last check(first) saw them
    auto check = [&](auto first) -> bool {
        // ---- narrow phase: one (partner, axis) item per lane --------------------------------------------------------
        double lo1, hi1;
        project(lo1, hi1);
        if constexpr (decltype(first)::value) { plo1 = lo1; phi1 = hi1; }
        double lo2 = alo2, hi2 = ahi2;
        if (lo1 > lo2) {                                                         // :815-817
            double t = lo1; lo1 = lo2; lo2 = t;
            t = hi1; hi1 = hi2; hi2 = t;
        }
        // (lanes without an item: see the sentinel bounds above)  a zero byte of sepbits <=> a colliding partner
        const bool sep = hi1 < lo2;                                              // :820
        const double cur = hi1 - lo2;                                            // :824
        const unsigned sepbits = __ballot_sync(FULL, sep);
        // first strict minimum in axis order (:827-830) over the 8 lanes of the live slot: three 32-bit redux.sync.min
        // (high word, low word, lane).  Overlaps of a non-separated pair are >= +0.0 -- or the -0.0 of (-0.0) - (+0.0),
        // hence the sign mask: `cur < overlap` is false between the two zeros and the push uses mag = fabs(overlap) --
        // so their bit patterns order like unsigned integers; +inf and NaN order above every finite overlap and never
        // win against one (`cur < overlap` with overlap = inf is false for them): if the minimum is one of them the
        // general path takes over.
        const bool mine = (sepbits & slot_bits) == 0u;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(cur);
        const unsigned hiw = (unsigned)(bits >> 32) & 0x7fffffffu, low = (unsigned)bits;
        const unsigned mhi = __reduce_min_sync(FULL, mine ? hiw : 0xffffffffu);
        if constexpr (!decltype(first)::value) {
            // ---- rest of the previous resolution, free body vs fixed body (:180-198) ------------------------------
            // :180-181 norm(nv) > VIBRATE_TOL.  The latency-tuned variant only applies pushes that are certainly longer
            // (see gate_min below): no test here
            if (MULTI || sq2(pnvx, pnvy) > vthr_r) { cpx += pnvx; cpy += pnvy; }
            if constexpr (TRACK) {
                bxlo = (cpx < bxlo) ? cpx : bxlo; bxhi = (cpx > bxhi) ? cpx : bxhi;
                bylo = (cpy < bylo) ? cpy : bylo; byhi = (cpy > byhi) ? cpy : byhi;
            }
            const double vdotN = dot2(cvx, cvy, pux, puy);                           // :183
            const double cc = ope_r * vdotN;                                         // :188
            cvx -= cc * pux;
            cvy -= cc * puy;
            heat += hm_r * (vdotN * vdotN);                                          // :191,:195
        }
        const bool m1_ = mine && hiw == mhi;
        const unsigned mlo = __reduce_min_sync(FULL, m1_ ? low : 0xffffffffu);
        // exactly one candidate slot without a separating axis: its lanes are the `mine` lanes
        liveb = __ballot_sync(FULL, mine);
        const bool one_live = __popc(liveb) == sw;
        // ---- broad phase of the pairs (b, *) on b's new snapshot (= com: finish_update follows every push) ----------
        const double dx = cpx - ox, dy = cpy - oy;
        const double d2 = sq2(dx, dy);
        const bool surv_l = !(d2 > tmax);
        // ---- memo check: edge k of b is (p2.y - p1.y, p1.x - p2.x) (:788-789) ----------------------------------------
        // (surv_l / miss_l stay local: kept alive across the loop they would be computed twice per pass; the code after
        //  the loop derives them again from the same registers)
        const double ax = vty[1] - vty[0], ay = vtx[0] - vtx[1];
        const bool miss_l = freeaxis & ((((unsigned long long)__double_as_longlong(ax) ^ (unsigned long long)__double_as_longlong(keyx)) |
                                         ((unsigned long long)__double_as_longlong(ay) ^ (unsigned long long)__double_as_longlong(keyy))) != 0ull);
        wmag = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));   // fabs(overlap) (:125-129)
        const double sc = wmag + eps_r;                                          // :158
        // (MULTI: nv = -n (mag + EPSILON) with a unit normal n, so norm(nv) = (mag + EPSILON)(1 + O(1e-15)): twice the
        //  tolerance settles the VIBRATE_TOL test of :180 without evaluating it; shorter pushes go to the general path)
        const bool long_push = !MULTI || sc > gate_min_r;
        okv = __all_sync(FULL, !((surv_l != surv0) | miss_l) && one_live && mhi < 0x7ff00000u && passes < budget && long_push);
        const bool m2_ = m1_ && low == mlo;
        const unsigned wl = __reduce_min_sync(FULL, m2_ ? (unsigned)lane : 32u);  // lowest axis index among equal minima
        // the winning axis (:828-830), negated (:838-839), and the push it asks for (:158): straight into the registers the
        // NEXT check's deferred part reads (their old values were consumed above; copies at the back edge cost issue slots)
        pux = __shfl_sync(FULL, anx, wl) * -1.0;
        puy = __shfl_sync(FULL, any_, wl) * -1.0;
        pnvx = pux * sc;
        pnvy = puy * sc;
        return okv;
    };

    // refresh the memo of the lanes whose edge vector changed: unit normal (:790), partner bounds (:806-811)
    auto refresh = [&]() {
        const double ax = vty[1] - vty[0], ay = vtx[0] - vtx[1];
        if (freeaxis && ((__double_as_longlong(ax) != __double_as_longlong(keyx)) ||
                         (__double_as_longlong(ay) != __double_as_longlong(keyy)))) {
            const double2 un = unit_normal(ax, ay);
            anx = un.x; any_ = un.y; keyx = ax; keyy = ay;
            const double *g = s.sfix + v2 * 6;
            alo2 = ahi2 = dot2(anx, any_, g[0], g[1]);
            for (int v = 1; v < n2; v++) {
                const double c = dot2(anx, any_, g[v * 6], g[v * 6 + 1]);
                if (c < alo2) alo2 = c;
                if (c > ahi2) ahi2 = c;
            }
        }
        __syncwarp();
    };
    auto memo_stale = [&]() -> bool {
        const double ax = vty[1] - vty[0], ay = vtx[0] - vtx[1];
        const bool miss_now = freeaxis && ((__double_as_longlong(ax) != __double_as_longlong(keyx)) ||
                                           (__double_as_longlong(ay) != __double_as_longlong(keyy)));
        return __any_sync(FULL, miss_now);
    };
    // b's vertices as pre_update rebuilds them from a moved centre of mass (:358-359), in this lane's rotation
    auto rebuild = [&](double cx, double cy) {
#pragma unroll
        for (int v = 0; v < NM; v++) {
            int src = rot + v;
            if (src >= n1) src -= n1;
            vtx[v] = s.off[2 * (v1 + src)] + cx;
            vty[v] = s.off[2 * (v1 + src) + 1] + cy;
        }
    };
    const bool step_pattern_ok = kstar >= 1 && dtk < kp.time_tol && kp.time_disc > kp.time_tol;

    // One copy of each check() in the instruction stream (the kernel is bound by instruction fetch wherever it is not
    // bound by this chain): phase says what the check at the top of the outer loop is --
    //   PH_RESOLVE  the :234 check of a resolve loop (entry from the general path, or again after a stale memo);
    //   PH_TRIAL0   the first halving trial of a new timestep (:124), vertices rebuilt from the moved centre of mass;
    //   PH_TRIALK   the last halving trial (and, same list, the fallback's check :137).
    enum { PH_RESOLVE = 0, PH_TRIAL0 = 1, PH_TRIALK = 2 };
    int phase = PH_RESOLVE;
    bool in_prologue_step = false;   // the current timestep was started here: other bodies' trial vertices are not in smem
    int ret = TRAP_BOUNDARY;
    bool need_rebuild = false;       // the check at the top runs on vertices rebuilt from the centre of mass (rbx, rby)
    double rbx = 0.0, rby = 0.0;
    for (;;) {
        if (need_rebuild) { rebuild(rbx, rby); need_rebuild = false; }
        refresh();
        bool ok = check(std::true_type{});
        if (MULTI && phase != PH_RESOLVE) {
            // ======================= prologue of a new update() (:99-137), speculative =============================
            // Nothing is stored anywhere but in this function's registers until the pattern is confirmed; if it is not,
            // the general path runs this timestep from its start (ret = TRAP_BOUNDARY).
            const double a0x = kp.accx, a0y = kp.accy;                      // step > 0 here
            if (!ok) break;                                                 // not exactly one tuple (b, fixed)
            if (phase == PH_TRIAL0) {
                if (!(wmag > kp.coll_tol)) break;                           // :99 the halving loop would end here
                // every halving trial fails (the contact kernel's proof, on this lane's axis of the colliding pair)
                double D = (fabs(cvx) + fabs(cvy)) * kp.time_disc + .5 * (fabs(a0x) + fabs(a0y)) * kp.time_disc * kp.time_disc;
                D = D * 1.000001 + 1e-9;
                bool clear = (wmag - D) > (kp.coll_tol + 1e-7);
                if (clear && active && (liveb & slot_bits) != 0u) {
                    const double fa = ahi2 - plo1, fb = phi1 - alo2;
                    const double fl = (fa < fb) ? fa : fb;
                    clear = (fl - D) > (kp.coll_tol + 1e-7 + 1e-9 * (fabs(fl) + D));  // (false for NaN)
                }
                if (!__all_sync(FULL, clear)) break;
                // trials 1 .. kstar-1 were made (and failed); trial kstar (:106-129)
                double tvx_, tvy_;
                FIZZ_MOVE_POINT(cpx, cpy, cvx, cvy, dtk, a0x, a0y, kp.accx, kp.accy, rbx, rby, tvx_, tvy_);
                need_rebuild = true;
                phase = PH_TRIALK;
                continue;
            }
            // the fallback (:133-137): the centre of mass stays where the timestep started, the vertices keep the last
            // trial's positions (Q5); its check returns the list of the last trial, whose resolve pass comes first and is
            // not a call of its own (passes = -1)
            k = kstar; dt = dtk;
            ncalls = kstar + 2;                                             // trials 0 .. kstar and the fallback's check
            ncalls_total += kstar + 2;
            budget = capped_budget();
            passes = -1;
            in_prologue_step = true;
            phase = PH_RESOLVE;
        }
        // =============================== the resolve loop of the current timestep ================================
        while (ok) {
            // ---- resolution, free body vs fixed body (:171-174); check() completes it (:180-198) --------------
            passes++;
#pragma unroll
            for (int v = 0; v < NM; v++) { vtx[v] += pnvx; vty[v] += pnvy; }         // :174
            ok = check(std::false_type{});
        }
        // (with a stale memo the check above decided nothing: refresh, check again)
        if (memo_stale()) continue;
        {   // the broad phase the last check() saw (same registers, same arithmetic)
            const double dx = cpx - ox, dy = cpy - oy;
            const double d2 = sq2(dx, dy);
            surv = !(d2 > tmax);
        }
        // The check that ended the loop, when it saw the same candidates and a separating axis for every one of them, IS
        // the check_collisions() of :234 that returns the empty list (pairs without b have not moved since the general
        // path found them apart): count it and let update() finish, instead of having the general path find the same
        // nothing again.
        if (passes < 0) passes = 0;   // (the uncounted first pass was not even made: the watchdog or the memo came first)
        const bool done_empty = !__any_sync(FULL, surv != surv0) && liveb == 0u && passes < budget;
        if (done_empty) passes++;
        ncalls += passes;
        ncalls_total += passes;
        passes = 0;
        const bool fell_back = k > 0 && dt < kp.time_tol;
        if (!MULTI || !done_empty || !fell_back || !step_pattern_ok) {
            ret = done_empty ? TRAP_EMPTY : TRAP_MIDSTEP;
            break;
        }

        // =============================== end of update() (:237-259), in registers ==================================
        {
            // :248-251 the rest of the timestep, no collision handling; the fallback reverted com.acc to the start-of-step
            // acceleration (Q5), the move installs the world's
            const bool acc_on = step > 0;
            const double b0x = acc_on ? kp.accx : 0.0, b0y = acc_on ? kp.accy : 0.0;
            const double rdt = kp.time_disc - dt;
            double nx_, ny_, nvx_, nvy_;
            FIZZ_MOVE_POINT(cpx, cpy, cvx, cvy, rdt, b0x, b0y, kp.accx, kp.accy, nx_, ny_, nvx_, nvy_);
            cpx = nx_; cpy = ny_; cvx = nvx_; cvy = nvy_;
            if (lane < F && lane != b) {
                double ox_, oy_, ovx_, ovy_;
                FIZZ_MOVE_POINT(s.px[lane], s.py[lane], s.vx[lane], s.vy[lane], rdt, b0x, b0y, kp.accx, kp.accy, ox_, oy_,
                                ovx_, ovy_);
                s.px[lane] = ox_; s.py[lane] = oy_; s.vx[lane] = ovx_; s.vy[lane] = ovy_;
                s.sx[lane] = ox_; s.sy[lane] = oy_;                     // finish_update (:251)
                ox = ox_; oy = oy_;                                     // this lane is b's broad-phase partner `lane`
            }
            if (lane == 0) {
                s.px[b] = cpx; s.py[b] = cpy; s.vx[b] = cvx; s.vy[b] = cvy; s.sx[b] = cpx; s.sy[b] = cpy;
            }
            __syncwarp();
            step++;                                                     // :259
            steps_here++;
            last_calls = ncalls;
            k = 0; dt = kp.time_disc; ncalls = 0;                       // the state of a fresh update()
            in_prologue_step = false;
            ret = TRAP_BOUNDARY;
        }
        if (step >= target || ncalls_total >= stop_at) break;
        {
            // broad phase of ALL pairs on the snapshots (:272-281): exactly the fixed partners of b, nothing else
            unsigned pm = 0u;
            bool other = false;
            const int P = F * (F - 1) / 2 + F * X;
            for (int p0 = 0; p0 < P; p0 += 32) {
                const int p = p0 + lane;
                if (p < P) {
                    const pairmeta_t m = s.pair_meta[p];
                    const int i = PM_II(m), j = PM_JJ(m);
                    double qx, qy, qt;
                    if (j < F) { qx = s.sx[j]; qy = s.sy[j]; qt = s.thr[j]; }
                    else { qx = s.fcx[j - F]; qy = s.fcy[j - F]; qt = s.fthr[j - F]; }
                    const double dx = s.sx[i] - qx, dy = s.sy[i] - qy;
                    const double d2 = sq2(dx, dy);
                    if (!(d2 > s.thr[i] && d2 > qt)) {
                        if (i == b && j >= F) pm |= 1u << j; else other = true;
                    }
                }
            }
            pm = __reduce_or_sync(FULL, pm);
            if (__any_sync(FULL, other) || pm != mask0) break;
        }
        {
            double tvx_, tvy_;                                          // trial 0 (:116-118)
            FIZZ_MOVE_POINT(cpx, cpy, cvx, cvy, kp.time_disc, kp.accx, kp.accy, kp.accx, kp.accy, rbx, rby, tvx_, tvy_);
            need_rebuild = true;
        }
        budget = 0x7fffffff;                                            // (the watchdog only counts in the resolve loop)
        phase = PH_TRIAL0;
    }
    // ---- write the state back for the general path ----------------------------------------------------------------
    __syncwarp();
    if (freeaxis && slot == 0 && keyx == keyx) {
        s.max_[v1 + kk] = keyx; s.may_[v1 + kk] = keyy; s.mnx[v1 + kk] = anx; s.mny[v1 + kk] = any_;
    }
    if (ret == TRAP_BOUNDARY) {
        // a timestep boundary: b's state is in shared memory already, the vertices are derived state again
        __syncwarp();
        k = 0; dt = kp.time_disc; ncalls = 0;
        return TRAP_BOUNDARY;
    }
    // inside a resolve loop, right before (TRAP_MIDSTEP) or after (TRAP_EMPTY) the check_collisions() call at :234
    if (MULTI && in_prologue_step) {
        // the timestep was started here: every OTHER body's vertices are where its last halving trial left them (Q5),
        // |r|cos(phi) + com moved by dt_kstar (:358-359) -- the general path reads and stores them
        const bool acc_on = step > 0;
        const double a0x = acc_on ? kp.accx : 0.0, a0y = acc_on ? kp.accy : 0.0;
        for (int v = lane; v < kp.V; v += 32) {
            const int bb = s.vbody[v];
            if (bb != b) {
                double qx, qy, qvx, qvy;
                FIZZ_MOVE_POINT(s.px[bb], s.py[bb], s.vx[bb], s.vy[bb], dt, a0x, a0y, kp.accx, kp.accy, qx, qy, qvx, qvy);
                s.vert[2 * v] = s.off[2 * v] + qx;
                s.vert[2 * v + 1] = s.off[2 * v + 1] + qy;
            }
        }
    }
    if (lane < n1) { s.vert[2 * (v1 + lane)] = vtx[0]; s.vert[2 * (v1 + lane) + 1] = vty[0]; }  // slot 0, k = lane
    if (lane == 0) {
        s.px[b] = cpx; s.py[b] = cpy; s.vx[b] = cvx; s.vy[b] = cvy; s.sx[b] = cpx; s.sy[b] = cpy;
    }
    if constexpr (TRACK) { box->xlo = bxlo; box->xhi = bxhi; box->ylo = bylo; box->yhi = byhi; }
    __syncwarp();
    return ret;
}

// Two instances of the same kernel, tuned for the two regimes of a batch:
// FAST = false: the GENERAL instance -- compact narrow phase (one item per lane and round, smaller instruction
//   footprint), 128 registers / 4 CTAs per SM: it owns the throughput phase (most worlds, short contact episodes).
//   A world whose timestep spent >= 32 calls in the single-contact fast path is flagged tier 3 ("trapped"): from the
//   next chunk on the other instance steps it.
// FAST = true : the TRAPPED instance -- two narrow-phase items per lane for ILP, 168 registers / 3 CTAs per SM: it owns
//   the latency-bound tail (tier 3 worlds) and hands a world back after 8 calm timesteps.
template <bool FAST>
#ifndef FIZZ_GEN_CTAS
#define FIZZ_GEN_CTAS 4
#endif
#ifndef FIZZ_FAST_CTAS
#define FIZZ_FAST_CTAS 3
#endif
__global__ void __launch_bounds__(CONTACT_BLOCK, FAST ? FIZZ_FAST_CTAS : FIZZ_GEN_CTAS)
contact_kernel(const __grid_constant__ KParams kp, const long long *__restrict__ list,
               const unsigned long long *__restrict__ count_ptr, unsigned long long *cursor,
               unsigned long long count_cap, int runner, const __grid_constant__ SpecParams sp) {
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
    if (tid == 0) {
        int p = 0;
        for (int i = 0; i < F; i++)
            for (int j = i + 1; j < F + X; j++)
                s.pair_meta[p++] = pack_pair(i, j, kp.nv[i], kp.nv[j], kp.vs[i], kp.vs[j]);
        for (int b = 0; b < F; b++)
            for (int v = 0; v < kp.nv[b]; v++) s.vbody[kp.vs[b] + v] = (unsigned char)b;
    }
    __syncthreads();

    // slot width of the narrow phase: smallest power of two >= the largest n1+n2 of the scene (host-computed)
    const int SWL = kp.swl;
    const int SW = 1 << SWL;
    const int PR = 32 >> SWL;                    // pairs per 32-lane set
    const int my_slot = lane >> SWL, my_k = lane & (SW - 1);
    const unsigned slot_mask = (SW == 32 ? FULL : ((1u << SW) - 1u)) << (my_slot * SW);

    // The trial that ends the halving search of a timestep in which EVERY trial fails: dt / 2 / 2 ... (:106) is dt * 2^-t
    // exactly as long as nothing gets subnormal; the first t >= 1 with dt_t <= TIME_TOL (:99).  (for trapped_run)
    int kstar_all;
    double dtk_all;
    halving_floor(kp, lane, kstar_all, dtk_all);

    const long long Wp = kp.Wp;
    // (runner bits 1, 2: this launch only runs when the list is at least / less than BULK_LIST worlds long -- the host
    //  launches both instances on the same list and the one that suits its length does the work)
    {
        const unsigned long long glen = sp.gate ? *sp.gate : *count_ptr;
        if (((runner & 2) && glen < BULK_LIST) || ((runner & 4) && glen >= BULK_LIST)) return;
    }
    // spec rounds (SpecParams, fizz_common.cuh): the items of this launch come from the work queue of the rounds -- (list
    // entry, slot) pairs -- not from the list
    const bool SPEC = FAST && sp.enabled != 0;
    const int SSW = spec_slot_words(F);
    const unsigned long long count = SPEC ? 0ull : min(*count_ptr, count_cap);  // (a capped list: classify_kernel's `cap`)
    const bool speculate = (kp.trace == nullptr);  // the contact trace needs every trial's tuples: sequential then
    // mutable per-world data: a spec launch re-reads what other SMs wrote earlier in the SAME launch -- past the L1
    auto ldm = [&](const auto *p) { return SPEC ? __ldcg(p) : *p; };
    // the clock as ONE value per warp: every decision taken on it (budgets of the spec items) must be warp-uniform, and
    // the lanes of a warp do not read the same cycle
    auto wclock = [&]() -> long long { return __shfl_sync(FULL, (long long)clock64(), 0); };

    // Work distribution: the first item of every warp is static (CTA c takes items 4c .. 4c+3), the rest is pulled from
    // the cursor.  A short list therefore keeps FEW CTAs resident (the others exit at once) instead of one warp in each:
    // the launches that run beside this one (spec rounds, the other instance) find room.
    bool first_item = true;
    for (;;) {
        unsigned long long idx = 0, wi = 0;
        int slot = 0;
        if constexpr (FAST) {
            if (SPEC) {
                // ---- next item of the rounds' queue: take a ticket, wait for its entry (or for the last world to finish) ----
                unsigned long long e = ~0ull;
                if (lane == 0) {
                    const unsigned long long t = atomicAdd(sp.state + 1, 1ULL);
                    volatile unsigned long long *qe = sp.queue + t % sp.qcap;
                    volatile unsigned long long *active = sp.state + 3;
                    const long long t_wait = clock64();
                    for (;;) {
                        const unsigned long long v = *qe;
                        if ((v >> 25) == t + 1) { e = v; break; }
                        if (*active == 0) break;
                        // (safety net, never seen: two seconds without an item -- leave; the launch after the rounds brings
                        //  every world to the chunk target anyway)
                        if (clock64() - t_wait > (1LL << FIZZ_SPEC_TIMEOUT_LOG2)) break;
                        __nanosleep(100);
                    }
                }
                e = __shfl_sync(FULL, e, 0);
                if (e == ~0ull) break;
                __threadfence();                       // (the entry was published after the state it refers to)
                wi = (e >> 5) & 0xfffffull;
                slot = (int)(e & 31);
            }
        }
        if (!SPEC) {
            if (first_item) {
                idx = (unsigned long long)blockIdx.x * CONTACT_WARPS + warp;
                first_item = false;
            } else {
                if (lane == 0) idx = (unsigned long long)gridDim.x * CONTACT_WARPS + atomicAdd(cursor, 1ULL);
                idx = __shfl_sync(FULL, idx, 0);
            }
            if (idx >= count) break;
            wi = idx;
        }
        FIZZ_ASSERT(wi < min(*count_ptr, count_cap));
        const long long w = list[wi];
        FIZZ_ASSERT(w >= 0 && w < kp.W);

        // ---- load the world into the warp's shared-memory slice ------------------------------------------------
        double heat = ldm(kp.heat + w);
        long long step = ldm(kp.steps + w);
        long long target = kp.target_steps;          // where this item stops
        // spec rounds: cost (units of 256 clock cycles) at which the item stops at the next timestep boundary / is given up
        // at once (only items that started from a PREDICTED state: a wrong prediction can be arbitrarily expensive to
        // evaluate, and nobody needs its result)
        const long long NOCAP = 0x3fffffffffffffffLL;
        long long call_budget = NOCAP, abort_cap = NOCAP;
        const long long t_item = wclock();   // cost is measured: 256 clock cycles a unit (one fast-path call, roughly)
        long long t_step = t_item;
        double *rec = nullptr;
        [[maybe_unused]] const double heat0 = heat;
        // spec rounds: this item is done (or had nothing to do).  The warp that finishes the LAST item of a world's round
        // commits the round (fizz_spec.cuh) and, unless the world is at the chunk target or has stopped, publishes the
        // items of its next round: a world's rounds follow one another without waiting for any other world's.
        auto spec_item_done = [&]() {
            __threadfence();
            __syncwarp();
            long long old = 0;
            long long *pending = sp.ctrl + w * SPEC_CTRL_WORDS + 4;
            if (lane == 0) old = (long long)atomicAdd(reinterpret_cast<unsigned long long *>(pending), ~0ULL);   // -1
            old = __shfl_sync(FULL, old, 0);
            FIZZ_ASSERT(old >= 1 && old <= SPEC_K);
            if (old == 1) {
                __threadfence();
                int knext = 0;
                spec_commit(kp, sp, w, sp.slots + wi * (unsigned)SPEC_K * (size_t)SSW, lane, knext);
                __syncwarp();
                if (knext > 0) {
                    unsigned long long base = 0;
                    if (lane == 0) {
                        atomicExch(reinterpret_cast<unsigned long long *>(pending), (unsigned long long)knext);
                        __threadfence();
                        base = atomicAdd(sp.state + 2, (unsigned long long)knext);
                    }
                    base = __shfl_sync(FULL, base, 0);
                    __threadfence();
                    if (lane < knext)
                        __stcg(sp.queue + (base + lane) % sp.qcap, ((base + lane + 1) << 25) | (wi << 5) | (unsigned long long)lane);
                } else if (lane == 0) {
                    __threadfence();
                    atomicAdd(sp.state + 3, ~0ULL);     // one world less in flight; at 0 the idle warps leave
                }
            }
            __syncwarp();
        };
        {
            double x = 0.0, y = 0.0, vx = 0.0, vy = 0.0;
            if (lane < F) {
                x = ldm(kp.pos + (2 * lane) * Wp + w); y = ldm(kp.pos + (2 * lane + 1) * Wp + w);
                vx = ldm(kp.vel + (2 * lane) * Wp + w); vy = ldm(kp.vel + (2 * lane + 1) * Wp + w);
            }
            bool skip = step >= kp.target_steps;
            // (a world that is at the target already -- listed by a reader that raced with another launch's write-back --
            //  must not be touched: update() below runs before the target is tested)
            if constexpr (FAST) {
                if (SPEC) {
                    // slot `slot` of this round: update() number step + slot * L ... of the world, from the PREDICTED state
                    const long long *ctrl = sp.ctrl + w * SPEC_CTRL_WORDS;
                    int Kw, Lw;
                    spec_shape(ldm(kp.tier + w), ldm(ctrl + 10), sp.S, sp.D, Kw, Lw);
                    const bool predictor = ldm(ctrl) > 1;
                    skip = skip || step < 1 || ldm(kp.status + w) != 0 || slot >= (predictor ? Kw : 1);
                    if (!skip && predictor) {
                        const long long my_start = step + (long long)slot * Lw;
                        if (my_start >= kp.target_steps) skip = true;
                        if (my_start + Lw < target) target = my_start + Lw;
                        if (!skip && lane < F && slot > 0) {
                            const unsigned long long cw = (unsigned long long)ldm(ctrl + 1 + (lane >> 2));
                            const int cx = (int)((cw >> (16 * (lane & 3))) & 255), cy = (int)((cw >> (16 * (lane & 3) + 8)) & 255);
                            const double rdt = kp.time_disc - dtk_all;
                            for (int t = slot * Lw; t > 0; t--) {
                                spec_predict(x, vx, cx, kp.time_disc, rdt, kp.accx);
                                spec_predict(y, vy, cy, kp.time_disc, rdt, kp.accy);
                            }
                        }
                        step = my_start;
                        // (abort_x4 / 4 times what the slot should cost by the world's own record, and never less than 2 D)
                        if (slot > 0) {
                            const long long expect = (long long)sp.abort_x4 * Lw * ((ldm(kp.tier + w) >> 32) & 0xffffff) / 4;
                            abort_cap = (expect > 2LL * sp.D ? expect : 2LL * sp.D) + 32;
                        }
                    } else if (!skip) {
                        // no usable predictor: this world advances sequentially, a bounded piece per round
                        if (step + sp.budget_steps < target) target = step + sp.budget_steps;
                        call_budget = sp.budget_cost;
                    }
                    if (skip) { spec_item_done(); continue; }
                    rec = sp.slots + (wi * (unsigned)SPEC_K + (unsigned)slot) * (size_t)SSW;
                    if (lane < F) { rec[4 * lane] = x; rec[4 * lane + 1] = y; rec[4 * lane + 2] = vx; rec[4 * lane + 3] = vy; }
                }
            }
            if (skip) continue;
            if (lane < F) {
                s.px[lane] = x; s.py[lane] = y; s.sx[lane] = x; s.sy[lane] = y;  // finish_update ended the last step
                s.vx[lane] = vx; s.vy[lane] = vy;
                s.thr[lane] = kp.thr[lane * Wp + w];
                s.mass[lane] = kp.mass[lane * Wp + w];
            }
        }
        for (int r = lane; r < 2 * V; r += 32) s.off[r] = kp.off[(size_t)r * Wp + w];
        for (int v = lane; v < V; v += 32) s.max_[v] = NAN;  // empty memo: NaN never compares equal
        const long long step_in = step;
        long long status = 0;
        long long ncalls_total = 0, fast_calls = 0;
        __syncwarp();

        bool resolve = false, verts_explicit = false;
        int k = 0, ncalls = 0, ns = 0;
        int fast_step = 0, calm = 0;  // tier-3 bookkeeping (scheduling hint only)
        int last_calls = 0;           // check_collisions() of the last completed timestep (scheduling hint only)
        int last_cost = 0;            // ... and what it cost, in units of 256 clock cycles (scheduling hint only)
        [[maybe_unused]] int two_fail = 0;  // consecutive failed attempts at the two-trapped-bodies shortcut
        bool trapped = false;
        int quiet = 0;        // consecutive contact-free timesteps: one check_collisions(), no tuple (scheduling hint)
        int quiet_empty = 0;  // ... of those, consecutive ones whose broad phase was empty
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
                    pairmeta_t m = 0;
                    if (p < P) {
                        m = s.pair_meta[p];
                        const int i = PM_II(m), j = PM_JJ(m);
                        double ox, oy, ot;
                        if (j < F) { ox = s.sx[j]; oy = s.sy[j]; ot = s.thr[j]; }
                        else { ox = s.fcx[j - F]; oy = s.fcy[j - F]; ot = s.fthr[j - F]; }
                        const double dx = s.sx[i] - ox, dy = s.sy[i] - oy;
                        const double d2 = sq2(dx, dy);
                        surv = !(d2 > s.thr[i] && d2 > ot);
                    }
                    const unsigned b = __ballot_sync(FULL, surv);
                    FIZZ_ASSERT(ns + __popc(b) <= SP_CAP);
                    if (surv) s.sp[ns + __popc(b & lt_mask)] = m;
                    ns += __popc(b);
                }
                __syncwarp();
            }
            ncalls++;
            ncalls_total++;
            if constexpr (FAST) {
                if (abort_cap != NOCAP && ((wclock() - t_item) >> 8) > abort_cap) { status = FIZZ_ST_SPEC_ABORT; break; }
            }

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
                // ---- narrow phase: (pair, axis) items in SW-wide slots, two sets of PR pairs per round ------------
                if constexpr (FAST) {
                    for (int p0 = 0; p0 < ns; p0 += 2 * PR) {
                        const int pa = p0 + my_slot, pb = p0 + PR + my_slot;
                        pairmeta_t ma = 0, mb = 0;
                        double nxa = 0.0, nya = 0.0, cura = INFINITY, nxb = 0.0, nyb = 0.0, curb = INFINITY;
                        bool sepa = false, sepb = false, acta = false, actb = false;
                        if (pa < ns) {
                            ma = s.sp[pa];
                            acta = my_k < PM_N1(ma) + PM_N2(ma);
                        }
                        if (pb < ns) {
                            mb = s.sp[pb];
                            actb = my_k < PM_N1(mb) + PM_N2(mb);
                        }
                        if (acta) sat_axis(s, F, ma, my_k, nxa, nya, cura, sepa);   // (FAST only)
                        if (actb) sat_axis(s, F, mb, my_k, nxb, nyb, curb, sepb);
    #pragma unroll
                        for (int set = 0; set < 2; set++) {
                            const pairmeta_t m = set ? mb : ma;
                            const bool act = set ? actb : acta, sep = set ? sepb : sepa;
                            const double nx = set ? nxb : nxa, ny = set ? nyb : nya, cur = set ? curb : cura;
                            const bool pair_live = (set ? pb : pa) < ns;
                            if (set == 1 && p0 + PR >= ns) break;  // uniform
                            const unsigned sepbits = __ballot_sync(FULL, sep);
                            const bool pair_sep = (sepbits & slot_mask) != 0;
                            // first strict minimum in axis order (:827-830) = lexicographic min of (overlap, axis)
                            const bool valid = act && (cur < INFINITY);
                            double bov = valid ? cur : INFINITY;
                            int bk = valid ? my_k : 99;
                            for (int o = 1; o < SW; o <<= 1) {
                                const double oov = __shfl_xor_sync(FULL, bov, o);
                                const int ok = __shfl_xor_sync(FULL, bk, o);
                                if (oov < bov || (oov == bov && ok < bk)) { bov = oov; bk = ok; }
                            }
                            const int src = (my_slot * SW + (bk == 99 ? 0 : bk)) & 31;
                            const double fxn = __shfl_sync(FULL, nx, src), fyn = __shfl_sync(FULL, ny, src);
                            const bool hit = pair_live && (my_k == 0) && !pair_sep && (bk != 99);  // :286-288
                            const unsigned hitb = __ballot_sync(FULL, hit);
                            if (hitb) {
                                if (hit) {
                                    const int slot = nh + __popc(hitb & lt_mask);
                                    const double hx = fxn * -1.0, hy = fyn * -1.0;  // :838-839 always negated (Q7)
                                    const int ii = PM_II(m), jj = PM_JJ(m);
                                    if (slot < FIZZ_MAX_HITS) {
                                        s.hi[slot] = (unsigned char)ii; s.hj[slot] = (unsigned char)jj;
                                        s.hnx[slot] = hx; s.hny[slot] = hy; s.hmag[slot] = bov;
                                    }
                                    if (kp.trace) trace_hit(kp, w, step, ncalls, ii, jj, hx, hy, bov);
                                }
                                nh += __popc(hitb);
                                // max |overlap| over this set's hits: slot leaders hold them
                                double am = hit ? fabs(bov) : -1.0;
                                for (int o = SW; o < 32; o <<= 1) {
                                    const double other = __shfl_xor_sync(FULL, am, o);
                                    if (other > am) am = other;
                                }
                                am = __shfl_sync(FULL, am, 0);
                                if (am > maxov) maxov = am;
                            }
                        }
                    }
                } else {
    #pragma unroll 1
                    for (int p0 = 0; p0 < ns; p0 += PR) {
                        const int pa = p0 + my_slot;
                        pairmeta_t m = 0;
                        double nx = 0.0, ny = 0.0, cur = INFINITY;
                        bool sep = false, act = false;
                        const bool pair_live = pa < ns;
                        if (pair_live) {
                            m = s.sp[pa];
                            act = my_k < PM_N1(m) + PM_N2(m);
                        }
                        if (act) sat_axis_compact(s, F, m, my_k, nx, ny, cur, sep);   // (!FAST only)
                        const unsigned sepbits = __ballot_sync(FULL, sep);
                        const bool pair_sep = (sepbits & slot_mask) != 0;
                        const bool valid = act && (cur < INFINITY);
                        double bov = valid ? cur : INFINITY;
                        int bk = valid ? my_k : 99;
                        for (int o = 1; o < SW; o <<= 1) {
                            const double oov = __shfl_xor_sync(FULL, bov, o);
                            const int ok = __shfl_xor_sync(FULL, bk, o);
                            if (oov < bov || (oov == bov && ok < bk)) { bov = oov; bk = ok; }
                        }
                        const int src = (my_slot * SW + (bk == 99 ? 0 : bk)) & 31;
                        const double fxn = __shfl_sync(FULL, nx, src), fyn = __shfl_sync(FULL, ny, src);
                        const bool hit = pair_live && (my_k == 0) && !pair_sep && (bk != 99);  // :286-288
                        const unsigned hitb = __ballot_sync(FULL, hit);
                        if (hitb) {
                            if (hit) {
                                const int slot = nh + __popc(hitb & lt_mask);
                                const double hx = fxn * -1.0, hy = fyn * -1.0;  // :838-839 always negated (Q7)
                                const int ii = PM_II(m), jj = PM_JJ(m);
                                if (slot < FIZZ_MAX_HITS) {
                                    s.hi[slot] = (unsigned char)ii; s.hj[slot] = (unsigned char)jj;
                                    s.hnx[slot] = hx; s.hny[slot] = hy; s.hmag[slot] = bov;
                                }
                                if (kp.trace) trace_hit(kp, w, step, ncalls, ii, jj, hx, hy, bov);
                            }
                            nh += __popc(hitb);
                            double am = hit ? fabs(bov) : -1.0;
                            for (int o = SW; o < 32; o <<= 1) {
                                const double other = __shfl_xor_sync(FULL, am, o);
                                if (other > am) am = other;
                            }
                            am = __shfl_sync(FULL, am, 0);
                            if (am > maxov) maxov = am;
                        }
                    }
                }
                __syncwarp();
                if (nh > FIZZ_MAX_HITS) { status = FIZZ_ST_HIT_OVERFLOW; break; }
            }

            if (!resolve) {
                if (maxov > kp.coll_tol && dt > kp.time_tol) {  // :99
                    if (k == 0 && speculate) {
                        // ---- speculative halving: lane t evaluates trial t (dt / 2^t) on its own -----------------
                        // dt / 2 / 2 ... (:106, `lane` times) is dt * 2^-lane exactly as long as nothing gets subnormal
                        double dtt = dt;
                        if (dt > 1e-280) dtt = dt * __longlong_as_double((long long)(1023 - lane) << 52);
                        else for (int t = 0; t < lane; t++) dtt = dtt / 2;
                        // Shortcut for a body that sits DEEP in a fixed one (the trapped pattern: every trial fails and the
                        // loop ends on TIME_TOL): every trial position of that body lies within D of the position the
                        // failed check saw -- pos(h) - pos(dt) = v (h - dt) + a0/2 (h^2 - dt^2), 0 < h <= dt -- and a
                        // translation by d moves both ends of its projection on a unit axis by the same s, |s| <= |d|.
                        // Whichever polygon :815 then calls "the first", the overlap the axis reports is one of
                        // hi2 - lo1 - s and hi1 - lo2 + s, so min(hi2 - lo1, hi1 - lo2) - D bounds it from below.  If
                        // that bound clears COLLISION_TOL (with a margin six orders above the rounding of the
                        // projections) on every axis of the pair, the pair collides deeper than the tolerance in
                        // every trial and no trial needs to be evaluated.  The broad phase is not an issue: all trials
                        // of a timestep run on the survivors of the same stale snapshot (Q3).
                        // (only in the latency-tuned instance: the first-chunk instance rarely sees deep contacts and pays for
                        //  every instruction of footprint; and only when the reported overlap itself, an upper bound of the
                        //  floor, clears the bar)
                        bool all_fail = false;
                        if constexpr (FAST) {
                            int hm = -1;
                            double best = kp.coll_tol;
                            for (int h = 0; h < nh; h++)
                                if (s.hj[h] >= F && fabs(s.hmag[h]) > best) { best = fabs(s.hmag[h]); hm = h; }
                            if (hm >= 0) {  // (warp-uniform)
                                const int ii = s.hi[hm], jj = s.hj[hm];
                                const int na = kp.nv[ii] + kp.nv[jj];
                                const pairmeta_t m = pack_pair(ii, jj, kp.nv[ii], kp.nv[jj], kp.vs[ii], kp.vs[jj]);
                                double D = (fabs(s.vx[ii]) + fabs(s.vy[ii])) * dt + .5 * (fabs(a0x) + fabs(a0y)) * dt * dt;
                                D = D * 1.000001 + 1e-9;
                                bool clear = (best - D) > (kp.coll_tol + 1e-7);
                                if (clear && lane < na) {
                                    double fnx, fny, fl;
                                    bool fsep;
                                    sat_axis<true>(s, F, m, lane, fnx, fny, fl, fsep);
                                    clear = (fl - D) > (kp.coll_tol + 1e-7 + 1e-9 * (fabs(fl) + D));  // (false for NaN)
                                }
                                all_fail = na <= 32 && __all_sync(FULL, clear);
                            }
                        }
                        bool done_t = true;                              // "this trial ends the :99 loop"
                        // trials exist while the previous one still had dt > time_tol; lane t >= 1 is a real trial iff
                        // dt_{t-1} > time_tol, which is monotone, so the first `done` lane below is always real
                        if (lane >= 1 && all_fail) {
                            done_t = !(dtt > kp.time_tol);
                        } else if (lane >= 1) {
                            // :99 only asks whether the largest overlap exceeds COLLISION_TOL: one pair above it decides
                            // the trial
                            double mo = 0.0;
                            for (int p = 0; p < ns && !(mo > kp.coll_tol); p++) {
                                const pairmeta_t m = s.sp[p];
                                const int ii = PM_II(m), jj = PM_JJ(m);
                                double c1x, c1y, c2x = 0.0, c2y = 0.0;
                                [[maybe_unused]] double dvx, dvy;
                                FIZZ_MOVE_POINT(s.px[ii], s.py[ii], s.vx[ii], s.vy[ii], dtt, a0x, a0y, kp.accx, kp.accy, c1x,
                                                c1y, dvx, dvy);
                                if (jj < F) {
                                    FIZZ_MOVE_POINT(s.px[jj], s.py[jj], s.vx[jj], s.vy[jj], dtt, a0x, a0y, kp.accx, kp.accy,
                                                    c2x, c2y, dvx, dvy);
                                }
                                double mag;
                                bool hit_t;
                                if constexpr (FAST) hit_t = sat_pair_trial(s, F, m, c1x, c1y, c2x, c2y, mag);
                                else hit_t = sat_pair_trial_compact(s, F, m, c1x, c1y, c2x, c2y, mag);
                                if (hit_t)
                                    if (mag > mo) mo = mag;
                            }
                            done_t = !(mo > kp.coll_tol && dtt > kp.time_tol);
                        } else {
                            done_t = false;  // trial 0 is the one that just failed
                        }
                        const unsigned doneb = __ballot_sync(FULL, done_t);
                        const int kstar = __ffs(doneb) - 1;  // >= 1: dt_31 < time_tol for any sane time_tol
                        if (kstar >= 1) {
                            // trials 1 .. kstar-1 were made (and failed); trial kstar is run for real next round
                            ncalls += kstar - 1;
                            ncalls_total += kstar - 1;
                            k = kstar;
                            dt = __shfl_sync(FULL, dtt, kstar);
                            continue;
                        }
                    }
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
                    const int n1 = kp.nv[ii], v1 = kp.vs[ii];
                    if (jj >= F) {                             // :171-198 free body vs fixed body
                        if (lane < n1) { s.vert[2 * (v1 + lane)] += nvx; s.vert[2 * (v1 + lane) + 1] += nvy; }  // :174
                        const bool gate = sq2(nvx, nvy) > kp.vib_thr;  // :180 norm(nv) > VIBRATE_TOL
                        double cpx = s.px[ii], cpy = s.py[ii], cvx = s.vx[ii], cvy = s.vy[ii];
                        const double m1 = s.mass[ii];
                        if (gate) { cpx += nvx; cpy += nvy; }                           // :181
                        const double vdotN = dot2(cvx, cvy, ux, uy);                    // :183
                        const double cc = kp.ope * vdotN;                               // :188
                        cvx -= cc * ux;
                        cvy -= cc * uy;
                        heat += .5 * m1 * kp.ome2 * (vdotN * vdotN);                    // :191,:195
                        if (kp.mu > 0.0) ext_friction_fixed(kp.mu, kp.ope, m1, ux, uy, vdotN, cvx, cvy, heat);  // (extension)
                        __syncwarp();
                        if (lane == 0) {
                            s.px[ii] = cpx; s.py[ii] = cpy; s.vx[ii] = cvx; s.vy[ii] = cvy;
                            s.sx[ii] = cpx; s.sy[ii] = cpy;                             // :197
                        }
                    } else {                                   // :201-223 two free bodies
                        const int n2 = kp.nv[jj], v2 = kp.vs[jj];
                        const double m1 = s.mass[ii], m2 = s.mass[jj];
                        const double mtotal = m1 + m2;
                        const double mprop1 = m1 / mtotal;
                        const double mprop2 = 1 - mprop1;
                        double c1x = s.px[ii], c1y = s.py[ii], v1x = s.vx[ii], v1y = s.vy[ii];
                        double c2x = s.px[jj], c2y = s.py[jj], v2x = s.vx[jj], v2y = s.vy[jj];
                        const double vn1 = dot2(v1x, v1y, ux, uy);                      // :207
                        const double vn2 = dot2(v2x, v2y, ux, uy);                      // :208
                        double v1u = (mprop1 - mprop2) * vn1 + 2 * mprop2 * vn2;        // :209
                        double v2u = 2 * mprop1 * vn1 + (mprop2 - mprop1) * vn2;        // :218
                        if (kp.ffr >= 0.0) ext_restitution_free(kp.ffr, m1, mprop1, mprop2, vn1, vn2, v1u, v2u, heat);  // (extension)
                        const double d1 = v1u - vn1, d2 = v2u - vn2;                    // :210,:219
                        const double p1x = mprop1 * nvx, p1y = mprop1 * nvy;
                        const double p2x = mprop2 * nvx, p2y = mprop2 * nvy;
                        c1x += p1x; c1y += p1y; v1x += d1 * ux; v1y += d1 * uy;         // :206,:210
                        c2x -= p2x; c2y -= p2y; v2x += d2 * ux; v2y += d2 * uy;         // :217,:219
                        if (kp.mu > 0.0)                                                // (extension)
                            ext_friction_free(kp.mu, kp.ffr >= 0.0 ? kp.ffr : 1.0, m1, m2, mprop2, ux, uy, vn1, vn2, v1x, v1y,
                                              v2x, v2y, heat);
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
#ifndef FIZZ_NO_TRAP
                if constexpr (FAST) {
                    // ---- TWO bodies trapped at once, each against fixed bodies --------------------------------------------
                    // Every check_collisions() of the resolve loop returns [(A, fixed), (B, fixed)] until one of them is out.
                    // A resolve pass applies each tuple to its own body and nothing else moves, so A's sequence of pushes
                    // is the one it would make alone, and so is B's -- provided the pair (A, B) stays outside the broad
                    // phase for every position the two go through (checked afterwards on the bounding boxes of the two
                    // trajectories; a failed check restores both bodies and the general path runs the loop).  The loop then
                    // makes max(n_A, n_B) checks.  Heat is added tuple by tuple in list order, which the two separate loops
                    // cannot reproduce: only taken when ELASTICITY == 1 (every term is +0; a non-finite one restores too).
                    if (speculate && nh == 2 && SW <= 16 && two_fail < 2 && s.hj[0] >= F && s.hj[1] >= F && s.hi[0] != s.hi[1] &&
                        kp.ome2 == 0.0 && !(kp.mu > 0.0)) {
                        const int bA = s.hi[0], bB = s.hi[1];
                        const int nA = kp.nv[bA], nB = kp.nv[bB];
                        if ((nA == 3 || nA == 4) && (nB == 3 || nB == 4)) {
                            // this lane's share of what a failed attempt has to put back
                            double *slotp = nullptr;
                            if (lane < 2 * nA) slotp = s.vert + 2 * kp.vs[bA] + lane;
                            else if (lane >= 8 && lane < 8 + 2 * nB) slotp = s.vert + 2 * kp.vs[bB] + (lane - 8);
                            else if (lane >= 16 && lane < 28) {
                                const int bb = (lane < 22) ? bA : bB, f = (lane - 16) % 6;
                                slotp = (f == 0 ? s.px : f == 1 ? s.py : f == 2 ? s.vx : f == 3 ? s.vy : f == 4 ? s.sx : s.sy) + bb;
                            }
                            const double saved = slotp ? *slotp : 0.0;
                            const int nc0 = ncalls;
                            const long long tot0 = ncalls_total;
                            int sh = 0, lc = 0, kk2 = k;
                            long long st2 = step;
                            double dt2 = dt;
                            TrapBox boxes[2];
                            // A's loop, then B's.  A run ends EMPTY (its :234 check found b free: counted) or MIDSTEP (it stands
                            // right before a check it could not classify: not counted).  The joint state the general path can
                            // continue from must have both bodies at the SAME check: a body that is out of contact stays where
                            // it is, so "A empty after pa checks, B before its check pb + 1 > pa" is such a state, and so is
                            // "A before check pa + 1, B run for at most pa checks" (B's budget is cut to pa).
                            int cons[2] = {0, 0}, rets[2] = {TRAP_MIDSTEP, TRAP_MIDSTEP};
                            bool valid = true;
#pragma unroll 1
                            for (int which = 0; which < 2 && valid; which++) {     // (one call site: instruction footprint)
                                long long tot = tot0;
                                double hx = 0.0;
                                int nc_in = nc0;
                                if (which == 1 && rets[0] == TRAP_MIDSTEP) nc_in = kp.max_calls - cons[0];   // budget = pa checks
                                int nc = nc_in;
                                rets[which] = trapped_run<false, true>(s, kp, F, X, lane, which ? bB : bA, hx, nc, tot, st2, kk2, dt2, sh,
                                                                       lc, kstar_all, dtk_all, target, NOCAP, NOCAP, &boxes[which]);
                                cons[which] = nc - nc_in;
                                valid = (hx == 0.0) && (which == 1 || rets[0] == TRAP_EMPTY || cons[0] >= 1);
                            }
                            const TrapBox boxA = boxes[0], boxB = boxes[1];
                            const int pa = cons[0], pb = cons[1];
                            int joint = -1;   // checks consumed by the joint state, or -1
                            bool finished = false;
                            if (valid) {
                                if (rets[0] == TRAP_EMPTY && rets[1] == TRAP_EMPTY) { joint = pa > pb ? pa : pb; finished = true; }
                                else if (rets[0] == TRAP_EMPTY && pb >= pa) joint = pb;
                                else if (rets[0] == TRAP_MIDSTEP && ((rets[1] == TRAP_EMPTY && pb <= pa) || (rets[1] == TRAP_MIDSTEP && pb == pa)))
                                    joint = pa;
                                valid = joint >= 0;
                            }
                            if (valid) {
                                // no position of A's trajectory within max(r_A, r_B) of any position of B's (:277-279)
                                const double gx = fmax(fmax(boxA.xlo - boxB.xhi, boxB.xlo - boxA.xhi), 0.0);
                                const double gy = fmax(fmax(boxA.ylo - boxB.yhi, boxB.ylo - boxA.yhi), 0.0);
                                const double tm = fmax(s.thr[bA], s.thr[bB]) * 1.000001 + 1e-9;
                                valid = (gx * gx > tm) || (gy * gy > tm);
                            }
                            if (valid) {
                                ncalls = nc0 + joint;
                                ncalls_total = tot0 + joint;
                                fast_calls += joint;
                                fast_step += joint;
                                verts_explicit = true;
                                two_fail = 0;
                                if (finished) goto end_of_update;   // the :234 call that found nothing has been made
                                continue;                           // the general path makes the next check of this loop
                            }
#ifdef FIZZ_DEBUG_TWO
                            if (lane == 0 && step < 400)
                                printf("two-body attempt failed: w %lld step %lld bA %d bB %d rets %d %d pa %d pb %d nc0 %d boxA [%g %g %g %g] boxB [%g %g %g %g] thr %g %g\n",
                                       w, step, bA, bB, rets[0], rets[1], pa, pb, nc0, boxA.xlo, boxA.xhi, boxA.ylo, boxA.yhi, boxB.xlo, boxB.xhi,
                                       boxB.ylo, boxB.yhi, s.thr[bA], s.thr[bB]);
#endif
                            __syncwarp();
                            if (slotp) *slotp = saved;         // put both bodies back: the general path runs this loop
                            two_fail++;
                            __syncwarp();
                        }
                    }
                }
                if (speculate && nh == 1 && SW <= 16 && s.hj[0] >= F && !(kp.mu > 0.0)) {   // (friction: general path only)
                    const int b = s.hi[0];
                    const long long before = ncalls_total;
                    const int nb = kp.nv[b];
                    int ret = TRAP_MIDSTEP, steps_here = 0;
                    // (one instantiation for triangles and quadrilaterals; anything else stays on the general path)
                    const long long t_call = wclock();
                    const long long cost_now = (t_call - t_item) >> 8;
                    if (nb == 3 || nb == 4)
                        ret = trapped_run<FAST>(s, kp, F, X, lane, b, heat, ncalls, ncalls_total, step, k, dt, steps_here,
                                                last_calls, kstar_all, dtk_all, target, call_budget - cost_now + ncalls_total,
                                                abort_cap - cost_now + ncalls_total);
                    fast_calls += ncalls_total - before;
                    if (steps_here) {
                        // whole timesteps went by in registers: all of them trapped ones
                        trapped = true; calm = 0; quiet = 0; quiet_empty = 0; fast_step = 0;
                        t_step = wclock();
                        last_cost = (int)(((t_step - t_call) >> 8) / steps_here);
                    }
                    if (ret == TRAP_BOUNDARY) {
                        // a fresh update() (or the target): the vertices are derived state again
                        resolve = false; verts_explicit = false;
                        if (step >= target || (call_budget != NOCAP && ((wclock() - t_item) >> 8) >= call_budget)) break;
                        continue;
                    }
                    verts_explicit = true;
                    fast_step += (int)(ncalls_total - before);
                    if (ret == TRAP_EMPTY) goto end_of_update;  // the fast path made the :234 call that found nothing
                }
#endif
                continue;
            }
        end_of_update:

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
            last_calls = ncalls;
            {
                const long long now = wclock();
                last_cost = (int)((now - t_step) >> 8);
                t_step = now;
            }
            quiet = (ncalls == 1) ? quiet + 1 : 0;  // (a step with a tuple makes at least two calls)
            quiet_empty = (ncalls == 1 && ns == 0) ? quiet_empty + 1 : 0;
            resolve = false; k = 0; dt = kp.time_disc; ncalls = 0;
            if constexpr (FAST) {
                calm = (fast_step < 32) ? calm + 1 : 0;
                fast_step = 0;
            } else {
                // a timestep with a long single-contact resolve loop: the latency-tuned instance takes the world
                const bool heavy = fast_step >= 32;
                fast_step = 0;
                if (heavy) trapped = true;  // flagged for the NEXT chunk; this chunk is finished here (no idle time)
            }
            if (step >= target || (call_budget != NOCAP && ((wclock() - t_item) >> 8) >= call_budget)) break;
            verts_explicit = false;
        }

        __syncwarp();
        if constexpr (FAST) {
            if (SPEC) {
                // ---- spec round: the result goes to this item's slot record; the commit decides (fizz_spec.cuh) ------------
                bool fin = true;
                if (lane < F) {
                    const double x = s.px[lane], y = s.py[lane];
                    double *e = rec + 4 * F + 4 * lane;
                    e[0] = x; e[1] = y; e[2] = s.vx[lane]; e[3] = s.vy[lane];
                    fin = (fabs(x) < INFINITY) && (fabs(y) < INFINITY);
                }
                if (status == 0 && !__all_sync(FULL, fin)) status = FIZZ_ST_NONFINITE;
                // slot 0 started from the real state: whatever it found is the world's (the vertices of a timestep that
                // ended in a resolve pass are explicit state, written where a normal launch writes them)
                if (slot == 0 && verts_explicit)
                    for (int r = lane; r < 2 * V; r += 32) kp.verts[(size_t)r * Wp + w] = s.vert[r];
                if (lane == 0) {
                    const bool clean = status == 0 && !verts_explicit &&
                                       __double_as_longlong(heat) == __double_as_longlong(heat0);
                    long long *m = reinterpret_cast<long long *>(rec + 8 * F);
                    m[1] = step - step_in; m[2] = ncalls_total; m[3] = fast_calls; m[4] = last_calls;
                    rec[8 * F + 5] = heat; m[6] = status; m[7] = verts_explicit ? 1 : 0;
                    m[8] = ((clock64() - t_item) >> 8) / (step > step_in ? step - step_in : 1);   // cost per timestep of this item
                    m[0] = 1 | (clean ? 2 : 0);
                    if (sp.dbg) {
                        const unsigned long long cyc = (unsigned long long)(clock64() - t_item);
                        atomicMax(sp.dbg + 4 * sp.round, cyc);
                        atomicAdd(sp.dbg + 4 * sp.round + 1, cyc);
                        atomicAdd(sp.dbg + 4 * sp.round + 2, 1ULL);
                        if (status == FIZZ_ST_SPEC_ABORT) atomicAdd(sp.dbg + 4 * sp.round + 3, 1ULL);
                    }
                }
                spec_item_done();
                continue;
            }
        }
#ifdef FIZZ_RUNNER_DEBUG
        if (lane == 0 && w < 65536 && step_in >= 2688) atomicAdd(&g_dbg_cycles[w], (unsigned long long)(clock64() - t_item));
#endif
        // ---- write the world back ----------------------------------------------------------------------------
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
            if (kp.general_calls) kp.general_calls[w] += ncalls_total - fast_calls;   // (introspection: which worlds cost time)
            kp.vflag[w] = verts_explicit ? 1 : 0;
            // hint for the next chunk: 0 quiescent (lean thread-per-world kernel), 1 candidates but no contact (prover
            // variant, when enabled), 2 contact phase (stay here)
            const long long calm_tier = (quiet_empty >= 8) ? 0 : ((quiet >= 2) ? 1 : 2);
            long long t = FAST ? ((calm >= 8) ? calm_tier : 3) : (trapped ? 3 : calm_tier);
            // a runner launch never releases its worlds itself (tier 4 = "owned by a runner"): it runs on its own stream,
            // and a work-list kernel of the chunk pipeline reading steps/tier while they are stored here could list a
            // finished world again.  The host demotes tier 4 after the join at the end of fizz_step().
            if (runner & 1) t = 4;
            // bits 8..31: check_collisions() of the last completed timestep: how heavy the world is NOW
            // bits 32..55: cost per timestep, averaged over this launch's timesteps (fizz_common.cuh: spec_shape)
            last_cost = (int)min((long long)0xffffff, ((clock64() - t_item) >> 8) / (step > step_in ? step - step_in : 1));
            // bit 56: "the spec rounds cannot predict this world" (set by their commit step, sticky)
            kp.tier[w] = t | ((long long)(last_calls & 0xffffff) << 8) | ((long long)(last_cost & 0xffffff) << 32) |
                         (kp.tier[w] & (1LL << 56));
            atomicAdd(kp.counters + 1, (unsigned long long)(step - step_in));
            atomicAdd(kp.counters + 2, (unsigned long long)ncalls_total);
            atomicAdd(kp.counters + 3, (unsigned long long)fast_calls);
        }
        __syncwarp();
    }
}

// Work list for the contact kernel: every running world that has not reached the chunk target.
// The tier word holds the scheduling hint in bits 0..7 and the check_collisions() count of the world's last completed
// timestep in bits 8..31 and its cost in bits 32..55.  `use_rate`: only worlds whose last timestep made [rate_lo, rate_hi)
// calls (1) or cost that much (2) are taken: lets the host append the heaviest worlds first / split the worlds between
// the spec rounds and the ordinary launch.
// `mark` >= 0: the selected worlds get tier = mark (runner mode: 4 = "owned by a runner launch", which keeps every other
// kernel of the pipeline away from them until the host releases them after the join).
__global__ void classify_kernel(const long long *status, const long long *steps, long long *tier, long long W,
                                long long target, long long tier_lo, long long tier_hi, int use_rate,
                                long long rate_lo, long long rate_hi, long long *list, unsigned long long *count,
                                long long mark = -1, unsigned long long cap = ~0ull, int flagmode = 0,
                                const unsigned long long *gate = nullptr) {
    if (gate && *gate >= BULK_LIST) return;      // (a throughput phase: no spec rounds, see fizz_step)
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tw = 0;
    bool lag = false;
    if (w < W) {
        tw = tier[w];                                  // (tier before steps: see the write-back of the contact kernel)
        const long long t = tw & 255;
        lag = status[w] == 0 && t >= tier_lo && t <= tier_hi && steps[w] < target;
    }
    if (lag && use_rate) {
        const long long r = (use_rate == 2) ? ((tw >> 32) & 0xffffff) : ((tw >> 8) & 0xffffff);   // 2: by cost
        const bool in = r >= rate_lo && r < rate_hi, nospec = ((tw >> 56) & 1) != 0;
        // flagmode 1: the spec list (in range and predictable); 2: its complement (below the range, or unpredictable);
        // ... 3: in range and unpredictable (the runner launches of spec mode)
        lag = (flagmode == 1) ? (in && !nospec) : ((flagmode == 2) ? (in || nospec) : ((flagmode == 3) ? (in && nospec) : in));
    }
    const unsigned b = __ballot_sync(0xffffffffu, lag);
    if (b == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned long long idx = base + __popc(b & ((1u << lane) - 1u));
    if (lag && idx < cap) {       // (beyond `cap` a world is neither listed nor marked: the count is clamped by the reader)
        list[idx] = w;
        if (mark >= 0) tier[w] = (tw & ~255LL) | mark;
    }
}

// End of a fizz_step() call, after the join with the runner launches: their worlds go back to the chunk pipeline as
// "trapped" (tier 3).
__global__ void release_kernel(long long *tier, long long W) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < W && (tier[w] & 255) == 4) tier[w] = (tier[w] & ~255LL) | 3;
}

// Spec rounds of a chunk, start: every listed world gets the items of its first round (fizz_common.cuh: SpecParams).
__global__ void spec_reset_kernel(SpecParams sp) {
    // tickets taken by warps that left when the last chunk's rounds ended were never produced: skip them
    if (sp.state[1] > sp.state[2]) sp.state[2] = sp.state[1];
    sp.state[3] = 0;
}
__global__ void spec_seed_kernel(const __grid_constant__ KParams kp, SpecParams sp, const long long *__restrict__ list,
                                 unsigned long long cap) {
    const unsigned long long n = min(sp.state[0], cap);
    const unsigned long long wi = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= n) return;
    const long long w = list[wi];
    long long *ctrl = sp.ctrl + w * SPEC_CTRL_WORDS;
    if (kp.steps[w] >= kp.target_steps || kp.steps[w] < 1 || kp.status[w] != 0) return;
    int Kw, Lw;
    spec_shape(kp.tier[w], ctrl[10], sp.S, sp.D, Kw, Lw);
    const int k = ctrl[0] > 1 ? Kw : 1;
    ctrl[4] = k;
    atomicAdd(sp.state + 3, 1ULL);
    const unsigned long long base = atomicAdd(sp.state + 2, (unsigned long long)k);
    for (int i = 0; i < k; i++)
        sp.queue[(base + i) % sp.qcap] = ((base + i + 1) << 25) | (wi << 5) | (unsigned long long)i;
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
