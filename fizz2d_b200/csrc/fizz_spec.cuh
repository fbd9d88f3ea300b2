// fizz_spec.cuh -- the commit step of the time-parallel "spec rounds" (SpecParams in fizz_common.cuh).  The rounds of a
// chunk are ONE launch of contact_kernel<true> whose warps take (listed world, slot) items from a ticket queue
// (fizz_contact_kernel.cuh); fizz_engine.cu: fizz_step lists the worlds, seeds the queue and launches it.
//
// After a round every slot record holds the state its update()s STARTED from and the state they ENDED in.  Slot 0
// started from the world's real state (the slab); slot k was started from a prediction.  The warp that finishes the last
// item of a world's round commits the round: it accepts the longest prefix of slots in which every slot's start equals
// the end of the slot before, bit for bit and for every free body -- the state update() t+1 depends on is (position,
// velocity) of the centres of mass: vertices are derived (physics.py:358-359), the snapshot of the broad phase is the
// position finish_update left (:453-464), the heat of a "clean" slot did not change; it is what a chunk boundary of the
// ordinary pipeline carries, too -- writes the last accepted end state to the slab, derives the predictor and the window
// of the next round from what was accepted, and says how many items that round has (the caller publishes them).
// World.update() semantics are untouched: an accepted slot IS update() on the real state, a rejected one is forgotten.
#pragma once

#include "fizz_common.cuh"

namespace fizz {

// Warp-wide (all 32 lanes).  `rec0`: slot records of the world's list entry.
// `knext`: items of the world's next round (0: the world is at the chunk target, or has stopped).
__device__ __forceinline__ void spec_commit(const KParams &kp, const SpecParams &sp, long long w, double *rec0, int lane,
                                            int &knext) {
    knext = 0;
    const unsigned FULL = 0xffffffffu;
    const int F = kp.F, SW = spec_slot_words(F);
    const long long Wp = kp.Wp;
    long long *ctrl = sp.ctrl + w * SPEC_CTRL_WORDS;

    // ---- lane k: is slot k the continuation of slot k-1? -------------------------------------------------------
    // (the records were written by other warps, possibly on other SMs, each followed by a __threadfence and the atomic that
    //  made this warp the last one: read them past the L1)
    double *rec = rec0 + (size_t)lane * SW;
    long long *meta = reinterpret_cast<long long *>(rec + 8 * F);
    const long long flags = __ldcg(meta);
    const bool ran = (flags & 1) != 0, clean = (flags & 2) != 0;
    bool link = ran;
    if (lane > 0 && ran) {
        const double *prev = rec - SW;
        const long long pf = __ldcg(reinterpret_cast<const long long *>(prev + 8 * F));
        link = clean && (pf & 3) == 3;           // (an unclean slot is only ever accepted as slot 0: it ends the prefix)
        for (int i = 0; i < 4 * F && link; i++)
            link = __double_as_longlong(__ldcg(rec + i)) == __double_as_longlong(__ldcg(prev + 4 * F + i));
    }
    const unsigned okb = __ballot_sync(FULL, link);
    const int m = (~okb == 0u) ? 32 : __ffs(~okb) - 1;             // leading accepted slots
    const unsigned ranb = __ballot_sync(FULL, ran);
    meta[0] = 0;                                                   // consumed
    if (m == 0) return;                                            // nothing ran (the world is at the target, or stopped)
    const long long old_steps = __ldcg(kp.steps + w);

    // ---- totals of the accepted prefix ----------------------------------------------------------------------------
    long long adv = (lane < m) ? __ldcg(meta + 1) : 0, calls = (lane < m) ? __ldcg(meta + 2) : 0;
    long long fast = (lane < m) ? __ldcg(meta + 3) : 0, evaluated = ran ? __ldcg(meta + 1) : 0;
    long long cost_sum = (lane < m) ? __ldcg(meta + 8) * __ldcg(meta + 1) : 0;    // (cost per timestep x timesteps)
    for (int o = 16; o > 0; o >>= 1) {
        cost_sum += __shfl_xor_sync(FULL, cost_sum, o);
        adv += __shfl_xor_sync(FULL, adv, o);
        calls += __shfl_xor_sync(FULL, calls, o);
        fast += __shfl_xor_sync(FULL, fast, o);
        evaluated += __shfl_xor_sync(FULL, evaluated, o);
    }
    const double *last = rec0 + (size_t)(m - 1) * SW;
    const long long *lmeta = reinterpret_cast<const long long *>(last + 8 * F);
    const bool last_clean = (__shfl_sync(FULL, (int)flags, m - 1) & 2) != 0;
    // ---- the world's new state --------------------------------------------------------------------------------------
    if (lane < F) {
        const double *e = last + 4 * F + 4 * lane;
        kp.pos[(2 * lane) * Wp + w] = __ldcg(e);
        kp.pos[(2 * lane + 1) * Wp + w] = __ldcg(e + 1);
        kp.vel[(2 * lane) * Wp + w] = __ldcg(e + 2);
        kp.vel[(2 * lane + 1) * Wp + w] = __ldcg(e + 3);
    }
    const long long new_status = last_clean ? 0 : __ldcg(lmeta + 6);
    const long long new_tier = 3 | ((__ldcg(lmeta + 4) & 0xffffff) << 8) | ((min(cost_sum / (adv > 0 ? adv : 1), 0xffffffLL)) << 32);
    if (lane == 0) {
        kp.steps[w] = old_steps + adv;
        // (read-modify-write of words another SM may have updated earlier in this launch: atomics work at the L2)
        atomicAdd(reinterpret_cast<unsigned long long *>(kp.calls + w), (unsigned long long)calls);
        if (kp.general_calls) atomicAdd(reinterpret_cast<unsigned long long *>(kp.general_calls + w), (unsigned long long)(calls - fast));
        if (last_clean) {
            kp.vflag[w] = 0;                       // vertices are derived: off + com
        } else {                                   // (m == 1: slot 0 ended in a way the slots after it cannot continue)
            kp.heat[w] = __ldcg(last + 8 * F + 5);
            kp.status[w] = new_status;
            kp.vflag[w] = __ldcg(lmeta + 7);
        }
        // still a trapped world as far as the schedule is concerned; its rate and cost as of the last accepted timestep
        kp.tier[w] = new_tier;
        atomicAdd(kp.counters + 1, (unsigned long long)adv);
        atomicAdd(kp.counters + 2, (unsigned long long)calls);
        atomicAdd(kp.counters + 3, (unsigned long long)fast);
        // (fizz_spec_stats)
        atomicAdd(kp.counters + 4, (unsigned long long)adv);
        atomicAdd(kp.counters + 5, (unsigned long long)evaluated);
        atomicAdd(kp.counters + 6, 1ULL);
        if (m < __popc(ranb)) atomicAdd(kp.counters + 7, 1ULL);
    }

    // ---- predictor for the next round: which move explains the last accepted slot, coordinate by coordinate? -------
    int kstar;
    double dtk;
    halving_floor(kp, lane, kstar, dtk);
    const long long n = __ldcg(lmeta + 1);
    int cls = 0;
    if (lane < 2 * F) {
        const int b = lane >> 1, c = lane & 1;
        const double p0 = __ldcg(last + 4 * b + c), v0 = __ldcg(last + 4 * b + 2 + c);
        const double p1 = __ldcg(last + 4 * F + 4 * b + c), v1 = __ldcg(last + 4 * F + 4 * b + 2 + c);
        const double acc = c ? kp.accy : kp.accx;
        const double rdt = kp.time_disc - dtk;
        cls = 3;
        if (n >= 1 && n <= 64) {
            for (int cand = 2; cand >= 0; cand--) {
                if (cand == 2 && kstar < 1) continue;
                double p = p0, v = v0;
                for (long long t = 0; t < n; t++) spec_predict(p, v, cand, kp.time_disc, rdt, acc);
                if (__double_as_longlong(p) == __double_as_longlong(p1) && __double_as_longlong(v) == __double_as_longlong(v1))
                    cls = cand;
            }
        }
    }
    const bool explained = last_clean && !__any_sync(FULL, cls == 3);
    unsigned long long w1 = 0, w2 = 0;
    for (int i = 0; i < 16; i++) {
        const unsigned long long c = (unsigned long long)(__shfl_sync(FULL, cls, i) & 255);
        if (i < 8) w1 |= c << (8 * i); else w2 |= c << (8 * (i - 8));
    }
    int use_i = 0;
    long long a16 = __ldcg(ctrl + 10);
    {
        // how far predictions hold (spec_shape): halfway to what this round got when a slot was rejected, growing by half
        // when every slot was accepted
        const long long got = 16 * adv;
        if (a16 <= 0) a16 = 16LL * sp.S;
        if (m < __popc(ranb)) a16 = (a16 + got) / 2;
        else if (__popc(ranb) > 1) { a16 = (a16 > got ? a16 : got) * 3 / 2; if (a16 > 16LL * 16 * SPEC_K) a16 = 16LL * 16 * SPEC_K; }
    }
    if (lane == 0) {
        ctrl[10] = a16;
        // A slot no move explains is usually an EVENT (the trapped body reached the other wall of its cage: one
        // coordinate jumps, one velocity flips) after which the same moves apply again: keep the predictor for one more
        // round.  Two such rounds in a row -- or an unclean slot -- and the world advances without one (slot 0 alone, a
        // bounded piece per round) until a slot is explained again.
        const bool had = __ldcg(ctrl) > 1;
        const long long streak = explained ? 0 : __ldcg(ctrl + 3) + 1;
        bool use = explained && !sp.nopred;
        if (!explained && last_clean && had && streak < 2 && !sp.nopred) use = true;        // (old classes stay in ctrl[1], ctrl[2])
        else { ctrl[1] = (long long)w1; ctrl[2] = (long long)w2; }
        ctrl[0] = use ? 2 : 1;
        ctrl[3] = streak;
        use_i = use ? 1 : 0;
        // no predictor for 8 rounds in a row: the world goes back to the ordinary launch (one warp, no round boundaries),
        // for good -- bit 56 of the tier word
        if (streak >= 8) { kp.tier[w] = new_tier | (1LL << 56); use_i = 2; }
        unsigned long long *st = reinterpret_cast<unsigned long long *>(ctrl);
        atomicAdd(st + 5, (unsigned long long)adv); atomicAdd(st + 6, (unsigned long long)evaluated);
        if (m < __popc(ranb)) atomicAdd(st + 7, 1ULL);
        if (!had) atomicAdd(st + 8, 1ULL);
        atomicAdd(st + 9, 1ULL);
    }
    use_i = __shfl_sync(FULL, use_i, 0);
    // the next round: none at the chunk target, for a stopped world, or for a world that just turned out unpredictable
    // (the launch after the rounds takes it to the chunk target; from the next chunk on a runner launch steps it)
    if (old_steps + adv < kp.target_steps && new_status == 0 && use_i != 2) {
        int Kw, Lw;
        spec_shape(new_tier, a16, sp.S, sp.D, Kw, Lw);
        knext = use_i ? Kw : 1;
    }
}

}  // namespace fizz
