// Lean decision logic of the fused rollout on the actor-relative register state of nfsp_fast.cuh:
// everything of Agent.play (agent.py:130-156) and of main.train's inner loop (main.py:28-67) around the
// network forward, shared by both first-layer variants.  The rules it applies are those of NfspW::step
// (nfsp_rules.cuh), which the general env kernel still runs on the packed word: two implementations of the same
// game that the parity tests and `tracefile check` hold against each other.
#pragma once
#include "nfsp_fast.cuh"
#include "rollout_common.cuh"

namespace nfsp {

// per-warp destinations of the staged records (the segment of the warp's 32 games is fixed for a launch)
struct WarpStage {
    uint4 *rl0, *rl1, *sl0, *sl1;
    uint32_t *cnt;  // &counts[seg]; the four counters are n_seg apart
    uint32_t n_seg, cap_rl, cap_sl;

    // direct = the RL records go into the players' rings (rl0 / rl1 = the rings, cap_rl = their capacity)
    __device__ __forceinline__ void init(const RolloutArgs &A, uint32_t seg, bool direct = false) {
        rl0 = direct ? A.ring[0] : A.rl[0] + (size_t)seg * A.cap_rl;
        rl1 = direct ? A.ring[1] : A.rl[1] + (size_t)seg * A.cap_rl;
        sl0 = A.sl[0] + (size_t)seg * A.cap_sl;
        sl1 = A.sl[1] + (size_t)seg * A.cap_sl;
        cnt = A.counts + seg;
        n_seg = A.n_seg;
        cap_rl = direct ? A.ring_cap : (uint32_t)A.cap_rl;
        cap_sl = (uint32_t)A.cap_sl;
    }
};

// per-thread counters of the lean kernels; `small` holds six 5-bit action counters (3*player + action) that
// spill into the wide ones every 16 steps
struct FastCounters {
    Counters wide;
    uint32_t small = 0u;

    __device__ __forceinline__ void spill() {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            wide.act0 += (unsigned long long)((small >> (5 * a)) & 31u) << (21 * a);
            wide.act1 += (unsigned long long)((small >> (5 * (3 + a))) & 31u) << (21 * a);
        }
        small = 0u;
    }
};

// What a game decided before the network runs (agent.py:130-141 + main.py:28-45).
struct FastDecision {
    uint32_t obs;      // observation of the player to act
    uint32_t snap_a;   // its previous snapshot (s of the transition it remembers now)
    uint32_t meta_a;   // last action | player << 16 of that transition
    bool vA;           // remember the previous transition (agent.py:132-136)
    bool pol;          // best-response policy this hand
    bool random;       // epsilon branch (agent.py:125-128)
    bool started;
    float r0, r1, r2;  // the random score vector, valid iff random
};

__device__ __forceinline__ void fast_begin(NfspFast &g, const FastLuts &L, const RolloutArgs &A, uint64_t game,
                                           uint64_t step, bool live, FastDecision &d, FastCounters &c) {
    const Philox4 x = game_block(A.keys, game, step, STREAM_STEP);
    d.started = false;
    if (g.need_reset()) {
        const uint32_t idx = __umulhi(x.y, 120u);
        g.redeal(L.deal[idx], L.deal[120u + idx], x.z < A.eta_u32, x.w < A.eta_u32);
        d.started = true;
        c.wide.hands += live;
    }
    d.obs = g.obs_a();
    d.vA = (g.PA & kPNz) != 0u;
    d.snap_a = g.SA;
    d.meta_a = ((g.PA >> 5) & 3u) | (g.p() << 16);
    d.pol = (g.PA & kPPol) != 0u;
    d.random = d.pol && x.x < (g.p() ? A.eps1_u32 : A.eps_u32);
    if (d.random) {  // np.random.rand(1,1,3): rare, so it has its own Philox block
        const Philox4 y = game_block(A.keys, game, step, STREAM_VECTOR);
        d.r0 = (float)(y.x >> 8) * (1.0f / 16777216.0f);
        d.r1 = (float)(y.y >> 8) * (1.0f / 16777216.0f);
        d.r2 = (float)(y.z >> 8) * (1.0f / 16777216.0f);
    }
}

// What a decision leaves behind for the memories (agent.py:134-136,151 + main.py:55-67): up to two RL records for the
// actor q (vA: the transition it remembered before acting, from FastDecision; vB: its terminal observation), one for the
// other player (vC) and one SL record (vS: the best-response policy acted).
struct FastRecords {
    uint32_t q;
    bool vA, vB, vC, vS;
    uint4 recB, recC;
};

// agent.py:142-156 after the forward + main.py:55-67 terminal observations; the records are returned, not appended.
// `live` masks everything a phantom lane could emit.
// np.argmax (first maximum) of a score vector | (np.average != 0, agent.py:134) << 2
__device__ __forceinline__ uint32_t score_action(float v0, float v1, float v2) {
    uint32_t a = 0u;
    float best = v0;
    if (v1 > best) { a = 1u; best = v1; }
    if (v2 > best) a = 2u;
    return a | (uint32_t)((v0 != 0.f) || (v1 != 0.f) || (v2 != 0.f)) << 2;
}

// `pre`: score_action() of (v0, v1, v2) when the caller already has it (the state table stores it beside the scores),
// kNoAction otherwise; a debug launch always recomputes it (the scores may be forced)
constexpr uint32_t kNoAction = 0xFFFFFFFFu;
template <bool kDebug>
__device__ __forceinline__ void fast_decide(NfspFast &g, const FastLuts &L, const RolloutArgs &A, const FastDecision &d,
                                            float &v0, float &v1, float &v2, bool live, int64_t at, int64_t plane,
                                            FastCounters &c, FastRecords &R, uint32_t pre = kNoAction) {
    if (kDebug && live) {
        if (A.vec) { A.vec[3 * at] = v0; A.vec[3 * at + 1] = v1; A.vec[3 * at + 2] = v2; }
        if (A.forced) { v0 = A.forced[3 * at]; v1 = A.forced[3 * at + 1]; v2 = A.forced[3 * at + 2]; }
    }
    const uint32_t q = g.p();
    const uint32_t an = (kDebug || pre == kNoAction) ? score_action(v0, v1, v2) : pre;
    const int a = (int)(an & 3u);
    const bool nz = (an & 4u) != 0u;
    const int eff = (int)(g.step(L.step, a, nz) & 3u);
    c.small += live ? 1u << (5u * (3u * q + (uint32_t)a)) : 0u;
    bool vB = false, vC = false;
    int ra = 0;
    if (g.terminated()) {  // main.py:55-67: both players observe the terminal state once (no swap happened)
        int ro;
        g.rewards(ra, ro);
        c.wide.rew0 += live ? (q ? ro : ra) : 0;
        c.wide.rew1 += live ? (q ? ra : ro) : 0;
        vB = nz;
        R.recB = make_uint4(g.SA, g.obs_a(), __float_as_uint(0.5f * (float)ra), (uint32_t)a | (1u << 8) | (q << 16));
        vC = (g.PO & kPNz) != 0u;
        R.recC = make_uint4(g.SO, g.obs_o(), __float_as_uint(0.5f * (float)ro), ((g.PO >> 5) & 3u) | (1u << 8) | ((q ^ 1u) << 16));
    }
    if (kDebug && live && A.trace) {
        A.trace[at] = (g.terminated() ? g.obs_a() : (q == g.p() ? g.obs_a() : g.obs_o())) | ((uint32_t)g.terminated() << 30) | (q << 31);
        A.trace[plane + at] = __float_as_uint(0.5f * (float)ra);
        A.trace[2 * plane + at] = g.trace_misc(a, eff, d.started);
    }
    R.q = q;
    R.vA = d.vA && live;
    R.vS = d.pol && live;
    R.vB = vB && live;
    R.vC = vC && live;
}

// records of one lane as byte counters rl0 | rl1 << 8 | sl0 << 16 | sl1 << 24, and their inclusive warp scan
__device__ __forceinline__ uint32_t record_counts(const FastRecords &R) {
    const uint32_t sh = R.q * 8u;
    return ((uint32_t)R.vA + (uint32_t)R.vB) << sh | (uint32_t)R.vC << (8u - sh) | (uint32_t)R.vS << (16u + sh);
}
__device__ __forceinline__ uint32_t warp_scan_bytes(uint32_t mine) {
    uint32_t incl = mine;
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        incl += lane >= (uint32_t)o ? up : 0u;
    }
    return incl;
}

// the warp-aggregated append: one warp scan, four atomics per warp -- on the segment's counters, or for the RL records
// of a direct launch on the rings' own totals (ticket -> slot = ticket % capacity; a launch never laps the ring, so one
// conditional subtraction wraps a record's slot).  All lanes call it.
template <bool kDirect = false>
__device__ __forceinline__ void warp_append(const RolloutArgs &A, const WarpStage &W, const FastDecision &d, const FastRecords &R,
                                            float v0, float v1, float v2, FastCounters &c) {
    const uint32_t q = R.q, sh = q * 8u;
    const uint32_t mine = record_counts(R);
    const uint32_t incl = warp_scan_bytes(mine);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
    uint32_t base = 0;
    if (lane < 4) {
        const uint32_t t = (tot >> (8u * lane)) & 0xFFu;
        if (t) {
            if (kDirect && lane < 2) base = ring_slot(atomicAdd(A.ring_total[lane], (unsigned long long)t), A.ring_cap, A.ring_magic);
            else base = atomicAdd(W.cnt + lane * W.n_seg, t);
        }
    }
    const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, base, 0), b1 = __shfl_sync(0xFFFFFFFFu, base, 1);
    const uint32_t b2 = __shfl_sync(0xFFFFFFFFu, base, 2), b3 = __shfl_sync(0xFFFFFFFFu, base, 3);
    const uint32_t excl = incl - mine;
    uint4 *rp = q ? W.rl1 : W.rl0, *ro = q ? W.rl0 : W.rl1, *sp = q ? W.sl1 : W.sl0;
    uint32_t off = (q ? b1 : b0) + ((excl >> sh) & 0xFFu);
    int drop = 0;
    if (R.vA) {
        if (kDirect) rp[off < W.cap_rl ? off : off - W.cap_rl] = make_uint4(d.snap_a, d.obs, 0u, d.meta_a);
        else if (off < W.cap_rl) rp[off] = make_uint4(d.snap_a, d.obs, 0u, d.meta_a);
        else ++drop;
        ++off;
    }
    if (R.vB) {
        if (kDirect) rp[off < W.cap_rl ? off : off - W.cap_rl] = R.recB;
        else if (off < W.cap_rl) rp[off] = R.recB;
        else ++drop;
    }
    if (R.vC) {
        const uint32_t o2 = (q ? b0 : b1) + ((excl >> (8u - sh)) & 0xFFu);
        if (kDirect) ro[o2 < W.cap_rl ? o2 : o2 - W.cap_rl] = R.recC;
        else if (o2 < W.cap_rl) ro[o2] = R.recC;
        else ++drop;
    }
    if (R.vS) {
        const uint32_t o3 = (q ? b3 : b2) + ((excl >> (16u + sh)) & 0xFFu);
        if (o3 < W.cap_sl) sp[o3] = make_uint4(d.obs, __float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2)); else ++drop;
    }
    c.wide.drop += drop;
}

// decide + append, as the warp-per-block kernels use it.  All lanes of the warp call it.
template <bool kDebug, bool kDirect = false>
__device__ __forceinline__ void fast_finish(NfspFast &g, const FastLuts &L, const RolloutArgs &A, const WarpStage &W, const FastDecision &d,
                                            float v0, float v1, float v2, bool live, int64_t at, int64_t plane,
                                            FastCounters &c) {
    FastRecords R;
    fast_decide<kDebug>(g, L, A, d, v0, v1, v2, live, at, plane, c, R);
    warp_append<kDirect>(A, W, d, R, v0, v1, v2, c);
}

}  // namespace nfsp
