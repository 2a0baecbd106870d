// Decision logic shared by the two variants of the fused rollout kernel (CUDA-core first layer in
// act_kernels.cu, tcgen05 first layer in act_tc_kernels.cu): everything of Agent.play (agent.py:130-156)
// and of main.train's inner loop (main.py:28-67) except the network forward itself.
#pragma once
#include "common.cuh"
#include "nfsp_rules.cuh"
#include "philox.cuh"

namespace nfsp { struct RolloutArgs; }
// launches the tcgen05 variant of the fused rollout (act_tc_kernels.cu)
int nfsp_rollout_tc_launch(nfsp_env_t h, const nfsp::RolloutArgs &A, bool debug, cudaStream_t st);

namespace nfsp {

// ---- warp-aggregated record append -------------------------------------------------------------
// The staging arrays are split into n_seg segments (a power of two), each with its own ticket counter; a
// warp's 32 consecutive games always append to segment (first game / 32) % n_seg, so which records land
// in which segment does not depend on the launch geometry, and counters are not shared across the chip.
// Per step a warp claims its slots in the four arrays (RL and SL of both players) with FOUR atomics issued
// by four lanes at once: one global round trip.
struct RolloutArgs {
    uint64_t *state;
    int64_t n;
    uint64_t seed, game0, step0;
    PhiloxKeys keys;         // round keys of `seed` (constant-bank operands of the lean kernels)
    int n_steps;
    uint32_t eta_u32, eps_u32;
    const void *pack;  // weight image of the variant
    uint4 *rl[2];
    uint4 *sl[2];
    int64_t cap_rl, cap_sl;  // per segment
    uint32_t n_seg;          // power of two
    uint32_t *counts;        // [4][n_seg]: rl0, rl1, sl0, sl1
    uint32_t *work;          // next block of 32 games to hand out (zeroed before the launch)
    unsigned long long *stats;
    uint32_t *trace;
    float *vec;
    const float *forced;
};

__device__ __forceinline__ uint4 make_rl(uint32_t s, uint32_t s2, int r_half, uint32_t a, uint32_t t, uint32_t p) {
    return make_uint4(s, s2, __float_as_uint(0.5f * (float)r_half), a | (t << 8) | (p << 16));
}

// per-thread counters; the action histogram packs 3 x 21-bit fields per player (flushed before overflow)
struct Counters {
    unsigned long long act0 = 0ull, act1 = 0ull;
    int rew0 = 0, rew1 = 0, hands = 0, trans = 0, drop = 0;

    __device__ __forceinline__ void flush_hist(unsigned long long *s_stats) {
        atomicAdd(&s_stats[0], act0 & 0x1FFFFFull); atomicAdd(&s_stats[1], (act0 >> 21) & 0x1FFFFFull);
        atomicAdd(&s_stats[2], act0 >> 42);
        atomicAdd(&s_stats[3], act1 & 0x1FFFFFull); atomicAdd(&s_stats[4], (act1 >> 21) & 0x1FFFFFull);
        atomicAdd(&s_stats[5], act1 >> 42);
        act0 = act1 = 0ull;
    }
    // warp shuffle reduction, one shared atomic per warp and counter, then 13 global atomics per CTA;
    // call from every thread of the CTA
    __device__ __forceinline__ void commit(unsigned long long *s_stats, unsigned long long *g_stats) {
        long long v[11] = {(long long)(act0 & 0x1FFFFFull), (long long)((act0 >> 21) & 0x1FFFFFull), (long long)(act0 >> 42),
                           (long long)(act1 & 0x1FFFFFull), (long long)((act1 >> 21) & 0x1FFFFFull), (long long)(act1 >> 42),
                           rew0, rew1, hands, trans, drop};
#pragma unroll
        for (int k = 0; k < 11; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xFFFFFFFFu, v[k], o);
        }
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int k = 0; k < 11; ++k)
                if (v[k]) atomicAdd(&s_stats[k < 6 ? k : k + 2], (unsigned long long)v[k]);
        }
        act0 = act1 = 0ull;
        __syncthreads();
        if (threadIdx.x == 6) s_stats[6] = s_stats[0] + s_stats[1] + s_stats[2];  // played = sum of the histogram
        if (threadIdx.x == 7) s_stats[7] = s_stats[3] + s_stats[4] + s_stats[5];
        __syncthreads();
        if (threadIdx.x < 13 && s_stats[threadIdx.x]) atomicAdd(g_stats + threadIdx.x, s_stats[threadIdx.x]);
    }
};

// what a game decided before the network runs
struct Decision {
    int p = 0;            // acting player
    uint32_t obs = 0;     // its observation mask
    uint32_t pol = 0;     // 1 = best-response policy this hand
    bool started = false; // a new hand was dealt at this step
    bool random = false;  // epsilon branch: the score vector is already in v0..v2
    bool vA = false;      // previous transition of p to remember
    uint4 recA;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
};

// agent.py:130-141 + main.py:28-45: re-deal if the hand is over, pick the actor, remember its previous
// transition, draw the epsilon test / random score vector
__device__ __forceinline__ void decide_begin(NfspW &g, const RolloutArgs &A, uint64_t game, uint64_t step, Decision &d,
                                             Counters &c) {
    const Philox4 x = game_block(A.seed, game, step, STREAM_STEP);
    if (g.need_reset()) {
        g.reset(g.dealer() ^ 1u, deal_ranks(__umulhi(x.y, 120u)), x.z < A.eta_u32, x.w < A.eta_u32);
        d.started = true;
        ++c.hands;
    }
    d.p = g.to_act();
    d.obs = g.obs(d.p);
    if (g.acted_nz(d.p)) {  // agent.py:132-136: remember the previous transition
        d.vA = true;
        d.recA = make_rl(g.snapshot(d.p), d.obs, 0, g.last_a(d.p), 0u, (uint32_t)d.p);
    }
    d.pol = g.policy(d.p);
    if (d.pol && x.x < A.eps_u32) {  // agent.py:125-128: np.random.rand(1,1,3); rare, so its own Philox block
        const Philox4 y = game_block(A.seed, game, step, STREAM_VECTOR);
        d.random = true;
        d.v0 = (float)(y.x >> 8) * (1.0f / 16777216.0f);
        d.v1 = (float)(y.y >> 8) * (1.0f / 16777216.0f);
        d.v2 = (float)(y.z >> 8) * (1.0f / 16777216.0f);
    }
}

// agent.py:142-156 after the forward + main.py:55-67 terminal observations; then the warp-aggregated append.
// Must be called by ALL lanes of a warp (live or not): it contains warp collectives.
template <bool kDebug>
__device__ __forceinline__ void decide_finish(NfspW &g, const RolloutArgs &A, const Decision &d, float v0, float v1,
                                              float v2, bool live, int64_t at, int64_t plane, uint32_t seg,
                                              Counters &c, unsigned long long *s_stats) {
    uint4 recB, recC, recS;
    bool vB = false, vC = false, vS = false;
    const int p = d.p;
    if (live) {
        if (kDebug) {
            if (A.vec) { A.vec[3 * at] = v0; A.vec[3 * at + 1] = v1; A.vec[3 * at + 2] = v2; }
            if (A.forced) { v0 = A.forced[3 * at]; v1 = A.forced[3 * at + 1]; v2 = A.forced[3 * at + 2]; }
        }
        if (d.pol) {  // agent.py:151: the raw score vector goes to the SL memory
            vS = true;
            recS = make_uint4(d.obs, __float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2));
        }
        int a = 0;  // np.argmax: first maximum
        float best = v0;
        if (v1 > best) { a = 1; best = v1; }
        if (v2 > best) a = 2;
        const bool nz = (v0 != 0.f) || (v1 != 0.f) || (v2 != 0.f);
        const int eff = g.step(a, nz, p);
        const unsigned long long inc = 1ull << (21 * a);
        c.act0 += p == 0 ? inc : 0ull;
        c.act1 += p == 1 ? inc : 0ull;
        if ((++c.trans & 0xFFFFF) == 0) c.flush_hist(s_stats);
        if (g.terminated()) {  // main.py:55-67: both players observe the terminal state once
            const int o = p ^ 1;
            c.rew0 += g.reward_half(0);
            c.rew1 += g.reward_half(1);
            if (g.acted_nz(p)) {
                vB = true;
                recB = make_rl(g.snapshot(p), g.obs(p), g.reward_half(p), g.last_a(p), 1u, (uint32_t)p);
            }
            if (g.acted_nz(o)) {
                vC = true;
                recC = make_rl(g.snapshot(o), g.obs(o), g.reward_half(o), g.last_a(o), 1u, (uint32_t)o);
            }
            g.w |= 1ull << 43;
        }
        if (kDebug && A.trace) {
            A.trace[at] = g.obs(p) | ((uint32_t)g.terminated() << 30) | ((uint32_t)p << 31);
            A.trace[plane + at] = __float_as_uint(0.5f * (float)g.reward_half(p));
            A.trace[2 * plane + at] = g.trace_misc(a, eff, d.started);
        }
    }
    // ---- append: counts per destination array k = rl0, rl1, sl0, sl1
    const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const int n_rl_p = (int)d.vA + (int)vB;  // records for the actor's ring (0..2)
    const uint32_t a1_0 = __ballot_sync(0xFFFFFFFFu, live && p == 0 && n_rl_p >= 1);
    const uint32_t a2_0 = __ballot_sync(0xFFFFFFFFu, live && p == 0 && n_rl_p >= 2);
    const uint32_t a1_1 = __ballot_sync(0xFFFFFFFFu, live && p == 1 && n_rl_p >= 1);
    const uint32_t a2_1 = __ballot_sync(0xFFFFFFFFu, live && p == 1 && n_rl_p >= 2);
    const uint32_t c_0 = __ballot_sync(0xFFFFFFFFu, live && vC && p == 1);  // opponent's terminal record -> ring 0
    const uint32_t c_1 = __ballot_sync(0xFFFFFFFFu, live && vC && p == 0);
    const uint32_t s_0 = __ballot_sync(0xFFFFFFFFu, live && vS && p == 0);
    const uint32_t s_1 = __ballot_sync(0xFFFFFFFFu, live && vS && p == 1);
    const uint32_t tot0 = __popc(a1_0) + __popc(a2_0) + __popc(c_0), tot1 = __popc(a1_1) + __popc(a2_1) + __popc(c_1);
    const uint32_t tot2 = __popc(s_0), tot3 = __popc(s_1);
    uint32_t base = 0;
    if (lane < 4) {
        const uint32_t t = lane == 0 ? tot0 : (lane == 1 ? tot1 : (lane == 2 ? tot2 : tot3));
        if (t) base = atomicAdd(A.counts + lane * A.n_seg + seg, t);
    }
    const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, base, 0), b1 = __shfl_sync(0xFFFFFFFFu, base, 1);
    const uint32_t b2 = __shfl_sync(0xFFFFFFFFu, base, 2), b3 = __shfl_sync(0xFFFFFFFFu, base, 3);
    if (live) {
        // slot order inside a warp's claim for ring q: [actor lanes' records in lane order][opponent records]
        const uint32_t a1 = p == 0 ? a1_0 : a1_1, a2 = p == 0 ? a2_0 : a2_1;
        const uint32_t bp = p == 0 ? b0 : b1, bo = p == 0 ? b1 : b0;
        uint4 *rp = A.rl[p] + (size_t)seg * A.cap_rl, *ro = A.rl[p ^ 1] + (size_t)seg * A.cap_rl;
        uint32_t off = bp + __popc(a1 & lt) + __popc(a2 & lt);
        if (d.vA) { if ((int64_t)off < A.cap_rl) rp[off] = d.recA; else ++c.drop; ++off; }
        if (vB) { if ((int64_t)off < A.cap_rl) rp[off] = recB; else ++c.drop; }
        if (vC) {
            const uint32_t oa1 = p == 0 ? a1_1 : a1_0, oa2 = p == 0 ? a2_1 : a2_0, oc = p == 0 ? c_1 : c_0;
            const uint32_t o2 = bo + __popc(oa1) + __popc(oa2) + __popc(oc & lt);
            if ((int64_t)o2 < A.cap_rl) ro[o2] = recC; else ++c.drop;
        }
        if (vS) {
            const uint32_t o3 = (p == 0 ? b2 : b3) + __popc((p == 0 ? s_0 : s_1) & lt);
            if ((int64_t)o3 < A.cap_sl) (A.sl[p] + (size_t)seg * A.cap_sl)[o3] = recS; else ++c.drop;
        }
    }
}

}  // namespace nfsp
