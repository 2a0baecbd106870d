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
// Every lane contributes cnt in {0,1,2} records for one destination array; one atomicAdd per warp
// claims the tickets, lanes write their 16-byte records at consecutive slots.
__device__ __forceinline__ uint32_t warp_claim(uint32_t *counter, int cnt, uint32_t &my_off) {
    const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const uint32_t m1 = __ballot_sync(0xFFFFFFFFu, cnt >= 1), m2 = __ballot_sync(0xFFFFFFFFu, cnt >= 2);
    const uint32_t total = __popc(m1) + __popc(m2);
    my_off = __popc(m1 & lt) + __popc(m2 & lt);
    uint32_t base = 0;
    if (total) {
        if (lane == 0) base = atomicAdd(counter, total);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
    }
    return base;
}

struct RolloutArgs {
    uint64_t *state;
    int64_t n;
    uint64_t seed, game0, step0;
    int n_steps;
    uint32_t eta_u32, eps_u32;
    const void *pack;  // weight image of the variant
    uint4 *rl[2];
    uint4 *sl[2];
    int64_t cap_rl, cap_sl;
    uint32_t *counts;
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
    // block-level reduction into s_stats, then 13 global atomics per CTA; call from every thread of the CTA
    __device__ __forceinline__ void commit(unsigned long long *s_stats, unsigned long long *g_stats) {
        flush_hist(s_stats);
        atomicAdd(&s_stats[8], (unsigned long long)(long long)rew0);
        atomicAdd(&s_stats[9], (unsigned long long)(long long)rew1);
        atomicAdd(&s_stats[10], (unsigned long long)hands);
        atomicAdd(&s_stats[11], (unsigned long long)trans);
        atomicAdd(&s_stats[12], (unsigned long long)drop);
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
    if (g.need_reset()) {
        const Philox4 y = game_block(A.seed, game, step, STREAM_RESET);
        g.reset(g.dealer() ^ 1u, deal_ranks(__umulhi(y.x, 120u)), y.y < A.eta_u32, y.z < A.eta_u32);
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
    const Philox4 x = game_block(A.seed, game, step, STREAM_STEP);
    if (d.pol && x.x < A.eps_u32) {  // agent.py:125-128: np.random.rand(1,1,3)
        d.random = true;
        d.v0 = (float)(x.y >> 8) * (1.0f / 16777216.0f);
        d.v1 = (float)(x.z >> 8) * (1.0f / 16777216.0f);
        d.v2 = (float)(x.w >> 8) * (1.0f / 16777216.0f);
    }
}

// agent.py:142-156 after the forward + main.py:55-67 terminal observations; then the warp-aggregated append.
// Must be called by ALL lanes of a warp (live or not): it contains warp collectives.
template <bool kDebug>
__device__ __forceinline__ void decide_finish(NfspW &g, const RolloutArgs &A, const Decision &d, float v0, float v1,
                                              float v2, bool live, int64_t at, int64_t plane, Counters &c,
                                              unsigned long long *s_stats) {
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
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const bool mine = live && (p == q);
        const int cnt = mine ? ((int)d.vA + (int)vB) : (int)(live && vC);
        uint32_t off;
        const uint32_t b = warp_claim(A.counts + q, cnt, off);
        uint4 *dst = A.rl[q];
        if (mine) {
            if (d.vA) { if ((int64_t)(b + off) < A.cap_rl) dst[b + off] = d.recA; else ++c.drop; ++off; }
            if (vB) { if ((int64_t)(b + off) < A.cap_rl) dst[b + off] = recB; else ++c.drop; }
        } else if (live && vC) {
            if ((int64_t)(b + off) < A.cap_rl) dst[b + off] = recC; else ++c.drop;
        }
        const int cs = (mine && vS) ? 1 : 0;
        const uint32_t bs = warp_claim(A.counts + 2 + q, cs, off);
        if (cs) { if ((int64_t)(bs + off) < A.cap_sl) A.sl[q][bs + off] = recS; else ++c.drop; }
    }
}

}  // namespace nfsp
