// K2, variant 5: the CUDA-core fused rollout with TWO games per lane and the second layer's weights loaded once for two
// decisions.
//
// rollout_kernel (variant 1) is bound by shared memory -> register bandwidth: an LDS.128 holds the pipe for four cycles
// whatever its lanes read, and a decision needs 32 of them for its first-layer rows and 48 for its net's W2
// (profiles/r02/sorted_rollout_notes.txt).  The only lever is fewer loads per decision: W2 fetched once for TWO decisions of
// the same net evaluated by one lane (32 + 48/2 = 56).  Variant 4 showed that sorting across warps costs more than it
// saves (barriers, lock-step phases); here everything stays inside a warp:
//   * a warp holds 64 games, two per lane (512-thread CTAs, 128 registers per thread), so a warp-step has 64 decisions;
//   * they are counting-sorted by net with four ballots; inside a bin consecutive decisions form PAIRS, pair j is evaluated by
//     lane j: two row pairs, one set of W2 loads (mlp_forward_tables_2);
//   * a bin with an odd count leaves one decision over (0, 2 or 4 per warp-step): those are evaluated by the whole warp
//     together, lane l holding hidden units 2l and 2l+1 (five LDS.64 and fifteen shuffles per decision instead of a full
//     80-load forward that would cost the same whatever its active lanes);
//   * descriptors and scores travel through 1.4 KB of shared memory per warp, ordered by __syncwarp only.
// Same table image, same game logic (rollout_fast.cuh) and the same per-warp append as variant 1, twice per warp-step.
#include "ptx_helpers.cuh"
#include "rollout_fast.cuh"
#include "rollout_tables.cuh"

namespace nfsp {

constexpr int kPairThreads = 512, kPairWarps = kPairThreads / 32;
constexpr int kPairImagePad = (kTabImageBytes + 127) / 128 * 128;

struct __align__(16) PairScratch {  // per warp
    float4 res[64];     // scores of the paired decisions, [2 * pair + member]
    float4 res_left[4];  // scores of the left-over decision of each bin
    uint32_t desc[64];  // paired decisions: xrow | (yrow - 75) << 7 | net << 13
    uint32_t left[4];   // left-over decision of each bin
};
constexpr int kPairSmemBytes = kPairImagePad + kPairWarps * (int)sizeof(PairScratch);

// two decisions of the same net: the W2 quads are loaded once
__device__ __forceinline__ void mlp_forward_tables_2(const float *__restrict__ st, uint32_t e0, uint32_t e1, uint32_t rot, float4 &o0,
                                                     float4 &o1) {
    const float4 *T = reinterpret_cast<const float4 *>(st);
    const uint32_t net = e0 >> 13;
    const float4 *x0 = T + (net * kNetRows + (e0 & 127u)) * kRowQuads + rot, *y0 = T + (net * kNetRows + 75u + ((e0 >> 7) & 63u)) * kRowQuads + rot;
    const float4 *x1 = T + (net * kNetRows + (e1 & 127u)) * kRowQuads + rot, *y1 = T + (net * kNetRows + 75u + ((e1 >> 7) & 63u)) * kRowQuads + rot;
    const float4 *wr = T + kTabRows * kRowQuads + net * (3 * kRowQuads) + rot;
    Layer2Acc a0, a1;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const float4 u0 = wr[q], u1 = wr[kRowQuads + q], u2 = wr[2 * kRowQuads + q];
        a0.quad_sum(x0[q], y0[q], u0, u1, u2);
        a1.quad_sum(x1[q], y1[q], u0, u1, u2);
    }
    const float4 b2 = reinterpret_cast<const float4 *>(st + kTabFloats + kTabW2Floats)[net];
    a0.head(b2, net & 1u, o0.x, o0.y, o0.z);
    a1.head(b2, net & 1u, o1.x, o1.y, o1.z);
    o0.w = o1.w = 0.f;
}

// one decision evaluated by the whole warp: lane l holds hidden units 2l and 2l + 1 (the first 16 quads of a row and of a
// W2 column are in natural order); every lane returns the scores
__device__ __forceinline__ float4 mlp_forward_tables_coop(const float *__restrict__ st, uint32_t e, uint32_t lane) {
    const uint32_t net = e >> 13;
    const float2 *x = reinterpret_cast<const float2 *>(st + ((net * kNetRows + (e & 127u)) * kRowQuads) * 4) + lane;
    const float2 *y = reinterpret_cast<const float2 *>(st + ((net * kNetRows + 75u + ((e >> 7) & 63u)) * kRowQuads) * 4) + lane;
    const float2 *w = reinterpret_cast<const float2 *>(st + kTabFloats + net * (3 * kRowQuads * 4)) + lane;
    const float2 xv = *x, yv = *y;
    const float h0 = fmaxf(xv.x + yv.x, 0.f), h1 = fmaxf(xv.y + yv.y, 0.f);
    float s[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float2 wv = w[c * (kRowQuads * 2)];
        s[c] = h0 * wv.x + h1 * wv.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) s[c] += __shfl_xor_sync(0xFFFFFFFFu, s[c], o);
    }
    float4 out;
    Layer2Acc::head_of_sums(s[0], s[1], s[2], reinterpret_cast<const float4 *>(st + kTabFloats + kTabW2Floats)[net], net & 1u, out.x, out.y,
                            out.z);
    out.w = 0.f;
    return out;
}

template <bool kDebug, bool kDirect>
__global__ void __launch_bounds__(kPairThreads, 1)
rollout_pairs_kernel(const RolloutArgs A) {
    extern __shared__ __align__(128) float sw[];  // table image, then the warps' exchange areas
    __shared__ unsigned long long s_stats[NFSP_STATS_FIELDS];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ FastLuts s_lut;
    __shared__ uint32_t s_next;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    PairScratch &S = reinterpret_cast<PairScratch *>(reinterpret_cast<uint8_t *>(sw) + kPairImagePad)[warp];
    if (threadIdx.x < NFSP_STATS_FIELDS) s_stats[threadIdx.x] = 0ull;
    s_lut.fill();
    const uint32_t bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        s_next = 0u;
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // the table image comes in as bulk async copies (TMA unit), completion on an mbarrier
        mbar_expect_tx(bar, kTabImageBytes);
        constexpr uint32_t kChunk = 32768;
        for (uint32_t off = 0; off < (uint32_t)kTabImageBytes; off += kChunk) {
            const uint32_t len = (uint32_t)kTabImageBytes - off < kChunk ? (uint32_t)kTabImageBytes - off : kChunk;
            bulk_g2s(smem_u32(sw) + off, reinterpret_cast<const uint8_t *>(A.pack) + off, len, bar);
        }
    }
    bool image_ready = false;

    FastCounters c;
    const int64_t plane = (int64_t)A.n_steps * A.n;
    const uint32_t rot = lane & 7u, lt_mask = (1u << lane) - 1u;
    // blocks of 64 consecutive games: CTA c owns blocks c, c + grid, ...; its warps take them through a shared-memory counter
    const int64_t n_blocks = (A.n + 63) >> 6;
    for (;;) {
        uint32_t blk = 0;
        if (lane == 0) blk = blockIdx.x + gridDim.x * atomicAdd(&s_next, 1u);
        blk = __shfl_sync(0xFFFFFFFFu, blk, 0);
        if ((int64_t)blk >= n_blocks) break;
        const int64_t base = (int64_t)blk << 6;
        int64_t i[2];
        bool live[2];
        uint64_t game[2];
        WarpStage W[2];
        NfspFast g[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            i[s] = base + 32 * s + lane;
            live[s] = i[s] < A.n;
            game[s] = A.game0 + (uint64_t)i[s];
            W[s].init(A, (uint32_t)((base >> 5) + s) & (A.n_seg - 1u), kDirect);
            g[s].unpack(live[s] ? A.state[i[s]] : 0ull);
        }
        for (int t = 0; t < A.n_steps; ++t) {
            FastDecision d[2];
            uint32_t key[2], desc[2], b0[2], b1[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                fast_begin(g[s], s_lut, A, game[s], A.step0 + (uint64_t)t, live[s], d[s], c);
                const uint32_t sg = g[s].sigma(), dl = g[s].dealer(), ca = (g[s].PA >> 11) & 3u;
                const uint32_t xrow = sg >= 9u ? 3u + ((((ca * 3u + g[s].pub()) * 2u + dl) << 2) | g[s].fin0()) : ca;
                const uint32_t net = g[s].p() * 2u + (uint32_t)d[s].pol;
                key[s] = net;
                desc[s] = xrow | ((dl * 18u + sg) << 7) | (net << 13);
                // a phantom lane (beyond the last game) takes no part in the sort: what a game's scores are summed with
                // depends on the games of its block only, never on the launch geometry
                b0[s] = __ballot_sync(0xFFFFFFFFu, live[s] && (net & 1u));
                b1[s] = __ballot_sync(0xFFFFFFFFu, live[s] && (net & 2u));
            }
            const uint32_t lv0 = __ballot_sync(0xFFFFFFFFu, live[0]), lv1 = __ballot_sync(0xFFFFFFFFu, live[1]);
            // per bin: members in set 0 / set 1, pairs, first pair, odd one out
            uint32_t m0[4], m1[4], first[4], npair[4], n_pairs = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                m0[k] = ((k & 1) ? b0[0] : ~b0[0]) & ((k & 2) ? b1[0] : ~b1[0]) & lv0;
                m1[k] = ((k & 1) ? b0[1] : ~b0[1]) & ((k & 2) ? b1[1] : ~b1[1]) & lv1;
                const uint32_t cnt = __popc(m0[k]) + __popc(m1[k]);
                first[k] = n_pairs;
                npair[k] = cnt >> 1;
                n_pairs += cnt >> 1;
            }
            uint32_t pos[2];  // where a decision's scores will be: 2 * pair + member, 64 + bin for a left-over one, 255 = none
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                pos[s] = 255u;
                if (live[s]) {
                    uint32_t mk0 = m0[0], mk1 = m1[0], f = first[0], np = npair[0];
#pragma unroll
                    for (int k = 1; k < 4; ++k)
                        if (key[s] == (uint32_t)k) { mk0 = m0[k]; mk1 = m1[k]; f = first[k]; np = npair[k]; }
                    const uint32_t r = s == 0 ? __popc(mk0 & lt_mask) : __popc(mk0) + __popc(mk1 & lt_mask);
                    if (r < 2u * np) {
                        pos[s] = 2u * (f + (r >> 1)) + (r & 1u);
                        S.desc[pos[s]] = desc[s];
                    } else {
                        pos[s] = 64u + key[s];
                        S.left[key[s]] = desc[s];
                    }
                }
            }
            __syncwarp();
            if (!image_ready) {
                mbar_wait(bar, 0);
                image_ready = true;
            }
            if (lane < n_pairs) {
                float4 o0, o1;
                mlp_forward_tables_2(sw, S.desc[2 * lane], S.desc[2 * lane + 1], rot, o0, o1);
                S.res[2 * lane] = o0;
                S.res[2 * lane + 1] = o1;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t cnt = __popc(m0[k]) + __popc(m1[k]);
                if (cnt & 1u) {  // warp-uniform
                    const float4 o = mlp_forward_tables_coop(sw, S.left[k], lane);
                    if (lane == 0) S.res_left[k] = o;
                }
            }
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pos[s] < 64u) r = S.res[pos[s]];
                else if (pos[s] < 68u) r = S.res_left[pos[s] - 64u];
                float v0 = r.x, v1 = r.y, v2 = r.z;
                if (d[s].random) { v0 = d[s].r0; v1 = d[s].r1; v2 = d[s].r2; }
                fast_finish<kDebug, kDirect>(g[s], s_lut, A, W[s], d[s], v0, v1, v2, live[s], (int64_t)t * A.n + i[s], plane, c);
            }
            __syncwarp();  // the scratch area is rewritten in the next step
            if ((t & 7) == 7) c.spill();  // two decisions per step: the 5-bit counters fill twice as fast
        }
        c.spill();
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (live[s]) A.state[i[s]] = g[s].pack();
            c.wide.trans += live[s] ? A.n_steps : 0;
        }
    }
    if (!image_ready) mbar_wait(bar, 0);
    c.spill();
    if (A.stats) c.wide.commit(s_stats, A.stats);
}

}  // namespace nfsp

using namespace nfsp;

int nfsp_rollout_pairs_configure() {
    NFSP_CUDA(cudaFuncSetAttribute(rollout_pairs_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    NFSP_CUDA(cudaFuncSetAttribute(rollout_pairs_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    NFSP_CUDA(cudaFuncSetAttribute(rollout_pairs_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    NFSP_CUDA(cudaFuncSetAttribute(rollout_pairs_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    return NFSP_OK;
}

int nfsp_rollout_pairs_launch(nfsp_env_t h, const RolloutArgs &A, const nfsp_rollout_io *io, bool debug, cudaStream_t st) {
    const bool direct = A.ring[0] != nullptr;
    int grid = h->sm_count - io->reserve_sms;
    const int64_t need = (A.n + 63) >> 6;  // at least one block of 64 games per CTA
    if (need < grid) grid = (int)need;
    if (grid < 1) grid = 1;
    if (debug && direct) rollout_pairs_kernel<true, true><<<grid, kPairThreads, kPairSmemBytes, st>>>(A);
    else if (debug) rollout_pairs_kernel<true, false><<<grid, kPairThreads, kPairSmemBytes, st>>>(A);
    else if (direct) rollout_pairs_kernel<false, true><<<grid, kPairThreads, kPairSmemBytes, st>>>(A);
    else rollout_pairs_kernel<false, false><<<grid, kPairThreads, kPairSmemBytes, st>>>(A);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}
