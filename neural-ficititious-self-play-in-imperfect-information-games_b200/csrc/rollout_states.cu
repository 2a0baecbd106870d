// K2, variant 6: the fused rollout over a table of the nets' outputs on every decision state.
//
// Under main.train's turn order (the precondition of the factorised first layer, rollout_tables.cuh) the observation a
// net sees at a decision is one of 702 values: the actor's card, the dealer and the betting sequence so far in round 0
// (54), or card, public card, dealer, the finished round-0 sequence and the round-1 sequence so far (648).  A net is a
// pure function of the observation, so its three outputs on all 702 of them -- 4 nets x 702 forwards, built by
// states_pack_kernel from the SAME table image and with the SAME arithmetic (mlp_forward_tables, rotation 0) every time
// the weights change -- stand for the 8.4 M forwards of a 2^20-game x 8-step launch.  A decision then costs ONE 16-byte
// shared-memory read instead of variant 1's 80 (32 row quads + 48 W2 quads, which bound that kernel: an LDS.128 holds
// the pipe for four cycles), and what is left is the game logic, Philox and the record append.
//
// Everything around the forward -- fast_begin / fast_finish, block ownership, the staged or direct append, the
// counters -- is variant 1's, so the records, game words and counters are those of variant 1 run on the same score
// vectors; the score vectors differ from variant 1's only by the summation order of a lane's rotation (1e-5 parity
// against the oracle, tests/test_gpu_act.py, tests/test_gpu_baseline_sizes.py).
#include "rollout_fast.cuh"
#include "rollout_tables.cuh"
#include "ptx_helpers.cuh"

namespace nfsp {

constexpr int kNetStates = 54 + 72 * 9;          // 702
constexpr int kStateQuads = 4 * kNetStates;      // one float4 {o0, o1, o2, 0} per net and state
constexpr int kStateBytes = kStateQuads * 16;    // 44 928
static_assert(kStateBytes % 16 == 0, "bulk copies move multiples of 16 bytes");

// decision state -> rows of the factorised first layer (the inverse of the index the rollout computes)
__device__ __forceinline__ void state_rows(int s, uint32_t &xrow, uint32_t &yrow) {
    if (s < 54) {  // (card * 2 + dealer) * 9 + sequence id
        const int sg = s % 9, cd = s / 9;
        xrow = (uint32_t)(cd >> 1);
        yrow = 75u + (uint32_t)(cd & 1) * 18u + (uint32_t)sg;
    } else {       // 54 + x * 9 + round-1 sequence id, x = ((card * 3 + pub) * 2 + dealer) * 4 + finished round-0 sequence
        const int sg = (s - 54) % 9, x = (s - 54) / 9;
        xrow = 3u + (uint32_t)x;
        yrow = 75u + (uint32_t)((x >> 2) & 1) * 18u + 9u + (uint32_t)sg;
    }
}

__global__ void states_pack_kernel(const float *__restrict__ tab, float4 *__restrict__ states) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= kStateQuads) return;
    uint32_t xrow, yrow;
    state_rows(e % kNetStates, xrow, yrow);
    float o0, o1, o2;
    mlp_forward_tables(tab, xrow, yrow, (uint32_t)(e / kNetStates), 0u, o0, o1, o2);
    states[e] = make_float4(o0, o1, o2, 0.f);
}

constexpr int kStatesThreads = 1024;

// ---- warp-private record buffers -----------------------------------------------------------------
// Variant 1 claims its slots with four global atomics per warp and step.  That is one round trip per step on every
// warp's critical path, and with the direct ring append every one of the launch's 262 144 warp-steps hits the SAME two
// words (the rings' totals): same-address atomics retire at ~1.4 ns each, 0.37 ms per launch -- invisible beside variant
// 1's 0.37 ms of shared-memory traffic, the bound of this kernel once the forward is a table read.  Here a warp collects
// its records in shared memory (the table leaves 180 KB free) and claims slots for a whole buffer at a time: one atomic
// and one coalesced copy (512 B per warp instruction) per ~100 records.
constexpr int kBufRL = 112, kBufSL = 32;                 // records per player and warp; >= a step's worst case (64, 32)
constexpr int kWarpBufQuads = 2 * kBufRL + 2 * kBufSL;   // 288 uint4 = 4 608 B per warp
constexpr int kStatesSmemBytes = kStateBytes + (kStatesThreads / 32) * kWarpBufQuads * 16;
static_assert(kBufRL >= 64 && kBufSL >= 32, "a buffer must take the records of one step");

struct WarpBuf {
    uint4 *rl0, *rl1, *sl0, *sl1;  // shared memory
    uint32_t n0 = 0u, n1 = 0u, m0 = 0u, m1 = 0u;  // records held (warp-uniform): RL of player 0 / 1, SL of player 0 / 1

    __device__ __forceinline__ void init(uint4 *base) {
        rl0 = base; rl1 = base + kBufRL; sl0 = base + 2 * kBufRL; sl1 = sl0 + kBufSL;
    }
};

// moves n buffered records of one list to its destination: one atomic for the warp, then a coalesced copy
template <bool kRing>
__device__ __forceinline__ void buf_flush(const uint4 *buf, uint32_t &n, uint4 *dst, void *counter, uint32_t cap, uint64_t magic,
                                          FastCounters &c) {
    __syncwarp();
    if (n) {
        const uint32_t lane = threadIdx.x & 31u;
        uint32_t base = 0u;
        if (lane == 0) {
            if (kRing) base = ring_slot(atomicAdd((unsigned long long *)counter, (unsigned long long)n), cap, magic);
            else base = atomicAdd((uint32_t *)counter, n);
        }
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        int drop = 0;
        for (uint32_t k = lane; k < n; k += 32u) {
            const uint32_t off = base + k;
            if (kRing) dst[off < cap ? off : off - cap] = buf[k];  // a launch never laps the ring
            else if (off < cap) dst[off] = buf[k];
            else ++drop;
        }
        c.wide.drop += drop;
        n = 0u;
    }
    __syncwarp();
}

template <bool kDirect>
__device__ __forceinline__ void buf_flush_rl(WarpBuf &B, const RolloutArgs &A, const WarpStage &W, FastCounters &c, bool p0, bool p1) {
    if (p0) buf_flush<kDirect>(B.rl0, B.n0, W.rl0, kDirect ? (void *)A.ring_total[0] : (void *)W.cnt, W.cap_rl, A.ring_magic, c);
    if (p1) buf_flush<kDirect>(B.rl1, B.n1, W.rl1, kDirect ? (void *)A.ring_total[1] : (void *)(W.cnt + W.n_seg), W.cap_rl, A.ring_magic, c);
}
__device__ __forceinline__ void buf_flush_sl(WarpBuf &B, const WarpStage &W, FastCounters &c, bool p0, bool p1) {
    if (p0) buf_flush<false>(B.sl0, B.m0, W.sl0, W.cnt + 2u * W.n_seg, W.cap_sl, 0ull, c);
    if (p1) buf_flush<false>(B.sl1, B.m1, W.sl1, W.cnt + 3u * W.n_seg, W.cap_sl, 0ull, c);
}

// a step's records into the warp's buffers (what warp_append does with global atomics).  All lanes call it.
template <bool kDirect>
__device__ __forceinline__ void buf_append(WarpBuf &B, const RolloutArgs &A, const WarpStage &W, const FastDecision &d,
                                           const FastRecords &R, float v0, float v1, float v2, FastCounters &c) {
    const uint32_t q = R.q, sh = q * 8u;
    const uint32_t mine = record_counts(R);
    const uint32_t incl = warp_scan_bytes(mine);
    const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
    const uint32_t t0 = tot & 0xFFu, t1 = (tot >> 8) & 0xFFu, t2 = (tot >> 16) & 0xFFu, t3 = tot >> 24;
    // make room first (warp-uniform branches; rare)
    buf_flush_rl<kDirect>(B, A, W, c, B.n0 + t0 > (uint32_t)kBufRL, B.n1 + t1 > (uint32_t)kBufRL);
    buf_flush_sl(B, W, c, B.m0 + t2 > (uint32_t)kBufSL, B.m1 + t3 > (uint32_t)kBufSL);
    const uint32_t excl = incl - mine;
    uint4 *rp = q ? B.rl1 : B.rl0, *ro = q ? B.rl0 : B.rl1, *sp = q ? B.sl1 : B.sl0;
    uint32_t off = (q ? B.n1 : B.n0) + ((excl >> sh) & 0xFFu);
    if (R.vA) rp[off++] = make_uint4(d.snap_a, d.obs, 0u, d.meta_a);
    if (R.vB) rp[off] = R.recB;
    if (R.vC) ro[(q ? B.n0 : B.n1) + ((excl >> (8u - sh)) & 0xFFu)] = R.recC;
    if (R.vS) sp[(q ? B.m1 : B.m0) + ((excl >> (16u + sh)) & 0xFFu)] = make_uint4(d.obs, __float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2));
    B.n0 += t0; B.n1 += t1; B.m0 += t2; B.m1 += t3;
}

template <bool kDebug, bool kDirect>
__global__ void __launch_bounds__(kStatesThreads, 1)
rollout_states_kernel(const RolloutArgs A) {
    extern __shared__ __align__(128) float4 s_tab[];  // kStateQuads, then the warps' record buffers
    __shared__ unsigned long long s_stats[NFSP_STATS_FIELDS];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ FastLuts s_lut;
    __shared__ uint32_t s_next;  // blocks of games this CTA has handed to its warps
    if (threadIdx.x < NFSP_STATS_FIELDS) s_stats[threadIdx.x] = 0ull;
    s_lut.fill();
    const uint32_t bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        s_next = 0u;
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // the state table comes in as two bulk async copies, completion on an mbarrier
        mbar_expect_tx(bar, kStateBytes);
        constexpr uint32_t kChunk = 32768;
        for (uint32_t off = 0; off < (uint32_t)kStateBytes; off += kChunk) {
            const uint32_t len = (uint32_t)kStateBytes - off < kChunk ? (uint32_t)kStateBytes - off : kChunk;
            bulk_g2s(smem_u32(s_tab) + off, reinterpret_cast<const uint8_t *>(A.pack) + off, len, bar);
        }
    }
    bool image_ready = false;

    FastCounters c;
    const int64_t plane = (int64_t)A.n_steps * A.n;
    const uint32_t lane = threadIdx.x & 31u;
    WarpBuf B;
    B.init(reinterpret_cast<uint4 *>(s_tab) + kStateQuads + (threadIdx.x >> 5) * kWarpBufQuads);
    WarpStage W;
    W.init(A, 0u, kDirect);
    const int64_t n_blocks = (A.n + 31) >> 5;  // CTA c owns blocks c, c + grid, ...; its warps take them dynamically
    for (;;) {
        uint32_t blk = 0;
        if (lane == 0) blk = blockIdx.x + gridDim.x * atomicAdd(&s_next, 1u);
        blk = __shfl_sync(0xFFFFFFFFu, blk, 0);
        if ((int64_t)blk >= n_blocks) break;
        const int64_t base = (int64_t)blk << 5;
        const int64_t i = base + lane;
        const bool live = i < A.n;
        const uint64_t game = A.game0 + (uint64_t)i;
        W.init(A, (uint32_t)(base >> 5) & (A.n_seg - 1u), kDirect);
        NfspFast g;
        g.unpack(live ? A.state[i] : 0ull);
        for (int t = 0; t < A.n_steps; ++t) {
            FastDecision d;
            fast_begin(g, s_lut, A, game, A.step0 + (uint64_t)t, live, d, c);
            const uint32_t sg = g.sigma(), dl = g.dealer(), ca = (g.PA >> 11) & 3u;
            const uint32_t s = sg >= 9u ? 54u - 9u + (((((ca * 3u + g.pub()) * 2u + dl) << 2) | g.fin0()) * 9u + sg)
                                        : (ca * 2u + dl) * 9u + sg;
            if (!image_ready) {
                mbar_wait(bar, 0);
                image_ready = true;
            }
            const float4 o = s_tab[(g.p() * 2u + (uint32_t)d.pol) * (uint32_t)kNetStates + s];
            float v0 = o.x, v1 = o.y, v2 = o.z;
            if (d.random) { v0 = d.r0; v1 = d.r1; v2 = d.r2; }
            FastRecords R;
            fast_decide<kDebug>(g, s_lut, A, d, v0, v1, v2, live, (int64_t)t * A.n + i, plane, c, R);
            buf_append<kDirect>(B, A, W, d, R, v0, v1, v2, c);
            if ((t & 15) == 15) c.spill();
        }
        c.spill();
        if (live) A.state[i] = g.pack();
        c.wide.trans += live ? A.n_steps : 0;
        // a block's staged records belong to the block's segment; the rings have no segments, their records stay buffered
        if (!kDirect) buf_flush_rl<false>(B, A, W, c, true, true);
        buf_flush_sl(B, W, c, true, true);
    }
    if (kDirect) buf_flush_rl<true>(B, A, W, c, true, true);
    if (!image_ready) mbar_wait(bar, 0);  // the copy into this CTA's shared memory must land before the CTA exits
    c.spill();
    if (A.stats) c.wide.commit(s_stats, A.stats);
}

}  // namespace nfsp

using namespace nfsp;

int nfsp_states_floats() { return kStateQuads * 4; }

int nfsp_rollout_states_configure() {
    NFSP_CUDA(cudaFuncSetAttribute(rollout_states_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStatesSmemBytes));
    NFSP_CUDA(cudaFuncSetAttribute(rollout_states_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStatesSmemBytes));
    NFSP_CUDA(cudaFuncSetAttribute(rollout_states_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStatesSmemBytes));
    NFSP_CUDA(cudaFuncSetAttribute(rollout_states_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStatesSmemBytes));
    return NFSP_OK;
}

// rebuilds the state table from the table image (both live in d_wpack) when the weights have changed since
int nfsp_states_ensure(nfsp_env_t h, const float *d_tab, float *d_states, cudaStream_t st) {
    if (!h->st_dirty) return NFSP_OK;
    states_pack_kernel<<<(kStateQuads + 127) / 128, 128, 0, st>>>(d_tab, reinterpret_cast<float4 *>(d_states));
    NFSP_LAUNCH_CHECK();
    h->st_dirty = false;
    return NFSP_OK;
}

int nfsp_rollout_states_launch(nfsp_env_t h, const RolloutArgs &A, const nfsp_rollout_io *io, bool debug, cudaStream_t st) {
    const bool direct = A.ring[0] != nullptr;
    const int grid = grid_for(h->n, 32, h->sm_count - io->reserve_sms, 1);  // at least one block of 32 games per CTA
    if (debug && direct) rollout_states_kernel<true, true><<<grid, kStatesThreads, kStatesSmemBytes, st>>>(A);
    else if (debug) rollout_states_kernel<true, false><<<grid, kStatesThreads, kStatesSmemBytes, st>>>(A);
    else if (direct) rollout_states_kernel<false, true><<<grid, kStatesThreads, kStatesSmemBytes, st>>>(A);
    else rollout_states_kernel<false, false><<<grid, kStatesThreads, kStatesSmemBytes, st>>>(A);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}
