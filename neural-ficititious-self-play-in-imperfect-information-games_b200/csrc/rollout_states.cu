// K2, variant 6: the fused rollout over a table of the nets' outputs on every decision state.
//
// Under main.train's turn order (the precondition of the factorised first layer, rollout_tables.cuh) the observation a
// net sees at a decision is one of 702 values: the actor's card, the dealer and the betting sequence so far in round 0
// (54), or card, public card, dealer, the finished round-0 sequence and the round-1 sequence so far (648).  A net is a
// pure function of the observation, so its three outputs on all of them -- computed by states_pack_kernel from the SAME
// table image and with the same operations as variant 1 (two rows added, relu, FFMA2 second layer, head; the 16
// hidden-unit quads summed in four interleaved groups) every time the weights change -- stand for the 8.4 M forwards
// of a 2^20-game x 8-step launch.  A decision then costs ONE 16-byte shared-memory read instead of variant 1's 80 (32
// row quads + 48 W2 quads, which bound that kernel: an LDS.128 holds the pipe for four cycles), and what is left is the
// game logic, Philox and the record append.
//
// The decision logic around the forward -- fast_begin / fast_decide, block ownership, the counters -- is variant 1's, so
// the records, game words and counters are those of variant 1 run on the same score vectors; the score vectors differ
// from variant 1's only by the summation order of a lane's rotation (held to 1e-5 of the CPU restatement in
// tests/test_gpu_act.py and tests/test_gpu_baseline_sizes.py).  The append is this kernel's own: warp-private buffers
// in shared memory (below).
#include "rollout_fast.cuh"
#include "rollout_tables.cuh"
#include "ptx_helpers.cuh"

namespace nfsp {

// Table slot of a decision state: x * 18 + sigma, x = ((card * 3 + public card) * 2 + dealer) * 4 + finished round-0
// sequence, sigma = 9 * round + sequence id.  One formula for both rounds, no branch in the step loop: round 0 does not see
// the public card and has no finished sequence yet (f = 0), so its 54 states are stored once per public card; slots with
// round 0 and f != 0 are never read.  702 distinct states in 72 * 18 = 1 296 slots per net.
constexpr int kNetStates = 72 * 18;
constexpr int kStateQuads = 4 * kNetStates;      // one float4 {o0, o1, o2, argmax | non-zero << 2} per net and slot
constexpr int kStateBytes = kStateQuads * 16;    // 82 944
static_assert(kStateBytes % 16 == 0, "bulk copies move multiples of 16 bytes");

// table slot -> rows of the factorised first layer (the inverse of the index the rollout computes)
__device__ __forceinline__ void state_rows(int s, uint32_t &xrow, uint32_t &yrow) {
    const uint32_t sg = (uint32_t)s % 18u, x = (uint32_t)s / 18u, dl = (x >> 2) & 1u;
    xrow = sg >= 9u ? 3u + x : x / 24u;  // round 0: the private card alone
    yrow = 75u + dl * 18u + sg;
}

// One table slot per FOUR lanes: lane j of a quad takes hidden-unit quads j, j + 4, j + 8, j + 12 of mlp_forward_tables'
// loop (two table rows added, relu, FFMA2 second layer), the three partial sums are combined with two shuffles, lane 0
// applies the head.  All 20 loads of a lane are in flight at once: the launch is ~3 us on the step path.
__global__ void states_pack_kernel(const float *__restrict__ tab, float4 *__restrict__ states) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, e = t >> 2, j = t & 3;
    const bool valid = e < kStateQuads;  // the grid is a whole number of warps: every lane reaches the shuffles
    uint32_t xrow, yrow;
    state_rows(valid ? e % kNetStates : 0, xrow, yrow);
    const uint32_t net = valid ? (uint32_t)(e / kNetStates) : 0u;
    const float4 *T = reinterpret_cast<const float4 *>(tab);
    const float4 *xr = T + (net * kNetRows + xrow) * kRowQuads, *yr = T + (net * kNetRows + yrow) * kRowQuads;
    const float4 *wr = T + kTabRows * kRowQuads + net * (3 * kRowQuads);
    Layer2Acc acc;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int q = j + 4 * k;
        acc.quad_sum(xr[q], yr[q], wr[q], wr[kRowQuads + q], wr[2 * kRowQuads + q]);
    }
    float s0 = hsum2(acc.z0), s1 = hsum2(acc.z1), s2 = hsum2(acc.z2);
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        s0 += __shfl_xor_sync(0xFFFFFFFFu, s0, o);
        s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, o);
        s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
    }
    if (valid && j == 0) {
        float o0, o1, o2;
        Layer2Acc::head_of_sums(s0, s1, s2, reinterpret_cast<const float4 *>(tab + kTabFloats + kTabW2Floats)[net], net & 1u, o0, o1, o2);
        states[e] = make_float4(o0, o1, o2, __uint_as_float(score_action(o0, o1, o2)));
    }
}

// 1 024 threads per CTA for launches that keep every warp busy with several blocks of games; 512 (16 warps, 90 registers)
// when the launch has no more blocks than that gives warps -- a warp then plays one block, and what counts is the latency
// of its steps, which fewer warps per SM shorten (65 536 games: kernel 24.6 -> 21.3 us)
constexpr int kStatesThreads = 1024, kStatesThreadsSmall = 512;

// ---- warp-private record buffers -----------------------------------------------------------------
// Variant 1 claims its slots with four global atomics per warp and step.  That is one round trip per step on every
// warp's critical path, and with the direct ring append every one of the launch's 262 144 warp-steps hits the SAME two
// words (the rings' totals): same-address atomics retire at ~1.4 ns each, 0.37 ms per launch -- invisible beside variant
// 1's 0.37 ms of shared-memory traffic, the bound of this kernel once the forward is a table read.  Here a warp collects
// its records in shared memory (the table leaves 180 KB free) and claims slots for a whole buffer at a time: one atomic
// and one coalesced copy (512 B per warp instruction) per ~100 records.
constexpr int kBufRL = 104, kBufSL = 32;                 // records per player and warp; >= a step's worst case (64, 32)
constexpr int kWarpBufQuads = 2 * kBufRL + 2 * kBufSL;   // 272 uint4 = 4 352 B per warp
constexpr int states_smem_bytes(int threads) { return kStateBytes + (threads / 32) * kWarpBufQuads * 16; }
static_assert(kBufRL >= 64 && kBufSL >= 32, "a buffer must take the records of one step");

// quad offsets of the four lists in a warp's buffer: RL of player 0 / 1, SL of player 0 / 1
constexpr uint32_t kOffRL1 = kBufRL, kOffSL0 = 2 * kBufRL, kOffSL1 = 2 * kBufRL + kBufSL;

struct WarpBuf {
    uint4 *b;            // this warp's kWarpBufQuads of shared memory
    uint32_t cnt = 0u;   // records held, one byte per list: rl0 | rl1 << 8 | sl0 << 16 | sl1 << 24 (warp-uniform)
};

// inclusive warp scan of four byte counters (no byte exceeds 255); the shuffle's own predicate says whether a lane has a
// source, so a step is SHFL + a predicated add
__device__ __forceinline__ uint32_t warp_scan_bytes_p(uint32_t v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
        asm volatile("{ .reg .pred p; .reg .u32 r; shfl.sync.up.b32 r|p, %0, %1, 0, 0xffffffff; @p add.u32 %0, %0, r; }" : "+r"(v) : "r"(o));
    return v;
}

// copies the n buffered records of one list to base, base + 1, ... of its destination (coalesced: 512 B per warp store)
template <bool kRing, int kCap>
__device__ __forceinline__ void buf_copy(const uint4 *src, uint32_t n, uint4 *dst, uint32_t base, uint32_t cap, int &drop) {
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (int it = 0; it < (kCap + 31) / 32; ++it) {
        const uint32_t k = lane + 32u * it;
        if (k < n) {
            uint32_t off = base + k;
            if (kRing) {
                off = off < cap ? off : off - cap;  // a launch never laps the ring
                dst[off] = src[k];
            } else if (off < cap) {
                dst[off] = src[k];
            } else {
                ++drop;
            }
        }
    }
}

// moves the lists named in `which` (bit l = list l; warp-uniform) to their destinations: the claims of all lists go out
// together (lanes 0-3, one round trip), then one coalesced copy per list
template <bool kDirect>
__device__ __forceinline__ void buf_flush(WarpBuf &B, const RolloutArgs &A, const WarpStage &W, FastCounters &c, uint32_t which) {
    __syncwarp();
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0u;
    if (lane < 4u) {
        const uint32_t n = (B.cnt >> (8u * lane)) & 0xFFu;
        if (n && ((which >> lane) & 1u)) {
            if (kDirect && lane < 2u) base = ring_slot(atomicAdd(A.ring_total[lane], (unsigned long long)n), A.ring_cap, A.ring_magic);
            else base = atomicAdd(W.cnt + lane * W.n_seg, n);
        }
    }
    const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, base, 0), b1 = __shfl_sync(0xFFFFFFFFu, base, 1);
    const uint32_t b2 = __shfl_sync(0xFFFFFFFFu, base, 2), b3 = __shfl_sync(0xFFFFFFFFu, base, 3);
    int drop = 0;
    if (which & 1u) buf_copy<kDirect, kBufRL>(B.b, B.cnt & 0xFFu, W.rl0, b0, W.cap_rl, drop);
    if (which & 2u) buf_copy<kDirect, kBufRL>(B.b + kOffRL1, (B.cnt >> 8) & 0xFFu, W.rl1, b1, W.cap_rl, drop);
    if (which & 4u) buf_copy<false, kBufSL>(B.b + kOffSL0, (B.cnt >> 16) & 0xFFu, W.sl0, b2, W.cap_sl, drop);
    if (which & 8u) buf_copy<false, kBufSL>(B.b + kOffSL1, B.cnt >> 24, W.sl1, b3, W.cap_sl, drop);
    c.wide.drop += drop;
    B.cnt &= ~(((which & 1u) ? 0xFFu : 0u) | ((which & 2u) ? 0xFF00u : 0u) | ((which & 4u) ? 0xFF0000u : 0u) | ((which & 8u) ? 0xFF000000u : 0u));
    __syncwarp();
}

// a step's records into the warp's buffers (what warp_append does with global atomics).  All lanes call it.
template <bool kDirect>
__device__ __forceinline__ void buf_append(WarpBuf &B, const RolloutArgs &A, const WarpStage &W, const FastDecision &d,
                                           const FastRecords &R, float v0, float v1, float v2, FastCounters &c) {
    const uint32_t q = R.q;
    // actor-relative counts: own ring | other ring << 8 | own reservoir << 16; swapping the bytes of each half makes them
    // absolute (rl0 | rl1 << 8 | sl0 << 16 | sl1 << 24) for player 1
    const uint32_t rel = ((uint32_t)R.vA + (uint32_t)R.vB) | (uint32_t)R.vC << 8 | (uint32_t)R.vS << 16;
    const uint32_t mine = q ? __byte_perm(rel, 0u, 0x2301u) : rel;
    const uint32_t incl = warp_scan_bytes_p(mine);
    const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
    // a list that cannot take this step's records is flushed first (warp-uniform, rare): byte > cap <=> bit 7 of
    // byte + 127 - cap (no byte carries: <= 104 + 64 and <= 32 + 32)
    constexpr uint32_t kOverRL = 127u - (uint32_t)kBufRL, kOverSL = 127u - (uint32_t)kBufSL;
    const uint32_t over = (B.cnt + tot + (kOverRL | kOverRL << 8 | kOverSL << 16 | kOverSL << 24)) & 0x80808080u;
    if (over) {
        uint32_t which = ((over >> 7) & 1u) | ((over >> 14) & 2u) | ((over >> 21) & 4u) | ((over >> 28) & 8u);
        if (which & 3u) which |= 3u;  // the two rings fill at the same pace: claim for both in one round trip
        buf_flush<kDirect>(B, A, W, c, which);
    }
    const uint32_t pos = B.cnt + (incl - mine);                       // absolute positions, bytewise
    const uint32_t prel = q ? __byte_perm(pos, 0u, 0x2301u) : pos;    // own ring | other ring | own reservoir
    uint32_t own = q * kOffRL1 + (prel & 0xFFu);
    if (R.vA) B.b[own++] = make_uint4(d.snap_a, d.obs, 0u, d.meta_a);
    if (R.vB) B.b[own] = R.recB;
    if (R.vC) B.b[(q ^ 1u) * kOffRL1 + ((prel >> 8) & 0xFFu)] = R.recC;
    if (R.vS) B.b[kOffSL0 + q * (uint32_t)kBufSL + ((prel >> 16) & 0xFFu)] = make_uint4(d.obs, __float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2));
    B.cnt += tot;
}

template <bool kDebug, bool kDirect, int kThreads>
__global__ void __launch_bounds__(kThreads, 1)
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
    // the warp's offset goes through a shuffle so that it lives in a register: left as arithmetic on threadIdx.x, it was
    // recomputed at each of the step's four buffer stores (16 instructions per warp-step in ncu's source view)
    B.b = reinterpret_cast<uint4 *>(s_tab) + kStateQuads + __shfl_sync(0xFFFFFFFFu, (threadIdx.x >> 5) * (uint32_t)kWarpBufQuads, 0);
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
        if (!image_ready) {  // the table's copy ran beside the set-up and the first loads of the game words
            mbar_wait(bar, 0);
            image_ready = true;
        }
        for (int t = 0; t < A.n_steps; ++t) {
            FastDecision d;
            fast_begin(g, s_lut, A, game, A.step0 + (uint64_t)t, live, d, c);
            const uint32_t ca = (g.PA >> 11) & 3u;
            const uint32_t x = (((ca * 3u + g.pub()) * 2u + g.dealer()) << 2) | g.fin0();
            const float4 o = s_tab[(g.p() * 2u + (uint32_t)d.pol) * (uint32_t)kNetStates + x * 18u + g.sigma()];
            float v0 = o.x, v1 = o.y, v2 = o.z;
            uint32_t pre = __float_as_uint(o.w);
            if (d.random) { v0 = d.r0; v1 = d.r1; v2 = d.r2; pre = score_action(v0, v1, v2); }
            FastRecords R;
            fast_decide<kDebug>(g, s_lut, A, d, v0, v1, v2, live, (int64_t)t * A.n + i, plane, c, R, pre);
            buf_append<kDirect>(B, A, W, d, R, v0, v1, v2, c);
            if ((t & 15) == 15) c.spill();
        }
        c.spill();
        if (live) A.state[i] = g.pack();
        c.wide.trans += live ? A.n_steps : 0;
        // a block's staged records belong to the block's segment; the rings have no segments, their records stay buffered
        buf_flush<kDirect>(B, A, W, c, kDirect ? 12u : 15u);
    }
    if (kDirect) buf_flush<true>(B, A, W, c, 3u);
    if (!image_ready) mbar_wait(bar, 0);  // the copy into this CTA's shared memory must land before the CTA exits
    c.spill();
    if (A.stats) c.wide.commit(s_stats, A.stats);
}

}  // namespace nfsp

using namespace nfsp;

int nfsp_states_floats() { return kStateQuads * 4; }

template <int kThreads>
static int configure_states() {
    constexpr int kBytes = states_smem_bytes(kThreads);
    NFSP_CUDA(cudaFuncSetAttribute(rollout_states_kernel<true, true, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes));
    NFSP_CUDA(cudaFuncSetAttribute(rollout_states_kernel<true, false, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes));
    NFSP_CUDA(cudaFuncSetAttribute(rollout_states_kernel<false, true, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes));
    NFSP_CUDA(cudaFuncSetAttribute(rollout_states_kernel<false, false, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes));
    return NFSP_OK;
}

int nfsp_rollout_states_configure() {
    const int rc = configure_states<kStatesThreads>();
    return rc != NFSP_OK ? rc : configure_states<kStatesThreadsSmall>();
}

// rebuilds the state table from the table image (both live in d_wpack) when the weights have changed since
int nfsp_states_ensure(nfsp_env_t h, const float *d_tab, float *d_states, cudaStream_t st) {
    if (!h->st_dirty) return NFSP_OK;
    states_pack_kernel<<<(4 * kStateQuads + 127) / 128, 128, 0, st>>>(d_tab, reinterpret_cast<float4 *>(d_states));
    NFSP_LAUNCH_CHECK();
    h->st_dirty = false;
    return NFSP_OK;
}

template <int kThreads>
static int launch_states(const RolloutArgs &A, int grid, bool debug, bool direct, cudaStream_t st) {
    constexpr int kBytes = states_smem_bytes(kThreads);
    if (debug && direct) rollout_states_kernel<true, true, kThreads><<<grid, kThreads, kBytes, st>>>(A);
    else if (debug) rollout_states_kernel<true, false, kThreads><<<grid, kThreads, kBytes, st>>>(A);
    else if (direct) rollout_states_kernel<false, true, kThreads><<<grid, kThreads, kBytes, st>>>(A);
    else rollout_states_kernel<false, false, kThreads><<<grid, kThreads, kBytes, st>>>(A);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

int nfsp_rollout_states_launch(nfsp_env_t h, const RolloutArgs &A, const nfsp_rollout_io *io, bool debug, cudaStream_t st) {
    const bool direct = A.ring[0] != nullptr;
    const int sms = h->sm_count - io->reserve_sms;
    const int grid = grid_for(h->n, 32, sms, 1);  // at least one block of 32 games per CTA
    const int64_t blocks = (h->n + 31) / 32;
    if (blocks <= (int64_t)sms * (kStatesThreadsSmall / 32)) return launch_states<kStatesThreadsSmall>(A, grid, debug, direct, st);
    return launch_states<kStatesThreads>(A, grid, debug, direct, st);
}
