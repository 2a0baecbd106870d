// K2, variant 4 (the library default): the CUDA-core fused rollout with the decisions of a warp GROUP sorted by net.
//
// rollout_kernel (act_kernels.cu) is bound by shared-memory wavefronts: of the 354 a warp-step costs, 192 are the second
// layer's weights, because the four nets are mixed inside a warp and every lane fetches "its" W2 (ncu r01).  Here the 128
// decisions a group of four warps takes per step are counting-sorted by net (two ballots per warp, one packed count
// word per warp in shared memory), each thread evaluates the MLP of the decision in ITS sorted slot and hands the three
// scores back through shared memory.  Warps are then net-homogeneous except at the three bin boundaries, the rotated W2
// reads of a warp collapse onto 8 distinct 16-byte addresses (1 wavefront instead of 4), and nothing else changes: same
// table image, same arithmetic per decision (agent.py:101-103,110-112,124-128,143), same game logic (rollout_fast.cuh).
//
// The records (agent.py:134-136,151, main.py:55-67) are appended with ONE set of global atomics per group and step
// instead of one per warp: warps add their totals to a shared-memory accumulator, the last one to arrive claims the
// group's slots.  That makes a single cursor per memory cheap enough (64 k atomics per launch and address), so
//   * the staging arrays are plain dense arrays (n_segments = 1), and
//   * the RL records can go STRAIGHT INTO THE PLAYERS' RINGS (nfsp_rollout_io.d_ring*): ticket = atomicAdd on the
//     ring's own "records ever inserted" counter, slot = ticket % capacity (replay_buffer.py:36-41) -- no staging copy,
//     no insert/commit launches for 12/13 of the records.  Precondition: a launch cannot lap the ring
//     (2 * n * n_steps <= capacity), checked by the host entry.
#include "ptx_helpers.cuh"
#include "rollout_fast.cuh"
#include "rollout_tables.cuh"

namespace nfsp {

constexpr int kSortThreads = 1024;
constexpr int kSortG = 4;                       // warps per sorting group
constexpr int kSortGT = 32 * kSortG;            // decisions per group and step
constexpr int kSortGroups = kSortThreads / kSortGT;
constexpr int kSortImagePad = (kTabImageBytes + 127) / 128 * 128;

struct __align__(16) GroupShared {
    uint32_t cnt[kSortG];   // per warp: decisions of sort key 0..3, one byte each
    uint32_t base[4];       // this step's first slot of the group in rl0, rl1, sl0, sl1
    uint32_t acc[2];        // records of the step so far: rl0 | rl1 << 16, sl0 | sl1 << 16
    uint32_t arrive, pad;
    uint32_t desc[kSortGT];  // sorted decisions: xrow | (yrow - 75) << 7 | net << 13
    float4 res[kSortGT];     // their scores
};
constexpr int kSortSmemBytes = kSortImagePad + kSortGroups * (int)sizeof(GroupShared);

__device__ __forceinline__ void group_bar(uint32_t id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kSortGT) : "memory");
}

template <bool kDebug, bool kDirect>
__global__ void __launch_bounds__(kSortThreads, 1)
rollout_sorted_kernel(const RolloutArgs A) {
    extern __shared__ __align__(128) float sw[];  // table image, then the groups' exchange areas
    __shared__ unsigned long long s_stats[NFSP_STATS_FIELDS];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ FastLuts s_lut;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t grp = warp / kSortG, wig = warp % kSortG, tig = threadIdx.x % kSortGT;
    GroupShared &S = reinterpret_cast<GroupShared *>(reinterpret_cast<uint8_t *>(sw) + kSortImagePad)[grp];
    if (threadIdx.x < NFSP_STATS_FIELDS) s_stats[threadIdx.x] = 0ull;
    if (tig == 0) { S.acc[0] = 0u; S.acc[1] = 0u; S.arrive = 0u; }
    s_lut.fill();
    const uint32_t bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
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
    const uint32_t rot = lane & 7u;
    const uint32_t bar_id = 1u + grp;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint4 *const rl0 = kDirect ? A.ring[0] : A.rl[0], *const rl1 = kDirect ? A.ring[1] : A.rl[1];
    const uint32_t cap_rl = kDirect ? A.ring_cap : (uint32_t)A.cap_rl, cap_sl = (uint32_t)A.cap_sl;
    const int64_t n_blocks = (A.n + kSortGT - 1) / kSortGT;
    // blocks of 128 consecutive games, dealt round-robin to the chip's warp groups (every SM gets its share of a small batch)
    for (uint32_t blk = blockIdx.x + gridDim.x * grp; (int64_t)blk < n_blocks; blk += gridDim.x * kSortGroups) {
        const int64_t i = (int64_t)blk * kSortGT + tig;
        const bool live = i < A.n;
        const uint64_t game = A.game0 + (uint64_t)i;
        NfspFast g;
        g.unpack(live ? A.state[i] : 0ull);
        for (int t = 0; t < A.n_steps; ++t) {
            FastDecision d;
            fast_begin(g, s_lut, A, game, A.step0 + (uint64_t)t, live, d, c);
            uint32_t slot;
            {
                const uint32_t sg = g.sigma(), dl = g.dealer(), ca = (g.PA >> 11) & 3u;
                const uint32_t xrow = sg >= 9u ? 3u + ((((ca * 3u + g.pub()) * 2u + dl) << 2) | g.fin0()) : ca;
                // a phantom lane (beyond the last game) sorts behind every live decision: the slot of a game, and with it
                // the summation order of its scores, depends on the games of its block only, never on the launch geometry
                const uint32_t net = live ? g.p() * 2u + (uint32_t)d.pol : 2u;
                // ---- counting sort of the group's 128 decisions by net.  Key order avg0, br0, br1, avg1: the two small
                // best-response bins sit between the two large ones, so at most three warps see two nets
                const uint32_t key = net ^ (net >> 1);
                const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, key & 1u), b1 = __ballot_sync(0xFFFFFFFFu, key & 2u);
                const uint32_t same = ((key & 1u) ? b0 : ~b0) & ((key & 2u) ? b1 : ~b1);
                if (lane == 0) {
                    const uint32_t c1 = __popc(b0 & ~b1), c2 = __popc(~b0 & b1), c3 = __popc(b0 & b1);
                    S.cnt[wig] = (32u - c1 - c2 - c3) | (c1 << 8) | (c2 << 16) | (c3 << 24);
                }
                group_bar(bar_id);
                const uint4 cw = *reinterpret_cast<const uint4 *>(S.cnt);
                const uint32_t tot = cw.x + cw.y + cw.z + cw.w;  // bytes <= 128: no carries
                const uint32_t before = (wig > 0 ? cw.x : 0u) + (wig > 1 ? cw.y : 0u) + (wig > 2 ? cw.z : 0u);
                slot = (((tot * 0x01010100u + before) >> (8u * key)) & 0xFFu) + __popc(same & lt_mask);
                S.desc[slot] = xrow | ((dl * 18u + sg) << 7) | (net << 13);
            }
            group_bar(bar_id);
            {
                const uint32_t e = S.desc[tig];
                if (!image_ready) {
                    mbar_wait(bar, 0);
                    image_ready = true;
                }
                float o0, o1, o2;
                mlp_forward_tables(sw, e & 127u, 75u + ((e >> 7) & 63u), e >> 13, rot, o0, o1, o2);
                S.res[tig] = make_float4(o0, o1, o2, 0.f);
            }
            group_bar(bar_id);
            const float4 r = S.res[slot];
            float v0 = r.x, v1 = r.y, v2 = r.z;
            if (d.random) { v0 = d.r0; v1 = d.r1; v2 = d.r2; }
            FastRecords R;
            fast_decide<kDebug>(g, s_lut, A, d, v0, v1, v2, live, (int64_t)t * A.n + i, plane, c, R);
            // ---- append: warp scan, the warp's totals into the group's accumulators, the last warp claims the slots
            const uint32_t mine = record_counts(R);
            const uint32_t incl = warp_scan_bytes(mine);
            const uint32_t wtot = __shfl_sync(0xFFFFFFFFu, incl, 31);
            uint32_t woff = 0u, arrived = 0u;
            if (lane < 2) {
                const uint32_t two = wtot >> (16u * lane);
                woff = atomicAdd(&S.acc[lane], (two & 0xFFu) | ((two & 0xFF00u) << 8));
            }
            __syncwarp();
            if (lane == 0) arrived = atomicAdd(&S.arrive, 1u);
            arrived = __shfl_sync(0xFFFFFFFFu, arrived, 0);
            if (arrived == (uint32_t)kSortG - 1u) {  // warp-uniform: every warp of the group has added its totals
                __threadfence_block();
                if (lane < 4) {
                    const uint32_t n_rec = (reinterpret_cast<volatile uint32_t *>(S.acc)[lane >> 1] >> (16u * (lane & 1u))) & 0xFFFFu;
                    uint32_t b = 0u;
                    if (n_rec) {
                        if (kDirect && lane < 2) b = ring_slot(atomicAdd(A.ring_total[lane], (unsigned long long)n_rec), A.ring_cap, A.ring_magic);
                        else b = atomicAdd(A.counts + lane, n_rec);
                    }
                    S.base[lane] = b;
                }
                __syncwarp();
                if (lane == 0) { S.acc[0] = 0u; S.acc[1] = 0u; S.arrive = 0u; }
            }
            group_bar(bar_id);
            {
                const uint4 B = *reinterpret_cast<const uint4 *>(S.base);
                const uint32_t wrl = __shfl_sync(0xFFFFFFFFu, woff, 0), wsl = __shfl_sync(0xFFFFFFFFu, woff, 1);
                const uint32_t excl = incl - mine, q = R.q, sh = 8u * q;
                uint4 *rp = q ? rl1 : rl0, *ro = q ? rl0 : rl1, *sp = q ? A.sl[1] : A.sl[0];
                uint32_t off = (q ? B.y : B.x) + ((wrl >> (2u * sh)) & 0xFFFFu) + ((excl >> sh) & 0xFFu);
                int drop = 0;
                if (R.vA) {
                    if (kDirect) rp[off < cap_rl ? off : off - cap_rl] = make_uint4(d.snap_a, d.obs, 0u, d.meta_a);
                    else if (off < cap_rl) rp[off] = make_uint4(d.snap_a, d.obs, 0u, d.meta_a);
                    else ++drop;
                    ++off;
                }
                if (R.vB) {
                    if (kDirect) rp[off < cap_rl ? off : off - cap_rl] = R.recB;
                    else if (off < cap_rl) rp[off] = R.recB;
                    else ++drop;
                }
                if (R.vC) {
                    const uint32_t o2 = (q ? B.x : B.y) + ((wrl >> (16u - 2u * sh)) & 0xFFFFu) + ((excl >> (8u - sh)) & 0xFFu);
                    if (kDirect) ro[o2 < cap_rl ? o2 : o2 - cap_rl] = R.recC;
                    else if (o2 < cap_rl) ro[o2] = R.recC;
                    else ++drop;
                }
                if (R.vS) {
                    const uint32_t o3 = (q ? B.w : B.z) + ((wsl >> (2u * sh)) & 0xFFFFu) + ((excl >> (16u + sh)) & 0xFFu);
                    if (o3 < cap_sl) sp[o3] = make_uint4(d.obs, __float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2));
                    else ++drop;
                }
                c.wide.drop += drop;
            }
            if ((t & 15) == 15) c.spill();
        }
        c.spill();
        if (live) A.state[i] = g.pack();
        c.wide.trans += live ? A.n_steps : 0;
    }
    if (!image_ready) mbar_wait(bar, 0);  // the copy into this CTA's shared memory must land before the CTA may exit
    c.spill();
    if (A.stats) c.wide.commit(s_stats, A.stats);
}

}  // namespace nfsp

using namespace nfsp;

template <bool kDebug, bool kDirect>
static int launch_sorted(const RolloutArgs &A, int grid, cudaStream_t st) {
    rollout_sorted_kernel<kDebug, kDirect><<<grid, kSortThreads, kSortSmemBytes, st>>>(A);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

// the dynamic shared-memory limit is a per-device attribute: called once per handle (nfsp_act_set_weights)
int nfsp_rollout_sorted_configure() {
    NFSP_CUDA(cudaFuncSetAttribute(rollout_sorted_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortSmemBytes));
    NFSP_CUDA(cudaFuncSetAttribute(rollout_sorted_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortSmemBytes));
    NFSP_CUDA(cudaFuncSetAttribute(rollout_sorted_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortSmemBytes));
    NFSP_CUDA(cudaFuncSetAttribute(rollout_sorted_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortSmemBytes));
    return NFSP_OK;
}

int nfsp_rollout_sorted_launch(nfsp_env_t h, const RolloutArgs &A, const nfsp_rollout_io *io, bool debug, cudaStream_t st) {
    NFSP_CHECK_ARG(io->n_segments == 1, "variant 4 appends through one cursor per memory: n_segments must be 1");
    const bool direct = A.ring[0] != nullptr;
    int grid = h->sm_count - io->reserve_sms;
    const int64_t need = (A.n + kSortGT - 1) / kSortGT;  // at least one block of 128 games per CTA
    if (need < grid) grid = (int)need;
    if (grid < 1) grid = 1;
    if (debug) return direct ? launch_sorted<true, true>(A, grid, st) : launch_sorted<true, false>(A, grid, st);
    return direct ? launch_sorted<false, true>(A, grid, st) : launch_sorted<false, false>(A, grid, st);
}
