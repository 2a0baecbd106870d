// K2, tensor-core variant of the first layer of the acting nets: tcgen05.mma with the accumulator in TMEM.
//
// Shape (from profiles/r01/umma_microbench.txt, B200): a tcgen05.mma costs ~90 cycles whatever N <= 128 and ~150
// cycles at N = 256, plus ~450 cycles issue->commit latency per chain, so the tile that minimises tensor time
// is the WIDE one: D[128 rows][256 = 4 nets x 64 hidden] = A[128][32] x B[256][32]^T.  A = the 30 observation
// bits + the constant 1 that carries b1 (+ a zero), exactly 0/1 in bf16; B = the fp32 weights split EXACTLY into
// three bf16 terms (w = hi + mid + lo), so 2 k-steps x 3 splits = 6 MMAs (M128 N256 K16, fp32 accumulation)
// reproduce the fp32 layer to accumulation-order rounding -- plain bf16 / tf32 products are ~1e-3 relative, far
// from the 1e-5 parity bound.  Layer 2 (64x3), the heads and everything else stay on CUDA cores.
//
// A row needs only its own net's 64 columns, so the rows of a tile are SORTED BY NET (counting sort over the
// group's 128 threads) before they are written as operand rows: a warp's tcgen05.ld then reads one 64-column
// block of TMEM (two at a segment boundary) and its W2 reads are warp-uniform.
//
// Both kernels here are persistent (one CTA per SM) and run kGroups groups of 128 threads; a group owns one operand
// tile in shared memory and takes tiles of 128 rows from a global counter.  The 512 TMEM columns are two
// accumulator slots; a group uses slot (group & 1) under a shared-memory lock taken by its MMA issuer and released
// by the last of its four warps to finish reading the accumulator, so the tensor pipe works on one slot while the
// other is being read.
//
// Operand images (K-major, no swizzle; core matrix = 8 rows x 16 bytes):
//   byte offset(row, k) = (row/8)*512 + (k/8)*128 + (row%8)*16 + (k%8)*2      (LBO = 128 B, SBO = 512 B)
//   A: 128 rows x 32 k = 8 KB per group, written by the threads (one sorted row each, 8 inputs per table lookup)
//   B: 3 splits x (256 n x 32 k) = 3 x 16 KB, built once per weight update, brought in by one bulk async copy
#include <cuda_bf16.h>

#include "mlp_math.cuh"
#include "ptx_helpers.cuh"
#include "rollout_fast.cuh"

namespace nfsp {

constexpr int kGroups = 6;
constexpr int kTcThreads = 128 * kGroups;
constexpr int kFwdThreads = 128 * 4;  // the forward kernel keeps a row's 64 accumulator columns in registers: 4 groups
constexpr int kTcW2Floats = 16 * 4 * 3 * 4 + 16;   // W2 as [16 quads][4 nets][3 outputs][4] + b2 [4][4]
constexpr int kWideABytes = 128 * 32 * 2;          // 8 KB operand tile per group
constexpr int kWideBSplitBytes = 256 * 32 * 2;     // 16 KB
constexpr int kWideBBytes = 3 * kWideBSplitBytes;  // 48 KB
constexpr int kWideImageBytes = kWideBBytes + kTcW2Floats * 4;
constexpr uint32_t kLBO = 128, kWideSBO = 512;     // 4 k-chunks of 128 B per 8-row group
constexpr uint32_t kIdescWide = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
// dynamic smem: A tiles | weight image | results float4[kGroups][128] | byte -> 8 x bf16 table (x8) | owner u8[kGroups][128]
//               | counts | tile ids | barriers | slot locks + TMEM base
constexpr int kOffImage = kGroups * kWideABytes;
constexpr int kOffResult = kOffImage + kWideImageBytes;
constexpr int kOffBits = kOffResult + kGroups * 128 * 16;
constexpr int kOffOwner = kOffBits + 256 * 8 * 16;  // 8 skewed copies of the byte table: lane & 7 picks the bank group
constexpr int kOffCnt = kOffOwner + kGroups * 128;
constexpr int kOffTile = kOffCnt + kGroups * 4 * 4;
constexpr int kOffBars = kOffTile + kGroups * 4 + 8;
constexpr int kOffLocks = kOffBars + 8 * (1 + kGroups);
constexpr int kTcSmemBytes = kOffLocks + 4 * 4 + 16;

// ---- weight image ---------------------------------------------------------------------------------
__global__ void pack_tc_wide_kernel(const float *__restrict__ w, uint8_t *__restrict__ img) {
    const int total = 3 * 256 * 32;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int split = e / (256 * 32), n = (e / 32) % 256, k = e % 32;
        const int net = n >> 6, jj = n & 63;
        float v = 0.f;
        if (k < 30) v = w[net * NFSP_NET_PARAMS + k * 64 + jj];
        else if (k == 30) v = w[net * NFSP_NET_PARAMS + 1920 + jj];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const float r1 = v - __bfloat162float(hi);
        const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
        const float r2 = r1 - __bfloat162float(mid);
        const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
        const __nv_bfloat16 pick = split == 0 ? hi : (split == 1 ? mid : lo);
        const size_t off = (size_t)split * kWideBSplitBytes + (n >> 3) * kWideSBO + (k >> 3) * kLBO + (n & 7) * 16 + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(img + off) = pick;
    }
    float *w2 = reinterpret_cast<float *>(img + kWideBBytes);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < kTcW2Floats; e += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (e < 16 * 4 * 3 * 4) {
            const int x = e & 3, g = e >> 2, c = g % 3, net = (g / 3) & 3, q = g / 12;
            v = w[net * NFSP_NET_PARAMS + 1984 + (q * 4 + x) * 3 + c];
        } else {
            const int f = e - 16 * 4 * 3 * 4, c = f & 3, net = f >> 2;
            if (c < 3) v = w[net * NFSP_NET_PARAMS + 2176 + c];
        }
        w2[e] = v;
    }
}

// ---- tcgen05 PTX wrappers (mbarrier / bulk copy wrappers: ptx_helpers.cuh) ------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_wide(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(kLBO >> 4) << 16) | ((uint64_t)(kWideSBO >> 4) << 32) |
           (1ull << 46);
}
// D f32, A/B bf16, both K-major, M = 128, N = 256
__device__ __forceinline__ void umma_f16_wide(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(a), "l"(b), "r"(kIdescWide), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void group_bar(uint32_t id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// 32 lanes x 16 consecutive columns: thread t of warp w reads TMEM lane 32*(w%4)+t.  The _nowait form only issues
// the load: the values may be used after tmem_wait_ld().
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    tmem_ld16_nowait(taddr, v);
    tmem_wait_ld();
}

// 8 input bits -> 8 bf16 values (0.0 / 1.0) = one 16-byte operand chunk
__device__ __forceinline__ uint4 bits_to_bf16x8(uint32_t b) {
    uint4 c;
    c.x = ((b >> 0) & 1u) * 0x3F80u + ((b >> 1) & 1u) * 0x3F800000u;
    c.y = ((b >> 2) & 1u) * 0x3F80u + ((b >> 3) & 1u) * 0x3F800000u;
    c.z = ((b >> 4) & 1u) * 0x3F80u + ((b >> 5) & 1u) * 0x3F800000u;
    c.w = ((b >> 6) & 1u) * 0x3F80u + ((b >> 7) & 1u) * 0x3F800000u;
    return c;
}

// ---- the machinery both kernels share -------------------------------------------------------------------
// Per-thread view of its group's shared-memory pieces and of the CTA-wide ones.
struct TcGroup {
    uint8_t *sA;
    const float *sW2;
    float4 *sResult;
    const uint4 *sBits;
    uint8_t *sOwner;
    uint32_t *sCnt;
    volatile uint32_t *sTile;
    uint32_t *slot_busy, *slot_readers;
    uint32_t a_smem, b_smem, bar_done, tmem_slot, tmem_acc;
    uint32_t group, gtid, wq, lane, ph_done;

    // CTA-wide set-up: byte table, barriers, weight image (bulk async copy), TMEM allocation.  Returns the TMEM base.
    __device__ __forceinline__ uint32_t setup(uint8_t *smem, const void *image) {
        group = threadIdx.x >> 7; gtid = threadIdx.x & 127u; wq = gtid >> 5; lane = gtid & 31u; ph_done = 0u;
        sA = smem + group * kWideABytes;
        uint8_t *sB = smem + kOffImage;
        sW2 = reinterpret_cast<const float *>(sB + kWideBBytes);
        sResult = reinterpret_cast<float4 *>(smem + kOffResult) + group * 128;
        sBits = reinterpret_cast<const uint4 *>(smem + kOffBits);
        sOwner = smem + kOffOwner + group * 128;
        sCnt = reinterpret_cast<uint32_t *>(smem + kOffCnt) + group * 4;
        sTile = reinterpret_cast<volatile uint32_t *>(smem + kOffTile) + group;
        uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBars);
        uint32_t *locks = reinterpret_cast<uint32_t *>(smem + kOffLocks);  // [0..1] slot busy, [2..3] warps done reading
        uint32_t *tmem_word = locks + 4;
        slot_busy = locks + (group & 1u);
        slot_readers = locks + 2 + (group & 1u);
        const uint32_t bar_w = smem_u32(bars);
        bar_done = smem_u32(bars + 1 + group);
        for (uint32_t e = threadIdx.x; e < 2048u; e += blockDim.x) reinterpret_cast<uint4 *>(smem + kOffBits)[e] = bits_to_bf16x8(e >> 3);
        if (threadIdx.x < 4) locks[threadIdx.x] = 0u;
        if (threadIdx.x == 0) {
            for (int k = 0; k < 1 + kGroups; ++k) mbar_init(smem_u32(bars + k), 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar_w, kWideImageBytes);
            bulk_g2s(smem_u32(sB), image, kWideImageBytes, bar_w);
        }
        if (threadIdx.x < 32) tmem_alloc(smem_u32(tmem_word), 512);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        const uint32_t tmem_base = *tmem_word;
        tmem_slot = tmem_base + (group & 1u) * 256u;
        tmem_acc = tmem_slot + (wq << 21);  // this warp's 32 TMEM lanes
        a_smem = smem_u32(sA);
        b_smem = smem_u32(sB);
        mbar_wait(bar_w, 0);
        return tmem_base;
    }

    // next tile of 128 rows from the global counter (one atomic per group and tile)
    __device__ __forceinline__ int64_t next_tile(uint32_t *work) {
        if (gtid == 0) *sTile = atomicAdd(work, 1u);
        group_bar(1 + group);
        return (int64_t)*sTile;
    }

    // 16 accumulator columns (chunk ch of the row's net) of this thread's sorted row
    __device__ __forceinline__ void read_columns(uint32_t my_key, uint32_t my_net, uint32_t k_lo, uint32_t k_hi, int ch, float *h) {
        if (k_lo == k_hi) {  // the whole warp reads one net's columns (the common case)
            tmem_ld16(tmem_acc + my_net * 64u + ch * 16u, h);
        } else {             // segment boundary inside the warp: one TMEM read per net present, select
#pragma unroll
            for (int e = 0; e < 16; ++e) h[e] = 0.f;
            for (uint32_t kk = k_lo; kk <= k_hi; ++kk) {
                const uint32_t nn = kk ^ (kk >> 1);
                float t16[16];
                tmem_ld16(tmem_acc + nn * 64u + ch * 16u, t16);
                if (kk == my_key) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) h[e] = t16[e];
                }
            }
        }
    }
    // the warp has its last accumulator columns in registers: the 4th warp of the group to say so frees the slot
    __device__ __forceinline__ void release_slot() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0 && atomicAdd(slot_readers, 1u) == 3u) {
            *reinterpret_cast<volatile uint32_t *>(slot_readers) = 0u;
            __threadfence_block();
            atomicExch(slot_busy, 0u);
        }
    }

    // One forward of the group's 128 rows.  Every thread passes ITS row (observation mask, net); on return o0..o2
    // are the outputs of SORTED row gtid and the return value is the thread that owns that row.  Group-collective.
    // kEarly: read all 64 accumulator columns into registers first and free the slot before layer 2 (needs 64 more
    // live registers: the forward kernel can afford them, the fused rollout cannot).
    template <bool kEarly>
    __device__ __forceinline__ uint32_t forward_sorted(uint32_t obs, uint32_t net, float &o0, float &o1, float &o2) {
        // ---- counting sort of the group's rows by net: packed byte counters, one word per warp.  Sort key order
        // avg0, br0, br1, avg1 keeps the two small best-response segments adjacent, so fewer warps straddle a
        // segment boundary.
        const uint32_t key = net ^ (net >> 1);  // net 0,1,2,3 -> key 0,1,3,2
        const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, key == 0), m1 = __ballot_sync(0xFFFFFFFFu, key == 1);
        const uint32_t m2 = __ballot_sync(0xFFFFFFFFu, key == 2), m3 = ~(m0 | m1 | m2);
        const uint32_t mine = key == 0 ? m0 : (key == 1 ? m1 : (key == 2 ? m2 : m3));
        const uint32_t rank = __popc(mine & ((1u << lane) - 1u));
        if (lane == 0) sCnt[wq] = __popc(m0) | (__popc(m1) << 8) | (__popc(m2) << 16) | (__popc(m3) << 24);
        group_bar(1 + group);  // also: every thread is done with the previous tile's owner map and results
        const uint32_t c0 = sCnt[0], c1 = sCnt[1], c2 = sCnt[2], c3 = sCnt[3];
        const uint32_t tot = c0 + c1 + c2 + c3;  // bytes: rows of key 0..3 (<= 128 each, no carry)
        const uint32_t before = (wq > 0 ? c0 : 0u) + (wq > 1 ? c1 : 0u) + (wq > 2 ? c2 : 0u);
        const uint32_t seg1 = tot & 0xFFu, seg2 = seg1 + ((tot >> 8) & 0xFFu), seg3 = seg2 + ((tot >> 16) & 0xFFu);
        const uint32_t seg_start = key == 0 ? 0u : (key == 1 ? seg1 : (key == 2 ? seg2 : seg3));
        const uint32_t pos = seg_start + ((before >> (8 * key)) & 0xFFu) + rank;
        {  // operand row `pos`: observation bits + the constant 1 that carries b1, 8 inputs per table lookup
            uint8_t *row = sA + (pos >> 3) * kWideSBO + (pos & 7u) * 16;
            const uint32_t x = (obs & 0x3FFFFFFFu) | (1u << 30);
#pragma unroll
            for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4 *>(row + k * kLBO) = sBits[((x >> (8 * k)) & 0xFFu) * 8u + (lane & 7u)];
            sOwner[pos] = (uint8_t)gtid;
        }
        fence_async_smem();
        group_bar(1 + group);
        if (gtid == 0) {
            while (atomicCAS(slot_busy, 0u, 1u) != 0u) __nanosleep(32);  // the accumulator slot is ours
            tc_fence_after();
#pragma unroll
            for (int split = 0; split < 3; ++split)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                    umma_f16_wide(tmem_slot, umma_desc_wide(a_smem + ks * 2 * kLBO),
                                  umma_desc_wide(b_smem + split * kWideBSplitBytes + ks * 2 * kLBO), (split | ks) != 0);
            umma_commit(bar_done);
        }
        mbar_wait_backoff(bar_done, ph_done);  // every thread waits on the mbarrier itself (measured faster than one
        ph_done ^= 1u;                         // polling warp + a group barrier: 0.67 vs 0.70 ms per rollout launch)
        tc_fence_after();
        // ---- epilogue of sorted row `gtid`: its net's 64 pre-activations -> layer 2 -> head
        const uint32_t r = gtid;
        const uint32_t my_key = (r >= seg1) + (r >= seg2) + (r >= seg3);
        const uint32_t my_net = my_key ^ (my_key >> 1);  // inverse of the key map
        const uint32_t r_lo = wq * 32u, r_hi = r_lo + 31u;
        const uint32_t k_lo = (r_lo >= seg1) + (r_lo >= seg2) + (r_lo >= seg3);
        const uint32_t k_hi = (r_hi >= seg1) + (r_hi >= seg2) + (r_hi >= seg3);
        const float4 *w2 = reinterpret_cast<const float4 *>(sW2) + my_net * 3;
        Layer2Acc acc;
        if (kEarly) {
            float h[64];
            if (k_lo == k_hi) {  // four loads in flight, one wait
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) tmem_ld16_nowait(tmem_acc + my_net * 64u + ch * 16u, h + 16 * ch);
                tmem_wait_ld();
            } else {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) read_columns(my_key, my_net, k_lo, k_hi, ch, h + 16 * ch);
            }
            release_slot();
#pragma unroll
            for (int q = 0; q < 16; ++q)
                acc.quad(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3], w2[q * 12], w2[q * 12 + 1], w2[q * 12 + 2]);
        } else {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float h[16];
                read_columns(my_key, my_net, k_lo, k_hi, ch, h);
                if (ch == 3) release_slot();
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    acc.quad(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3], w2[(ch * 4 + q) * 12],
                             w2[(ch * 4 + q) * 12 + 1], w2[(ch * 4 + q) * 12 + 2]);
            }
        }
        acc.head(reinterpret_cast<const float4 *>(sW2 + 16 * 4 * 3 * 4)[my_net], my_net & 1u, o0, o1, o2);
        return (uint32_t)sOwner[r];
    }
};

// ---- batched Model.predict on the tensor cores ----------------------------------------------------------
__global__ void __launch_bounds__(kFwdThreads, 1)
act_forward_tc_kernel(const uint8_t *__restrict__ img, const uint32_t *__restrict__ obs, const int8_t *__restrict__ net,
                      int64_t n, float *__restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    TcGroup G;
    const uint32_t tmem_base = G.setup(smem, img);
    // every tile costs the same here, so tiles are dealt statically (no atomic, no barrier) and the inputs of a
    // group's next tile are loaded while the current one is in flight: the ~1000-cycle global-load latency is off
    // the group's critical path
    const int64_t tiles = (n + 127) >> 7, stride = (int64_t)gridDim.x * (kFwdThreads / 128);
    int64_t tile = (int64_t)blockIdx.x * (kFwdThreads / 128) + G.group;
    int64_t i = tile * 128 + G.gtid;
    uint32_t o_cur = (tile < tiles && i < n) ? obs[i] : 0u, k_cur = (tile < tiles && i < n) ? (uint32_t)(net[i] & 3) : 0u;
    for (; tile < tiles; tile += stride) {
        const int64_t i_next = (tile + stride) * 128 + G.gtid;
        const bool more = tile + stride < tiles && i_next < n;
        const uint32_t o_next = more ? obs[i_next] : 0u, k_next = more ? (uint32_t)(net[i_next] & 3) : 0u;
        float o0, o1, o2;
        const uint32_t owner = G.forward_sorted<true>(o_cur, k_cur, o0, o1, o2);
        const int64_t dst = tile * 128 + owner;
        if (dst < n) { out[3 * dst] = o0; out[3 * dst + 1] = o1; out[3 * dst + 2] = o2; }
        o_cur = o_next;
        k_cur = k_next;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

// ---- fused rollout, tcgen05 first layer -------------------------------------------------------------------
// Thread t of a group owns game t of the tile (actor-relative state in registers over the launch's steps,
// nfsp_fast.cuh) and is the epilogue worker of sorted row t; results travel back to the owners through shared memory.
template <bool kDebug>
__global__ void __launch_bounds__(kTcThreads, 1)
rollout_tc_kernel(const RolloutArgs A) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ unsigned long long s_stats[NFSP_STATS_FIELDS];
    __shared__ FastLuts s_lut;
    if (threadIdx.x < NFSP_STATS_FIELDS) s_stats[threadIdx.x] = 0ull;
    s_lut.fill();
    TcGroup G;
    const uint32_t tmem_base = G.setup(smem, A.pack);

    FastCounters c;
#ifdef NFSP_TC_TIMING
    long long tm_mma = 0, tm_epi = 0, tm_fin = 0, tm_beg = 0; long long t0, t1;
#define TC_T0() t0 = clock64()
#define TC_T1(acc) do { t1 = clock64(); acc += t1 - t0; t0 = t1; } while (0)
#else
#define TC_T0()
#define TC_T1(acc)
#endif
    const int64_t plane = (int64_t)A.n_steps * A.n;
    const int64_t tiles = (A.n + 127) >> 7;
    for (;;) {
        const int64_t tile = G.next_tile(A.work);  // tiles of 128 games are handed out dynamically
        if (tile >= tiles) break;
        const int64_t i = tile * 128 + G.gtid;
        const bool live = i < A.n;
        const uint64_t game = A.game0 + (uint64_t)i;
        WarpStage W;
        W.init(A, (uint32_t)((tile * 4 + G.wq) & (int64_t)(A.n_seg - 1u)));  // = (first game of the warp / 32) % n_seg
        NfspFast g;
        g.unpack(live ? A.state[i] : 0ull);
        for (int s = 0; s < A.n_steps; ++s) {
            FastDecision d;
            TC_T0();
            fast_begin(g, s_lut, A, game, A.step0 + (uint64_t)s, live, d, c);
            TC_T1(tm_beg);
            float o0, o1, o2;
            const uint32_t owner = G.forward_sorted<false>(d.obs, g.p() * 2u + (uint32_t)d.pol, o0, o1, o2);
            TC_T1(tm_mma);
            G.sResult[owner] = make_float4(o0, o1, o2, 0.f);
            group_bar(1 + G.group);  // results are visible
            float4 v = G.sResult[G.gtid];
            TC_T1(tm_epi);
            if (d.random) { v.x = d.r0; v.y = d.r1; v.z = d.r2; }
            fast_finish<kDebug>(g, s_lut, A, W, d, v.x, v.y, v.z, live, (int64_t)s * A.n + i, plane, c);
            if ((s & 15) == 15) c.spill();
            TC_T1(tm_fin);
        }
        c.spill();  // the 5-bit counters must not run on into the group's next tile of games
        if (live) A.state[i] = g.pack();
        c.wide.trans += live ? A.n_steps : 0;
    }
#ifdef NFSP_TC_TIMING
    if (G.gtid == 0 && A.stats) {  // per-phase cycles of the group leaders, summed over groups and CTAs
        atomicAdd(A.stats + 13, (unsigned long long)tm_beg);
        atomicAdd(A.stats + 14, (unsigned long long)tm_mma);
        atomicAdd(A.stats + 15, (unsigned long long)tm_epi | ((unsigned long long)tm_fin << 32));
    }
#endif
    c.spill();
    tc_fence_before();
    __syncthreads();
    if (A.stats) c.wide.commit(s_stats, A.stats);
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

}  // namespace nfsp

using namespace nfsp;

extern "C" int nfsp_act_forward_tc(nfsp_env_t h, const uint32_t *d_obs, const int8_t *d_net, int64_t n, float *d_out,
                                   void *stream) {
    NFSP_CHECK_ARG(h != nullptr && d_obs && d_net && d_out && n >= 0, "bad arguments");
    if (!h->has_weights) return set_error(NFSP_E_STATE, "nfsp_act_set_weights has not been called");
    if (n == 0) return NFSP_OK;
    DeviceGuard guard(h->device);
    if (!guard.ok) return set_error(NFSP_E_CUDA, "cannot select device %d", h->device);
    cudaStream_t st = (cudaStream_t)stream;
    {
        const int rc = nfsp_ensure_tc_images(h, false, st);
        if (rc != NFSP_OK) return rc;
    }
    const int64_t ctas = (n + kFwdThreads - 1) / kFwdThreads;
    const int grid = (int)(ctas < h->sm_count ? ctas : h->sm_count);
    act_forward_tc_kernel<<<grid, kFwdThreads, kTcSmemBytes, st>>>((const uint8_t *)h->d_wtc_wide, d_obs, d_net, n, d_out);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

// called from nfsp_rollout (act_kernels.cu) when the tensor-core variant is selected
int nfsp_rollout_tc_launch(nfsp_env_t h, const nfsp::RolloutArgs &A0, bool debug, cudaStream_t st) {
    RolloutArgs A = A0;
    A.pack = h->d_wtc_wide;
    NFSP_CUDA(cudaMemsetAsync(h->d_work, 0, sizeof(uint32_t), st));
    const int64_t ctas = (A.n + kTcThreads - 1) / kTcThreads;
    const int grid = (int)(ctas < h->sm_count ? ctas : h->sm_count);
    if (debug) rollout_tc_kernel<true><<<grid, kTcThreads, kTcSmemBytes, st>>>(A);
    else rollout_tc_kernel<false><<<grid, kTcThreads, kTcSmemBytes, st>>>(A);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

// called from nfsp_act_set_weights (act_kernels.cu)
int nfsp_pack_tc_image(nfsp_env_t h, const float *d_weights, cudaStream_t st) {
    if (!h->d_wtc_wide) {
        NFSP_CUDA(cudaMalloc(&h->d_wtc_wide, kWideImageBytes));
        NFSP_CUDA(cudaFuncSetAttribute(act_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        NFSP_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        NFSP_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
    }
    pack_tc_wide_kernel<<<96, 256, 0, st>>>(d_weights, (uint8_t *)h->d_wtc_wide);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}
