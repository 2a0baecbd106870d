// K2, tensor-core variant of the first layer: tcgen05.mma with the accumulator in TMEM.
//
// The four acting nets are stacked along K: input k = net*32 + i (i < 30 observation bits, i = 30 the
// constant 1 that carries b1, i = 31 zero), so a row's net is selected by WHERE its 32 inputs sit and the
// result D[128 rows][64 hidden] needs a single tcgen05.ld per warp whatever mix of nets the tile holds.
// Inputs are exactly 0/1 in bf16; the fp32 weights are split exactly into three bf16 terms
// (w = hi + mid + lo), so 3 x 8 MMAs of M128 x N64 x K16 with fp32 accumulation reproduce the fp32 layer
// to accumulation-order rounding.  Layer 2 (64x3), the heads and everything else stay on CUDA cores.
//
// Operand images (K-major, no swizzle; core matrix = 8 rows x 16 bytes, see DESIGN.md "UMMA layout"):
//   byte offset(row, k) = (row/8)*2048 + (k/8)*128 + (row%8)*16 + (k%8)*2      (LBO = 128 B, SBO = 2048 B)
//   A: 128 rows  x 128 k = 32 KB, written by the threads (one row each)
//   B: 3 splits x (64 n x 128 k) = 3 x 16 KB, built once per weight update, brought in by a bulk copy
#include <cuda_bf16.h>

#include "mlp_math.cuh"
#include "ptx_helpers.cuh"
#include "rollout_fast.cuh"

namespace nfsp {

constexpr int kTcThreads = 128;
constexpr int kTcK = 128;                       // 4 nets x 32 inputs
constexpr int kTcABytes = 128 * kTcK * 2;       // 32 KB
constexpr int kTcBSplitBytes = 64 * kTcK * 2;   // 16 KB
constexpr int kTcBBytes = 3 * kTcBSplitBytes;   // 48 KB
constexpr int kTcW2Floats = 16 * 4 * 3 * 4 + 16;  // W2 as [16 quads][4 nets][3 outputs][4] + b2 [4][4]
constexpr int kTcImageBytes = kTcBBytes + kTcW2Floats * 4;
constexpr int kTcSmemBytes = kTcABytes + kTcImageBytes + 64;  // + mbarriers / tmem slot
constexpr uint32_t kLBO = 128, kSBO = 2048;

// ---- weight image ---------------------------------------------------------------------------------
__global__ void pack_tc_kernel(const float *__restrict__ w, uint8_t *__restrict__ img) {
    const int total = 3 * 64 * kTcK;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int split = e / (64 * kTcK), n = (e / kTcK) % 64, k = e % kTcK;
        const int net = k >> 5, i = k & 31;
        float v = 0.f;
        if (i < 30) v = w[net * NFSP_NET_PARAMS + i * 64 + n];
        else if (i == 30) v = w[net * NFSP_NET_PARAMS + 1920 + n];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const float r1 = v - __bfloat162float(hi);
        const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
        const float r2 = r1 - __bfloat162float(mid);
        const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
        const __nv_bfloat16 pick = split == 0 ? hi : (split == 1 ? mid : lo);
        const size_t off = (size_t)split * kTcBSplitBytes + (n >> 3) * kSBO + (k >> 3) * kLBO + (n & 7) * 16 + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(img + off) = pick;
    }
    float *w2 = reinterpret_cast<float *>(img + kTcBBytes);
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
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(kLBO >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor: D f32, A/B bf16, both K-major, M = 128, N = 64
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(a), "l"(b), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive columns: thread t of warp w reads TMEM lane 32*(w%4)+t
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
        "%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
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

// One CTA-wide first layer: every thread has written its A row; returns with h[64] = pre-activations of the
// thread's row.  Collective over the CTA (128 threads).  `phase` is the mbarrier parity of this tile.
struct TcTile {
    uint32_t a_smem, b_smem, bar, tmem;
};

// `bar_id`: named barrier of the 128 threads that share this tile (0 = the whole 128-thread CTA)
__device__ __forceinline__ void tc_layer1(const TcTile &t, uint32_t phase, float *h, uint32_t bar_id = 0) {
    fence_async_smem();  // generic-proxy writes of A -> visible to the tensor core (async proxy)
    tc_fence_before();   // earlier tcgen05.ld of the accumulator is ordered before the barrier
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    if ((threadIdx.x & 127u) == 0) {
        tc_fence_after();
#pragma unroll
        for (int split = 0; split < 3; ++split) {
#pragma unroll
            for (int s = 0; s < kTcK / 16; ++s) {
                const uint64_t da = umma_desc(t.a_smem + s * 2 * kLBO);
                const uint64_t db = umma_desc(t.b_smem + split * kTcBSplitBytes + s * 2 * kLBO);
                umma_f16(t.tmem, da, db, (split | s) != 0);
            }
        }
        umma_commit(t.bar);
    }
    mbar_wait(t.bar, phase);
    tc_fence_after();
    const uint32_t lane_base = t.tmem + (((threadIdx.x >> 5) & 3u) << 21);  // lane 32*(warp%4) in bits 31..16
    tmem_ld32(lane_base, h);
    tmem_ld32(lane_base + 32, h + 32);
}

__device__ __forceinline__ void layer2_head(const float *__restrict__ w2img, const float *h, int net, float out[3]) {
    const float4 *w2 = reinterpret_cast<const float4 *>(w2img) + net * 3;
    Layer2Acc acc;
#pragma unroll
    for (int q = 0; q < 16; ++q)
        acc.quad(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3], w2[q * 12], w2[q * 12 + 1], w2[q * 12 + 2]);
    acc.head(reinterpret_cast<const float4 *>(w2img + 16 * 4 * 3 * 4)[net], net & 1, out[0], out[1], out[2]);
}

// writes the thread's operand row: the 4 chunks of `net` hold obs|bias, the 4 chunks of the net the row held
// before are cleared (all other chunks are already zero)
__device__ __forceinline__ void write_a_row(uint8_t *a_row, uint32_t obs, int net, int &prev_net) {
    if (prev_net >= 0 && prev_net != net) {
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4 *>(a_row + (prev_net * 4 + c) * kLBO) = make_uint4(0, 0, 0, 0);
    }
    const uint32_t x = (obs & 0x3FFFFFFFu) | (1u << 30);
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4 *>(a_row + (net * 4 + c) * kLBO) = bits_to_bf16x8((x >> (8 * c)) & 0xFFu);
    prev_net = net;
}

__global__ void __launch_bounds__(kTcThreads)
act_forward_tc_kernel(const uint8_t *__restrict__ img, const uint32_t *__restrict__ obs, const int8_t *__restrict__ net,
                      int64_t n, float *__restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sA = smem;
    uint8_t *sB = smem + kTcABytes;
    const float *sW2 = reinterpret_cast<const float *>(sB + kTcBBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kTcABytes + kTcImageBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const uint32_t bar_w = smem_u32(bars), bar_m = smem_u32(bars + 1);

    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_m, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = threadIdx.x; e < kTcABytes / 16; e += blockDim.x) reinterpret_cast<uint4 *>(sA)[e] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (threadIdx.x == 0) {  // weight image: one bulk async copy (TMA unit), completion on an mbarrier
        mbar_expect_tx(bar_w, kTcImageBytes);
        bulk_g2s(smem_u32(sB), img, kTcImageBytes, bar_w);
    }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(tmem_slot), 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    TcTile t;
    t.a_smem = smem_u32(sA); t.b_smem = smem_u32(sB); t.bar = bar_m; t.tmem = *tmem_slot;
    mbar_wait(bar_w, 0);

    uint8_t *a_row = sA + (threadIdx.x >> 3) * kSBO + (threadIdx.x & 7) * 16;
    int prev_net = -1;
    uint32_t phase = 0;
    const int64_t tiles = (n + kTcThreads - 1) / kTcThreads;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t i = tile * kTcThreads + threadIdx.x;
        const bool live = i < n;
        const int k = live ? (int)(net[i] & 3) : 0;
        write_a_row(a_row, live ? obs[i] : 0u, k, prev_net);
        float h[64];
        tc_layer1(t, phase, h);
        phase ^= 1u;
        float v[3];
        layer2_head(sW2, h, k, v);
        if (live) { out[3 * i] = v[0]; out[3 * i + 1] = v[1]; out[3 * i + 2] = v[2]; }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(t.tmem, 64);
}

// ---- fused rollout, tcgen05 first layer ---------------------------------------------------------------
// Shape chosen from profiles/r01/umma_microbench.txt (B200): a tcgen05.mma costs ~90 cycles whatever N <= 128
// and ~150 cycles at N = 256, plus ~450 cycles issue->commit latency per chain, so the tile that minimises
// tensor time is the WIDE one: D[128 rows][256 = 4 nets x 64 hidden] = A[128][32] x B[256][32]^T, 2 k-steps
// x 3 bf16 splits = 6 MMAs per 128 decisions.  A row needs only its own net's 64 columns, so the rows of a
// tile are SORTED BY NET (counting sort over the group's 128 threads) before they are written as operand rows:
// warps then read one 64-column block of TMEM (two at a segment boundary) and their W2 reads are warp-uniform.
//
// One persistent CTA per SM, kGroups groups of 128 threads.  A group takes tiles of 128 consecutive games from a
// global counter; thread t owns game t of the tile (actor-relative state in registers over the launch's steps,
// nfsp_fast.cuh) and is the epilogue worker of sorted row t.  The 512 TMEM columns are two accumulator slots;
// a group uses slot (group & 1) under a shared-memory lock taken by its MMA issuer and released by the last
// of its four warps to finish reading the accumulator.
constexpr int kGroups = 6;
constexpr int kRtcThreads = 128 * kGroups;
constexpr int kWideABytes = 128 * 32 * 2;          // 8 KB operand tile per group
constexpr int kWideBSplitBytes = 256 * 32 * 2;     // 16 KB
constexpr int kWideBBytes = 3 * kWideBSplitBytes;  // 48 KB
constexpr int kWideImageBytes = kWideBBytes + kTcW2Floats * 4;
constexpr uint32_t kWideSBO = 512;                 // 4 k-chunks of 128 B per 8-row group
constexpr uint32_t kIdescWide = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
// dynamic smem: A tiles | weight image | results float4[kGroups][128] | byte -> 8 x bf16 table | owner u8[kGroups][128]
//               | counts | tile ids | barriers | slot locks
constexpr int kRtcOffImage = kGroups * kWideABytes;
constexpr int kRtcOffResult = kRtcOffImage + kWideImageBytes;
constexpr int kRtcOffBits = kRtcOffResult + kGroups * 128 * 16;
constexpr int kRtcOffOwner = kRtcOffBits + 256 * 16;
constexpr int kRtcOffCnt = kRtcOffOwner + kGroups * 128;
constexpr int kRtcOffTile = kRtcOffCnt + kGroups * 4 * 4;
constexpr int kRtcOffBars = kRtcOffTile + kGroups * 4 + 8;
constexpr int kRtcOffLocks = kRtcOffBars + 8 * (1 + kGroups);
constexpr int kRtcSmemBytes = kRtcOffLocks + 4 * 4 + 16;

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

__device__ __forceinline__ uint64_t umma_desc_wide(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(kLBO >> 4) << 16) | ((uint64_t)(kWideSBO >> 4) << 32) |
           (1ull << 46);
}
__device__ __forceinline__ void umma_f16_wide(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(a), "l"(b), "r"(kIdescWide), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void group_bar(uint32_t id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <bool kDebug>
__global__ void __launch_bounds__(kRtcThreads, 1)
rollout_tc_kernel(const RolloutArgs A) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ unsigned long long s_stats[NFSP_STATS_FIELDS];
    __shared__ FastLuts s_lut;
    const uint32_t group = threadIdx.x >> 7, gtid = threadIdx.x & 127u, wq = gtid >> 5, lane = gtid & 31u;
    uint8_t *sA = smem + group * kWideABytes;
    uint8_t *sB = smem + kRtcOffImage;
    const float *sW2 = reinterpret_cast<const float *>(sB + kWideBBytes);
    float4 *sResult = reinterpret_cast<float4 *>(smem + kRtcOffResult) + group * 128;
    const uint4 *sBits = reinterpret_cast<const uint4 *>(smem + kRtcOffBits);
    uint8_t *sOwner = smem + kRtcOffOwner + group * 128;
    uint32_t *sCnt = reinterpret_cast<uint32_t *>(smem + kRtcOffCnt) + group * 4;
    volatile uint32_t *sTile = reinterpret_cast<volatile uint32_t *>(smem + kRtcOffTile) + group;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kRtcOffBars);
    uint32_t *locks = reinterpret_cast<uint32_t *>(smem + kRtcOffLocks);  // [0..1] slot busy, [2..3] warps done reading
    uint32_t *tmem_slot = locks + 4;
    const uint32_t bar_w = smem_u32(bars), bar_done = smem_u32(bars + 1 + group);
    uint32_t *slot_busy = locks + (group & 1u), *slot_readers = locks + 2 + (group & 1u);

    if (threadIdx.x < NFSP_STATS_FIELDS) s_stats[threadIdx.x] = 0ull;
    s_lut.fill();
    for (uint32_t e = threadIdx.x; e < 256u; e += blockDim.x) reinterpret_cast<uint4 *>(smem + kRtcOffBits)[e] = bits_to_bf16x8(e);
    if (threadIdx.x < 4) locks[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        for (int k = 0; k < 1 + kGroups; ++k) mbar_init(smem_u32(bars + k), 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar_w, kWideImageBytes);
        bulk_g2s(smem_u32(sB), A.pack, kWideImageBytes, bar_w);
    }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_acc = tmem_base + (group & 1u) * 256u + (wq << 21);  // this group's slot, this warp's lanes
    const uint32_t a_smem = smem_u32(sA), b_smem = smem_u32(sB);
    mbar_wait(bar_w, 0);

    uint32_t ph_done = 0;
    FastCounters c;
#ifdef NFSP_TC_TIMING
    long long tm_lock = 0, tm_mma = 0, tm_epi = 0, tm_fin = 0, tm_beg = 0, tm_sort = 0; long long t0, t1;
#define TC_T0() t0 = clock64()
#define TC_T1(acc) do { t1 = clock64(); acc += t1 - t0; t0 = t1; } while (0)
#else
#define TC_T0()
#define TC_T1(acc)
#endif
    const int64_t plane = (int64_t)A.n_steps * A.n;
    const int64_t tiles = (A.n + 127) >> 7;
    for (;;) {
        if (gtid == 0) *sTile = atomicAdd(A.work, 1u);  // tiles of 128 games are handed out dynamically
        group_bar(1 + group);
        const int64_t tile = (int64_t)*sTile;
        if (tile >= tiles) break;
        const int64_t i = tile * 128 + gtid;
        const bool live = i < A.n;
        const uint64_t game = A.game0 + (uint64_t)i;
        WarpStage W;
        W.init(A, (uint32_t)((tile * 4 + wq) & (int64_t)(A.n_seg - 1u)));  // = (first game of the warp / 32) % n_seg
        NfspFast g;
        g.unpack(live ? A.state[i] : 0ull);
        for (int s = 0; s < A.n_steps; ++s) {
            FastDecision d;
            TC_T0();
            fast_begin(g, s_lut, A, game, A.step0 + (uint64_t)s, live, d, c);
            TC_T1(tm_beg);
            const uint32_t net = g.p() * 2u + (uint32_t)d.pol;
            // ---- counting sort of the group's rows by net: packed byte counters, one word per warp.  Sort key
            // order avg0, br0, br1, avg1 keeps the two small best-response segments adjacent, so fewer warps
            // straddle a segment boundary.
            const uint32_t key = net ^ (net >> 1);  // net 0,1,2,3 -> key 0,1,3,2
            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, key == 0), m1 = __ballot_sync(0xFFFFFFFFu, key == 1);
            const uint32_t m2 = __ballot_sync(0xFFFFFFFFu, key == 2), m3 = ~(m0 | m1 | m2);
            const uint32_t mine = key == 0 ? m0 : (key == 1 ? m1 : (key == 2 ? m2 : m3));
            const uint32_t rank = __popc(mine & ((1u << lane) - 1u));
            if (lane == 0) sCnt[wq] = __popc(m0) | (__popc(m1) << 8) | (__popc(m2) << 16) | (__popc(m3) << 24);
            group_bar(1 + group);
            const uint32_t c0 = sCnt[0], c1 = sCnt[1], c2 = sCnt[2], c3 = sCnt[3];
            const uint32_t tot = c0 + c1 + c2 + c3;  // bytes: rows of key 0..3 (<= 128 each, no carry)
            const uint32_t before = (wq > 0 ? c0 : 0u) + (wq > 1 ? c1 : 0u) + (wq > 2 ? c2 : 0u);
            const uint32_t seg1 = tot & 0xFFu, seg2 = seg1 + ((tot >> 8) & 0xFFu), seg3 = seg2 + ((tot >> 16) & 0xFFu);
            const uint32_t seg_start = key == 0 ? 0u : (key == 1 ? seg1 : (key == 2 ? seg2 : seg3));
            const uint32_t pos = seg_start + ((before >> (8 * key)) & 0xFFu) + rank;
            {  // operand row `pos`: observation bits + the constant 1 that carries b1, 8 inputs per table lookup
                uint8_t *row = sA + (pos >> 3) * kWideSBO + (pos & 7u) * 16;
                const uint32_t x = (d.obs & 0x3FFFFFFFu) | (1u << 30);
#pragma unroll
                for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4 *>(row + k * kLBO) = sBits[(x >> (8 * k)) & 0xFFu];
                sOwner[pos] = (uint8_t)gtid;
            }
            fence_async_smem();
            group_bar(1 + group);
            TC_T1(tm_sort);
            if (gtid == 0) {
                while (atomicCAS(slot_busy, 0u, 1u) != 0u) __nanosleep(32);  // the accumulator slot is ours
                tc_fence_after();
                TC_T1(tm_lock);
#pragma unroll
                for (int split = 0; split < 3; ++split)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        umma_f16_wide(tmem_base + (group & 1u) * 256u, umma_desc_wide(a_smem + ks * 2 * kLBO),
                                      umma_desc_wide(b_smem + split * kWideBSplitBytes + ks * 2 * kLBO), (split | ks) != 0);
                umma_commit(bar_done);
            }
            mbar_wait_backoff(bar_done, ph_done);  // every thread waits on the mbarrier itself (measured faster than one
            ph_done ^= 1u;                         // polling warp + a group barrier: 0.67 vs 0.70 ms per launch)
            tc_fence_after();
            TC_T1(tm_mma);
            // ---- epilogue of sorted row `gtid`: its net's 64 pre-activations -> layer 2 -> head
            const uint32_t r = gtid;
            const uint32_t my_key = (r >= seg1) + (r >= seg2) + (r >= seg3);
            const uint32_t my_net = my_key ^ (my_key >> 1);  // inverse of the key map
            const uint32_t r_lo = wq * 32u, r_hi = r_lo + 31u;
            const uint32_t k_lo = (r_lo >= seg1) + (r_lo >= seg2) + (r_lo >= seg3);
            const uint32_t k_hi = (r_hi >= seg1) + (r_hi >= seg2) + (r_hi >= seg3);
            const float4 *w2 = reinterpret_cast<const float4 *>(sW2) + my_net * 3;
            Layer2Acc acc;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float h[16];
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
                if (ch == 3) {  // the warp has its last accumulator columns in registers: the 4th warp frees the slot
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0 && atomicAdd(slot_readers, 1u) == 3u) {
                        *reinterpret_cast<volatile uint32_t *>(slot_readers) = 0u;
                        __threadfence_block();
                        atomicExch(slot_busy, 0u);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    acc.quad(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3], w2[(ch * 4 + q) * 12],
                             w2[(ch * 4 + q) * 12 + 1], w2[(ch * 4 + q) * 12 + 2]);
            }
            float o0, o1, o2;
            acc.head(reinterpret_cast<const float4 *>(sW2 + 16 * 4 * 3 * 4)[my_net], my_net & 1u, o0, o1, o2);
            sResult[sOwner[r]] = make_float4(o0, o1, o2, 0.f);
            group_bar(1 + group);  // results are visible; A tile, owner map and counters may be rewritten
            float4 v = sResult[gtid];
            TC_T1(tm_epi);
            if (d.random) { v.x = d.r0; v.y = d.r1; v.z = d.r2; }
            fast_finish<kDebug>(g, s_lut, A, W, d, v.x, v.y, v.z, live, (int64_t)s * A.n + i, plane, c);
            if ((s & 15) == 15) c.spill();
            TC_T1(tm_fin);
        }
        if (live) A.state[i] = g.pack();
        c.wide.trans += live ? A.n_steps : 0;
    }
#ifdef NFSP_TC_TIMING
    if (gtid == 0 && A.stats) {  // per-phase cycles of the group leaders, summed over groups and CTAs
        atomicAdd(A.stats + 13, (unsigned long long)(tm_beg + tm_sort));
        atomicAdd(A.stats + 14, (unsigned long long)(tm_lock) | ((unsigned long long)tm_mma << 32));
        atomicAdd(A.stats + 15, (unsigned long long)(tm_epi) | ((unsigned long long)tm_fin << 32));
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
    if (!h->has_weights || !h->d_wtc) return set_error(NFSP_E_STATE, "nfsp_act_set_weights has not been called");
    if (n == 0) return NFSP_OK;
    DeviceGuard guard(h->device);
    if (!guard.ok) return set_error(NFSP_E_CUDA, "cannot select device %d", h->device);
    const int64_t tiles = (n + kTcThreads - 1) / kTcThreads;
    const int64_t full = (int64_t)h->sm_count * 2;
    const int grid = (int)(tiles < full ? tiles : full);
    act_forward_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, (cudaStream_t)stream>>>((const uint8_t *)h->d_wtc, d_obs, d_net,
                                                                                    n, d_out);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

// called from nfsp_rollout (act_kernels.cu) when the tensor-core variant is selected
int nfsp_rollout_tc_launch(nfsp_env_t h, const nfsp::RolloutArgs &A0, bool debug, cudaStream_t st) {
    RolloutArgs A = A0;
    A.pack = h->d_wtc_wide;
    NFSP_CUDA(cudaMemsetAsync(h->d_work, 0, sizeof(uint32_t), st));
    const int64_t ctas = (A.n + kRtcThreads - 1) / kRtcThreads;
    const int grid = (int)(ctas < h->sm_count ? ctas : h->sm_count);
    if (debug) rollout_tc_kernel<true><<<grid, kRtcThreads, kRtcSmemBytes, st>>>(A);
    else rollout_tc_kernel<false><<<grid, kRtcThreads, kRtcSmemBytes, st>>>(A);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

// called from nfsp_act_set_weights (act_kernels.cu)
int nfsp_pack_tc_image(nfsp_env_t h, const float *d_weights, cudaStream_t st) {
    if (!h->d_wtc) {
        NFSP_CUDA(cudaMalloc(&h->d_wtc, kTcImageBytes));
        NFSP_CUDA(cudaMalloc(&h->d_wtc_wide, kWideImageBytes));
        NFSP_CUDA(cudaFuncSetAttribute(act_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        NFSP_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRtcSmemBytes));
        NFSP_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRtcSmemBytes));
    }
    pack_tc_kernel<<<96, 256, 0, st>>>(d_weights, (uint8_t *)h->d_wtc);
    NFSP_LAUNCH_CHECK();
    pack_tc_wide_kernel<<<96, 256, 0, st>>>(d_weights, (uint8_t *)h->d_wtc_wide);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}
