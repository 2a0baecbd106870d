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

#include "rollout_common.cuh"

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

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
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
    float z0 = 0.f, z1 = 0.f, z2 = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const float hx = fmaxf(h[4 * q], 0.f), hy = fmaxf(h[4 * q + 1], 0.f);
        const float hz = fmaxf(h[4 * q + 2], 0.f), hw = fmaxf(h[4 * q + 3], 0.f);
        const float4 u0 = w2[q * 12], u1 = w2[q * 12 + 1], u2 = w2[q * 12 + 2];
        z0 = fmaf(hx, u0.x, z0); z0 = fmaf(hy, u0.y, z0); z0 = fmaf(hz, u0.z, z0); z0 = fmaf(hw, u0.w, z0);
        z1 = fmaf(hx, u1.x, z1); z1 = fmaf(hy, u1.y, z1); z1 = fmaf(hz, u1.z, z1); z1 = fmaf(hw, u1.w, z1);
        z2 = fmaf(hx, u2.x, z2); z2 = fmaf(hy, u2.y, z2); z2 = fmaf(hz, u2.z, z2); z2 = fmaf(hw, u2.w, z2);
    }
    const float4 b2 = reinterpret_cast<const float4 *>(w2img + 16 * 4 * 3 * 4)[net];
    z0 += b2.x; z1 += b2.y; z2 += b2.z;
    if (net & 1) {
        out[0] = fmaxf(z0, 0.f); out[1] = fmaxf(z1, 0.f); out[2] = fmaxf(z2, 0.f);
    } else {
        const float m = fmaxf(z0, fmaxf(z1, z2));
        const float e0 = expf(z0 - m), e1 = expf(z1 - m), e2 = expf(z2 - m);
        const float inv = 1.0f / (e0 + e1 + e2);
        out[0] = e0 * inv; out[1] = e1 * inv; out[2] = e2 * inv;
    }
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
// One persistent CTA per SM, kGroups independent groups of 128 threads.  A group owns 128 games at a time
// (thread = game = operand row = TMEM lane), its own A tile, 64 TMEM columns and an mbarrier; the groups
// share one weight image and synchronise only among themselves (named barriers), so while one group waits
// for its MMAs the others run their CUDA-core work (env step, layer 2, record emission).
constexpr int kGroups = 4;
constexpr int kRtcThreads = 128 * kGroups;
constexpr int kRtcSmemBytes = kGroups * kTcABytes + kTcImageBytes + 128;

template <bool kDebug>
__global__ void __launch_bounds__(kRtcThreads, 1)
rollout_tc_kernel(const RolloutArgs A) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ unsigned long long s_stats[NFSP_STATS_FIELDS];
    uint8_t *sA = smem;
    uint8_t *sB = smem + kGroups * kTcABytes;
    const float *sW2 = reinterpret_cast<const float *>(sB + kTcBBytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + kTcImageBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 1 + kGroups);
    const uint32_t bar_w = smem_u32(bars);
    const uint32_t group = threadIdx.x >> 7, gtid = threadIdx.x & 127u;

    if (threadIdx.x < NFSP_STATS_FIELDS) s_stats[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        for (int k = 0; k < kGroups; ++k) mbar_init(smem_u32(bars + 1 + k), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = threadIdx.x; e < kGroups * kTcABytes / 16; e += blockDim.x)
        reinterpret_cast<uint4 *>(sA)[e] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar_w, kTcImageBytes);
        bulk_g2s(smem_u32(sB), A.pack, kTcImageBytes, bar_w);
    }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(tmem_slot), 64 * kGroups);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    TcTile t;
    t.a_smem = smem_u32(sA) + group * kTcABytes;
    t.b_smem = smem_u32(sB);
    t.bar = smem_u32(bars + 1 + group);
    t.tmem = *tmem_slot + group * 64;
    const uint32_t tmem_base = *tmem_slot;
    mbar_wait(bar_w, 0);

    uint8_t *a_row = sA + group * kTcABytes + (gtid >> 3) * kSBO + (gtid & 7) * 16;
    int prev_net = -1;
    uint32_t phase = 0;
    Counters c;
    const int64_t plane = (int64_t)A.n_steps * A.n;
    const int64_t tiles = (A.n + 127) / 128;
    for (int64_t tile = (int64_t)blockIdx.x * kGroups + group; tile < tiles; tile += (int64_t)gridDim.x * kGroups) {
        const int64_t i = tile * 128 + gtid;
        const bool live = i < A.n;
        const uint64_t game = A.game0 + (uint64_t)i;
        NfspW g{live ? A.state[i] : 0ull};
        for (int s = 0; s < A.n_steps; ++s) {
            Decision d;
            if (live) decide_begin(g, A, game, A.step0 + (uint64_t)s, d, c);
            const int net = live ? d.p * 2 + (int)d.pol : 0;
            write_a_row(a_row, d.obs, net, prev_net);
            float h[64];
            tc_layer1(t, phase, h, 1 + group);
            phase ^= 1u;
            float v[3];
            layer2_head(sW2, h, net, v);
            if (d.random) { v[0] = d.v0; v[1] = d.v1; v[2] = d.v2; }
            decide_finish<kDebug>(g, A, d, v[0], v[1], v[2], live, (int64_t)s * A.n + i, plane, c, s_stats);
        }
        if (live) A.state[i] = g.w;
    }
    tc_fence_before();
    __syncthreads();
    if (A.stats) c.commit(s_stats, A.stats);
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 64 * kGroups);
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
    A.pack = h->d_wtc;
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
        NFSP_CUDA(cudaFuncSetAttribute(act_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        NFSP_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRtcSmemBytes));
        NFSP_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRtcSmemBytes));
    }
    pack_tc_kernel<<<96, 256, 0, st>>>(d_weights, (uint8_t *)h->d_wtc);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}
