// K2 (CUDA-core variant): NFSP acting fused with the env step.
//
// One thread per game.  Per step a game does exactly one Agent.play decision (agent.py:130-156):
// observation -> one of four 30-64-3 MLPs (player x {average, best-response}) -> eta-mixed /
// epsilon-greedy score vector -> argmax -> newenv step, and emits the RL / SL memory records of
// agent.py:134-136,151 and the terminal observations of main.py:55-67.  The observation is a
// 30-bit mask that never leaves registers; because it is binary and sparse (<= 9 bits set) the
// first layer is a SUM OF WEIGHT ROWS read from shared memory as float4s instead of a dense 30x64
// product: <= 9 rows for arbitrary masks (act_forward_kernel), 3 precombined rows in the rollout.
#include "mlp_math.cuh"
#include "rollout_common.cuh"

namespace nfsp {

constexpr int kActThreads = 128;   // act_forward_kernel (generic observation masks)
constexpr int kRollThreads = 256;  // rollout_kernel
// packed weight image (floats): W1 rows as [16 col-quads][128 rows = net*32 + input][4], where input 30 is
// the bias b1; then W2 as [64 hidden][4 nets][4 = 3 outputs + pad]; then b2 as [4 nets][4].
constexpr int kW1Floats = 16 * 128 * 4;
constexpr int kW2Floats = 64 * 4 * 4;
constexpr int kB2Floats = 4 * 4;
constexpr int kPackFloats = kW1Floats + kW2Floats + kB2Floats;

__global__ void pack_weights_kernel(const float *__restrict__ w, float *__restrict__ pack) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < kPackFloats; e += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (e < kW1Floats) {
            const int c = e & 3, row = (e >> 2) & 127, q = e >> 9;
            const int net = row >> 5, i = row & 31, j = q * 4 + c;
            const float *wn = w + net * NFSP_NET_PARAMS;
            if (i < 30) v = wn[i * 64 + j];
            else if (i == 30) v = wn[1920 + j];
        } else if (e < kW1Floats + kW2Floats) {
            const int f = e - kW1Floats, c = f & 3, net = (f >> 2) & 3, j = f >> 4;
            if (c < 3) v = w[net * NFSP_NET_PARAMS + 1984 + j * 3 + c];
        } else {
            const int f = e - kW1Floats - kW2Floats, c = f & 3, net = f >> 2;
            if (c < 3) v = w[net * NFSP_NET_PARAMS + 2176 + c];
        }
        pack[e] = v;
    }
}

// ---- group-factorised first layer (rollout path) --------------------------------------------------
// Under main.train's turn order an observation is (cards of the actor, round-0 betting sequence,
// round-1 betting sequence); each group takes few values, so W1^T x + b1 is the sum of THREE
// precombined rows: T_cards[c_p][public state] (+ b1), T_r0[dealer][sequence], T_r1[dealer][sequence].
// The rows are sums of W1 rows only (input independent), rebuilt whenever the weights change.
//   per net 48 rows: 0-11 cards (c_p*4 + (revealed ? 1+pub : 0)), 12-29 round 0, 30-47 round 1
//   (dealer*9 + sequence id: 0 -, 1 C, 2 R, 3 CC, 4 CR, 5 RC, 6 RR, 7 CRC, 8 RRC).
// Image: rows as [16 col-quads][192 rows][4] floats, then W2 as [16 quads][4 nets][3 outputs][4], then b2.
constexpr int kTabRows = 4 * 48;
constexpr int kTabFloats = 16 * kTabRows * 4;
constexpr int kTabW2Floats = 16 * 4 * 3 * 4;
constexpr int kTabImageFloats = kTabFloats + kTabW2Floats + kB2Floats;
constexpr int kTabImageBytes = kTabImageFloats * 4;

__global__ void pack_tables_kernel(const float *__restrict__ w, float *__restrict__ img) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < kTabImageFloats; e += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (e < kTabFloats) {
            const int c = e & 3, row = (e >> 2) % kTabRows, q = (e >> 2) / kTabRows;
            const int net = row / 48, r = row % 48, j = q * 4 + c;
            const float *W1 = w + net * NFSP_NET_PARAMS;  // W1[i*64 + j]
            uint32_t bits = 0;                            // observation bits this row stands for
            if (r < 12) {
                const int cp = r >> 2, ps = r & 3;
                if (cp < 3) {
                    bits = 1u << (24 + cp);
                    if (ps) bits |= (1u << (27 + cp)) | (1u << (27 + (ps - 1)));
                }
            } else {
                const int rr = (r - 12) / 18, d = ((r - 12) % 18) / 9, id = (r - 12) % 9;
                // sequence id -> actions of slots 0..2 (1 = call, 2 = raise, 0 = none)
                const int s0 = id == 0 ? 0 : (id == 1 || id == 3 || id == 4 || id == 7 ? 1 : 2);
                const int s1 = id < 3 ? 0 : (id == 3 || id == 5 ? 1 : 2);
                const int s2 = id >= 7 ? 1 : 0;
                const int sl[3] = {s0, s1, s2};
                for (int k = 0; k < 3; ++k)
                    if (sl[k]) bits |= 1u << (((k & 1) ^ d) * 12 + rr * 6 + k * 2 + (sl[k] - 1));
            }
            for (int i = 0; i < 30; ++i)
                if ((bits >> i) & 1u) v += W1[i * 64 + j];
            if (r < 12 && (r >> 2) < 3) v += W1[1920 + j];  // b1 rides on the card row
        } else if (e < kTabFloats + kTabW2Floats) {
            const int f = e - kTabFloats, x = f & 3, c = (f >> 2) % 3, net = ((f >> 2) / 3) & 3, q = (f >> 2) / 12;
            v = w[net * NFSP_NET_PARAMS + 1984 + (q * 4 + x) * 3 + c];
        } else {
            const int f = e - kTabFloats - kTabW2Floats, c = f & 3, net = f >> 2;
            if (c < 3) v = w[net * NFSP_NET_PARAMS + 2176 + c];
        }
        img[e] = v;
    }
}

__device__ __forceinline__ uint32_t seq_id(uint32_t rnd) {  // 6 slot bits of one round -> sequence id 0..8
    const uint32_t s0 = rnd & 3u, s1 = (rnd >> 2) & 3u, s2 = (rnd >> 4) & 3u;
    return s2 ? 6u + s0 : (s1 ? 2u * s0 + s1 : s0);
}

// one decision of net `net` on the game word: layer 1 as 3 row reads streamed into layer 2
__device__ __forceinline__ void mlp_forward_tables(const float *__restrict__ st, uint32_t hist, uint32_t cp,
                                                   uint32_t pubstate, uint32_t dealer, int net, float out[3]) {
    const float4 *T = reinterpret_cast<const float4 *>(st);
    const uint32_t both = hist | (hist >> 12);
    const float4 *rc = T + net * 48 + cp * 4 + pubstate;
    const float4 *r0 = T + net * 48 + 12 + dealer * 9 + seq_id(both & 63u);
    const float4 *r1 = T + net * 48 + 30 + dealer * 9 + seq_id((both >> 6) & 63u);
    const float4 *w2 = reinterpret_cast<const float4 *>(st + kTabFloats) + net * 3;
    Layer2Acc acc;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const float4 a = rc[q * kTabRows], b = r0[q * kTabRows], c = r1[q * kTabRows];
        acc.quad((b.x + c.x) + a.x, (b.y + c.y) + a.y, (b.z + c.z) + a.z, (b.w + c.w) + a.w, w2[q * 12], w2[q * 12 + 1],
                 w2[q * 12 + 2]);
    }
    acc.head(reinterpret_cast<const float4 *>(st + kTabFloats + kTabW2Floats)[net], net & 1, out[0], out[1], out[2]);
}

// forward of one net on one observation mask; sw = packed image in shared memory
__device__ __forceinline__ void mlp_forward(const float *__restrict__ sw, uint32_t obs, int net, float out[3]) {
    const float4 *w1 = reinterpret_cast<const float4 *>(sw);
    const int row0 = net * 32;
    float h[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) h[j] = 0.f;
    uint32_t bits = (obs & 0x3FFFFFFFu) | (1u << 30);  // input 30 = constant 1 -> adds the bias row last
    while (bits) {
        const int i = __ffs(bits) - 1;
        bits &= bits - 1u;
        const float4 *r = w1 + row0 + i;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 v = r[q * 128];
            h[4 * q + 0] += v.x; h[4 * q + 1] += v.y; h[4 * q + 2] += v.z; h[4 * q + 3] += v.w;
        }
    }
    const float4 *w2 = reinterpret_cast<const float4 *>(sw + kW1Floats) + net;
    float z0 = 0.f, z1 = 0.f, z2 = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        const float hj = fmaxf(h[j], 0.f);  // Dense(64, relu), agent.py:102,111
        const float4 v = w2[j * 4];
        z0 = fmaf(hj, v.x, z0); z1 = fmaf(hj, v.y, z1); z2 = fmaf(hj, v.z, z2);
    }
    const float4 b2 = reinterpret_cast<const float4 *>(sw + kW1Floats + kW2Floats)[net];
    z0 += b2.x; z1 += b2.y; z2 += b2.z;
    if (net & 1) {  // best-response net: Dense(3, relu), agent.py:103
        out[0] = fmaxf(z0, 0.f); out[1] = fmaxf(z1, 0.f); out[2] = fmaxf(z2, 0.f);
    } else {        // average-policy net: Dense(3, softmax), agent.py:112
        const float m = fmaxf(z0, fmaxf(z1, z2));
        const float e0 = expf(z0 - m), e1 = expf(z1 - m), e2 = expf(z2 - m);
        const float inv = 1.0f / (e0 + e1 + e2);
        out[0] = e0 * inv; out[1] = e1 * inv; out[2] = e2 * inv;
    }
}

__device__ __forceinline__ void load_pack(float *sw, const float *__restrict__ pack) {
    const float4 *src = reinterpret_cast<const float4 *>(pack);
    float4 *dst = reinterpret_cast<float4 *>(sw);
    for (int e = threadIdx.x; e < kPackFloats / 4; e += blockDim.x) dst[e] = src[e];
    __syncthreads();
}

__global__ void __launch_bounds__(kActThreads)
act_forward_kernel(const float *__restrict__ pack, const uint32_t *__restrict__ obs, const int8_t *__restrict__ net,
                   int64_t n, float *__restrict__ out) {
    __shared__ __align__(16) float sw[kPackFloats];
    load_pack(sw, pack);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v[3];
        mlp_forward(sw, obs[i], (int)(net[i] & 3), v);
        out[3 * i] = v[0]; out[3 * i + 1] = v[1]; out[3 * i + 2] = v[2];
    }
}

template <bool kDebug>
__global__ void __launch_bounds__(kRollThreads, 3)
rollout_kernel(const RolloutArgs A) {
    extern __shared__ __align__(16) float sw[];  // table image, kTabImageBytes
    __shared__ unsigned long long s_stats[NFSP_STATS_FIELDS];
    if (threadIdx.x < NFSP_STATS_FIELDS) s_stats[threadIdx.x] = 0ull;
    {
        const float4 *src = reinterpret_cast<const float4 *>(A.pack);
        float4 *dst = reinterpret_cast<float4 *>(sw);
        for (int e = threadIdx.x; e < kTabImageFloats / 4; e += blockDim.x) dst[e] = src[e];
        __syncthreads();
    }
    Counters c;
    const int64_t plane = (int64_t)A.n_steps * A.n;
    // all lanes of a warp stay in the loop together (warp collectives in decide_finish)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x & ~31); base < A.n; base += stride) {
        const int64_t i = base + (threadIdx.x & 31);
        const bool live = i < A.n;
        const uint32_t seg = (uint32_t)(base >> 5) & (A.n_seg - 1u);
        const uint64_t game = A.game0 + (uint64_t)i;
        NfspW g{live ? A.state[i] : 0ull};
        for (int t = 0; t < A.n_steps; ++t) {
            Decision d;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            if (live) {
                decide_begin(g, A, game, A.step0 + (uint64_t)t, d, c);
                if (d.random) {
                    v0 = d.v0; v1 = d.v1; v2 = d.v2;
                } else {
                    float v[3];
                    mlp_forward_tables(sw, g.hist(), g.card(d.p), g.round() ? 1u + g.pub() : 0u, g.dealer(),
                                       d.p * 2 + (int)d.pol, v);
                    v0 = v[0]; v1 = v[1]; v2 = v[2];
                }
            }
            decide_finish<kDebug>(g, A, d, v0, v1, v2, live, (int64_t)t * A.n + i, plane, seg, c, s_stats);
        }
        if (live) A.state[i] = g.w;
    }
    if (A.stats) c.commit(s_stats, A.stats);
}

}  // namespace nfsp

using namespace nfsp;

extern "C" int nfsp_act_set_weights(nfsp_env_t h, const float *d_weights, void *stream) {
    NFSP_CHECK_ARG(h != nullptr && d_weights != nullptr, "null argument");
    NFSP_CHECK_ARG(h->rules == NFSP_RULES_NFSP, "acting nets need NFSP rules");
    DeviceGuard guard(h->device);
    if (!guard.ok) return set_error(NFSP_E_CUDA, "cannot select device %d", h->device);
    if (!h->d_wpack) {
        NFSP_CUDA(cudaMalloc(&h->d_wpack, sizeof(float) * (kPackFloats + kTabImageFloats)));
        NFSP_CUDA(cudaFuncSetAttribute(rollout_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTabImageBytes));
        NFSP_CUDA(cudaFuncSetAttribute(rollout_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTabImageBytes));
    }
    pack_weights_kernel<<<(kPackFloats + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_weights, h->d_wpack);
    NFSP_LAUNCH_CHECK();
    pack_tables_kernel<<<(kTabImageFloats + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_weights,
                                                                                         h->d_wpack + kPackFloats);
    NFSP_LAUNCH_CHECK();
    const int rc = nfsp_pack_tc_image(h, d_weights, (cudaStream_t)stream);
    if (rc != NFSP_OK) return rc;
    h->has_weights = true;
    return NFSP_OK;
}

extern "C" int nfsp_act_forward(nfsp_env_t h, const uint32_t *d_obs, const int8_t *d_net, int64_t n, float *d_out,
                                void *stream) {
    NFSP_CHECK_ARG(h != nullptr && d_obs && d_net && d_out && n >= 0, "bad arguments");
    if (!h->has_weights) return set_error(NFSP_E_STATE, "nfsp_act_set_weights has not been called");
    if (n == 0) return NFSP_OK;
    DeviceGuard guard(h->device);
    if (!guard.ok) return set_error(NFSP_E_CUDA, "cannot select device %d", h->device);
    const int grid = grid_for(n, kActThreads, h->sm_count, 4);
    act_forward_kernel<<<grid, kActThreads, 0, (cudaStream_t)stream>>>(h->d_wpack, d_obs, d_net, n, d_out);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_rollout(nfsp_env_t h, int n_steps, double eta, double epsilon, const nfsp_rollout_io *io,
                            void *stream) {
    NFSP_CHECK_ARG(h != nullptr && io != nullptr, "null argument");
    NFSP_CHECK_ARG(h->rules == NFSP_RULES_NFSP, "rollout needs NFSP rules");
    NFSP_CHECK_ARG(n_steps >= 1 && n_steps <= 4096, "n_steps must be in [1,4096]");
    NFSP_CHECK_ARG(io->d_rl[0] && io->d_rl[1] && io->d_sl[0] && io->d_sl[1] && io->d_counts, "missing staging arrays");
    NFSP_CHECK_ARG(io->cap_rl > 0 && io->cap_sl > 0 && io->cap_rl < ((int64_t)1 << 32) && io->cap_sl < ((int64_t)1 << 32),
                   "staging capacity out of range");
    NFSP_CHECK_ARG(io->n_segments >= 1 && io->n_segments <= 65536 && (io->n_segments & (io->n_segments - 1)) == 0,
                   "n_segments must be a power of two in [1,65536]");
    if (!h->has_weights) return set_error(NFSP_E_STATE, "nfsp_act_set_weights has not been called");
    DeviceGuard guard(h->device);
    if (!guard.ok) return set_error(NFSP_E_CUDA, "cannot select device %d", h->device);
    RolloutArgs A;
    A.state = h->d_state; A.n = h->n; A.seed = h->seed; A.game0 = h->game0; A.step0 = h->step; A.n_steps = n_steps;
    A.eta_u32 = frac_u32(eta); A.eps_u32 = frac_u32(epsilon); A.pack = h->d_wpack + kPackFloats;
    for (int q = 0; q < 2; ++q) { A.rl[q] = (uint4 *)io->d_rl[q]; A.sl[q] = (uint4 *)io->d_sl[q]; }
    A.cap_rl = io->cap_rl; A.cap_sl = io->cap_sl; A.n_seg = (uint32_t)io->n_segments; A.counts = io->d_counts;
    A.stats = (unsigned long long *)io->d_stats; A.trace = io->d_trace; A.vec = io->d_vec; A.forced = io->d_forced_vec;
    const bool debug = io->d_trace || io->d_vec || io->d_forced_vec;
    NFSP_CHECK_ARG(io->variant >= 0 && io->variant <= 2, "variant must be 0 (default), 1 (CUDA cores) or 2 (tcgen05)");
    const int variant = io->variant == 0 ? NFSP_ROLLOUT_DEFAULT_VARIANT : io->variant;
    if (variant == 2) {
        const int rc = nfsp_rollout_tc_launch(h, A, debug, (cudaStream_t)stream);
        if (rc != NFSP_OK) return rc;
        h->step += (uint64_t)n_steps;
        return NFSP_OK;
    }
    const int grid = grid_for(h->n, kRollThreads, h->sm_count, 3);
    if (debug) rollout_kernel<true><<<grid, kRollThreads, kTabImageBytes, (cudaStream_t)stream>>>(A);
    else rollout_kernel<false><<<grid, kRollThreads, kTabImageBytes, (cudaStream_t)stream>>>(A);
    NFSP_LAUNCH_CHECK();
    h->step += (uint64_t)n_steps;
    return NFSP_OK;
}
