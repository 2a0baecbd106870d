// Group-factorised first layer of the fused rollout (CUDA-core variants): the table image, its builder and the
// per-decision forward that streams two rotated table rows into the second layer.  Shared by rollout_kernel
// (act_kernels.cu) and rollout_sorted_kernel (rollout_sorted.cu).
#pragma once
#include "mlp_math.cuh"
#include "../../include/nfsp_b200.h"

namespace nfsp {

constexpr int kB2Floats = 4 * 4;  // b2 as [4 nets][4]

// ---- group-factorised first layer (rollout path) --------------------------------------------------
// Under main.train's turn order an observation is (cards of the actor, round-0 betting sequence,
// round-1 betting sequence); each group takes few values, so W1^T x + b1 is the sum of TWO precombined
// rows X + Y.  The rows are sums of W1 rows only (input independent), rebuilt whenever the weights change.
//   per net 111 rows:
//     X  0-2    round 0: private card c (+ b1)
//        3-74   round 1: 3 + ((c*3 + pub)*2 + dealer)*4 + f, f = finished round-0 sequence CC, RC, CRC, RRC
//               = card rows + public card + round-0 history bits (+ b1)
//     Y  75-110 75 + dealer*18 + sigma, sigma = 9*round + sequence id so far in that round
//               (0 -, 1 C, 2 R, 3 CC, 4 CR, 5 RC, 6 RR, 7 CRC, 8 RRC)
// Shared-memory layout: a row is 24 float4 "quads": the 16 quads of the 64 hidden units followed by a copy of
// the first 8, so a thread can read its row starting at quad rot = game & 7 with immediate offsets.  The 8
// threads of a quarter-warp then hit 8 different 16-byte bank groups whatever rows they gather: the row
// gathers are bank-conflict free (the straight layout ran at 2.0-2.5x the ideal wavefront count, ncu r01a).
// W2 uses the same rotated layout, [net][output][24 quads], so a quad of h meets its own weights.
constexpr int kNetRows = 111, kRowQuads = 24;
constexpr int kTabRows = 4 * kNetRows;
constexpr int kTabFloats = kTabRows * kRowQuads * 4;
constexpr int kTabW2Floats = 4 * 3 * kRowQuads * 4;
constexpr int kTabImageFloats = kTabFloats + kTabW2Floats + kB2Floats;
constexpr int kTabImageBytes = kTabImageFloats * 4;

// history bits of betting sequence `id` in round rr when `d` deals (the dealer opens, players alternate)
__device__ __forceinline__ uint32_t seq_bits(int rr, int d, int id) {
    const int s0 = id == 0 ? 0 : (id == 1 || id == 3 || id == 4 || id == 7 ? 1 : 2);
    const int s1 = id < 3 ? 0 : (id == 3 || id == 5 ? 1 : 2);
    const int s2 = id >= 7 ? 1 : 0;
    const int sl[3] = {s0, s1, s2};
    uint32_t bits = 0;
    for (int k = 0; k < 3; ++k)
        if (sl[k]) bits |= 1u << (((k & 1) ^ d) * 12 + rr * 6 + k * 2 + (sl[k] - 1));
    return bits;
}

// element e of the rollout's table image
__device__ __forceinline__ float pack_tables_value(const float *__restrict__ w, int e) {
    {
        float v = 0.f;
        if (e < kTabFloats) {
            const int x = e & 3, pos = (e >> 2) % kRowQuads, row = (e >> 2) / kRowQuads;
            const int net = row / kNetRows, r = row % kNetRows, j = (pos & 15) * 4 + x;
            const float *W1 = w + net * NFSP_NET_PARAMS;  // W1[i*64 + j]
            uint32_t bits = 0;                            // observation bits this row stands for
            bool bias = false;
            if (r < 3) {
                bits = 1u << (24 + r);
                bias = true;
            } else if (r < 75) {
                const int idx = r - 3, f = idx & 3, d = (idx >> 2) & 1, cp = idx >> 3, c = cp / 3, pub = cp % 3;
                const int fin[4] = {3, 5, 7, 8};
                bits = (1u << (24 + c)) | (1u << (27 + c)) | (1u << (27 + pub)) | seq_bits(0, d, fin[f]);
                bias = true;
            } else {
                bits = seq_bits(((r - 75) % 18) / 9, (r - 75) / 18, (r - 75) % 9);
            }
            for (int i = 0; i < 30; ++i)
                if ((bits >> i) & 1u) v += W1[i * 64 + j];
            if (bias) v += W1[1920 + j];
        } else if (e < kTabFloats + kTabW2Floats) {
            const int f = e - kTabFloats, x = f & 3, pos = (f >> 2) % kRowQuads, c = ((f >> 2) / kRowQuads) % 3;
            const int net = (f >> 2) / (3 * kRowQuads);
            v = w[net * NFSP_NET_PARAMS + 1984 + ((pos & 15) * 4 + x) * 3 + c];
        } else {
            const int f = e - kTabFloats - kTabW2Floats, c = f & 3, net = f >> 2;
            if (c < 3) v = w[net * NFSP_NET_PARAMS + 2176 + c];
        }
        return v;
    }
}

// one decision of net `net`: layer 1 as two rotated row reads streamed into layer 2
__device__ __forceinline__ void mlp_forward_tables(const float *__restrict__ st, uint32_t xrow, uint32_t yrow,
                                                   uint32_t net, uint32_t rot, float &o0, float &o1, float &o2) {
    const float4 *T = reinterpret_cast<const float4 *>(st);
    const float4 *xr = T + (net * kNetRows + xrow) * kRowQuads + rot;
    const float4 *yr = T + (net * kNetRows + yrow) * kRowQuads + rot;
    const float4 *wr = T + kTabRows * kRowQuads + net * (3 * kRowQuads) + rot;
    Layer2Acc acc;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        acc.quad_sum(xr[q], yr[q], wr[q], wr[kRowQuads + q], wr[2 * kRowQuads + q]);
    }
    acc.head(reinterpret_cast<const float4 *>(st + kTabFloats + kTabW2Floats)[net], net & 1u, o0, o1, o2);
}

}  // namespace nfsp
