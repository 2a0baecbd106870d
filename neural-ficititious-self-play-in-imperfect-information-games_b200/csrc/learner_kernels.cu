// Learner half of agent.Agent (agent.py:209-273), SURVEY 8 f-1: one launch computes the minibatch
// gradients of all four acting nets (player x {average policy, best response}) straight from the
// packed records of the memories, into ONE flat fp32 buffer that the host all-reduces over GPUs
// (NCCL) together with the exploitability statistics; a second kernel applies the SGD step.
//
//   best response (agent.py:90-107, 209-253): Q = relu(W2^T relu(W1^T s + b1) + b2), Huber loss on the
//     taken action against r + gamma * (1 - t) * max_a' Q_target(s2), mean over the 3 outputs and the rows;
//   average policy (agent.py:109-116, 255-264): softmax head, categorical cross-entropy against the stored
//     score vector a (Keras semantics: the target is used as is, not renormalised).
// One CTA per net, kRowGroups groups of 64 threads, one thread per hidden unit and group: the thread keeps its column
// of W1, its row of W2 and the matching gradient accumulators in registers; the 64-wide reductions of layer 2 go
// through shuffles and a 64-thread named barrier.  The groups take the minibatch rows round-robin (the kernel is
// latency-bound: three dependent reductions per row) and their partial gradients are summed in a fixed order.
#include "common.cuh"

namespace nfsp {

constexpr int kRowGroups = 4;
constexpr int kLearnThreads = 64 * kRowGroups;
constexpr int kGradFloats = 4 * NFSP_NET_PARAMS;

struct LearnerArgs {
    const float *w;          // [4][2179] acting nets, index player*2 + policy
    const float *w_target;   // [2][2179] target best-response nets (agent.py:71-72)
    const uint4 *rl[2];      // rings (M_RL) of the two players
    const int64_t *rl_idx[2];
    const uint4 *sl[2];      // reservoirs (M_SL)
    const int64_t *sl_idx[2];
    int row0, rows;          // minibatch = sampled rows [row0, row0 + rows)
    float gamma;
    int net_mask;            // bit k set = net k takes part (memory large enough, agent.py:215,259)
    int terminal_bootstraps; // reference quirk agent.py:227: `t is True` never holds, terminals bootstrap too
    float *grad;             // [4][2179] mean gradients (+= nothing: overwritten)
    float *stats;            // [NFSP_LEARNER_STATS]
};

struct NetRegs {
    float w1[30], b1, w2[3];
    __device__ __forceinline__ void load(const float *w, int j) {
#pragma unroll
        for (int i = 0; i < 30; ++i) w1[i] = w[i * 64 + j];
        b1 = w[1920 + j];
#pragma unroll
        for (int c = 0; c < 3; ++c) w2[c] = w[1984 + j * 3 + c];
    }
    __device__ __forceinline__ float hidden(uint32_t obs) const {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 30; ++i) acc += ((obs >> i) & 1u) ? w1[i] : 0.f;
        return fmaxf(acc + b1, 0.f);
    }
};

// sum over the 64 threads of a group (2 warps); every thread of the group gets the 3 totals
__device__ __forceinline__ void group_sync(int group) { asm volatile("bar.sync %0, 64;" ::"r"(1 + group) : "memory"); }
__device__ __forceinline__ void cta_sum3(float &a, float &b, float &c, float (*red)[2][4], int slot, int group) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
        c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    }
    const int warp = (threadIdx.x >> 5) & 1;
    if ((threadIdx.x & 31) == 0) { red[slot][warp][0] = a; red[slot][warp][1] = b; red[slot][warp][2] = c; }
    group_sync(group);
    a = red[slot][0][0] + red[slot][1][0];
    b = red[slot][0][1] + red[slot][1][1];
    c = red[slot][0][2] + red[slot][1][2];
}

__global__ void __launch_bounds__(kLearnThreads)
learner_grad_kernel(const LearnerArgs A) {
    __shared__ float red_all[kRowGroups][4][2][4];
    __shared__ float part[kRowGroups][40][64];  // per group: gw1[30], gb1, gw2[3] per hidden unit; row 34: gb2, loss, expl
    const int net = blockIdx.x, player = net >> 1, is_br = net & 1, group = threadIdx.x >> 6, j = threadIdx.x & 63;
    float (*red)[2][4] = red_all[group];
    float *g = A.grad + net * NFSP_NET_PARAMS;
    if (!((A.net_mask >> net) & 1)) {
        for (int e = threadIdx.x; e < NFSP_NET_PARAMS; e += kLearnThreads) g[e] = 0.f;
        return;
    }
    NetRegs W, T;
    W.load(A.w + net * NFSP_NET_PARAMS, j);
    const float b2_0 = A.w[net * NFSP_NET_PARAMS + 2176], b2_1 = A.w[net * NFSP_NET_PARAMS + 2177],
                b2_2 = A.w[net * NFSP_NET_PARAMS + 2178];
    float t2_0 = 0.f, t2_1 = 0.f, t2_2 = 0.f;
    if (is_br) {
        const float *wt = A.w_target + player * NFSP_NET_PARAMS;
        T.load(wt, j);
        t2_0 = wt[2176]; t2_1 = wt[2177]; t2_2 = wt[2178];
    }
    float gw1[30], gb1 = 0.f, gw2[3] = {0.f, 0.f, 0.f}, gb2[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 30; ++i) gw1[i] = 0.f;
    float loss = 0.f, expl = 0.f;
    const float inv_rows = 1.0f / (float)A.rows;
    for (int r = group; r < A.rows; r += kRowGroups) {
        const int row = A.row0 + r;
        uint32_t s;
        float dz0, dz1, dz2;
        if (is_br) {
            const uint4 rec = A.rl[player][A.rl_idx[player][row]];
            s = rec.x;
            const uint32_t a = rec.w & 0xFFu, term = (rec.w >> 8) & 0xFFu;
            const float rew = __uint_as_float(rec.z);
            // online Q(s), target Q(s2), target Q(s) (exploitability proxy, agent.py:234-238)
            const float h = W.hidden(s);
            float z0 = h * W.w2[0], z1 = h * W.w2[1], z2 = h * W.w2[2];
            cta_sum3(z0, z1, z2, red, 0, group);
            const float ht = T.hidden(rec.y);
            float y0 = ht * T.w2[0], y1 = ht * T.w2[1], y2 = ht * T.w2[2];
            cta_sum3(y0, y1, y2, red, 1, group);
            const float hs = T.hidden(s);
            float e0 = hs * T.w2[0], e1 = hs * T.w2[1], e2 = hs * T.w2[2];
            cta_sum3(e0, e1, e2, red, 2, group);
            expl += fmaxf(fmaxf(fmaxf(e0 + t2_0, 0.f), fmaxf(e1 + t2_1, 0.f)), fmaxf(e2 + t2_2, 0.f));
            const float qn = fmaxf(fmaxf(fmaxf(y0 + t2_0, 0.f), fmaxf(y1 + t2_1, 0.f)), fmaxf(y2 + t2_2, 0.f));
            const float target = rew + ((term && !A.terminal_bootstraps) ? 0.f : A.gamma * qn);
            z0 += b2_0; z1 += b2_1; z2 += b2_2;
            const float za = a == 0 ? z0 : (a == 1 ? z1 : z2);
            const float err = target - fmaxf(za, 0.f);                     // y - Q(s,a), other outputs have zero error
            const float ae = fabsf(err);
            loss += (ae > 1.f ? ae - 0.5f : 0.5f * err * err) * (1.0f / 3.0f);
            const float dq = -(ae > 1.f ? copysignf(1.f, err) : err) * (1.0f / 3.0f) * (za > 0.f ? 1.f : 0.f);
            dz0 = a == 0 ? dq : 0.f; dz1 = a == 1 ? dq : 0.f; dz2 = a == 2 ? dq : 0.f;
            group_sync(group);  // red[] slots are reused by the next row
            const float dh = (h > 0.f) ? (W.w2[0] * dz0 + W.w2[1] * dz1 + W.w2[2] * dz2) : 0.f;
            gw2[0] += h * dz0; gw2[1] += h * dz1; gw2[2] += h * dz2;
            gb1 += dh;
#pragma unroll
            for (int i = 0; i < 30; ++i) gw1[i] += ((s >> i) & 1u) ? dh : 0.f;
        } else {
            const uint4 rec = A.sl[player][A.sl_idx[player][row]];
            s = rec.x;
            const float ya = __uint_as_float(rec.y), yb = __uint_as_float(rec.z), yc = __uint_as_float(rec.w);
            const float h = W.hidden(s);
            float z0 = h * W.w2[0], z1 = h * W.w2[1], z2 = h * W.w2[2];
            cta_sum3(z0, z1, z2, red, 0, group);
            z0 += b2_0; z1 += b2_1; z2 += b2_2;
            const float m = fmaxf(z0, fmaxf(z1, z2));
            const float x0 = expf(z0 - m), x1 = expf(z1 - m), x2 = expf(z2 - m);
            const float inv = 1.0f / (x0 + x1 + x2);
            const float p0 = x0 * inv, p1 = x1 * inv, p2 = x2 * inv;
            const float ysum = ya + yb + yc;
            loss += -(ya * logf(fmaxf(p0, 1e-7f)) + yb * logf(fmaxf(p1, 1e-7f)) + yc * logf(fmaxf(p2, 1e-7f)));
            dz0 = p0 * ysum - ya; dz1 = p1 * ysum - yb; dz2 = p2 * ysum - yc;
            group_sync(group);
            const float dh = (h > 0.f) ? (W.w2[0] * dz0 + W.w2[1] * dz1 + W.w2[2] * dz2) : 0.f;
            gw2[0] += h * dz0; gw2[1] += h * dz1; gw2[2] += h * dz2;
            gb1 += dh;
#pragma unroll
            for (int i = 0; i < 30; ++i) gw1[i] += ((s >> i) & 1u) ? dh : 0.f;
        }
        gb2[0] += dz0; gb2[1] += dz1; gb2[2] += dz2;
    }
    // the groups' partial sums meet in shared memory and are added in group order (deterministic)
#pragma unroll
    for (int i = 0; i < 30; ++i) part[group][i][j] = gw1[i];
    part[group][30][j] = gb1;
#pragma unroll
    for (int c = 0; c < 3; ++c) part[group][31 + c][j] = gw2[c];
    if (j == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) part[group][34][c] = gb2[c];
        part[group][34][3] = loss;
        part[group][34][4] = expl;
    }
    __syncthreads();
    if (group != 0) return;
    float tot[34];
#pragma unroll
    for (int i = 0; i < 34; ++i) {
        float v = part[0][i][j];
#pragma unroll
        for (int k = 1; k < kRowGroups; ++k) v += part[k][i][j];
        tot[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 30; ++i) g[i * 64 + j] = tot[i] * inv_rows;
    g[1920 + j] = tot[30] * inv_rows;
#pragma unroll
    for (int c = 0; c < 3; ++c) g[1984 + j * 3 + c] = tot[31 + c] * inv_rows;
    if (j < 5) {
        float v = part[0][34][j];
#pragma unroll
        for (int k = 1; k < kRowGroups; ++k) v += part[k][34][j];
        if (j < 3) g[2176 + j] = v * inv_rows;
        // stats: [0,1] exploitability-proxy sums of the BR nets, [2,3] their row counts, [4..7] loss sums per net
        if (j == 3) A.stats[4 + net] = v;
        if (j == 4 && is_br) { A.stats[player] = v; A.stats[2 + player] = (float)A.rows; }
    }
}

// w[k] -= lr[k] * scale * grad[k]   (keras.optimizers.SGD without momentum, agent.py:45-46)
__global__ void sgd_apply_kernel(float *__restrict__ w, const float *__restrict__ grad, float lr0, float lr1, float lr2,
                                 float lr3, float scale) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= kGradFloats) return;
    const int net = e / NFSP_NET_PARAMS;
    const float lr = net == 0 ? lr0 : (net == 1 ? lr1 : (net == 2 ? lr2 : lr3));
    w[e] -= lr * scale * grad[e];
}

}  // namespace nfsp

using namespace nfsp;

extern "C" int nfsp_learner_grads(const nfsp_learner_io *io, void *stream) {
    NFSP_CHECK_ARG(io != nullptr && io->d_weights && io->d_target_weights && io->d_grad && io->d_stats, "null argument");
    NFSP_CHECK_ARG(io->rows >= 1 && io->row0 >= 0, "bad minibatch slice");
    LearnerArgs A;
    A.w = io->d_weights; A.w_target = io->d_target_weights;
    for (int p = 0; p < 2; ++p) {
        A.rl[p] = (const uint4 *)io->d_rl[p]; A.rl_idx[p] = io->d_rl_idx[p];
        A.sl[p] = (const uint4 *)io->d_sl[p]; A.sl_idx[p] = io->d_sl_idx[p];
        if ((io->net_mask >> (2 * p + 1)) & 1) NFSP_CHECK_ARG(A.rl[p] && A.rl_idx[p], "missing RL batch of player %d", p);
        if ((io->net_mask >> (2 * p)) & 1) NFSP_CHECK_ARG(A.sl[p] && A.sl_idx[p], "missing SL batch of player %d", p);
    }
    A.row0 = io->row0; A.rows = io->rows; A.gamma = io->gamma; A.net_mask = io->net_mask;
    A.terminal_bootstraps = io->terminal_bootstraps; A.grad = io->d_grad; A.stats = io->d_stats;
    learner_grad_kernel<<<4, kLearnThreads, 0, (cudaStream_t)stream>>>(A);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_sgd_apply(float *d_weights, const float *d_grad, const float lr[4], float scale, void *stream) {
    NFSP_CHECK_ARG(d_weights && d_grad && lr, "null argument");
    sgd_apply_kernel<<<(kGradFloats + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_weights, d_grad, lr[0], lr[1], lr[2],
                                                                                  lr[3], scale);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}
