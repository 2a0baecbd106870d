// K3/K4/K5: the agent memories.
//   K3 circular replay insert  (utils/replay_buffer.py:30-41)   16-byte records, streaming copy
//   K4 reservoir insert        (utils/ReservoirBuffer.py:18-28) Philox Algorithm R, two passes
//   K5 minibatch sample        (replay_buffer.py:46-59, ReservoirBuffer.py:33-43) Floyd + gather/expand
// All counts live in device memory: no host synchronisation on the hot path.
#include "common.cuh"
#include "philox.cuh"

namespace nfsp {

constexpr int kBufThreads = 256;

__device__ __forceinline__ uint64_t mulhi64(uint64_t a, uint64_t b) { return __umul64hi(a, b); }

// ---- staged batches --------------------------------------------------------------------------------
// A batch is n_seg segments of seg_cap slots with a device count each (n_seg = 1: a plain dense array).
// Record i of segment s has batch index prefix(s) + i, where prefix(s) = sum of the counts of the segments
// before s; its ticket is total + that index.  Each CTA row (blockIdx.y = segment) recomputes its prefix
// with one block reduction over the <= 65536 counts.
struct Batch {
    const uint4 *recs;
    const uint32_t *counts;
    uint32_t n_seg;
    uint64_t seg_cap;
};

__device__ __forceinline__ void batch_prefix(const Batch &B, uint32_t seg, uint64_t &before, uint64_t &all) {
    __shared__ unsigned long long s_before, s_all;
    if (threadIdx.x == 0) { s_before = 0ull; s_all = 0ull; }
    __syncthreads();
    unsigned long long lb = 0ull, la = 0ull;
    for (uint32_t k = threadIdx.x; k < B.n_seg; k += blockDim.x) {
        unsigned long long cnt = B.counts[k];
        if (cnt > B.seg_cap) cnt = B.seg_cap;
        la += cnt;
        if (k < seg) lb += cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lb += __shfl_xor_sync(0xFFFFFFFFu, lb, o);
        la += __shfl_xor_sync(0xFFFFFFFFu, la, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_before, lb); atomicAdd(&s_all, la); }
    __syncthreads();
    before = s_before;
    all = s_all;
}

// ---- memories of one launch ----------------------------------------------------------------------
// The insert kernels take up to NFSP_MAX_INSERT_REQS memories at once (blockIdx.z = memory): the two players' rings
// travel in one launch, their reservoirs in another -- at 64k games the flush is launch latency, not bandwidth.
struct Mem {
    uint4 *data;
    unsigned long long *stamp;  // reservoirs only
    uint64_t cap;
    uint64_t *total;
    Batch B;
    uint64_t seed;
    int mode;
};
struct MemSet {
    Mem m[NFSP_MAX_INSERT_REQS];
};

// ---- K3 ------------------------------------------------------------------------------------------
constexpr int kRingUnroll = 4;
// slot = ticket % cap; of a batch larger than the ring only the last `cap` records survive (the others
// would be evicted by popleft(), replay_buffer.py:40).
__global__ void __launch_bounds__(kBufThreads) ring_insert_kernel(const MemSet S) {
    const Mem &M = S.m[blockIdx.z];
    const Batch &B = M.B;
    const uint32_t seg = blockIdx.y;
    if (seg >= B.n_seg) return;
    uint64_t before, m;
    batch_prefix(B, seg, before, m);
    const uint64_t total = *M.total, cap = M.cap;
    const uint64_t first = m > cap ? m - cap : 0;
    uint64_t cnt = B.counts[seg];
    if (cnt > B.seg_cap) cnt = B.seg_cap;
    const uint4 *src = B.recs + (uint64_t)seg * B.seg_cap;
    uint4 *ring = M.data;
    // slot of batch index idx = (total + idx) % cap without a 64-bit division per record: head = total % cap once, then
    // one conditional subtraction (a second reduction only for batches larger than the ring).  kRingUnroll records per
    // thread and iteration, the loads issued before the first store, so several 16-byte reads are in flight per thread.
    const uint64_t head = total % cap;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i0 < cnt; i0 += stride * kRingUnroll) {
        uint4 v[kRingUnroll];
#pragma unroll
        for (int u = 0; u < kRingUnroll; ++u) {
            const uint64_t i = i0 + (uint64_t)u * stride;
            if (i < cnt) v[u] = src[i];
        }
#pragma unroll
        for (int u = 0; u < kRingUnroll; ++u) {
            const uint64_t i = i0 + (uint64_t)u * stride, idx = before + i;
            if (i < cnt && idx >= first) {
                uint64_t slot = head + idx;
                if (slot >= cap) slot -= cap;
                if (slot >= cap) slot %= cap;
                ring[slot] = v[u];
            }
        }
    }
}

// runs after the insert kernel(s) of a batch, one CTA per memory: total += n, staged counts cleared for the next rollout
__global__ void __launch_bounds__(kBufThreads) commit_kernel(const MemSet S) {
    const Mem &M = S.m[blockIdx.x];
    uint64_t before, m;
    batch_prefix(M.B, 0, before, m);
    __syncthreads();
    uint32_t *counts = const_cast<uint32_t *>(M.B.counts);
    for (uint32_t k = threadIdx.x; k < M.B.n_seg; k += blockDim.x) counts[k] = 0;
    if (threadIdx.x == 0) *M.total += m;
}

// ---- K4 ------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t reservoir_slot(uint64_t seed, uint64_t ticket, uint64_t cap, int mode) {
    if (ticket < cap) return (int64_t)ticket;  // fill phase, ReservoirBuffer.py:22-24
    const uint64_t u = buffer_u64(seed, ticket, 0, STREAM_RESERVOIR);
    if (mode == 1) {  // the reference's law: j = randrange(1, B+1); replace iff j < B (ReservoirBuffer.py:26-28)
        const uint64_t j = 1u + mulhi64(u, cap);
        return j < cap ? (int64_t)j : -1;
    }
    const uint64_t j = mulhi64(u, ticket + 1u);  // Algorithm R: j ~ U[0, ticket]
    return j < cap ? (int64_t)j : -1;
}

// Both passes are random 8/16-byte accesses whose latency (not bandwidth) is the bound with one access in flight
// per thread, so each thread handles kResUnroll records per iteration: the slot draws are independent and the
// atomics / loads of an iteration are all issued before the first result is needed.
constexpr int kResUnroll = 4;

// pass 1: every accepted record stamps its slot with ticket+1; atomicMax keeps the latest
__global__ void __launch_bounds__(kBufThreads) reservoir_stamp_kernel(const MemSet S) {
    const Mem &M = S.m[blockIdx.z];
    const Batch &B = M.B;
    const uint32_t seg = blockIdx.y;
    if (seg >= B.n_seg) return;
    uint64_t before, m;
    batch_prefix(B, seg, before, m);
    const uint64_t total = *M.total, cap = M.cap, seed = M.seed;
    const int mode = M.mode;
    unsigned long long *stamp = M.stamp;
    uint64_t cnt = B.counts[seg];
    if (cnt > B.seg_cap) cnt = B.seg_cap;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i0 < cnt; i0 += stride * kResUnroll) {
        int64_t slot[kResUnroll];
#pragma unroll
        for (int u = 0; u < kResUnroll; ++u) {
            const uint64_t i = i0 + (uint64_t)u * stride;
            slot[u] = i < cnt ? reservoir_slot(seed, total + before + i, cap, mode) : -1;
        }
#pragma unroll
        for (int u = 0; u < kResUnroll; ++u)
            if (slot[u] >= 0) atomicMax(stamp + slot[u], (unsigned long long)(total + before + i0 + (uint64_t)u * stride + 1u));
    }
}

// pass 2: the record whose ticket owns the stamp writes the payload (== sequential order of adds)
__global__ void __launch_bounds__(kBufThreads) reservoir_write_kernel(const MemSet S) {
    const Mem &M = S.m[blockIdx.z];
    const Batch &B = M.B;
    const uint32_t seg = blockIdx.y;
    if (seg >= B.n_seg) return;
    uint64_t before, m;
    batch_prefix(B, seg, before, m);
    const uint64_t total = *M.total, cap = M.cap, seed = M.seed;
    const int mode = M.mode;
    const unsigned long long *stamp = M.stamp;
    uint4 *res = M.data;
    uint64_t cnt = B.counts[seg];
    if (cnt > B.seg_cap) cnt = B.seg_cap;
    const uint4 *src = B.recs + (uint64_t)seg * B.seg_cap;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i0 < cnt; i0 += stride * kResUnroll) {
        int64_t slot[kResUnroll];
        unsigned long long owner[kResUnroll];
#pragma unroll
        for (int u = 0; u < kResUnroll; ++u) {
            const uint64_t i = i0 + (uint64_t)u * stride;
            slot[u] = i < cnt ? reservoir_slot(seed, total + before + i, cap, mode) : -1;
        }
#pragma unroll
        for (int u = 0; u < kResUnroll; ++u) owner[u] = slot[u] >= 0 ? stamp[slot[u]] : 0ull;
#pragma unroll
        for (int u = 0; u < kResUnroll; ++u) {
            const uint64_t i = i0 + (uint64_t)u * stride;
            if (slot[u] >= 0 && owner[u] == (unsigned long long)(total + before + i + 1u)) res[slot[u]] = src[i];
        }
    }
}

// ---- K5 ------------------------------------------------------------------------------------------
// Floyd's algorithm (1987) for `batch` distinct positions out of `count`: for i = count-batch .. count-1
// draw t ~ U[0,i]; take t unless already taken, else take i.  One CTA: the draws are computed in
// parallel, the "already taken" scan is a warp-parallel compare over the prefix.
constexpr int kMaxBatch = 1024;

// draw / pick: kMaxBatch words of shared memory each; slot_out[m] (shared or global) = storage slot of sample m, -1
// beyond the records held.  Returns the number of samples (min(count, batch)).  All threads of the CTA call it.
template <class SlotT>
__device__ __forceinline__ int floyd_sample(uint64_t seed, uint64_t call_idx, uint64_t total, uint64_t cap, int is_ring,
                                            int batch, uint64_t *draw, uint64_t *pick, SlotT *slot_out) {
    __shared__ int s_collision;
    const uint64_t count = total < cap ? total : cap;
    const int b = (uint64_t)batch < count ? batch : (int)count;
    const uint64_t lo = count - (uint64_t)b;
    if (threadIdx.x == 0) s_collision = 0;
    for (int m = threadIdx.x; m < b; m += blockDim.x)
        draw[m] = mulhi64(buffer_u64(seed, (uint64_t)m, call_idx, STREAM_SAMPLE), lo + (uint64_t)m + 1u);
    __syncthreads();
    // Floyd: step m takes t = draw[m] unless t was taken before, else lo + m (always new).  Every earlier draw is in the
    // set whether it was taken or replaced, so "taken before" = (some earlier draw equals t) or (t = lo + k for an
    // earlier step k that was replaced).  The first term is b*b/2 independent compares; the second refers to ONE
    // earlier step, k = t - lo < m, so the replaced flags are the least fixpoint of dup[m] |= dup[draw[m] - lo], reached
    // by a few parallel sweeps (flags only turn on, references only point backwards) -- same picks as the sequential
    // algorithm, no sequential scan even when draws collide (a 200 000-record ring and 256 draws: 15 % of the calls).
    __shared__ unsigned char s_dup[kMaxBatch];
    for (int m = threadIdx.x; m < b; m += blockDim.x) {
        const uint64_t t = draw[m];
        bool dup = false;
        for (int k = 0; k < m; ++k) dup |= (draw[k] == t);
        if (dup) s_collision = 1;
        s_dup[m] = dup;
    }
    __syncthreads();
    while (s_collision) {  // uniform: every thread reads the same flag between barriers
        __syncthreads();
        if (threadIdx.x == 0) s_collision = 0;
        __syncthreads();
        for (int m = threadIdx.x; m < b; m += blockDim.x) {
            const uint64_t t = draw[m];
            if (!s_dup[m] && t >= lo && t - lo < (uint64_t)m && s_dup[t - lo]) {
                s_dup[m] = 1;
                s_collision = 1;
            }
        }
        __syncthreads();
    }
    for (int m = threadIdx.x; m < b; m += blockDim.x) pick[m] = s_dup[m] ? lo + (uint64_t)m : draw[m];
    __syncthreads();
    // deque position (oldest first) -> storage slot
    const uint64_t head = (is_ring && total >= cap) ? total % cap : 0;
    for (int m = threadIdx.x; m < batch; m += blockDim.x) {
        int64_t slot = -1;
        if (m < b) slot = is_ring ? (int64_t)((head + pick[m]) % cap) : (int64_t)pick[m];
        slot_out[m] = (SlotT)slot;
    }
    return b;
}

__global__ void __launch_bounds__(kBufThreads)
sample_kernel(uint64_t seed, uint64_t call_idx, const uint64_t *__restrict__ total_p, uint64_t cap, int is_ring,
              int batch, int64_t *__restrict__ idx_out, uint32_t *__restrict__ n_out) {
    __shared__ uint64_t draw[kMaxBatch];
    __shared__ uint64_t pick[kMaxBatch];
    const int b = floyd_sample(seed, call_idx, *total_p, cap, is_ring, batch, draw, pick, idx_out);
    if (threadIdx.x == 0 && n_out) *n_out = (uint32_t)b;
}

// one output element of a sampled RL / SL row (the dense views of replay_buffer.py:46-59 / ReservoirBuffer.py:33-43)
__device__ __forceinline__ void expand_rl(const uint4 rec, int row, int col, float *__restrict__ s, float *__restrict__ a,
                                          float *__restrict__ r, float *__restrict__ s2, float *__restrict__ t) {
    if (col < 30) s[row * 30 + col] = (float)((rec.x >> col) & 1u);
    else if (col < 60) s2[row * 30 + (col - 30)] = (float)((rec.y >> (col - 30)) & 1u);
    else if (col < 63) a[row * 3 + (col - 60)] = ((rec.w & 0xFFu) == (uint32_t)(col - 60)) ? 1.f : 0.f;
    else if (col == 63) r[row] = __uint_as_float(rec.z);
    else t[row] = (float)((rec.w >> 8) & 0xFFu);
}
__device__ __forceinline__ void expand_sl(const uint4 rec, int row, int col, float *__restrict__ s, float *__restrict__ a) {
    if (col < 30) s[row * 30 + col] = (float)((rec.x >> col) & 1u);
    else a[row * 3 + (col - 30)] = __uint_as_float(col == 30 ? rec.y : (col == 31 ? rec.z : rec.w));
}

// sample_batch of several memories in ONE launch: blockIdx.x = memory.  A CTA draws the positions of its memory, fetches
// its share of the sampled records once (one random 16-byte read per row, all in flight together) and expands them
// into out[m] -- an RL block is s[b][30] a[b][3] r[b] s2[b][30] t[b], an SL block s[b][30] a[b][3] (b = batch).  The
// rows of a memory are split over gridDim.y CTAs; each repeats the (cheap) draw.
struct SampleReqs {
    const uint4 *mem[NFSP_MAX_SAMPLE_REQS];
    const uint64_t *total[NFSP_MAX_SAMPLE_REQS];
    uint64_t cap[NFSP_MAX_SAMPLE_REQS], seed[NFSP_MAX_SAMPLE_REQS], call_idx[NFSP_MAX_SAMPLE_REQS];
    float *out[NFSP_MAX_SAMPLE_REQS];
    int is_ring[NFSP_MAX_SAMPLE_REQS];
};
__global__ void __launch_bounds__(kBufThreads)
sample_gather_kernel(const SampleReqs R, int batch, int64_t *__restrict__ idx_out, uint32_t *__restrict__ n_out) {
    __shared__ uint64_t draw[kMaxBatch];
    __shared__ uint64_t pick[kMaxBatch];
    __shared__ int64_t slot[kMaxBatch];
    __shared__ uint4 s_rec[kMaxBatch];
    const int m = blockIdx.x;
    const int b = floyd_sample(R.seed[m], R.call_idx[m], *R.total[m], R.cap[m], R.is_ring[m], batch, draw, pick, slot);
    __syncthreads();
    const int per = (batch + (int)gridDim.y - 1) / (int)gridDim.y;
    const int row0 = (int)blockIdx.y * per, row1 = min(batch, row0 + per);
    if (threadIdx.x == 0 && blockIdx.y == 0 && n_out) n_out[m] = (uint32_t)b;
    const uint4 *mem = R.mem[m];
    float *o = R.out[m];
    for (int k = row0 + threadIdx.x; k < row1; k += blockDim.x) {
        if (idx_out) idx_out[(int64_t)m * batch + k] = slot[k];
        if (o) s_rec[k] = slot[k] >= 0 ? mem[slot[k]] : make_uint4(0, 0, 0, 0);
    }
    if (!o) return;  // positions only (the learner reads the packed records itself)
    __syncthreads();
    const int rows = max(row1 - row0, 0);
    if (R.is_ring[m]) {
        float *s = o, *a = s + 30 * batch, *r = a + 3 * batch, *s2 = r + batch, *t = s2 + 30 * batch;
        for (int e = threadIdx.x; e < rows * 65; e += blockDim.x) {
            const int row = row0 + e / 65, col = e % 65;
            expand_rl(s_rec[row], row, col, s, a, r, s2, t);
        }
    } else {
        float *s = o, *a = s + 30 * batch;
        for (int e = threadIdx.x; e < rows * 33; e += blockDim.x) {
            const int row = row0 + e / 33, col = e % 33;
            expand_sl(s_rec[row], row, col, s, a);
        }
    }
}

__global__ void __launch_bounds__(kBufThreads)
gather_rl_kernel(const uint4 *__restrict__ ring, const int64_t *__restrict__ idx, int batch, float *__restrict__ s,
                 float *__restrict__ a, float *__restrict__ r, float *__restrict__ s2, float *__restrict__ t) {
    // one thread per output element of the widest rows (30 + 30 + 3 + 1 + 1 = 65 columns)
    const int total = batch * 65;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int row = e / 65, col = e - row * 65;
        const int64_t slot = idx[row];
        expand_rl(slot >= 0 ? ring[slot] : make_uint4(0, 0, 0, 0), row, col, s, a, r, s2, t);
    }
}

__global__ void __launch_bounds__(kBufThreads)
gather_sl_kernel(const uint4 *__restrict__ res, const int64_t *__restrict__ idx, int batch, float *__restrict__ s,
                 float *__restrict__ a) {
    const int total = batch * 33;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int row = e / 33, col = e - row * 33;
        const int64_t slot = idx[row];
        expand_sl(slot >= 0 ? res[slot] : make_uint4(0, 0, 0, 0), row, col, s, a);
    }
}

static dim3 insert_grid(int64_t seg_cap, int n_seg, int n_mems) {
    int64_t g = (seg_cap + kBufThreads - 1) / kBufThreads;
    const int64_t rows = (int64_t)n_seg * n_mems;
    const int64_t lim = rows >= 148 * 8 ? 1 : (148 * 8 + rows - 1) / rows;  // about 8 CTAs per SM in total
    if (g > lim) g = lim;
    if (g < 1) g = 1;
    return dim3((unsigned)g, (unsigned)n_seg, (unsigned)n_mems);
}

}  // namespace nfsp

using namespace nfsp;

// validates the requests and fills the kernels' argument; *seg_cap_max / *n_seg_max size the grid
static int make_memset(const nfsp_insert_req *reqs, int n, bool reservoir, MemSet &S, int64_t *seg_cap_max, int *n_seg_max) {
    NFSP_CHECK_ARG(reqs && n >= 1 && n <= NFSP_MAX_INSERT_REQS, "1..%d memories per call", NFSP_MAX_INSERT_REQS);
    *seg_cap_max = 0;
    *n_seg_max = 0;
    for (int k = 0; k < NFSP_MAX_INSERT_REQS; ++k) S.m[k] = Mem{};
    for (int k = 0; k < n; ++k) {
        const nfsp_insert_req &r = reqs[k];
        NFSP_CHECK_ARG(r.d_mem && r.d_total && r.cap > 0, "bad memory %d", k);
        NFSP_CHECK_ARG(r.d_recs && r.d_counts, "null staged batch %d", k);
        NFSP_CHECK_ARG(r.n_segments >= 1 && r.n_segments <= 65535 && r.seg_cap >= 0, "bad segment geometry %d", k);
        if (reservoir) {
            NFSP_CHECK_ARG(r.d_stamp, "reservoir %d has no stamp array", k);
            NFSP_CHECK_ARG(r.mode == 0 || r.mode == 1, "mode must be 0 (Algorithm R) or 1 (reference law)");
        }
        S.m[k].data = (uint4 *)r.d_mem;
        S.m[k].stamp = (unsigned long long *)r.d_stamp;
        S.m[k].cap = (uint64_t)r.cap;
        S.m[k].total = r.d_total;
        S.m[k].B = Batch{(const uint4 *)r.d_recs, r.d_counts, (uint32_t)r.n_segments, (uint64_t)r.seg_cap};
        S.m[k].seed = r.seed;
        S.m[k].mode = r.mode;
        if (r.seg_cap > *seg_cap_max) *seg_cap_max = r.seg_cap;
        if (r.n_segments > *n_seg_max) *n_seg_max = r.n_segments;
    }
    return NFSP_OK;
}

extern "C" int nfsp_ring_insert_multi(const nfsp_insert_req *reqs, int n, void *stream) {
    MemSet S;
    int64_t seg_cap;
    int n_seg;
    const int rc = make_memset(reqs, n, false, S, &seg_cap, &n_seg);
    if (rc != NFSP_OK) return rc;
    if (seg_cap == 0) return NFSP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ring_insert_kernel<<<insert_grid(seg_cap, n_seg, n), kBufThreads, 0, st>>>(S);
    NFSP_LAUNCH_CHECK();
    commit_kernel<<<n, kBufThreads, 0, st>>>(S);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_reservoir_insert_multi(const nfsp_insert_req *reqs, int n, void *stream) {
    MemSet S;
    int64_t seg_cap;
    int n_seg;
    const int rc = make_memset(reqs, n, true, S, &seg_cap, &n_seg);
    if (rc != NFSP_OK) return rc;
    if (seg_cap == 0) return NFSP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid = insert_grid(seg_cap, n_seg, n);
    reservoir_stamp_kernel<<<grid, kBufThreads, 0, st>>>(S);
    NFSP_LAUNCH_CHECK();
    reservoir_write_kernel<<<grid, kBufThreads, 0, st>>>(S);
    NFSP_LAUNCH_CHECK();
    commit_kernel<<<n, kBufThreads, 0, st>>>(S);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_ring_insert(void *d_ring, int64_t cap, uint64_t *d_total, const void *d_recs, uint32_t *d_counts,
                                int n_segments, int64_t seg_cap, void *stream) {
    nfsp_insert_req r{};
    r.d_mem = d_ring; r.cap = cap; r.d_total = d_total; r.d_recs = d_recs; r.d_counts = d_counts;
    r.n_segments = n_segments; r.seg_cap = seg_cap;
    return nfsp_ring_insert_multi(&r, 1, stream);
}

extern "C" int nfsp_reservoir_insert(void *d_res, int64_t cap, uint64_t *d_total, uint64_t *d_stamp,
                                     const void *d_recs, uint32_t *d_counts, int n_segments, int64_t seg_cap,
                                     uint64_t seed, int mode, void *stream) {
    nfsp_insert_req r{};
    r.d_mem = d_res; r.cap = cap; r.d_total = d_total; r.d_stamp = d_stamp; r.d_recs = d_recs; r.d_counts = d_counts;
    r.n_segments = n_segments; r.seg_cap = seg_cap; r.seed = seed; r.mode = mode;
    return nfsp_reservoir_insert_multi(&r, 1, stream);
}

extern "C" int nfsp_sample_indices(uint64_t seed, uint64_t call_idx, const uint64_t *d_total, int64_t cap,
                                   int is_ring, int batch, int64_t *d_idx, uint32_t *d_n_out, void *stream) {
    NFSP_CHECK_ARG(d_total && d_idx && cap > 0, "bad arguments");
    NFSP_CHECK_ARG(batch >= 1 && batch <= kMaxBatch, "batch must be in [1,%d]", kMaxBatch);
    sample_kernel<<<1, kBufThreads, 0, (cudaStream_t)stream>>>(seed, call_idx, d_total, (uint64_t)cap, is_ring, batch,
                                                               d_idx, d_n_out);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_sample_minibatches(const nfsp_sample_req *reqs, int n_reqs, int batch, int64_t *d_idx, uint32_t *d_n_out,
                                       void *stream) {
    NFSP_CHECK_ARG(reqs && n_reqs >= 1 && n_reqs <= NFSP_MAX_SAMPLE_REQS, "1..%d requests", NFSP_MAX_SAMPLE_REQS);
    NFSP_CHECK_ARG(batch >= 1 && batch <= kMaxBatch, "batch must be in [1,%d]", kMaxBatch);
    SampleReqs R;
    for (int m = 0; m < n_reqs; ++m) {
        NFSP_CHECK_ARG(reqs[m].d_mem && reqs[m].d_total && reqs[m].cap > 0, "bad request %d", m);
        NFSP_CHECK_ARG(reqs[m].d_out || d_idx, "request %d has neither an output block nor an index array", m);
        R.mem[m] = (const uint4 *)reqs[m].d_mem;
        R.total[m] = reqs[m].d_total;
        R.cap[m] = (uint64_t)reqs[m].cap;
        R.seed[m] = reqs[m].seed;
        R.call_idx[m] = reqs[m].call_idx;
        R.out[m] = reqs[m].d_out;
        R.is_ring[m] = reqs[m].is_ring;
    }
    sample_gather_kernel<<<dim3((unsigned)n_reqs, 8u, 1u), kBufThreads, 0, (cudaStream_t)stream>>>(R, batch, d_idx, d_n_out);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_gather_rl(const void *d_ring, const int64_t *d_idx, int batch, float *d_s, float *d_a, float *d_r,
                              float *d_s2, float *d_t, void *stream) {
    NFSP_CHECK_ARG(d_ring && d_idx && d_s && d_a && d_r && d_s2 && d_t && batch >= 1, "bad arguments");
    gather_rl_kernel<<<(batch * 65 + kBufThreads - 1) / kBufThreads, kBufThreads, 0, (cudaStream_t)stream>>>(
        (const uint4 *)d_ring, d_idx, batch, d_s, d_a, d_r, d_s2, d_t);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_gather_sl(const void *d_res, const int64_t *d_idx, int batch, float *d_s, float *d_a,
                              void *stream) {
    NFSP_CHECK_ARG(d_res && d_idx && d_s && d_a && batch >= 1, "bad arguments");
    gather_sl_kernel<<<(batch * 33 + kBufThreads - 1) / kBufThreads, kBufThreads, 0, (cudaStream_t)stream>>>(
        (const uint4 *)d_res, d_idx, batch, d_s, d_a);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}
