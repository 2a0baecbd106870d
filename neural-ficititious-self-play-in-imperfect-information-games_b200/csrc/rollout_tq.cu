// K2, production variant: the fused NFSP rollout as a warp-specialised pipeline around tcgen05.
//
// What round 1 measured (profiles/r01, profiles/r02/lds_patterns.txt): the CUDA-core rollout is bound by shared-memory
// wavefronts -- every decision gathers two 256-byte weight rows and 768 bytes of its net's W2, and a warp's LDS.128
// costs 4 wavefronts whether or not its lanes share rows -- and the first tcgen05 rollout lost because its tiles mixed
// the four nets (4x the MMA work, non-uniform W2) and every 128-thread group ran sort -> MMA -> read -> layer 2 ->
// game logic as one dependent chain under three group barriers.  Here the chain is cut into roles that only meet
// through shared-memory queues:
//
//   env warps (23 x 32 games)   one game per lane, the state machine of nfsp_fsm.cuh in registers.  Per step a lane
//                               builds its observation, takes a TICKET in the queue of its net (player x policy) and
//                               writes its operand row (30 input bits + the constant 1 as bf16) straight into row
//                               ticket % 128 of that net's current A tile; then it sleeps on its warp's mbarrier until
//                               its three outputs are back, steps the game and appends the memory records.
//   scheduler (one thread)      watches the four queues; a tile that is full -- or has waited `patience` cycles, so
//                               that the rare best-response nets do not stall everybody -- gets a free TMEM slot and
//                               its 6 MMAs (M128 N64 K16: 2 k-steps x the exact 3-way bf16 split of the fp32 weights).
//                               Tiles are NET-HOMOGENEOUS: no redundant columns, 8 accumulator slots in flight.
//   epilogue warpgroups (2 x 4) wait for a slot, read their 32 rows of pre-activations from TMEM (tcgen05.ld), run
//                               relu + the 64x3 second layer with W2 as CONSTANT-BANK operands (the rows of a warp
//                               share one net, so W2 costs no shared-memory bandwidth at all), the head, and hand the
//                               outputs back: one 16-byte store per row, one mbarrier arrive per owner warp.
//
// Replaces agent.py:118-156 (Agent.play), main.py:28-67 (turn order, re-deal) and newenv.py:76-349 for the rollout.
#include <cuda_bf16.h>

#include "mlp_math.cuh"
#include "nfsp_fsm.cuh"
#include "ptx_helpers.cuh"
#include "rollout_fast.cuh"

namespace nfsp {
namespace tq {

constexpr int kThreads = 1024;
constexpr int kMlpGroups = 2;                     // epilogue warpgroups
// Roles by warp index.  The issue arbiter of an SM sub-partition favours the HIGHER warp ids: the warps that spend most
// of their time polling (epilogue, scheduler) get the low ids, so their retries only take issue slots nobody else wants.
constexpr int kMlpWarp0 = 0;                      // first epilogue warp; a multiple of 4 (TMEM lane quadrant = warp % 4)
constexpr int kSchedWarp = 4 * kMlpGroups;
constexpr int kEnvWarp0 = kSchedWarp + 1;         // warps kEnvWarp0 .. 31 play
constexpr int kEnvWarps = 32 - kEnvWarp0;
constexpr int kSlots = 8;                         // TMEM accumulator slots of 64 columns
constexpr int kBufs = 12;                         // A tile buffers: 4 per average net, 2 per best-response net
constexpr int kTileBytes = 128 * 32 * 2;          // 128 rows x 32 k, bf16
constexpr uint32_t kLBO = 128, kSBO = 512;        // K-major, no swizzle: core matrix = 8 rows x 16 bytes
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr int kBSplitBytes = 256 * 32 * 2;        // one split of the four nets' W1 (image of pack_tc_wide_kernel)
constexpr int kBBytes = 3 * kBSplitBytes;
constexpr uint32_t kSpinLimit = 1u << 22;

// second layer of the four nets as constant-bank operands: [slot] = one handle's nets
struct __align__(16) W2Image {
    float w[4][3][64];  // [net][output][hidden]
    float b[4][4];
};
constexpr int kW2Slots = 16;
__constant__ W2Image c_w2[kW2Slots];

constexpr int kW2Floats = 16 * 4 * 3 * 4 + 16;   // W2 as [16 quads][4 nets][3 outputs][4] + b2 [4][4] (pack_tc_wide_kernel)
constexpr int kImageBytes = kBBytes + kW2Floats * 4;

// control block in shared memory
struct Control {
    uint32_t tail[4];         // tickets handed out per net
    uint32_t gen[kBufs];      // times a tile buffer has been released
    uint32_t filled[kBufs];   // rows written into the buffer's current tile
    uint32_t rel[kBufs];      // epilogue warps that are done with the buffer's owner table
    uint32_t meta[kSlots];    // net | rows << 8 | buffer << 16 of the tile in a TMEM slot
    uint32_t quit, active_env, tmem_base, lanes;  // lanes: live games of the sets being played right now
    uint32_t delivered, pad[3];                   // rows whose outputs the epilogue has handed back
    uint64_t full[kSlots], empty[kSlots], env[2 * kEnvWarps], wbar;
};

constexpr int kOffA = 0;
constexpr int kOffOwner = kOffA + kBufs * kTileBytes;
constexpr int kOffB = kOffOwner + kBufs * 128 * 2;
constexpr int kOffBits = kOffB + ((kImageBytes + 127) / 128) * 128;
constexpr int kOffFsm = kOffBits + 256 * 8 * 16;
constexpr int kOffRes = kOffFsm + ((fsm::kImageBytes + 15) / 16) * 16;
constexpr int kOffCtl = kOffRes + 2 * kEnvWarps * 32 * 16;
constexpr int kSmemBytes = kOffCtl + (int)sizeof(Control) + 16;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
static_assert(kMlpWarp0 % 4 == 0, "epilogue warps must start a warpgroup");

__device__ __forceinline__ uint32_t buf_of(uint32_t net, uint32_t seq) {  // buffer of tile `seq` of `net`
    const uint32_t odd = net & 1u;
    return (net >> 1) * 6u + odd * 4u + (seq & (odd ? 1u : 3u));
}
__device__ __forceinline__ uint32_t gen_of(uint32_t net, uint32_t seq) { return seq >> ((net & 1u) ? 1 : 2); }

__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void st_vol(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }

__device__ __forceinline__ void mbar_arrive(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// One try.  (A suspend-time hint as fourth operand makes the retries rarer but the wake-up much later: 0.61 -> 1.0 ms.)
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0u;
}
// bounded wait: a protocol error must end the kernel, never hang the GPU
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity, uint32_t *err, uint32_t code) {
    for (uint32_t it = 0; it < kSpinLimit; ++it)
        if (mbar_try(bar, parity)) return true;
    atomicOr(err, code);
    return false;
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0u;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(kLBO >> 4) << 16) | ((uint64_t)(kSBO >> 4) << 32) | (1ull << 46);
}
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

__device__ __forceinline__ uint4 bits_to_bf16x8(uint32_t b) {  // 8 input bits -> 8 bf16 values 0.0 / 1.0
    uint4 c;
    c.x = ((b >> 0) & 1u) * 0x3F80u + ((b >> 1) & 1u) * 0x3F800000u;
    c.y = ((b >> 2) & 1u) * 0x3F80u + ((b >> 3) & 1u) * 0x3F800000u;
    c.z = ((b >> 4) & 1u) * 0x3F80u + ((b >> 5) & 1u) * 0x3F800000u;
    c.w = ((b >> 6) & 1u) * 0x3F80u + ((b >> 7) & 1u) * 0x3F800000u;
    return c;
}

// ---- second layer + head of 32 rows of ONE net (agent.py:102-103,111-112) -----------------------------------------
// The 64 pre-activations of this lane's row come from TMEM 16 columns at a time; W2 and b2 are constant-bank operands
// (the net is a template parameter, the handle's slot a kernel argument: uniform-register loads, no LDS).  The TMEM
// slot is handed back as soon as the last columns are in registers.
// W2 comes from shared memory with a WARP-UNIFORM address (the rows of a warp share one net): 48 LDS.128 of 2 wavefronts
// each per 32 rows, against 4 wavefronts each and 80 of them in the CUDA-core rollout.  (The constant bank would cost no
// shared-memory bandwidth at all, and ptxas does emit uniform-datapath LDCU for it when the net comes out of a warp
// reduction -- but only while this code sits in a convergent region; inside the role branches of this kernel it falls
// back to one register-indexed LDC per lane and value, 4 800 cycles per 32 rows.  profiles/r02/epilogue_w2_paths.txt)
__device__ __forceinline__ float4 rows_forward(const float4 *__restrict__ w2, const float4 b2, bool is_br, uint32_t taddr, uint32_t empty_bar) {
    Layer2Acc acc;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        float h[16];
        tmem_ld16(taddr + ch * 16, h);
        if (ch == 3) {
            tc_fence_before();
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(empty_bar, 1u);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            acc.quad(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3], w2[(ch * 4 + q) * 12], w2[(ch * 4 + q) * 12 + 1],
                     w2[(ch * 4 + q) * 12 + 2]);
    }
    float o0, o1, o2;
    acc.head(b2, is_br, o0, o1, o2);
    return make_float4(o0, o1, o2, 0.f);
}

// The same with W2 / b2 as constant-bank operands: `slot` and `net` must be values the compiler KNOWS to be uniform
// (kernel argument; result of a warp reduction) and the call must sit in convergent code -- then the reads are
// uniform-datapath LDCU c[3][UR + imm] feeding FFMA2 directly and the second layer costs no shared-memory bandwidth.
__device__ __forceinline__ float4 rows_forward_const(uint32_t slot, uint32_t net_v, uint32_t taddr, uint32_t empty_bar) {
    // its own reduction, used for nothing but these addresses: a value that also feeds per-lane arithmetic is copied
    // to a vector register first and the loads follow it there
    const uint32_t net = __reduce_max_sync(0xFFFFFFFFu, net_v) & 3u;
    const W2Image &I = c_w2[slot];
    unsigned long long z0 = 0ull, z1 = 0ull, z2 = 0ull;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        float h[16];
        tmem_ld16(taddr + ch * 16, h);
        if (ch == 3) {
            tc_fence_before();
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(empty_bar, 1u);
        }
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const int u = ch * 16 + j;
            const unsigned long long hh = pack2(fmaxf(h[j], 0.f), fmaxf(h[j + 1], 0.f));
            fma2(z0, hh, pack2(I.w[net][0][u], I.w[net][0][u + 1]));
            fma2(z1, hh, pack2(I.w[net][1][u], I.w[net][1][u + 1]));
            fma2(z2, hh, pack2(I.w[net][2][u], I.w[net][2][u + 1]));
        }
    }
    const float y0 = hsum2(z0) + I.b[net][0], y1 = hsum2(z1) + I.b[net][1], y2 = hsum2(z2) + I.b[net][2];
    float o0, o1, o2;
    if (net & 1u) {  // best-response head: relu (agent.py:103)
        o0 = fmaxf(y0, 0.f); o1 = fmaxf(y1, 0.f); o2 = fmaxf(y2, 0.f);
    } else {         // average-policy head: softmax with max subtraction (agent.py:112)
        const float m = fmaxf(y0, fmaxf(y1, y2));
        const float e0 = expf(y0 - m), e1 = expf(y1 - m), e2 = expf(y2 - m);
        const float inv = 1.0f / (e0 + e1 + e2);
        o0 = e0 * inv; o1 = e1 * inv; o2 = e2 * inv;
    }
    return make_float4(o0, o1, o2, 0.f);
}

#ifdef NFSP_TQ_PROF
#define TQ_CLK() clock64()
#define TQ_DECL() long long tq_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define TQ_ADD(k, v) tq_acc[(k) & 7] += (long long)(v)   /* local accumulators: role r owns counters 8r .. 8r+7 */
#define TQ_FLUSH(r) do { if ((threadIdx.x & 31) == 0) for (int q_ = 0; q_ < 8; ++q_) atomicAdd(T.prof + 8 * (r) + q_, (unsigned long long)tq_acc[q_]); } while (0)
#else
#define TQ_CLK() 0ll
#define TQ_DECL() do { } while (0)
#define TQ_ADD(k, v) do { } while (0)
#define TQ_FLUSH(r) do { } while (0)
#endif

struct TqArgs {
    RolloutArgs R;
    uint32_t w2_slot;          // constant-bank slot of this handle's second layers
    unsigned long long *prof;  // NFSP_TQ_PROF builds: cycle counters, see nfsp_rollout_profile
    const uint32_t *fsm_image;
    uint32_t *err;
    uint32_t patience_avg, patience_br;  // cycles a partly filled tile may wait for more rows
};

// warp-aggregated append of one step's records to the staging segments (same order rules as rollout_fast.cuh)
__device__ __forceinline__ int append_records(const WarpStage &W, uint32_t q, bool vA, const uint4 &recA, bool vB, const uint4 &recB,
                                              bool vC, const uint4 &recC, bool vS, const uint4 &recS) {
    const uint32_t sh = q * 8u;
    const uint32_t mine = ((uint32_t)vA + (uint32_t)vB) << sh | (uint32_t)vC << (8u - sh) | (uint32_t)vS << (16u + sh);
    uint32_t incl = mine;
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        incl += lane >= (uint32_t)o ? up : 0u;
    }
    const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
    uint32_t base = 0;
    if (lane < 4) {
        const uint32_t t = (tot >> (8u * lane)) & 0xFFu;
        if (t) base = atomicAdd(W.cnt + lane * W.n_seg, t);
    }
    const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, base, 0), b1 = __shfl_sync(0xFFFFFFFFu, base, 1);
    const uint32_t b2 = __shfl_sync(0xFFFFFFFFu, base, 2), b3 = __shfl_sync(0xFFFFFFFFu, base, 3);
    const uint32_t excl = incl - mine;
    uint4 *rp = q ? W.rl1 : W.rl0, *ro = q ? W.rl0 : W.rl1, *sp = q ? W.sl1 : W.sl0;
    uint32_t off = (q ? b1 : b0) + ((excl >> sh) & 0xFFu);
    int drop = 0;
    if (vA) {
        if (off < W.cap_rl) rp[off] = recA; else ++drop;
        ++off;
    }
    if (vB) { if (off < W.cap_rl) rp[off] = recB; else ++drop; }
    if (vC) {
        const uint32_t o2 = (q ? b0 : b1) + ((excl >> (8u - sh)) & 0xFFu);
        if (o2 < W.cap_rl) ro[o2] = recC; else ++drop;
    }
    if (vS) {
        const uint32_t o3 = (q ? b3 : b2) + ((excl >> (16u + sh)) & 0xFFu);
        if (o3 < W.cap_sl) sp[o3] = recS; else ++drop;
    }
    return drop;
}

// what an env warp needs besides the kernel arguments
struct EnvCtx {
    Control *ctl;
    uint8_t *smem;
    const uint4 *bits;  // byte -> 8 x bf16 table, this lane's skewed copy
    uint32_t fsm_base, dl_sum, rew_base, lane;
    int64_t plane;
};

constexpr int kSets = 2;  // game sets per env warp

// 32 games (one per lane) of an env warp: the state machine of nfsp_fsm.cuh in registers + what Agent.play keeps
// between its observation and its step.
struct GameSet {
    fsm::Game g;
    uint32_t SA, SO;   // snapshots s[p] of the player to act / the other one (newenv.py:200-202)
    uint32_t q;        // player to act
    uint32_t obs;      // its observation of this step
    uint32_t flags;    // 1 live, 2 policy 'b' this step, 4 epsilon branch, 8 hand started at this step (<< 21 in the trace)
    uint32_t owner;    // result slot = index of this set's mbarrier * 32 + lane
    uint32_t phase, seg, n_live;
    int64_t i;
    bool any;          // the set holds at least one game
    long long tq_wait;

    __device__ __forceinline__ void load(const TqArgs &T, const EnvCtx &X, int64_t blk, uint32_t bar_idx) {
        const RolloutArgs &A = T.R;
        i = (blk << 5) + X.lane;
        const bool live = i < A.n;
        n_live = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, live));
        any = n_live != 0u;
        flags = live ? 1u : 0u;
        owner = bar_idx * 32u + X.lane;
        seg = (uint32_t)blk & (A.n_seg - 1u);
        const uint64_t w = live ? A.state[i] : 0ull;
        const NfspW gw{w};
        g.unpack(w, X.fsm_base);
        q = (uint32_t)gw.to_act();
        SA = gw.snapshot((int)q);
        SO = gw.snapshot((int)(q ^ 1u));
        obs = 0u;
        tq_wait = 0;
    }

    __device__ __forceinline__ void store(const TqArgs &T, const EnvCtx &X, FastCounters &c) {
        if (flags & 1u) {
            T.R.state[i] = g.pack(X.fsm_base);
            c.wide.trans += T.R.n_steps;
        }
    }

    // agent.py:130-141 up to the network: re-deal, observe, choose the net, submit the operand row
    template <bool kDebug>
    __device__ __forceinline__ void begin(const TqArgs &T, const EnvCtx &X, FastCounters &c, int t) {
        if (!any) return;
        const RolloutArgs &A = T.R;
        Control *ctl = X.ctl;
        const bool live = flags & 1u;
        const uint64_t game = A.game0 + (uint64_t)i;
        const Philox4 x = game_block(A.keys, game, A.step0 + (uint64_t)t, STREAM_STEP);
        flags &= 1u;
        if (g.HX & fsm::kHxOver) {  // newenv.py:76-114 + main.py:28-45: the other player deals
            g.dl = X.dl_sum - g.dl;
            const uint4 D = fsm::lds128(g.dl + __umulhi(x.y, 120u) * 16u);
            uint32_t pa = g.dl;
            if (x.z < A.eta_u32) pa += 16u;
            if (x.w < A.eta_u32) pa += 32u;
            const uint4 P = fsm::lds128(pa + (uint32_t)fsm::kPolOff);
            g.PA = D.x | P.x;
            g.PO = D.y | P.y;
            g.ms = D.z | P.z;
            g.tix = P.w;
            g.HX = fsm::kCm0;
            q = (D.z >> 5) & 1u;  // the dealer opens
            SA = 0u;
            SO = 0u;
            flags |= 8u;
            c.wide.hands += live;
        }
        obs = g.HX & (g.PA | 0x00FFFFFFu) & 0x3FFFFFFFu;
        const uint32_t pol = (g.PA >> 12) & 1u;
        flags |= pol << 1;
        if (pol && x.x < (q ? A.eps1_u32 : A.eps_u32)) flags |= 4u;
        const uint32_t net = q * 2u + pol;
        // ---- ticket in the net's queue, operand row into the tile ------------------------------------------------
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, live ? net : 4u);
        const uint32_t rank = (uint32_t)__popc(peers & ((1u << X.lane) - 1u));
        const uint32_t leader = (uint32_t)__ffs(peers) - 1u, cnt = (uint32_t)__popc(peers);
        uint32_t tb = 0u;
        if (live && X.lane == leader) tb = atomicAdd(&ctl->tail[net], cnt);
        tb = __shfl_sync(0xFFFFFFFFu, tb, leader);
        if (live) {
            const uint32_t ticket = tb + rank, seq = ticket >> 7, row = ticket & 127u;
            const uint32_t bid = buf_of(net, seq), need = gen_of(net, seq);
            uint32_t it = 0;
            while (ld_vol(&ctl->gen[bid]) != need) {  // the buffer still holds an older tile
                __nanosleep(64);
                if (++it > kSpinLimit) { atomicOr(T.err, 1u); break; }
            }
            uint8_t *rowp = X.smem + kOffA + bid * kTileBytes + (row >> 3) * kSBO + (row & 7u) * 16u;
            const uint32_t xin = obs | (1u << 30);  // input 30 = the constant 1 that carries b1
#pragma unroll
            for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4 *>(rowp + k * kLBO) = X.bits[((xin >> (8 * k)) & 0xFFu) * 8u];
            reinterpret_cast<uint16_t *>(X.smem + kOffOwner)[bid * 128u + row] = (uint16_t)owner;
            fence_async_smem();
        }
        __syncwarp();
        if (live && X.lane == leader) {  // the group's rows are written: count them in (at most two tiles)
            const uint32_t seq0 = tb >> 7, first = min(cnt, 128u - (tb & 127u));
            __threadfence_block();
            atomicAdd(&ctl->filled[buf_of(net, seq0)], first);
            if (cnt > first) atomicAdd(&ctl->filled[buf_of(net, seq0 + 1u)], cnt - first);
        }
        if (X.lane == 0 && n_live < 32u) mbar_arrive(smem_u32(&ctl->env[owner >> 5]), 32u - n_live);
    }

    // agent.py:142-156 after the network + newenv.py:192-349 + main.py:55-67 + the memory records
    template <bool kDebug>
    __device__ __forceinline__ void finish(const TqArgs &T, const EnvCtx &X, FastCounters &c, int t) {
        if (!any) return;
        const RolloutArgs &A = T.R;
        const bool live = flags & 1u;
        const uint64_t game = A.game0 + (uint64_t)i;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
        if (flags & 4u) {  // agent.py:125-128, np.random.rand(1,1,3): rare, its own Philox block
            const Philox4 y = game_block(A.keys, game, A.step0 + (uint64_t)t, STREAM_VECTOR);
            r0 = (float)(y.x >> 8) * (1.0f / 16777216.0f);
            r1 = (float)(y.y >> 8) * (1.0f / 16777216.0f);
            r2 = (float)(y.z >> 8) * (1.0f / 16777216.0f);
        }
        const long long w0 = TQ_CLK();
        mbar_wait_bounded(smem_u32(&X.ctl->env[owner >> 5]), phase, T.err, 2u);
        tq_wait = TQ_CLK() - w0;
        phase ^= 1u;
        const float4 res = reinterpret_cast<const float4 *>(X.smem + kOffRes)[owner];
        float v0 = res.x, v1 = res.y, v2 = res.z;
        if (flags & 4u) { v0 = r0; v1 = r1; v2 = r2; }
        const int64_t at = (int64_t)t * A.n + i;
        if (kDebug && live) {
            if (A.vec) { A.vec[3 * at] = v0; A.vec[3 * at + 1] = v1; A.vec[3 * at + 2] = v2; }
            if (A.forced) { v0 = A.forced[3 * at]; v1 = A.forced[3 * at + 1]; v2 = A.forced[3 * at + 2]; }
        }
        const bool vA = live && ((g.PA >> 15) & 1u);  // agent.py:132-136: remember the previous transition
        const uint4 recA = make_uint4(SA, obs, 0u, ((g.PA >> 13) & 3u) | (q << 16));
        uint32_t a = 0u;  // np.argmax: first maximum
        float best = v0;
        if (v1 > best) { a = 1u; best = v1; }
        if (v2 > best) a = 2u;
        const bool nz = (v0 != 0.f) || (v1 != 0.f) || (v2 != 0.f);
        const uint32_t ea = g.tix + a * (uint32_t)fsm::kEntryBytes;
        const uint4 E = fsm::lds128(ea);        // hc, pa, pm, nx
        const uint4 M = fsm::lds128(ea + 16u);  // mw, mb, ma, mo
        g.PA = (g.PA & ~fsm::kPClear) + E.y - (nz ? 0u : 0x8000u);
        g.HX = (g.HX & 0xFFFFFFu) | E.x;
        c.small += live ? 1u << (5u * (3u * q + a)) : 0u;
        const bool term = (g.HX & fsm::kHxOver) != 0u;
        const uint32_t rew_a = fsm::lds32(X.rew_base + ((g.PA & M.z) | (g.PO & M.w)));  // 0 while the hand is live
        bool vB = false, vC = false;
        uint4 recB = make_uint4(0, 0, 0, 0), recC = make_uint4(0, 0, 0, 0);
        if (term) {  // main.py:55-67: both players observe the terminal state once
            const float ra = __uint_as_float(rew_a), ro = 0.f - ra;  // zero-sum (newenv.py:250-298)
            const int ra2 = __float2int_rn(ra + ra);
            c.wide.rew0 += live ? (q ? -ra2 : ra2) : 0;
            c.wide.rew1 += live ? (q ? ra2 : -ra2) : 0;
            vB = live && nz;
            recB = make_uint4(obs, g.HX & (g.PA | 0x00FFFFFFu) & 0x3FFFFFFFu, rew_a, a | (1u << 8) | (q << 16));
            vC = live && ((g.PO >> 15) & 1u);
            recC = make_uint4(SO, g.HX & (g.PO | 0x00FFFFFFu) & 0x3FFFFFFFu, __float_as_uint(ro),
                              ((g.PO >> 13) & 3u) | (1u << 8) | ((q ^ 1u) << 16));
        }
        if (kDebug) {
            g.ms += M.y;
            if (live && A.trace) {
                A.trace[at] = g.HX & (g.PA | 0xC0FFFFFFu);
                A.trace[X.plane + at] = rew_a;
                A.trace[2 * X.plane + at] = g.ms | M.x | ((flags & 8u) << 18);
            }
        }
        const bool vS = live && (flags & 2u);
        const uint4 recS = make_uint4(obs, __float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2));
        WarpStage W;
        W.init(A, seg);
        c.wide.drop += append_records(W, q, vA, recA, vB, recB, vC, recC, vS, recS);
        // the turn passes (or not): swap the players' words and snapshots
        const uint32_t pa = g.PA, sa = obs;
        g.PA = (g.PO & E.z) | (pa & ~E.z);
        g.PO = (pa & E.z) | (g.PO & ~E.z);
        SA = (SO & E.z) | (sa & ~E.z);
        SO = (sa & E.z) | (SO & ~E.z);
        q ^= E.z & 1u;
        g.tix = E.w;
    }
};

template <bool kDebug>
__global__ void __launch_bounds__(kThreads, 1)
rollout_tq_kernel(const TqArgs T) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ unsigned long long s_stats[NFSP_STATS_FIELDS];
    const RolloutArgs &A = T.R;
    Control *ctl = reinterpret_cast<Control *>(smem + kOffCtl);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

    // ---- set-up: control block, barriers, tables, weight image, TMEM -------------------------------------------
    if (threadIdx.x < NFSP_STATS_FIELDS) s_stats[threadIdx.x] = 0ull;
    for (uint32_t e = threadIdx.x; e < sizeof(Control) / 4; e += blockDim.x) reinterpret_cast<uint32_t *>(ctl)[e] = 0u;
    for (uint32_t e = threadIdx.x; e < 2048u; e += blockDim.x) reinterpret_cast<uint4 *>(smem + kOffBits)[e] = bits_to_bf16x8(e >> 3);
    const uint32_t fsm_base = smem_u32(smem + kOffFsm);
    for (int w = threadIdx.x; w < fsm::kImageWords; w += blockDim.x)
        reinterpret_cast<uint32_t *>(smem + kOffFsm)[w] = T.fsm_image[w] + (fsm::is_address(w) ? fsm_base : 0u);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < kSlots; ++k) { mbar_init(smem_u32(&ctl->full[k]), 1); mbar_init(smem_u32(&ctl->empty[k]), 4); }
        for (int k = 0; k < 2 * kEnvWarps; ++k) mbar_init(smem_u32(&ctl->env[k]), 32);
        mbar_init(smem_u32(&ctl->wbar), 1);
        ctl->active_env = kEnvWarps;
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(smem_u32(&ctl->wbar), kImageBytes);
        bulk_g2s(smem_u32(smem + kOffB), A.pack, kImageBytes, smem_u32(&ctl->wbar));
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ld_vol(&ctl->tmem_base);

    FastCounters c;
    if (warp >= (uint32_t)kEnvWarp0) {
        // ================================ env warps ================================================================
        // Two independent sets of 32 games per warp: while the requests of one set are in flight (queue, MMAs,
        // epilogue: several thousand cycles) the warp finishes and re-submits the other one.
        EnvCtx X;
        X.ctl = ctl;
        X.smem = smem;
        X.fsm_base = fsm_base;
        X.dl_sum = 2u * (fsm_base + (uint32_t)fsm::kDealOff) + (uint32_t)fsm::kDealHalf;
        X.rew_base = fsm_base + (uint32_t)fsm::kRewardOff;
        X.bits = reinterpret_cast<const uint4 *>(smem + kOffBits) + (lane & 7u);
        X.lane = lane;
        X.plane = (int64_t)A.n_steps * A.n;
        TQ_DECL();
        const long long tq_t0 = TQ_CLK();
        const int64_t n_blocks = (A.n + 31) >> 5;
        GameSet S[kSets];
#pragma unroll
        for (int k = 0; k < kSets; ++k) S[k].phase = 0u;
        for (;;) {
            uint32_t blk = 0;
            if (lane == 0) blk = atomicAdd(A.work, (uint32_t)kSets);
            blk = __shfl_sync(0xFFFFFFFFu, blk, 0);
            if ((int64_t)blk >= n_blocks) break;
#pragma unroll
            for (int k = 0; k < kSets; ++k) {
                S[k].load(T, X, (int64_t)blk + k, (uint32_t)k * (uint32_t)kEnvWarps + warp - (uint32_t)kEnvWarp0);
                if (lane == 0 && S[k].any) atomicAdd(&ctl->lanes, S[k].n_live);
            }
#pragma unroll
            for (int k = 0; k < kSets; ++k) S[k].template begin<kDebug>(T, X, c, 0);
            for (int t = 0; t < A.n_steps; ++t) {
#pragma unroll
                for (int k = 0; k < kSets; ++k) {
                    const long long tq_t2 = TQ_CLK();
                    S[k].template finish<kDebug>(T, X, c, t);
                    TQ_ADD(2, S[k].tq_wait);
                    const long long tq_t3 = TQ_CLK();
                    if (t + 1 < A.n_steps) S[k].template begin<kDebug>(T, X, c, t + 1);
                    TQ_ADD(1, TQ_CLK() - tq_t3);
                    TQ_ADD(3, S[k].any ? 1 : 0);
                    (void)tq_t2;
                }
                if ((t & 7) == 7) c.spill();  // six 5-bit counters, two sets: at most 16 increments between spills
            }
            c.spill();
#pragma unroll
            for (int k = 0; k < kSets; ++k) {
                S[k].store(T, X, c);
                if (lane == 0 && S[k].any) atomicSub(&ctl->lanes, S[k].n_live);
            }
        }
        c.spill();
        TQ_ADD(0, TQ_CLK() - tq_t0);
        TQ_FLUSH(0);
        if (lane == 0) atomicSub(&ctl->active_env, 1u);
    } else if (warp == (uint32_t)kSchedWarp) {
        // ================================ scheduler ================================================================
        // The whole warp runs this loop in lockstep (a single divergent thread makes every tcgen05.mma an
        // elect / R2UR.BROADCAST waterfall: 190 cycles per MMA measured).  Lane n < 4 watches the queue of net n; tiles
        // that are ready are then issued one by one with warp-uniform operands by an elected lane.
        mbar_wait_bounded(smem_u32(&ctl->wbar), 0u, T.err, 4u);  // the weight image has landed
        const uint32_t a_smem = smem_u32(smem + kOffA), b_smem = smem_u32(smem + kOffB);
        uint32_t issued = 0u, closed = 0u, seen = 0u, padding = 0u;  // of this lane's net
        uint32_t k = 0u;
        TQ_DECL();
        const long long tq_s0 = TQ_CLK();
        for (;;) {
            TQ_ADD(6, 1);
            const uint32_t tail = lane < 4u ? ld_vol(&ctl->tail[lane]) : 0u;
            // every game in play already waits for its outputs: no tile can gain a row, so none should wait
            const uint32_t asked = __reduce_add_sync(0xFFFFFFFFu, tail - padding);
            const bool drained = asked - ld_vol(&ctl->delivered) >= ld_vol(&ctl->lanes);
            uint32_t rows = 0u, bid = 0u;
            if (lane < 4u) {
                rows = closed;
                if (!rows) {
                    const uint32_t avail = tail - (issued << 7);
                    if (avail == 0u) {
                        seen = 0u;
                    } else if (avail >= 128u) {
                        rows = 128u;
                    } else {  // a partly filled tile goes out once it has waited long enough
                        const uint32_t now = (uint32_t)clock64() | 1u;
                        if (!seen) seen = now;
                        if ((drained || now - seen >= ((lane & 1u) ? T.patience_br : T.patience_avg)) &&
                            atomicCAS(&ctl->tail[lane], tail, (issued + 1u) << 7) == tail) {
                            rows = avail;
                            padding += 128u - avail;
                        }
                    }
                    closed = rows;
                }
                if (rows) {
                    bid = buf_of(lane, issued);
                    if (ld_vol(&ctl->filled[bid]) != rows) rows = 0u;  // its writers are still busy
                    else __threadfence_block();
                }
            }
            uint32_t ready = __ballot_sync(0xFFFFFFFFu, rows != 0u);
            if (!ready) {
                if (ld_vol(&ctl->active_env) == 0u) break;
                __nanosleep(32);
                continue;
            }
            while (ready) {
                const uint32_t net = (uint32_t)__ffs(ready) - 1u;
                ready &= ready - 1u;
                const uint32_t job = __reduce_max_sync(0xFFFFFFFFu, lane == net ? (rows | (bid << 8)) : 0u);  // uniform
                const uint32_t rows_u = job & 0xFFu, bid_u = job >> 8;
                const uint32_t slot = k & (kSlots - 1u);
                const long long tq_s1 = TQ_CLK();
                if (k >= (uint32_t)kSlots) mbar_wait_bounded(smem_u32(&ctl->empty[slot]), ((k >> 3) - 1u) & 1u, T.err, 8u);
                TQ_ADD(1, TQ_CLK() - tq_s1);
                TQ_ADD(2, 1);
                TQ_ADD(3, rows_u);
                const long long tq_s2 = TQ_CLK();
                if (lane == 0) {
                    st_vol(&ctl->meta[slot], net | (rows_u << 8) | (bid_u << 16));
                    st_vol(&ctl->filled[bid_u], 0u);
                    __threadfence_block();
                }
                __syncwarp();
                tc_fence_after();
                const long long tq_s3 = TQ_CLK();
                TQ_ADD(5, tq_s3 - tq_s2);
                if (elect_one()) {
                    const uint32_t tacc = tmem_base + slot * 64u;
#pragma unroll
                    for (int split = 0; split < 3; ++split)
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            umma_f16(tacc, umma_desc(a_smem + bid_u * kTileBytes + ks * 2 * kLBO),
                                     umma_desc(b_smem + split * kBSplitBytes + net * 4096u + ks * 2 * kLBO), (split | ks) != 0);
                    umma_commit(smem_u32(&ctl->full[slot]));
                }
                __syncwarp();
                TQ_ADD(4, TQ_CLK() - tq_s3);
                if (lane == net) {
                    ++issued;
                    closed = 0u;
                    seen = 0u;
                }
                ++k;
            }
        }
        TQ_ADD(0, TQ_CLK() - tq_s0);
        TQ_FLUSH(1);
        if (lane == 0) st_vol(&ctl->quit, 1u);
    } else {
        // ================================ epilogue warpgroups ======================================================
        const uint32_t grp = (warp - (uint32_t)kMlpWarp0) >> 2, wq = warp & 3u;
        float4 *res = reinterpret_cast<float4 *>(smem + kOffRes);
        const float4 *sW2 = reinterpret_cast<const float4 *>(smem + kOffB + kBBytes);
        mbar_wait_bounded(smem_u32(&ctl->wbar), 0u, T.err, 4u);  // the weight image has landed
        const uint16_t *owners = reinterpret_cast<const uint16_t *>(smem + kOffOwner);
        TQ_DECL();
        const long long tq_e0 = TQ_CLK();
        for (uint32_t k = grp;; k += (uint32_t)kMlpGroups) {
            const uint32_t slot = k & (kSlots - 1u), parity = (k >> 3) & 1u;
            const uint32_t full = smem_u32(&ctl->full[slot]);
            bool quit = false;
            const long long tq_e1 = TQ_CLK();
            for (uint32_t it = 0;; ++it) {
                if (mbar_try(full, parity)) break;
                if (ld_vol(&ctl->quit)) { quit = true; break; }
                if (it > kSpinLimit) { atomicOr(T.err, 16u); quit = true; break; }
            }
            if (quit) break;
            TQ_ADD(1, TQ_CLK() - tq_e1);
            TQ_ADD(2, 1);
            tc_fence_after();
            const uint32_t meta = ld_vol(&ctl->meta[slot]);
            const uint32_t net = meta & 0xFFu, rows = (meta >> 8) & 0xFFu, bid = meta >> 16;
            const uint32_t row = wq * 32u + lane;
            const bool valid = row < rows;
            const uint32_t owner = valid ? (uint32_t)owners[bid * 128u + row] : 0xFFFFu;
            __syncwarp();
            if (lane == 0) {  // this warp is done with the tile's buffer; the fourth one releases it to the writers
                if (atomicAdd(&ctl->rel[bid], 1u) == 3u) {
                    st_vol(&ctl->rel[bid], 0u);
                    __threadfence_block();
                    atomicAdd(&ctl->gen[bid], 1u);
                }
            }
            const uint32_t empty = smem_u32(&ctl->empty[slot]);
            if (wq * 32u >= rows) {  // nothing but padding in this warp's rows
                if (lane == 0) mbar_arrive(empty, 1u);
                continue;
            }
            const uint32_t taddr = tmem_base + slot * 64u + (wq << 21);
#ifndef NFSP_TQ_W2_CONST
            const float4 out = rows_forward(sW2 + net * 3u, sW2[16 * 12 + net], (net & 1u) != 0u, taddr, empty);
#else
            const float4 out = rows_forward_const(T.w2_slot, net, taddr, empty);
#endif
            if (valid) res[owner] = out;
            const uint32_t ow = owner >> 5;  // rows of one owner warp are neighbours: one arrive per owner warp
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, ow);
            __syncwarp();
            if (valid && lane == (uint32_t)__ffs(peers) - 1u) mbar_arrive(smem_u32(&ctl->env[ow]), (uint32_t)__popc(peers));
            if (lane == 0) atomicAdd(&ctl->delivered, min(32u, rows - wq * 32u));
            TQ_ADD(3, 1);
        }
        TQ_ADD(0, TQ_CLK() - tq_e0);
        TQ_FLUSH(2);
    }
    tc_fence_before();
    __syncthreads();
    if (A.stats) c.wide.commit(s_stats, A.stats);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace tq
}  // namespace nfsp

using namespace nfsp;

namespace nfsp {
namespace tq {
// a handle's W2 image from the flat weights: [4][2179] -> W2Image
__global__ void pack_w2_kernel(const float *__restrict__ w, W2Image *__restrict__ img) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 4 * 3 * 64 + 16; e += gridDim.x * blockDim.x) {
        if (e < 4 * 3 * 64) {
            const int net = e / 192, c = (e / 64) % 3, j = e % 64;
            img->w[net][c][j] = w[net * NFSP_NET_PARAMS + 1984 + j * 3 + c];
        } else {
            const int f = e - 4 * 3 * 64, net = f >> 2, c = f & 3;
            img->b[net][c] = c < 3 ? w[net * NFSP_NET_PARAMS + 2176 + c] : 0.f;
        }
    }
}
}  // namespace tq
}  // namespace nfsp

// ---- constant-bank slots: one per handle with acting nets, per device --------------------------------------------
static uint32_t g_slot_mask[64] = {0};  // bit s set: slot s of that device is taken

// called from nfsp_act_set_weights: set-up of the warp-specialised rollout for this handle + its second-layer image
int nfsp_tq_set_weights(nfsp_env_t h, const float *d_weights, cudaStream_t st) {
    if (h->w2_slot < 0) {
        NFSP_CHECK_ARG(h->device >= 0 && h->device < 64, "device index out of range");
        int s = 0;
        while (s < tq::kW2Slots && (g_slot_mask[h->device] >> s & 1u)) ++s;
        if (s == tq::kW2Slots)
            return set_error(NFSP_E_STATE, "more than %d handles with acting nets on device %d", tq::kW2Slots, h->device);
        g_slot_mask[h->device] |= 1u << s;
        h->w2_slot = s;
        NFSP_CUDA(cudaMalloc(&h->d_w2img, sizeof(tq::W2Image)));
        NFSP_CUDA(cudaMalloc(&h->d_err, 8 + 24 * sizeof(unsigned long long)));
        NFSP_CUDA(cudaMemset(h->d_err, 0, 8 + 24 * sizeof(unsigned long long)));
        NFSP_CUDA(cudaFuncSetAttribute(tq::rollout_tq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tq::kSmemBytes));
        NFSP_CUDA(cudaFuncSetAttribute(tq::rollout_tq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tq::kSmemBytes));
    }
    tq::pack_w2_kernel<<<4, 256, 0, st>>>(d_weights, (tq::W2Image *)h->d_w2img);
    NFSP_LAUNCH_CHECK();
    NFSP_CUDA(cudaMemcpyToSymbolAsync(tq::c_w2, h->d_w2img, sizeof(tq::W2Image), (size_t)h->w2_slot * sizeof(tq::W2Image),
                                      cudaMemcpyDeviceToDevice, st));
    return NFSP_OK;
}

void nfsp_tq_release(nfsp_env_t h) {
    if (h->w2_slot >= 0 && h->device >= 0 && h->device < 64) g_slot_mask[h->device] &= ~(1u << h->w2_slot);
    h->w2_slot = -1;
    if (h->d_w2img) cudaFree(h->d_w2img);
    if (h->d_err) cudaFree(h->d_err);
    h->d_w2img = nullptr;
    h->d_err = nullptr;
}

int nfsp_rollout_tq_launch(nfsp_env_t h, const nfsp::RolloutArgs &A0, bool debug, int reserve_sms, cudaStream_t st) {
    tq::TqArgs T;
    T.R = A0;
    T.R.pack = h->d_wtc_wide;
    T.fsm_image = h->d_fsm;
    T.err = h->d_err;
    T.w2_slot = (uint32_t)h->w2_slot;
    T.prof = reinterpret_cast<unsigned long long *>(h->d_err + 2);
    T.patience_avg = h->tq_patience[0];
    T.patience_br = h->tq_patience[1];
    NFSP_CUDA(cudaMemsetAsync(h->d_work, 0, sizeof(uint32_t), st));
    const int64_t ctas = (h->n + tq::kEnvWarps * 32 - 1) / (tq::kEnvWarps * 32);
    const int sms = h->sm_count - reserve_sms;
    const int grid = (int)(ctas < sms ? ctas : sms);
    if (debug) tq::rollout_tq_kernel<true><<<grid, tq::kThreads, tq::kSmemBytes, st>>>(T);
    else tq::rollout_tq_kernel<false><<<grid, tq::kThreads, tq::kSmemBytes, st>>>(T);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

// NFSP_TQ_PROF builds only: the kernel's cycle counters since the handle's first launch (all zero otherwise):
// env: 0 total, 1 ticket + operand row, 2 waiting for the outputs, 3 set-steps; scheduler: 8 total, 9 waiting for a
// TMEM slot, 10 tiles, 11 rows, 12 MMA issue + commit, 13 fences + meta, 14 loop iterations; epilogue: 16 total,
// 17 waiting for a tile, 18 tiles x warps, 19 computed
extern "C" int nfsp_rollout_profile(nfsp_env_t h, uint64_t *out24) {
    NFSP_CHECK_ARG(h != nullptr && out24 != nullptr, "null argument");
    for (int k = 0; k < 24; ++k) out24[k] = 0;
    if (!h->d_err) return NFSP_OK;
    DeviceGuard guard(h->device);
    NFSP_CUDA(cudaMemcpy(out24, h->d_err + 2, 24 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return NFSP_OK;
}
