// K1: batched Leduc environments (both rule sets) -- one thread per game, the packed word lives in
// a register across the steps of a launch, HBM sees one coalesced 8-byte load + store per game per
// launch plus the (optional) SoA trace planes.  Replaces leduc/env.py, leduc/newenv.py,
// leduc/deck.py and the turn-order part of main.py:28-67.
#include "common.cuh"
#include "legacy_rules.cuh"
#include "nfsp_fast.cuh"
#include "legacy_fsm.cuh"
#include "nfsp_fsm.cuh"
#include "nfsp_rules.cuh"
#include "philox.cuh"

namespace nfsp {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------
// ENV_NFSP
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void nfsp_redeal(NfspW &g, uint32_t dealer, const Philox4 &x, uint32_t eta_u32) {
    g.reset(dealer, deal_ranks(__umulhi(x.y, 120u)), x.z < eta_u32, x.w < eta_u32);
}

__global__ void __launch_bounds__(kThreads)
nfsp_reset_kernel(uint64_t *__restrict__ state, int64_t n, uint64_t seed, uint64_t game0, uint64_t step,
                  const int8_t *__restrict__ dealer, uint32_t eta_u32) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t game = game0 + (uint64_t)i;
        NfspW g;
        nfsp_redeal(g, dealer ? (uint32_t)(dealer[i] & 1) : (uint32_t)(game & 1u), game_block(seed, game, step, STREAM_STEP),
                    eta_u32);
        state[i] = g.w;
    }
}

__global__ void __launch_bounds__(kThreads)
nfsp_set_hands_kernel(uint64_t *__restrict__ state, int64_t n, const int8_t *__restrict__ dealer,
                      const int8_t *__restrict__ cards, const int8_t *__restrict__ policy) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        NfspW g;
        const uint32_t c = (uint32_t)(cards[3 * i] & 3) | ((uint32_t)(cards[3 * i + 1] & 3) << 2) |
                           ((uint32_t)(cards[3 * i + 2] & 3) << 4);
        g.reset((uint32_t)(dealer[i] & 1), c, policy ? (uint32_t)(policy[2 * i] & 1) : 0u,
                policy ? (uint32_t)(policy[2 * i + 1] & 1) : 0u);
        state[i] = g.w;
    }
}

template <bool kTrace>
__global__ void __launch_bounds__(kThreads)
nfsp_step_kernel(uint64_t *__restrict__ state, int64_t n, uint64_t seed, uint64_t game0, uint64_t step0, int n_steps,
                 const int8_t *__restrict__ actions, const int8_t *__restrict__ players, int auto_reset,
                 uint32_t eta_u32, uint32_t *__restrict__ trace) {
    const int64_t plane = (int64_t)n_steps * n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t game = game0 + (uint64_t)i;
        NfspW g{state[i]};
        for (int t = 0; t < n_steps; ++t) {
            const uint64_t step = step0 + (uint64_t)t;
            const int64_t at = (int64_t)t * n + i;
            int code = actions ? (int)actions[at] : -1;
            bool started = false;
            const bool redeal = auto_reset && (g.need_reset() || g.terminated());  // also hands finished with auto_reset = 0
            if (redeal || code < 0) {  // one Philox block serves the re-deal and the random action
                const Philox4 x = game_block(seed, game, step, STREAM_STEP);
                if (redeal) {
                    nfsp_redeal(g, g.dealer() ^ 1u, x, eta_u32);
                    started = true;
                }
                if (code < 0) code = (int)__umulhi(x.x, 3u);
            }
            const int p = players ? (int)(players[at] & 1) : g.to_act();
            if (code == 4) continue;  // this game sits the step out (per-game call sequences)
            const int raw = code == 3 ? 0 : code;
            const int eff = g.step(raw, code != 3, p);
            if (auto_reset && g.terminated()) g.w |= 1ull << 43;
            if (kTrace) {
                trace[at] = g.obs(p) | ((uint32_t)g.terminated() << 30) | ((uint32_t)p << 31);
                trace[plane + at] = __float_as_uint(0.5f * (float)g.reward_half(p));
                trace[2 * plane + at] = g.trace_misc(raw, eff, started);
            }
        }
        state[i] = g.w;
    }
}

// The same step for the configuration the throughput figure is quoted on (uniform Philox actions, main.train's turn
// order, auto re-deal) as a table-driven state machine (nfsp_fsm.cuh): HBM sees the packed word once per launch and the
// 12-byte trace record per transition.  Bit-identical to nfsp_step_kernel; 96 instructions per transition (43 on the
// integer pipe) against 201 (126) for the register formulation of nfsp_fast.cuh it replaced -- 0.81 instead of 0.46 of
// the HBM peak.  `image` is the table image built on the host at nfsp_env_create.
template <bool kTrace>
__global__ void __launch_bounds__(kThreads)
nfsp_step_fsm_kernel(uint64_t *__restrict__ state, int64_t n, const uint32_t *__restrict__ image, const PhiloxKeys keys,
                     uint64_t game0, uint64_t step0, int n_steps, uint32_t eta_u32, uint32_t *__restrict__ trace) {
    __shared__ __align__(16) uint32_t s_tab[fsm::kImageWords];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(s_tab);
    for (int w = threadIdx.x; w < fsm::kImageWords; w += blockDim.x) s_tab[w] = image[w] + (fsm::is_address(w) ? base : 0u);
    __syncthreads();
    const uint32_t dl_sum = 2u * (base + (uint32_t)fsm::kDealOff) + (uint32_t)fsm::kDealHalf;
    const uint32_t rew_base = base + (uint32_t)fsm::kRewardOff;
    const int64_t plane = (int64_t)n_steps * n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t game = game0 + (uint64_t)i;
        fsm::Game g;
        g.unpack(state[i], base);
        uint32_t *p0 = trace + i, *p1 = p0 + plane, *p2 = p1 + plane;
        for (int t = 0; t < n_steps; ++t) {
            const Philox4 x = game_block(keys, game, step0 + (uint64_t)t, STREAM_STEP);
            uint32_t started = 0u;
            if (g.HX & fsm::kHxOver) {  // newenv.py:76-114 + main.py:28-45: the other player deals
                g.dl = dl_sum - g.dl;
                const uint4 D = fsm::lds128(g.dl + __umulhi(x.y, 120u) * 16u);
                uint32_t pa = g.dl;
                if (x.z < eta_u32) pa += 16u;
                if (x.w < eta_u32) pa += 32u;
                const uint4 P = fsm::lds128(pa + (uint32_t)fsm::kPolOff);
                g.PA = D.x | P.x;
                g.PO = D.y | P.y;
                g.ms = D.z | P.z;
                g.tix = P.w;
                g.HX = fsm::kCm0;
                started = 1u << 21;
            }
            const uint32_t ea = g.tix + __umulhi(x.x, 3u) * (uint32_t)fsm::kEntryBytes;
            const uint4 E = fsm::lds128(ea);  // hc, pa, pm, nx
            g.PA = (g.PA & ~fsm::kPClear) + E.y;
            g.HX = (g.HX & 0xFFFFFFu) | E.x;
            if (kTrace) {
                const uint4 T = fsm::lds128(ea + 16u);  // mw, mb, ma, mo
                g.ms += T.y;
                *p0 = g.HX & (g.PA | 0xC0FFFFFFu);  // observation of the actor | hand over << 30 | actor << 31
                *p1 = fsm::lds32(rew_base + ((g.PA & T.z) | (g.PO & T.w)));
                *p2 = g.ms | T.x | started;
                p0 += n;
                p1 += n;
                p2 += n;
            }
            const uint32_t a = g.PA;
            g.PA = (g.PO & E.z) | (a & ~E.z);
            g.PO = (a & E.z) | (g.PO & ~E.z);
            g.tix = E.w;
        }
        state[i] = g.pack(base);
    }
}

__global__ void __launch_bounds__(kThreads)
nfsp_observe_kernel(const uint64_t *__restrict__ state, int64_t n, const int8_t *__restrict__ players, int player,
                    uint32_t *__restrict__ s, uint32_t *__restrict__ s2, float *__restrict__ reward,
                    uint8_t *__restrict__ term, uint8_t *__restrict__ last_a) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const NfspW g{state[i]};
        const int p = player >= 0 ? player : (int)(players[i] & 1);
        if (s) s[i] = g.snapshot(p);
        if (s2) s2[i] = g.obs(p);
        if (reward) reward[i] = 0.5f * (float)g.reward_half(p);
        if (term) term[i] = g.terminated();
        if (last_a) last_a[i] = g.acted_nz(p) ? (uint8_t)g.last_a(p) : (uint8_t)3;
    }
}

// newenv.Env.do_action (newenv.py:131-178) called on its own: code 0..2 = np.argmax(action), 3 = the all-zero vector
__global__ void __launch_bounds__(kThreads)
nfsp_do_action_kernel(uint64_t *__restrict__ state, int64_t n, const int8_t *__restrict__ actions, const int8_t *__restrict__ players,
                      int8_t *__restrict__ fold) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (actions[i] < 0) {  // negative code: this game makes no call
            if (fold) fold[i] = 0;
            continue;
        }
        NfspW g{state[i]};
        const int code = actions[i] & 3;
        const bool f = g.do_action(code == 3 ? 0 : code, code != 3, players[i] & 1);
        state[i] = g.w;
        if (fold) fold[i] = (int8_t)f;
    }
}

__global__ void __launch_bounds__(kThreads)
nfsp_export_kernel(const uint64_t *__restrict__ state, int64_t n, int32_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const NfspW g{state[i]};
        int32_t *o = out + i * NFSP_EXPORT_FIELDS;
        o[0] = (int32_t)g.hist();   o[1] = (int32_t)g.card(0);  o[2] = (int32_t)g.card(1);  o[3] = (int32_t)g.pub();
        o[4] = (int32_t)g.dealer(); o[5] = (int32_t)g.round();  o[6] = (int32_t)g.k();
        o[7] = (int32_t)g.bets(0);  o[8] = (int32_t)g.bets(1);  o[9] = g.terminated();      o[10] = g.need_reset();
        o[11] = (int32_t)g.policy(0); o[12] = (int32_t)g.policy(1);
        o[13] = (int32_t)g.last_a(0); o[14] = (int32_t)g.last_a(1);
        o[15] = g.acted_nz(0);      o[16] = g.acted_nz(1);
        o[17] = (int32_t)g.obs(0);  o[18] = (int32_t)g.obs(1);
        o[19] = (int32_t)g.snapshot(0); o[20] = (int32_t)g.snapshot(1);
        o[21] = g.reward_half(0);   o[22] = g.reward_half(1);   o[23] = g.anomaly();
    }
}

// ------------------------------------------------------------------------------------------------
// ENV_LEGACY
// ------------------------------------------------------------------------------------------------
constexpr int kPenalty = -1;  // config.ini:12

__device__ __forceinline__ void legacy_redeal(LegacyW &g, const Philox4 &x) {
    const uint32_t c = deal_ranks(__umulhi(x.z, 120u));
    g.reset(c & 3u, (c >> 2) & 3u);
}

__global__ void __launch_bounds__(kThreads)
legacy_reset_kernel(uint64_t *__restrict__ state, int64_t n, uint64_t seed, uint64_t game0, uint64_t step,
                    const int8_t *__restrict__ cards) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        LegacyW g;
        if (cards) g.reset((uint32_t)(cards[2 * i] & 3), (uint32_t)(cards[2 * i + 1] & 3));
        else legacy_redeal(g, game_block(seed, game0 + (uint64_t)i, step, STREAM_STEP));
        state[i] = g.w;
    }
}

__global__ void __launch_bounds__(kThreads)
legacy_step_kernel(uint64_t *__restrict__ state, int64_t n, const int8_t *__restrict__ actions,
                   const int8_t *__restrict__ players, int player) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (actions[i] == 4 || (player < 0 && players[i] > 1)) continue;  // game sits this call out
        LegacyW g{state[i]};
        g.step((int)actions[i], player >= 0 ? player : (int)(players[i] & 1), kPenalty);
        state[i] = g.w;
    }
}

__global__ void __launch_bounds__(kThreads)
legacy_gns_kernel(uint64_t *__restrict__ state, int64_t n, const int8_t *__restrict__ players, int player,
                  int32_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (player < 0 && players[i] > 1) continue;  // game sits this call out
        LegacyW g{state[i]};
        const int p = player >= 0 ? player : (int)(players[i] & 1);
        g.get_new_state(p);
        state[i] = g.w;
        if (out) {
            int32_t *o = out + 5 * i;
            o[0] = g.card(p); o[1] = -1; o[2] = g.st_pot(p); o[3] = g.reward(p); o[4] = g.terminal(p);
        }
    }
}

// One game through n_iters README iterations on the packed word: the rules as written in legacy_rules.cuh.  The rollout
// kernel plays a game this way only when its word is not a state a rollout can produce (see legacy_fsm.cuh).
template <bool kTrace>
__device__ __noinline__ uint64_t legacy_generic_game(uint64_t w, const PhiloxKeys &keys, uint64_t game, uint64_t step0, int n_iters,
                                                      const int8_t *__restrict__ actions, uint32_t *__restrict__ rec, int64_t n,
                                                      int64_t i) {
    const int64_t plane = (int64_t)n_iters * n * 2;
    LegacyW g{w};
    for (int t = 0; t < n_iters; ++t) {
        bool started = false;
        const int64_t at = ((int64_t)t * n + i) * 2;
        int a0 = actions ? (int)actions[at] : -1, a1 = actions ? (int)actions[at + 1] : -1;
        if (g.need_reset() || a0 < 0 || a1 < 0) {
            const Philox4 x = game_block(keys, game, step0 + (uint64_t)t, STREAM_STEP);
            if (g.need_reset()) {
                legacy_redeal(g, x);
                started = true;
            }
            if (a0 < 0) a0 = (int)__umulhi(x.x, 3u);
            if (a1 < 0) a1 = (int)__umulhi(x.y, 3u);
        }
        // README.md:15-38: both players step, then both read the new state
        g.step(a0, 0, kPenalty);
        g.step(a1, 1, kPenalty);
        g.get_new_state(0);
        const int t0 = g.terminal(0), r0 = g.reward(0), sp0 = g.st_pot(0);
        g.get_new_state(1);
        if (kTrace) {
            uint2 w0, w1, w2;
            w0.x = (uint32_t)g.card(0) | (0xFFu << 8) | ((uint32_t)sp0 << 16) | ((uint32_t)t0 << 24);
            w0.y = (uint32_t)g.card(1) | (0xFFu << 8) | ((uint32_t)g.st_pot(1) << 16) | ((uint32_t)g.terminal(1) << 24);
            *reinterpret_cast<uint2 *>(rec + at) = w0;
            w1.x = (uint32_t)r0;
            w1.y = (uint32_t)g.reward(1);
            *reinterpret_cast<uint2 *>(rec + plane + at) = w1;
            w2.x = (uint32_t)a0 | ((uint32_t)(g.left(0) + 1) << 2) | ((uint32_t)g.pot(0) << 5) | ((uint32_t)started << 8);
            w2.y = (uint32_t)a1 | ((uint32_t)(g.left(1) + 1) << 2) | ((uint32_t)g.pot(1) << 5) | ((uint32_t)started << 8);
            *reinterpret_cast<uint2 *>(rec + 2 * plane + at) = w2;
        }
        if (t0 | g.terminal(1)) g.w |= 1ull << 30;
    }
    return g.w;
}

// The README iteration for every game, n_iters times, with auto re-deal: one table entry per (state, a0, a1)
// (legacy_fsm.cuh).  `image` is the table image built on the host at nfsp_env_create.
template <bool kTrace>
__global__ void __launch_bounds__(kThreads, 6)
legacy_rollout_kernel(uint64_t *__restrict__ state, int64_t n, const uint32_t *__restrict__ image, const PhiloxKeys keys,
                      uint64_t game0, uint64_t step0, int n_iters, const int8_t *__restrict__ actions,
                      uint32_t *__restrict__ rec) {
    // rows and deals in shared memory; the state-key table is read once per game and launch and stays in global memory
    __shared__ __align__(16) uint32_t s_tab[lfsm::kSharedWords];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(s_tab);
    for (int w = threadIdx.x; w < lfsm::kSharedWords; w += blockDim.x) {
        const int src = w < lfsm::kKeyOff / 4 ? w : w + lfsm::kKeys / 4;
        s_tab[w] = image[src] + (lfsm::is_address(src) ? base : 0u);
    }
    __syncthreads();
    const uint8_t *g_keys = reinterpret_cast<const uint8_t *>(image) + lfsm::kKeyOff;
    const uint32_t deal_base = base + (uint32_t)lfsm::kKeyOff;
    const int64_t plane = (int64_t)n_iters * n;  // in uint2 records
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t game = game0 + (uint64_t)i;
        const LegacyW g{state[i]};
        // a hand that is over is re-dealt before anything is looked up: any row will do
        const int key = g.need_reset() ? -1 : lfsm::state_key(g, kPenalty);
        const uint32_t row = g.need_reset() ? 0u : (key >= 0 ? (uint32_t)__ldg(g_keys + key) : 0xFFu);
        if (row == 0xFFu) {
            state[i] = legacy_generic_game<kTrace>(g.w, keys, game, step0, n_iters, actions, rec, n, i);
            continue;
        }
        uint32_t six = base + (uint32_t)lfsm::kRowsOff + row * (uint32_t)lfsm::kRowBytes, ea = six;
        uint32_t c0 = (uint32_t)g.card(0), c1 = (uint32_t)g.card(1);
        int G = c0 > c1, L = c0 < c1, r0 = 0, r1 = 0;
        uint32_t need = g.need_reset() ? lfsm::kTermBit : 0u;
        uint2 *p0 = reinterpret_cast<uint2 *>(rec) + i, *p1 = p0 + plane, *p2 = p1 + plane;
        for (int t = 0; t < n_iters; ++t) {
            int a0 = -1, a1 = -1;
            if (actions) {  // values above 2 act as a raise (env.py:130: the else branch)
                const int64_t at = ((int64_t)t * n + i) * 2;
                a0 = min((int)actions[at], 2);
                a1 = min((int)actions[at + 1], 2);
            }
            uint32_t started = 0u;
            if (need | (uint32_t)((a0 | a1) < 0)) {
                const Philox4 x = game_block(keys, game, step0 + (uint64_t)t, STREAM_STEP);
                if (need) {  // env.py:46-72
                    const uint4 D = fsm::lds128(deal_base + __umulhi(x.z, 120u) * 16u);
                    c0 = D.x; c1 = D.y; G = (int)D.z; L = (int)D.w;
                    six = base + (uint32_t)lfsm::kRowsOff;
                    started = 1u << 8;
                }
                if (a0 < 0) a0 = (int)__umulhi(x.x, 3u);
                if (a1 < 0) a1 = (int)__umulhi(x.y, 3u);
            }
            ea = six + (uint32_t)(a0 * 3 + a1) * (uint32_t)lfsm::kEntryBytes;
            const uint4 A = fsm::lds128(ea);  // nx, w0, rb0, rb1
            const uint4 B = fsm::lds128(ea + 16u);  // w2x, w2y, p0t, p1t
            const int u = L * (int)B.w - G * (int)B.z;  // env.py:184-199 for both players
            r0 = (int)A.z + u;
            r1 = (int)A.w - u;
            if (kTrace) {
                *p0 = make_uint2(A.y | c0, A.y | c1);
                *p1 = make_uint2((uint32_t)r0, (uint32_t)r1);
                *p2 = make_uint2(B.x | started, B.y | started);
                p0 += n;
                p1 += n;
                p2 += n;
            }
            need = A.y & lfsm::kTermBit;
            six = A.x;
        }
        const uint32_t lo = fsm::lds32(ea + 32u) | c0 | (c1 << 2);
        const uint32_t hi = ((uint32_t)r0 & 0xFFFFu) | ((uint32_t)r1 << 16);
        state[i] = (uint64_t)lo | ((uint64_t)hi << 32);
    }
}

__global__ void __launch_bounds__(kThreads)
legacy_export_kernel(const uint64_t *__restrict__ state, int64_t n, int32_t *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const LegacyW g{state[i]};
        int32_t *o = out + i * NFSP_LEGACY_EXPORT_FIELDS;
        for (int p = 0; p < 2; ++p) {
            o[0 + p] = g.card(p);   o[2 + p] = g.left(p);    o[4 + p] = g.pot(p);  o[6 + p] = g.terminal(p);
            o[8 + p] = g.st_pot(p); o[10 + p] = g.st_action(p); o[12 + p] = g.reward(p);
        }
    }
}

__global__ void __launch_bounds__(kThreads)
expand_obs_kernel(const uint32_t *__restrict__ masks, int64_t n, float *__restrict__ out) {
    const int64_t total = n * NFSP_OBS_DIM;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / NFSP_OBS_DIM;
        const int bit = (int)(e - row * NFSP_OBS_DIM);
        out[e] = (float)((masks[row] >> bit) & 1u);
    }
}

}  // namespace nfsp

using namespace nfsp;

// grid of a table-driven kernel: exactly the CTAs that are resident at once (each copies the table image into its
// shared memory once and then strides over the games), never more than the games need
template <class Kernel>
static int resident_grid(Kernel kernel, int64_t n, int sm_count) {
    static const int per_sm = [kernel] {
        int v = 0;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kThreads, 0) == cudaSuccess && v >= 1 ? v : 4;
    }();
    return grid_for(n, kThreads, sm_count, per_sm);
}

#define ENV_PROLOGUE(h, want_rules)                                                              \
    NFSP_CHECK_ARG(h != nullptr, "null handle");                                                 \
    NFSP_CHECK_ARG(h->rules == (want_rules), "handle has rules %d, call needs %d", h->rules, (want_rules)); \
    DeviceGuard guard__(h->device);                                                              \
    if (!guard__.ok) return set_error(NFSP_E_CUDA, "cannot select device %d", h->device);       \
    cudaStream_t st = (cudaStream_t)stream;                                                      \
    const int grid = grid_for(h->n, kThreads, h->sm_count, 8);                                   \
    (void)grid

int nfsp_fsm_upload(nfsp_env_t h) {
    static uint32_t image[fsm::kImageWords], legacy_image[lfsm::kImageWords];
    static const int legacy_rows = (fsm::build_image(image), lfsm::build_image(legacy_image, kPenalty));
    if (legacy_rows < 0) return set_error(NFSP_E_ARG, "legacy state machine does not fit its image");
    const bool legacy = h->rules == NFSP_RULES_LEGACY;
    const size_t bytes = legacy ? sizeof(legacy_image) : sizeof(image);
    NFSP_CUDA(cudaMalloc(&h->d_fsm, bytes));
    NFSP_CUDA(cudaMemcpy(h->d_fsm, legacy ? legacy_image : image, bytes, cudaMemcpyHostToDevice));
    return NFSP_OK;
}

// the table images of the state-machine kernels, for inspection on the host (no device needed)
extern "C" int nfsp_fsm_image(uint32_t *out, int capacity_words) {
    if (out && capacity_words >= fsm::kImageWords) fsm::build_image(out);
    return fsm::kImageWords;
}
extern "C" int nfsp_legacy_fsm_image(uint32_t *out, int capacity_words, int *n_rows) {
    if (out && capacity_words >= lfsm::kImageWords) {
        const int rows = lfsm::build_image(out, kPenalty);
        if (n_rows) *n_rows = rows;
    }
    return lfsm::kImageWords;
}

extern "C" int nfsp_env_reset(nfsp_env_t h, const int8_t *d_dealer, double eta, void *stream) {
    ENV_PROLOGUE(h, NFSP_RULES_NFSP);
    nfsp_reset_kernel<<<grid, kThreads, 0, st>>>(h->d_state, h->n, h->seed, h->game0, h->step, d_dealer,
                                                 frac_u32(eta));
    NFSP_LAUNCH_CHECK();
    h->step += 1;
    return NFSP_OK;
}

extern "C" int nfsp_env_set_hands(nfsp_env_t h, const int8_t *d_dealer, const int8_t *d_cards,
                                  const int8_t *d_policy, void *stream) {
    ENV_PROLOGUE(h, NFSP_RULES_NFSP);
    NFSP_CHECK_ARG(d_dealer && d_cards, "set_hands needs dealer and cards");
    nfsp_set_hands_kernel<<<grid, kThreads, 0, st>>>(h->d_state, h->n, d_dealer, d_cards, d_policy);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_env_step(nfsp_env_t h, const int8_t *d_actions, const int8_t *d_players, int n_steps,
                             int auto_reset, double eta, uint32_t *d_trace, void *stream) {
    ENV_PROLOGUE(h, NFSP_RULES_NFSP);
    NFSP_CHECK_ARG(n_steps >= 1, "n_steps must be >= 1");
    if (!d_actions && !d_players && auto_reset) {  // the benchmark configuration has its own lean kernel
        if (d_trace)
            nfsp_step_fsm_kernel<true><<<resident_grid(nfsp_step_fsm_kernel<true>, h->n, h->sm_count), kThreads, 0, st>>>(
                h->d_state, h->n, h->d_fsm, philox_keys(h->seed), h->game0, h->step, n_steps, frac_u32(eta), d_trace);
        else
            nfsp_step_fsm_kernel<false><<<resident_grid(nfsp_step_fsm_kernel<false>, h->n, h->sm_count), kThreads, 0, st>>>(
                h->d_state, h->n, h->d_fsm, philox_keys(h->seed), h->game0, h->step, n_steps, frac_u32(eta), nullptr);
        NFSP_LAUNCH_CHECK();
        h->step += (uint64_t)n_steps;
        return NFSP_OK;
    }
    if (d_trace)
        nfsp_step_kernel<true><<<grid, kThreads, 0, st>>>(h->d_state, h->n, h->seed, h->game0, h->step, n_steps,
                                                          d_actions, d_players, auto_reset, frac_u32(eta), d_trace);
    else
        nfsp_step_kernel<false><<<grid, kThreads, 0, st>>>(h->d_state, h->n, h->seed, h->game0, h->step, n_steps,
                                                           d_actions, d_players, auto_reset, frac_u32(eta), nullptr);
    NFSP_LAUNCH_CHECK();
    h->step += (uint64_t)n_steps;
    return NFSP_OK;
}

extern "C" int nfsp_env_do_action(nfsp_env_t h, const int8_t *d_actions, const int8_t *d_players, int8_t *d_fold, void *stream) {
    NFSP_CHECK_ARG(h != nullptr && d_actions != nullptr && d_players != nullptr, "null argument");
    NFSP_CHECK_ARG(h->rules == NFSP_RULES_NFSP, "do_action is a method of leduc.newenv.Env (NFSP rules)");
    DeviceGuard guard(h->device);
    if (!guard.ok) return set_error(NFSP_E_CUDA, "cannot select device %d", h->device);
    nfsp_do_action_kernel<<<grid_for(h->n, kThreads, h->sm_count, 8), kThreads, 0, (cudaStream_t)stream>>>(h->d_state, h->n, d_actions,
                                                                                                        d_players, d_fold);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_env_observe(nfsp_env_t h, const int8_t *d_players, int player, uint32_t *d_s, uint32_t *d_s2,
                                float *d_reward, uint8_t *d_term, uint8_t *d_last_a, void *stream) {
    ENV_PROLOGUE(h, NFSP_RULES_NFSP);
    NFSP_CHECK_ARG(player == 0 || player == 1 || (player < 0 && d_players), "player must be 0/1 or per-game");
    nfsp_observe_kernel<<<grid, kThreads, 0, st>>>(h->d_state, h->n, d_players, player, d_s, d_s2, d_reward, d_term,
                                                   d_last_a);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_env_export(nfsp_env_t h, int32_t *d_fields, void *stream) {
    ENV_PROLOGUE(h, NFSP_RULES_NFSP);
    NFSP_CHECK_ARG(d_fields, "null output");
    nfsp_export_kernel<<<grid, kThreads, 0, st>>>(h->d_state, h->n, d_fields);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_legacy_reset(nfsp_env_t h, void *stream) {
    ENV_PROLOGUE(h, NFSP_RULES_LEGACY);
    legacy_reset_kernel<<<grid, kThreads, 0, st>>>(h->d_state, h->n, h->seed, h->game0, h->step, nullptr);
    NFSP_LAUNCH_CHECK();
    h->step += 1;
    return NFSP_OK;
}

extern "C" int nfsp_legacy_set_hands(nfsp_env_t h, const int8_t *d_cards, void *stream) {
    ENV_PROLOGUE(h, NFSP_RULES_LEGACY);
    NFSP_CHECK_ARG(d_cards, "null cards");
    legacy_reset_kernel<<<grid, kThreads, 0, st>>>(h->d_state, h->n, h->seed, h->game0, h->step, d_cards);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_legacy_step(nfsp_env_t h, const int8_t *d_actions, const int8_t *d_players, int player,
                                void *stream) {
    ENV_PROLOGUE(h, NFSP_RULES_LEGACY);
    NFSP_CHECK_ARG(d_actions, "null actions");
    NFSP_CHECK_ARG(player == 0 || player == 1 || (player < 0 && d_players), "player must be 0/1 or per-game");
    legacy_step_kernel<<<grid, kThreads, 0, st>>>(h->d_state, h->n, d_actions, d_players, player);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_legacy_get_new_state(nfsp_env_t h, const int8_t *d_players, int player, int32_t *d_out,
                                         void *stream) {
    ENV_PROLOGUE(h, NFSP_RULES_LEGACY);
    NFSP_CHECK_ARG(player == 0 || player == 1 || (player < 0 && d_players), "player must be 0/1 or per-game");
    legacy_gns_kernel<<<grid, kThreads, 0, st>>>(h->d_state, h->n, d_players, player, d_out);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_legacy_rollout(nfsp_env_t h, const int8_t *d_actions, int n_iters, uint32_t *d_rec,
                                   void *stream) {
    ENV_PROLOGUE(h, NFSP_RULES_LEGACY);
    NFSP_CHECK_ARG(n_iters >= 1, "n_iters must be >= 1");
    if (d_rec)
        legacy_rollout_kernel<true><<<resident_grid(legacy_rollout_kernel<true>, h->n, h->sm_count), kThreads, 0, st>>>(
            h->d_state, h->n, h->d_fsm, philox_keys(h->seed), h->game0, h->step, n_iters, d_actions, d_rec);
    else
        legacy_rollout_kernel<false><<<resident_grid(legacy_rollout_kernel<false>, h->n, h->sm_count), kThreads, 0, st>>>(
            h->d_state, h->n, h->d_fsm, philox_keys(h->seed), h->game0, h->step, n_iters, d_actions, nullptr);
    NFSP_LAUNCH_CHECK();
    h->step += (uint64_t)n_iters;
    return NFSP_OK;
}

extern "C" int nfsp_legacy_export(nfsp_env_t h, int32_t *d_fields, void *stream) {
    ENV_PROLOGUE(h, NFSP_RULES_LEGACY);
    NFSP_CHECK_ARG(d_fields, "null output");
    legacy_export_kernel<<<grid, kThreads, 0, st>>>(h->d_state, h->n, d_fields);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}

extern "C" int nfsp_expand_obs(const uint32_t *d_masks, int64_t n, float *d_out, void *stream) {
    NFSP_CHECK_ARG(d_masks && d_out && n >= 0, "bad arguments");
    if (n == 0) return NFSP_OK;
    const int64_t total = n * NFSP_OBS_DIM;
    const int grid = (int)((total + kThreads - 1) / kThreads < 148 * 16 ? (total + kThreads - 1) / kThreads : 148 * 16);
    expand_obs_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(d_masks, n, d_out);
    NFSP_LAUNCH_CHECK();
    return NFSP_OK;
}
