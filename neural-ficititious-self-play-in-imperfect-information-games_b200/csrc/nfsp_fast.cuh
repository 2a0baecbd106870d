// ENV_NFSP in registers, actor-relative: the representation the hot kernels (fused rollout, env-only fast
// path) keep across the steps of a launch.  The packed 64-bit word of nfsp_rules.cuh stays the HBM format
// (one load + one store per game per launch); here the same fields are spread over 32-bit registers laid
// out so that one decision costs a few dozen integer instructions:
//
//   * per-player fields live in two words PA (the player to act) / PO (the other one) that SWAP when the turn
//     passes, so no field is ever addressed with a variable shift;
//   * the betting state is one number sigma = 9*round + sequence id of the current round (0 -, 1 C, 2 R, 3 CC,
//     4 CR, 5 RC, 6 RR, 7 CRC, 8 RRC), and newenv.Env.do_action / game_or_round_has_terminated
//     (newenv.py:131-190: raise->call coercions, chips added, history bit, round over) are ONE lookup in a
//     54-entry shared-memory table indexed by (sigma, raw action) -- the integer pipe is what bounds these
//     kernels (ncu r01: alu pipe 75% busy), the load/store pipe is idle;
//   * the snapshots s[p] (newenv.py:200-202) are explicit registers (the observation the player acted on).
//
// Semantics are those of NfspW::step (nfsp_rules.cuh, the packed-word implementation the general env kernel
// runs); the parity tests run the golden hands and seeded rollouts through both.
#pragma once
#include "nfsp_rules.cuh"
#include "philox.cuh"

namespace nfsp {

// per-player word P
//   0-3  bets in half chips      4 policy ('b' = 1)      5-6 last raw action     7 acted with a non-zero vector
//   8-10 time of the last step() call (7 = never)        11-12 card rank       13-14 public card rank (copy)
//   15-16 outcome code of a showdown in which THIS player is the actor (newenv.py:261-298): 1 it wins, 2 the other player
//         wins, 3 nobody -- precomputed with the deal, so that the step reads it instead of comparing winner and actor
//   24-26 one-hot private card (obs bits 24-26)          27-29 private | public one-hot (obs bits 27-29, shown in round 1)
constexpr uint32_t kPBets = 0xFu, kPPol = 1u << 4, kPNz = 1u << 7;
constexpr uint32_t kRound0Cards = 0x07000000u, kRound1Cards = 0x3F000000u;

// flag word F
//   0-2 tt = 3*round + k (k up to 3 after a showdown)    3 dealer    4 player to act    5-6 public card rank
//   7 need_reset    8-12 sigma    13-14 f = finished round-0 sequence (0 CC, 1 RC, 2 CRC, 3 RRC), set when round 1 starts
//   16 terminated   17 actor of the terminating step     18-19 outcome   20 anomaly
constexpr uint32_t kFNeedReset = 1u << 7, kFTerm = 1u << 16, kFStepMask = 0x1F07u;

__device__ __forceinline__ uint32_t seq_id6(uint32_t rnd) {  // 6 slot bits of one round -> sequence id 0..8
    const uint32_t s0 = rnd & 3u, s1 = (rnd >> 2) & 3u, s2 = (rnd >> 4) & 3u;
    return s2 ? 6u + s0 : (s1 ? 2u * s0 + s1 : s0);
}

// ---- step table -------------------------------------------------------------------------------------
// entry(sigma, raw):  0-1 effective action   2-4 half chips added   5 history bit valid (not a fold)
//   6-9 history bit position 2*t + action - 1 within the actor's 12-bit block   10-11 kind: 0 the turn passes,
//   1 round 0 over (public card, round 1), 2 showdown, 3 fold   12-14 t of this step
//   16-31 the new tt / sigma / f fields in F's layout
constexpr int kStepLutWords = 54;
enum : uint32_t { KIND_PASS = 0, KIND_ROUND = 1, KIND_SHOWDOWN = 2, KIND_FOLD = 3 };

__device__ __forceinline__ uint32_t step_entry(uint32_t sigma, uint32_t raw) {
    const uint32_t r = sigma >= 9u ? 1u : 0u, s = sigma - 9u * r;
    const uint32_t k = s == 0u ? 0u : (s < 3u ? 1u : 2u), t = 3u * r + k;
    uint32_t av = raw;
    if (av == A_RAISE && s >= 3u) av = A_CALL;  // newenv.py:141-145: [C,R] and "already raised" both mean s in {4,6}
    if (av == A_FOLD) return av | (KIND_FOLD << 10) | (t << 12) | ((t | (sigma << 8)) << 16);
    const bool prev_raise = s != 0u && !(s & 1u);                                           // s in {2,4,6}
    const uint32_t add = (av == A_RAISE ? 2u : 0u) + (prev_raise ? 2u : 0u) + (t == 0u ? 1u : 0u);  // newenv.py:157-176
    const uint32_t sn = s < 3u ? 2u * s + av : 5u + (s >> 1);
    const bool over = s >= 3u || (s != 0u && av == A_CALL);                                 // newenv.py:180-190
    uint32_t kind, nf;
    if (!over) {
        kind = KIND_PASS;
        nf = (t + 1u) | ((9u * r + sn) << 8);
    } else if (r == 0u) {
        kind = KIND_ROUND;  // newenv.py:215-242
        nf = 3u | (9u << 8) | (((sn >> 1) - 1u) << 13);
    } else {
        kind = KIND_SHOWDOWN;  // newenv.py:261-298; k counts the last action too
        nf = (t + 1u) | ((9u + sn) << 8);
    }
    return av | (add << 2) | (1u << 5) | ((2u * t + av - 1u) << 6) | (kind << 10) | (t << 12) | (nf << 16);
}
__device__ __forceinline__ void fill_step_lut(uint32_t *lut) {
    for (uint32_t i = threadIdx.x; i < (uint32_t)kStepLutWords; i += blockDim.x) lut[i] = step_entry(i / 3u, i % 3u);
}

// ---- deal table -------------------------------------------------------------------------------------
// deal index 0..119 (deck.py:35-50: uniform ordered draw of 3 of 6 cards) -> the card part of both players'
// P words, lut[idx] for player 0 and lut[120 + idx] for player 1:
//   7 << 8 (never acted) | card << 11 | public card << 13 | one-hot rows (bits 24-29)
constexpr int kDealLutWords = 360;  // + [240 + idx]: the cards as the trace record's word 3 holds them (c0<<6 | c1<<8 | pub<<10)
__device__ __forceinline__ uint32_t showdown_winner(uint32_t c0, uint32_t c1, uint32_t pub) {
    // newenv.py:261-298: the actor pairing the board wins, else the opponent pairing it, else the lower rank index
    // (0 = Ace).  3 = both pair the board: the actor wins whoever it is (cannot be dealt from two cards per rank,
    // but set_hands accepts it and the reference's order of tests decides)
    if (c0 == pub && c1 == pub) return 3u;
    return c0 == pub ? 0u : (c1 == pub ? 1u : (c0 < c1 ? 0u : (c0 > c1 ? 1u : 2u)));
}
// outcome code of a showdown for player p as the actor, from showdown_winner's value
__device__ __forceinline__ uint32_t showdown_code(uint32_t winner, uint32_t p) {
    return winner == 2u ? 3u : ((winner == p || winner == 3u) ? 1u : 2u);
}
__device__ __forceinline__ uint32_t deal_word(uint32_t c, uint32_t pub, uint32_t code) {
    const uint32_t row0 = 1u << c;
    return (7u << 8) | (c << 11) | (pub << 13) | (code << 15) | (row0 << 24) | ((row0 | (1u << pub)) << 27);
}
__device__ __forceinline__ void fill_deal_lut(uint32_t *lut) {
    for (uint32_t i = threadIdx.x; i < 120u; i += blockDim.x) {
        const uint32_t c = deal_ranks(i);
        const uint32_t w = showdown_winner(c & 3u, (c >> 2) & 3u, (c >> 4) & 3u);
        lut[i] = deal_word(c & 3u, (c >> 4) & 3u, showdown_code(w, 0u));
        lut[120u + i] = deal_word((c >> 2) & 3u, (c >> 4) & 3u, showdown_code(w, 1u));
        lut[240u + i] = ((c & 3u) << 6) | (((c >> 2) & 3u) << 8) | (((c >> 4) & 3u) << 10);
    }
}

// both tables, one shared array per CTA
struct FastLuts {
    uint32_t step[kStepLutWords];
    uint32_t deal[kDealLutWords];
    __device__ __forceinline__ void fill() {
        fill_step_lut(step);
        fill_deal_lut(deal);
    }
};

struct NfspFast {
    uint32_t H, F, PA, PO, SA, SO, cmask;

    __device__ __forceinline__ uint32_t tt() const { return F & 7u; }
    __device__ __forceinline__ uint32_t dealer() const { return (F >> 3) & 1u; }
    __device__ __forceinline__ uint32_t p() const { return (F >> 4) & 1u; }
    __device__ __forceinline__ uint32_t pub() const { return (F >> 5) & 3u; }
    __device__ __forceinline__ bool need_reset() const { return F & kFNeedReset; }
    __device__ __forceinline__ uint32_t sigma() const { return (F >> 8) & 31u; }
    __device__ __forceinline__ uint32_t fin0() const { return (F >> 13) & 3u; }
    __device__ __forceinline__ bool terminated() const { return F & kFTerm; }
    __device__ __forceinline__ uint32_t obs_a() const { return H | (PA & cmask); }
    __device__ __forceinline__ uint32_t obs_o() const { return H | (PO & cmask); }

    __device__ __forceinline__ static uint32_t make_p(const NfspW &g, int q) {
        const uint32_t c = g.card(q), row0 = 1u << c, row1 = row0 | (1u << g.pub());
        return g.bets(q) | (g.policy(q) << 4) | (g.last_a(q) << 5) | ((uint32_t)g.acted_nz(q) << 7) | (g.t_snap(q) << 8) |
               (c << 11) | (g.pub() << 13) | (showdown_code(showdown_winner(g.card(0), g.card(1), g.pub()), (uint32_t)q) << 15) |
               (row0 << 24) | (row1 << 27);
    }

    __device__ __forceinline__ void unpack(uint64_t w) {
        const NfspW g{w};
        H = g.hist();
        const uint32_t both = H | (H >> 12);
        const uint32_t id0 = seq_id6(both & 63u), id1 = seq_id6((both >> 6) & 63u);
        const uint32_t r = g.round(), q = (uint32_t)g.to_act();
        // a hand that finished without the re-deal flag (stepped with auto_reset = 0 before) is re-dealt all the same:
        // the kernels that use this representation always run with auto re-deal
        F = (3u * r + g.k()) | (g.dealer() << 3) | (q << 4) | (g.pub() << 5) | ((uint32_t)(g.need_reset() || g.terminated()) << 7) |
            ((r ? 9u + id1 : id0) << 8) | ((r ? (id0 >> 1) - 1u : 0u) << 13) | ((uint32_t)g.terminated() << 16) |
            (g.term_actor() << 17) | (g.outcome() << 18) | ((uint32_t)g.anomaly() << 20);
        PA = make_p(g, (int)q);
        PO = make_p(g, (int)(q ^ 1u));
        SA = g.snapshot((int)q);
        SO = g.snapshot((int)(q ^ 1u));
        cmask = r ? kRound1Cards : kRound0Cards;
    }

    __device__ __forceinline__ uint64_t pack() const {
        const uint32_t q = p();
        const uint32_t P0 = q ? PO : PA, P1 = q ? PA : PO;
        const uint32_t t = tt(), r = t >= 3u ? 1u : 0u, k = t - 3u * r;
        const uint32_t lo = H | (((P0 >> 11) & 3u) << 24) | (((P1 >> 11) & 3u) << 26) | (pub() << 28) | (dealer() << 30) | (r << 31);
        const uint32_t hi = k | ((P0 & 15u) << 2) | ((P1 & 15u) << 6) | (((F >> 16) & 1u) << 10) | (((F >> 7) & 1u) << 11) |
                            (((P0 >> 4) & 1u) << 12) | (((P1 >> 4) & 1u) << 13) | (((P0 >> 5) & 3u) << 14) |
                            (((P1 >> 5) & 3u) << 16) | (((P0 >> 7) & 1u) << 18) | (((P1 >> 7) & 1u) << 19) |
                            (((P0 >> 8) & 7u) << 20) | (((P1 >> 8) & 7u) << 23) | (((F >> 17) & 1u) << 26) |
                            (((F >> 18) & 3u) << 27) | (((F >> 20) & 1u) << 30);
        return (uint64_t)lo | ((uint64_t)hi << 32);
    }

    // newenv.py:76-114 + main.py:28-45: the other player deals, fresh cards, per-hand policy draws.
    // q0 / q1 = deal-table words of players 0 / 1
    __device__ __forceinline__ void redeal(uint32_t q0, uint32_t q1, uint32_t pol0, uint32_t pol1) {
        const uint32_t d = dealer() ^ 1u;  // the dealer opens: it is the player to act
        PA = (d ? q1 : q0) | 1u | ((d ? pol1 : pol0) << 4);  // small blind 0.5
        PO = (d ? q0 : q1) | 2u | ((d ? pol0 : pol1) << 4);  // big blind 1.0
        H = 0u;
        SA = 0u;
        SO = 0u;
        F = d * 0x18u | (((q0 >> 13) & 3u) << 5);
        cmask = kRound0Cards;
    }

    // newenv.py:192-349 for the player to act.  raw = np.argmax(action), nz = (np.average(action) != 0).
    // Returns the step-table entry (bits 0-1 the effective action, bits 2-4 the half chips added); on return the
    // turn has passed (PA is the next player to act) unless the hand terminated.  Sets need_reset with terminated
    // (rollout mode, main.py:55-67 ends the hand).
    __device__ __forceinline__ uint32_t step(const uint32_t *step_lut, int raw, bool nz) {
        const uint32_t e = step_lut[sigma() * 3u + (uint32_t)raw];
        const uint32_t q = p(), kind = (e >> 10) & 3u;
        SA = obs_a();
        PA = ((PA & ~0x7E0u) | ((uint32_t)raw << 5) | ((uint32_t)nz << 7) | ((e >> 4) & 0x700u)) + ((e >> 2) & 7u);
        H |= ((e >> 5) & 1u) << (12u * q + ((e >> 6) & 15u));
        F = (F & ~kFStepMask) | (e >> 16);
        // branch-free tail: a warp's 32 games are in different phases, so branches would run every side anyway
        cmask = kind == KIND_ROUND ? kRound1Cards : cmask;
        const bool term = kind >= KIND_SHOWDOWN;
        const uint32_t oc = kind != KIND_SHOWDOWN ? 0u : (PA >> 15) & 3u;  // 1 the actor wins, 2 the opponent, 3 draw; 0 = fold
        F |= term ? (kFTerm | kFNeedReset | (q << 17) | (oc << 18)) : 0u;
        // round 1 opens with the dealer, otherwise players alternate; nobody moves after the end of the hand
        const bool pass = !term && (kind == KIND_PASS || q != dealer());
        F ^= pass ? 1u << 4 : 0u;
        const uint32_t a = PA, b = SA;
        PA = pass ? PO : a; PO = pass ? a : PO;
        SA = pass ? SO : b; SO = pass ? b : SO;
        return e;
    }

    // 12-byte trace record, word 3 (layout: DESIGN.md "trace record"), same value as NfspW::trace_misc
    __device__ __forceinline__ uint32_t trace_misc(int raw, int eff, bool started) const {
        const uint32_t q = p(), P0 = q ? PO : PA, P1 = q ? PA : PO, r = tt() >= 3u ? 1u : 0u;
        return (uint32_t)raw | ((uint32_t)eff << 2) | (r << 4) | (dealer() << 5) | (((P0 >> 11) & 3u) << 6) |
               (((P1 >> 11) & 3u) << 8) | (pub() << 10) | (r << 12) | ((P0 & 15u) << 13) | ((P1 & 15u) << 17) |
               ((uint32_t)started << 21) | (((P0 >> 4) & 1u) << 22) | (((P1 >> 4) & 1u) << 23);
    }

    // rewards in half chips of the actor / the opponent of the terminating step (newenv.py:250-298)
    __device__ __forceinline__ void rewards(int &ra, int &ro) const {
        const int ba = (int)(PA & 15u), bo = (int)(PO & 15u);
        const uint32_t oc = (F >> 18) & 3u;
        ra = oc == 0u ? -ba : (oc == 1u ? bo : (oc == 2u ? -bo : 0));
        ro = oc == 0u ? ba : (oc == 1u ? -ba : (oc == 2u ? ba : 0));
    }
    // reward of the actor only, 0 while the hand is live (newenv.py:125-129); branch-free
    __device__ __forceinline__ int reward_actor() const {
        const int ba = (int)(PA & 15u), bo = (int)(PO & 15u);
        const uint32_t oc = (F >> 18) & 3u;
        const int mag = oc == 0u ? ba : bo, v = oc == 1u ? mag : -mag;
        return (terminated() && oc != 3u) ? v : 0;
    }
};

}  // namespace nfsp
