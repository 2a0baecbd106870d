// ENV_NFSP in registers, actor-relative: the representation the hot kernels (fused rollout, env-only fast
// path) keep across the steps of a launch.  The packed 64-bit word of nfsp_rules.cuh stays the HBM format
// (one load + one store per game per launch); here the same fields are spread over 32-bit registers laid
// out so that one decision costs a few dozen integer instructions:
//
//   * per-player fields live in two words PA (the player to act) / PO (the other one) that SWAP when the turn
//     passes, so no field is ever addressed with a variable shift;
//   * the slot index tt = 3*round + k is kept directly (history bit = 12*player + 2*tt + raise, newenv.py:53);
//   * the betting sequence of the current round is kept as a sequence id sq (0 -, 1 C, 2 R, 3 CC, 4 CR, 5 RC,
//     6 RR, 7 CRC, 8 RRC): the raise->call coercions (newenv.py:141-145), "previous action was a raise"
//     (newenv.py:157,169) and "round over" (newenv.py:180-190) are comparisons on sq;
//   * the snapshots s[p] (newenv.py:200-202) are explicit registers (the observation the player acted on).
//
// Semantics are those of NfspW::step / decide_begin / decide_finish; tests run the golden hands and seeded
// rollouts through both.
#pragma once
#include "nfsp_rules.cuh"
#include "philox.cuh"

namespace nfsp {

// per-player word P
//   0-3  bets in half chips      4 policy ('b' = 1)      5-6 last raw action     7 acted with a non-zero vector
//   8-10 time of the last step() call (7 = never)        11-12 card rank       13-14 public card rank (copy)
//   24-26 one-hot private card (obs bits 24-26)          27-29 private | public one-hot (obs bits 27-29, shown in round 1)
constexpr uint32_t kPBets = 0xFu, kPPol = 1u << 4, kPNz = 1u << 7;
constexpr uint32_t kRound0Cards = 0x07000000u, kRound1Cards = 0x3F000000u;

// flag word F
//   0-2 tt = 3*round + k (k up to 3 after a showdown)    3 dealer    4 player to act    5-6 public card rank
//   7 need_reset    8-11 sq (current round)    12-15 sq0 (round 0; == sq while in round 0)
//   16 terminated   17 actor of the terminating step     18-19 outcome   20 anomaly
constexpr uint32_t kFNeedReset = 1u << 7, kFTerm = 1u << 16;

__device__ __forceinline__ uint32_t seq_id6(uint32_t rnd) {  // 6 slot bits of one round -> sequence id 0..8
    const uint32_t s0 = rnd & 3u, s1 = (rnd >> 2) & 3u, s2 = (rnd >> 4) & 3u;
    return s2 ? 6u + s0 : (s1 ? 2u * s0 + s1 : s0);
}

// Deal table in shared memory: deal index 0..119 (deck.py:35-50: uniform ordered draw of 3 of 6 cards) ->
// the card part of both players' P words, lut[idx] for player 0 and lut[120 + idx] for player 1:
//   7 << 8 (never acted) | card << 11 | public card << 13 | one-hot rows (bits 24-29)
constexpr int kDealLutWords = 240;
__device__ __forceinline__ uint32_t deal_word(uint32_t c, uint32_t pub) {
    const uint32_t row0 = 1u << c;
    return (7u << 8) | (c << 11) | (pub << 13) | (row0 << 24) | ((row0 | (1u << pub)) << 27);
}
__device__ __forceinline__ void fill_deal_lut(uint32_t *lut) {
    for (uint32_t i = threadIdx.x; i < 120u; i += blockDim.x) {
        const uint32_t c = deal_ranks(i);
        lut[i] = deal_word(c & 3u, (c >> 4) & 3u);
        lut[120u + i] = deal_word((c >> 2) & 3u, (c >> 4) & 3u);
    }
}

struct NfspFast {
    uint32_t H, F, PA, PO, SA, SO, cmask;

    __device__ __forceinline__ uint32_t tt() const { return F & 7u; }
    __device__ __forceinline__ uint32_t dealer() const { return (F >> 3) & 1u; }
    __device__ __forceinline__ uint32_t p() const { return (F >> 4) & 1u; }
    __device__ __forceinline__ uint32_t pub() const { return (F >> 5) & 3u; }
    __device__ __forceinline__ bool need_reset() const { return F & kFNeedReset; }
    __device__ __forceinline__ uint32_t sq() const { return (F >> 8) & 15u; }
    __device__ __forceinline__ uint32_t sq0() const { return (F >> 12) & 15u; }
    __device__ __forceinline__ bool terminated() const { return F & kFTerm; }
    __device__ __forceinline__ uint32_t obs_a() const { return H | (PA & cmask); }
    __device__ __forceinline__ uint32_t obs_o() const { return H | (PO & cmask); }

    __device__ __forceinline__ static uint32_t make_p(const NfspW &g, int q) {
        const uint32_t c = g.card(q), row0 = 1u << c, row1 = row0 | (1u << g.pub());
        return g.bets(q) | (g.policy(q) << 4) | (g.last_a(q) << 5) | ((uint32_t)g.acted_nz(q) << 7) | (g.t_snap(q) << 8) |
               (c << 11) | (row0 << 24) | (row1 << 27);
    }

    __device__ __forceinline__ void unpack(uint64_t w) {
        const NfspW g{w};
        H = g.hist();
        const uint32_t both = H | (H >> 12);
        const uint32_t id0 = seq_id6(both & 63u), id1 = seq_id6((both >> 6) & 63u);
        const uint32_t r = g.round(), q = (uint32_t)g.to_act();
        F = (3u * r + g.k()) | (g.dealer() << 3) | (q << 4) | (g.pub() << 5) | ((uint32_t)g.need_reset() << 7) |
            ((r ? id1 : id0) << 8) | (id0 << 12) | ((uint32_t)g.terminated() << 16) | (g.term_actor() << 17) |
            (g.outcome() << 18) | ((uint32_t)g.anomaly() << 20);
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
    // Returns the effective action; on return the turn has passed (PA is the next player to act) unless the
    // hand terminated.  Sets need_reset with terminated (rollout mode, main.py:55-67 ends the hand).
    __device__ __forceinline__ int step(int raw, bool nz) {
        const uint32_t t = tt(), s = sq(), q = p();
        SA = obs_a();
        PA = (PA & ~0x7E0u) | ((uint32_t)raw << 5) | ((uint32_t)nz << 7) | (t << 8);
        int av = raw;
        if (av == A_RAISE && s >= 3u) av = A_CALL;  // newenv.py:141-145: [C,R] and "p already raised" both mean s in {4,6}
        if (av == A_FOLD) {
            F |= kFTerm | kFNeedReset | (q << 17);  // outcome 0
            return av;
        }
        const bool prev_raise = s != 0u && !(s & 1u);                                          // s in {2,4,6}
        PA += (av == A_RAISE ? 2u : 0u) + (prev_raise ? 2u : 0u) + (t == 0u ? 1u : 0u);        // newenv.py:157-176
        H |= 1u << (12u * q + 2u * t + (uint32_t)av - 1u);
        const uint32_t sn = s < 3u ? 2u * s + (uint32_t)av : 5u + (s >> 1);
        const bool over = s >= 3u || (s != 0u && av == A_CALL);                                // newenv.py:180-190
        if (!over) {  // same round, the turn passes
            uint32_t f = ((F & ~0xF00u) + 1u) | (sn << 8);
            if (t < 3u) f = (f & ~0xF000u) | (sn << 12);
            F = f ^ (1u << 4);
            swap_players();
            return av;
        }
        if (t < 3u) {  // newenv.py:215-242: the public card is revealed, round 1 starts with the dealer
            F = (F & ~0xFF07u) | 3u | (sn << 12);
            cmask = kRound1Cards;
            if (q != dealer()) {
                F ^= 1u << 4;
                swap_players();
            }
            return av;
        }
        // showdown, newenv.py:261-298
        const uint32_t cp = (PA >> 11) & 3u, co = (PO >> 11) & 3u, pb = pub();
        const uint32_t oc = cp == pb ? 1u : (co == pb ? 2u : (cp < co ? 1u : (cp > co ? 2u : 3u)));
        F = (((F & ~0xF00u) + 1u) | (sn << 8)) | kFTerm | kFNeedReset | (q << 17) | (oc << 18);
        return av;
    }

    __device__ __forceinline__ void swap_players() {
        const uint32_t a = PA, b = SA;
        PA = PO; PO = a;
        SA = SO; SO = b;
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
};

}  // namespace nfsp
