// ENV_NFSP rule set: leduc/newenv.py (alternating-move Leduc with blinds, 30-bit observation)
// plus the turn order / terminal-observation logic of main.py:28-67, on ONE packed 64-bit word
// per game.  Everything the reference keeps in numpy arrays and Python lists is either a bit
// field here or derived from the 24 history bits on the fly.
//
// word layout (bit: field)
//   0-23  history  == observation bits 0-23: bit q*12 + r*6 + k*2 + b  (newenv.py:53,155,167)
//  24-25  c0   26-27 c1   28-29 pub   card ranks 0=Ace 1=King 2=Queen (pub is dealt at reset, revealed in round 1)
//  30     dealer                      31     round
//  32-33  k = round_raises (actions taken this round, newenv.py:156,173)
//  34-37  bets[0] in half chips       38-41  bets[1]           (overall_raises, newenv.py:80-82)
//  42     terminated                  43     need_reset (rollout mode: hand over, re-deal on next step)
//  44     policy[0]  45 policy[1]     1 = 'b' best response, 0 = 'a' average (main.py:38-45)
//  46-47  last raw action of p0       48-49  of p1             (last_action, newenv.py:136; argmax only)
//  50     p0 acted with a non-zero vector   51 same for p1     (np.average(a) != 0 gate, agent.py:134)
//  52-54  t0   55-57  t1   time 3*round+k of the player's last step() call, 7 = never: the snapshot
//                          s[p] (newenv.py:200-202) is history & bits-older-than-t plus the card rows
//  58     actor of the terminating step    59-60  outcome: 0 fold, 1 actor wins, 2 opponent wins, 3 draw
//  62     anomaly: step() on a terminated hand (newenv.py:346-348)
#pragma once
#include <cstdint>

namespace nfsp {

enum : int { A_FOLD = 0, A_CALL = 1, A_RAISE = 2 };

struct NfspW {
    uint64_t w;

    __device__ __forceinline__ uint32_t hist() const { return (uint32_t)w & 0xFFFFFFu; }
    __device__ __forceinline__ uint32_t card(int p) const { return (uint32_t)(w >> (24 + 2 * p)) & 3u; }
    __device__ __forceinline__ uint32_t pub() const { return (uint32_t)(w >> 28) & 3u; }
    __device__ __forceinline__ uint32_t dealer() const { return (uint32_t)(w >> 30) & 1u; }
    __device__ __forceinline__ uint32_t round() const { return (uint32_t)(w >> 31) & 1u; }
    __device__ __forceinline__ uint32_t k() const { return (uint32_t)(w >> 32) & 3u; }
    __device__ __forceinline__ uint32_t bets(int p) const { return (uint32_t)(w >> (34 + 4 * p)) & 15u; }
    __device__ __forceinline__ bool terminated() const { return (w >> 42) & 1u; }
    __device__ __forceinline__ bool need_reset() const { return (w >> 43) & 1u; }
    __device__ __forceinline__ uint32_t policy(int p) const { return (uint32_t)(w >> (44 + p)) & 1u; }
    __device__ __forceinline__ uint32_t last_a(int p) const { return (uint32_t)(w >> (46 + 2 * p)) & 3u; }
    __device__ __forceinline__ bool acted_nz(int p) const { return (w >> (50 + p)) & 1u; }
    __device__ __forceinline__ uint32_t t_snap(int p) const { return (uint32_t)(w >> (52 + 3 * p)) & 7u; }
    __device__ __forceinline__ uint32_t term_actor() const { return (uint32_t)(w >> 58) & 1u; }
    __device__ __forceinline__ uint32_t outcome() const { return (uint32_t)(w >> 59) & 3u; }
    __device__ __forceinline__ bool anomaly() const { return (w >> 62) & 1u; }

    // main.py:55-65: the dealer opens both rounds and the players alternate
    __device__ __forceinline__ int to_act() const { return (int)((k() & 1u) ^ dealer()); }

    // newenv.py:76-114.  cards = c0 | c1<<2 | pub<<4
    __device__ __forceinline__ void reset(uint32_t dealer, uint32_t cards, uint32_t pol0, uint32_t pol1) {
        const uint32_t b0 = dealer == 0 ? 1u : 2u;  // small blind 0.5 / big blind 1.0, in half chips
        const uint32_t b1 = dealer == 0 ? 2u : 1u;
        w = ((uint64_t)(cards & 63u) << 24) | ((uint64_t)dealer << 30) | ((uint64_t)b0 << 34) | ((uint64_t)b1 << 38) |
            ((uint64_t)pol0 << 44) | ((uint64_t)pol1 << 45) | (63ull << 52);
    }

    __device__ __forceinline__ static uint32_t card_rows(uint32_t c, uint32_t pub, bool second_row) {
        const uint32_t row0 = 1u << c;
        const uint32_t row1 = second_row ? (row0 | (1u << pub)) : 0u;
        return (row0 << 24) | (row1 << 27);
    }

    // newenv.py:118: concat(history.flatten(), specific_cards[p].flatten()) as a 30-bit mask
    __device__ __forceinline__ uint32_t obs(int p) const { return hist() | card_rows(card(p), pub(), round() != 0); }

    // the snapshot s[p] taken at the top of p's last step() (newenv.py:200-202); zeros before it
    __device__ __forceinline__ uint32_t snapshot(int p) const {
        const uint32_t t = t_snap(p);
        if (t == 7u) return 0u;
        const uint32_t m = (1u << (2u * t)) - 1u;  // history slots strictly older than t, per player block
        return (hist() & (m | (m << 12))) | card_rows(card(p), pub(), t >= 3u);
    }

    // reward[p] in half chips (newenv.py:250-298), 0 while the hand is live (newenv.py:125-129)
    __device__ __forceinline__ int reward_half(int p) const {
        if (!terminated()) return 0;
        const int a = (int)term_actor(), o = a ^ 1;
        const int ba = (int)bets(a), bo = (int)bets(o);
        int ra, ro;
        switch (outcome()) {
            case 0: ra = -ba; ro = ba; break;   // fold: newenv.py:252-255
            case 1: ra = bo;  ro = -ba; break;  // actor wins the showdown
            case 2: ra = -bo; ro = ba; break;   // opponent wins
            default: ra = 0;  ro = 0; break;    // draw: newenv.py:291-293
        }
        return p == a ? ra : ro;
    }

    // newenv.py:192-349 (+ do_action 131-178, game_or_round_has_terminated 180-190).
    // raw = np.argmax(action) in {0,1,2}; nonzero = (np.average(action) != 0).
    // Returns the effective action after the raise->call coercions.
    __device__ __forceinline__ int step(int raw, bool nonzero, int p) {
        const uint32_t r = round(), kk = k();
        const uint32_t t_now = 3u * r + kk;
        w = (w & ~(7ull << (52 + 3 * p))) | ((uint64_t)t_now << (52 + 3 * p));  // snapshot s[p]
        if (terminated()) {
            w |= 1ull << 62;
            return raw;
        }
        w = (w & ~((3ull << (46 + 2 * p)) | (1ull << (50 + p)))) | ((uint64_t)raw << (46 + 2 * p)) |
            ((uint64_t)nonzero << (50 + p));
        const uint32_t h = hist();
        const uint32_t rnd = (h | (h >> 12)) >> (6u * r);  // slot bits of this round, both players OR-ed
        // bit 2k' = someone called in slot k', bit 2k'+1 = someone raised in slot k'
        const bool p_raised = ((h >> (12u * p + 6u * r)) & 0x2Au) != 0u;  // raises[p] > 0
        int av = raw;
        if (av == A_RAISE && p_raised) av = A_CALL;                                           // newenv.py:141-142
        if (av == A_RAISE && kk == 2u && (rnd & 1u) && (rnd & 8u)) av = A_CALL;               // newenv.py:143-145
        if (av == A_FOLD) {
            w |= (1ull << 42) | ((uint64_t)p << 58);  // outcome 0 = fold
            return av;
        }
        const bool prev_raise = kk > 0u && ((rnd >> (2u * (kk - 1u) + 1u)) & 1u);
        uint32_t add = (av == A_CALL) ? (prev_raise ? 2u : 0u) : (prev_raise ? 4u : 2u);     // newenv.py:157-158,169-172
        if (r == 0u && kk == 0u) add += 1u;                                                  // newenv.py:159-161,174-176
        w |= 1ull << (12u * p + 6u * r + 2u * kk + (av == A_RAISE ? 1u : 0u));
        w += (uint64_t)add << (34 + 4 * p);
        const uint32_t k2 = kk + 1u;
        // round over on [C,C] [R,C] [C,R,C] [R,R,C]: second action a call, or any third action
        const bool over = (k2 == 3u) || (k2 == 2u && av == A_CALL);
        if (!over) {
            w = (w & ~(3ull << 32)) | ((uint64_t)k2 << 32);
            return av;
        }
        if (r == 0u) {  // newenv.py:215-242: reveal the public card, start round 1
            w = (w & ~(3ull << 32)) | (1ull << 31);
            return av;
        }
        // showdown, newenv.py:261-298 (raw != fold here)
        w = (w & ~(3ull << 32)) | ((uint64_t)k2 << 32);
        const uint32_t cp = card(p), co = card(p ^ 1), pb = pub();
        uint32_t oc;
        if (cp == pb) oc = 1u;
        else if (co == pb) oc = 2u;
        else if (cp < co) oc = 1u;
        else if (cp > co) oc = 2u;
        else oc = 3u;
        w |= (1ull << 42) | ((uint64_t)p << 58) | ((uint64_t)oc << 59);
        return av;
    }

    // newenv.py:131-178 ALONE, as a caller of the bare method sees it: the coercions, the history bit, round_raises,
    // the chips and last_action -- no snapshot, no round change, no showdown, and `terminated` is neither read nor set
    // (step() does that with the return value).  Returns true for a fold.  A fold leaves no trace but last_action (the
    // reference appends 'Fold' to actions_done, which only game_or_round_has_terminated reads, and it answers falsy
    // either way); a fourth action in a round (history[p][round][3]: IndexError in the reference) sets the anomaly bit.
    __device__ __forceinline__ bool do_action(int raw, bool nonzero, int p) {
        const uint32_t r = round(), kk = k();
        w = (w & ~((3ull << (46 + 2 * p)) | (1ull << (50 + p)))) | ((uint64_t)raw << (46 + 2 * p)) |
            ((uint64_t)nonzero << (50 + p));
        const uint32_t h = hist();
        const uint32_t rnd = (h | (h >> 12)) >> (6u * r);
        const bool p_raised = ((h >> (12u * p + 6u * r)) & 0x2Au) != 0u;
        int av = raw;
        if (av == A_RAISE && p_raised) av = A_CALL;
        if (av == A_RAISE && kk == 2u && (rnd & 1u) && (rnd & 8u)) av = A_CALL;
        if (av == A_FOLD) return true;
        if (kk == 3u) {
            w |= 1ull << 62;
            return false;
        }
        const bool prev_raise = kk > 0u && ((rnd >> (2u * (kk - 1u) + 1u)) & 1u);
        uint32_t add = (av == A_CALL) ? (prev_raise ? 2u : 0u) : (prev_raise ? 4u : 2u);
        if (r == 0u && kk == 0u) add += 1u;
        w |= 1ull << (12u * p + 6u * r + 2u * kk + (av == A_RAISE ? 1u : 0u));
        w += (uint64_t)add << (34 + 4 * p);
        w = (w & ~(3ull << 32)) | ((uint64_t)(kk + 1u) << 32);
        return false;
    }

    // 12-byte trace record, word 3 (layout: DESIGN.md "trace record")
    __device__ __forceinline__ uint32_t trace_misc(int raw, int eff, bool started) const {
        return (uint32_t)raw | ((uint32_t)eff << 2) | (round() << 4) | (dealer() << 5) | (card(0) << 6) |
               (card(1) << 8) | (pub() << 10) | (round() << 12) | (bets(0) << 13) | (bets(1) << 17) |
               ((uint32_t)started << 21) | (policy(0) << 22) | (policy(1) << 23);
    }
};

}  // namespace nfsp
