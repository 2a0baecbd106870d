// ENV_LEGACY (leduc/env.py) README iteration -- step(a0, 0); step(a1, 1); get_new_state(0); get_new_state(1) -- as a
// finite-state machine, for legacy_rollout_kernel.
//
// What the two steps decide depends on (left[0], left[1], pot[0], pot[1]) and on whether a player already carries the
// -1 penalty (env.py:114-139); from a fresh hand only 70 such states can be reached.  Cards matter only when the hand
// ends (env.py:184-199), and then linearly: with G = (c0 > c1), L = (c0 < c1) and u = L*pot[1] - G*pot[0] the rewards
// after both get_new_state calls are rb0 + u and rb1 - u.  So one iteration is ONE table entry per (state, a0, a1),
// loaded with two LDS.128 and used as loaded:
//   A: nx   shared-memory address of the next state's row
//      w0   record word 1 without the card: 0xFF << 8 (public card -1) | pot sum << 16 | terminal << 24 (both players
//           read the same pot sum and, after get_new_state, the same terminal flag: env.py:174)
//      rb0, rb1  the rewards after the two steps (0 or the penalty)
//   B: w2x, w2y  record word 3 of each player: action | (left + 1) << 2 | pot << 5
//      p0t, p1t  pot[0], pot[1] if the hand ends here, else 0
//   C: lo   bits 4-30 of the packed word after the iteration (read when the game is packed)
// The image is built on the host by running LegacyW itself (legacy_rules.cuh) breadth-first from the reset word, so the
// tables cannot drift from the packed-word rules that every other entry point uses.  A word that no rollout can
// produce (hands stepped by hand through nfsp_legacy_step in another order, accumulated rewards) is played by the
// packed-word code instead.
#pragma once
#include <cstdint>
#include <vector>

#include "legacy_rules.cuh"

namespace nfsp {
namespace lfsm {

constexpr int kEntryBytes = 48, kRowBytes = 9 * kEntryBytes, kMaxRows = 72;
constexpr int kRowsOff = 0;
constexpr int kKeyOff = kRowsOff + kMaxRows * kRowBytes;  // 31104: uint8 row of a 12-bit state key, 0xFF = not a rollout state
constexpr int kKeys = 4096;
constexpr int kDealOff = kKeyOff + kKeys;                 // 35200: deal index -> c0, c1, G, L
constexpr int kImageBytes = kDealOff + 120 * 16;          // 37120
constexpr int kImageWords = kImageBytes / 4;
constexpr int kSharedWords = (kImageBytes - kKeys) / 4;   // what a CTA keeps in shared memory: rows, then deals
constexpr uint32_t kTermBit = 1u << 24;
constexpr uint32_t kLoMask = 0x7FFFFFF0u;  // everything of the low word but the cards and the overflow flag

// left + 1 (3 bits each) | pots (2 bits each) | "carries the penalty" per player; -1 = not representable
__host__ __device__ inline int state_key(const LegacyW &g, int penalty) {
    const int r0 = g.reward(0), r1 = g.reward(1);
    if ((r0 != 0 && r0 != penalty) || (r1 != 0 && r1 != penalty) || g.pot(0) > 3 || g.pot(1) > 3 || ((g.w >> 31) & 1u)) return -1;
    return (g.left(0) + 1) | ((g.left(1) + 1) << 3) | (g.pot(0) << 6) | (g.pot(1) << 8) | ((r0 != 0) << 10) | ((r1 != 0) << 11);
}

inline uint32_t deal_ranks_host(uint32_t idx) {  // deck.py:35-50, as deal_ranks in philox.cuh
    const uint32_t i0 = idx / 20u, r = idx - i0 * 20u, j1 = r >> 2, j2 = r & 3u;
    const uint32_t i1 = j1 + (j1 >= i0 ? 1u : 0u), lo = i0 < i1 ? i0 : i1, hi = i0 < i1 ? i1 : i0;
    uint32_t i2 = j2;
    i2 += (i2 >= lo) ? 1u : 0u;
    i2 += (i2 >= hi) ? 1u : 0u;
    return (i0 >> 1) | ((i1 >> 1) << 2) | ((i2 >> 1) << 4);
}

// Returns the number of rows, or -1 if the reachable set does not fit the image (never with the reference's rules).
inline int build_image(uint32_t *img, int penalty) {
    for (int i = 0; i < kImageWords; ++i) img[i] = 0u;
    uint8_t *keys = reinterpret_cast<uint8_t *>(img) + kKeyOff;
    for (int i = 0; i < kKeys; ++i) keys[i] = 0xFFu;
    std::vector<uint64_t> rows;  // a representative word per state, equal cards so that get_new_state adds nothing
    LegacyW g0;
    g0.reset(0u, 0u);
    rows.push_back(g0.w);
    keys[state_key(g0, penalty)] = 0;
    for (size_t s = 0; s < rows.size(); ++s)
        for (int a0 = 0; a0 < 3; ++a0)
            for (int a1 = 0; a1 < 3; ++a1) {
                LegacyW g{rows[s]};
                g.step(a0, 0, penalty);
                g.step(a1, 1, penalty);
                const int pot0 = g.pot(0), pot1 = g.pot(1);
                const int rb0 = g.reward(0), rb1 = g.reward(1);
                g.get_new_state(0);
                g.get_new_state(1);
                const uint32_t term = (uint32_t)(g.terminal(0) | g.terminal(1));
                if (term) g.w |= 1ull << 30;
                uint32_t nx = 0u;
                if (!term) {
                    const int key = state_key(g, penalty);
                    if (key < 0) return -1;
                    if (keys[key] == 0xFFu) {
                        if ((int)rows.size() >= kMaxRows) return -1;
                        keys[key] = (uint8_t)rows.size();
                        rows.push_back(g.w);
                    }
                    nx = (uint32_t)(kRowsOff + keys[key] * kRowBytes);
                }
                uint32_t *e = img + (kRowsOff + (int)s * kRowBytes + (a0 * 3 + a1) * kEntryBytes) / 4;
                e[0] = nx;
                e[1] = (0xFFu << 8) | ((uint32_t)g.st_pot(0) << 16) | (term << 24);
                e[2] = (uint32_t)rb0;
                e[3] = (uint32_t)rb1;
                e[4] = (uint32_t)a0 | ((uint32_t)(g.left(0) + 1) << 2) | ((uint32_t)pot0 << 5);
                e[5] = (uint32_t)a1 | ((uint32_t)(g.left(1) + 1) << 2) | ((uint32_t)pot1 << 5);
                e[6] = term ? (uint32_t)pot0 : 0u;
                e[7] = term ? (uint32_t)pot1 : 0u;
                e[8] = (uint32_t)g.w & kLoMask;
            }
    for (uint32_t idx = 0; idx < 120u; ++idx) {
        const uint32_t c = deal_ranks_host(idx), c0 = c & 3u, c1 = (c >> 2) & 3u;
        uint32_t *e = img + kDealOff / 4 + 4 * idx;
        e[0] = c0;
        e[1] = c1;
        e[2] = c0 > c1 ? 1u : 0u;
        e[3] = c0 < c1 ? 1u : 0u;
    }
    return (int)rows.size();
}
// word w of the image is a byte offset that becomes a shared-memory address (nx of every entry)
__host__ __device__ inline bool is_address(int w) { return w < kKeyOff / 4 && (w % (kEntryBytes / 4)) == 0; }

}  // namespace lfsm
}  // namespace nfsp
