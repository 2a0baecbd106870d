// ENV_LEGACY rule set: leduc/env.py (the README's documented simultaneous-move env, 3-int state
// [card, public=-1, pot]) on one packed 64-bit word per game.  Bit-exactness includes the
// reference's quirks: the public card is never dealt (env.py:97-101 is dead code), a fold does
// not decide the winner, get_new_state() overwrites a player's terminal flag with the
// opponent's (env.py:174) and accumulates the reward on every call (env.py:184-199).
//
// word layout
//   0-1 c0   2-3 c1                      card ranks (cardmatrix.py:8)
//   4-6 left[0]+1   7-9 left[1]+1        _left_choices in [-1,4]
//  10-12 pot[0]    13-15 pot[1]          _pot
//  16 terminal[0]  17 terminal[1]        _specific_state[p][3]
//  18-21 state pot of p0  22-25 of p1    _specific_state[p][0][0][2] (pot sum as of p's last call)
//  26-27 stored action p0 28-29 p1       argmax of _specific_state[p][1]; 3 = never acted
//  30 need_reset (rollout mode)          31 reward overflow
//  32-47 reward[0] int16  48-63 reward[1] int16   _specific_state[p][2]
#pragma once
#include <cstdint>

namespace nfsp {

struct LegacyW {
    uint64_t w;

    __host__ __device__ __forceinline__ int card(int p) const { return (int)((w >> (2 * p)) & 3u); }
    __host__ __device__ __forceinline__ int left(int p) const { return (int)((w >> (4 + 3 * p)) & 7u) - 1; }
    __host__ __device__ __forceinline__ int pot(int p) const { return (int)((w >> (10 + 3 * p)) & 7u); }
    __host__ __device__ __forceinline__ int terminal(int p) const { return (int)((w >> (16 + p)) & 1u); }
    __host__ __device__ __forceinline__ int st_pot(int p) const { return (int)((w >> (18 + 4 * p)) & 15u); }
    __host__ __device__ __forceinline__ int st_action(int p) const { return (int)((w >> (26 + 2 * p)) & 3u); }
    __host__ __device__ __forceinline__ bool need_reset() const { return (w >> 30) & 1u; }
    __host__ __device__ __forceinline__ int reward(int p) const { return (int)(int16_t)(uint16_t)(w >> (32 + 16 * p)); }

    __host__ __device__ __forceinline__ void set_left(int p, int v) {
        w = (w & ~(7ull << (4 + 3 * p))) | ((uint64_t)(uint32_t)(v + 1) << (4 + 3 * p));
    }
    __host__ __device__ __forceinline__ void set_terminal(int p, int v) {
        w = (w & ~(1ull << (16 + p))) | ((uint64_t)v << (16 + p));
    }
    __host__ __device__ __forceinline__ void set_st_pot(int p, int v) {
        w = (w & ~(15ull << (18 + 4 * p))) | ((uint64_t)v << (18 + 4 * p));
    }
    __host__ __device__ __forceinline__ void set_reward(int p, int v) {
        if (v > 32767 || v < -32768) {
            w |= 1ull << 31;
            v = v > 0 ? 32767 : -32768;
        }
        w = (w & ~(0xFFFFull << (32 + 16 * p))) | ((uint64_t)(uint16_t)(int16_t)v << (32 + 16 * p));
    }

    // env.py:46-72: Choices = 4 (config.ini:22), pots 0, rewards 0, stored action "3" (env.py:66)
    __host__ __device__ __forceinline__ void reset(uint32_t c0, uint32_t c1) {
        w = (uint64_t)c0 | ((uint64_t)c1 << 2) | (5ull << 4) | (5ull << 7) | (3ull << 26) | (3ull << 28);
    }

    // env.py:84-158; av = np.argmax(action); penalty = config.ini Agent.Penalty
    __host__ __device__ __forceinline__ void step(int av, int p, int penalty) {
        const int o = p ^ 1;
        int lp = left(p), term = 0;
        if (lp > 0) {
            if (av == 0) {  // env.py:114-120
                if (lp == 4 && left(o) == 4) set_reward(p, penalty);
                term = 1;
            } else if (av == 1) {  // env.py:122-128
                lp -= 1;
                if (pot(p) < pot(o)) w += 1ull << (10 + 3 * p);
                term = lp == 0;
            } else {  // env.py:130-139
                if (lp & 1) set_reward(p, penalty);
                lp -= 2;
                if (pot(p) <= pot(o)) w += 1ull << (10 + 3 * p);
                term = lp == 0;
            }
            set_left(p, lp);
        } else {
            term = 1;  // env.py:140-141
        }
        set_terminal(p, term);
        set_st_pot(p, pot(0) + pot(1));
        w = (w & ~(3ull << (26 + 2 * p))) | ((uint64_t)av << (26 + 2 * p));
    }

    // env.py:160-205
    __host__ __device__ __forceinline__ void get_new_state(int p) {
        const int o = p ^ 1;
        set_st_pot(p, pot(0) + pot(1));
        const int t = terminal(o);
        set_terminal(p, t);  // env.py:174
        if (t) {
            const int cp = card(p), co = card(o);
            // env.py:179,186: the public card is the fake rank -1 and never matches
            if (cp > co) set_reward(p, reward(p) - pot(p));       // env.py:189-190
            else if (cp < co) set_reward(p, reward(p) + pot(o));  // env.py:192-199
        }
    }
};

}  // namespace nfsp
