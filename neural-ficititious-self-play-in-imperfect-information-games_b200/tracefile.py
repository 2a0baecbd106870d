"""Trace files: hands on disk + a replay/check CLI.

A hand of leduc/newenv.py under main.train's turn order is fully described by who deals, the three
dealt ranks, the two per-hand policies and at most six raw actions: one uint32 per hand,

    bit 0 dealer | 1-2 c0 | 3-4 c1 | 5-6 public | 7 policy0 | 8 policy1 | 9-11 number of actions |
    12-23 raw actions, 2 bits each, in the order they were taken

next to the per-transition outputs of the kernel that played it (the three trace planes of DESIGN.md,
restricted to the hand's transitions).  Parity is defined on such traces, never on seeds (the reference
shares one global `random` stream between deck, policies and buffers).

    python -m nfsp_b200.tracefile record --games 4096 --steps 32 --out hands.npz
    python -m nfsp_b200.tracefile check hands.npz

`record` plays Philox-dealt uniform-random hands with the state-machine env kernel (nfsp_step_fsm_kernel) and writes
every hand that started and finished inside the window.  `check` replays the file through the GENERAL env
kernel (explicit deck via set_hands, explicit action codes, no auto re-deal) and compares every transition
word for word: the two kernels share no game logic beyond the packed word, so this is an independent
end-to-end check that needs nothing but the GPU; the parity tests additionally replay the same file on the CPU.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np
import torch

MAGIC = "leduc-newenv-hands-v1"
STARTED_BIT = 1 << 21


def pack_hand(dealer, c0, c1, pub, pol0, pol1, actions):
    w = int(dealer) | (int(c0) << 1) | (int(c1) << 3) | (int(pub) << 5) | (int(pol0) << 7) | (int(pol1) << 8) | (len(actions) << 9)
    for k, a in enumerate(actions):
        w |= int(a) << (12 + 2 * k)
    return w


def unpack_hands(words):
    w = np.asarray(words, dtype=np.uint32)
    n_act = ((w >> 9) & 7).astype(np.int64)
    acts = np.stack([((w >> (12 + 2 * k)) & 3) for k in range(6)], axis=1).astype(np.int8)
    return dict(dealer=(w & 1).astype(np.int8), cards=np.stack([(w >> 1) & 3, (w >> 3) & 3, (w >> 5) & 3], 1).astype(np.int8),
                policy=np.stack([(w >> 7) & 1, (w >> 8) & 1], 1).astype(np.int8), n_actions=n_act, actions=acts)


def hands_from_trace(planes):
    """planes uint32 [3, T, n] (host) -> (hand words [m], transitions uint32 [m, 6, 3]) for every hand that
    starts and ends inside the window."""
    obs, rew, misc = (np.asarray(planes[k]).view(np.uint32) for k in range(3))
    T, n = obs.shape
    word = np.zeros(n, np.uint32)
    length = np.zeros(n, np.int64)
    active = np.zeros(n, bool)
    rows = np.zeros((n, 6, 3), np.uint32)
    out_w, out_t = [], []
    for t in range(T):
        m = misc[t]
        started = (m & STARTED_BIT) != 0
        if started.any():
            s = started
            word[s] = ((m[s] >> 5) & 1) | (((m[s] >> 6) & 3) << 1) | (((m[s] >> 8) & 3) << 3) | (((m[s] >> 10) & 3) << 5) | \
                      (((m[s] >> 22) & 1) << 7) | (((m[s] >> 23) & 1) << 8)
            length[s] = 0
            active[s] = True
            rows[s] = 0
        a = np.nonzero(active)[0]
        k = length[a]
        word[a] |= (m[a] & 3) << (12 + 2 * k).astype(np.uint32)
        rows[a, k, 0], rows[a, k, 1], rows[a, k, 2] = obs[t, a], rew[t, a], m[a] & ~np.uint32(STARTED_BIT)
        length[a] = k + 1
        done = a[((obs[t, a] >> 30) & 1) != 0]
        if done.size:
            out_w.append(word[done] | (length[done].astype(np.uint32) << 9))
            out_t.append(rows[done].copy())
            active[done] = False
    if not out_w:
        return np.zeros(0, np.uint32), np.zeros((0, 6, 3), np.uint32)
    return np.concatenate(out_w), np.concatenate(out_t)


def save(path, words, transitions, **meta):
    np.savez_compressed(path, magic=MAGIC, hands=np.asarray(words, np.uint32), transitions=np.asarray(transitions, np.uint32),
                        **{k: np.asarray(v) for k, v in meta.items()})


def load(path):
    with np.load(path) as z:
        if str(z["magic"]) != MAGIC:
            raise ValueError("%s is not a %s file" % (path, MAGIC))
        return {k: z[k] for k in z.files}


def action_codes(h):
    """int8 [6, m] for nfsp_env_step: the raw action while the hand lasts, 4 = the game sits the step out."""
    k = np.arange(6)[:, None]
    return np.where(k < h["n_actions"][None, :], h["actions"].T, 4).astype(np.int8)


def replay_on_gpu(words, device=None):
    """Replays hands through the general env kernel; returns transitions uint32 [m, 6, 3] (zeros past a hand's end)."""
    from .batched import BatchedNfspEnv

    h = unpack_hands(words)
    m = len(h["dealer"])
    env = BatchedNfspEnv(m, device=device)
    env.set_hands(h["dealer"], h["cards"], h["policy"])
    tr = env.step(actions=action_codes(h), n_steps=6, auto_reset=False, trace=True)["raw"]
    got = tr.cpu().numpy().view(np.uint32).transpose(2, 1, 0).copy()  # [m, 6, 3]
    got[:, :, 2] &= ~np.uint32(STARTED_BIT)
    got[np.arange(6)[None, :] >= h["n_actions"][:, None]] = 0
    return got


def record(games, steps, seed=1234, eta=0.1, device=None):
    from .batched import BatchedNfspEnv

    env = BatchedNfspEnv(games, seed=seed, device=device, eta=eta)
    env.reset()
    planes = env.step(n_steps=steps, trace=True)["raw"].cpu().numpy()
    return hands_from_trace(planes)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)
    r = sub.add_parser("record")
    r.add_argument("--games", type=int, default=4096)
    r.add_argument("--steps", type=int, default=32)
    r.add_argument("--seed", type=int, default=1234)
    r.add_argument("--out", required=True)
    c = sub.add_parser("check")
    c.add_argument("path")
    args = ap.parse_args(argv)
    if args.cmd == "record":
        words, trans = record(args.games, args.steps, args.seed)
        save(args.out, words, trans, games=args.games, steps=args.steps, seed=args.seed)
        print("%s: %d hands, %d transitions" % (args.out, len(words), int(unpack_hands(words)["n_actions"].sum())))
        return 0
    f = load(args.path)
    got = replay_on_gpu(f["hands"])
    bad = np.nonzero((got != f["transitions"]).any(axis=(1, 2)))[0]
    print("%s: %d hands replayed, %d differ" % (args.path, len(f["hands"]), len(bad)))
    if len(bad):
        i = int(bad[0])
        print("first difference: hand %d word 0x%08x\n recorded %s\n replayed %s" % (i, int(f["hands"][i]), f["transitions"][i], got[i]))
    return 1 if len(bad) else 0


if __name__ == "__main__":
    sys.exit(main())
