"""Drop-in for the reference's leduc/newenv.py `Env` (the env main.py actually runs), backed by
the batched CUDA engine.  With n_games == 1 (default) it keeps the reference's numpy shapes; with
n_games > 1 every method takes / returns arrays with a leading games axis.
"""
import numpy as np
import torch

from ..batched import BatchedNfspEnv, expand_obs
from ..config import load_config


def action_code(action):
    """np.argmax(action) (newenv.py:135) plus code 3 for the all-zero vector (agent.py:134 gate)."""
    a = np.asarray(action, dtype=np.float64).reshape(-1, 3)
    code = np.argmax(a, axis=1).astype(np.int8)
    code[(a == 0).all(axis=1)] = 3
    return code


class Env:
    def __init__(self, n_games=1, seed=None, device=None, config_path="./config.ini"):
        self.config = load_config(config_path)
        cfg = self.config
        self.player_count = cfg.getint("Environment", "Playercount")
        self.decksize = cfg.getint("Environment", "Decksize")
        self.max_rounds = cfg.getint("Environment", "MaxRounds")
        self.suits = cfg.getint("Environment", "Suits")
        self.max_raises = cfg.getint("Environment", "MaxRaises")
        self._action_space = cfg.getint("Environment", "ActionSpace")
        self.total_action_space = cfg.getint("Environment", "TotalActionSpace")
        seed = cfg.getint("Utils", "Seed") if seed is None else seed
        self.n_games = int(n_games)
        self.batched = BatchedNfspEnv(self.n_games, seed=seed, device=device, eta=cfg.getfloat("Agent", "Eta"))
        self.dealer = 0
        self.last_action = np.zeros((self.n_games, self.player_count, self.total_action_space))

    # ---- properties (newenv.py:59-74)
    @property
    def round_index(self):
        r = self.batched.export()["round"].cpu().numpy()
        return int(r[0]) if self.n_games == 1 else r

    @property
    def action_space(self):
        return (self.total_action_space,)

    @property
    def observation_space(self):
        return (1, 30)

    @property
    def terminated(self):
        t = self.batched.export()["terminated"].cpu().numpy().astype(bool)
        return bool(t[0]) if self.n_games == 1 else t

    # ---- newenv.py:76-114
    def reset(self, dealer):
        self.dealer = dealer
        self.last_action[:] = 0
        self.batched.reset(dealer)

    def load_hand(self, dealer, ranks_in_pop_order):
        """Replay mode: ranks popped by player 0, player 1, public card (deck.py:49-50)."""
        d = np.broadcast_to(np.asarray(dealer, np.int8), (self.n_games,))
        c = np.broadcast_to(np.asarray(ranks_in_pop_order, np.int8), (self.n_games, 3))
        self.dealer = dealer
        self.last_action[:] = 0
        self.batched.set_hands(d.copy(), c.copy())

    # ---- newenv.py:116-129
    def get_state(self, p_index):
        s, _, r, s2, t = self.batched.get_state(p_index)
        s_d = expand_obs(s).cpu().numpy().astype(np.float64)
        s2_d = expand_obs(s2).cpu().numpy().astype(np.float64)
        r, t = r.cpu().numpy().astype(np.float64), t.cpu().numpy().astype(bool)
        if self.n_games == 1:
            p = int(p_index)
            return (s_d.reshape(1, 30), self.last_action[0, p].reshape(1, 1, 3).copy(), float(r[0]) if t[0] else 0,
                    s2_d.reshape(1, 1, 30), bool(t[0]))
        pl = np.broadcast_to(np.asarray(p_index), (self.n_games,))
        a = self.last_action[np.arange(self.n_games), pl]
        return s_d, a.copy(), r, s2_d, t

    # ---- newenv.py:192-349
    def step(self, action, p_index):
        a = np.asarray(action, dtype=np.float64).reshape(self.n_games, 3)
        pl = np.broadcast_to(np.asarray(p_index, np.int8), (self.n_games,)).copy()
        live = ~np.atleast_1d(self.terminated)
        self.batched.step(actions=action_code(a), players=pl, n_steps=1, auto_reset=False)
        idx = np.arange(self.n_games)[live]
        self.last_action[idx, pl[live]] = a[live]  # newenv.py:136 runs only on a live hand

    def do_action(self, action, p_index):
        """newenv.py:131-178 on its own (step() runs it fused with the rest on the device): the raise->call coercions,
        the history bit, round_raises, the chips and last_action; returns True when the player folds.  Like the
        reference's method it neither reads nor sets `terminated`, takes no snapshot and never changes the round."""
        a = np.asarray(action, dtype=np.float64).reshape(self.n_games, 3)
        pl = np.broadcast_to(np.asarray(p_index, np.int8), (self.n_games,)).copy()
        fold = self.batched.do_action(action_code(a), pl).cpu().numpy()
        self.last_action[np.arange(self.n_games), pl] = a  # newenv.py:136
        return bool(fold[0]) if self.n_games == 1 else fold

    def game_or_round_has_terminated(self):
        """newenv.py:180-190 on the current round's actions; None (falsy) where the reference returns None."""
        e = {k: v.cpu().numpy() for k, v in self.batched.export().items()}
        out = []
        for g in range(self.n_games):
            h, r, k = int(e["hist"][g]), int(e["round"][g]), int(e["k"][g])
            rnd = ((h | (h >> 12)) >> (6 * r)) & 0x3F
            acts = ["Raise" if (rnd >> (2 * j + 1)) & 1 else "Call" for j in range(k)]
            if len(acts) == 2:
                out.append(True if acts[1] == "Call" else None)
            elif len(acts) == 3:
                out.append(True if acts[1] == "Raise" and acts[2] == "Call" else None)
            else:
                out.append(False)
        return out[0] if self.n_games == 1 else out
