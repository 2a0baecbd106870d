"""Drop-in for the reference's leduc/env.py `Env` (the README's documented API:
reset / init_state / step(action, player) / get_new_state(player)), backed by the batched CUDA
engine.  n_games == 1 keeps the reference's return shapes."""
import numpy as np

from ..batched import BatchedLegacyEnv
from ..config import load_config


class Env:
    def __init__(self, n_games=1, seed=None, device=None, config_path="./config.ini"):
        self.config = load_config(config_path)
        self.player_count = self.config.getint("Environment", "Playercount")
        self._decksize = self.config.getint("Environment", "Decksize")
        seed = self.config.getint("Utils", "Seed") if seed is None else seed
        self.n_games = int(n_games)
        self.batched = BatchedLegacyEnv(self.n_games, seed=seed, device=device)
        self._state_shape = np.zeros((1, 3))
        self._observation_state = 3
        self._action_space = 3
        self._info = ""
        self._actions = [[3, 3] for _ in range(self.n_games)]  # env.py:66: stored "action" starts as 3

    @property
    def dim_shape(self):
        return self._state_shape.shape

    @property
    def observation_space(self):
        return self._observation_state

    @property
    def action_space(self):
        return self._action_space

    def reset(self):
        self._actions = [[3, 3] for _ in range(self.n_games)]
        self.batched.reset()

    def load_hand(self, c0, c1):
        self._actions = [[3, 3] for _ in range(self.n_games)]
        cards = np.broadcast_to(np.asarray([c0, c1], np.int8), (self.n_games, 2)).copy()
        self.batched.set_hands(cards)

    def init_state(self, player_index):
        e = self.batched.export()
        card = e["c%d" % player_index].cpu().numpy()
        st = np.stack([card, np.full_like(card, -1), e["st_pot%d" % player_index].cpu().numpy()], 1).astype(np.int64)
        return st[0:1] if self.n_games == 1 else st

    def step(self, action, player_index):
        a = np.asarray(action, dtype=np.float64).reshape(self.n_games, -1)
        # a scalar action argmaxes to 0 = fold, as in the reference (env.py:107)
        code = np.argmax(a, axis=1).astype(np.int8)
        self.batched.step(code, int(player_index))
        for g in range(self.n_games):
            self._actions[g][player_index] = action if self.n_games == 1 else a[g]

    def get_new_state(self, player_index):
        out = self.batched.get_new_state(int(player_index)).cpu().numpy()
        if self.n_games == 1:
            card, pub, pot, reward, terminal = (int(x) for x in out[0])
            tup = np.empty(5, dtype=object)
            tup[0] = np.array([[card, pub, pot]])
            tup[1] = self._actions[0][player_index]
            tup[2], tup[3], tup[4] = reward, terminal, self._info
            return tup
        return out
