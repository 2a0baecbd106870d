"""Drop-in for agent/agent.py `Agent`: the acting half (play / act_best_response / boltzmann /
remember_*) over the CUDA forward and the GPU memories.  Single-game semantics follow
agent.py:118-166 decision by decision; the batched hot path is `nfsp_b200.SelfPlay`.

The reference's README names `act` and `remember_opponent_behaviour` (README.md:75); the code has
`play` and `remember_for_rl`.  Both spellings are exported.
"""
import math
import random

import numpy as np
import torch

from .. import _lib
from ..batched import BatchedNfspEnv, _as, _ptr, _stream, check, glorot_nets, lib, obs_to_mask
from ..config import load_config
from ..utils.replay_buffer import ReplayBuffer
from ..utils.ReservoirBuffer import ReservoirBuffer


class _Model:
    """`keras.Model.predict` stand-in bound to one of the agent's nets (agent.py:101-116)."""

    def __init__(self, agent, slot):
        self.agent, self.slot = agent, slot

    def predict(self, x):
        x = np.asarray(x, dtype=np.float64).reshape(-1, 30)
        out = self.agent._forward(x, self.slot)
        return out.reshape(x.shape[0], 1, 3) if x.shape[0] > 1 else out.reshape(1, 1, 3)

    def get_weights(self):
        w = self.agent.weights[self.slot].cpu().numpy()
        return [w[:1920].reshape(30, 64).copy(), w[1920:1984].copy(), w[1984:2176].reshape(64, 3).copy(),
                w[2176:].copy()]

    def set_weights(self, ws):
        flat = np.concatenate([np.asarray(x, np.float32).reshape(-1) for x in ws])
        self.agent.weights[self.slot] = torch.from_numpy(flat).to(self.agent.device)
        self.agent._upload()


class Agent:
    AVG, BR, TARGET = 0, 1, 2

    def __init__(self, sess, state_dim, action_dim, name, env, config_path="./config.ini", device=None, seed=None):
        self.sess, self.s_dim, self.a_dim, self.name, self.env = sess, state_dim, action_dim, name, env
        cfg = self.config = load_config(config_path)
        self.exploitability = 0
        self.iteration = 0
        self.minibatch_size = cfg.getint("Agent", "MiniBatchSize")
        self.n_hidden = cfg.getint("Agent", "HiddenLayer")
        self.lr_br = cfg.getfloat("Agent", "LearningRateBR")
        self.lr_ar = cfg.getfloat("Agent", "LearningRateAR")
        self.epsilon = cfg.getfloat("Agent", "Epsilon")
        self.epsilon_min = cfg.getfloat("Agent", "EpsilonMin")
        self.gamma = cfg.getfloat("Agent", "Gamma")
        self.omega = cfg.getfloat("Agent", "Omega")
        self.target_model_update_rate = cfg.getint("Agent", "TargetModelUpdateRate")
        self.temp = (1 + 0.02 * np.sqrt(self.iteration)) ** (-1)
        self.eta = cfg.getfloat("Agent", "Eta")
        self.target_br_model_update_count = 0
        buf, bseed = cfg.getint("Utils", "Buffersize"), cfg.getint("Utils", "Seed")
        self._rl_memory = ReplayBuffer(buf, bseed, device)        # agent.py:59-60
        self._sl_memory = ReservoirBuffer(buf, bseed, device)     # agent.py:63-64
        self.device = self._rl_memory.memory.device
        # three nets: average, best response, target (a copy of BR) -- agent.py:66-72
        seed = bseed if seed is None else seed
        g = glorot_nets(seed + sum(map(ord, name)), self.device)
        self.weights = [g[0].clone(), g[1].clone(), g[1].clone()]
        self._fwd = BatchedNfspEnv(1, seed=seed, device=self.device)  # owns the packed weight image
        self._upload()
        self.avg_strategy_model = _Model(self, self.AVG)
        self.best_response_model = _Model(self, self.BR)
        self.target_br_model = _Model(self, self.TARGET)
        self.actions = np.zeros(3)
        self.played = 0
        self.reward = 0
        self.game_step = 0

    # ---- device plumbing
    def _upload(self):
        # forward slots: 0 = average (softmax head), 1 = BR (relu head), 3 = target BR (relu head)
        w = torch.stack([self.weights[0], self.weights[1], self.weights[0], self.weights[2]]).contiguous()
        self._wdev = w
        check(lib().nfsp_act_set_weights(self._fwd._h, _ptr(w), _stream(self.device)))

    def _forward(self, x, slot):
        m = obs_to_mask(torch.from_numpy(x)).to(self.device)
        net = torch.full((m.numel(),), {0: 0, 1: 1, 2: 3}[slot], dtype=torch.int8, device=self.device)
        out = torch.empty((m.numel(), 3), dtype=torch.float32, device=self.device)
        check(lib().nfsp_act_forward(self._fwd._h, _ptr(m), _ptr(net), m.numel(), _ptr(out), _stream(self.device)))
        return out.cpu().numpy().astype(np.float64)

    # ---- agent.py:118-128
    def remember_best_response(self, state, action):
        self._sl_memory.add(state, action)

    def remember_for_rl(self, state, action, reward, nextstate, terminal):
        self._rl_memory.add(state, action, reward, nextstate, terminal)

    remember_opponent_behaviour = remember_for_rl  # README.md:75 spelling

    def act_best_response(self, state):
        if random.random() > self.epsilon:
            return self.best_response_model.predict(state)
        return np.random.rand(1, 1, 3)

    # ---- agent.py:130-156
    def play(self, policy, index, s2=None):
        if s2 is None:
            s, a, r, s2, t = self.env.get_state(index)
            self.reward += r
            if np.average(a) != 0:
                self.remember_for_rl(s, a, r, s2, t)
                self.game_step += 1
            if t:
                return t
        else:
            t = False
        if policy == "a":
            a = self.avg_strategy_model.predict(np.reshape(s2, (1, 1, 30)))
            self.env.step(a, index)
            self.played += 1
        else:
            a_t = self.act_best_response(np.reshape(s2, (1, 1, 30)))
            a = self.boltzmann(a_t)
            self.env.step(a_t, index)
            self.remember_best_response(s2, a_t)
            self.played += 1
        if self.game_step % 128 == 0:
            self.update_strategy()
        self.actions[np.argmax(a)] += 1
        return t

    act = play  # BASELINE.json north_star spelling

    def boltzmann(self, actions):  # agent.py:158-166
        q = np.asarray(actions, dtype=np.float64).reshape(3)
        e = np.exp(q / self.temp)
        return (e / e.sum()).reshape(1, 1, 3)

    # ---- agent.py:168-190 (evaluation hooks; their caller is commented out in main.py)
    def update_strategy(self):
        self.update_avg_response_network()
        self.update_best_response_network()

    def sampled_actions(self):
        print("{} played {} times: Folds: {}, Calls: {}, Raises: {} - Reward: {}".format(
            self.name, self.played, self.actions[0], self.actions[1], self.actions[2], self.reward))
        self.actions = np.zeros(3)
        self.played = 0

    def average_payoff_br(self):
        return np.average(self.exploitability)

    # ---- learner half (agent.py:209-273): SURVEY section 8 f-1, "next" -- see learner.py
    def update_best_response_network(self):
        if self._rl_memory.size() > self.minibatch_size:
            from ..learner import update_best_response

            update_best_response(self)

    def update_avg_response_network(self):
        if self._sl_memory.size() > self.minibatch_size:
            from ..learner import update_average_policy

            update_average_policy(self)

    def update_br_target_network(self):  # agent.py:266-273
        if self.target_br_model_update_count % self.target_model_update_rate == 0:
            self.weights[self.TARGET] = self.weights[self.BR].clone()
            self._upload()
        self.target_br_model_update_count += 1
