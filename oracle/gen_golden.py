"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (oracle/ref_runner.py).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/

The reference has no tests and no golden vectors, so the fixtures
are exhaustive enumerations of its tiny game trees plus call-order fuzzing:

  nfsp_exhaustive.npz   every (rank-deal, dealer, policy pair, raw decision sequence) hand of
                        leduc/newenv.py driven by the reference's own main.train loop and
                        Agent.play (extracted with ast), decisions scripted through stub models
                        ("F","C","R" score vectors, incl. argmax ties, and "Z" = all-zero vector).
                        Logged: every env.step with post-step env internals, every RL / SL
                        memory add (by value at add time), agent counters.
  legacy_exhaustive.npz every (card pair, joint action sequence) hand of leduc/env.py under the
                        README.md:15-38 driver.
  fuzz_calls.npz        random call sequences (arbitrary players / repeated calls) on both envs:
                        pins the single-game drop-in API outside the canonical drivers.
  buffers.npz           ReplayBuffer / ReservoirBuffer op sequences with scripted `random`.
  kat.npz               boltzmann() / temperature KATs (leduc/test.py:57-77, agent.py:50,158-166).
"""
from __future__ import annotations

import ast
import itertools
import os
import random as pyrandom
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_runner as rr  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

KIND_VECS = {  # dyadic values: exact in fp32
    "F": [np.array([0.75, 0.125, 0.125]), np.array([0.5, 0.5, 0.0])],
    "C": [np.array([0.125, 0.75, 0.125]), np.array([0.0, 0.5, 0.5])],
    "R": [np.array([0.125, 0.125, 0.75]), np.array([0.0, 0.25, 1.5])],
    "Z": [np.array([0.0, 0.0, 0.0]), np.array([0.0, 0.0, 0.0])],
}
KIND_ID = {"F": 0, "C": 1, "R": 2, "Z": 3}


def mask30(v) -> int:
    v = np.asarray(v).reshape(-1)
    assert v.shape[0] == 30
    m = 0
    for i in range(30):
        if v[i] != 0:
            assert v[i] == 1
            m |= 1 << i
    return m


class NeedMore(Exception):
    pass


# ------------------------------------------------------------------------------------------
# NFSP exhaustive enumeration through the reference's train() + Agent.play()
# ------------------------------------------------------------------------------------------
class HandScript:
    def __init__(self, dealer, cards, policies, kinds, variant):
        self.dealer, self.cards, self.policies, self.kinds, self.variant = dealer, cards, policies, kinds, variant
        self.pos = 0  # next decision
        self.eps_flags = []

    def eps_for(self, i):
        return (self.variant + i) % 3 == 0

    def next_vec(self):
        if self.pos >= len(self.kinds):
            raise NeedMore()
        k = self.kinds[self.pos]
        v = KIND_VECS[k][(self.variant + self.pos) % 2]
        self.pos += 1
        return v.copy()


class TrainRandom:
    """`random` as seen by main.train: initial dealer + the two per-hand policy draws."""

    def __init__(self, script: HandScript):
        self.script = script
        self.calls = 0

    def randint(self, a, b):
        return 1 - self.script.dealer  # main.py:24,28-31 flips before the first hand

    def random(self):
        # main.py:38-45: first draw -> policy[dealer], second -> policy[lhand]; 'a' iff > eta
        who = self.script.dealer if self.calls == 0 else 1 - self.script.dealer
        self.calls += 1
        return 0.05 if self.script.policies[who] == 1 else 0.5


class AgentRandom:
    def __init__(self, script: HandScript):
        self.script = script

    def random(self):  # agent.py:125
        eps = self.script.eps_for(self.script.pos)
        self.script.eps_flags.append((self.script.pos, eps))
        return 0.01 if eps else 0.9


class NpShim:
    def __init__(self, script: HandScript, log):
        self.script, self.log = script, log

        class _R:
            @staticmethod
            def rand(*shape):  # agent.py:128
                assert shape == (1, 1, 3)
                log["src"].append(2)
                return script.next_vec().reshape(1, 1, 3)

        self.random = _R()

    def __getattr__(self, name):
        return getattr(np, name)


class ScriptModel:
    def __init__(self, script, log, src_id):
        self.script, self.log, self.src_id = script, log, src_id

    def predict(self, x):
        assert x.shape == (1, 1, 30)
        v = self.script.next_vec()
        self.log["src"].append(self.src_id)
        self.log["x_in"].append(mask30(x))
        return v.reshape(1, 1, 3)


def run_nfsp_hand(mods, script: HandScript):
    """Runs ONE hand through the reference's train(); returns the log or raises NeedMore."""
    log = {"src": [], "x_in": [], "steps": [], "rl": [], "sl": [], "anomalies": 0}
    newenv = mods["newenv"]

    def counting_print(*a, **k):
        log["anomalies"] += 1

    newenv.print = counting_print
    env = newenv.Env()
    orig_step = env.step

    def logged_step(action, p_index):
        before = np.concatenate((env.history.flatten(), env.specific_cards[p_index].flatten()))
        orig_step(action, p_index)
        log["steps"].append(dict(
            player=p_index, vec=np.asarray(action, dtype=np.float64).reshape(3).copy(), obs_before=mask30(before),
            terminated=int(bool(env.terminated)), round=int(env.round),
            bets=np.array(env.overall_raises, dtype=np.float64).copy(),
            reward=np.array(env.reward, dtype=np.float64).copy(),
            obs_after=[mask30(np.concatenate((env.history.flatten(), env.specific_cards[q].flatten())))
                       for q in (0, 1)],
            snap=[mask30(env.s[q]) for q in (0, 1)],
            round_raises=int(env.round_raises)))

    env.step = logged_step
    Agent = rr.agent_class(random_module=AgentRandom(script), np_module=NpShim(script, log))
    agents = []
    for p in (0, 1):
        rl = mods["replay_buffer"].ReplayBuffer(1000, 1234)
        sl = mods["ReservoirBuffer"].ReservoirBuffer(1000, 1234)
        orig_rl_add, orig_sl_add = rl.add, sl.add

        def rl_add(s, a, r, s2, t, _p=p, _orig=orig_rl_add):
            log["rl"].append(dict(player=_p, s=mask30(s), a=np.asarray(a, dtype=np.float64).reshape(3).copy(),
                                  r=float(r), s2=mask30(s2), t=int(bool(t))))
            _orig(s, a, r, s2, t)

        def sl_add(s, a, _p=p, _orig=orig_sl_add):
            log["sl"].append(dict(player=_p, s=mask30(s), a=np.asarray(a, dtype=np.float64).reshape(3).copy()))
            _orig(s, a)

        rl.add, sl.add = rl_add, sl_add
        agents.append(Agent("Player%d" % p, env, ScriptModel(script, log, 0), ScriptModel(script, log, 1), rl, sl,
                            epsilon=0.06, temp=1.0))
    train = rr.train_function(episodes=1, eta=0.1, random_module=TrainRandom(script))
    with rr.ScriptedShuffle(lambda: list(script.cards)):
        train(env, agents[0], agents[1])
    if script.pos != len(script.kinds):
        return None  # hand ended before consuming the whole script: not a canonical sequence
    log["agent"] = dict(actions=np.array([a.actions for a in agents]), played=[a.played for a in agents],
                        reward=[a.reward for a in agents], game_step=[a.game_step for a in agents],
                        updates=[a.updates for a in agents])
    log["final_env"] = env
    return log


def enumerate_sequences(mods):
    """All complete raw decision sequences over the alphabet F,C,R,Z (deal-independent)."""
    done, frontier = [], [()]
    while frontier:
        nxt = []
        for prefix in frontier:
            for k in "FCRZ":
                seq = prefix + (k,)
                try:
                    r = run_nfsp_hand(mods, HandScript(0, (0, 1, 2), (0, 0), seq, 0))
                    assert r is not None
                    done.append(seq)
                except NeedMore:
                    nxt.append(seq)
        frontier = nxt
    return done


def gen_nfsp(mods):
    seqs = enumerate_sequences(mods)
    n_fcr = sum(1 for s in seqs if "Z" not in s)
    deals = [(a, b, c) for a in range(3) for b in range(3) for c in range(3) if not (a == b == c)]
    assert len(deals) == 24
    rows = []
    variant = 0
    for deal in deals:
        for dealer in (0, 1):
            for pol in itertools.product((0, 1), repeat=2):
                for seq in seqs:
                    sc = HandScript(dealer, deal, pol, seq, variant)
                    log = run_nfsp_hand(mods, sc)
                    assert log is not None
                    rows.append((sc, log))
                    variant += 1
    T = len(rows)
    D, RL, SL = 6, 12, 6
    g = dict(
        dealer=np.zeros(T, np.int8), cards=np.zeros((T, 3), np.int8), policy=np.zeros((T, 2), np.int8),
        n_dec=np.zeros(T, np.int8), kind=np.full((T, D), -1, np.int8), src=np.full((T, D), -1, np.int8),
        vec=np.zeros((T, D, 3), np.float32), player=np.full((T, D), -1, np.int8),
        obs_before=np.zeros((T, D), np.uint32), terminated=np.zeros((T, D), np.int8),
        round=np.zeros((T, D), np.int8), round_raises=np.zeros((T, D), np.int8),
        bets=np.zeros((T, D, 2), np.float32), reward=np.zeros((T, D, 2), np.float32),
        obs_after=np.zeros((T, D, 2), np.uint32), snap=np.zeros((T, D, 2), np.uint32),
        n_rl=np.zeros(T, np.int8), rl_player=np.full((T, RL), -1, np.int8), rl_s=np.zeros((T, RL), np.uint32),
        rl_a=np.zeros((T, RL, 3), np.float32), rl_r=np.zeros((T, RL), np.float32),
        rl_s2=np.zeros((T, RL), np.uint32), rl_t=np.zeros((T, RL), np.int8),
        n_sl=np.zeros(T, np.int8), sl_player=np.full((T, SL), -1, np.int8), sl_s=np.zeros((T, SL), np.uint32),
        sl_a=np.zeros((T, SL, 3), np.float32),
        ag_actions=np.zeros((T, 2, 3), np.int32), ag_played=np.zeros((T, 2), np.int32),
        ag_reward=np.zeros((T, 2), np.float32), ag_game_step=np.zeros((T, 2), np.int32),
        anomalies=np.zeros(T, np.int32))
    for i, (sc, log) in enumerate(rows):
        g["dealer"][i] = sc.dealer
        g["cards"][i] = sc.cards
        g["policy"][i] = sc.policies
        n = len(sc.kinds)
        assert n == len(log["steps"]) == len(log["src"]) <= D
        g["n_dec"][i] = n
        for d in range(n):
            st = log["steps"][d]
            g["kind"][i, d] = KIND_ID[sc.kinds[d]]
            g["src"][i, d] = log["src"][d]
            g["vec"][i, d] = st["vec"]
            assert np.array_equal(g["vec"][i, d].astype(np.float64), st["vec"])
            g["player"][i, d] = st["player"]
            g["obs_before"][i, d] = st["obs_before"]
            g["terminated"][i, d] = st["terminated"]
            g["round"][i, d] = st["round"]
            g["round_raises"][i, d] = st["round_raises"]
            g["bets"][i, d] = st["bets"]
            g["reward"][i, d] = st["reward"]
            g["obs_after"][i, d] = st["obs_after"]
            g["snap"][i, d] = st["snap"]
        assert len(log["rl"]) <= RL and len(log["sl"]) <= SL
        g["n_rl"][i] = len(log["rl"])
        for j, r in enumerate(log["rl"]):
            g["rl_player"][i, j] = r["player"]
            g["rl_s"][i, j] = r["s"]
            g["rl_a"][i, j] = r["a"]
            g["rl_r"][i, j] = r["r"]
            g["rl_s2"][i, j] = r["s2"]
            g["rl_t"][i, j] = r["t"]
        g["n_sl"][i] = len(log["sl"])
        for j, r in enumerate(log["sl"]):
            g["sl_player"][i, j] = r["player"]
            g["sl_s"][i, j] = r["s"]
            g["sl_a"][i, j] = r["a"]
        ag = log["agent"]
        g["ag_actions"][i] = ag["actions"]
        g["ag_played"][i] = ag["played"]
        g["ag_reward"][i] = ag["reward"]
        g["ag_game_step"][i] = ag["game_step"]
        g["anomalies"][i] = log["anomalies"]
    np.savez_compressed(os.path.join(OUT, "nfsp_exhaustive.npz"), **g)
    print("nfsp_exhaustive: %d traces (%d raw F/C/R sequences, %d with Z), anomalies=%d" %
          (T, n_fcr, len(seqs), int(g["anomalies"].sum())))


# ------------------------------------------------------------------------------------------
# legacy env exhaustive enumeration under the README driver
# ------------------------------------------------------------------------------------------
def onehot(a):
    v = np.zeros(3)
    v[a] = 1
    return v


def run_legacy_hand(mods, cards, joint_actions):
    """README.md:15-38 driver. Returns per-iteration logs, or None if the script is too short, or
    'long' if the hand ended before the script was consumed."""
    env = mods["env"].Env()
    with rr.ScriptedShuffle(lambda: [cards[0], cards[1], 0, 0, 0, 0]):
        env.reset()
    init = [env.init_state(p).copy() for p in (0, 1)]
    its = []
    for i, (a0, a1) in enumerate(joint_actions):
        env.step(onehot(a0), 0)
        env.step(onehot(a1), 1)
        left_after_step = list(env._left_choices)
        pot_after_step = list(env._pot)
        term_after_step = [int(env._specific_state[p][3]) for p in (0, 1)]
        outs = []
        for p in (0, 1):
            r = env.get_new_state(p)
            outs.append((int(r[0][0][0]), int(r[0][0][1]), int(r[0][0][2]), int(r[2]), int(r[3])))
        its.append(dict(left=left_after_step, pot=pot_after_step, term_step=term_after_step, out=outs))
        if outs[0][4] or outs[1][4]:
            return ("done" if i == len(joint_actions) - 1 else "long"), init, its
    return "more", init, its


def gen_legacy(mods):
    pairs = [(a, b) for a in range(3) for b in range(3)]
    joint = [(a, b) for a in range(3) for b in range(3)]
    # enumerate sequences once (card-independent termination) then run for every card pair
    done, frontier = [], [()]
    while frontier:
        nxt = []
        for prefix in frontier:
            for ja in joint:
                seq = prefix + (ja,)
                status, _, _ = run_legacy_hand(mods, (0, 1), seq)
                if status == "done":
                    done.append(seq)
                elif status == "more":
                    nxt.append(seq)
                else:
                    raise AssertionError(status)
        frontier = nxt
    I = max(len(s) for s in done)
    T = len(done) * len(pairs)
    g = dict(cards=np.zeros((T, 2), np.int8), n_it=np.zeros(T, np.int8), actions=np.full((T, I, 2), -1, np.int8),
             left=np.zeros((T, I, 2), np.int8), pot=np.zeros((T, I, 2), np.int8),
             term_step=np.zeros((T, I, 2), np.int8), out=np.zeros((T, I, 2, 5), np.int32),
             init=np.zeros((T, 2, 3), np.int32))
    i = 0
    for cards in pairs:
        for seq in done:
            status, init, its = run_legacy_hand(mods, cards, seq)
            assert status == "done"
            g["cards"][i] = cards
            g["n_it"][i] = len(seq)
            g["init"][i] = np.array(init).reshape(2, 3)
            for k, it in enumerate(its):
                g["actions"][i, k] = seq[k]
                g["left"][i, k] = it["left"]
                g["pot"][i, k] = it["pot"]
                g["term_step"][i, k] = it["term_step"]
                g["out"][i, k] = it["out"]
            i += 1
    np.savez_compressed(os.path.join(OUT, "legacy_exhaustive.npz"), **g)
    zs = 0
    for t in range(T):
        k = g["n_it"][t] - 1
        zs += int(g["out"][t, k, 0, 3] + g["out"][t, k, 1, 3] == 0)
    print("legacy_exhaustive: %d traces, max %d iterations, reward range [%d,%d], zero-sum %.1f%%" %
          (T, I, g["out"][..., 3].min(), g["out"][..., 3].max(), 100.0 * zs / T))


# ------------------------------------------------------------------------------------------
# call-order fuzz: arbitrary players, repeated calls, steps after terminal
# ------------------------------------------------------------------------------------------
def gen_fuzz(mods, n_hands=1500, n_calls=14, seed=20181018):
    rng = pyrandom.Random(seed)
    newenv = mods["newenv"]
    newenv.print = lambda *a, **k: None
    # ---- newenv: ops 0 = step(vec,p), 1 = get_state(p)
    N = n_hands
    nf = dict(dealer=np.zeros(N, np.int8), cards=np.zeros((N, 3), np.int8), op=np.zeros((N, n_calls), np.int8),
              player=np.zeros((N, n_calls), np.int8), vec=np.zeros((N, n_calls, 3), np.float32),
              terminated=np.zeros((N, n_calls), np.int8), round=np.zeros((N, n_calls), np.int8),
              bets=np.zeros((N, n_calls, 2), np.float32), reward=np.zeros((N, n_calls, 2), np.float32),
              obs=np.zeros((N, n_calls, 2), np.uint32), snap=np.zeros((N, n_calls, 2), np.uint32),
              gs_r=np.zeros((N, n_calls), np.float32), gs_a=np.zeros((N, n_calls, 3), np.float32))
    env = newenv.Env()
    deals = [(a, b, c) for a in range(3) for b in range(3) for c in range(3) if not (a == b == c)]
    for i in range(N):
        deal = rng.choice(deals)
        dealer = rng.randrange(2)
        with rr.ScriptedShuffle(lambda: list(deal)):
            env.reset(dealer)
        nf["dealer"][i], nf["cards"][i] = dealer, deal
        for c in range(n_calls):
            op = 0 if rng.random() < 0.7 else 1
            p = rng.randrange(2)
            nf["op"][i, c], nf["player"][i, c] = op, p
            if op == 0:
                kind = rng.choice("FCCRRRZ" if rng.random() < 0.5 else "CCRR")
                v = KIND_VECS[kind][rng.randrange(2)]
                nf["vec"][i, c] = v
                env.step(v.copy(), p)
            else:
                s, a, r, s2, t = env.get_state(p)
                nf["gs_r"][i, c] = r
                nf["gs_a"][i, c] = np.asarray(a).reshape(3)
            nf["terminated"][i, c] = int(bool(env.terminated))
            nf["round"][i, c] = env.round
            nf["bets"][i, c] = env.overall_raises
            nf["reward"][i, c] = env.reward
            nf["obs"][i, c] = [mask30(np.concatenate((env.history.flatten(), env.specific_cards[q].flatten())))
                               for q in (0, 1)]
            nf["snap"][i, c] = [mask30(env.s[q]) for q in (0, 1)]
    np.savez_compressed(os.path.join(OUT, "fuzz_nfsp_calls.npz"), **nf)
    # ---- legacy env: ops 0 = step(onehot,p), 1 = get_new_state(p)
    lf = dict(cards=np.zeros((N, 2), np.int8), op=np.zeros((N, n_calls), np.int8),
              player=np.zeros((N, n_calls), np.int8), action=np.zeros((N, n_calls), np.int8),
              left=np.zeros((N, n_calls, 2), np.int8), pot=np.zeros((N, n_calls, 2), np.int8),
              st=np.zeros((N, n_calls, 2, 5), np.int32))
    lenv = mods["env"].Env()
    for i in range(N):
        cards = (rng.randrange(3), rng.randrange(3))
        with rr.ScriptedShuffle(lambda: [cards[0], cards[1], 0, 0, 0, 0]):
            lenv.reset()
        lf["cards"][i] = cards
        for c in range(n_calls):
            op = 0 if rng.random() < 0.6 else 1
            p = rng.randrange(2)
            a = rng.choice((0, 1, 1, 2, 2))
            lf["op"][i, c], lf["player"][i, c], lf["action"][i, c] = op, p, a
            if op == 0:
                lenv.step(onehot(a), p)
            else:
                lenv.get_new_state(p)
            lf["left"][i, c] = lenv._left_choices
            lf["pot"][i, c] = lenv._pot
            for q in (0, 1):
                st = lenv._specific_state[q]
                lf["st"][i, c, q] = (int(st[0][0][0]), int(st[0][0][1]), int(st[0][0][2]), int(st[2]), int(st[3]))
    np.savez_compressed(os.path.join(OUT, "fuzz_legacy_calls.npz"), **lf)
    print("fuzz: %d hands x %d calls for each env" % (N, n_calls))


# ------------------------------------------------------------------------------------------
# buffers
# ------------------------------------------------------------------------------------------
class ScriptedBufRandom:
    def __init__(self, seed):
        self.rng = pyrandom.Random(seed)
        self.randrange_log, self.sample_log = [], []

    def seed(self, s):
        pass

    def randrange(self, a, b):
        j = self.rng.randrange(a, b)
        self.randrange_log.append(j)
        return j

    def sample(self, population, k):
        idx = self.rng.sample(range(len(population)), k)
        self.sample_log.append(list(idx))
        return [population[i] for i in idx]


def gen_do_action(mods, n_hands=1200, seed=20261018):
    """newenv.Env.do_action (newenv.py:131-178) called on its own: a few step() calls bring the hand somewhere (possibly
    into round 1), then up to three bare do_action calls by arbitrary players; after every call the return value and
    everything the method touches (history, round_raises, overall_raises, last_action)."""
    rng = pyrandom.Random(seed)
    newenv = mods["newenv"]
    newenv.print = lambda *a, **k: None
    N, C = n_hands, 6
    g = dict(dealer=np.zeros(N, np.int8), cards=np.zeros((N, 3), np.int8), op=np.full((N, C), -1, np.int8),
             player=np.zeros((N, C), np.int8), vec=np.zeros((N, C, 3), np.float32), ret=np.zeros((N, C), np.int8),
             hist=np.zeros((N, C), np.uint32), round=np.zeros((N, C), np.int8), round_raises=np.zeros((N, C), np.int8),
             bets=np.zeros((N, C, 2), np.float32), terminated=np.zeros((N, C), np.int8),
             last_action=np.zeros((N, C, 2, 3), np.float32))
    env = newenv.Env()
    deals = [(a, b, c) for a in range(3) for b in range(3) for c in range(3) if not (a == b == c)]
    for i in range(N):
        deal, dealer = rng.choice(deals), rng.randrange(2)
        with rr.ScriptedShuffle(lambda: list(deal)):
            env.reset(dealer)
        g["dealer"][i], g["cards"][i] = dealer, deal
        n_steps = rng.randrange(0, 4)
        folded = False
        for c in range(C):
            # a bare fold ends the hand for the caller, as it does for step() (the reference appends 'Fold' to actions_done,
            # a state no later call is meant to see); a fourth action in a round is an IndexError in the reference
            if env.terminated or env.round_raises > 2 or folded:
                break
            p = rng.randrange(2)
            v = KIND_VECS[rng.choice("FCCRRRZ")][rng.randrange(2)]
            if c < n_steps:  # ops: 0 = step(vec, p), 1 = do_action(vec, p)
                g["op"][i, c] = 0
                env.step(v.copy(), p)
            else:
                g["op"][i, c] = 1
                g["ret"][i, c] = int(bool(env.do_action(v.copy(), p)))
                folded = bool(g["ret"][i, c])
            g["player"][i, c], g["vec"][i, c] = p, v
            g["hist"][i, c] = mask30(np.concatenate((env.history.flatten(), np.zeros(6)))) & 0xFFFFFF
            g["round"][i, c], g["round_raises"][i, c] = env.round, env.round_raises
            g["bets"][i, c], g["terminated"][i, c] = env.overall_raises, int(bool(env.terminated))
            g["last_action"][i, c] = np.asarray(env.last_action, dtype=np.float64).reshape(2, 3)
    np.savez_compressed(os.path.join(OUT, "do_action.npz"), **g)
    print("do_action.npz: %d hands, %d bare do_action calls, %d folds" % (N, int((g["op"] == 1).sum()), int(g["ret"].sum())))


def gen_buffers(mods, seed=7):
    rng = np.random.RandomState(seed)
    cap, n_add = 50, 173
    states = (rng.rand(n_add, 30) < 0.3).astype(np.float64)
    states2 = (rng.rand(n_add, 30) < 0.3).astype(np.float64)
    acts = np.eye(3)[rng.randint(0, 3, n_add)]
    rews = rng.randint(-10, 11, n_add) * 0.5
    terms = rng.rand(n_add) < 0.3
    # ring
    rb_mod = mods["replay_buffer"]
    sr = ScriptedBufRandom(11)
    rb_mod.random = sr
    rb = rb_mod.ReplayBuffer(cap, 1234)
    sizes = []
    for i in range(n_add):
        rb.add(states[i].reshape(1, 1, 30), acts[i].reshape(1, 1, 3), rews[i], states2[i].reshape(1, 1, 30), terms[i])
        sizes.append(rb.size())
    ring_s = np.array([mask30(e[0]) for e in rb.buffer], np.uint32)  # deque order: oldest first
    ring_s2 = np.array([mask30(e[3]) for e in rb.buffer], np.uint32)
    ring_a = np.array([int(np.argmax(e[1])) for e in rb.buffer], np.int8)
    ring_r = np.array([e[2] for e in rb.buffer], np.float32)
    ring_t = np.array([int(e[4]) for e in rb.buffer], np.int8)
    sb = rb.sample_batch(16)
    small = rb_mod.ReplayBuffer(cap, 1234)
    for i in range(5):
        small.add(states[i], acts[i], rews[i], states2[i], terms[i])
    sb_small = small.sample_batch(16)  # count < batch -> returns `count` rows (replay_buffer.py:48-49)
    # reservoir (reference mode)
    rs_mod = mods["ReservoirBuffer"]
    sr2 = ScriptedBufRandom(13)
    rs_mod.random = sr2
    rs = rs_mod.ReservoirBuffer(cap, 1234)
    avecs = rng.rand(n_add, 3).astype(np.float32).astype(np.float64)
    rsizes = []
    for i in range(n_add):
        rs.add(states[i].reshape(1, 1, 30), avecs[i].reshape(1, 1, 3))
        rsizes.append(rs.size())
    res_s = np.array([mask30(e[0]) for e in rs.buffer], np.uint32)
    res_a = np.array([e[1].reshape(3) for e in rs.buffer], np.float32)
    ssb = rs.sample_batch(16)
    np.savez_compressed(
        os.path.join(OUT, "buffers.npz"), cap=cap, n_add=n_add,
        in_s=np.array([mask30(s) for s in states], np.uint32), in_s2=np.array([mask30(s) for s in states2], np.uint32),
        in_a=np.argmax(acts, 1).astype(np.int8), in_r=rews.astype(np.float32), in_t=terms.astype(np.int8),
        in_avec=avecs.astype(np.float32),
        ring_sizes=np.array(sizes), ring_s=ring_s, ring_s2=ring_s2, ring_a=ring_a, ring_r=ring_r, ring_t=ring_t,
        ring_sample_idx=np.array(sr.sample_log[0]), ring_sample_s=np.array([mask30(x) for x in sb[0]], np.uint32),
        ring_sample_shapes=np.array([sb[0].shape + (0,) * (3 - sb[0].ndim), sb[1].shape, sb[3].shape]),
        ring_sample_r=sb[2].astype(np.float32), ring_sample_t=sb[4].astype(np.int8),
        ring_small_rows=sb_small[0].shape[0],
        res_sizes=np.array(rsizes), res_j=np.array(sr2.randrange_log), res_s=res_s, res_a=res_a,
        res_sample_idx=np.array(sr2.sample_log[0]), res_sample_s=np.array([mask30(x) for x in ssb[0]], np.uint32),
        res_sample_shapes=np.array([ssb[0].shape, ssb[1].shape]))
    print("buffers: ring %d adds into %d; reservoir j-draws=%d" % (n_add, cap, len(sr2.randrange_log)))


def gen_kat(mods):
    src = open(os.path.join(rr.REF_ROOT, "leduc", "test.py")).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "boltzmann"]
    ns = {"np": np}
    exec(compile(ast.Module(body=fn, type_ignores=[]), "test.py", "exec"), ns)
    q = np.array([0.0, -0.45, 0.23])
    b = ns["boltzmann"](q, 0.99)  # leduc/test.py:71-75
    temp = (1 + 0.02 * np.sqrt(100000)) ** (-1)  # leduc/test.py:77
    Agent = rr.agent_class()
    ag = Agent.__new__(Agent)
    qs = np.array([[0.0, 0.0, 0.0], [1.25, 0.5, 0.0], [0.3, 0.3, 0.1], [2.0, 0.0, 4.5]])
    temps = np.array([1.0, 0.5, (1 + 0.02 * np.sqrt(7)) ** (-1), temp])
    outs = np.zeros((len(temps), len(qs), 3))
    for i, t in enumerate(temps):
        ag.temp = t
        for j, qq in enumerate(qs):
            outs[i, j] = ag.boltzmann(qq.reshape(1, 1, 3)).reshape(3)  # agent.py:158-166
    np.savez(os.path.join(OUT, "kat.npz"), test_q=q, test_t=0.99, test_boltzmann=b, test_temp=temp,
             agent_q=qs, agent_temps=temps, agent_boltzmann=outs)
    print("kat: boltzmann", b, "temp", temp)


def main():
    os.makedirs(OUT, exist_ok=True)
    os.chdir(rr.REF_ROOT)
    mods = rr.load()
    gen_kat(mods)
    gen_do_action(mods, n_hands=2400)
    gen_buffers(mods)
    gen_legacy(mods)
    gen_fuzz(mods)
    gen_nfsp(mods)


if __name__ == "__main__":
    main()
