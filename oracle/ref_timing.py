"""BASELINE.json configs[0]: the REFERENCE'S OWN Python timed on this machine's host cores.

TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference legs).  Nothing here is a
restatement: the code that runs is the reference's, loaded by oracle/ref_runner.py (five one-token Py2 -> Py3
edits applied in memory) from NFSP_REFERENCE_ROOT -- /root/reference in the build container, or the verbatim,
git-ignored copy under baseline/_ref/ that `materialise()` writes so that it travels to the GPU box.

Two drivers:
  legacy   leduc.env.Env under the README loop (README.md:15-38): reset(); step(a0, 0); step(a1, 1);
           get_new_state(0); get_new_state(1) with one-hot uniform-random actions, until either reports terminal
  nfsp     leduc.newenv.Env driven by main.train (main.py:21-67) with agent.Agent.play (agent.py:118-156) and both
           memories (utils/replay_buffer.py, utils/ReservoirBuffer.py).  TensorFlow / Keras are not installed, so the
           two networks of an agent are replaced by np.random.rand(1, 1, 3) score vectors and update_strategy() is a
           no-op: the figure is an UPPER bound of the reference's rollout rate (every real decision adds a batch-1
           Keras predict, every 128 decisions a fit)
A transition = one Env.step(action, player).  `run()` plays `hands` hands on each of `procs` independent processes.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import random
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_COPY = os.path.join(ROOT, "baseline", "_ref")
FILES = ("config.ini", "main.py", "agent/agent.py", "leduc/cardmatrix.py", "leduc/deck.py", "leduc/env.py", "leduc/newenv.py",
         "utils/replay_buffer.py", "utils/ReservoirBuffer.py")


def materialise(src="/root/reference", dst=REF_COPY) -> bool:
    """Verbatim copy of the nine reference files the two drivers execute into baseline/_ref/ (git-ignored, NOT
    gpurun-ignored).  Returns False when the reference tree is not there (the GPU box: the copy already travelled)."""
    if not os.path.isfile(os.path.join(src, "leduc", "newenv.py")):
        return False
    for rel in FILES:
        os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), os.path.join(dst, rel))
    return True


def reference_root():
    """Where the reference's files are on this machine, or None."""
    for cand in (os.environ.get("NFSP_REFERENCE_ROOT"), "/root/reference", REF_COPY):
        if cand and os.path.isfile(os.path.join(cand, "leduc", "newenv.py")):
            return cand
    return None


def _load():
    root = reference_root()
    if root is None:
        raise RuntimeError("no reference tree (neither /root/reference nor baseline/_ref)")
    os.environ["NFSP_REFERENCE_ROOT"] = root
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import ref_runner as rr

    rr.REF_ROOT = root
    return rr, rr.load()


def play_legacy(hands: int, seed: int = 0):
    """README driver on leduc.env.Env; returns (transitions, seconds)."""
    import numpy as np

    rr, mods = _load()
    env = mods["env"].Env()
    rng = random.Random(seed)
    onehot = [np.array([1.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0]), np.array([0.0, 0.0, 1.0])]
    transitions = 0
    t0 = time.perf_counter()
    for _ in range(hands):
        env.reset()
        for _ in range(16):  # a hand ends within 5 README iterations; the bound only guards the loop
            env.step(onehot[rng.randrange(3)], 0)
            env.step(onehot[rng.randrange(3)], 1)
            transitions += 2
            r0 = env.get_new_state(0)
            r1 = env.get_new_state(1)
            if r0[3] or r1[3]:
                break
    return transitions, time.perf_counter() - t0


class _RandomNet:
    """Stands where a Keras model would: predict() is the np.random.rand(1, 1, 3) of SURVEY 8d config 1."""

    def predict(self, x):
        import numpy as np

        return np.random.rand(1, 1, 3)


def play_nfsp(hands: int, seed: int = 0):
    """main.train + Agent.play + both memories on leduc.newenv.Env; returns (transitions, seconds)."""
    import numpy as np

    rr, mods = _load()
    random.seed(seed)
    np.random.seed(seed)
    env = mods["newenv"].Env()
    Base = rr.agent_class()

    class Agent(Base):
        total_played = 0

        def average_payoff_br(self):  # learner statistic (agent.py:234-238): outside the rollout path
            return 0.0

        def sampled_actions(self):  # agent.py:196-204 zeroes `played` every 100 hands: keep the sum
            self.total_played += self.played
            Base.sampled_actions(self)

    buf = 40000  # config.ini Utils.Buffersize, what agent.py:59-64 passes to both memories
    players = [Agent("Player%d" % p, env, _RandomNet(), _RandomNet(), mods["replay_buffer"].ReplayBuffer(buf, 1234),
                     mods["ReservoirBuffer"].ReservoirBuffer(buf, 1234)) for p in range(2)]
    train = rr.train_function(episodes=hands, eta=0.1)
    t0 = time.perf_counter()
    train(env, players[0], players[1])
    dt = time.perf_counter() - t0
    return int(sum(p.total_played + p.played for p in players)), dt


def _worker(args):
    kind, hands, seed = args
    return (play_legacy if kind == "legacy" else play_nfsp)(hands, seed)


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run(kind: str, hands: int, procs: int = 1):
    """`hands` hands on each of `procs` processes.  Rate = all transitions / the slowest process' loop time (imports and
    the construction of the env are outside the clock, as the GPU arm's set-up is)."""
    if procs == 1:
        res = [_worker((kind, hands, 0))]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_worker, [(kind, hands, s) for s in range(procs)])
    trans = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return {"kind": kind, "procs": procs, "hands_per_proc": hands, "transitions": trans, "seconds": slowest,
            "transitions_per_sec": trans / slowest, "hands_per_sec": procs * hands / slowest, "cpu": cpu_model()}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "materialise":
        print("copied" if materialise() else "reference tree not found")
    else:
        import json

        n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
        for kind in ("legacy", "nfsp"):
            print(json.dumps(run(kind, n, int(sys.argv[1]) if len(sys.argv) > 1 else 1)))
