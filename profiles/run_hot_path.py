"""Short driver for ncu: a few passes of the hot path at bench size (1M games; 8 decisions per launch for the
fused rollout, 32 transitions per game and launch for the env-only kernels, as in bench.py).

    python profiles/run_hot_path.py [rollout|env|legacy|all] [passes]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfsp_b200  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, T = 1 << 20, 8
if what in ("rollout", "all"):
    sp = nfsp_b200.SelfPlay(n, seed=1234, rl_capacity=1 << 25, sl_capacity=1 << 23, max_steps_per_call=T, direct_rings=True)
    for _ in range(passes):
        sp.rollout(T)
    torch.cuda.synchronize()
    print("rollout", sp.read_stats())
if what in ("env", "all"):
    env = nfsp_b200.BatchedNfspEnv(n, seed=1234)
    env.reset()
    for _ in range(passes):
        env.step(n_steps=32, trace=True)
    torch.cuda.synchronize()
    print("env ok")
if what in ("legacy", "all"):
    leg = nfsp_b200.BatchedLegacyEnv(n, seed=1234)
    leg.reset()
    for _ in range(passes):
        leg.rollout(16, trace=True)
    torch.cuda.synchronize()
    print("legacy ok")
