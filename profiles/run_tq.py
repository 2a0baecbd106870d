"""A few launches of the warp-specialised rollout at bench size (driver for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfsp_b200  # noqa: E402

n, T = 1 << 20, 8
sp = nfsp_b200.SelfPlay(n, seed=1234, rl_capacity=1 << 25, sl_capacity=1 << 23, max_steps_per_call=T,
                        variant=sys.argv[1] if len(sys.argv) > 1 else "tcgen05_ws")
for _ in range(3):
    sp.rollout(T, insert=False)
    sp.counts.zero_()
torch.cuda.synchronize()
print(sp.read_stats())
