"""ncu driver: python profiles/run_variant.py {cuda|tcgen05} [passes] -- fused rollout at bench size."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfsp_b200  # noqa: E402

variant = sys.argv[1]
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sp = nfsp_b200.SelfPlay(1 << 20, seed=1234, rl_capacity=1 << 25, sl_capacity=1 << 23, max_steps_per_call=8, variant=variant)
for _ in range(passes):
    sp.rollout(8, insert=False)
    sp.counts.zero_()
torch.cuda.synchronize()
print(variant, sp.read_stats())
