"""A few launches of the state-table rollout (variant 6) at the bench shape, for ncu.

    ncu --set full --clock-control none --import-source on -k regex:rollout_states --launch-skip 3 -c 1 -o gpurun_out/states python profiles/run_states.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfsp_b200  # noqa: E402

direct = (sys.argv[1] if len(sys.argv) > 1 else "1") == "1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
sp = nfsp_b200.SelfPlay(n, seed=1234, rl_capacity=1 << 25, sl_capacity=1 << 23, max_steps_per_call=8, variant="states", direct_rings=direct)
for _ in range(6):
    sp.rollout(8, insert=False)
    sp.flush()
torch.cuda.synchronize()
print("ok", sp.read_stats()["transitions"])
