"""The insert launch of a bench step in its steady state (both reservoirs full), for ncu.

    ncu --set full --clock-control none --import-source on -k regex:insert_kernel --launch-skip 34 -c 1 -o gpurun_out/insert python profiles/run_insert.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfsp_b200  # noqa: E402

sp = nfsp_b200.SelfPlay(1 << 20, seed=1234, rl_capacity=1 << 25, sl_capacity=1 << 23, max_steps_per_call=8, direct_rings=True)
for _ in range(38):
    sp.rollout(8)
torch.cuda.synchronize()
print("ok", [int(m.total.item()) for m in sp.sl])
