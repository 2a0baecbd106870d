"""Device time of the insert launch of a bench step (SL records of 2^20 games x 8 decisions into both reservoirs) and of
the same launch on an empty batch (its fixed cost: cooperative launch, count scan, grid barrier, commit).

    python profiles/time_flush.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfsp_b200  # noqa: E402

sp = nfsp_b200.SelfPlay(1 << 20, seed=1234, rl_capacity=1 << 25, sl_capacity=1 << 23, max_steps_per_call=8, direct_rings=True)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
full, empty = [], []
for k in range(40):
    sp.rollout(8, insert=False)
    a, b, c = ev(), ev(), ev()
    a.record()
    sp.flush()
    b.record()
    sp.flush()
    c.record()
    c.synchronize()
    if k >= 5:
        full.append(a.elapsed_time(b) * 1e3)
        empty.append(b.elapsed_time(c) * 1e3)
    if k % 10 == 9:
        print("after %2d steps: reservoir totals %s  insert %.1f us, empty insert %.1f us" % (
            k + 1, [int(m.total.item()) for m in sp.sl], sum(full[-5:]) / 5, sum(empty[-5:]) / 5))
