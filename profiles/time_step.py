"""Device time of one bench step -- rollout(8) + the move of the records into the memories -- staged vs direct ring append.

    python profiles/time_step.py [reps] [games]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfsp_b200  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
T = 8
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for variant, direct in (("cuda", False), ("cuda", True), ("states", False), ("states", True)):
    sp = nfsp_b200.SelfPlay(n, seed=1234, rl_capacity=1 << 25, sl_capacity=1 << 23, max_steps_per_call=T, variant=variant,
                            direct_rings=direct)
    tot, ker = [], []
    for k in range(reps + 3):
        flush.zero_()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        sp.rollout(T, insert=False)
        b.record()
        sp.flush()
        c.record()
        c.synchronize()
        if k >= 3:
            tot.append(a.elapsed_time(c))
            ker.append(a.elapsed_time(b))
    tot.sort()
    ker.sort()
    print("%-7s direct=%d  step median %.4f ms (rollout kernel %.4f)  -> %.3e transitions/s   rl totals %s dropped %d" % (
        variant, direct, tot[len(tot) // 2], ker[len(ker) // 2], n * T / (tot[len(tot) // 2] * 1e-3),
        [int(sp.rl[p].total.item()) for p in range(2)], sp.read_stats()["dropped"]))
