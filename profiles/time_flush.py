"""Where the flush of a bench step goes: rollout(8) for 1M games, then each insert timed on its own with CUDA
events (warm, L2 not flushed: the staged records are as the rollout left them).

    python profiles/time_flush.py [n_games] [rl_capacity] [sl_capacity]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfsp_b200  # noqa: E402

n, T = (int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20), 8
rl_cap = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 25
sl_cap = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 23
sp = nfsp_b200.SelfPlay(n, seed=1234, rl_capacity=rl_cap, sl_capacity=sl_cap, max_steps_per_call=T)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
acc = {}
for it in range(8):
    marks = [ev()]
    marks[0].record()
    sp.rollout(T, insert=False)
    marks.append(ev()); marks[-1].record()
    names = ["rollout"]
    for p in range(2):
        sp.rl[p].insert(sp.stage_rl[p], sp.counts[p], sp.cap_rl)
        marks.append(ev()); marks[-1].record(); names.append("ring%d" % p)
        sp.sl[p].insert(sp.stage_sl[p], sp.counts[2 + p], sp.cap_sl)
        marks.append(ev()); marks[-1].record(); names.append("reservoir%d" % p)
    sp.sample_minibatches(256)
    marks.append(ev()); marks[-1].record(); names.append("sample_4x256")
    sp.rollout(T, insert=False)
    marks.append(ev()); marks[-1].record(); names.append("rollout_again")
    sp.flush()
    marks.append(ev()); marks[-1].record(); names.append("flush_two_streams")
    torch.cuda.synchronize()
    if it >= 3:
        for k, name in enumerate(names):
            acc.setdefault(name, []).append(marks[k].elapsed_time(marks[k + 1]) * 1e3)
        acc.setdefault("total", []).append(marks[0].elapsed_time(marks[-1]) * 1e3)
print("n_seg", sp.n_seg, "cap_rl", sp.cap_rl, "cap_sl", sp.cap_sl, "records per player: rl", [int(sp.rl[p].total.item()) // 8 for p in range(2)],
      "sl", [int(sp.sl[p].total.item()) // 8 for p in range(2)])
for k, v in acc.items():
    print("%-12s %8.1f us" % (k, sum(v) / len(v)))
