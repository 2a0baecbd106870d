"""Timeline of the e2e step of bench.py (1M games): device time between CUDA events after each call, and the host
time the calls take, with the GPU idle at the start of every step as in the timed region.

    python profiles/time_e2e.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfsp_b200  # noqa: E402

n, T = 1 << 20, 8
sp = nfsp_b200.SelfPlay(n, seed=1234, rl_capacity=1 << 25, sl_capacity=1 << 23, max_steps_per_call=T, direct_rings=True)
w_host = sp.weights.cpu().pin_memory()
stats_host = torch.empty(sp.stats.shape, dtype=sp.stats.dtype).pin_memory()
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
acc, host = {}, {}
for it in range(12):
    torch.cuda.synchronize()
    names, marks, ht = [], [ev()], [time.perf_counter()]
    marks[0].record()

    def mark(name):
        marks.append(ev()); marks[-1].record(); names.append(name); ht.append(time.perf_counter())

    sp.set_weights(w_host); mark("set_weights (H2D + 2 pack kernels)")
    sp.rollout(T, insert=False); mark("rollout")
    sp.flush(); mark("flush")
    sp.sample_minibatches(256, to_host=True); mark("sample + D2H slab")
    stats_host.copy_(sp.stats, non_blocking=True); mark("D2H counters")
    marks[-1].synchronize()
    t_end = time.perf_counter()
    if it >= 4:
        for k, name in enumerate(names):
            acc.setdefault(name, []).append(marks[k].elapsed_time(marks[k + 1]) * 1e3)
            host.setdefault(name, []).append((ht[k + 1] - ht[k]) * 1e6)
        acc.setdefault("total (events)", []).append(marks[0].elapsed_time(marks[-1]) * 1e3)
        host.setdefault("total (events)", []).append((t_end - ht[0]) * 1e6)
print("%-38s %10s %10s" % ("", "device us", "host us"))
for k in acc:
    print("%-38s %10.1f %10.1f" % (k, sum(acc[k]) / len(acc[k]), sum(host[k]) / len(host[k])))
