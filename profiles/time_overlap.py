"""The insert launch beside the next rollout (SelfPlay(overlap_insert=True)) against the plain sequence, step by step as the
reservoirs' tickets / capacity grows (Algorithm R accepts capacity / ticket of the records: the insert gets cheaper).
CUDA events; the insert's window is taken on its own stream.  python profiles/time_overlap.py [blocks_of_steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nfsp_b200  # noqa: E402

N, T, CAP = 1 << 20, 8, 1 << 23
blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda", 0)
mk = lambda ov: nfsp_b200.SelfPlay(N, seed=1234, device=dev, eta=0.1, epsilon=0.06, rl_capacity=1 << 25, sl_capacity=CAP,  # noqa: E731
                                   max_steps_per_call=T, direct_rings=True, overlap_insert=ov)
objs = {"sequential": mk(False), "overlap": mk(True)}
for sp in objs.values():
    while min(int(m.total.item()) for m in sp.sl) < CAP:
        sp.rollout(T)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
main = torch.cuda.current_stream(dev)
fill = lambda sp: min(int(m.total.item()) for m in sp.sl) / float(CAP)  # noqa: E731


def steps(sp, k=10, trace=False):
    """mean step time in us over the last k - 2 of k steps; with trace (overlap object only) also the insert's window"""
    marks, orig = {}, getattr(sp, "_flush_set", None)

    def traced(kk, beside=False):
        s = torch.cuda.current_stream(dev)
        marks["i0"], marks["i1"] = ev(), ev()
        marks["i0"].record(s)
        orig(kk, beside)
        marks["i1"].record(s)

    if trace:
        sp._flush_set = traced
    rows = []
    for _ in range(k):
        flush.zero_()  # L2 flush, outside the timed events
        a, c = ev(), ev()
        a.record(main)
        sp.rollout(T, refresh_weights=True)
        c.record(main)
        c.synchronize()
        rows.append([a.elapsed_time(c) * 1e3] + ([a.elapsed_time(marks["i0"]) * 1e3, a.elapsed_time(marks["i1"]) * 1e3] if trace else []))
    if trace:
        sp._flush_set = orig
    rows = rows[2:]
    return [sum(r[j] for r in rows) / len(rows) for j in range(len(rows[0]))]


print("2^20 games x 8 decisions, reservoirs of 2^23 records per player, L2 flushed between steps; us per step")
print("%-22s %12s %12s %28s" % ("tickets / capacity", "sequential", "overlap", "insert window (traced steps)"))
for b in range(blocks):
    f0 = fill(objs["overlap"])
    s = steps(objs["sequential"])[0]
    o = steps(objs["overlap"])[0]
    f1 = fill(objs["overlap"])
    t = steps(objs["overlap"], k=6, trace=True)
    steps(objs["sequential"], k=6)  # keep the two objects' reservoirs equally full
    print("%5.2f -> %5.2f          %9.1f    %9.1f        %5.1f -> %5.1f of %5.1f" % (f0, f1, s, o, t[1], t[2], t[0]))
