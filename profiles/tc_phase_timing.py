import os, sys
sys.path.insert(0, "/root/repo")
import torch, nfsp_b200
n, T = 1 << 20, 8
sp = nfsp_b200.SelfPlay(n, seed=1234, rl_capacity=1 << 25, sl_capacity=1 << 23, max_steps_per_call=T, variant="tcgen05")
for _ in range(3):
    sp.rollout(T, insert=False); sp.counts.zero_()
torch.cuda.synchronize(); sp.stats.zero_()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); sp.rollout(T, insert=False); b.record(); b.synchronize()
st = sp.stats.cpu().numpy().astype("uint64")
steps = n * T / 128
print("ms", a.elapsed_time(b))
print("per group-step cycles: begin %.0f  sort+lock+mma+epilogue %.0f  result exchange %.0f  finish %.0f" % (
    int(st[13]) / steps, int(st[14]) / steps, (int(st[15]) & 0xFFFFFFFF) / steps, (int(st[15]) >> 32) / steps))
