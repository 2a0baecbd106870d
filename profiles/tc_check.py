"""tcgen05 forward vs CUDA-core forward vs oracle: accuracy on every reachable observation, and throughput."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import nfsp_b200
from oracle import orc

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "nfsp_exhaustive.npz"))
obs = np.unique(np.concatenate([g["obs_before"].ravel(), g["obs_after"].ravel()])).astype(np.uint32)
rng = np.random.RandomState(0)
w = np.zeros((4, 2179), np.float32)
for k in range(4):
    w[k, :1920] = rng.uniform(-0.25, 0.25, 1920)
    w[k, 1920:1984] = rng.uniform(-0.2, 0.2, 64)
    w[k, 1984:2176] = rng.uniform(-0.3, 0.3, 192)
    w[k, 2176:] = rng.uniform(-0.2, 0.2, 3)
sp = nfsp_b200.SelfPlay(64, weights=torch.from_numpy(w), rl_capacity=1024, sl_capacity=1024)
nets = orc.Nets([nfsp_b200.split_net(w[k]) for k in range(4)])
x = ((obs[:, None] >> np.arange(30)) & 1).astype(np.float32)
o_all = np.tile(obs, 4).astype(np.int32)
k_all = np.repeat(np.arange(4), len(obs)).astype(np.int8)
perm = rng.permutation(len(o_all))
o_all, k_all = o_all[perm], k_all[perm]
ref = np.zeros((len(o_all), 3), np.float32)
for k in range(4):
    m = k_all == k
    ref[m] = nets.forward(k, ((o_all[m].astype(np.uint32)[:, None] >> np.arange(30)) & 1).astype(np.float32), "br" if k & 1 else "avg")
a = sp.forward(torch.from_numpy(o_all), torch.from_numpy(k_all)).cpu().numpy()
b = sp.forward(torch.from_numpy(o_all), torch.from_numpy(k_all), tensor_cores=True).cpu().numpy()
print("rows", len(o_all), "cuda-core max err", np.abs(a - ref).max(), "tcgen05 max err", np.abs(b - ref).max())
n = 1 << 23
oo = torch.from_numpy(o_all).cuda()[torch.randint(0, len(o_all), (n,), device="cuda")]
mixes = {
    "uniform-random nets": torch.randint(0, 4, (n,), device="cuda", dtype=torch.int8),
    "NFSP mix 45/5/5/45": torch.tensor([0, 1, 3, 2], device="cuda", dtype=torch.int8)[
        torch.bucketize(torch.rand(n, device="cuda"), torch.tensor([0.45, 0.50, 0.55], device="cuda"))],
    "one net (Model.predict)": torch.zeros(n, device="cuda", dtype=torch.int8),
}
only = sys.argv[1] if len(sys.argv) > 1 else None
for mix, kk in mixes.items():
    if only and only not in mix:
        continue
    for name, tc in (("cuda-core", False), ("tcgen05", True)):
        for _ in range(3):
            sp.forward(oo, kk, tensor_cores=tc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            sp.forward(oo, kk, tensor_cores=tc)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("%-24s %-10s %.3f ms for %d rows -> %.3e rows/s" % (mix, name, ms, n, n / ms * 1e3))
