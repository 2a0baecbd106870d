"""Two or more GPUs of one box (not collected by pytest; run it under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_learner_check.py

The fit with the all-reduce inside the kernel (peer memory, nfsp_learner_fit_peers) against the same fit as eight
launch pairs with one NCCL all-reduce each, on identical memories and sampled rows: same statistics, weights equal to
1e-6 absolute (the two differ in the order the ranks' gradients are summed), and -- for the peer path -- weights
bit-identical on every rank."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import nfsp_b200  # noqa: E402
from nfsp_b200.learner import Learner  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, steps = 4096, 8


def filled():
    sp = nfsp_b200.SelfPlay(n, seed=5, game0=rank * n, device=dev, eta=0.3, epsilon=0.2, rl_capacity=1 << 16,
                            sl_capacity=1 << 16, max_steps_per_call=steps, deterministic=True)
    w = sp.weights.clone()
    w[:, 1920:1984] = 0.05
    w[:, 2176:] = 0.1
    sp.set_weights(w)
    for _ in range(4):
        sp.rollout(steps)
    return sp


a, b, c, e4 = filled(), filled(), filled(), filled()
La, Lb, Lc = Learner(a, fused=True), Learner(b, fused=False), Learner(c, fused=True, use_multicast=False)
Ld = Learner(e4, fused=True, peer_transport="ipc")  # buffers shared as plain CUDA IPC handles by the library itself
assert La._peers is not None, getattr(La, "_peer_note", "no peers")
assert Ld._peers is not None and Ld._peer_transport == "ipc", getattr(Ld, "_peer_note", "no ipc peers")
if rank == 0:
    print("transport:", La._peer_transport, " multicast (NVLS) pushes:", La._multicast)
for k in range(4):
    ra, rb, rc, rd = La.update(), Lb.update(), Lc.update(), Ld.update()
    assert torch.equal(a.weights, c.weights), "multicast and unicast pushes must give identical weights"
    assert torch.equal(a.weights, e4.weights), "symmetric-memory and IPC buffers must give identical weights"
    assert ra["trained"] == rb["trained"] == 0xF, (ra, rb)
    assert np.allclose(ra["loss"], rb["loss"], rtol=1e-5, atol=1e-5), (ra["loss"], rb["loss"])
    assert abs(ra["exploitability"] - rb["exploitability"]) < 1e-5
    d = (a.weights - b.weights).abs().max().item()
    assert d < 1e-6, (k, d)
    mine = a.weights.clone()
    ref = mine.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(mine, ref), "peer path: weights differ between ranks"
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
for name, L in (("peer exchange", La), ("peer, unicast pushes", Lc), ("peer, CUDA IPC buffers", Ld), ("nccl per step", Lb)):
    dist.barrier()
    s, e = ev(), ev()
    s.record()
    for _ in range(10):
        L.update(sync=False)
    e.record()
    e.synchronize()
    if rank == 0:
        print("%-22s %.1f us per update" % (name, s.elapsed_time(e) * 100))
assert int(La._peer_err.item()) == 0 and int(Lc._peer_err.item()) == 0 and int(Ld._peer_err.item()) == 0
Ld.close_peers()  # unmaps and frees the IPC buffers; the learner carries on over NCCL
assert Ld._peers is None
Ld.update(), Lb.update()
assert (e4.weights - b.weights).abs().max().item() < 1e-5
if rank == 0:
    print("mgpu learner check ok: world", world)
dist.destroy_process_group()
