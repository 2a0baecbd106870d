"""-m gpu: learner kernels against a float64 numpy restatement (oracle/learner_oracle.py).
Floating point: 1e-5 absolute on gradients (values are O(0.1)); NN parity is unpinned (no Keras)."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import learner_oracle as lo  # noqa: E402
from oracle import orc  # noqa: E402


@pytest.fixture(scope="module")
def nb():
    import nfsp_b200

    assert torch.cuda.is_available()
    return nfsp_b200


def _filled_selfplay(nb, n=4096, steps=8, rounds=4, **kw):
    kw.setdefault("deterministic", True)  # several tests hold two such objects against each other: same record order
    sp = nb.SelfPlay(n, seed=5, eta=0.3, epsilon=0.2, rl_capacity=1 << 16, sl_capacity=1 << 16,
                     max_steps_per_call=steps, **kw)
    w = sp.weights.clone()
    w[:, 1920:1984] = 0.05   # non-zero biases so that relu gates are exercised on both sides
    w[:, 2176:] = 0.1
    sp.set_weights(w)
    for _ in range(rounds):
        sp.rollout(steps)
    return sp


def test_gradients_match_float64_oracle(nb):
    from nfsp_b200 import _lib
    from nfsp_b200.learner import GRAD, Learner

    sp = _filled_selfplay(nb)
    L = Learner(sp, minibatch=128, gamma=0.95)
    L.target = (sp.weights[[1, 3]] * 0.9 + 0.01).contiguous()
    idx_rl = [sp.rl[p].sample_slots(128)[0] for p in range(2)]
    idx_sl = [sp.sl[p].sample_slots(128)[0] for p in range(2)]
    w0 = sp.weights.clone()
    for row0, rows in ((0, 32), (96, 32), (0, 128)):
        sp.weights.copy_(w0)
        L.lr_br, L.lr_ar = [0.0, 0.0], 0.0          # gradients only
        L._step(idx_rl, idx_sl, row0, rows, 0xF)
        flat = L.flat.cpu().numpy().astype(np.float64)
        W = w0.cpu().numpy()
        T = L.target.cpu().numpy()
        for p in range(2):
            rl = sp.rl[p].data.cpu().numpy().view(np.uint8).reshape(-1).view(orc.RL_DT)[idx_rl[p].cpu().numpy()][row0:row0 + rows]
            g, expl, loss = lo.br_grad(W[2 * p + 1], T[p], rl["s"], rl["s2"], rl["a"].astype(int), rl["r"], rl["t"], 0.95)
            got = flat[(2 * p + 1) * 2179:(2 * p + 2) * 2179]
            assert np.abs(got - g).max() < 1e-5, (p, np.abs(got - g).max())
            assert abs(flat[GRAD + p] - expl) < 1e-3 and flat[GRAD + 2 + p] == rows
            assert abs(flat[GRAD + 4 + 2 * p + 1] - loss) < 1e-3
            sl = sp.sl[p].data.cpu().numpy().view(np.uint8).reshape(-1).view(orc.SL_DT)[idx_sl[p].cpu().numpy()][row0:row0 + rows]
            g, loss = lo.avg_grad(W[2 * p], sl["s"], sl["a"])
            got = flat[(2 * p) * 2179:(2 * p + 1) * 2179]
            assert np.abs(got - g).max() < 1e-5, (p, np.abs(got - g).max())
            assert abs(flat[GRAD + 4 + 2 * p] - loss) < 1e-2
    assert float(np.abs(flat[:GRAD]).max()) > 1e-4  # not vacuous


@pytest.mark.parametrize("others", [False, True])
def test_gradients_match_torch_autograd(nb, others):
    """The same gradients against torch.autograd in float64 on the CPU: the backward pass is DERIVED by autograd from the
    forward definition (Dense-relu-Dense with a relu head and the Huber loss of agent.py:91-99 averaged over the three
    outputs as Keras does, resp. a softmax head with categorical cross-entropy), so nothing of the hand-written backward --
    kernel or numpy restatement -- is shared."""
    import torch.nn.functional as F

    from nfsp_b200.learner import GRAD, Learner

    sp = _filled_selfplay(nb)
    L = Learner(sp, minibatch=128, gamma=0.95, others_to_target=others)
    L.target = (sp.weights[[1, 3]] * 0.9 + 0.01).contiguous()
    idx_rl = [sp.rl[p].sample_slots(128)[0] for p in range(2)]
    idx_sl = [sp.sl[p].sample_slots(128)[0] for p in range(2)]
    L.lr_br, L.lr_ar = [0.0, 0.0], 0.0
    row0, rows = 32, 64
    L._step(idx_rl, idx_sl, row0, rows, 0xF)
    flat = L.flat.cpu().double()
    W, T = sp.weights.cpu().double(), L.target.cpu().double()

    def parts(w):
        return w[:1920].reshape(30, 64), w[1920:1984], w[1984:2176].reshape(64, 3), w[2176:2179]

    def bits(m):
        return torch.from_numpy(((np.asarray(m, np.uint32)[:, None] >> np.arange(30)) & 1).astype(np.float64))

    def net(w, x):
        W1, b1, W2, b2 = parts(w)
        return F.relu(x @ W1 + b1) @ W2 + b2

    for p in range(2):
        rl = np.ascontiguousarray(sp.rl[p].data.cpu().numpy()).view(np.uint8).reshape(-1).view(orc.RL_DT)[idx_rl[p].cpu().numpy()][row0:row0 + rows]
        w = W[2 * p + 1].clone().requires_grad_(True)
        q = F.relu(net(w, bits(rl["s"])))
        with torch.no_grad():
            qn = F.relu(net(T[p], bits(rl["s2"]))).max(1).values
            y = torch.from_numpy(rl["r"].astype(np.float64)) + 0.95 * (1.0 - torch.from_numpy(rl["t"].astype(np.float64))) * qn
        a_idx = torch.from_numpy(rl["a"].astype(np.int64))
        with torch.no_grad():  # the fit's target matrix: the target net's predictions on s (agent.py:220) or, by default,
            tgt = F.relu(net(T[p], bits(rl["s"]))) if others else q.detach().clone()   # the net's own (zero error)
            tgt[torch.arange(rows), a_idx] = y                                          # agent.py:241, every row its own
        err = tgt - q
        huber = torch.where(err.abs() > 1, err.abs() - 0.5, 0.5 * err * err)   # agent.py:91-99
        huber.mean(dim=1).mean().backward()                                      # Keras: mean over the 3 outputs, mean over rows
        got = flat[(2 * p + 1) * 2179:(2 * p + 2) * 2179]
        assert float((got - w.grad).abs().max()) < 1e-5 and float(w.grad.abs().max()) > 1e-4
        sl = np.ascontiguousarray(sp.sl[p].data.cpu().numpy()).view(np.uint8).reshape(-1).view(orc.SL_DT)[idx_sl[p].cpu().numpy()][row0:row0 + rows]
        w = W[2 * p].clone().requires_grad_(True)
        logp = F.log_softmax(net(w, bits(sl["s"])), dim=1)
        (-(torch.from_numpy(sl["a"].astype(np.float64)) * logp).sum(1)).mean().backward()   # categorical cross-entropy, target as is
        got = flat[(2 * p) * 2179:(2 * p + 1) * 2179]
        assert float((got - w.grad).abs().max()) < 1e-5 and float(w.grad.abs().max()) > 1e-4


def test_sgd_apply_and_net_mask(nb):
    from nfsp_b200.learner import GRAD, Learner

    sp = _filled_selfplay(nb, n=1024, rounds=2)
    L = Learner(sp, lr_br=0.05, lr_ar=0.1)
    idx_rl = [sp.rl[p].sample_slots(128)[0] for p in range(2)]
    idx_sl = [sp.sl[p].sample_slots(128)[0] for p in range(2)]
    w0 = sp.weights.clone()
    L._step(idx_rl, idx_sl, 0, 32, 0b0110)          # only p0's BR net and p1's average net train
    g = L.flat[:GRAD].reshape(4, 2179)
    assert float(g[0].abs().max()) == 0 and float(g[3].abs().max()) == 0
    lr = torch.tensor([0.1, 0.05, 0.1, 0.05], device=sp.device)[:, None]
    assert torch.allclose(sp.weights, w0 - lr * g, atol=1e-7)
    assert torch.equal(sp.weights[0], w0[0]) and not torch.equal(sp.weights[1], w0[1])


def test_update_runs_schedules(nb):
    """Learner.update(): the reference's schedules (agent.py:245-253) and the target-net sync."""
    from nfsp_b200.learner import Learner

    sp = _filled_selfplay(nb)
    L = Learner(sp, cfg=nb.load_config(None))
    eps0 = sp.epsilon
    first = L.update()
    assert first["trained"] == 0xF and L.iteration == [2, 2] and L.target_update_count == [1, 1]
    assert abs(sp.epsilon - eps0 / 2) < 1e-12                      # epsilon ** 1 / iteration
    assert abs(L.lr_br[0] - 0.05 / (1 + 0.003 * 2 ** 0.5)) < 1e-12
    t0 = L.target.clone()                                           # synced at count 0 (before this update's copy)
    for _ in range(5):
        L.update()
    assert L.target_update_count == [6, 6] and torch.equal(L.target, t0)   # next sync only at count 150
    assert not torch.equal(L.target[0], sp.weights[1])
    out = sp.rollout(4, insert=False, debug=True)                   # the rollout runs on the trained weights
    assert torch.isfinite(out["vec"]).all()


@pytest.mark.parametrize("minibatch,fit_batch", [(128, 32), (100, 24), (320, 80)])
def test_fit_in_one_launch_equals_the_step_by_step_sequence(nb, minibatch, fit_batch):
    """nfsp_learner_fit (8 SGD steps, weights in registers) against 8 x (nfsp_learner_grads, nfsp_sgd_apply) on the same
    sampled rows: same statistics, weights equal to fp32 rounding of the same arithmetic (tolerance 1e-6 absolute on
    weights of magnitude <= 0.3; the two paths differ at most in fused-multiply-add contraction)."""
    from nfsp_b200.learner import Learner

    a, b = _filled_selfplay(nb), _filled_selfplay(nb)
    assert torch.equal(a.weights, b.weights)
    # (128, 32): the reference's sizes; (100, 24): ragged last slice; (320, 80): beyond the row-parallel kernel's shared
    # memory, played by the register kernel
    La, Lb = (Learner(x, minibatch=minibatch, fit_batch=fit_batch, fused=f) for x, f in ((a, True), (b, False)))
    for k in range(3):
        ra, rb = La.update(), Lb.update()
        assert ra["trained"] == rb["trained"] == 0xF
        assert np.allclose(ra["loss"], rb["loss"], rtol=1e-5, atol=1e-6)
        assert abs(ra["exploitability"] - rb["exploitability"]) < 1e-5
        d = (a.weights - b.weights).abs().max().item()
        assert d < 1e-6, (k, d)


def test_peer_exchange_protocol_two_ranks_on_one_gpu(nb):
    """nfsp_learner_fit_peers (the all-reduce of every SGD step inside the kernel) with TWO ranks played by two kernels
    running side by side on two streams of one GPU, each with its own exchange buffer.  Reference: the same eight steps
    as nfsp_learner_grads of both ranks, their gradients added on the device, nfsp_sgd_apply with scale 1/2.  Weights
    must be bit-identical between the ranks and within 1e-6 of the reference; the statistics are the sums over ranks."""
    from nfsp_b200 import _lib
    from nfsp_b200._lib import check
    from nfsp_b200.batched import _ptr, _stream
    from nfsp_b200.learner import GRAD, Learner

    sps = [_filled_selfplay(nb), _filled_selfplay(nb, n=2048)]      # different games -> different memories
    w0 = sps[0].weights.clone()
    sps[1].set_weights(w0.clone())
    Ls = [Learner(sp, fused=False) for sp in sps]
    idx = [L._sample_positions() for L in Ls]
    lr = (C.c_float * 4)(0.1, 0.05, 0.1, 0.05)
    # reference, step by step
    stats_ref = None
    for _ in range(2):
        for row0 in range(0, 128, 32):
            for L, (irl, isl) in zip(Ls, idx):
                io = L._io(irl, isl, row0, 32, 0xF)
                check(nb.lib().nfsp_learner_grads(C.byref(io), _stream(L.device)))
            total = Ls[0].flat + Ls[1].flat
            if stats_ref is None:
                stats_ref = total[GRAD:].clone()
            for L in Ls:
                check(nb.lib().nfsp_sgd_apply(_ptr(L.sp.weights), _ptr(total), C.byref(lr), 0.5, _stream(L.device)))
    ref = [L.sp.weights.clone() for L in Ls]
    assert torch.equal(ref[0], ref[1])
    # the same through the peer exchange: two kernels that wait for each other
    dev = sps[0].device
    bufs = [torch.zeros(_lib.PEER_BUF_FLOATS, dtype=torch.float32, device=dev) for _ in range(2)]
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    for epoch0 in (0, 8):   # two calls: the parities and epochs carry on from one fit to the next
        for L in Ls:
            L.sp.weights.copy_(w0)
        torch.cuda.synchronize()
        for r, (L, (irl, isl)) in enumerate(zip(Ls, idx)):
            io = L._io(irl, isl, 0, 128, 0xF)
            p = _lib.Peers()
            p.world, p.rank, p.epoch0, p.d_err = 2, r, epoch0, err.data_ptr()
            for q in range(2):
                p.d_buf[q] = bufs[q].data_ptr()
            with torch.cuda.stream(streams[r]):
                check(nb.lib().nfsp_learner_fit_peers(C.byref(io), 128, 32, 2, lr, _ptr(L.sp.weights), C.byref(p),
                                                         _stream(dev)))
        torch.cuda.synchronize()
        assert int(err.item()) == 0, "a rank timed out waiting for its peer"
        assert torch.equal(Ls[0].sp.weights, Ls[1].sp.weights)
        assert (Ls[0].sp.weights - ref[0]).abs().max().item() < 1e-6
        for L in Ls:
            assert torch.allclose(L.flat[GRAD:], stats_ref, rtol=1e-5, atol=1e-5)


def test_peer_exchange_on_real_gpus_when_there_are_two():
    """tests/mgpu_learner_check.py under torchrun (peer memory over NVLink against the NCCL path); needs >= 2 GPUs."""
    import os
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(root, "tests", "mgpu_learner_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "mgpu learner check ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_sgd_reduces_loss_on_a_fixed_minibatch(nb):
    from nfsp_b200.learner import GRAD, Learner

    sp = _filled_selfplay(nb)
    L = Learner(sp, lr_br=0.05, lr_ar=0.1)
    idx_rl = [sp.rl[p].sample_slots(128)[0] for p in range(2)]
    idx_sl = [sp.sl[p].sample_slots(128)[0] for p in range(2)]

    def losses():
        lr = (L.lr_br, L.lr_ar)
        L.lr_br, L.lr_ar = [0.0, 0.0], 0.0
        L._step(idx_rl, idx_sl, 0, 128, 0xF)
        L.lr_br, L.lr_ar = lr
        return L.flat[GRAD + 4:GRAD + 8].cpu().numpy().copy()

    before = losses()
    for _ in range(40):
        L._step(idx_rl, idx_sl, 0, 128, 0xF)
    after = losses()
    assert (after < before).all(), (before, after)


def test_empty_memories_do_not_train(nb):
    from nfsp_b200.learner import Learner

    sp = nb.SelfPlay(64, rl_capacity=1024, sl_capacity=1024)
    w0 = sp.weights.clone()
    assert Learner(sp).update() == {"trained": 0} and torch.equal(sp.weights, w0)


def test_dropin_agent_update_networks(nb):
    """agent.agent.Agent.update_best_response_network / update_avg_response_network on its own memories."""
    nb.install_dropin()
    import agent.agent as agent
    import leduc.newenv as leduc

    env = leduc.Env()
    ag = agent.Agent(None, env.observation_space, env.action_space, "Player0", env)
    rng = np.random.RandomState(0)
    for _ in range(3):
        s = (rng.rand(100, 30) < 0.2).astype(np.float64)
        s2 = (rng.rand(100, 30) < 0.2).astype(np.float64)
        ag._rl_memory.add(s, np.eye(3)[rng.randint(0, 3, 100)], rng.randint(-4, 5, 100) * 0.5, s2, rng.rand(100) < 0.3)
        ag._sl_memory.add(s, rng.rand(100, 3))
    w_br, w_avg = ag.weights[ag.BR].clone(), ag.weights[ag.AVG].clone()
    ag.update_strategy()
    assert not torch.equal(ag.weights[ag.BR], w_br) and not torch.equal(ag.weights[ag.AVG], w_avg)
    assert ag.iteration == 2 and ag.target_br_model_update_count == 1 and ag.average_payoff_br() >= 0
