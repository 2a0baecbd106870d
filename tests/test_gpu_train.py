"""-m gpu: the outer training loop (main.train drop-in) and the exact exploitability (8 f-2)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import orc  # noqa: E402


@pytest.fixture(scope="module")
def nb():
    import nfsp_b200

    assert torch.cuda.is_available()
    return nfsp_b200


def test_policy_tables_come_from_the_cuda_forward(nb):
    """Every information state of the game through the CUDA forward: probabilities within 1e-5 of the oracle MLP,
    and the exact exploitability of freshly initialised nets is a positive number of chips per hand."""
    from nfsp_b200 import exploitability as ex

    sp = nb.SelfPlay(64, seed=5)
    w = sp.weights.cpu().numpy()
    nets = orc.Nets([dict(W1=w[k, :1920].reshape(30, 64), b1=w[k, 1920:1984], W2=w[k, 1984:2176].reshape(64, 3),
                          b2=w[k, 2176:]) for k in range(4)])
    states = ex.information_states()
    assert len(states[0]) == len(states[1]) == 195   # 15 in round 0 + 180 in round 1
    for p in range(2):
        m = np.asarray(states[p], np.int64).astype(np.int32)
        got = sp.forward(torch.from_numpy(m), torch.full((len(m),), 2 * p, dtype=torch.int8)).cpu().numpy()
        x = ((m[:, None] >> np.arange(30)) & 1).astype(np.float32)
        ref = nets.forward(2 * p, x, "avg")
        assert np.abs(got - ref).max() <= 1e-5
    a = ex.exploitability(sp)
    b = ex.exploitability(sp)
    assert a == b and a["value"] > 0 and a["nash_conv"] == a["br_value"][0] + a["br_value"][1]
    assert a["value"] < 5.0   # nobody can win more than the largest pot per hand


@pytest.mark.parametrize("seed,scale", [(5, 1.0), (11, 6.0)])
def test_best_response_value_matches_a_play_out(nb, seed, scale):
    """Independent check of the exact best-response value (replaces agent.py:234-238's proxy): the
    best-response TABLE the expectimax produces is played against the other seat's average-policy net through the
    batched CUDA env (Philox deals, the CUDA forward on every decision) for 1.15e7 hands; the mean reward must land within
    its confidence interval of br_value.  Glorot nets, and nets with 6x larger weights (a less uniform policy)."""
    from nfsp_b200 import exploitability as ex

    sp = nb.SelfPlay(64, seed=seed)
    if scale != 1.0:
        sp.set_weights(sp.weights * scale)
    full = ex.exploitability(sp)
    tables = ex.policy_tables(lambda m, k: sp.forward(torch.from_numpy(m), torch.from_numpy(k)).cpu().numpy())
    for b in (0, 1):
        v, pol = ex.best_response_policy(b, tables[b ^ 1])
        assert v == full["br_value"][b] and len(pol) > 0
        mean, se, hands = ex.play_best_response(sp, b, pol, hands=11 << 20, games=1 << 20)
        assert hands >= 10_000_000 and 0 < se < 2e-3
        assert abs(mean - v) <= 4.0 * se, (b, v, mean, se)
        # and it really is a BEST response: the same seat playing its own average net earns no more
        own = {k: tables[b][k] for k in pol}
        mean_own, se_own, _ = ex.play_best_response(sp, b, own, hands=4 << 20, games=1 << 20)
        assert mean_own <= v + 4.0 * se_own


@pytest.mark.parametrize("pipelined", [False, True])
def test_train_loop_runs_and_reports(nb, pipelined):
    from nfsp_b200 import main as drv
    from nfsp_b200.learner import Learner

    sp = nb.SelfPlay(4096, seed=1234, rl_capacity=200000, sl_capacity=200000, max_steps_per_call=8)
    w0 = sp.weights.clone()
    learner = Learner(sp, cfg=nb.load_config(None))
    lines = []
    rows = drv.train(sp, learner, episodes=60_000, steps_per_call=8, report_every=10, log=lines.append, pipelined=pipelined)
    assert rows and rows[-1]["hands"] >= 60_000
    assert rows[-1]["transitions"] == rows[-1]["calls"] * 8 * 4096
    for p in range(2):
        assert abs(sum(rows[-1]["actions"][p]) - 1.0) < 1e-9
    assert np.isfinite(rows[-1]["exploitability"]["value"]) and rows[-1]["exploitability"]["value"] >= 0
    assert not torch.equal(w0, sp.weights)            # the learner moved the nets
    assert any("Exploitability" in s for s in lines)  # main.py:73's line is printed


def test_pipelined_trainer_equals_the_sequential_loop_run_with_one_update_of_lag(nb):
    """PipelinedTrainer (update j on its own stream beside rollout j+1, the rollout grid four SMs short) against the
    plain loop made to act with the same nets (rollout j with W_{j-1}): the events must order everything, so weights,
    game words and memories come out bit-identical (deterministic=True: one staging segment per block of 32 games, so the
    order of the staged records does not depend on how the blocks were scheduled)."""
    from nfsp_b200.learner import Learner, PipelinedTrainer

    def make():
        sp = nb.SelfPlay(4096, seed=21, eta=0.3, epsilon=0.2, rl_capacity=1 << 15, sl_capacity=1 << 15, max_steps_per_call=8,
                         deterministic=True)
        for _ in range(3):
            sp.rollout(8)
        return sp

    K = 7
    a, b = make(), make()
    La, Lb = Learner(a, cfg=nb.load_config(None)), Learner(b, cfg=nb.load_config(None))
    chain = a.weights.clone()                    # W_j, trained in place by the plain learner
    hist = [chain.clone()]                       # hist[j] = W_j
    for j in range(K):
        a.set_weights(hist[j - 1 if j >= 1 else 0].clone())
        a.rollout(8)
        a.set_weights(chain)
        assert La.update(sync=False)["trained"] == 0xF
        hist.append(chain.clone())
    T = PipelinedTrainer(b, Lb)
    for j in range(K):
        assert T.step(8)["trained"] == 0xF
    final = T.finish()
    torch.cuda.synchronize()
    assert torch.equal(final, hist[K]) and not torch.equal(hist[K], hist[0])
    assert torch.equal(a.env.state_words(), b.env.state_words())
    for p in range(2):
        assert int(a.rl[p].total.item()) == int(b.rl[p].total.item()) and torch.equal(a.rl[p].data, b.rl[p].data)
        assert int(a.sl[p].total.item()) == int(b.sl[p].total.item()) and torch.equal(a.sl[p].data, b.sl[p].data)
    assert La.iteration == Lb.iteration and abs(a.epsilon - b.epsilon) < 1e-15


def test_state_table_rollout_trains_like_the_per_decision_rollout(nb):
    """Self-play with the learner in the loop, once with the nets tabulated over the decision states (the default) and
    once with a forward per decision: the same games, counters and RL memories after every iteration (the actions agree
    although the scores differ in their last bits), SL records and trained nets within float rounding of each other."""
    from nfsp_b200.learner import Learner

    def make(variant):
        return nb.SelfPlay(4096, seed=17, eta=0.3, epsilon=0.2, rl_capacity=1 << 15, sl_capacity=1 << 15, max_steps_per_call=8,
                           deterministic=True, variant=variant)

    a, b = make("states"), make("cuda")
    La, Lb = Learner(a, cfg=nb.load_config(None)), Learner(b, cfg=nb.load_config(None))
    w0 = a.weights.clone()
    for j in range(10):
        a.rollout(8)
        b.rollout(8)
        if j >= 2:
            assert La.update(sync=False)["trained"] == Lb.update(sync=False)["trained"] == 0xF
        assert torch.equal(a.env.state_words(), b.env.state_words()), j
        assert a.read_stats() == b.read_stats()
    torch.cuda.synchronize()
    for p in range(2):
        assert torch.equal(a.rl[p].data, b.rl[p].data) and int(a.sl[p].total.item()) == int(b.sl[p].total.item())
        sa, sb = a.sl[p].data.cpu(), b.sl[p].data.cpu()
        assert torch.equal(sa[:, 0], sb[:, 0])                                           # the observations
        assert (sa[:, 1:].view(torch.float32) - sb[:, 1:].view(torch.float32)).abs().max() <= 1e-5   # the score vectors
    assert (a.weights - b.weights).abs().max() <= 1e-4 and (a.weights - w0).abs().max() > 1e-3


def test_pipelined_trainer_with_direct_rings(nb):
    """The production combination of bench.py: the rollout writes the rings in place while the previous update is still
    running.  The update works on a copy of its sampled records taken before the rollout may start (the `gathered`
    event), so nothing it reads can change under it: nets stay finite and move, every record is accounted for."""
    from nfsp_b200.learner import Learner, PipelinedTrainer

    sp = nb.SelfPlay(8192, seed=3, eta=0.3, epsilon=0.2, rl_capacity=1 << 18, sl_capacity=1 << 16, max_steps_per_call=8,
                     direct_rings=True)
    for _ in range(3):
        sp.rollout(8)
    w0 = sp.weights.clone()
    T = PipelinedTrainer(sp, Learner(sp, cfg=nb.load_config(None)))
    for _ in range(12):
        assert T.step(8)["trained"] == 0xF
    w = T.finish()
    torch.cuda.synchronize()
    st = sp.read_stats()
    assert torch.isfinite(w).all() and not torch.equal(w, w0)
    assert st["transitions"] == 15 * 8 * 8192 and st["dropped"] == 0
    assert int(sp.counts.sum().item()) == 0 and all(int(sp.rl[p].total.item()) > 8192 for p in range(2))
