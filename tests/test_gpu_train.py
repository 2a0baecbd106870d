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


def test_train_loop_runs_and_reports(nb):
    from nfsp_b200 import main as drv
    from nfsp_b200.learner import Learner

    sp = nb.SelfPlay(4096, seed=1234, rl_capacity=200000, sl_capacity=200000, max_steps_per_call=8)
    w0 = sp.weights.clone()
    learner = Learner(sp, cfg=nb.load_config(None))
    lines = []
    rows = drv.train(sp, learner, episodes=60_000, steps_per_call=8, report_every=10, log=lines.append)
    assert rows and rows[-1]["hands"] >= 60_000
    assert rows[-1]["transitions"] == rows[-1]["calls"] * 8 * 4096
    for p in range(2):
        assert abs(sum(rows[-1]["actions"][p]) - 1.0) < 1e-9
    assert np.isfinite(rows[-1]["exploitability"]["value"]) and rows[-1]["exploitability"]["value"] >= 0
    assert not torch.equal(w0, sp.weights)            # the learner moved the nets
    assert any("Exploitability" in s for s in lines)  # main.py:73's line is printed
