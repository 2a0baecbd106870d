"""Trace files: hands on disk, replayed on the CPU oracle (here) and through the general env
kernel (-m gpu).  Integer work => word-for-word equality."""
import numpy as np
import pytest

from oracle import orc


def _planes_from_oracle(n, steps, seed=99, eta=0.25):
    b = orc.NfspBatch(n, seed)
    b.reset(0, orc.u32_frac(eta))
    tr = b.rollout_env(1, steps, orc.u32_frac(eta))
    return np.stack([tr["obs"], tr["reward"].view(np.uint32), tr["misc"]])


def _replay_on_oracle(tf, words, transitions):
    h = tf.unpack_hands(words)
    for i in range(len(words)):
        g = orc.NfspSingle()
        g.reset(int(h["dealer"][i]), *(int(c) for c in h["cards"][i]))
        p = int(h["dealer"][i])
        for k in range(int(h["n_actions"][i])):
            rec = transitions[i, k]
            p = int(rec[0] >> 31)                      # the actor the kernel recorded ...
            vec = np.zeros(3)
            vec[int(h["actions"][i, k])] = 1.0
            g.step(vec, p)
            s, a, r, s2, t = g.get_state(p)
            assert g.obs(p) == int(rec[0] & 0x3FFFFFFF), (i, k)
            assert bool(t) == bool((rec[0] >> 30) & 1), (i, k)
            assert np.float32(r if t else 0.0).view(np.uint32) == rec[1], (i, k)
            assert int(rec[2] & 3) == int(h["actions"][i, k])
        assert bool(t), i                              # ... and every recorded hand ends exactly at its last action


def test_pack_unpack_roundtrip():
    import nfsp_b200.tracefile as tf

    rng = np.random.RandomState(3)
    words = []
    for _ in range(200):
        acts = list(rng.randint(0, 3, rng.randint(1, 7)))
        words.append(tf.pack_hand(rng.randint(2), rng.randint(3), rng.randint(3), rng.randint(3), rng.randint(2), rng.randint(2), acts))
    h = tf.unpack_hands(words)
    for i, w in enumerate(words):
        acts = list(h["actions"][i, : h["n_actions"][i]])
        assert tf.pack_hand(h["dealer"][i], *h["cards"][i], *h["policy"][i], acts) == w
    codes = tf.action_codes(h)
    assert codes.shape == (6, 200) and ((codes == 4) == (np.arange(6)[:, None] >= h["n_actions"][None])).all()


def test_hands_from_oracle_trace_replay_on_oracle(tmp_path):
    import nfsp_b200.tracefile as tf

    words, trans = tf.hands_from_trace(_planes_from_oracle(300, 40))
    assert len(words) > 2000 and trans.shape == (len(words), 6, 3)
    h = tf.unpack_hands(words)
    assert h["n_actions"].min() >= 1 and h["n_actions"].max() <= 6
    assert abs(h["policy"].mean() - 0.25) < 0.03         # per-hand policy draws, P('b') = eta
    path = str(tmp_path / "hands.npz")
    tf.save(path, words, trans, games=300, steps=40)
    f = tf.load(path)
    assert np.array_equal(f["hands"], words) and np.array_equal(f["transitions"], trans)
    _replay_on_oracle(tf, words[:400], trans[:400])
    with pytest.raises(ValueError):
        np.savez(str(tmp_path / "x.npz"), magic="nope", hands=words, transitions=trans)
        tf.load(str(tmp_path / "x.npz"))


@pytest.mark.gpu
def test_record_with_lean_kernel_check_with_general_kernel(tmp_path):
    import nfsp_b200.tracefile as tf

    words, trans = tf.record(5000, 48, seed=7)
    assert len(words) > 50_000
    got = tf.replay_on_gpu(words)
    assert np.array_equal(got, trans)
    _replay_on_oracle(tf, words[:300], trans[:300])
    path = str(tmp_path / "hands.npz")
    tf.save(path, words, trans, games=5000, steps=48, seed=7)
    assert tf.main(["check", path]) == 0
    bad = trans.copy()
    bad[17, 0, 1] ^= 1
    tf.save(path, words, bad)
    assert tf.main(["check", path]) == 1
