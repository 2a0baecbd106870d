"""numpy float64 restatement of the learner arithmetic (agent.py:90-116, 209-264) -- TEST INFRASTRUCTURE ONLY.

NN PARITY UNPINNED: Keras/TensorFlow are not installed and the reference pins no learner output, so this
follows the documented semantics of Dense/relu/softmax, the reference's huber_loss (agent.py:91-99), Keras'
categorical_crossentropy (target used as is) and plain SGD.  Each row gets its own TD target (the reference's
row-0-only overwrite, agent.py:241, is a bug that is not reproduced -- see DESIGN.md).
"""
import numpy as np


def split(w):
    w = np.asarray(w, np.float64)
    return w[:1920].reshape(30, 64), w[1920:1984], w[1984:2176].reshape(64, 3), w[2176:2179]


def bits(m):
    return ((np.asarray(m, np.uint32)[:, None] >> np.arange(30)) & 1).astype(np.float64)


def br_grad(w, w_target, s, s2, a, r, t, gamma, terminal_bootstraps=False, others_to_target=False):
    """Mean gradient (flat, 2179) of the Huber loss on the taken action; also the exploitability proxy sum."""
    W1, b1, W2, b2 = split(w)
    T1, tb1, T2, tb2 = split(w_target)
    x, x2 = bits(s), bits(s2)
    pre = x @ W1 + b1
    h = np.maximum(pre, 0)
    z = h @ W2 + b2
    q = np.maximum(z, 0)
    qn = np.maximum(np.maximum(x2 @ T1 + tb1, 0) @ T2 + tb2, 0).max(1)
    expl = np.maximum(np.maximum(x @ T1 + tb1, 0) @ T2 + tb2, 0).max(1).sum()
    live = np.ones(len(s)) if terminal_bootstraps else 1.0 - np.asarray(t, np.float64)
    y = np.asarray(r, np.float64) + gamma * live * qn
    n = len(s)
    rows = np.arange(n)
    qs = np.maximum(np.maximum(x @ T1 + tb1, 0) @ T2 + tb2, 0)       # target net on s: agent.py:220
    tgt = qs.copy() if others_to_target else q.copy()                   # outputs not taken: to qs, or zero error
    tgt[rows, a] = y
    err = tgt - q
    dz = -np.where(np.abs(err) > 1, np.sign(err), err) / 3.0 * (z > 0)
    loss = (np.where(np.abs(err) > 1, np.abs(err) - 0.5, 0.5 * err * err) / 3.0).sum()
    return _backprop(x, h, W2, dz, n), expl, loss


def avg_grad(w, s, y):
    W1, b1, W2, b2 = split(w)
    x = bits(s)
    h = np.maximum(x @ W1 + b1, 0)
    z = h @ W2 + b2
    e = np.exp(z - z.max(1, keepdims=True))
    p = e / e.sum(1, keepdims=True)
    y = np.asarray(y, np.float64)
    dz = p * y.sum(1, keepdims=True) - y
    loss = -(y * np.log(np.maximum(p, 1e-7))).sum()
    return _backprop(x, h, W2, dz, len(s)), loss


def _backprop(x, h, W2, dz, n):
    dh = (dz @ W2.T) * (h > 0)
    g = np.concatenate([(x.T @ dh).ravel(), dh.sum(0), (h.T @ dz).ravel(), dz.sum(0)]) / n
    return g
