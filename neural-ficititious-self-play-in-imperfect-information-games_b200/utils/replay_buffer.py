"""Drop-in for utils/replay_buffer.py `ReplayBuffer` (circular replay memory M_RL), stored on the
GPU as 16-byte packed records {s mask, s2 mask, reward, argmax(a), terminal}.

Divergences from the reference, both documented in DESIGN.md:
  * records are stored BY VALUE (the reference stores numpy views of env buffers that later mutate,
    replay_buffer.py:31-35 -- an aliasing bug);
  * `a` is kept as its argmax (the only function of it the learner reads, agent.py:241) and comes
    back one-hot; states must be the env's 0/1 observation vectors.
"""
import numpy as np
import torch

from ..batched import DeviceRing, RL_DT, obs_to_mask


def _pack_rl(s, a, r, s2, t):
    s = np.asarray(s, dtype=np.float64).reshape(-1, 30)
    s2 = np.asarray(s2, dtype=np.float64).reshape(-1, 30)
    a = np.asarray(a, dtype=np.float64).reshape(-1, 3)
    n = s.shape[0]
    rec = np.zeros(n, RL_DT)
    rec["s"] = obs_to_mask(torch.from_numpy(s)).numpy().astype(np.uint32)
    rec["s2"] = obs_to_mask(torch.from_numpy(s2)).numpy().astype(np.uint32)
    rec["a"] = np.argmax(a, axis=1)
    rec["r"] = np.asarray(r, dtype=np.float32).reshape(-1)
    rec["t"] = np.asarray(t).reshape(-1).astype(bool)
    return rec


class ReplayBuffer(object):
    def __init__(self, buffer_size, random_seed=123, device=None):
        self.buffer_size = buffer_size
        self.memory = DeviceRing(buffer_size, random_seed, device)
        self.last_recent_batch = 0

    @property
    def count(self):
        return self.memory.size()

    def add(self, s, a, r, s2, t):
        """replay_buffer.py:30-41 (one transition, or a leading batch axis)."""
        rec = _pack_rl(s, a, r, s2, t)
        dev = self.memory.device
        d = torch.from_numpy(rec.view(np.int32).reshape(-1, 4)).to(dev)
        n = torch.tensor([len(rec)], dtype=torch.int32, device=dev)
        self.memory.insert(d, n)

    def size(self):
        return self.memory.size()

    def sample_batch(self, batch_size):
        """replay_buffer.py:46-59: uniform without replacement, min(count, batch_size) rows;
        shapes s (B,1,30), a (B,1,3), r (B,), s2 (B,1,30), t (B,)."""
        b = min(self.size(), int(batch_size))
        if b == 0:
            z = np.zeros
            return z((0, 1, 30)), z((0, 1, 3)), z((0,)), z((0, 1, 30)), z((0,), dtype=bool)
        s, a, r, s2, t, _, _ = self.memory.sample(b)
        c = lambda x: x.cpu().numpy().astype(np.float64)  # noqa: E731
        return (c(s).reshape(b, 1, 30), c(a).reshape(b, 1, 3), c(r), c(s2).reshape(b, 1, 30),
                t.cpu().numpy().astype(bool))

    def recent_batch(self, batch_size):
        """replay_buffer.py:61-77: the reference slices from last_recent_batch = 0, i.e. samples the whole
        buffer; unused by the agent."""
        return self.sample_batch(batch_size)

    def reservoir_sample(self, batch_size):
        """replay_buffer.py:79-93 (unused by the agent; off-by-one quirks not reproduced): a uniform
        sample with the reference's output shapes (a as (B,3))."""
        s, a, r, s2, t = self.sample_batch(batch_size)
        return s, a.reshape(-1, 3), r, s2, t

    def clear(self):
        self.memory.clear()
