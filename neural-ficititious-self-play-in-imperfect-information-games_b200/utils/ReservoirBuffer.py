"""Drop-in for utils/ReservoirBuffer.py `ReservoirBuffer` (supervised memory M_SL) on the GPU:
16-byte records {s mask, a[3] float32}.

mode="R" (default) is true reservoir sampling (Vitter's Algorithm R), which is what NFSP specifies;
mode="reference" reproduces the reference's law (ReservoirBuffer.py:26-28: a full buffer replaces a
uniformly drawn slot in 1..B-1 with probability (B-1)/B; slot 0 is never replaced)."""
import numpy as np
import torch

from ..batched import DeviceReservoir, SL_DT, obs_to_mask


class ReservoirBuffer(object):
    def __init__(self, buffer_size, random_seed=123, device=None, mode="R"):
        self.buffer_size = buffer_size
        self.memory = DeviceReservoir(buffer_size, random_seed, device, mode)
        self.last_recent_batch = 0

    @property
    def count(self):
        return self.memory.size()

    def add(self, s, a):
        s = np.asarray(s, dtype=np.float64).reshape(-1, 30)
        a = np.asarray(a, dtype=np.float32).reshape(-1, 3)
        rec = np.zeros(s.shape[0], SL_DT)
        rec["s"] = obs_to_mask(torch.from_numpy(s)).numpy().astype(np.uint32)
        rec["a"] = a
        dev = self.memory.device
        d = torch.from_numpy(rec.view(np.int32).reshape(-1, 4)).to(dev)
        n = torch.tensor([len(rec)], dtype=torch.int32, device=dev)
        self.memory.insert(d, n)

    def size(self):
        return self.memory.size()

    def sample_batch(self, batch_size):
        """ReservoirBuffer.py:33-43: s (B,1,30), a (B,1,3); min(count, batch_size) rows."""
        b = min(self.size(), int(batch_size))
        if b == 0:
            return np.zeros((0, 1, 30)), np.zeros((0, 1, 3))
        s, a, _, _ = self.memory.sample(b)
        return (s.cpu().numpy().astype(np.float64).reshape(b, 1, 30),
                a.cpu().numpy().astype(np.float64).reshape(b, 1, 3))
