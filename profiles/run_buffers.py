"""ncu driver for K3/K4 at bandwidth size: 2^23 staged records into 2^24-slot memories (as bench.py's buffers block)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfsp_b200  # noqa: E402

dev = torch.device("cuda", 0)
n = 1 << 23
recs = torch.randint(0, 1 << 30, (n, 4), dtype=torch.int32, device=dev)
ring = nfsp_b200.DeviceRing(1 << 24, 1, dev)
res = nfsp_b200.DeviceReservoir(1 << 24, 2, dev)
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    for mem in (ring, res):
        mem.insert(recs, torch.tensor([n], dtype=torch.int32, device=dev))
torch.cuda.synchronize()
print("buffers ok", int(ring.total.item()), int(res.total.item()))
