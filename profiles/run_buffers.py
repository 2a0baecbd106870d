"""ncu driver for the insert kernel: batches of 2^23 records into a 2^24-slot ring / reservoir (bench.py's buffer leg).

    python profiles/run_buffers.py [ring|reservoir|all] [batches]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfsp_b200  # noqa: E402

if os.environ.get("NFSP_L2_FETCH"):  # cudaLimitMaxL2FetchGranularity (0x05), a hint: 32, 64 or 128 bytes
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    print("cudaDeviceSetLimit ->", rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["NFSP_L2_FETCH"]))))
    v = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
    print("L2 fetch granularity now", v.value)
what = sys.argv[1] if len(sys.argv) > 1 else "all"
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda")
n_rec = 1 << 23
recs = torch.randint(0, 1 << 30, (n_rec, 4), dtype=torch.int32, device=dev)
for name, mem in (("ring", nfsp_b200.DeviceRing(1 << 24, 1, dev)), ("reservoir", nfsp_b200.DeviceReservoir(1 << 24, 2, dev))):
    if what not in (name, "all"):
        continue
    for k in range(batches):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cnt = torch.tensor([n_rec], dtype=torch.int32, device=dev)
        a.record()
        mem.insert(recs, cnt)
        b.record()
        b.synchronize()
        print(name, "batch", k, "%.4f ms" % a.elapsed_time(b), flush=True)
