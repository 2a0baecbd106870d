"""Probe: can two ranks on one box exchange peer pointers (torch symmetric memory) and read each other's memory?"""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem

    t = symm_mem.empty(1024, dtype=torch.float32, device=dev)
    t.fill_(float(rank + 1))
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok; attrs:", [a for a in dir(hdl) if not a.startswith("_")][:40], flush=True)
    print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs], flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
    print(rank, "peer value", float(peer[0].item()), flush=True)
    dist.barrier()
except Exception as e:  # noqa: BLE001
    print(rank, "symmetric memory failed:", repr(e), flush=True)
    # fallback probe: CUDA IPC through torch's reductions
    try:
        x = torch.full((1024,), float(rank + 1), device=dev)
        import torch.multiprocessing.reductions as red
        h = red.reduce_tensor(x)
        print(rank, "ipc handle ok", type(h), flush=True)
    except Exception as e2:  # noqa: BLE001
        print(rank, "ipc failed:", repr(e2), flush=True)
dist.destroy_process_group()
