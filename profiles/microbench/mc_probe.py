import os, torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem
buf = symm_mem.empty(1 << 16, dtype=torch.float32, device=f"cuda:{local}")
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
if rank == 0:
    print("attrs", [a for a in dir(hdl) if not a.startswith("_")])
    print("multicast_ptr", getattr(hdl, "multicast_ptr", None))
    print("buffer_ptrs", hdl.buffer_ptrs)
dist.barrier()
dist.destroy_process_group()
