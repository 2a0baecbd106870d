"""Multi-GPU layout of the rollout path: games shard by contiguous GLOBAL id range, one
process per GPU, and the rollout / insert / sample kernels exchange nothing.  The only collective on
the path is the all-reduce of the per-rank counters (the reference's per-100-hand stats print,
main.py:69-75 / agent.py:196-204, and the exploitability-stat gather)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_games(total_games: int, rank: int, world: int):
    """(first global game id, number of games) of `rank`; contiguous, sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, rem = divmod(int(total_games), int(world))
    n = base + (1 if rank < rem else 0)
    game0 = rank * base + min(rank, rem)
    return game0, n


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def allreduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """Sum the int64 rollout counters over ranks (NCCL on GPUs, gloo in the CPU tests)."""
    out = stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def max_over_ranks(value: float, device=None) -> float:
    """Timing rule: a multi-GPU number is the MAX over ranks of the device-measured time."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
