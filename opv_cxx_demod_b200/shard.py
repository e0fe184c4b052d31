"""Multi-GPU host logic: streams are independent (the reference runs one process per stream and
shares nothing, src/opv-demod.cpp:999-1001), so ranks own contiguous blocks of stream ids and the
only exchange is a sum of per-rank counters plus a max of elapsed time."""
from __future__ import annotations

import torch
import torch.distributed as dist


def stream_range(rank: int, world: int, n_streams: int) -> tuple[int, int]:
    """Block partition: rank r owns [lo, hi); sizes differ by at most one."""
    base, rem = divmod(int(n_streams), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_counters(local: dict, device: torch.device) -> dict:
    """Sum integer counters over all ranks (NCCL on GPUs, gloo in the CPU tests)."""
    keys = sorted(local)
    t = torch.tensor([int(local[k]) for k in keys], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: int(v) for k, v in zip(keys, t.tolist())}


def reduce_max_ms(ms: float, device: torch.device) -> float:
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
