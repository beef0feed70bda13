"""Batch sharding across the GPUs of one box.

Every polynomial pair is independent (new_reference/cg_ntt.py:78-92 has no cross-row
dependency), so the batch is cut into contiguous row ranges, one per rank, and NO collective
touches the data path: NVLink / NVSwitch stay idle by design.  torch.distributed is used only
for the start barrier and the max-over-ranks timing of bench.py.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of rank `rank`: ceil(total/world) rows each, the tail rank(s) get less."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    if total_rows < 0:
        raise ValueError("negative batch")
    per = -(-total_rows // world_size)
    lo = min(rank * per, total_rows)
    return lo, min(lo + per, total_rows)


def shard_rows(x, world_size: int, rank: int):
    """The rows of a [batch, n] tensor / array that rank `rank` owns (a view, no copy)."""
    lo, hi = shard_range(x.shape[0], world_size, rank)
    return x[lo:hi]


def max_over_ranks(value: float) -> float:
    """Max of a per-rank scalar (device time); identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
