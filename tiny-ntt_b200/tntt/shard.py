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


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_thread_to_gpu(device: int) -> dict:
    """Pin the calling process to the CPUs that sit next to GPU `device` (its PCIe root / NUMA node).

    The host pipeline (tntt_polymul_host) streams 96 KB per polynomial over PCIe; when the feeding process
    runs on the other socket, every pinned page is first-touched remotely and the copies cross the inter-socket
    link.  One process per GPU, each bound to its GPU's `local_cpulist`, keeps the staging buffers local.
    Returns what was done (for the benchmark's log); never raises -- an unknown topology just leaves the
    affinity alone."""
    import os

    info = {"device": device, "bound": False}
    try:
        import torch

        p = torch.cuda.get_device_properties(device)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        with open(os.path.join(base, "local_cpulist")) as fh:
            local = _parse_cpulist(fh.read())
        try:
            with open(os.path.join(base, "numa_node")) as fh:
                info["numa_node"] = int(fh.read().strip())
        except (OSError, ValueError):
            pass
        allowed = os.sched_getaffinity(0)
        target = (local & allowed) or set()
        info["pci"] = bdf
        if target and target != allowed:
            os.sched_setaffinity(0, target)
            info["bound"] = True
        info["cpus"] = len(target or allowed)
    except Exception as exc:  # topology not visible (container without /sys, no such attribute, ...)
        info["error"] = type(exc).__name__
    return info
