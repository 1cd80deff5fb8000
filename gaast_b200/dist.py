"""Multi-GPU plumbing: one process per GPU (torch.distributed), batch sharding.

Evaluation has no exchange step: batch elements are independent, so every rank
evaluates a contiguous slice of the batch with its own plan replica and its own
copy of the broadcast operands.  The only collective of the
path is the optional batch-sum: each rank reduces its slice on the device
(fused in the kernel epilogue) and the per-rank vectors -- sum over root grades
of C(n,k) doubles, 66 for cfg5 -- are combined with one all-reduce (NCCL over
NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Tuple


def rank_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n: int, rank: int, world: int, align: int = 2) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of a batch of n elements for `rank`.

    Slices start on multiples of `align` elements (2 keeps every per-grade row
    segment 16-byte aligned for the 128-bit accesses); the sizes differ by at
    most `align`; the slices tile [0, n) exactly."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    units = (n + align - 1) // align
    base, extra = divmod(units, world)
    begin_u = rank * base + min(rank, extra)
    end_u = begin_u + base + (1 if rank < extra else 0)
    return min(n, begin_u * align), min(n, end_u * align)


def all_reduce_sum(t):
    """In-place sum over ranks of a small tensor (the batch-sum vector)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def max_over_ranks(value: float, device=None) -> float:
    """Max over ranks of a scalar (device timings are reported as the max)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
