"""Instance sharding across the GPUs of one box.

MHPC instances are fully independent (each MHPCLocomotion object of the reference is self-contained), so the job is
partitioned into contiguous blocks of ceil(N / G) instances per rank; there is NO collective inside the solve.  The only
communication is one final gather of the per-instance result table (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the global instance indices owned by `rank` (contiguous blocks of ceil(N/G))."""
    per = (n_total + world - 1) // world
    b = min(rank * per, n_total)
    return b, min(b + per, n_total)


def gather_result_table(local, n_total: int, world: int, dist=None):
    """All-gather the per-rank result tables ([rows, per] tensors, `per` = ceil(N/G) columns, padded) and return the
    global [rows, n_total] table on every rank.  `dist` = torch.distributed (already initialised) or None for world 1."""
    import torch
    if dist is None or world == 1:
        return local[:, :n_total].clone()
    parts = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous())
    per = (n_total + world - 1) // world
    return torch.cat([p[:, :per] for p in parts], dim=1)[:, :n_total]
