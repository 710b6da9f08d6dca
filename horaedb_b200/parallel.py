"""Multi-GPU plumbing for the scan path: SSTs shard by file, one process per GPU, ONE collective that
combines the per-rank partial aggregates.  Works on any torch.distributed backend (NCCL on GPUs, gloo in CPU tests).

Group keys that contain the series id are disjoint across ranks when every SST belongs to exactly one rank and SSTs are
PK-disjoint, so the combine is an all-gather of (key, bucket, count, sum, min, max) rows followed by a sort by key —
exact, no floating-point re-association."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_files(n_files: int, rank: int, world: int) -> List[int]:
    """Contiguous block of file indices owned by `rank` (files are ordered by PK range, so blocks stay PK-disjoint)."""
    per, rem = divmod(n_files, world)
    lo = rank * per + min(rank, rem)
    return list(range(lo, lo + per + (1 if rank < rem else 0)))


def combine_partials(gkey: torch.Tensor, bucket: torch.Tensor, count: torch.Tensor, sum_: torch.Tensor, mn: torch.Tensor,
                     mx: torch.Tensor):
    """All ranks receive the concatenation of every rank's partial aggregate rows, ordered by (gkey, bucket).

    Inputs are 1-D tensors of equal length on the rank's device (int64 keys/buckets/counts, float64 sum/min/max).
    One size exchange + one all_gather of a padded [6, cap] float64/int64-bit-cast block."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return gkey, bucket, count, sum_, mn, mx
    dev = gkey.device
    n = torch.tensor([gkey.numel()], device=dev, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    cap = int(max(int(s.item()) for s in sizes))
    block = torch.zeros(6, max(cap, 1), device=dev, dtype=torch.int64)
    k = gkey.numel()
    if k:
        block[0, :k] = gkey.to(torch.int64)
        block[1, :k] = bucket.to(torch.int64)
        block[2, :k] = count.to(torch.int64)
        block[3, :k] = sum_.to(torch.float64).view(torch.int64)   # bit-cast: sums travel exactly
        block[4, :k] = mn.to(torch.float64).view(torch.int64)
        block[5, :k] = mx.to(torch.float64).view(torch.int64)
    gathered = [torch.zeros_like(block) for _ in range(world)]
    dist.all_gather(gathered, block)
    parts = [g[:, : int(s.item())] for g, s in zip(gathered, sizes)]
    allb = torch.cat(parts, dim=1)
    # stable order by (gkey as unsigned, bucket): ranks own ascending PK ranges, so concatenation is already sorted when
    # files were sharded with shard_files(); sort anyway to be independent of the sharding
    order = torch.argsort(allb[1], stable=True)
    order = order[torch.argsort(allb[0][order], stable=True)]
    allb = allb[:, order]
    return (allb[0], allb[1], allb[2], allb[3].view(torch.float64), allb[4].view(torch.float64), allb[5].view(torch.float64))
