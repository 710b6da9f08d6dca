"""Multi-GPU plumbing for the scan path: SSTs shard by file, one process per GPU, ONE collective that
combines the per-rank partial aggregates.  Works on any torch.distributed backend (NCCL on GPUs, gloo in CPU tests).

Group keys that contain the series id are disjoint across ranks when every SST belongs to exactly one rank and SSTs are
PK-disjoint, so the combine is an all-gather of (key, bucket, count, sum, min, max) rows — exact, no floating-point
re-association (sums travel as their bit patterns)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def shard_files(n_files: int, rank: int, world: int) -> List[int]:
    """Contiguous block of file indices owned by `rank` (files are ordered by PK range, so blocks stay PK-disjoint)."""
    per, rem = divmod(n_files, world)
    lo = rank * per + min(rank, rem)
    return list(range(lo, lo + per + (1 if rank < rem else 0)))


class PartialCombiner:
    """One all-gather per call of a padded [6, cap] int64 block; `cap` (max groups of any rank) is agreed once and
    re-agreed only when a rank outgrows it, so the steady state has no size exchange and no host synchronisation."""

    def __init__(self):
        self.cap = 0
        self.block = None
        self.out = None

    def gather(self, gkey, bucket, count, sum_, mn, mx, check_cap: bool = True) -> torch.Tensor:
        """Returns the [world, 6, cap] block (rows: key, bucket, count, sum bits, min bits, max bits); count == 0 marks
        padding.  With `check_cap` (collective: every rank must pass the same flag) the ranks first agree on the padded
        capacity; `check_cap=False` skips that synchronising exchange when the caller knows the sizes are stable."""
        k = gkey.numel()
        if check_cap or self.block is None:
            t = torch.tensor([k], device=gkey.device, dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            need = max(int(t.item()), 1)
            if self.block is None or need > self.cap:
                self.cap = need
                self.block = torch.zeros(6, self.cap, device=gkey.device, dtype=torch.int64)
                self.out = torch.zeros(dist.get_world_size() * 6, self.cap, device=gkey.device, dtype=torch.int64)
        assert k <= self.cap, "partial aggregate outgrew the agreed capacity: call gather(check_cap=True)"
        self.block.zero_()
        if k:
            self.block[:, :k] = torch.stack([gkey.to(torch.int64), bucket.to(torch.int64), count.to(torch.int64),
                                             sum_.to(torch.float64).view(torch.int64), mn.to(torch.float64).view(torch.int64),
                                             mx.to(torch.float64).view(torch.int64)])
        dist.all_gather_into_tensor(self.out, self.block)          # concatenation along dim 0 (works on NCCL and gloo)
        return self.out.view(dist.get_world_size(), 6, self.cap)


    def gather_packed(self, engine, num_groups: int, device, check_cap: bool = False) -> torch.Tensor:
        """Same collective, but the engine packs its last aggregate straight into the send block (one kernel, no
        per-field tensor views): the steady-state step is one C call + one all-gather."""
        if check_cap or self.block is None:
            t = torch.tensor([num_groups], device=device, dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            need = max(int(t.item()), 1)
            if self.block is None or need > self.cap:
                self.cap = need
                self.block = torch.zeros(6, self.cap, device=device, dtype=torch.int64)
                self.out = torch.zeros(dist.get_world_size() * 6, self.cap, device=device, dtype=torch.int64)
        assert num_groups <= self.cap
        engine.export_packed(self.block.data_ptr(), self.cap)
        dist.all_gather_into_tensor(self.out, self.block)
        return self.out.view(dist.get_world_size(), 6, self.cap)


def combine_partials(gkey: torch.Tensor, bucket: torch.Tensor, count: torch.Tensor, sum_: torch.Tensor, mn: torch.Tensor,
                     mx: torch.Tensor, combiner: Optional[PartialCombiner] = None):
    """All ranks receive every rank's partial aggregate rows, ordered by (gkey, bucket).
    Inputs: 1-D tensors of equal length on the rank's device (int64 keys/buckets/counts, float64 sum/min/max)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return gkey, bucket, count, sum_, mn, mx
    combiner = combiner or PartialCombiner()
    out = combiner.gather(gkey, bucket, count, sum_, mn, mx)
    allb = out.permute(1, 0, 2).reshape(6, -1)
    allb = allb[:, allb[2] > 0]
    order = torch.argsort(allb[1], stable=True)
    order = order[torch.argsort(allb[0][order], stable=True)]
    allb = allb[:, order]
    return (allb[0], allb[1], allb[2], allb[3].view(torch.float64), allb[4].view(torch.float64), allb[5].view(torch.float64))
