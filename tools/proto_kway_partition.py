"""Prototype (numpy, CPU) of the partition step of a SINGLE-PASS k-way merge — groundwork for replacing the log2(k)
pairwise merge-path passes of csrc/kernels.cu (merge_partition_kernel / merge_pass_kernel) with one pass:

  1. for every output tile boundary p = t * TILE find split positions (i_0 .. i_{k-1}), sum = p, such that every element left
     of the splits precedes every element right of them in the total order (key, run index, position) — `kway_split`;
  2. each CTA then loads its k sub-ranges (TILE records) into shared memory and merges them with a log2(k)-level merge
     tree there; global memory is read once and written once (64 B per record instead of 64 B x log2 k).

`kway_split` is multiway selection by bisection, written the way a warp would run it: lane j owns run j, every round one
pivot element is ranked in all runs at once (k binary searches, one per lane, over ever smaller windows) and all windows
shrink.  The largest window at least halves every round, so a warp needs O(log n) rounds in practice (asserted in the
test).  Not product code: nothing in horaedb_b200/ imports it."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def _lower_bound(run: np.ndarray, lo: int, hi: int, key, strict_before: bool) -> int:
    """First index in [lo, hi) whose element does NOT precede the pivot.  Elements of an EARLIER run that compare equal to
    the pivot key precede it (ties go to the lower run index, read.rs:412-427 / SortPreservingMergeExec), elements of a
    LATER run do not: `strict_before` = the run is earlier than the pivot's run."""
    side = "right" if strict_before else "left"
    return lo + int(np.searchsorted(run[lo:hi], key, side=side))


def kway_split(runs: Sequence[np.ndarray], p: int) -> Tuple[List[int], int]:
    """Splits (i_j) with sum(i_j) == p separating the p smallest elements of the k sorted runs under the total order
    (key, run, position).  Returns (splits, rounds)."""
    k = len(runs)
    lo = [0] * k
    hi = [len(r) for r in runs]
    total = sum(hi)
    assert 0 <= p <= total
    rounds = 0
    while True:
        below = sum(lo)
        if below == p:
            return lo, rounds                     # everything left of lo is known to be among the p smallest
        if sum(hi) == p:
            return hi, rounds
        rounds += 1
        # pivot: the middle element of the widest open window (one lane wins a warp-wide argmax)
        j = max(range(k), key=lambda q: hi[q] - lo[q])
        assert hi[j] > lo[j]
        m = (lo[j] + hi[j]) // 2
        key = runs[j][m]
        # rank of the pivot = elements preceding it: in its own run exactly m (positions break ties), elsewhere a
        # binary search inside the open window (everything left of lo precedes, everything right of hi does not)
        pos = [0] * k
        for q in range(k):
            pos[q] = m if q == j else _lower_bound(runs[q], lo[q], hi[q], key, strict_before=q < j)
        rank = sum(pos)
        if rank < p:                              # pivot and everything before it are among the p smallest
            for q in range(k):
                lo[q] = max(lo[q], pos[q] + (1 if q == j else 0))
        else:                                     # pivot is not among them, nor is anything after it
            for q in range(k):
                hi[q] = min(hi[q], pos[q])


def merge_by_tiles(runs: Sequence[np.ndarray], tile: int) -> np.ndarray:
    """Reference use of the splits: merge tile by tile (each tile only touches its own sub-ranges) and concatenate."""
    total = sum(len(r) for r in runs)
    bounds = [kway_split(runs, min(p, total))[0] for p in range(0, total + tile, tile)]
    out = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        parts = [(runs[q][a[q]:b[q]], q) for q in range(len(runs))]
        keys = np.concatenate([x for x, _ in parts]) if parts else np.empty(0)
        tags = np.concatenate([np.full(len(x), q) for x, q in parts])
        order = np.lexsort((tags, keys))          # stable in (key, run); positions are already ascending inside a run
        out.append(keys[order])
    return np.concatenate(out) if out else np.empty(0)
