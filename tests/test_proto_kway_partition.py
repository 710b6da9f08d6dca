"""The multiway-selection prototype for the planned single-pass k-way merge (tools/proto_kway_partition.py) against a
plain stable merge: random run lengths, heavy key duplication inside and across runs, empty runs, every boundary."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from proto_kway_partition import kway_split, merge_by_tiles  # noqa: E402


def _reference_order(runs):
    keys = np.concatenate(runs)
    tags = np.concatenate([np.full(len(r), q) for q, r in enumerate(runs)])
    pos = np.concatenate([np.arange(len(r)) for r in runs])
    order = np.lexsort((pos, tags, keys))
    return keys[order], tags[order], pos[order]


def test_splits_separate_the_p_smallest_for_every_p():
    rng = np.random.default_rng(0)
    for trial in range(30):
        k = int(rng.integers(1, 18))
        runs = [np.sort(rng.integers(0, int(rng.integers(2, 50)), int(rng.integers(0, 40)))) for _ in range(k)]
        keys, tags, pos = _reference_order(runs)
        total = len(keys)
        for p in range(total + 1):
            splits, rounds = kway_split(runs, p)
            assert sum(splits) == p
            want = [int(((tags[:p] == q)).sum()) for q in range(k)]     # how many of the p smallest come from run q
            assert splits == want, (trial, p)


def test_tile_merge_equals_full_merge_and_round_count_is_logarithmic():
    rng = np.random.default_rng(1)
    runs = [np.sort(rng.integers(0, 5000, int(rng.integers(1000, 6000)))) for _ in range(16)]
    runs[3] = np.empty(0, dtype=np.int64)                              # an empty run
    runs[7] = np.full(3000, 1234)                                      # one key only
    keys, _, _ = _reference_order(runs)
    assert np.array_equal(merge_by_tiles(runs, 1024), keys)
    total = len(keys)
    worst = max(kway_split(runs, p)[1] for p in range(0, total, 997))
    assert worst <= 4 * math.ceil(math.log2(total)), worst              # a warp-round per halving of the widest window
