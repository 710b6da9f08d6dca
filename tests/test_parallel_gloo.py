"""CPU test of the N>1 path: two gloo ranks each aggregate their shard of the SSTs (with the oracle as the
per-rank engine stand-in — no GPU here) and combine with the same collective code bench.py uses over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from horaedb_b200 import sstgen
from horaedb_b200.parallel import combine_partials, shard_files


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ssts, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    schema = sstgen.metric_storage_schema()
    mine = [ssts[i] for i in shard_files(len(ssts), rank, world)]
    preds = [("tag", "eq", 3)]
    a = oracle.scan_aggregate(mine, schema.arrow_schema, 2, preds, group_col=0, ts_col=1, window_ms=60_000, value_col=2)
    t = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x.astype(dt)))
    res = combine_partials(t(a.gkey, np.int64), t(a.bucket, np.int64), t(a.count, np.int64), t(a.sum, np.float64),
                           t(a.min, np.float64), t(a.max, np.float64))
    if rank == 0:
        out.put([r.numpy() for r in res])
    dist.barrier()
    dist.destroy_process_group()


def test_shard_files_partitions_everything():
    for n in (1, 5, 16, 17):
        for w in (1, 2, 3, 8):
            got = [i for r in range(w) for i in shard_files(n, r, w)]
            assert got == list(range(n))


def test_two_rank_combine_matches_single_rank():
    from oracle import oracle
    ssts = [sstgen.synth_sst(lo, lo + 10, 300, 10_000, seq=50 + i)[0] for i, lo in enumerate(range(0, 50, 10))]
    schema = sstgen.metric_storage_schema()
    exp = oracle.scan_aggregate(ssts, schema.arrow_schema, 2, [("tag", "eq", 3)], group_col=0, ts_col=1, window_ms=60_000, value_col=2)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ssts, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gkey, bucket, count, sum_, mn, mx = got
    assert gkey.astype(np.uint64).tolist() == exp.gkey.tolist()
    assert bucket.tolist() == exp.bucket.tolist()
    assert count.astype(np.uint64).tolist() == exp.count.tolist()
    assert np.array_equal(sum_, exp.sum) and np.array_equal(mn, exp.min) and np.array_equal(mx, exp.max)
