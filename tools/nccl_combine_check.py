"""Multi-GPU combine check (hg_comm_* / hg_agg_combine, csrc/comm.cu) against the CPU oracle.

  single process:  python tools/nccl_combine_check.py                (world 1: the REDUCE path on one GPU)
  N ranks:         python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
                       tools/nccl_combine_check.py

Every rank scans its own SSTs (contiguous series ranges), then
  GATHER  per-series partials (disjoint): the gathered blocks must equal every rank's oracle result, bit for bit;
  REDUCE  per-(tag, bucket) partials (HG_AGG_HASH; keys cross ranks): the combined table must equal the oracle's multi-shard
          definition — per-shard sequential sums, shards added in rank order — on every rank.
torch.distributed is only the host-side channel that ships the 128-byte NCCL id (what a Rust host does over its own RPC)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from horaedb_b200 import sstgen
    from oracle import oracle
    schema = sstgen.metric_storage_schema()
    files_per_rank, per = 3, 200
    all_datas = [[sstgen.synth_sst((r * files_per_rank + f) * per, (r * files_per_rank + f + 1) * per, 300, 1000, seq=100 + r * 10 + f,
                                   compression="snappy" if f % 2 == 0 else "none")[0] for f in range(files_per_rank)] for r in range(world)]
    import torch
    import torch.distributed as dist
    from horaedb_b200._ffi import HG_AGG_HASH, HG_COMBINE_GATHER, HG_COMBINE_REDUCE, DeviceArray, Engine, SchemaHandle, SstInput
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("gloo")          # the id travels over a CPU channel: the data path is the library's own NCCL communicator
    eng = Engine(device=local_rank)
    uid = [Engine.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    eng.comm_init(uid[0], rank, world)
    handle = SchemaHandle(schema.arrow_schema, 2)
    mine = [SstInput(id=1000 * rank + i, data=d) for i, d in enumerate(all_datas[rank])]
    t0 = sstgen.T0_MS
    preds = [("ts", "ge", t0 + 20_000), ("ts", "lt", t0 + 250_000)]

    def f64bits(a):
        return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)

    def to_host(ptr, n):
        t = torch.as_tensor(DeviceArray(ptr, n, "<i8"), device=f"cuda:{local_rank}")
        return t.cpu().numpy()

    # ---------------- GATHER: per-series sums
    kw = dict(group_col=0, ts_col=-1, window_ms=0, value_col=2)
    for hint in (0, 0, 777):
        eng.scan_aggregate_device(handle, mine, preds, **kw)
        cmb = eng.combine(HG_COMBINE_GATHER, hint)
        eng.comm_sync()
        cap = cmb.capacity
        blocks = to_host(cmb.d_blocks, world * 6 * cap).reshape(world, 6, cap)
        for r in range(world):
            exp = oracle.scan_aggregate(all_datas[r], schema.arrow_schema, 2, preds, **kw)
            g = len(exp.count)
            b = blocks[r]
            assert (b[2, g:] == 0).all() and g <= cap
            assert np.array_equal(b[0, :g], exp.gkey.astype(np.int64)) and np.array_equal(b[2, :g], exp.count.astype(np.int64))
            assert np.array_equal(b[3, :g], f64bits(exp.sum)) and np.array_equal(b[4, :g], f64bits(exp.min)) and np.array_equal(b[5, :g], f64bits(exp.max))
    # ---------------- REDUCE: per-(tag, bucket), keys cross ranks
    for kw in (dict(group_col=3, ts_col=1, window_ms=60_000, value_col=2), dict(group_col=3, ts_col=-1, window_ms=0, value_col=2),
               dict(group_col=-1, ts_col=1, window_ms=30_000, value_col=2)):
        eng.scan_aggregate_device(handle, mine, preds, mode=HG_AGG_HASH, **kw)
        cmb = eng.combine(HG_COMBINE_REDUCE, 0)
        eng.comm_sync()
        G, rc = cmb.num_groups, cmb.reduced_capacity
        tbl = to_host(cmb.d_reduced, 6 * rc).reshape(6, rc)[:, :G]
        # oracle, multi-shard definition: per-shard hash aggregation, shards combined in rank order
        acc = {}
        for r in range(world):
            e = oracle.scan_aggregate(all_datas[r], schema.arrow_schema, 2, preds, mode=1, **kw)
            for i in range(len(e.count)):
                key = (int(e.gkey[i]), int(e.bucket[i]))
                if key not in acc:
                    acc[key] = [int(e.count[i]), float(e.sum[i]), float(e.min[i]), float(e.max[i])]
                else:
                    a = acc[key]
                    a[0] += int(e.count[i]); a[1] = a[1] + float(e.sum[i]); a[2] = min(a[2], float(e.min[i])); a[3] = max(a[3], float(e.max[i]))
        keys = sorted(acc)
        assert G == len(keys), (G, len(keys))
        assert tbl[0].tolist() == [k[0] for k in keys] and tbl[1].tolist() == [k[1] for k in keys]
        assert tbl[2].tolist() == [acc[k][0] for k in keys]
        assert np.array_equal(tbl[3], f64bits(np.array([acc[k][1] for k in keys])))
        assert np.array_equal(tbl[4], f64bits(np.array([acc[k][2] for k in keys]))) and np.array_equal(tbl[5], f64bits(np.array([acc[k][3] for k in keys])))
    eng.comm_destroy()
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    print(f"rank {rank}/{world}: combine ok", flush=True)


if __name__ == "__main__":
    main()
