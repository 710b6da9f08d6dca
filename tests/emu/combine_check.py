"""TEST INFRASTRUCTURE: the library's multi-GPU combine (hg_comm_init / hg_agg_combine, csrc/comm.cu) with WORLD ranks as threads of this
process, each with its own engine of the emulated build (cuda_emu.h) and tests/emu/nccl_emu.cpp standing in for NCCL.  Same checks as
tools/nccl_combine_check.py runs on real GPUs:

  GATHER  per-series partials (disjoint keys): every rank's gathered blocks equal each rank's oracle result, bit for bit;
  REDUCE  per-(tag, bucket) / per-bucket partials (HG_AGG_HASH; keys cross ranks): the combined table equals the oracle's multi-shard
          definition — per-shard sequential sums, shards added in rank order — on every rank.

    python tests/emu/combine_check.py WORLD"""
import ctypes as C
import os
import sys
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
os.environ.setdefault("HORAE_EMU_CRASH_REPORT", "1")

import numpy as np  # noqa: E402

import build_engine_emu  # noqa: E402
from horaedb_b200 import _ffi  # noqa: E402

_ffi.LIB_PATH = build_engine_emu.build()
_ffi._lib = None

from horaedb_b200 import sstgen  # noqa: E402
from horaedb_b200._ffi import HG_AGG_HASH, HG_COMBINE_GATHER, HG_COMBINE_REDUCE, Engine, SchemaHandle, SstInput  # noqa: E402
from oracle import oracle  # noqa: E402


def f64bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def to_host(ptr, n):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int64)), shape=(n,)).copy()      # "device" memory of the emulated build


def rank_main(rank, world, uid, all_datas, schema, errors):
    try:
        eng = Engine(device=0)
        eng.comm_init(uid, rank, world)
        handle = SchemaHandle(schema.arrow_schema, 2)
        mine = [SstInput(id=1000 * rank + i, data=d) for i, d in enumerate(all_datas[rank])]
        t0 = sstgen.T0_MS
        preds = [("ts", "ge", t0 + 20_000), ("ts", "lt", t0 + 250_000)]
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
        for kw in (dict(group_col=3, ts_col=1, window_ms=60_000, value_col=2), dict(group_col=3, ts_col=-1, window_ms=0, value_col=2),
                   dict(group_col=-1, ts_col=1, window_ms=30_000, value_col=2)):
            eng.scan_aggregate_device(handle, mine, preds, mode=HG_AGG_HASH, **kw)
            cmb = eng.combine(HG_COMBINE_REDUCE, 0)
            eng.comm_sync()
            G, rc = cmb.num_groups, cmb.reduced_capacity
            tbl = to_host(cmb.d_reduced, 6 * rc).reshape(6, rc)[:, :G]
            acc = {}
            for r in range(world):
                e = oracle.scan_aggregate(all_datas[r], schema.arrow_schema, 2, preds, mode=1, **kw)
                for i in range(len(e.count)):
                    key = (int(e.gkey[i]), int(e.bucket[i]))
                    if key not in acc:
                        acc[key] = [int(e.count[i]), float(e.sum[i]), float(e.min[i]), float(e.max[i])]
                    else:
                        a = acc[key]
                        a[0] += int(e.count[i])
                        a[1] = a[1] + float(e.sum[i])
                        a[2] = min(a[2], float(e.min[i]))
                        a[3] = max(a[3], float(e.max[i]))
            keys = sorted(acc)
            assert G == len(keys), (G, len(keys))
            assert tbl[0].tolist() == [k[0] for k in keys] and tbl[1].tolist() == [k[1] for k in keys]
            assert tbl[2].tolist() == [acc[k][0] for k in keys]
            assert np.array_equal(tbl[3], f64bits(np.array([acc[k][1] for k in keys])))
            assert np.array_equal(tbl[4], f64bits(np.array([acc[k][2] for k in keys]))) and np.array_equal(tbl[5], f64bits(np.array([acc[k][3] for k in keys])))
        eng.comm_destroy()
        eng.close()
        print(f"rank {rank}/{world}: combine ok", flush=True)
    except BaseException as e:                                   # a failed rank must not leave the others waiting in an all-gather
        errors.append((rank, repr(e)))
        import traceback
        traceback.print_exc()
        os._exit(1)


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    schema = sstgen.metric_storage_schema()
    files_per_rank, per = 2, 60
    all_datas = [[sstgen.synth_sst((r * files_per_rank + f) * per, (r * files_per_rank + f + 1) * per, 300, 1000, seq=100 + r * 10 + f,
                                   compression="snappy" if f % 2 == 0 else "none")[0] for f in range(files_per_rank)] for r in range(world)]
    uid = Engine.comm_unique_id()
    errors = []
    threads = [threading.Thread(target=rank_main, args=(r, world, uid, all_datas, schema, errors)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise SystemExit(f"failed: {errors}")
    print(f"world {world}: ok")


if __name__ == "__main__":
    main()
