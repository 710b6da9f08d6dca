"""Config-5 measurement: k overlapping SSTs of one segment -> one sorted, deduplicated run
(hg_compact_open = Executor::do_compaction's plan, keep_builtin), and the same through the GPU SST writer (hg_compact_to_sst).
Prints one JSON line.

Usage: bench_compaction.py [k=16] [series=4000] [points=1000] [keep=0.5] [codec=snappy] [procs=16]
BASELINE config 5 at size: bench_compaction.py 64 15625 1000 0.25 snappy 32   (64 SSTs x 3.9 M rows = 250 M rows in)"""
import json
import os
import sys
import tempfile
import time
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
series = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
points = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
keep = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
codec = sys.argv[5] if len(sys.argv) > 5 else "snappy"
procs = int(sys.argv[6]) if len(sys.argv) > 6 else 16


def _one(f):
    from horaedb_b200 import sstgen
    return sstgen.synth_overlapping_ssts(1, series, points, 1000, keep, compression=codec, base_seq=1000 + f)[0]


if __name__ == "__main__":
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=procs) as ex:
        ssts = list(ex.map(_one, range(k)))
    gen_s = time.perf_counter() - t0

    import numpy as np
    from horaedb_b200 import sstgen
    from horaedb_b200._ffi import HG_FLAG_PAIRWISE_MERGE, Engine, SchemaHandle, SstInput

    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    inputs = []
    for i, (data, n, seq) in enumerate(ssts):
        eng.load_sst(handle, SstInput(id=seq, data=data, num_rows=n))
        inputs.append(SstInput(id=seq, num_rows=n, time_start=0, time_end=1, max_sequence=seq))
    rows_in = sum(n for _, n, _ in ssts)
    peak = 6578.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    def run(reps=4):
        res = []
        out = None
        for it in range(reps):
            t = time.perf_counter()
            out = eng.compact(handle, inputs).read_all()
            wall = time.perf_counter() - t
            st = eng.stats()
            res.append((st["merge_ms"], st["kernel_ms"], st["gpu_ms"], wall * 1e3))
        return np.median(np.array(res[1:]), axis=0), out, eng.stats()

    m, out, st = run()
    sid, ts = out["series_id"].to_numpy(), out["ts"].to_numpy()
    key = sid.astype(np.uint64) * np.uint64(1 << 32) + (ts - sstgen.T0_MS).astype(np.uint64)
    assert np.all(key[1:] > key[:-1]), "output must be sorted and duplicate-free"
    eng.set_flags(HG_FLAG_PAIRWISE_MERGE)
    mp, outp, _ = run(3)
    assert outp.num_rows == out.num_rows
    eng.set_flags(0)
    # end to end on the GPU: merge + dedup + Parquet encode (Snappy) + file written
    path = os.path.join(tempfile.mkdtemp(), "out.sst")
    wt = []
    for it in range(3):
        t = time.perf_counter()
        meta = eng.compact_to_sst(handle, inputs, path)
        wt.append(((time.perf_counter() - t) * 1e3, eng.stats()["gpu_ms"]))
    wt = np.median(np.array(wt[1:]), axis=0)
    alg = rows_in * 64
    print(json.dumps({"workload": f"merge-compaction: {k} overlapping SSTs, {rows_in} rows in, {out.num_rows} rows out, codec {codec}",
                      "rows_in": rows_in, "rows_out": out.num_rows, "merge_ms": float(m[0]), "decode_ms": float(m[1]), "call_gpu_ms": float(m[2]),
                      "wall_ms": float(m[3]), "merge_rows_per_s": rows_in / (m[0] / 1e3), "call_rows_per_s": rows_in / (m[3] / 1e3),
                      "roofline_merge": {"alg_bytes": alg, "bytes_model": "64 B per input row (SURVEY 8d: read 4 columns incl. __seq__ + write them once)",
                                         "achieved_GBps": alg / (m[0] / 1e3) / 1e9, "peak_GBps": peak, "frac": alg / (m[0] / 1e3) / 1e9 / peak},
                      "pairwise_passes": {"merge_ms": float(mp[0]), "frac": alg / (mp[0] / 1e3) / 1e9 / peak,
                                          "note": "HG_FLAG_PAIRWISE_MERGE: log2(k) merge-path passes over 32-byte records (round 1)"},
                      "compact_to_sst": {"wall_ms": float(wt[0]), "gpu_ms": float(wt[1]), "file_bytes": int(meta.size), "rows": int(meta.num_rows),
                                         "rows_in_per_s": rows_in / (wt[0] / 1e3)},
                      "kernel_launches": st["kernel_launches"], "generate_s": gen_s}))
    eng.close()
