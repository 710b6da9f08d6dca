"""Config-5 style measurement: k overlapping SSTs of one segment -> one sorted, deduplicated run
(hg_compact_open = Executor::do_compaction's plan, keep_builtin).  Prints one JSON line.

Usage: bench_compaction.py [k=16] [series=4000] [points=1000] [keep=0.5] [codec=snappy]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from horaedb_b200 import sstgen  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
series = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
points = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
keep = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
codec = sys.argv[5] if len(sys.argv) > 5 else "snappy"
ssts = sstgen.synth_overlapping_ssts(k, series, points, 1000, keep, compression=codec)

import numpy as np  # noqa: E402
from horaedb_b200._ffi import Engine, SchemaHandle, SstInput  # noqa: E402

schema = sstgen.metric_storage_schema()
handle = SchemaHandle(schema.arrow_schema, 2)
eng = Engine(device=0)
inputs = []
for i, (data, n, seq) in enumerate(ssts):
    eng.load_sst(handle, SstInput(id=seq, data=data, num_rows=n))
    inputs.append(SstInput(id=seq, num_rows=n))
rows_in = sum(n for _, n, _ in ssts)
res = []
for it in range(5):
    t = time.perf_counter()
    out = eng.compact(handle, inputs).read_all()
    wall = time.perf_counter() - t
    st = eng.stats()
    res.append((st["merge_ms"], st["kernel_ms"], st["gpu_ms"], wall * 1e3))
m = np.median(np.array(res[1:]), axis=0)
key = list(zip(out["series_id"].to_pylist()[:200000], out["ts"].to_pylist()[:200000]))
assert key == sorted(set(key)), "output must be sorted and duplicate-free"
peak = 6578.0
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
alg = rows_in * 64
print(json.dumps({"workload": f"merge-compaction: {k} overlapping SSTs, {rows_in} rows in, {out.num_rows} rows out, codec {codec}",
                  "rows_in": rows_in, "rows_out": out.num_rows, "merge_ms": float(m[0]), "decode_ms": float(m[1]), "call_gpu_ms": float(m[2]),
                  "wall_ms": float(m[3]), "merge_rows_per_s": rows_in / (m[0] / 1e3), "call_rows_per_s": rows_in / (m[3] / 1e3),
                  "roofline_merge": {"alg_bytes": alg, "achieved_GBps": alg / (m[0] / 1e3) / 1e9, "peak_GBps": peak,
                                     "frac": alg / (m[0] / 1e3) / 1e9 / peak},
                  "kernel_launches": st["kernel_launches"]}))
eng.close()
