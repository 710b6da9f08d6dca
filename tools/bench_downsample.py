"""Config-3 style measurement: 1-minute downsample sum/min/max/count per (series, bucket), no predicate.
Usage: bench_downsample.py [series_per_file=6250] [points=1000] [delta_ms=10000] [files=16] [codec=none]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

spf = int(sys.argv[1]) if len(sys.argv) > 1 else 6250
points = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
delta = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
nfiles = int(sys.argv[4]) if len(sys.argv) > 4 else 16
codec = sys.argv[5] if len(sys.argv) > 5 else "none"
bench.SERIES_PER_FILE, bench.POINTS, bench.DELTA_MS = spf, points, delta
ssts = bench.gen_ssts(0, codec, nfiles, min(os.cpu_count(), 16))

import numpy as np  # noqa: E402
from horaedb_b200 import sstgen  # noqa: E402
from horaedb_b200._ffi import Engine, SchemaHandle, SstInput  # noqa: E402

schema = sstgen.metric_storage_schema()
handle = SchemaHandle(schema.arrow_schema, 2)
eng = Engine(device=0)
for sid, data, n in ssts:
    eng.load_sst(handle, SstInput(id=sid, data=data, num_rows=n))
res = [SstInput(id=sid, num_rows=n) for sid, _, n in ssts]
rows = sum(n for _, _, n in ssts)
ks, gs = [], []
for it in range(6):
    eng.scan_aggregate_device(handle, res, [], group_col=0, ts_col=1, window_ms=60_000, value_col=2)
    st = eng.stats()
    ks.append(st["kernel_ms"]); gs.append(st["gpu_ms"])
peak = 6578.0
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
k, g = float(np.median(ks[1:])), float(np.median(gs[1:]))
alg = rows * 24 + st["groups_out"] * 32
print(json.dumps({"workload": f"1-min downsample sum/min/max/count: {nfiles} SSTs, {rows} rows, delta {delta} ms, codec {codec}",
                  "rows": rows, "groups": st["groups_out"], "path": "fused" if st["path"] == 1 else "general", "kernel_ms": k, "call_gpu_ms": g,
                  "rows_per_s_call": rows / (g / 1e3), "roofline_kernel": {"alg_bytes": alg, "achieved_GBps": alg / (k / 1e3) / 1e9, "frac": alg / (k / 1e3) / 1e9 / peak}}))
eng.close()
