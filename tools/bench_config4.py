"""Config-4 style measurement: one GPU's share of "256 SST files, 10 M series / 1 B points" — 32 SSTs x 39 062 series x 100
points = 125 M rows — with config 2's query shape (tag equality + time range, sum(value), count per series).  Series are short (100
rows), so every 8192-row group holds ~5 series that pass the tag predicate: the row-group gate drops nothing and the late-materialising
kernel touches most 64-row blocks — the opposite regime of config 2.  Prints one JSON line per codec.
Usage: bench_config4.py [files=32] [series_per_file=39062] [points=100] [codecs=snappy,none]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

nfiles = int(sys.argv[1]) if len(sys.argv) > 1 else 32
spf = int(sys.argv[2]) if len(sys.argv) > 2 else 39062
points = int(sys.argv[3]) if len(sys.argv) > 3 else 100
codecs = (sys.argv[4] if len(sys.argv) > 4 else "snappy,none").split(",")
bench.SERIES_PER_FILE, bench.POINTS, bench.DELTA_MS = spf, points, 10_000

import numpy as np  # noqa: E402
from horaedb_b200 import sstgen  # noqa: E402
from horaedb_b200._ffi import Engine, SchemaHandle, SstInput  # noqa: E402
from oracle import oracle  # noqa: E402  (inline parity of the measured query on two of the files)

schema = sstgen.metric_storage_schema()
handle = SchemaHandle(schema.arrow_schema, 2)
t0 = sstgen.T0_MS
P = [("tag", "eq", 3), ("ts", "ge", t0 + 250_000), ("ts", "lt", t0 + 750_000)]
peak = 6578.0
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
for codec in codecs:
    ssts = bench.gen_ssts(0, codec, nfiles, min(os.cpu_count(), 16))
    eng = Engine(device=0)
    for sid, data, n in ssts:
        eng.load_sst(handle, SstInput(id=sid, data=data, num_rows=n))
    res = [SstInput(id=sid, num_rows=n) for sid, _, n in ssts]
    rows = sum(n for _, _, n in ssts)
    ks, gs, ds = [], [], []
    for it in range(8):
        eng.scan_aggregate_device(handle, res, P, group_col=0, ts_col=-1, window_ms=0, value_col=2)
        st = eng.stats()
        ks.append(st["kernel_ms"]); gs.append(st["gpu_ms"]); ds.append(st["decomp_ms"])
    # parity on a sample: the first two files through the host-result entry point against the oracle
    tbl = eng.scan_aggregate(handle, [SstInput(id=900_000 + i, data=ssts[i][1], num_rows=ssts[i][2]) for i in range(2)], P, group_col=0, ts_col=-1,
                             window_ms=0, value_col=2)
    ex = oracle.scan_aggregate([ssts[i][1] for i in range(2)], schema.arrow_schema, 2, P, group_col=0, ts_col=-1, window_ms=0, value_col=2)
    ok = tbl["series_id"].to_numpy().tolist() == ex.gkey.tolist() and tbl["count"].to_numpy().tolist() == ex.count.tolist() and \
        bool(np.array_equal(tbl["sum"].to_numpy(), ex.sum))
    g, k, d = float(np.median(gs[2:])), float(np.median(ks[2:])), float(np.median(ds[2:]))
    print(json.dumps({"workload": f"config-4 share of one GPU: {nfiles} SSTs x {spf} series x {points} pts = {rows} rows, tag = 3 AND ts range, sum/count per series, codec {codec}",
                      "rows": rows, "rows_decoded": st["rows_decoded"], "rows_materialized": st["rows_materialized"], "rows_filtered": st["rows_filtered"],
                      "groups": st["groups_out"], "path": "fused" if st["path"] == 1 else "general", "call_gpu_ms": g, "scan_kernel_ms": k, "decomp_ms": d,
                      "rows_per_s": rows / (g / 1e3), "decoded_GBps_28B": st["rows_decoded"] * 28 / (g / 1e3) / 1e9, "frac_of_peak_28B": st["rows_decoded"] * 28 / (g / 1e3) / 1e9 / peak,
                      "parity_sample_ok": ok}), flush=True)
    eng.close()
