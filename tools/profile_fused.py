"""Developer tool: run the config-2 query a few times in-process (for ncu).  Usage: profile_fused.py [codec] [nfiles] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

codec = sys.argv[1] if len(sys.argv) > 1 else "none"
nfiles = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ssts = bench.gen_ssts(0, codec, nfiles, min(os.cpu_count(), 16))
from horaedb_b200 import sstgen  # noqa: E402
from horaedb_b200._ffi import Engine, SchemaHandle, SstInput  # noqa: E402

schema = sstgen.metric_storage_schema()
handle = SchemaHandle(schema.arrow_schema, 2)
eng = Engine(device=0)
for sid, data, n in ssts:
    eng.load_sst(handle, SstInput(id=sid, data=data, num_rows=n))
res = [SstInput(id=sid, num_rows=n) for sid, _, n in ssts]
for _ in range(reps):
    eng.scan_aggregate_device(handle, res, bench.preds(), group_col=0, ts_col=-1, window_ms=0, value_col=2)
    print(eng.stats())
eng.close()
