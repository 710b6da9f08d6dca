"""Developer tool: run the config-2 query a few times in-process (for ncu / HORAE_TRACE).
Usage: profile_fused.py [codec=none] [nfiles=16] [reps=4] [mode=resident|e2e]
  resident: SSTs loaded into HBM once, then `reps` scans;  e2e: every scan gets pinned HOST buffers (transient loads over PCIe)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

codec = sys.argv[1] if len(sys.argv) > 1 else "none"
nfiles = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
mode = sys.argv[4] if len(sys.argv) > 4 else "resident"
ssts = bench.gen_ssts(0, codec, nfiles, min(os.cpu_count(), 16))
from horaedb_b200 import sstgen  # noqa: E402
from horaedb_b200._ffi import Engine, SchemaHandle, SstInput  # noqa: E402

schema = sstgen.metric_storage_schema()
handle = SchemaHandle(schema.arrow_schema, 2)
eng = Engine(device=0)
if mode == "e2e":
    import numpy as np
    import torch
    pinned, inputs = [], []
    for sid, data, n in ssts:
        t = torch.empty(len(data), dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = np.frombuffer(data, dtype=np.uint8)
        pinned.append(t)
        inputs.append(SstInput(id=sid, ptr=t.data_ptr(), size=len(data), num_rows=n))
    for _ in range(reps):
        t0 = time.perf_counter()
        tbl = eng.scan_aggregate(handle, inputs, bench.preds(), group_col=0, ts_col=-1, window_ms=0, value_col=2)
        dt = time.perf_counter() - t0
        st = eng.stats()
        print(f"e2e {dt * 1e3:.2f} ms, groups {tbl.num_rows}, h2d {st['bytes_h2d'] / 1e6:.1f} MB, gpu_ms {st['gpu_ms']:.2f}, decomp_ms {st['decomp_ms']:.2f}")
else:
    for sid, data, n in ssts:
        eng.load_sst(handle, SstInput(id=sid, data=data, num_rows=n))
    res = [SstInput(id=sid, num_rows=n) for sid, _, n in ssts]
    for _ in range(reps):
        eng.scan_aggregate_device(handle, res, bench.preds(), group_col=0, ts_col=-1, window_ms=0, value_col=2)
        print(eng.stats())
eng.close()
