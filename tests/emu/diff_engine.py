"""TEST INFRASTRUCTURE: randomized differential test — random VALID tables (schema, NULL rates, key density, duplicates inside and across
files), written with random writer settings (codec per file, row-group size, dictionaries, DELTA_BINARY_PACKED on integer columns), scanned
/ aggregated / compacted by the emulated build of the library (cuda_emu.h; guard pages on) and compared with the CPU oracle: rows, batch
boundaries, builtin columns, group keys, counts, f64 sums bit for bit.

    python tests/emu/diff_engine.py SEED CASES

The fixed GPU tests pick their shapes by hand; this walks the combinations nobody picked."""
import io
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("HORAE_EMU_GUARD", "1")
os.environ.setdefault("HORAE_EMU_CRASH_REPORT", "1")

import numpy as np  # noqa: E402
import pyarrow as pa  # noqa: E402
import pyarrow.parquet as pq  # noqa: E402

import build_engine_emu  # noqa: E402
from horaedb_b200 import _ffi  # noqa: E402

_ffi.LIB_PATH = build_engine_emu.build()
_ffi._lib = None

from helpers import check_stream  # noqa: E402
from horaedb_b200 import sstgen  # noqa: E402
from horaedb_b200._ffi import HG_AGG_HASH, HG_FLAG_NO_FUSED, HG_FLAG_NO_LATE_MATERIALIZATION, HG_FLAG_PAIRWISE_MERGE, Engine, SchemaHandle, SstInput  # noqa: E402
from horaedb_b200.config import ColumnOptions, WriteConfig  # noqa: E402
from horaedb_b200.types import StorageSchema  # noqa: E402
from oracle import oracle  # noqa: E402

VALUE_TYPES = [pa.int32(), pa.uint32(), pa.int64(), pa.uint64(), pa.float32(), pa.float64()]
PK0_TYPES = [pa.uint64(), pa.int64(), pa.uint32(), pa.int32()]


def rand_values(rng, t, n):
    if pa.types.is_floating(t):
        v = rng.choice([rng.random(n) * 100 - 50, np.round(rng.random(n) * 8) / 4, rng.integers(-5, 5, n).astype(np.float64)])
        return v.astype(np.float32 if t == pa.float32() else np.float64)
    lo, hi = {pa.int32(): (-2**31, 2**31), pa.uint32(): (0, 2**32), pa.int64(): (-2**62, 2**62), pa.uint64(): (0, 2**63)}[t]
    if rng.random() < 0.5:
        lo, hi = max(lo, -20), min(hi, 20)                       # few distinct values (dictionary / RLE friendly, predicate hits)
    return rng.integers(lo, hi, n).astype(t.to_pandas_dtype())


def make_case(rng):
    pk0_t = PK0_TYPES[int(rng.integers(0, len(PK0_TYPES)))]
    nval = int(rng.integers(1, 4))
    vts = [VALUE_TYPES[int(rng.integers(0, len(VALUE_TYPES)))] for _ in range(nval)]
    user = pa.schema([pa.field("k", pa_t, True) for pa_t in [pk0_t]] + [pa.field("t", pa.int64(), True)] + [pa.field(f"v{i}", t, True) for i, t in enumerate(vts)])
    schema = StorageSchema.try_new(user, 2)
    nfiles = int(rng.integers(1, 5))
    keyspace = int(rng.choice([3, 40, 400]))
    t_step = int(rng.choice([1, 1000, 60_000]))
    t_base = int(rng.choice([0, -500_000, 1_700_000_000_000]))
    files = []
    for f in range(nfiles):
        n = int(rng.choice([1, 37, 900, 5000]))
        k = np.sort(rng.integers(0, keyspace, n)) + (int(rng.integers(0, keyspace)) if rng.random() < 0.5 else 0)
        if pk0_t in (pa.int64(), pa.int32()):
            k = k - keyspace // 2
        t = t_base + rng.integers(0, max(2, n // max(1, keyspace) * 2 + 2), n) * t_step
        order = np.lexsort((t, k))                                # sorted by (k, t); duplicates stay (read.rs keeps the last per PK)
        k, t = k[order], t[order]
        cols = [pa.array(k.astype(pk0_t.to_pandas_dtype())), pa.array(t.astype(np.int64))]
        for vt in vts:
            v = rand_values(rng, vt, n)
            p_null = float(rng.choice([0.0, 0.0, 0.05, 0.6]))
            cols.append(pa.array(v, mask=(rng.random(n) < p_null)) if p_null else pa.array(v))
        opts = {}
        if rng.random() < 0.3:
            for name in ["k", "t", "__seq__"] + [f"v{i}" for i, vt in enumerate(vts) if pa.types.is_integer(vt)]:
                if rng.random() < 0.7:
                    opts[name] = ColumnOptions(encoding="DELTA_BINARY_PACKED")
        cfg = WriteConfig(compression=str(rng.choice(["snappy", "snappy", "none", "zstd"])), max_row_group_size=int(rng.choice([8192, 1000, 97])),
                          enable_dict=bool(rng.random() < 0.25) and not opts, column_options=opts or None)
        files.append(sstgen.write_sst(schema, pa.RecordBatch.from_arrays(cols, schema=user), 100 + f, cfg, presorted=True))
    preds = []
    for _ in range(int(rng.integers(0, 3))):
        c = int(rng.integers(0, 2 + nval))
        name = user.names[c]
        op = str(rng.choice(["eq", "ne", "lt", "le", "gt", "ge", "in"]))
        t = user.field(c).type
        if pa.types.is_floating(t):
            lit = float(rng.choice([0.0, 0.5, -1.0, 2.25, 50.0]))
        elif c == 1:
            lit = t_base + int(rng.integers(0, 50)) * t_step
        elif c == 0:
            lit = int(rng.integers(0, keyspace)) - (keyspace // 2 if pk0_t in (pa.int64(), pa.int32()) else 0)
        else:
            lit = int(rng.integers(-3, 10)) if pa.types.is_signed_integer(t) else int(rng.integers(0, 10))
        if op == "in":
            lit = [lit, lit + 1, lit + 3] if not isinstance(lit, float) else [lit, 0.25, -1.0]
        preds.append((name, op, lit))
    return schema, files, preds, nval, t_step


def _nan_equal(a, b):
    x, y = a.combine_chunks(), b.combine_chunks()
    if not pa.types.is_floating(x.type) or x.null_count != y.null_count:
        return False
    xv, yv = x.to_numpy(zero_copy_only=False), y.to_numpy(zero_copy_only=False)
    return np.array_equal(np.isnan(xv), np.isnan(yv)) and np.array_equal(xv[~np.isnan(xv)], yv[~np.isnan(yv)])


def main():
    tmpdir = tempfile.mkdtemp(prefix="horae_diff_")
    seed, cases = int(sys.argv[1]), int(sys.argv[2])
    rng = np.random.default_rng(seed)
    eng = Engine(device=0)
    next_id = [1]
    for case in range(cases):
        schema, files, preds, nval, t_step = make_case(rng)
        handle = SchemaHandle(schema.arrow_schema, 2)

        # half of the cases keep (some of) their files RESIDENT in device memory (hg_sst_load: the SST cache keyed by file id) and pass
        # them by id; the others hand host bytes to every call (transient loads: only the needed column chunks travel)
        resident = {}
        if rng.random() < 0.5:
            for i, d in enumerate(files):
                if rng.random() < 0.7:
                    resident[i] = next_id[0]
                    eng.load_sst(handle, SstInput(id=next_id[0], data=d))
                    next_id[0] += 1

        def ins():
            out = []
            for i, d in enumerate(files):
                if i in resident:
                    out.append(SstInput(id=resident[i]))
                else:
                    out.append(SstInput(id=next_id[0], data=d))
                    next_id[0] += 1
            return out

        flags = int(rng.choice([0, 0, HG_FLAG_NO_FUSED, HG_FLAG_NO_LATE_MATERIALIZATION, HG_FLAG_PAIRWISE_MERGE]))
        eng.set_flags(flags)
        tag = f"case {case} (seed {seed}): {len(files)} files ({len(resident)} resident), preds {preds}, flags {flags}"
        try:
            # scan: rows, batch boundaries, builtin columns
            keep = bool(rng.random() < 0.5)
            got = list(eng.scan(handle, ins(), preds, None, keep))
            exp = oracle.scan(files, schema.arrow_schema, 2, preds, keep, 8192).batches
            check_stream(got, exp)
            # aggregate: per key / per (key, bucket), value column = a random value column
            vc = 2 + int(rng.integers(0, nval))
            with_ts = bool(rng.random() < 0.6)
            kw = dict(group_col=0, ts_col=1 if with_ts else -1, window_ms=int(rng.choice([1, 7, 60, 3600])) * t_step if with_ts else 0, value_col=vc)
            a = eng.scan_aggregate(handle, ins(), preds, **kw)
            b = oracle.scan_aggregate(files, schema.arrow_schema, 2, preds, **kw)
            assert a.num_rows == len(b.gkey), (a.num_rows, len(b.gkey))
            assert a["count"].to_numpy().tolist() == b.count.tolist()
            if with_ts:
                assert a["bucket"].to_numpy().tolist() == b.bucket.tolist()
            assert [int(x) & 0xffffffffffffffff for x in a[a.schema.names[0]].to_pylist()] == [int(x) for x in b.gkey.tolist()]
            assert np.array_equal(a["sum"].to_numpy().view(np.uint64), b.sum.view(np.uint64))
            assert np.array_equal(a["min"].to_numpy().view(np.uint64), b.min.view(np.uint64)) and np.array_equal(a["max"].to_numpy().view(np.uint64), b.max.view(np.uint64))
            # GROUP BY a value column (not a prefix of the sort order): radix-partitioned aggregation, groups sorted by (key, bucket)
            gcol = 2 + int(rng.integers(0, nval))
            if not pa.types.is_floating(schema.arrow_schema.field(gcol).type) or rng.random() < 0.5:
                kwh = dict(group_col=gcol, ts_col=1 if with_ts else -1, window_ms=kw["window_ms"], value_col=vc)
                a = eng.scan_aggregate(handle, ins(), preds, mode=HG_AGG_HASH, **kwh)
                b = oracle.scan_aggregate(files, schema.arrow_schema, 2, preds, mode=1, **kwh)
                assert a.num_rows == len(b.count), (a.num_rows, len(b.count))
                assert a["count"].to_numpy().tolist() == b.count.tolist()
                assert np.array_equal(a["sum"].to_numpy().view(np.uint64), b.sum.view(np.uint64))
            # compaction = scan without predicates, builtin columns kept
            got = list(eng.compact(handle, ins()))
            exp = oracle.scan(files, schema.arrow_schema, 2, (), True, 8192).batches
            check_stream(got, exp)
            # ... and written as an SST on the device: pyarrow and the oracle read back the same rows
            if exp and rng.random() < 0.5:
                path = os.path.join(tmpdir, "out.sst")
                eng.compact_to_sst(handle, ins(), path, max_row_group_size=int(rng.choice([8192, 500])), compression=str(rng.choice(["snappy", "none"])))
                with open(path, "rb") as f:
                    written = f.read()
                want = pa.Table.from_batches(exp)
                back = pq.read_table(io.BytesIO(written))
                assert back.num_rows == want.num_rows
                for c in range(want.num_columns):
                    assert back.column(c).combine_chunks().equals(want.column(c).combine_chunks()) or \
                        back.column(c).to_pylist() == want.column(c).to_pylist() or _nan_equal(back.column(c), want.column(c)), (c, "device-written SST differs")
                again = pa.Table.from_batches(oracle.scan([written], schema.arrow_schema, 2, (), True, 8192).batches)     # (batch boundaries follow the new row groups)
                check_stream([again.combine_chunks().to_batches()[0]], [want.combine_chunks().to_batches()[0]])
            for rid in resident.values():
                eng.unload_sst(rid)
        except Exception:
            print("FAILED", tag, flush=True)
            for i, d in enumerate(files):
                with open(f"/tmp/diff_case_{i}.sst", "wb") as f:
                    f.write(d)
            raise
    print(f"{cases} cases ok")


if __name__ == "__main__":
    main()
