"""TEST INFRASTRUCTURE: damaged SSTs through the WHOLE library on the emulated build (cuda_emu.h), with guard pages behind every device
allocation (HORAE_EMU_GUARD) and cudaMalloc'ed memory filled with 0xCD.  Every call must end in a result or an HgError; a kernel that
reads or writes outside its buffers because a footer, a page header or page bytes lie is a segfault here (reported with kernel, block and
thread) — on the GPU it would be silent corruption or a sticky illegal-address error.

    python tests/emu/fuzz_engine.py SEED ITERATIONS [KIND ...]      KIND: metric-snappy metric-none metric-zstd dict delta nulls pages binary binary-append

Damage: 1-3 places with a byte overwritten / a bit flipped / a short range zeroed, set to 0xff or copied from elsewhere, 40 % of them in the footer (statistics, sizes, offsets, encodings), the rest anywhere in the
page area (page headers, level runs, compressed streams, dictionary indices, delta headers, values -> rows that contradict their chunk
statistics and the sort order).  Operations: aggregate on the fused path, aggregate on the general pipeline, scan, scan with predicates,
merge-compaction to a stream and to an SST written on the device."""
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
os.environ.setdefault("HORAE_EMU_GUARD", "1")
os.environ.setdefault("HORAE_EMU_CRASH_REPORT", "1")

import numpy as np  # noqa: E402
import pyarrow as pa  # noqa: E402

import build_engine_emu  # noqa: E402
from horaedb_b200 import _ffi  # noqa: E402

_ffi.LIB_PATH = build_engine_emu.build()
_ffi._lib = None

from horaedb_b200 import sstgen  # noqa: E402
from horaedb_b200._ffi import HG_FLAG_NO_FUSED, Engine, HgError, SchemaHandle, SstInput  # noqa: E402
from horaedb_b200.config import ColumnOptions, WriteConfig  # noqa: E402
from horaedb_b200.types import StorageSchema  # noqa: E402


def _nulls(rng, a, p, t):
    return pa.array([None if rng.random() < p else x for x in a], t)


def make_case(kind, rng):
    """-> (schema handle, [base SST bytes, partner SST bytes], predicates, aggregate kwargs)"""
    if kind.startswith("metric-"):
        codec = kind.split("-")[1]
        schema = sstgen.metric_storage_schema()
        a, _ = sstgen.synth_sst(0, 8, 300, 1000, seq=7, compression=codec)
        b, _ = sstgen.synth_sst(4, 12, 300, 1000, seq=8, compression=codec)
        return (SchemaHandle(schema.arrow_schema, 2), [a, b], [("tag", "eq", 3), ("ts", "ge", sstgen.T0_MS + 100_000)],
                dict(group_col=0, ts_col=1, window_ms=60_000, value_col=2))
    n = 4000
    if kind == "dict":
        user = pa.schema([pa.field("k", pa.uint64()), pa.field("t", pa.int64()), pa.field("lowcard", pa.int32()), pa.field("f", pa.float64()), pa.field("u", pa.uint32())])
        cols = [pa.array(np.arange(n, dtype=np.uint64) // 5), pa.array(np.arange(n, dtype=np.int64) * 1000),
                _nulls(rng, rng.integers(-3, 4, n).tolist(), 0.1, pa.int32()), _nulls(rng, rng.choice([0.5, -1.25, 3.0, 1e10], n).tolist(), 0.2, pa.float64()),
                pa.array(np.repeat(rng.integers(0, 50, n // 100 + 1), 100)[:n].astype(np.uint32))]
        cfg = [WriteConfig(compression=c, max_row_group_size=1500, enable_dict=True) for c in ("snappy", "none")]
        preds, kw = [("lowcard", "ge", 0)], dict(group_col=0, ts_col=1, window_ms=600_000, value_col=3)
    elif kind == "delta":
        user = pa.schema([pa.field("k", pa.uint64()), pa.field("t", pa.int64()), pa.field("i32", pa.int32()), pa.field("u32", pa.uint32()), pa.field("v", pa.float64())])
        cols = [pa.array(np.arange(n, dtype=np.uint64) // 7), pa.array(np.arange(n, dtype=np.int64) * 1000 - 500_000 + rng.integers(0, 300, n)),
                _nulls(rng, [int(x) for x in rng.integers(-2**31, 2**31, n)], 0.1, pa.int32()), _nulls(rng, [int(x) for x in rng.integers(0, 2**32, n)], 0.3, pa.uint32()),
                pa.array(rng.random(n))]
        opts = {c: ColumnOptions(encoding="DELTA_BINARY_PACKED") for c in ("k", "t", "i32", "u32", "__seq__")}
        cfg = [WriteConfig(compression=c, max_row_group_size=1500, column_options=opts) for c in ("snappy", "none")]
        preds, kw = [("t", "ge", 0), ("u32", "lt", 2**31)], dict(group_col=0, ts_col=1, window_ms=600_000, value_col=4)
    elif kind == "nulls":
        user = pa.schema([pa.field("k", pa.uint32()), pa.field("t", pa.int64()), pa.field("a", pa.int64()), pa.field("b", pa.float32()), pa.field("c", pa.float64())])
        cols = [pa.array((np.arange(n) // 9).astype(np.uint32)), pa.array(np.arange(n, dtype=np.int64) * 10),
                _nulls(rng, [int(x) for x in rng.integers(-2**40, 2**40, n)], 0.3, pa.int64()), _nulls(rng, rng.random(n).astype(np.float32).tolist(), 0.5, pa.float32()),
                _nulls(rng, rng.random(n).tolist(), 0.05, pa.float64())]
        cfg = [WriteConfig(compression=c, max_row_group_size=700) for c in ("snappy", "zstd")]
        preds, kw = [("a", "gt", 0)], dict(group_col=0, ts_col=1, window_ms=5000, value_col=4)
    elif kind == "pages":
        # many small data pages per chunk, page versions 1.0 and 2.0 (levels outside the compressed part), NULLs: written with pyarrow directly
        import io
        import pyarrow.parquet as pq
        user = pa.schema([pa.field("k", pa.uint64()), pa.field("t", pa.int64()), pa.field("a", pa.int32()), pa.field("c", pa.float64())])
        schema = StorageSchema.try_new(user, 2)
        cols = [pa.array(np.arange(n, dtype=np.uint64) // 6), pa.array(np.arange(n, dtype=np.int64) * 100),
                _nulls(rng, [int(x) for x in rng.integers(-50, 50, n)], 0.2, pa.int32()), _nulls(rng, rng.random(n).tolist(), 0.1, pa.float64())]
        files = []
        for i, (ver, comp) in enumerate((("2.0", "snappy"), ("1.0", "none"), ("2.0", "zstd"))):
            full = schema.fill_builtin_columns(pa.RecordBatch.from_arrays(cols, schema=user), 40 + i)
            sink = io.BytesIO()
            pq.write_table(pa.Table.from_batches([full]), sink, row_group_size=1500, compression=comp, use_dictionary=False, data_page_size=700, data_page_version=ver)
            files.append(sink.getvalue())
        return SchemaHandle(schema.arrow_schema, 2), files, [("a", "ge", 0)], dict(group_col=0, ts_col=1, window_ms=5000, value_col=3)
    elif kind in ("binary", "binary-append"):
        # Binary value columns (BYTE_ARRAY PLAIN / DELTA_LENGTH_BYTE_ARRAY), LastValue or BytesMerge (operator.rs:47-111)
        from horaedb_b200.types import UpdateMode
        mode = UpdateMode.Append if kind == "binary-append" else UpdateMode.Overwrite
        user = pa.schema([pa.field("pk1", pa.uint64()), pa.field("pk2", pa.int32()), pa.field("blob", pa.binary()), pa.field("idx", pa.binary())])
        n = 1500
        files = []
        for i, (codec, enc) in enumerate((("snappy", None), ("none", "DELTA_LENGTH_BYTE_ARRAY"), ("zstd", None))):
            pk1 = np.unique(rng.integers(0, 4000, n))
            cols = [pa.array(pk1.astype(np.uint64)), pa.array((pk1 % 5 - 2).astype(np.int32)),
                    pa.array([None if (int(k) + i) % 7 == 0 else rng.bytes(int(rng.integers(1, 40))) for k in pk1], pa.binary()),
                    pa.array([bytes([i + 1]) * int(rng.integers(1, 5)) for _ in pk1], pa.binary())]
            opts = {c: ColumnOptions(encoding=enc) for c in ("blob", "idx")} if enc else {}
            sch = StorageSchema.try_new(user, 2, mode)
            files.append(sstgen.write_sst(sch, pa.RecordBatch.from_arrays(cols, schema=user), 30 + i,
                                          WriteConfig(compression=codec, max_row_group_size=600, column_options=opts), presorted=True))
        return SchemaHandle(sch.arrow_schema, 2, mode), files, [("pk2", "ge", 0)], None
    else:
        raise SystemExit("unknown kind " + kind)
    schema = StorageSchema.try_new(user, 2)
    batch = pa.RecordBatch.from_arrays(cols, schema=user)
    files = [sstgen.write_sst(schema, batch, 20 + i, c, presorted=True) for i, c in enumerate(cfg)]
    return SchemaHandle(schema.arrow_schema, 2), files, preds, kw


def damage(rng, data):
    b = bytearray(data)
    flen = int.from_bytes(data[-8:-4], "little")
    for _ in range(int(rng.integers(1, 4))):
        p = len(b) - 8 - flen + int(rng.integers(0, flen)) if rng.random() < 0.4 else int(rng.integers(4, len(b) - 8 - flen))
        how = rng.random()
        if how < 0.4:
            b[p] = int(rng.integers(0, 256))
        elif how < 0.8:
            b[p] ^= 1 << int(rng.integers(0, 8))
        else:
            # a short range: zeroed, set to 0xff (huge varints / lengths / offsets), or overwritten with bytes from elsewhere in the file
            n = min(int(rng.integers(2, 48)), len(b) - 8 - p)
            kind = int(rng.integers(0, 3))
            if kind == 2:
                q = int(rng.integers(4, len(b) - 8 - n))
                b[p:p + n] = b[q:q + n]
            else:
                b[p:p + n] = bytes([0x00 if kind == 0 else 0xFF]) * n
    return bytes(b)


def main():
    seed, iters = int(sys.argv[1]), int(sys.argv[2])
    kinds = sys.argv[3:] or ["metric-snappy", "metric-none", "metric-zstd", "dict", "delta", "nulls", "pages", "binary", "binary-append"]
    rng = np.random.default_rng(seed)
    eng = Engine(device=0)
    tmp = tempfile.mkdtemp(prefix="horae_fuzz_")
    accepted = rejected = 0
    for kind in kinds:
        handle, files, preds, kw = make_case(kind, rng)
        for it in range(iters):
            which = it % len(files)
            bad = damage(rng, files[which])
            with open(os.path.join(tmp, "current.sst"), "wb") as f:          # the input of a crash stays on disk
                f.write(bad)
            ins = [SstInput(id=10_000 + it, data=bad), SstInput(id=5, data=files[(which + 1) % len(files)])]
            op = it % 6
            if kw is None and op in (0, 1, 5):                   # Binary tables: no aggregation; the device writer takes fixed-width columns
                op = 2 + it % 3
            try:
                eng.set_flags(HG_FLAG_NO_FUSED if op == 1 else 0)
                if op in (0, 1):
                    eng.scan_aggregate(handle, ins, preds, **kw)
                elif op == 2:
                    eng.scan(handle, ins, preds).read_all()
                elif op == 3:
                    eng.scan(handle, ins[:1], [], None, True).read_all()
                elif op == 4:
                    eng.compact(handle, ins).read_all()
                else:
                    eng.compact_to_sst(handle, ins, os.path.join(tmp, "out.sst"), max_row_group_size=1000)
                accepted += 1
            except HgError:
                rejected += 1
        print(f"{kind}: done ({iters} files)", flush=True)
    print(f"accepted {accepted} rejected {rejected}")


if __name__ == "__main__":
    main()
