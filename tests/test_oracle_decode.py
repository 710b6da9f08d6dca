"""CPU tests: the oracle's Parquet restatement (S1/S2) against pyarrow 24 — an independent implementation of the
same format — and its filter / pruning / aggregation arithmetic against numpy."""
import io

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pyarrow.parquet as pq
import pytest

from helpers import arrow_schema, record_batch
from horaedb_b200 import sstgen
from horaedb_b200.config import ColumnOptions, ParquetCompression, WriteConfig
from horaedb_b200.types import StorageSchema
from oracle import oracle


def _random_batch(rng, n, null_frac):
    sch = arrow_schema([("k1", "uint64"), ("k2", "int64"), ("a", "float64"), ("b", "uint32"), ("c", "int8"),
                        ("d", "uint8"), ("e", "int32"), ("f", "float32"), ("g", "int16"), ("h", "uint16")])
    k1 = np.sort(rng.integers(0, max(n // 7, 1), n).astype(np.uint64))
    k2 = np.arange(n, dtype=np.int64) - n // 2
    cols = [pa.array(k1), pa.array(k2)]
    gens = [lambda: rng.standard_normal(n), lambda: rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32),
            lambda: rng.integers(-128, 128, n).astype(np.int8), lambda: rng.integers(0, 256, n).astype(np.uint8),
            lambda: rng.integers(-2**31, 2**31, n).astype(np.int32), lambda: rng.standard_normal(n).astype(np.float32),
            lambda: rng.integers(-2**15, 2**15, n).astype(np.int16), lambda: rng.integers(0, 2**16, n).astype(np.uint16)]
    for g in gens:
        v = g()
        mask = rng.random(n) < null_frac
        cols.append(pa.array(v, mask=mask))
    return sch, pa.RecordBatch.from_arrays(cols, schema=sch)


@pytest.mark.parametrize("compression", [ParquetCompression.Snappy, ParquetCompression.Uncompressed, ParquetCompression.Zstd])
@pytest.mark.parametrize("n,null_frac,rg", [(0, 0.0, 8192), (1, 0.0, 8192), (5, 0.5, 2), (1000, 0.0, 8192), (20000, 0.3, 8192),
                                            (20000, 1.0, 8192), (70000, 0.01, 8192), (9000, 0.9, 100)])
def test_decode_matches_pyarrow(compression, n, null_frac, rg):
    rng = np.random.default_rng(n + int(null_frac * 100))
    user, batch = _random_batch(rng, n, null_frac)
    schema = StorageSchema.try_new(user, 2)
    data = sstgen.write_sst(schema, batch, seq=77, cfg=WriteConfig(compression=compression, max_row_group_size=rg),
                            presorted=True)
    got = oracle.decode_sst(data, schema.arrow_schema)
    ref = pq.read_table(io.BytesIO(data)).cast(schema.arrow_schema)
    assert got.num_rows == n
    assert got.equals(ref)


def test_decode_small_pages_and_v2():
    """Multi-page chunks (tiny data_page_size) and DataPage V2 — page counts must come from headers (SURVEY §8 caveat)."""
    rng = np.random.default_rng(3)
    user, batch = _random_batch(rng, 30000, 0.2)
    schema = StorageSchema.try_new(user, 2)
    full = schema.fill_builtin_columns(batch, 5)
    for version in ("1.0", "2.0"):
        for comp in ("snappy", "none", "zstd"):
            sink = io.BytesIO()
            pq.write_table(pa.Table.from_batches([full]), sink, row_group_size=8192, compression=comp, use_dictionary=False,
                           data_page_size=3000, data_page_version=version)
            data = sink.getvalue()
            got = oracle.decode_sst(data, schema.arrow_schema)
            assert got.equals(pq.read_table(io.BytesIO(data)).cast(schema.arrow_schema)), (version, comp)


def test_filter_and_pruning_match_numpy():
    schema = sstgen.metric_storage_schema()
    data, n = sstgen.synth_sst(0, 32, 2000, 1000, seq=9)
    tbl = pq.read_table(io.BytesIO(data))
    t_lo, t_hi = sstgen.T0_MS + 500 * 1000, sstgen.T0_MS + 1500 * 1000
    preds = [("tag", "eq", 3), ("ts", "ge", t_lo), ("ts", "lt", t_hi)]
    mask = pc.and_(pc.and_(pc.equal(tbl["tag"], 3), pc.greater_equal(tbl["ts"], t_lo)), pc.less(tbl["ts"], t_hi))
    expect = tbl.filter(mask).select(["series_id", "ts", "value", "tag"])
    for prune in (False, True):
        res = oracle.scan([data], schema.arrow_schema, 2, preds=preds, prune=prune)
        got = pa.Table.from_batches(res.batches) if res.batches else expect.slice(0, 0)
        assert got.equals(expect.cast(got.schema))
        assert res.rows_in_files == n
        assert (res.rows_decoded < n) == prune  # stats pruning drops row groups whose tag range excludes 3
    for op, fn in (("ne", pc.not_equal), ("le", pc.less_equal), ("gt", pc.greater)):
        res = oracle.scan([data], schema.arrow_schema, 2, preds=[("value", op, 0.5)])
        assert sum(b.num_rows for b in res.batches) == pc.sum(fn(tbl["value"], 0.5)).as_py()


def test_null_predicate_is_false():
    user = arrow_schema([("pk", "int64"), ("v", "float64")])
    schema = StorageSchema.try_new(user, 1)
    b = pa.RecordBatch.from_arrays([pa.array([1, 2, 3, 4], pa.int64()), pa.array([1.0, None, 3.0, None])], schema=user)
    data = sstgen.write_sst(schema, b, seq=1)
    for op in ("eq", "ne", "lt", "ge"):
        res = oracle.scan([data], schema.arrow_schema, 1, preds=[("v", op, 2.0)])
        rows = pa.Table.from_batches(res.batches).to_pydict() if res.batches else {"pk": []}
        assert all(p in (1, 3) for p in rows["pk"])  # NULL => false (arrow filter semantics, read.rs:467-469)


def test_aggregate_sequential_sum_and_buckets():
    schema = sstgen.metric_storage_schema()
    ssts = [sstgen.synth_sst(lo, lo + 8, 300, 10_000, seq=20 + i)[0] for i, lo in enumerate((0, 8, 16))]
    w = 60_000
    res = oracle.scan_aggregate(ssts, schema.arrow_schema, 2, group_col=0, ts_col=1, window_ms=w, value_col=2)
    tbl = pa.concat_tables([pq.read_table(io.BytesIO(s)) for s in ssts])
    sid = tbl["series_id"].to_numpy(); ts = tbl["ts"].to_numpy(); v = tbl["value"].to_numpy()
    bucket = ts // w * w
    key_change = np.r_[True, (sid[1:] != sid[:-1]) | (bucket[1:] != bucket[:-1])]
    starts = np.flatnonzero(key_change)
    assert res.count.tolist() == np.diff(np.r_[starts, len(sid)]).tolist()
    assert res.gkey.tolist() == sid[starts].tolist() and res.bucket.tolist() == bucket[starts].tolist()
    ends = np.r_[starts[1:], len(sid)]
    for g in range(0, len(starts), 37):
        acc = 0.0
        for x in v[starts[g]:ends[g]]:
            acc += x  # sequential order
        assert res.sum[g] == acc
        assert res.min[g] == v[starts[g]:ends[g]].min() and res.max[g] == v[starts[g]:ends[g]].max()
    assert int(res.count.sum()) == len(sid)


def test_dedup_newest_seq_wins_across_files():
    ssts = sstgen.synth_overlapping_ssts(6, series=40, points=50, delta_ms=1000, keep_frac=0.5)
    schema = sstgen.metric_storage_schema()
    res = oracle.scan([s[0] for s in ssts], schema.arrow_schema, 2, keep_builtin=True, batch_size=256)
    out = pa.Table.from_batches(res.batches)
    full = pa.concat_tables([pq.read_table(io.BytesIO(s[0])) for s in ssts])
    exp = full.group_by(["series_id", "ts"]).aggregate([("__seq__", "max")]).sort_by([("series_id", "ascending"), ("ts", "ascending")])
    assert out.num_rows == exp.num_rows
    assert out["series_id"].to_pylist() == exp["series_id"].to_pylist()
    assert out["ts"].to_pylist() == exp["ts"].to_pylist()
    assert out["__seq__"].to_pylist() == exp["__seq___max"].to_pylist()
    assert res.rows_merged == full.num_rows


def test_delta_binary_packed_columns_match_pyarrow():
    """DELTA_BINARY_PACKED (config.rs:54-75: a per-column writer option) on INT64 / INT32-backed columns, with NULLs, Snappy and
    uncompressed pages, several pages per chunk: the oracle's decoder against pyarrow's reading of the same bytes."""
    import io

    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    from horaedb_b200 import sstgen
    from horaedb_b200._ffi import SchemaHandle, parquet_inspect, plan_row_groups
    from horaedb_b200.config import ColumnOptions, WriteConfig
    from horaedb_b200.types import StorageSchema
    from oracle import oracle
    rng = np.random.default_rng(8)
    n = 20_000
    user = pa.schema([pa.field("k", pa.uint64()), pa.field("t", pa.int64()), pa.field("i32", pa.int32()), pa.field("u32", pa.uint32()),
                      pa.field("i64", pa.int64()), pa.field("v", pa.float64())])
    schema = StorageSchema.try_new(user, 2)

    def nulls(a, p):
        return pa.array([None if rng.random() < p else int(x) for x in a], type=None)

    batch = pa.RecordBatch.from_arrays(
        [pa.array(np.arange(n, dtype=np.uint64) // 7), pa.array(np.arange(n, dtype=np.int64) * 1000 - 5_000_000 + rng.integers(0, 300, n)),
         pa.array(nulls(rng.integers(-2**31, 2**31, n), 0.1), pa.int32()), pa.array(nulls(rng.integers(0, 2**32, n), 0.3), pa.uint32()),
         pa.array(nulls(rng.integers(-2**62, 2**62, n), 0.02), pa.int64()), pa.array(rng.random(n))], schema=user)
    opts = {c: ColumnOptions(encoding="DELTA_BINARY_PACKED") for c in ("k", "t", "i32", "u32", "i64", "__seq__")}
    for comp in ("snappy", "none"):
        for rg in (8192, 3000):
            data = sstgen.write_sst(schema, batch, 9, WriteConfig(compression=comp, max_row_group_size=rg, column_options=opts), presorted=True)
            md = pq.ParquetFile(io.BytesIO(data)).metadata
            assert "DELTA_BINARY_PACKED" in md.row_group(0).column(1).encodings
            got = oracle.decode_sst(data, schema.arrow_schema)
            ref = pq.read_table(io.BytesIO(data))
            for c in ref.schema.names:
                assert got[c].combine_chunks().equals(ref[c].combine_chunks()), (comp, rg, c)
            # the library's host-side reader accepts the file (planning facts only; decoding is GPU work)
            assert parquet_inspect(data)["num_rows"] == n
            keep = plan_row_groups(SchemaHandle(schema.arrow_schema, 2), data, [("t", "lt", 0)])
            assert 0 < sum(keep) < len(keep)


def test_dictionary_encoded_columns_match_pyarrow():
    """RLE_DICTIONARY (`enable_dict`, config.rs:98-103,127): the oracle's dictionary-page + index-run decoder against pyarrow, and the
    library's host-side reader accepting such files."""
    import io

    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    from horaedb_b200 import sstgen
    from horaedb_b200._ffi import parquet_inspect
    from horaedb_b200.config import WriteConfig
    from horaedb_b200.types import StorageSchema
    from oracle import oracle
    rng = np.random.default_rng(2)
    n = 25_000
    user = pa.schema([pa.field("k", pa.uint64()), pa.field("t", pa.int64()), pa.field("c", pa.int32()), pa.field("f", pa.float32()),
                      pa.field("w", pa.int64())])
    schema = StorageSchema.try_new(user, 2)
    batch = pa.RecordBatch.from_arrays(
        [pa.array(np.arange(n, dtype=np.uint64) // 3), pa.array(np.arange(n, dtype=np.int64)),
         pa.array([None if rng.random() < 0.15 else int(x) for x in rng.integers(-5, 5, n)], pa.int32()),
         pa.array(rng.choice([0.5, -2.0, 7.25], n).astype(np.float32)), pa.array(rng.integers(-2**60, 2**60, n))], schema=user)
    for comp in ("snappy", "none"):
        for rg in (8192, 1111):
            data = sstgen.write_sst(schema, batch, 3, WriteConfig(compression=comp, max_row_group_size=rg, enable_dict=True), presorted=True)
            md = pq.ParquetFile(io.BytesIO(data)).metadata
            assert md.row_group(0).column(2).has_dictionary_page
            got = oracle.decode_sst(data, schema.arrow_schema)
            ref = pq.read_table(io.BytesIO(data))
            for c in ref.schema.names:
                assert got[c].combine_chunks().equals(ref[c].combine_chunks()), (comp, rg, c)
            assert parquet_inspect(data)["num_rows"] == n


@pytest.mark.parametrize("level", [1, 3, 9])
def test_zstd_restatement_matches_libzstd(level):
    """oracle/zstd_oracle.h (sequential restatement of RFC 8878) against libzstd itself (pyarrow's codec): frames of every shape the
    device decoder is tested on — raw / RLE / Huffman literals in one and four streams, predefined / RLE / described / repeated
    sequence tables, repeat offsets, overlapping matches, several blocks."""
    import ctypes as C
    L = oracle.lib()
    L.orc_decompress.argtypes = [C.c_int, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64]
    L.orc_decompress.restype = C.c_int
    rng = np.random.default_rng(5)
    n = 8192
    ts = (1_700_000_000_000 + np.tile(np.arange(1000), 9)[:n].astype(np.int64) * 1000 + rng.integers(0, 500, n)).astype(np.int64)
    cases = {
        "empty": b"", "three_bytes": b"abc", "short_run": b"a" * 70,
        "jitter_ts": b"\x02\x00\x00\x00\x03\x10" + ts.tobytes(),
        "series_id": np.repeat(np.arange(9, dtype=np.uint64) + 77, 1000)[:n].tobytes(),
        "random": rng.integers(0, 2**63, n, dtype=np.uint64).tobytes(),
        "sawtooth": ((np.arange(n, dtype=np.uint64) % 1000) * 37).tobytes(),
        "slow_counter": (np.arange(n, dtype=np.uint64) // 3).tobytes(),
        "few_values": rng.choice(rng.integers(0, 2**60, 5, dtype=np.uint64), n).tobytes(),
        "small_u32": rng.integers(0, 16, n).astype(np.uint32).tobytes(),
        "text": (b"the quick brown fox jumps over the lazy dog. " * 500)[:20000],
        "f64_round": np.round(rng.random(n), 2).tobytes(),
        "skewed_bytes_300k": rng.integers(0, 50, 300_000).astype(np.uint8).tobytes(),
    }
    for name, raw in cases.items():
        comp = pa.Codec("zstd", compression_level=level).compress(raw, asbytes=True)
        out = np.zeros(len(raw) + 8, dtype=np.uint8)
        rc = L.orc_decompress(6, comp, len(comp), out.ctypes.data, len(raw))
        assert rc == 0, (name, L.orc_last_error())
        assert bytes(out[:len(raw)]) == raw, name
    comp = pa.Codec("zstd").compress(cases["jitter_ts"], asbytes=True)
    out = np.zeros(len(cases["jitter_ts"]) + 8, dtype=np.uint8)
    assert L.orc_decompress(6, comp[: len(comp) // 2], len(comp) // 2, out.ctypes.data, len(cases["jitter_ts"])) != 0       # truncated frame


def test_oracle_reads_zstd_ssts_like_their_snappy_twins():
    """The GPU parity tests of Zstd SSTs (tests/test_gpu_zstd.py) compare against the oracle's result on a Snappy TWIN of every file;
    here the oracle itself reads the Zstd files: scan (merge + dedup over overlapping files, batch boundaries) and aggregates are
    identical to the twins'."""
    rng = np.random.default_rng(3)
    schema = sstgen.metric_storage_schema()
    zs, ss = [], []
    for f in range(3):
        sid = np.repeat(np.arange(20 * f, 20 * f + 40), 600)
        ts = sstgen.T0_MS + np.tile(np.arange(600) * 1000 + rng.integers(0, 300, 600), 40)
        b = pa.RecordBatch.from_arrays([pa.array(sid.astype(np.uint64)), pa.array(ts.astype(np.int64)), pa.array(rng.random(len(sid))),
                                        pa.array((sid % 16).astype(np.uint32))], schema=sstgen.METRIC_SCHEMA)
        zs.append(sstgen.write_sst(schema, b, seq=700 + f, cfg=WriteConfig(compression=ParquetCompression.Zstd), presorted=True))
        ss.append(sstgen.write_sst(schema, b, seq=700 + f, cfg=WriteConfig(compression=ParquetCompression.Snappy), presorted=True))
    for preds in ([], [("tag", "eq", 3)], [("ts", "ge", sstgen.T0_MS + 100_000), ("ts", "lt", sstgen.T0_MS + 400_000)]):
        a = oracle.scan(zs, schema.arrow_schema, 2, preds).batches
        b = oracle.scan(ss, schema.arrow_schema, 2, preds).batches
        assert len(a) == len(b) and all(x.equals(y) for x, y in zip(a, b))
        kw = dict(group_col=0, ts_col=1, window_ms=60_000, value_col=2)
        x, y = oracle.scan_aggregate(zs, schema.arrow_schema, 2, preds, **kw), oracle.scan_aggregate(ss, schema.arrow_schema, 2, preds, **kw)
        assert x.gkey.tolist() == y.gkey.tolist() and x.count.tolist() == y.count.tolist() and np.array_equal(x.sum, y.sum)
