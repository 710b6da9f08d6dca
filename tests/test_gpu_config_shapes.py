"""GPU parity tests at the SHAPES of BASELINE.json's configs 3, 4 and 5, sized so that the CPU oracle still
finishes in seconds, plus the window arithmetic around bucket 0 (Timestamp::truncate_by, types.rs:82-85: truncating
division, bucket 0 spans (-w, w)) on both aggregators.  Everything goes through the C ABI and is compared bit for bit."""
import numpy as np
import pyarrow as pa
import pytest

from horaedb_b200 import sstgen
from horaedb_b200._ffi import HG_FLAG_NO_FUSED, HG_FLAG_PAIRWISE_MERGE, Engine, SchemaHandle, SstInput
from horaedb_b200.config import ParquetCompression, WriteConfig
from oracle import oracle

from helpers import check_stream

pytestmark = pytest.mark.gpu
_ids = iter(range(80_000_000, 90_000_000))


def _inputs(datas):
    return [SstInput(id=next(_ids), data=d) for d in datas]


def _metric_batch(sid, ts, value, tag):
    return pa.RecordBatch.from_arrays([pa.array(sid.astype(np.uint64)), pa.array(ts.astype(np.int64)), pa.array(value.astype(np.float64)),
                                       pa.array(tag.astype(np.uint32))], schema=sstgen.METRIC_SCHEMA)


def _check(got, exp, bucket=True):
    assert got.num_rows == len(exp.count)
    assert got["series_id"].to_numpy().tolist() == exp.gkey.tolist()
    if bucket:
        assert got["bucket"].to_numpy().tolist() == exp.bucket.tolist()
    assert got["count"].to_numpy().tolist() == exp.count.tolist()
    assert np.array_equal(got["sum"].to_numpy(), exp.sum) and np.array_equal(got["min"].to_numpy(), exp.min) and np.array_equal(got["max"].to_numpy(), exp.max)


@pytest.mark.parametrize("codec", ["snappy", "none"])
def test_config3_shape_one_minute_downsample(codec):
    """Config 3: 1-minute buckets over a dense series x time grid (groups ~ rows / 6: the group-dense aggregation), and its
    primary variant with ~1 row per bucket.  2 000 series x 5 000 points = 10 M rows."""
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    datas = [sstgen.synth_sst(500 * i, 500 * i + 500, 5000, 10_000, seq=100 + i, compression=codec)[0] for i in range(4)]
    for window, preds in ((60_000, []), (60_000, [("tag", "le", 7)]), (10_000, [])):
        kw = dict(group_col=0, ts_col=1, window_ms=window, value_col=2)
        got = eng.scan_aggregate(handle, _inputs(datas), preds, **kw)
        exp = oracle.scan_aggregate(datas, schema.arrow_schema, 2, preds, **kw)
        assert len(exp.count) > 800_000
        _check(got, exp)
    eng.close()


@pytest.mark.parametrize("codec", ["snappy", "none"])
def test_config4_shape_many_files_short_series(codec):
    """Config 4: many PK-disjoint files, 100-point series, predicate + per-series aggregate (a) and 1-minute buckets."""
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    nfiles, per = 64, 400                                      # 64 files x 400 series x 100 points = 2.56 M rows
    datas = [sstgen.synth_sst(per * i, per * (i + 1), 100, 1000, seq=1000 + i, compression=codec)[0] for i in range(nfiles)]
    t0 = sstgen.T0_MS
    preds = [("tag", "eq", 3), ("ts", "ge", t0 + 25_000), ("ts", "lt", t0 + 75_000)]
    for kw in (dict(group_col=0, ts_col=-1, window_ms=0, value_col=2), dict(group_col=0, ts_col=1, window_ms=60_000, value_col=2)):
        got = eng.scan_aggregate(handle, _inputs(datas), preds, **kw)
        assert eng.stats()["path"] == 1
        exp = oracle.scan_aggregate(datas, schema.arrow_schema, 2, preds, **kw)
        assert len(exp.count) >= nfiles * per // 16
        _check(got, exp, kw["ts_col"] >= 0)
    eng.close()


@pytest.mark.parametrize("codec", ["none", "snappy"])
def test_config5_shape_64_way_merge_compaction(codec):
    """Config 5: 64 overlapping SSTs of one segment, every file a 25 % sample of the same PK universe -> one sorted,
    deduplicated stream with builtin columns kept (Executor::do_compaction, executor.rs:155-222).  10 M rows in."""
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    ssts = sstgen.synth_overlapping_ssts(64, series=625, points=1000, delta_ms=1000, keep_frac=0.25, compression=codec)
    datas = [s[0] for s in ssts]
    rows_in = sum(s[1] for s in ssts)
    assert rows_in > 9_000_000
    exp = oracle.scan(datas, schema.arrow_schema, 2, (), True, 8192).batches
    eng = Engine(device=0)
    got = list(eng.compact(handle, _inputs(datas)))
    st = eng.stats()
    check_stream(got, exp)                                      # contents AND MergeStream batch boundaries
    out = pa.Table.from_batches(got)
    sid, ts, seq = out["series_id"].to_numpy(), out["ts"].to_numpy(), out["__seq__"].to_numpy()
    key = sid.astype(np.uint64) * np.uint64(1 << 32) + (ts - sstgen.T0_MS).astype(np.uint64)
    assert np.all(key[1:] > key[:-1])                           # sorted, one row per distinct PK
    assert st["rows_out"] == len(key) and st["rows_decoded"] == rows_in
    # every survivor carries the largest sequence among its versions: recompute from the inputs
    import io

    import pyarrow.parquet as pq
    best = {}
    for d in datas[:8]:                                         # a sample of the inputs is enough to catch a wrong winner
        t = pq.read_table(io.BytesIO(d), columns=["series_id", "ts", "__seq__"])
        k2 = t["series_id"].to_numpy().astype(np.uint64) * np.uint64(1 << 32) + (t["ts"].to_numpy() - sstgen.T0_MS).astype(np.uint64)
        for kk, s in zip(k2[:20000].tolist(), t["__seq__"].to_numpy()[:20000].tolist()):
            best[kk] = max(best.get(kk, 0), s)
    pos = np.searchsorted(key, np.array(list(best), dtype=np.uint64))
    assert np.all(seq[pos] >= np.array(list(best.values()), dtype=np.uint64))
    # A/B: the pairwise merge passes give the identical stream
    eng2 = Engine(device=0, flags=HG_FLAG_PAIRWISE_MERGE)
    got2 = list(eng2.compact(handle, _inputs(datas)))
    check_stream(got2, exp)
    eng2.close()
    eng.close()


def test_merge_packed_keys_edge_cases():
    """The packed-key single pass on awkward inputs: empty streams, one long + many tiny streams, all-equal keys across
    streams (ties -> lower stream index, i.e. the highest __seq__ wins only because it sorts last), a filter in front."""
    rng = np.random.default_rng(3)
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)

    def sst(sid, ts, seq, rg=8192):
        return sstgen.write_sst(schema, _metric_batch(sid, ts, rng.random(len(sid)), sid % 4), seq=seq,
                                cfg=WriteConfig(compression=ParquetCompression.Uncompressed, max_row_group_size=rg), presorted=True)

    base_sid = np.repeat(np.arange(200), 500)
    base_ts = sstgen.T0_MS + np.tile(np.arange(500) * 1000, 200)
    cases = []
    # identical PK sets in 5 files (every PK has 5 versions)
    cases.append([sst(base_sid, base_ts, 10 + f) for f in range(5)])
    # one long stream, many tiny ones, one empty
    tiny = [sst(base_sid[i * 997:i * 997 + 3], base_ts[i * 997:i * 997 + 3], 30 + i) for i in range(20)]
    cases.append([sst(base_sid, base_ts, 20)] + tiny + [sst(base_sid[:0], base_ts[:0], 99)])
    # interleaved halves: stream A holds even points, B odd points, C a random 10 %
    m = rng.random(len(base_sid)) < 0.1
    cases.append([sst(base_sid[0::2], base_ts[0::2], 40), sst(base_sid[1::2], base_ts[1::2], 41), sst(base_sid[m], base_ts[m], 42, rg=100)])
    for datas in cases:
        for preds in ((), [("tag", "eq", 1)], [("value", "lt", 0.3)]):
            for keep_builtin in (True, False):
                got = list(eng.scan(handle, _inputs(datas), preds, None, keep_builtin))
                exp = oracle.scan(datas, schema.arrow_schema, 2, preds, keep_builtin, 8192).batches
                check_stream(got, exp)
    eng.close()


@pytest.mark.parametrize("flags", [0, HG_FLAG_NO_FUSED])
def test_negative_timestamps_and_bucket_zero(flags):
    """Timestamp::truncate_by divides toward zero: bucket 0 covers (-w, w); negative timestamps fall into buckets -w, -2w, ...
    Checked on the fused kernel (bucket_range) and on the general pipeline (bucket_of), with and without predicates."""
    rng = np.random.default_rng(9)
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0, flags=flags)
    w = 60_000
    datas = []
    for f in range(2):
        sid = np.repeat(np.arange(f * 20, f * 20 + 20), 2000)
        ts = np.tile(np.arange(-1000, 1000) * 250 + rng.integers(0, 100), 20)      # -250 s .. +250 s around zero
        datas.append(sstgen.write_sst(schema, _metric_batch(sid, ts, rng.random(len(sid)), sid % 4), seq=700 + f,
                                      cfg=WriteConfig(compression=ParquetCompression.Uncompressed, max_row_group_size=3000), presorted=True))
    for preds in ([], [("ts", "ge", -90_000), ("ts", "lt", 90_001)], [("ts", "lt", 0)], [("tag", "eq", 2), ("ts", "gt", -w)]):
        for window in (w, 1000, 7):
            kw = dict(group_col=0, ts_col=1, window_ms=window, value_col=2)
            got = eng.scan_aggregate(handle, _inputs(datas), preds, **kw)
            exp = oracle.scan_aggregate(datas, schema.arrow_schema, 2, preds, **kw)
            if window == w:
                assert eng.stats()["path"] == (0 if flags else 1)
                if not preds:
                    b = exp.bucket[exp.gkey == 0].tolist()
                    assert 0 in b and -w in b and w in b and exp.count[(exp.gkey == 0) & (exp.bucket == 0)][0] > exp.count[(exp.gkey == 0) & (exp.bucket == w)][0]
            _check(got, exp)
    eng.close()


def _check_hash(got, exp, gname, has_bucket=True, has_group=True):
    assert got.num_rows == len(exp.count)
    if has_group:
        g = got[gname].to_numpy()
        assert g.view(np.uint64).tolist() == exp.gkey.tolist() if g.dtype.itemsize == 8 else g.astype(np.int64).astype(np.uint64).tolist() == exp.gkey.tolist()
    if has_bucket:
        assert got["bucket"].to_numpy().tolist() == exp.bucket.tolist()
    assert got["count"].to_numpy().tolist() == exp.count.tolist()
    assert np.array_equal(got["sum"].to_numpy(), exp.sum) and np.array_equal(got["min"].to_numpy(), exp.min) and np.array_equal(got["max"].to_numpy(), exp.max)


def test_config4b_group_by_tag_and_bucket_hash_mode():
    """Config 4(b): per-(tag, bucket) aggregates — the group key is not a prefix of the sort order, so groups are scattered
    through the stream: radix-partitioned aggregation (HG_AGG_HASH).  Sums are bit-exact: every group's rows are added in
    stream order, as a single-partition hash aggregation over the scan output would."""
    from horaedb_b200._ffi import HG_AGG_HASH
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    datas = [sstgen.synth_sst(300 * i, 300 * (i + 1), 400, 1000, seq=2000 + i, compression="snappy" if i % 2 else "none")[0] for i in range(6)]
    t0 = sstgen.T0_MS
    for preds in ([], [("ts", "ge", t0 + 50_000), ("ts", "lt", t0 + 333_000)], [("tag", "le", 5)]):
        for kw, gname, hb, hg in ((dict(group_col=3, ts_col=1, window_ms=60_000, value_col=2), "tag", True, True),
                                  (dict(group_col=3, ts_col=-1, window_ms=0, value_col=2), "tag", False, True),
                                  (dict(group_col=-1, ts_col=1, window_ms=30_000, value_col=2), None, True, False),
                                  (dict(group_col=0, ts_col=1, window_ms=60_000, value_col=2), "series_id", True, True)):
            got = eng.scan_aggregate(handle, _inputs(datas), preds, mode=HG_AGG_HASH, **kw)
            exp = oracle.scan_aggregate(datas, schema.arrow_schema, 2, preds, mode=1, **kw)
            assert len(exp.count) > 0
            _check_hash(got, exp, gname, hb, hg)
    eng.close()


def test_hash_mode_signed_and_float_group_columns():
    from helpers import arrow_schema, record_batch
    from horaedb_b200._ffi import HG_AGG_HASH
    from horaedb_b200.types import StorageSchema
    rng = np.random.default_rng(21)
    user = arrow_schema([("k", "uint64"), ("t", "int64"), ("g32", "int32"), ("gf", "float64"), ("g8", "int8"), ("v", "float64")])
    schema = StorageSchema.try_new(user, 2)
    n = 20_000
    b = record_batch(user, {"k": np.arange(n).tolist(), "t": (np.arange(n) * 10 - 70_000).tolist(), "g32": rng.integers(-5, 5, n).tolist(),
                            "gf": rng.choice([-1.5, -0.0, 0.0, 2.25, 1e300], n).tolist(), "g8": rng.integers(-128, 128, n).tolist(),
                            "v": rng.random(n).tolist()})
    data = sstgen.write_sst(schema, b, seq=5, cfg=WriteConfig(compression=ParquetCompression.Snappy, max_row_group_size=3000))
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    for gcol, gname in ((2, "g32"), (3, "gf"), (4, "g8")):
        for ts_col, w in ((-1, 0), (1, 25_000)):
            kw = dict(group_col=gcol, ts_col=ts_col, window_ms=w, value_col=5)
            got = eng.scan_aggregate(handle, _inputs([data]), [], mode=HG_AGG_HASH, **kw)
            exp = oracle.scan_aggregate([data], schema.arrow_schema, 2, [], mode=1, **kw)
            assert got.num_rows == len(exp.count)
            gv = got[gname].to_numpy()
            if gname == "gf":
                assert gv.view(np.uint64).tolist() == exp.gkey.tolist()
            else:
                assert gv.astype(np.int64).astype(np.uint64).tolist() == exp.gkey.tolist()
            if ts_col >= 0:
                assert got["bucket"].to_numpy().tolist() == exp.bucket.tolist()
            assert got["count"].to_numpy().tolist() == exp.count.tolist() and np.array_equal(got["sum"].to_numpy(), exp.sum)
    eng.close()


def test_in_list_and_float_total_order_predicates():
    """HG_OP_IN (DataFusion InListExpr; pruning = OR of equalities) and float predicates in IEEE totalOrder (arrow-rs
    comparison kernels: NaN above +inf, -0.0 < +0.0) — rows with NaN / signed zeros / NULLs in the predicate column."""
    from helpers import arrow_schema, record_batch
    from horaedb_b200.types import StorageSchema
    rng = np.random.default_rng(33)
    user = arrow_schema([("k", "uint64"), ("t", "int64"), ("f", "float64"), ("g", "float32"), ("u", "uint32"), ("i", "int32")])
    schema = StorageSchema.try_new(user, 2)
    n = 30_000
    f = rng.choice([float("nan"), -float("nan"), -0.0, 0.0, 1.5, -2.5, float("inf"), -float("inf"), 3.0], n).tolist()
    f = [None if rng.random() < 0.05 else x for x in f]
    b = record_batch(user, {"k": np.arange(n).tolist(), "t": np.zeros(n, dtype=np.int64).tolist(), "f": f,
                            "g": rng.choice([float("nan"), -0.0, 0.0, 1.0, -1.0], n).tolist(), "u": rng.integers(0, 50, n).tolist(),
                            "i": rng.integers(-20, 20, n).tolist()})
    data = sstgen.write_sst(schema, b, seq=9, cfg=WriteConfig(compression=ParquetCompression.Snappy, max_row_group_size=4000))
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    cases = [[("f", op, lit)] for op in ("eq", "ne", "lt", "le", "gt", "ge") for lit in (0.0, -0.0, 1.5, float("inf"), float("nan"))]
    cases += [[("g", "ge", 0.0)], [("g", "eq", float("nan"))], [("u", "in", [3, 7, 49, 1000])], [("i", "in", [-20, 0, 19])], [("i", "in", [])],
              [("f", "in", [1.5, -0.0])], [("u", "in", [5]), ("i", "lt", 3)]]
    for preds in cases:
        got = list(eng.scan(handle, _inputs([data]), preds, None, False))
        exp = oracle.scan([data], schema.arrow_schema, 2, preds, False, 8192).batches
        check_stream(got, exp)
    # row counts against a direct model of totalOrder for one case
    tbl = pa.Table.from_batches(list(eng.scan(handle, _inputs([data]), [("f", "gt", 3.0)], None, False)))
    want = sum(1 for x in f if x is not None and (x == float("inf") or (x != x and np.signbit(x) == False)))
    assert tbl.num_rows == want
    eng.close()


def test_delta_binary_packed_scan_matches_oracle():
    """DELTA_BINARY_PACKED pages (config.rs:54-75) on the PK / integer value columns, with NULLs, Snappy or not, several pages:
    decoded on the GPU (block-parallel unpack + prefix sums) and compared with the oracle through scan and aggregate."""
    from horaedb_b200.config import ColumnOptions
    from horaedb_b200.types import StorageSchema
    rng = np.random.default_rng(8)
    n = 40_000
    user = pa.schema([pa.field("k", pa.uint64()), pa.field("t", pa.int64()), pa.field("i32", pa.int32()), pa.field("u32", pa.uint32()),
                      pa.field("i64", pa.int64()), pa.field("v", pa.float64())])
    schema = StorageSchema.try_new(user, 2)

    def nulls(a, p, t):
        return pa.array([None if rng.random() < p else int(x) for x in a], t)

    batch = pa.RecordBatch.from_arrays(
        [pa.array(np.arange(n, dtype=np.uint64) // 7), pa.array(np.arange(n, dtype=np.int64) * 1000 - 5_000_000 + rng.integers(0, 300, n)),
         nulls(rng.integers(-2**31, 2**31, n), 0.1, pa.int32()), nulls(rng.integers(0, 2**32, n), 0.3, pa.uint32()),
         nulls(rng.integers(-2**62, 2**62, n), 0.02, pa.int64()), pa.array(rng.random(n))], schema=user)
    opts = {c: ColumnOptions(encoding="DELTA_BINARY_PACKED") for c in ("k", "t", "i32", "u32", "i64", "__seq__")}
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    for comp in ("snappy", "none"):
        for rg in (8192, 3000, 100):
            data = sstgen.write_sst(schema, batch, 9, WriteConfig(compression=comp, max_row_group_size=rg, column_options=opts), presorted=True)
            for preds in ((), [("t", "ge", 0), ("u32", "lt", 2**31)], [("i64", "gt", 0)]):
                got = list(eng.scan(handle, _inputs([data]), preds, None, True))
                exp = oracle.scan([data], schema.arrow_schema, 2, preds, True, 8192).batches
                check_stream(got, exp)
            kw = dict(group_col=0, ts_col=1, window_ms=3_600_000, value_col=5)
            a = eng.scan_aggregate(handle, _inputs([data]), [("t", "ge", -1_000_000)], **kw)
            b = oracle.scan_aggregate([data], schema.arrow_schema, 2, [("t", "ge", -1_000_000)], **kw)
            assert a["count"].to_numpy().tolist() == b.count.tolist() and np.array_equal(a["sum"].to_numpy(), b.sum)
    # the metric schema with delta-encoded series_id / ts: what the reference would pick for these columns
    mschema = sstgen.metric_storage_schema()
    mh = SchemaHandle(mschema.arrow_schema, 2)
    sid, ts, value, tag = sstgen.synth_columns(0, 50, 2000, 1000)
    mb = pa.RecordBatch.from_arrays([pa.array(sid), pa.array(ts), pa.array(value), pa.array(tag)], schema=sstgen.METRIC_SCHEMA)
    mopts = {c: ColumnOptions(encoding="DELTA_BINARY_PACKED") for c in ("series_id", "ts")}
    data = sstgen.write_sst(mschema, mb, 3, WriteConfig(column_options=mopts), presorted=True)
    plain = sstgen.write_sst(mschema, mb, 3, WriteConfig(), presorted=True)
    assert len(data) < 0.8 * len(plain)
    preds = [("tag", "eq", 3), ("ts", "ge", sstgen.T0_MS + 500_000)]
    kw = dict(group_col=0, ts_col=-1, window_ms=0, value_col=2)
    a = eng.scan_aggregate(mh, _inputs([data]), preds, **kw)
    b = oracle.scan_aggregate([data], mschema.arrow_schema, 2, preds, **kw)
    assert a["series_id"].to_numpy().tolist() == b.gkey.tolist() and np.array_equal(a["sum"].to_numpy(), b.sum)
    eng.close()


def test_dictionary_encoded_columns_scan_matches_oracle():
    """RLE_DICTIONARY pages (`enable_dict`, config.rs:98-103,127): dictionary page + RLE / bit-packed index runs, Snappy or not, with
    NULLs and PLAIN fall-back pages — decoded on the GPU and compared with the oracle (which is pinned on pyarrow)."""
    from horaedb_b200.types import StorageSchema
    rng = np.random.default_rng(15)
    n = 30_000
    user = pa.schema([pa.field("k", pa.uint64()), pa.field("t", pa.int64()), pa.field("lowcard", pa.int32()), pa.field("f", pa.float64()),
                      pa.field("u", pa.uint32()), pa.field("wide", pa.int64())])
    schema = StorageSchema.try_new(user, 2)

    def nulls(a, p, t):
        return pa.array([None if rng.random() < p else x for x in a], t)

    batch = pa.RecordBatch.from_arrays(
        [pa.array(np.arange(n, dtype=np.uint64) // 5), pa.array(np.arange(n, dtype=np.int64) * 1000),
         nulls(rng.integers(-3, 4, n).tolist(), 0.1, pa.int32()), nulls(rng.choice([0.5, -1.25, 3.0, 1e10], n).tolist(), 0.2, pa.float64()),
         pa.array(np.repeat(rng.integers(0, 50, n // 100 + 1), 100)[:n].astype(np.uint32)), pa.array(rng.integers(-2**62, 2**62, n))], schema=user)
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    for comp in ("snappy", "none"):
        for rg in (8192, 2500):
            data = sstgen.write_sst(schema, batch, 11, WriteConfig(compression=comp, max_row_group_size=rg, enable_dict=True), presorted=True)
            for preds in ((), [("lowcard", "ge", 0)], [("u", "in", [1, 7, 30]), ("t", "ge", 5_000_000)], [("f", "gt", 0.75)]):
                got = list(eng.scan(handle, _inputs([data]), preds, None, True))
                exp = oracle.scan([data], schema.arrow_schema, 2, preds, True, 8192).batches
                check_stream(got, exp)
    # the metric schema written with dictionaries on every column: aggregate parity
    mschema = sstgen.metric_storage_schema()
    mh = SchemaHandle(mschema.arrow_schema, 2)
    sid, ts, value, tag = sstgen.synth_columns(0, 40, 1500, 1000)
    mb = pa.RecordBatch.from_arrays([pa.array(sid), pa.array(ts), pa.array(np.round(value * 50) / 50), pa.array(tag)], schema=sstgen.METRIC_SCHEMA)
    data = sstgen.write_sst(mschema, mb, 3, WriteConfig(enable_dict=True), presorted=True)
    preds = [("tag", "eq", 3), ("ts", "ge", sstgen.T0_MS + 500_000)]
    for kw in (dict(group_col=0, ts_col=-1, window_ms=0, value_col=2), dict(group_col=0, ts_col=1, window_ms=60_000, value_col=2)):
        a = eng.scan_aggregate(mh, _inputs([data]), preds, **kw)
        b = oracle.scan_aggregate([data], mschema.arrow_schema, 2, preds, **kw)
        assert a["series_id"].to_numpy().tolist() == b.gkey.tolist() and a["count"].to_numpy().tolist() == b.count.tolist()
        assert np.array_equal(a["sum"].to_numpy(), b.sum)
    eng.close()
