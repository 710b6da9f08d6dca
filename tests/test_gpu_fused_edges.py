"""GPU parity tests aimed at the fused single-pass kernel (fused_scan.cu): duplicate primary keys inside a file (the
look-ahead dedup path), ragged row-group sizes, runs crossing row-group / file boundaries, and BASELINE.json's full
config-2 size checked through size-independent properties."""
import numpy as np
import pyarrow as pa
import pytest

from horaedb_b200 import sstgen
from horaedb_b200._ffi import HG_FLAG_NO_FUSED, Engine, SchemaHandle, SstInput
from horaedb_b200.config import ParquetCompression, WriteConfig
from oracle import oracle

pytestmark = pytest.mark.gpu
_ids = iter(range(50_000_000, 60_000_000))


def _metric_batch(sid, ts, value, tag):
    return pa.RecordBatch.from_arrays([pa.array(sid.astype(np.uint64)), pa.array(ts.astype(np.int64)), pa.array(value.astype(np.float64)),
                                       pa.array(tag.astype(np.uint32))], schema=sstgen.METRIC_SCHEMA)


def _agg(eng, handle, datas, preds, **kw):
    return eng.scan_aggregate(handle, [SstInput(id=next(_ids), data=d) for d in datas], preds, **kw)


def _check(got, exp, bucket):
    assert got.num_rows == len(exp.count)
    assert got["series_id"].to_numpy().tolist() == exp.gkey.tolist()
    if bucket:
        assert got["bucket"].to_numpy().tolist() == exp.bucket.tolist()
    assert got["count"].to_numpy().tolist() == exp.count.tolist()
    assert np.array_equal(got["sum"].to_numpy(), exp.sum) and np.array_equal(got["min"].to_numpy(), exp.min) and np.array_equal(got["max"].to_numpy(), exp.max)


@pytest.mark.parametrize("rg", [8192, 1000, 97, 33])
def test_fused_intra_file_duplicates_and_ragged_row_groups(rg):
    """Duplicate (series_id, ts) rows inside one file: LastValue keeps the LAST one that passes the filter
    (read.rs:459-480: the filter runs first), including runs that straddle slices, blocks and row groups."""
    rng = np.random.default_rng(rg)
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    datas = []
    for f, lo in enumerate((0, 40)):
        sid = np.repeat(np.arange(lo, lo + 40), 300)
        ts = sstgen.T0_MS + np.tile(np.arange(300) * 1000, 40)
        # duplicate ~20% of the rows 1-4 times, keep file order (stable) = sort order
        reps = np.where(rng.random(len(sid)) < 0.2, rng.integers(2, 6, len(sid)), 1)
        sid, ts = np.repeat(sid, reps), np.repeat(ts, reps)
        value = rng.random(len(sid))
        tag = rng.integers(0, 4, len(sid))          # per-row tag so that duplicates differ in whether they pass the filter
        datas.append(sstgen.write_sst(schema, _metric_batch(sid, ts, value, tag), seq=900 + f,
                                      cfg=WriteConfig(compression=ParquetCompression.Uncompressed, max_row_group_size=rg), presorted=True))
    for preds in ([], [("tag", "eq", 3)], [("tag", "le", 1), ("ts", "ge", sstgen.T0_MS + 50_000)]):
        for kw in (dict(group_col=0, ts_col=-1, window_ms=0, value_col=2), dict(group_col=0, ts_col=1, window_ms=60_000, value_col=2),
                   dict(group_col=-1, ts_col=-1, window_ms=0, value_col=-1)):
            got = _agg(eng, handle, datas, preds, **kw)
            assert rg < 97 or eng.stats()["path"] == 1, "expected the fused path"   # tiny row groups: the planner picks the general pipeline
            exp = oracle.scan_aggregate(datas, schema.arrow_schema, 2, preds, **kw)
            if kw["group_col"] < 0:
                assert got["count"].to_pylist() == exp.count.tolist()
            else:
                _check(got, exp, kw["ts_col"] >= 0)
    eng.close()


def test_fused_long_runs_cross_row_groups_and_files():
    """One series longer than several row groups, tiny series, and a file boundary in the middle of the key space."""
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    rng = np.random.default_rng(1)
    lens = [30_000, 1, 2, 50, 9000, 1, 1, 20_000]
    datas = []
    sid0 = 0
    for f in range(2):
        sid = np.concatenate([np.full(n, sid0 + i) for i, n in enumerate(lens)])
        ts = np.concatenate([sstgen.T0_MS + np.arange(n) * 700 for n in lens])
        sid0 += len(lens)
        datas.append(sstgen.write_sst(schema, _metric_batch(sid, ts, rng.random(len(sid)), sid % 5), seq=950 + f,
                                      cfg=WriteConfig(compression=ParquetCompression.Uncompressed, max_row_group_size=4096), presorted=True))
    for preds in ([], [("tag", "ne", 0)], [("ts", "lt", sstgen.T0_MS + 5_000_000)]):
        for kw in (dict(group_col=0, ts_col=-1, window_ms=0, value_col=2), dict(group_col=0, ts_col=1, window_ms=3_600_000, value_col=2)):
            got = _agg(eng, handle, datas, preds, **kw)
            exp = oracle.scan_aggregate(datas, schema.arrow_schema, 2, preds, **kw)
            _check(got, exp, kw["ts_col"] >= 0)
    eng.close()


def test_config2_full_size_properties():
    """BASELINE.json configs[1] at full size (100 k series x 1 k points, 16 SSTs): selection count, per-series counts and
    a sample of sequential f64 sums against numpy; fused and general pipelines agree bit for bit."""
    import bench
    ssts = bench.gen_ssts(0, "none", 16, 16)
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    preds = bench.preds()
    eng = Engine(device=0)
    inputs = []
    for sid, data, n in ssts:
        eng.load_sst(handle, SstInput(id=sid, data=data, num_rows=n))
        inputs.append(SstInput(id=sid))
    got = eng.scan_aggregate(handle, inputs, preds, group_col=0, ts_col=-1, window_ms=0, value_col=2)
    st = eng.stats()
    assert st["path"] == 1 and st["rows_in_files"] == 100_000_000 and st["rows_decoded"] < 100_000_000
    cnt = eng.scan_aggregate(handle, inputs, [], group_col=-1, ts_col=-1, window_ms=0, value_col=-1)
    assert cnt["count"].to_pylist() == [100_000_000]
    # expected from the generator itself (numpy), series with tag == 3 only
    series = np.arange(3, 100_000, 16)
    assert got["series_id"].to_numpy().tolist() == series.tolist()
    t_lo, t_hi = preds[1][2], preds[2][2]
    counts, sums = [], []
    for s0 in range(0, len(series), 500):
        for s in series[s0:s0 + 500:97]:                      # a sample of series for the exact sums
            sid, ts, value, tag = sstgen.synth_columns(int(s), int(s) + 1, 1000, 1000)
            m = (ts >= t_lo) & (ts < t_hi)
            counts.append((int(s), int(m.sum())))
            sums.append((int(s), float(np.cumsum(value[m])[-1]) if m.any() else 0.0))   # cumsum = sequential addition
    gmap = {int(k): (int(c), float(x)) for k, c, x in zip(got["series_id"].to_numpy(), got["count"].to_numpy(), got["sum"].to_numpy())}
    for (s, c), (_, x) in zip(counts, sums):
        assert gmap[s][0] == c and gmap[s][1] == x
    assert int(got["count"].to_numpy().sum()) == st["rows_filtered"] == st["rows_out"]
    e2 = Engine(device=0, flags=HG_FLAG_NO_FUSED)
    for sid, data, n in ssts:
        e2.load_sst(handle, SstInput(id=sid, data=data, num_rows=n))
    ref = e2.scan_aggregate(handle, inputs, preds, group_col=0, ts_col=-1, window_ms=0, value_col=2)
    assert e2.stats()["path"] == 0 and ref.equals(got)
    e2.close()
    eng.close()


def test_fused_late_materialization_gate():
    """The fused kernel loads the narrowest predicate column first and the other columns only for blocks with a passing
    row (fused_scan.cu process_block GATED).  Same answers as the oracle and as the ungated kernel, for every gate
    position: a 4-byte extra column, the narrower of two extra columns, pk1 alone; selective and dense predicates."""
    from horaedb_b200._ffi import HG_FLAG_NO_LATE_MATERIALIZATION
    from horaedb_b200.types import StorageSchema
    rng = np.random.default_rng(11)
    user = pa.schema([pa.field("series_id", pa.uint64(), True), pa.field("ts", pa.int64(), True), pa.field("value", pa.float64(), True),
                      pa.field("tag", pa.int32(), True), pa.field("big", pa.int64(), True)])
    schema = StorageSchema.try_new(user, 2)
    handle = SchemaHandle(schema.arrow_schema, 2)
    datas = []
    for f, lo in enumerate((0, 700)):
        lens = rng.integers(1, 400, 700)
        sid = np.repeat(np.arange(lo, lo + 700), lens)
        ts = sstgen.T0_MS + np.concatenate([np.arange(n) * 1000 for n in lens])
        tag = np.repeat(rng.integers(-8, 8, 700), lens)                 # per-series tag: long all-fail stretches
        big = rng.integers(-5, 5, len(sid)) * (1 << 40)                 # per-row 8-byte predicate column
        batch = pa.RecordBatch.from_arrays([pa.array(sid.astype(np.uint64)), pa.array(ts.astype(np.int64)), pa.array(rng.random(len(sid))),
                                            pa.array(tag.astype(np.int32)), pa.array(big.astype(np.int64))], schema=user)
        datas.append(sstgen.write_sst(schema, batch, seq=970 + f,
                                      cfg=WriteConfig(compression=ParquetCompression.Uncompressed, max_row_group_size=5000), presorted=True))
    eng = Engine(device=0)
    ref = Engine(device=0, flags=HG_FLAG_NO_LATE_MATERIALIZATION)
    cases = [([("tag", "eq", -3)], True),
             ([("big", "ge", 3 << 40), ("tag", "eq", 5)], True),                       # two extra columns: gate = the 4-byte one
             ([("tag", "lt", 7), ("big", "gt", -(6 << 40))], False),                   # nearly everything passes (dense blocks)
             ([("ts", "ge", sstgen.T0_MS + 390_000)], True),                           # gate on pk1
             ([("big", "eq", 4 << 40)], False),                                       # 8-byte gate, ~10 % of rows scattered
             ([("tag", "eq", 99)], True)]                                              # nothing passes (pruned by statistics)
    for preds, selective in cases:
        for kw in (dict(group_col=0, ts_col=-1, window_ms=0, value_col=2), dict(group_col=0, ts_col=1, window_ms=300_000, value_col=2),
                   dict(group_col=-1, ts_col=-1, window_ms=0, value_col=-1)):
            got = _agg(eng, handle, datas, preds, **kw)
            st = eng.stats()
            want = _agg(ref, handle, datas, preds, **kw)
            sr = ref.stats()
            assert st["path"] == 1 and sr["path"] == 1
            # (the inputs are host bytes: the transient load already drops row groups without a passing gate row)
            assert sr["rows_materialized"] == sr["rows_decoded"] >= st["rows_decoded"]
            assert st["rows_materialized"] <= st["rows_decoded"]
            if selective and sr["rows_decoded"]:
                assert st["rows_materialized"] < sr["rows_decoded"] // 2
            assert got.equals(want)
            exp = oracle.scan_aggregate(datas, schema.arrow_schema, 2, preds, **kw)
            if kw["group_col"] < 0:
                assert got["count"].to_pylist() == exp.count.tolist()
            else:
                _check(got, exp, kw["ts_col"] >= 0)
    eng.close()
    ref.close()
