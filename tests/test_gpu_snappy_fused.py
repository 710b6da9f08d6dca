"""GPU parity tests of the fused scan over Snappy SSTs (the reference's default codec, WriteConfig::default,
config.rs:120-133): pages are decompressed into scratch regions first (gate column first, the other columns only for
row groups that hold a passing row), the incompressible value column is read in place (stored pages), and the same
single-pass kernel as for uncompressed SSTs runs on top.  Everything is compared bit for bit with the CPU oracle."""
import numpy as np
import pyarrow as pa
import pytest

from horaedb_b200 import sstgen
from horaedb_b200._ffi import HG_FLAG_NO_FUSED, HG_FLAG_NO_LATE_MATERIALIZATION, Engine, SchemaHandle, SstInput
from horaedb_b200.config import ParquetCompression, WriteConfig
from oracle import oracle

pytestmark = pytest.mark.gpu
_ids = iter(range(70_000_000, 80_000_000))


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


KWS = (dict(group_col=0, ts_col=-1, window_ms=0, value_col=2), dict(group_col=0, ts_col=1, window_ms=300_000, value_col=2),
       dict(group_col=-1, ts_col=-1, window_ms=0, value_col=-1))


def _preds():
    t0 = sstgen.T0_MS
    return ([], [("tag", "eq", 3)], [("tag", "eq", 3), ("ts", "ge", t0 + 100_000), ("ts", "lt", t0 + 700_000)],
            [("ts", "ge", t0 + 100_000)], [("tag", "gt", 100)], [("series_id", "ge", 30), ("tag", "le", 7)])


@pytest.mark.parametrize("codecs", [("snappy", "snappy", "snappy"), ("snappy", "none", "snappy")])
def test_snappy_fused_matches_oracle(codecs):
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    datas = [sstgen.synth_sst(40 * i, 40 * i + 40, 1000, 1000, seq=200 + i, compression=c)[0] for i, c in enumerate(codecs)]
    for preds in _preds():
        for kw in KWS:
            exp = oracle.scan_aggregate(datas, schema.arrow_schema, 2, preds, **kw)
            for flags in (0, HG_FLAG_NO_LATE_MATERIALIZATION):
                eng.set_flags(flags)
                got = _agg(eng, handle, datas, preds, **kw)
                st = eng.stats()
                assert st["path"] == 1, "Snappy SSTs must take the fused path"
                if kw["group_col"] < 0:
                    assert got["count"].to_pylist() == exp.count.tolist()
                else:
                    _check(got, exp, kw["ts_col"] >= 0)
            eng.set_flags(0)
    eng.close()


def test_snappy_fused_value_column_variants():
    """The value column as stored pages (random f64), as compressible pages (constant runs), as a 4-byte column, and with
    row groups whose value page splits into two literals at a non-trivial row."""
    rng = np.random.default_rng(5)
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    n = 30_000
    sid = np.repeat(np.arange(30), 1000)
    ts = sstgen.T0_MS + np.tile(np.arange(1000) * 1000, 30)
    tag = (sid % 4).astype(np.uint32)
    # "mixed": random values with one compressible stretch, so one file holds stored pages (read in place) next to pages with copy
    # elements (decompressed into scratch) — what a large random f64 column looks like in practice (the odd 4-byte match)
    mixed = rng.random(n)
    mixed[9000:15000] = 0.25
    for name, value, rg in (("random", rng.random(n), 8192), ("runs", np.repeat(rng.random(n // 100), 100), 8192), ("ragged", rng.random(n), 5000),
                            ("big-rg", rng.random(n), 20_000), ("mixed", mixed, 4096), ("mixed-ragged", mixed, 5000)):
        data = sstgen.write_sst(schema, _metric_batch(sid, ts, value, tag), seq=300,
                                cfg=WriteConfig(compression=ParquetCompression.Snappy, max_row_group_size=rg), presorted=True)
        for preds in ([], [("tag", "eq", 1)], [("tag", "eq", 1), ("ts", "lt", sstgen.T0_MS + 400_000)]):
            for kw in KWS[:2]:
                got = _agg(eng, handle, [data], preds, **kw)
                assert eng.stats()["path"] == 1, name
                _check(got, oracle.scan_aggregate([data], schema.arrow_schema, 2, preds, **kw), kw["ts_col"] >= 0)
    # the tag column as the aggregated value: a 4-byte, highly compressible value column
    data = sstgen.write_sst(schema, _metric_batch(sid, ts, rng.random(n), tag), seq=301, cfg=WriteConfig(compression=ParquetCompression.Snappy), presorted=True)
    kw = dict(group_col=0, ts_col=-1, window_ms=0, value_col=3)
    got = _agg(eng, handle, [data], [("ts", "ge", sstgen.T0_MS + 5000)], **kw)
    _check(got, oracle.scan_aggregate([data], schema.arrow_schema, 2, [("ts", "ge", sstgen.T0_MS + 5000)], **kw), False)
    eng.close()


def test_snappy_fused_duplicates_and_gate_dropped_row_groups():
    """Intra-file duplicate PKs whose copies differ in whether they pass the filter, with runs crossing row groups that the
    gate drops entirely (no passing row): LastValue must still pick the last PASSING copy (read.rs:459-480)."""
    rng = np.random.default_rng(11)
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    sid = np.repeat(np.arange(12), 2500)
    ts = sstgen.T0_MS + np.tile(np.repeat(np.arange(500) * 1000, 5), 12)      # every PK five times
    value = rng.random(len(sid))
    # tag passes (== 3) only in a few short stretches, so most 1000-row groups hold no passing row at all
    tag = np.where((np.arange(len(sid)) // 700) % 9 == 0, 3, 5)
    tag[rng.random(len(sid)) < 0.002] = 3
    data = sstgen.write_sst(schema, _metric_batch(sid, ts, value, tag), seq=400,
                            cfg=WriteConfig(compression=ParquetCompression.Snappy, max_row_group_size=1000), presorted=True)
    for preds in ([("tag", "eq", 3)], [("tag", "eq", 3), ("ts", "ge", sstgen.T0_MS + 100_000)]):
        for kw in KWS:
            got = _agg(eng, handle, [data], preds, **kw)
            assert eng.stats()["path"] == 1
            exp = oracle.scan_aggregate([data], schema.arrow_schema, 2, preds, **kw)
            if kw["group_col"] < 0:
                assert got["count"].to_pylist() == exp.count.tolist()
            else:
                _check(got, exp, kw["ts_col"] >= 0)
    eng.close()


def test_snappy_transient_prefix_transfer_and_retry():
    """Host-buffer SSTs (transient loads): pages the fused scan decodes only up to the last gate-passing row cross PCIe as a
    PREFIX of their compressed stream, sized by the share of the output that is needed.  A lopsided page (incompressible first
    half, constant second half) makes that estimate too short: the decoder must notice the stream ending early and the call must
    repeat with whole pages — same result as the oracle either way."""
    rng = np.random.default_rng(21)
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    n = 40_000
    sid = np.repeat(np.arange(40), 1000)
    ts = sstgen.T0_MS + np.tile(np.arange(1000) * 1000 + rng.integers(0, 400, 1000), 40)
    for name, tag, value in (
            # the passing rows sit in the first half of every 8192-row group; the value page is lopsided (random, then constant)
            ("lopsided", np.where((np.arange(n) % 8192) < 3500, 3, 5), np.where((np.arange(n) % 8192) < 4000, rng.random(n), 0.5)),
            # evenly compressible columns: the prefix estimate holds
            ("even", np.where((np.arange(n) % 8192) < 3500, 3, 5), np.round(rng.random(n), 2))):
        data = sstgen.write_sst(schema, _metric_batch(sid, ts, value, tag.astype(np.uint32)), seq=450,
                                cfg=WriteConfig(compression=ParquetCompression.Snappy), presorted=True)
        for preds in ([("tag", "eq", 3)], [("tag", "eq", 3), ("ts", "ge", sstgen.T0_MS + 200_000)]):
            kw = KWS[0]
            got = _agg(eng, handle, [data], preds, **kw)
            assert eng.stats()["path"] == (3 if name == "lopsided" else 1), name      # bit 1: the call had to be repeated
            _check(got, oracle.scan_aggregate([data], schema.arrow_schema, 2, preds, **kw), False)
    eng.close()


def test_snappy_general_pipeline_still_agrees():
    """HG_FLAG_NO_FUSED: the materialising pipeline (decompress -> decode -> filter -> dedup -> reduce) on the same files."""
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0, flags=HG_FLAG_NO_FUSED)
    datas = [sstgen.synth_sst(40 * i, 40 * i + 40, 600, 1000, seq=500 + i, compression="snappy")[0] for i in range(2)]
    preds = _preds()[2]
    for kw in KWS[:2]:
        got = _agg(eng, handle, datas, preds, **kw)
        assert eng.stats()["path"] == 0
        _check(got, oracle.scan_aggregate(datas, schema.arrow_schema, 2, preds, **kw), kw["ts_col"] >= 0)
    eng.close()


def test_snappy_decoder_torture():
    """The Snappy decompressor on streams of every shape it special-cases: periodic copies of every small offset (run mode
    with the one-word pattern for offsets 1/2/4/8, the generic path otherwise), short elements with sources inside the
    batch / straddling elements / far back (word mode and its prefix splits), long literals, and mixtures — decoded through
    the general pipeline and compared with pyarrow's reading of the same bytes."""
    import io

    import pyarrow.parquet as pq
    rng = np.random.default_rng(77)
    n = 50_000
    cols = {}
    for period in (1, 2, 3, 4, 5, 6, 7, 8, 12, 16, 24):               # byte-periodic u8-ish patterns inside u64/u32 values
        pat = rng.integers(0, 256, period, dtype=np.uint8)
        raw = np.resize(pat, n * 8)
        cols[f"p{period}_u64"] = raw.view(np.uint64)
        cols[f"p{period}_u32"] = np.resize(pat, n * 4).view(np.uint32)
    cols["sawtooth"] = (np.arange(n, dtype=np.uint64) % 1000) * 37
    cols["slow_counter"] = np.arange(n, dtype=np.uint64) // 3
    cols["jitter_ts"] = (1_700_000_000_000 + np.arange(n, dtype=np.uint64) * 1000 + rng.integers(0, 500, n).astype(np.uint64))
    cols["random"] = rng.integers(0, 2**63, n, dtype=np.uint64)
    cols["few_values"] = rng.choice(rng.integers(0, 2**60, 5, dtype=np.uint64), n)
    mix = rng.integers(0, 2**63, n, dtype=np.uint64)
    mix[1000:20000] = 7
    mix[30000:30100] = np.arange(100, dtype=np.uint64)
    cols["mixed"] = mix
    cols["small_ints_u32"] = rng.integers(0, 16, n).astype(np.uint32)
    cols["runs_u32"] = np.repeat(rng.integers(0, 2**31, n // 50 + 1), 50)[:n].astype(np.uint32)
    all_names = list(cols)
    eng = Engine(device=0)
    for part in (all_names[:16], all_names[16:]):          # the ABI caps a schema at 32 columns
        _torture_part(eng, cols, part, n)
    eng.close()


def _torture_part(eng, cols, names, n):
    import io

    import pyarrow.parquet as pq
    spec = pa.schema([pa.field("k0", pa.uint64()), pa.field("k1", pa.int64())] + [pa.field(c, pa.uint64() if cols[c].dtype == np.uint64 else pa.uint32()) for c in names])
    from horaedb_b200.types import StorageSchema
    schema = StorageSchema.try_new(spec, 2)
    batch = pa.RecordBatch.from_arrays([pa.array(np.arange(n, dtype=np.uint64)), pa.array(np.zeros(n, dtype=np.int64))] + [pa.array(cols[c]) for c in names], schema=spec)
    handle = SchemaHandle(schema.arrow_schema, 2)
    for rg in (8192, 50_000, 777):
        data = sstgen.write_sst(schema, batch, seq=600, cfg=WriteConfig(compression=ParquetCompression.Snappy, max_row_group_size=rg), presorted=True)
        got = eng.scan(handle, [SstInput(id=next(_ids), data=data)]).read_all()
        ref = pq.read_table(io.BytesIO(data))
        for c in ["k0"] + names:
            assert got[c].to_numpy().tolist() == ref[c].to_numpy().tolist(), (rg, c)
