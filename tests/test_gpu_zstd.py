"""GPU parity tests of ParquetCompression::Zstd SSTs (config.rs:78-94): pages are decompressed by zstd_chunks_kernel (zstd_core.h) into
the general pipeline's scratch.  The CPU oracle does not read Zstandard, so every Zstd SST has a TWIN — the same rows written with
Snappy — and the GPU's result on the Zstd file must equal the oracle's on the twin (identical rows => identical stream, boundaries and
aggregates); raw decode is also compared with pyarrow's reading of the Zstd bytes."""
import io

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from helpers import check_stream
from horaedb_b200 import sstgen
from horaedb_b200._ffi import Engine, SchemaHandle, SstInput
from horaedb_b200.config import ParquetCompression, WriteConfig
from horaedb_b200.types import StorageSchema
from oracle import oracle

pytestmark = pytest.mark.gpu
_ids = iter(range(90_000_000, 99_000_000))


def _twins(schema, batch, seq, rg=8192, presorted=True):
    out = []
    for comp in (ParquetCompression.Zstd, ParquetCompression.Snappy):
        out.append(sstgen.write_sst(schema, batch, seq=seq, cfg=WriteConfig(compression=comp, max_row_group_size=rg), presorted=presorted))
    return out


def _metric_batch(sid, ts, value, tag):
    return pa.RecordBatch.from_arrays([pa.array(sid.astype(np.uint64)), pa.array(ts.astype(np.int64)), pa.array(value.astype(np.float64)),
                                       pa.array(tag.astype(np.uint32))], schema=sstgen.METRIC_SCHEMA)


def test_zstd_scan_and_aggregate_match_the_oracle_on_snappy_twins():
    rng = np.random.default_rng(3)
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    zs, ss = [], []
    for f in range(3):                                           # overlapping files: a real merge + dedup over Zstd inputs
        sid = np.repeat(np.arange(20 * f, 20 * f + 40), 600)
        ts = sstgen.T0_MS + np.tile(np.arange(600) * 1000 + rng.integers(0, 300, 600), 40)
        z, s = _twins(schema, _metric_batch(sid, ts, rng.random(len(sid)), sid % 16), seq=700 + f)
        zs.append(z); ss.append(s)
    t0 = sstgen.T0_MS
    for preds in ([], [("tag", "eq", 3)], [("ts", "ge", t0 + 100_000), ("ts", "lt", t0 + 400_000)]):
        got = list(eng.scan(handle, [SstInput(id=next(_ids), data=d) for d in zs], preds))
        exp = oracle.scan(ss, schema.arrow_schema, 2, preds).batches
        check_stream(got, exp)
        for kw in (dict(group_col=0, ts_col=-1, window_ms=0, value_col=2), dict(group_col=0, ts_col=1, window_ms=60_000, value_col=2)):
            tbl = eng.scan_aggregate(handle, [SstInput(id=next(_ids), data=d) for d in zs], preds, **kw)
            ex = oracle.scan_aggregate(ss, schema.arrow_schema, 2, preds, **kw)
            assert tbl["series_id"].to_numpy().tolist() == ex.gkey.tolist() and tbl["count"].to_numpy().tolist() == ex.count.tolist()
            assert np.array_equal(tbl["sum"].to_numpy(), ex.sum) and np.array_equal(tbl["min"].to_numpy(), ex.min) and np.array_equal(tbl["max"].to_numpy(), ex.max)
    # mixed codecs in one call (Zstd + Snappy + uncompressed), PK-disjoint: the fused path must decline, the result must not change
    datas, twins = [], []
    for f, comp in enumerate((ParquetCompression.Zstd, ParquetCompression.Snappy, ParquetCompression.Uncompressed)):
        sid = np.repeat(np.arange(100 * f, 100 * f + 30), 500)
        ts = sstgen.T0_MS + np.tile(np.arange(500) * 1000, 30)
        b = _metric_batch(sid, ts, rng.random(len(sid)), sid % 4)
        datas.append(sstgen.write_sst(schema, b, seq=800 + f, cfg=WriteConfig(compression=comp), presorted=True))
        twins.append(sstgen.write_sst(schema, b, seq=800 + f, cfg=WriteConfig(compression=ParquetCompression.Snappy), presorted=True))
    kw = dict(group_col=0, ts_col=-1, window_ms=0, value_col=2)
    tbl = eng.scan_aggregate(handle, [SstInput(id=next(_ids), data=d) for d in datas], [("tag", "eq", 1)], **kw)
    assert eng.stats()["path"] == 0
    ex = oracle.scan_aggregate(twins, schema.arrow_schema, 2, [("tag", "eq", 1)], **kw)
    assert tbl["series_id"].to_numpy().tolist() == ex.gkey.tolist() and np.array_equal(tbl["sum"].to_numpy(), ex.sum)
    # resident Zstd SSTs + compaction (merge + dedup + GPU Parquet encode) of Zstd inputs
    ids = []
    for d in zs:
        i = next(_ids)
        eng.load_sst(handle, SstInput(id=i, data=d))
        ids.append(i)
    got = list(eng.compact(handle, [SstInput(id=i) for i in ids]))
    exp = oracle.scan(ss, schema.arrow_schema, 2, [], True).batches
    check_stream(got, exp)
    eng.close()


def test_zstd_decode_of_every_column_shape_matches_pyarrow():
    """Column shapes that drive the frame format through its modes (raw / RLE / Huffman literals, predefined / FSE / RLE / repeated
    sequence tables, long overlapping matches), all primitive widths, NULLs, small pages and V2 pages — scanned through the engine and
    compared with pyarrow's reading of the same bytes."""
    rng = np.random.default_rng(77)
    n = 30_000
    cols = {
        "jitter_ts": (1_700_000_000_000 + np.arange(n, dtype=np.uint64) * 1000 + rng.integers(0, 500, n).astype(np.uint64)),
        "random": rng.integers(0, 2**63, n, dtype=np.uint64),
        "few_values": rng.choice(rng.integers(0, 2**60, 5, dtype=np.uint64), n),
        "slow_counter": np.arange(n, dtype=np.uint64) // 3,
        "constant": np.full(n, 7, dtype=np.uint64),
        "small_u32": rng.integers(0, 16, n).astype(np.uint32),
        "runs_u32": np.repeat(rng.integers(0, 2**31, n // 50 + 1), 50)[:n].astype(np.uint32),
        "f64_round": np.round(rng.random(n), 2),
        "i16": rng.integers(-300, 300, n).astype(np.int16),
        "u8": rng.integers(0, 5, n).astype(np.uint8),
    }
    names = list(cols)
    fields = [pa.field("k0", pa.uint64()), pa.field("k1", pa.int64())] + [pa.field(c, pa.from_numpy_dtype(cols[c].dtype)) for c in names]
    spec = pa.schema(fields)
    schema = StorageSchema.try_new(spec, 2)
    arrays = [pa.array(np.arange(n, dtype=np.uint64)), pa.array(np.zeros(n, dtype=np.int64))]
    for c in names:
        mask = None
        if c in ("few_values", "small_u32", "f64_round"):
            mask = rng.random(n) < 0.1                          # NULLs: definition levels are bit-packed / RLE runs inside the compressed page
        arrays.append(pa.array(cols[c], mask=mask))
    batch = pa.RecordBatch.from_arrays(arrays, schema=spec)
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    for rg in (8192, 30_000, 777):
        data = sstgen.write_sst(schema, batch, seq=600, cfg=WriteConfig(compression=ParquetCompression.Zstd, max_row_group_size=rg), presorted=True)
        got = eng.scan(handle, [SstInput(id=next(_ids), data=data)]).read_all()
        ref = pq.read_table(io.BytesIO(data))
        for c in ["k0"] + names:
            assert got[c].to_pylist() == ref[c].to_pylist(), (rg, c)
    # data page V2 (levels outside the compressed part) and small pages, written by pyarrow directly
    full = schema.fill_builtin_columns(batch, 601)
    for kw in (dict(data_page_version="2.0"), dict(data_page_size=4096)):
        sink = io.BytesIO()
        pq.write_table(pa.Table.from_batches([full]), sink, row_group_size=8192, use_dictionary=False, compression="zstd", write_statistics=True, **kw)
        data = sink.getvalue()
        got = eng.scan(handle, [SstInput(id=next(_ids), data=data)]).read_all()
        ref = pq.read_table(io.BytesIO(data))
        for c in ["k0"] + names:
            assert got[c].to_pylist() == ref[c].to_pylist(), (kw, c)
    eng.close()
