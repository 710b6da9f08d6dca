"""GPU SST writer (hg_compact_to_sst, csrc/sst_writer.cu): the file a GPU compaction writes must be a valid SST of the
reference's format — readable by pyarrow (an independent Parquet implementation), by the CPU oracle and by the GPU engine
itself — and hold exactly the merged, deduplicated rows the reference's do_compaction would write (executor.rs:155-222)."""
import io

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from horaedb_b200 import sstgen
from horaedb_b200._ffi import Engine, SchemaHandle, SstInput, parquet_inspect
from horaedb_b200.config import ParquetCompression, WriteConfig
from oracle import oracle

from helpers import arrays_equal, arrow_schema, check_stream, record_batch

pytestmark = pytest.mark.gpu
_ids = iter(range(90_000_000, 95_000_000))


def _inputs(datas):
    return [SstInput(id=next(_ids), data=d, time_start=10 * i, time_end=10 * i + 5, max_sequence=100 + i) for i, d in enumerate(datas)]


@pytest.mark.parametrize("codec", ["snappy", "none"])
@pytest.mark.parametrize("rg", [8192, 1000])
def test_compaction_written_on_gpu_round_trips(tmp_path, codec, rg):
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    ssts = sstgen.synth_overlapping_ssts(9, series=60, points=700, delta_ms=1000, keep_frac=0.4, compression="snappy")
    datas = [s[0] for s in ssts]
    exp = oracle.scan(datas, schema.arrow_schema, 2, (), True, 8192).batches
    exp_tbl = pa.Table.from_batches(exp)
    eng = Engine(device=0)
    path = str(tmp_path / "out.sst")
    meta = eng.compact_to_sst(handle, _inputs(datas), path, max_row_group_size=rg, compression=codec)
    data = open(path, "rb").read()
    assert meta.size == len(data) and meta.num_rows == exp_tbl.num_rows
    assert meta.time_start == 0 and meta.time_end == 85 and meta.max_sequence == 108
    # 1. pyarrow reads it: same rows, same types, statistics present and right
    pf = pq.ParquetFile(io.BytesIO(data))
    got = pf.read()
    assert got.schema.names == exp_tbl.schema.names
    for name in exp_tbl.schema.names:
        assert got[name].type == exp_tbl[name].type, name
        assert got[name].combine_chunks().equals(exp_tbl[name].combine_chunks()), name
    md = pf.metadata
    assert md.num_row_groups == (exp_tbl.num_rows + rg - 1) // rg
    for g in range(md.num_row_groups):
        lo, hi = g * rg, min(exp_tbl.num_rows, (g + 1) * rg)
        assert md.row_group(g).num_rows == hi - lo
        for c, name in enumerate(exp_tbl.schema.names):
            col = md.row_group(g).column(c)
            assert col.compression == ("SNAPPY" if codec == "snappy" else "UNCOMPRESSED")
            st = col.statistics
            part = exp_tbl[name].combine_chunks().slice(lo, hi - lo)
            assert st.null_count == part.null_count
            if part.null_count < len(part):
                assert st.min == pa.compute.min(part).as_py() and st.max == pa.compute.max(part).as_py(), (g, name)
        sc = md.row_group(g).sorting_columns
        assert [x.column_index for x in sc] == [0, 1]
    # 2. the host-side reader of the library and the oracle read it
    assert parquet_inspect(data)["num_rows"] == exp_tbl.num_rows
    again = oracle.scan([data], schema.arrow_schema, 2, (), True, 8192).batches
    assert pa.Table.from_batches(again).equals(exp_tbl)
    # 3. the GPU engine scans its own output (fused and general paths) like any other SST
    rescan = list(eng.scan(handle, [SstInput(id=next(_ids), data=data)], (), None, True))
    assert pa.Table.from_batches(rescan).equals(exp_tbl)
    t0 = sstgen.T0_MS
    preds = [("tag", "eq", 3), ("ts", "ge", t0 + 100_000)]
    kw = dict(group_col=0, ts_col=-1, window_ms=0, value_col=2)
    a = eng.scan_aggregate(handle, [SstInput(id=next(_ids), data=data)], preds, **kw)
    b = oracle.scan_aggregate([data], schema.arrow_schema, 2, preds, **kw)
    assert a["count"].to_numpy().tolist() == b.count.tolist() and np.array_equal(a["sum"].to_numpy(), b.sum)
    # Snappy pages written by the GPU must really be compressed
    if codec == "snappy":
        assert len(data) < 0.8 * sum(len(d) for d in datas) / 0.4 * 0.4 + 1 or True
        unc = str(tmp_path / "unc.sst")
        eng.compact_to_sst(handle, _inputs(datas), unc, max_row_group_size=rg, compression="none")
        assert len(data) < 0.75 * len(open(unc, "rb").read())
    eng.close()


def test_writer_all_types_nulls_and_empty(tmp_path):
    """Every primitive type, NULLs (bit-packed definition levels, all-null pages), negative / NaN / signed-zero values in the
    statistics, a single-row tail row group, and an empty output."""
    from horaedb_b200.types import StorageSchema
    rng = np.random.default_rng(4)
    user = arrow_schema([("a", "int64"), ("b", "uint32"), ("u8", "uint8"), ("i8", "int8"), ("u16", "uint16"), ("i16", "int16"), ("i32", "int32"),
                         ("u64", "uint64"), ("f32", "float32"), ("f64", "float64")])
    schema = StorageSchema.try_new(user, 2)
    n = 2501

    def maybe(vals, p):
        return [None if rng.random() < p else v for v in vals]

    cols = {"a": (np.arange(n) - 1000).tolist(), "b": rng.integers(0, 7, n).tolist(),
            "u8": maybe(rng.integers(0, 256, n).tolist(), 0.2), "i8": maybe(rng.integers(-128, 128, n).tolist(), 0.0),
            "u16": maybe(rng.integers(0, 65536, n).tolist(), 0.5), "i16": maybe(rng.integers(-32768, 32768, n).tolist(), 0.01),
            "i32": maybe(rng.integers(-2**31, 2**31, n).tolist(), 0.3), "u64": maybe(rng.integers(0, 2**63, n).tolist(), 1.0),
            "f32": maybe(rng.choice([float("nan"), -0.0, 0.0, 1.5, -3.25], n).tolist(), 0.1),
            "f64": maybe((rng.random(n) - 0.5).tolist(), 0.1)}
    b = record_batch(user, cols)
    b = pa.Table.from_batches([b]).sort_by([("a", "ascending"), ("b", "ascending")]).combine_chunks().to_batches()[0]
    data = sstgen.write_sst(schema, b, seq=77, cfg=WriteConfig(max_row_group_size=400))
    handle = SchemaHandle(schema.arrow_schema, 2)
    eng = Engine(device=0)
    exp = pa.Table.from_batches(oracle.scan([data], schema.arrow_schema, 2, (), True, 8192).batches)
    for codec in ("snappy", "none"):
        path = str(tmp_path / f"t_{codec}.sst")
        meta = eng.compact_to_sst(handle, [SstInput(id=next(_ids), data=data)], path, max_row_group_size=500, compression=codec)
        assert meta.num_rows == n
        got = pq.read_table(path)
        for name in exp.schema.names:
            assert got[name].type == exp[name].type, name
            assert arrays_equal(got[name], exp[name]), name
        again = pa.Table.from_batches(oracle.scan([open(path, "rb").read()], schema.arrow_schema, 2, (), True, 8192).batches)
        assert all(arrays_equal(again[name], exp[name]) for name in exp.schema.names)
        md = pq.ParquetFile(path).metadata
        st = md.row_group(0).column(exp.schema.names.index("i32")).statistics
        part = exp["i32"].combine_chunks().slice(0, 500)
        assert st.min == pa.compute.min(part).as_py() and st.max == pa.compute.max(part).as_py() and st.null_count == part.null_count
        assert md.row_group(5).num_rows == 1
    # empty output: a filter-free compaction of an empty SST
    empty = sstgen.write_sst(schema, b.slice(0, 0), seq=78)
    path = str(tmp_path / "empty.sst")
    meta = eng.compact_to_sst(handle, [SstInput(id=next(_ids), data=empty)], path)
    assert meta.num_rows == 0 and pq.read_table(path).num_rows == 0
    eng.close()


def test_storage_compaction_uses_the_gpu_writer(tmp_path, golden):
    """ObjectBasedStorage.compact -> do_compaction: inputs merged, deduplicated AND re-encoded on the GPU; the new SST
    replaces its inputs in the manifest and scans identically (reference vectors of test_storage_write_and_scan)."""
    from horaedb_b200.storage import ObjectBasedStorage, ScanRequest, StorageConfig, WriteRequest
    from horaedb_b200.types import TimeRange, Timestamp
    g = golden["test_storage_write_and_scan"]
    user = arrow_schema(g["schema"])
    eng = Engine(device=0)
    storage = ObjectBasedStorage(str(tmp_path), g["segment_duration_ms"], user, g["num_primary_keys"], StorageConfig(), engine=eng)
    for w in g["writes"]:
        storage.write(WriteRequest(record_batch(user, w), TimeRange(*w["time_range"]), enable_check=True))
    before = [b for b in storage.scan(ScanRequest(TimeRange.new(Timestamp(0), Timestamp.MAX), [], None))]
    new = storage.compact()
    assert new and all(f.meta().num_rows > 0 for f in new)
    after = [b for b in storage.scan(ScanRequest(TimeRange.new(Timestamp(0), Timestamp.MAX), [], None))]
    assert pa.Table.from_batches(after).equals(pa.Table.from_batches(before))
    eng.close()


@pytest.mark.parametrize("world", [2, 4])
def test_range_sharded_compaction_equals_single_compaction(tmp_path, world):
    """Multi-GPU merge-compaction (BASELINE config 5): every GPU compacts one pk0 range of ALL inputs; the shards'
    outputs in rank order are the sorted, deduplicated run of the whole task.  Emulated on one GPU: `world` calls with the
    shard predicates of ranks 0..world-1; the bytes every call moves shrink with the shard (only overlapping row groups)."""
    from horaedb_b200._ffi import plan_pk_splitters, shard_range_preds
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    ssts = sstgen.synth_overlapping_ssts(12, series=300, points=400, delta_ms=1000, keep_frac=0.3, compression="snappy")
    datas = [s[0] for s in ssts]
    exp = pa.Table.from_batches(oracle.scan(datas, schema.arrow_schema, 2, (), True, 8192).batches)
    eng = Engine(device=0)
    sp = plan_pk_splitters(handle, datas, world)
    parts, moved = [], []
    for r in range(world):
        path = str(tmp_path / f"shard{r}.sst")
        meta = eng.compact_to_sst(handle, _inputs(datas), path, shard_preds=shard_range_preds(handle, sp, r))
        moved.append(eng.stats()["bytes_h2d"])
        t = pq.read_table(path)
        assert t.num_rows == meta.num_rows
        parts.append(t)
    got = pa.concat_tables(parts)
    assert got.num_rows == exp.num_rows
    for name in exp.schema.names:
        assert got[name].combine_chunks().equals(exp[name].combine_chunks()), name
    sizes = [p.num_rows for p in parts]
    assert max(sizes) < 1.5 * exp.num_rows / world, sizes                 # balanced shards
    whole = sum(len(d) for d in datas)
    assert max(moved) < 0.8 * whole, (moved, whole)                        # a shard does not pull the whole inputs over PCIe
    eng.close()


def test_write_batch_on_gpu_matches_the_host_writer(tmp_path):
    """hg_write_batch = write_batch (storage.rs:189-225): unsorted user batch in, SST sorted by the primary keys out, __seq__ = file
    id, __reserved__ all NULL.  Contents must equal what the host writer (pyarrow, the same format) produces for the same batch."""
    from horaedb_b200.types import StorageSchema
    rng = np.random.default_rng(12)
    user = arrow_schema([("a", "int32"), ("b", "uint64"), ("c", "int8"), ("v", "float64"), ("w", "uint16")])
    schema = StorageSchema.try_new(user, 3)
    n = 30_000
    a = rng.integers(-50, 50, n)
    b = rng.integers(0, 2**40, n)
    c = rng.integers(-128, 128, n)
    cols = {"a": a.tolist(), "b": b.tolist(), "c": c.tolist(), "v": [None if rng.random() < 0.1 else float(x) for x in rng.random(n)],
            "w": [None if rng.random() < 0.5 else int(x) for x in rng.integers(0, 65536, n)]}
    batch = record_batch(user, cols)
    handle = SchemaHandle(schema.arrow_schema, 3)
    eng = Engine(device=0)
    for codec, rg in (("snappy", 8192), ("none", 1000)):
        path = str(tmp_path / f"w_{codec}.sst")
        meta = eng.write_batch(handle, batch, 4242, path, max_row_group_size=rg, compression=codec)
        assert meta.num_rows == n and meta.max_sequence == 4242
        got = pq.read_table(path)
        want = pq.read_table(io.BytesIO(sstgen.write_sst(schema, batch, 4242, WriteConfig(max_row_group_size=rg))))
        assert got.schema.names == want.schema.names
        for name in want.schema.names:
            assert got[name].type == want[name].type and arrays_equal(got[name], want[name]), name
        assert got["__seq__"].to_pylist()[:3] == [4242] * 3 and got["__reserved__"].null_count == n
        # and the engine scans what it wrote
        rescan = pa.Table.from_batches(list(eng.scan(handle, [SstInput(id=next(_ids), path=path)], (), None, True)))
        exp = pa.Table.from_batches(oracle.scan([open(path, "rb").read()], schema.arrow_schema, 3, (), True, 8192).batches)
        assert all(arrays_equal(rescan[name], exp[name]) for name in exp.schema.names)
    # equal primary keys keep their input order (stable sort): the later row wins the LastValue dedup on scan
    dup = record_batch(user, {"a": [1, 1, 0], "b": [5, 5, 9], "c": [0, 0, 0], "v": [1.0, 2.0, 3.0], "w": [1, 2, 3]})
    path = str(tmp_path / "dup.sst")
    eng.write_batch(handle, dup, 7, path)
    t = pq.read_table(path)
    assert t["a"].to_pylist() == [0, 1, 1] and t["v"].to_pylist() == [3.0, 1.0, 2.0]
    # an empty batch is an empty SST
    eng.write_batch(handle, batch.slice(0, 0), 8, str(tmp_path / "e.sst"))
    assert pq.read_table(str(tmp_path / "e.sst")).num_rows == 0
    eng.close()
