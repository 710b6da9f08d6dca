"""CPU tests: the oracle and the host-side mirror replay the reference's own unit-test vectors."""
import numpy as np
import pyarrow as pa
import pytest

from helpers import arrow_schema, check_stream, record_batch
from horaedb_b200 import sstgen
from horaedb_b200.config import ParquetCompression, WriteConfig
from horaedb_b200.types import (RESERVED_COLUMN_NAME, SEQ_COLUMN_NAME, HoraeError, StorageSchema, TimeRange, Timestamp,
                                UpdateMode)
from oracle import oracle
from oracle.merge_stream import BytesMergeOperator, LastValueOperator, MergeStream


def test_timestamp_truncate_by(golden):  # types.rs:246-261
    for ts, seg, expected in golden["test_timestamp_truncate_by"]["cases"]:
        assert Timestamp(ts).truncate_by(seg).value == expected
        assert oracle.truncate_by(ts, seg) == expected
    # Rust i64 division truncates toward zero (types.rs:82-85) — not floor
    assert Timestamp(-10).truncate_by(20).value == 0
    assert oracle.truncate_by(-10, 20) == 0
    assert Timestamp(-30).truncate_by(20).value == -20


def test_time_range_overlaps(golden):  # types.rs:125-127
    for a, b, exp in golden["test_pick_candidate_time_range_overlaps"]["cases"]:
        assert TimeRange(*a).overlaps(TimeRange(*b)) is exp


def test_build_storage_schema(golden):  # types.rs:263-302
    g = golden["test_build_storage_schema"]
    user = arrow_schema(g["schema"])
    schema = StorageSchema.try_new(user, g["num_primary_keys"], UpdateMode.Append)
    assert schema.value_idxes == g["value_idxes"]
    assert schema.seq_idx == g["seq_idx"] and schema.reserved_idx == g["reserved_idx"]
    with pytest.raises(HoraeError):
        StorageSchema.try_new(user, 3, UpdateMode.Append)  # no value column
    batch = record_batch(user, g["batch"])
    nb = schema.fill_builtin_columns(batch, g["sequence"])
    assert nb.schema.names == ["pk1", "pk2", "value", SEQ_COLUMN_NAME, RESERVED_COLUMN_NAME]
    assert nb.column(3).to_pylist() == [g["sequence"]] * 4
    assert nb.column(4).null_count == 4 and nb.column(4).type == pa.uint64()
    for inp, exp in g["projections"]:
        assert schema.fill_required_projections(inp) == exp


def test_last_value_operator(golden):  # operator.rs:119-137
    g = golden["test_last_value_operator"]
    sch = arrow_schema([("pk1", "uint8"), ("pk2", "uint8"), ("value", "int64")])
    out = LastValueOperator().merge(record_batch(sch, g["input"]))
    assert out.equals(record_batch(sch, g["expected"]))


def test_bytes_merge_operator(golden):  # operator.rs:139-159
    g = golden["test_bytes_merge_operator"]
    sch = arrow_schema([("pk1", "uint8"), ("pk2", "uint8"), ("value", "binary")])
    out = BytesMergeOperator(g["value_idxes"]).merge(record_batch(sch, g["input"]))
    assert out.equals(record_batch(sch, g["expected"]))


def _merge_stream_inputs(g):
    sch = arrow_schema([("pk1", "uint8"), ("value", "binary"), (SEQ_COLUMN_NAME, "uint8"), (RESERVED_COLUMN_NAME, "uint8")])
    ins = []
    for b in g["input_batches"]:
        cols = dict(b)
        cols[RESERVED_COLUMN_NAME] = [None] * len(b["pk1"])
        ins.append(record_batch(sch, cols))
    return ins


def test_merge_stream(golden):  # read.rs:512-573
    g = golden["test_merge_stream"]
    out_sch = arrow_schema([("pk1", "uint8"), ("value", "binary")])
    for op, key in ((LastValueOperator(), "expected_last_value"), (BytesMergeOperator([1]), "expected_bytes_merge")):
        stream = MergeStream(_merge_stream_inputs(g), g["num_primary_keys"], op, keep_builtin=False)
        check_stream(stream, [record_batch(out_sch, e) for e in g[key]])


def test_storage_sort_batch(golden):  # storage.rs:493-536
    g = golden["test_storage_sort_batch"]
    sch = arrow_schema([("a", "uint8"), ("b", "uint8"), ("c", "uint8"), ("d", "uint8")])
    schema = StorageSchema.try_new(sch, 1)
    out = sstgen.sort_batch(schema, record_batch(sch, g["input"]))
    assert out.equals(record_batch(sch, g["expected"]))


def _write_and_scan_ssts(g, compression):
    user = arrow_schema(g["schema"])
    schema = StorageSchema.try_new(user, g["num_primary_keys"])
    ssts = []
    for i, w in enumerate(g["writes"]):
        batch = record_batch(user, w)
        ssts.append(sstgen.write_sst(schema, batch, seq=1000 + i, cfg=WriteConfig(compression=compression)))
    return user, schema, ssts


@pytest.mark.parametrize("compression", [ParquetCompression.Snappy, ParquetCompression.Uncompressed])
def test_storage_write_and_scan(golden, compression):  # storage.rs:391-491
    g = golden["test_storage_write_and_scan"]
    user, schema, ssts = _write_and_scan_ssts(g, compression)
    res = oracle.scan(ssts, schema.arrow_schema, schema.num_primary_keys)
    check_stream(res.batches, [record_batch(user, e) for e in g["scan_all_expected"]])
    res = oracle.scan(ssts, schema.arrow_schema, schema.num_primary_keys, preds=[("pk1", "eq", 11)])
    check_stream(res.batches, [record_batch(user, e) for e in g["scan_pk1_eq_11_expected"]])
    # keep_builtin (compaction's view of the same plan, executor.rs:164-169): builtin columns retained
    res = oracle.scan(ssts, schema.arrow_schema, schema.num_primary_keys, keep_builtin=True)
    assert res.batches[0].schema.names[-2:] == [SEQ_COLUMN_NAME, RESERVED_COLUMN_NAME]
    assert res.batches[0].column(3).to_pylist() == [1000, 1001, 1001, 1001]


def test_merge_stream_c_oracle_matches_python(golden):
    """The C oracle's MergeStream restatement and the line-by-line Python one agree on batch boundaries."""
    rng = np.random.default_rng(7)
    user = arrow_schema([("pk1", "uint8"), ("pk2", "int64"), ("value", "int64")])
    schema = StorageSchema.try_new(user, 2)
    ssts, tables = [], []
    for f in range(4):
        n = 300
        b = record_batch(user, {"pk1": rng.integers(0, 6, n).tolist(), "pk2": rng.integers(-5, 5, n).tolist(),
                                "value": rng.integers(0, 1000, n).tolist()})
        # intra-file duplicate PKs have unspecified order in the reference: drop them
        tbl = pa.Table.from_batches([b]).group_by(["pk1", "pk2"]).aggregate([("value", "max")]).rename_columns(["pk1", "pk2", "value"])
        b = tbl.combine_chunks().to_batches()[0]
        b = pa.RecordBatch.from_arrays([b.column(0), b.column(1), b.column(2)], schema=user)
        ssts.append(sstgen.write_sst(schema, b, seq=50 + f, cfg=WriteConfig(max_row_group_size=64)))
    for bs in (7, 64, 8192):
        res = oracle.scan(ssts, schema.arrow_schema, 2, batch_size=bs)
        # python restatement over the same merged stream
        full = pa.concat_tables([oracle.decode_sst(s, schema.arrow_schema) for s in ssts])
        idx = pa.compute.sort_indices(full, sort_keys=[("pk1", "ascending"), ("pk2", "ascending"), (SEQ_COLUMN_NAME, "ascending")])
        merged = full.take(idx).combine_chunks()
        chunks = [merged.slice(o, bs).combine_chunks().to_batches()[0] for o in range(0, merged.num_rows, bs)]
        expected = list(MergeStream(chunks, 2, LastValueOperator(), keep_builtin=False))
        check_stream(res.batches, expected)
