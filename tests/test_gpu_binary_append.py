"""Binary value columns and UpdateMode::Append on the GPU path: BytesMergeOperator (operator.rs:47-111) concatenates the Binary
values of a primary-key run in (pk, seq) order and takes the other columns from the run's first row; LastValueOperator keeps
the last row.  Oracle: pyarrow decodes the SSTs (pinned independent decoder), the rows are merged by (pk.., __seq__), cut into
8192-row batches like SortPreservingMergeExec's output, and fed to the line-by-line Python `MergeStream`
(oracle/merge_stream.py, which replays the reference's own test_merge_stream / test_bytes_merge_operator vectors)."""
import io

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from horaedb_b200 import sstgen
from horaedb_b200._ffi import Engine, HgError, SchemaHandle, SstInput
from horaedb_b200.config import ParquetCompression, WriteConfig
from horaedb_b200.types import StorageSchema, UpdateMode
from oracle.merge_stream import BytesMergeOperator, LastValueOperator, MergeStream

from helpers import arrays_equal, arrow_schema, record_batch

pytestmark = pytest.mark.gpu
_ids = iter(range(95_000_000, 99_000_000))


def _reference_scan(schema: StorageSchema, datas, append: bool, keep_builtin: bool, batch_size=8192):
    tables = [pq.read_table(io.BytesIO(d)) for d in datas if pq.read_metadata(io.BytesIO(d)).num_rows]
    if not tables:
        return []
    names = schema.arrow_schema.names
    if len(tables) == 1:
        # ParquetExec hands MergeExec the reader's batches: <= batch_size rows, never across a row group
        md = pq.ParquetFile(io.BytesIO([d for d in datas if pq.read_metadata(io.BytesIO(d)).num_rows][0])).metadata
        t = tables[0]
        batches, row = [], 0
        for g in range(md.num_row_groups):
            n = md.row_group(g).num_rows
            for lo in range(0, n, batch_size):
                batches.append(t.slice(row + lo, min(batch_size, n - lo)).combine_chunks().to_batches()[0])
            row += n
    else:
        allrows = pa.concat_tables(tables).combine_chunks()
        src = np.concatenate([np.full(t.num_rows, i) for i, t in enumerate(tables)])
        pos = np.concatenate([np.arange(t.num_rows) for t in tables])
        keys = [pos, src, allrows["__seq__"].to_numpy()] + [allrows[names[k]].to_numpy() for k in reversed(range(schema.num_primary_keys))]
        order = np.lexsort(keys)                    # (pk.., seq, stream, position): SortPreservingMergeExec's total order
        merged = allrows.take(pa.array(order))
        batches = [merged.slice(lo, batch_size).combine_chunks().to_batches()[0] for lo in range(0, merged.num_rows, batch_size)]
    op = BytesMergeOperator(schema.value_idxes) if append else LastValueOperator()
    return list(MergeStream(batches, schema.num_primary_keys, op, keep_builtin))


def _check(got, exp):
    assert len(got) == len(exp), ([b.num_rows for b in got], [b.num_rows for b in exp])
    for a, e in zip(got, exp):
        assert a.schema.names == e.schema.names and a.num_rows == e.num_rows
        for c in range(a.num_columns):
            if not arrays_equal(a.column(c), e.column(c)):
                ga, ea = a.column(c).to_pylist(), e.column(c).to_pylist()
                bad = [i for i in range(len(ea)) if ga[i] != ea[i]]
                raise AssertionError(f"column {a.schema.names[c]}: {len(bad)} of {len(ea)} rows differ, first {[(i, ga[i], ea[i]) for i in bad[:4]]}")


def test_reference_merge_stream_vectors_through_ssts(golden):
    """read.rs:512-573: the literal rows of test_merge_stream, one SST per __seq__ value, scanned with both operators."""
    g = golden["test_merge_stream"]
    user = arrow_schema([("pk1", "uint8"), ("value", "binary")])
    rows = [(pk, v.encode(), s) for b in g["input_batches"] for pk, v, s in zip(b["pk1"], b["value"], b["__seq__"])]
    eng = Engine(device=0)
    for append, key in ((False, "expected_last_value"), (True, "expected_bytes_merge")):
        schema = StorageSchema.try_new(user, 1, UpdateMode.Append if append else UpdateMode.Overwrite)
        datas = [sstgen.write_sst(schema, record_batch(user, {"pk1": [pk], "value": [v]}), seq=s) for pk, v, s in rows]
        handle = SchemaHandle(schema.arrow_schema, 1, UpdateMode.Append if append else UpdateMode.Overwrite)
        got = pa.Table.from_batches(list(eng.scan(handle, [SstInput(id=next(_ids), data=d) for d in datas])))
        want_pk = [x for b in g[key] for x in b["pk1"]]
        want_v = [x.encode() for b in g[key] for x in b["value"]]
        assert got["pk1"].to_pylist() == want_pk and got["value"].to_pylist() == want_v
    eng.close()


@pytest.mark.parametrize("codec", [ParquetCompression.Snappy, ParquetCompression.Uncompressed])
def test_binary_columns_overwrite_and_append_match_merge_stream(codec):
    rng = np.random.default_rng(31)
    user = arrow_schema([("pk1", "uint64"), ("pk2", "int32"), ("blob", "binary"), ("idx", "binary")])

    def make(nrows, seq, keyspace, f=0, min_len=0):
        pk1 = np.sort(rng.integers(0, keyspace, nrows))
        pk2 = rng.integers(-2, 3, nrows)
        order = np.lexsort((pk2, pk1))
        pk1, pk2 = pk1[order], pk2[order]
        keep = np.ones(nrows, bool)
        keep[1:] = (pk1[1:] != pk1[:-1]) | (pk2[1:] != pk2[:-1])          # no duplicate PKs inside a file
        pk1, pk2 = pk1[keep], pk2[keep]
        n = len(pk1)
        # a key is NULL in at most one file ((pk1 + f) % 7): a run of SEVERAL rows never has zero bytes in Append mode, which the
        # reference cannot assemble (operator.rs:80-92, tested apart in test_append_mode_rules); one-row NULL runs do occur
        blob = [None if (int(k) + f) % 7 == 0 else rng.bytes(int(rng.integers(min_len, 40))) for k in pk1]
        idx = [bytes([seq % 251]) * int(rng.integers(1, 5)) for _ in range(n)]
        return record_batch(user, {"pk1": pk1.tolist(), "pk2": pk2.tolist(), "blob": blob, "idx": idx})

    eng = Engine(device=0)
    for append in (False, True):
        mode = UpdateMode.Append if append else UpdateMode.Overwrite
        schema = StorageSchema.try_new(user, 2, mode)
        handle = SchemaHandle(schema.arrow_schema, 2, mode)
        ml = 1 if append else 0                                                          # Overwrite also sees empty (non-NULL) values
        cases = [[make(3000, 5, 400, 0, 0)],                                                      # one file, several row groups
                 [make(2500, 10 + f, 300, f, ml) for f in range(5)],                              # overlapping files: real merge + runs
                 [make(9000, 20 + f, 2000, f, ml) for f in range(3)] + [make(0, 30, 10)]]         # > 8192 merged rows: MergeStream carry; an empty file
        for batches in cases:
            datas = [sstgen.write_sst(schema, b, seq=100 + i, cfg=WriteConfig(compression=codec, max_row_group_size=1000), presorted=True)
                     for i, b in enumerate(batches)]
            for keep_builtin in (False, True):
                got = list(eng.scan(handle, [SstInput(id=next(_ids), data=d) for d in datas], (), None, keep_builtin))
                exp = _reference_scan(schema, datas, append, keep_builtin)
                _check(got, exp)
            # a predicate on a fixed-width column in front of the merge (filter first, read.rs:459-480): same as scanning pre-filtered files
            got = list(eng.scan(handle, [SstInput(id=next(_ids), data=d) for d in datas], [("pk2", "ge", 0)], None, False))
            filt = []
            for d in datas:
                t = pq.read_table(io.BytesIO(d))
                t = t.filter(pa.compute.greater_equal(t["pk2"], 0)).combine_chunks()
                b = t.to_batches()[0] if t.num_rows else pa.RecordBatch.from_arrays([pa.array([], f.type) for f in schema.arrow_schema], schema=schema.arrow_schema)
                filt.append(sstgen.write_sst_with_seq(schema, b, WriteConfig(compression=codec, max_row_group_size=1000)))
            exp_tbl = pa.Table.from_batches(_reference_scan(schema, filt, append, False), schema=schema.user_schema())
            got_tbl = pa.Table.from_batches(got, schema=schema.user_schema())
            assert got_tbl.num_rows == exp_tbl.num_rows and exp_tbl.num_rows > 0
            for name in exp_tbl.schema.names:
                assert arrays_equal(got_tbl[name], exp_tbl[name]), name
    eng.close()


@pytest.mark.parametrize("codec", [ParquetCompression.Snappy, ParquetCompression.Uncompressed, ParquetCompression.Zstd])
def test_delta_length_byte_array_binary_columns(codec):
    """ParquetEncoding::DeltaLengthByteArray on the Binary value columns (config.rs:54-75): the lengths are a DELTA_BINARY_PACKED run in
    front of the concatenated bytes.  Same data, same expectation as the PLAIN case (pyarrow decodes the reference stream)."""
    from horaedb_b200.config import ColumnOptions
    rng = np.random.default_rng(47)
    user = arrow_schema([("pk1", "uint64"), ("pk2", "int32"), ("blob", "binary"), ("idx", "binary")])

    def make(nrows, seq, keyspace, f):
        pk1 = np.sort(rng.integers(0, keyspace, nrows))
        pk2 = rng.integers(-2, 3, nrows)
        order = np.lexsort((pk2, pk1))
        pk1, pk2 = pk1[order], pk2[order]
        keep = np.ones(nrows, bool)
        keep[1:] = (pk1[1:] != pk1[:-1]) | (pk2[1:] != pk2[:-1])
        pk1, pk2 = pk1[keep], pk2[keep]
        blob = [None if (int(k) + f) % 7 == 0 else rng.bytes(int(rng.integers(1, 60))) for k in pk1]
        idx = [bytes([seq % 251]) * int(rng.integers(1, 5)) for _ in range(len(pk1))]
        return record_batch(user, {"pk1": pk1.tolist(), "pk2": pk2.tolist(), "blob": blob, "idx": idx})

    opts = {"blob": ColumnOptions(encoding="DELTA_LENGTH_BYTE_ARRAY"), "idx": ColumnOptions(encoding="DELTA_LENGTH_BYTE_ARRAY")}
    eng = Engine(device=0)
    for append in (False, True):
        mode = UpdateMode.Append if append else UpdateMode.Overwrite
        schema = StorageSchema.try_new(user, 2, mode)
        handle = SchemaHandle(schema.arrow_schema, 2, mode)
        for batches, rg in (([make(3000, 5, 400, 0)], 1000), ([make(2500, 10 + f, 300, f) for f in range(4)], 700)):
            datas = [sstgen.write_sst(schema, b, seq=100 + i, cfg=WriteConfig(compression=codec, max_row_group_size=rg, column_options=opts), presorted=True)
                     for i, b in enumerate(batches)]
            md = pq.ParquetFile(io.BytesIO(datas[0])).metadata
            assert "DELTA_LENGTH_BYTE_ARRAY" in md.row_group(0).column(2).encodings
            for keep_builtin in (False, True):
                got = list(eng.scan(handle, [SstInput(id=next(_ids), data=d) for d in datas], (), None, keep_builtin))
                _check(got, _reference_scan(schema, datas, append, keep_builtin))
    eng.close()


def test_append_mode_rules():
    """read.rs:485-490 + operator.rs:66-73: Append merges EVERY value column, which must be Binary; Binary keys, predicates on
    Binary columns, aggregation and the GPU writer refuse Binary / Append (error codes, no fallback)."""
    user = arrow_schema([("pk1", "uint64"), ("v", "int64")])
    schema = StorageSchema.try_new(user, 1)
    data = sstgen.write_sst(schema, record_batch(user, {"pk1": [1, 2], "v": [5, 6]}), seq=1)
    eng = Engine(device=0)
    with pytest.raises(HgError) as ei:
        list(eng.scan(SchemaHandle(schema.arrow_schema, 1, UpdateMode.Append), [SstInput(id=next(_ids), data=data)]))
    assert "only used for binary column" in str(ei.value)
    ub = arrow_schema([("pk1", "uint64"), ("b", "binary")])
    sb = StorageSchema.try_new(ub, 1)
    db = sstgen.write_sst(sb, record_batch(ub, {"pk1": [1, 2], "b": [b"x", b"yy"]}), seq=1)
    hb = SchemaHandle(sb.arrow_schema, 1)
    assert pa.Table.from_batches(list(eng.scan(hb, [SstInput(id=next(_ids), data=db)])))["b"].to_pylist() == [b"x", b"yy"]
    with pytest.raises(HgError):
        list(eng.scan(hb, [SstInput(id=next(_ids), data=db)], [("b", "eq", 1)]))
    with pytest.raises(HgError):
        eng.scan_aggregate(hb, [SstInput(id=next(_ids), data=db)], [], group_col=0, value_col=1)
    with pytest.raises(HgError):
        eng.compact_to_sst(hb, [SstInput(id=next(_ids), data=db)], "/tmp/never_written.sst")
    # operator.rs:80-92: a run without bytes returns its column unchanged — one row keeps its validity, several rows cannot form the
    # one-row batch: the reference fails with "failed to construct RecordBatch in BytesMergeOperator", and so does this path
    sa = StorageSchema.try_new(ub, 1, UpdateMode.Append)
    ha = SchemaHandle(sa.arrow_schema, 1, UpdateMode.Append)
    one = [sstgen.write_sst(sa, record_batch(ub, {"pk1": [1, 2, 3], "b": [None, b"", b"q"]}), seq=1)]
    assert pa.Table.from_batches(list(eng.scan(ha, [SstInput(id=next(_ids), data=one[0])])))["b"].to_pylist() == [None, b"", b"q"]
    assert [r for b in _reference_scan(sa, one, True, False) for r in b.column(1).to_pylist()] == [None, b"", b"q"]
    two = one + [sstgen.write_sst(sa, record_batch(ub, {"pk1": [1, 3], "b": [b"", b"r"]}), seq=2)]
    with pytest.raises(HgError) as ei:
        list(eng.scan(ha, [SstInput(id=next(_ids), data=d) for d in two]))
    assert ei.value.code == 1 and "failed to construct RecordBatch in BytesMergeOperator" in str(ei.value)
    with pytest.raises(Exception):
        _reference_scan(sa, two, True, False)
    eng.close()
