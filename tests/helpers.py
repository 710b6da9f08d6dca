"""Shared helpers for the parity tests (fixtures in the style of the reference's test_util.rs:31-165)."""
import pyarrow as pa

_T = {"uint8": pa.uint8(), "int8": pa.int8(), "uint32": pa.uint32(), "int32": pa.int32(), "uint64": pa.uint64(),
      "int64": pa.int64(), "float64": pa.float64(), "binary": pa.binary(), "uint16": pa.uint16(), "int16": pa.int16(),
      "float32": pa.float32()}


def arrow_schema(spec):
    """`arrow_schema!` (test_util.rs:31-43): every field nullable."""
    return pa.schema([pa.field(n, _T[t], True) for n, t in spec])


def record_batch(schema: pa.Schema, cols: dict) -> pa.RecordBatch:
    """`record_batch!` (test_util.rs:45-73)."""
    arrays = []
    for f in schema:
        v = cols[f.name]
        if f.type == pa.binary():
            v = [x.encode() if isinstance(x, str) else x for x in v]
        arrays.append(pa.array(v, f.type))
    return pa.RecordBatch.from_arrays(arrays, schema=schema)


def arrays_equal(a, b) -> bool:
    """Array equality that treats floats by BIT PATTERN (NaN == NaN, -0.0 != +0.0): parity means identical bytes."""
    import numpy as np
    if isinstance(a, pa.ChunkedArray):
        a = a.combine_chunks()
    if isinstance(b, pa.ChunkedArray):
        b = b.combine_chunks()
    if a.type != b.type:
        a = a.cast(b.type)
    if not pa.types.is_floating(b.type):
        return a.equals(b)
    if len(a) != len(b) or a.null_count != b.null_count:
        return False
    va, vb = a.is_valid().to_numpy(zero_copy_only=False), b.is_valid().to_numpy(zero_copy_only=False)
    if not np.array_equal(va, vb):
        return False
    it = np.uint64 if b.type == pa.float64() else np.uint32
    xa = a.fill_null(0).to_numpy(zero_copy_only=False).view(it)
    xb = b.fill_null(0).to_numpy(zero_copy_only=False).view(it)
    return bool(np.array_equal(xa[va], xb[vb]))


def check_stream(actual_batches, expected_batches):
    """`check_stream` (test_util.rs:150-165): batch-by-batch equality INCLUDING batch boundaries."""
    actual_batches = list(actual_batches)
    assert len(actual_batches) == len(expected_batches), (
        f"batch count {len(actual_batches)} != {len(expected_batches)}: "
        f"{[b.num_rows for b in actual_batches]} vs {[b.num_rows for b in expected_batches]}")
    for i, (a, e) in enumerate(zip(actual_batches, expected_batches)):
        assert a.schema.names == e.schema.names, f"batch {i} names {a.schema.names} != {e.schema.names}"
        assert a.num_rows == e.num_rows, f"batch {i} rows {a.num_rows} != {e.num_rows}"
        for c in range(a.num_columns):
            assert arrays_equal(a.column(c), e.column(c)), (
                f"batch {i} column {a.schema.names[c]}: {a.column(c)} != {e.column(c)}")
