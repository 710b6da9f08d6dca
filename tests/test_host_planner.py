"""Host logic without a GPU: the planner's statistics pruning and the validation every scan call runs
(csrc/engine.cu: rg_may_match, validate_schema, validate_preds, prepare_sst) through hg_plan_row_groups.

Pinned semantics: DataFusion's PruningPredicate as the reference's plan text shows it (read.rs:613):
  CASE WHEN null_count = row_count THEN false ELSE <min/max rewrite of the comparison> END
which must never drop a row group that holds a matching row."""
import io

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from horaedb_b200 import sstgen
from horaedb_b200._ffi import HgError, SchemaHandle, plan_row_groups
from horaedb_b200.config import ParquetCompression, WriteConfig
from horaedb_b200.types import StorageSchema
from oracle import oracle

OPS = {"eq": lambda a, b: a == b, "ne": lambda a, b: a != b, "lt": lambda a, b: a < b, "le": lambda a, b: a <= b,
       "gt": lambda a, b: a > b, "ge": lambda a, b: a >= b}


def _expected_from_stats(data, preds, names):
    """The pinned rewrite, evaluated from pyarrow's reading of the chunk statistics."""
    md = pq.ParquetFile(io.BytesIO(data)).metadata
    out = []
    for g in range(md.num_row_groups):
        rg = md.row_group(g)
        keep = rg.num_rows > 0
        for col, op, lit in preds:
            st = rg.column(names.index(col)).statistics
            if st is None:
                continue
            if st.has_null_count and st.null_count == rg.num_rows:
                keep = False
                break
            if not st.has_min_max:
                continue
            mn, mx = st.min, st.max
            ok = {"eq": mn <= lit <= mx, "ne": mn != lit or mx != lit, "lt": mn < lit, "le": mn <= lit, "gt": mx > lit, "ge": mx >= lit}[op]
            keep = keep and ok
        out.append(int(keep))
    return out


def _rows_matching_per_rg(data, preds):
    pf = pq.ParquetFile(io.BytesIO(data))
    res = []
    for g in range(pf.metadata.num_row_groups):
        t = pf.read_row_group(g)
        m = np.ones(t.num_rows, dtype=bool)
        for col, op, lit in preds:
            v = t[col].to_numpy(zero_copy_only=False)
            valid = ~np.asarray(t[col].is_null())
            m &= valid & OPS[op](np.where(valid, v, 0), lit)
        res.append(int(m.sum()))
    return res


def test_pruning_is_exact_and_sound_on_metric_ssts():
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    names = schema.arrow_schema.names
    rng = np.random.default_rng(5)
    data, n = sstgen.synth_sst(7, 60, 900, 1000, seq=3, compression=ParquetCompression.Uncompressed)   # tags wrap 15 -> 0 inside row groups
    assert plan_row_groups(handle, data, []) == [1] * 6
    cases = [[("tag", "eq", 3)], [("tag", "ne", 3)], [("tag", "gt", 14)], [("tag", "lt", 0)], [("series_id", "le", 20)],
             [("series_id", "ge", 30), ("series_id", "lt", 31)], [("ts", "ge", sstgen.T0_MS + 899_000)], [("ts", "lt", sstgen.T0_MS)],
             [("value", "gt", 0.999999)], [("value", "lt", 0.0)], [("tag", "eq", 3), ("ts", "ge", sstgen.T0_MS + 450_000), ("value", "le", 0.5)]]
    for _ in range(20):
        cases.append([("series_id", rng.choice(list(OPS)), int(rng.integers(0, 70))), ("tag", rng.choice(list(OPS)), int(rng.integers(0, 17)))])
    for preds in cases:
        got = plan_row_groups(handle, data, preds)
        assert got == _expected_from_stats(data, preds, names), preds
        hits = _rows_matching_per_rg(data, preds)
        assert all(k == 1 for k, h in zip(got, hits) if h > 0), preds                       # soundness
        # the oracle prunes the same way: rows it decodes = rows of the kept row groups
        res = oracle.scan([data], schema.arrow_schema, 2, preds, materialize=False)
        rg_rows = [pq.ParquetFile(io.BytesIO(data)).metadata.row_group(g).num_rows for g in range(len(got))]
        assert res.rows_decoded == sum(r for r, k in zip(rg_rows, got) if k), preds
        assert res.rows_filtered == sum(hits), preds


def test_all_null_column_prunes_and_nulls_never_match():
    user = pa.schema([pa.field("k", pa.int64(), True), pa.field("t", pa.int64(), True), pa.field("v", pa.float64(), True), pa.field("w", pa.int32(), True)])
    schema = StorageSchema.try_new(user, 2)
    handle = SchemaHandle(schema.arrow_schema, 2)
    n = 3000
    w = pa.array([None] * 1000 + list(range(1000)) + [None if i % 2 else -5 for i in range(1000)], pa.int32())
    batch = pa.RecordBatch.from_arrays([pa.array(np.arange(n)), pa.array(np.arange(n) * 10), pa.array(np.linspace(0, 1, n)), w], schema=user)
    data = sstgen.write_sst(schema, batch, seq=1, cfg=WriteConfig(compression=ParquetCompression.Snappy, max_row_group_size=1000), presorted=True)
    assert plan_row_groups(handle, data, [("w", "ge", -100)]) == [0, 1, 1]          # row group 0: null_count == row_count -> false
    assert plan_row_groups(handle, data, [("w", "eq", -5)]) == [0, 0, 1]
    assert plan_row_groups(handle, data, [("w", "lt", -5)]) == [0, 0, 0]
    assert plan_row_groups(handle, data, [("__reserved__", "eq", 1)]) == [0, 0, 0]  # the builtin all-null column (types.rs:178-187)


def test_validation_errors_without_a_gpu():
    schema = sstgen.metric_storage_schema()
    handle = SchemaHandle(schema.arrow_schema, 2)
    data, _ = sstgen.synth_sst(0, 4, 100, 1000, seq=1)
    with pytest.raises(HgError):
        plan_row_groups(handle, data[: len(data) // 2], [])                          # malformed Parquet
    other = StorageSchema.try_new(pa.schema([pa.field("a", pa.int64(), True), pa.field("b", pa.int64(), True), pa.field("c", pa.int64(), True)]), 2)
    with pytest.raises(HgError):
        plan_row_groups(SchemaHandle(other.arrow_schema, 2), data, [])               # file does not have this schema
    fkey = StorageSchema.try_new(pa.schema([pa.field("a", pa.float64(), True), pa.field("b", pa.int64(), True), pa.field("c", pa.int64(), True)]), 1)
    with pytest.raises(HgError):                                                     # primary_key_eq has no float arm (read.rs:269-286)
        plan_row_groups(SchemaHandle(fkey.arrow_schema, 1), data, [])


def test_pk_splitters_balance_rows_and_are_deterministic():
    """hg_plan_pk_splitters (multi-GPU compaction): splitters from row-group statistics only; every shard's share of
    the input rows is close to 1 / parts on uniformly overlapping inputs; signed keys; skewed inputs stay a valid partition."""
    import numpy as np
    import pyarrow as pa
    from horaedb_b200 import sstgen
    from horaedb_b200._ffi import SchemaHandle, plan_pk_splitters, shard_range_preds
    from horaedb_b200.types import StorageSchema
    schema = sstgen.metric_storage_schema()
    h = SchemaHandle(schema.arrow_schema, 2)
    ssts = sstgen.synth_overlapping_ssts(8, series=400, points=200, delta_ms=1000, keep_frac=0.3, compression="none")
    datas = [s[0] for s in ssts]
    import io

    import pyarrow.parquet as pq
    sid = np.concatenate([pq.read_table(io.BytesIO(d), columns=["series_id"])["series_id"].to_numpy() for d in datas])
    for parts in (2, 4, 8):
        sp = plan_pk_splitters(h, datas, parts)
        assert sp == plan_pk_splitters(h, datas, parts) and sp == sorted(sp) and len(sp) == parts - 1
        edges = [-1] + sp + [1 << 62]
        shares = [np.count_nonzero((sid >= max(edges[i], 0)) & (sid < edges[i + 1])) / len(sid) for i in range(parts)]
        assert abs(sum(shares) - 1.0) < 1e-9 and max(shares) < 1.25 / parts, shares
        preds = [shard_range_preds(h, sp, r) for r in range(parts)]
        assert preds[0] == [("series_id", "lt", sp[0])] and preds[-1] == [("series_id", "ge", sp[-1])]
    # signed first key, heavy skew: still ordered, inside the key range
    user = pa.schema([pa.field("a", pa.int64()), pa.field("b", pa.int64()), pa.field("v", pa.float64())])
    s2 = StorageSchema.try_new(user, 2)
    a = np.sort(np.concatenate([np.full(5000, -7), np.arange(-3000, 3000)])).astype(np.int64)
    b = pa.RecordBatch.from_arrays([pa.array(a), pa.array(np.arange(len(a), dtype=np.int64)), pa.array(np.zeros(len(a)))], schema=user)
    d = sstgen.write_sst(s2, b, 4, presorted=True)
    sp = plan_pk_splitters(SchemaHandle(s2.arrow_schema, 2), [d], 4)
    assert sp == sorted(sp) and -3000 <= sp[0] and sp[-1] <= 2999
