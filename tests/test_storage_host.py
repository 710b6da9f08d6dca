"""CPU tests of the host-side mirror of `ObjectBasedStorage` (storage.rs:106-375): write path checks, SST path scheme,
file selection by time range and the per-segment scan plan — with a recording stand-in for the GPU engine."""
import os

import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from helpers import arrow_schema, record_batch
from horaedb_b200.config import StorageConfig
from horaedb_b200.storage import ObjectBasedStorage, ScanRequest, Task, WriteRequest, col, lit
from horaedb_b200.types import HoraeError, TimeRange, Timestamp


class FakeEngine:
    def __init__(self):
        self.calls = []

    def scan(self, handle, inputs, preds, projection, keep_builtin):
        self.calls.append(("scan", [i.id for i in inputs], [i.path for i in inputs], list(preds), projection, keep_builtin))
        return iter(())

    def compact(self, handle, inputs):
        self.calls.append(("compact", [i.id for i in inputs]))
        raise RuntimeError("boom")

    def compact_to_sst(self, handle, inputs, out_path, **kw):
        self.calls.append(("compact_to_sst", [i.id for i in inputs], out_path, kw))
        raise RuntimeError("boom")

    def unload_sst(self, id):
        pass


def _storage(tmp_path, seg_ms=100):
    user = arrow_schema([("pk1", "uint8"), ("pk2", "uint8"), ("value", "int64")])
    eng = FakeEngine()
    return user, eng, ObjectBasedStorage(str(tmp_path), seg_ms, user, 2, StorageConfig(), engine=eng)


def test_write_rejects_segment_crossing_and_lays_out_files(tmp_path):
    user, eng, st = _storage(tmp_path)
    b = record_batch(user, {"pk1": [3, 1, 2], "pk2": [0, 0, 0], "value": [30, 10, 20]})
    with pytest.raises(HoraeError) as ei:                                  # storage.rs:308-316
        st.write(WriteRequest(b, TimeRange(90, 110), enable_check=True))
    assert "time range can't cross segment" in str(ei.value)
    st.write(WriteRequest(b, TimeRange(90, 110), enable_check=False))      # the check is optional, as in the reference
    st.write(WriteRequest(b, TimeRange(10, 20)))
    files = st.manifest.all_ssts()
    assert len(files) == 2 and files[0].id() < files[1].id()               # ids only grow (sst.rs:35-46)
    for f in files:
        path = st.sst_path_gen.generate(f.id())
        assert path == f"{tmp_path}/data/{f.id()}.sst" and os.path.getsize(path) == f.meta().size   # sst.rs:202-204
        t = pq.read_table(path)
        assert t.schema.names == ["pk1", "pk2", "value", "__seq__", "__reserved__"]
        assert t["pk1"].to_pylist() == [1, 2, 3]                           # sorted by PK (storage.rs:244-256)
        assert t["__seq__"].to_pylist() == [f.id()] * 3 and t["__reserved__"].null_count == 3
        md = pq.ParquetFile(path).metadata
        assert md.row_group(0).column(0).compression == "SNAPPY"           # WriteConfig::default (config.rs:120-133)


def test_scan_selects_files_by_range_and_plans_per_segment(tmp_path):
    user, eng, st = _storage(tmp_path)
    b = record_batch(user, {"pk1": [1], "pk2": [0], "value": [1]})
    for rng in ((0, 10), (10, 20), (100, 150), (150, 199), (300, 310)):
        st.write(WriteRequest(b, TimeRange(*rng)))
    ids = [f.id() for f in st.manifest.all_ssts()]
    list(st.scan(ScanRequest(TimeRange.new(Timestamp(0), Timestamp(200)), [col("pk1").eq(lit(1))], None)))
    scans = [c for c in eng.calls if c[0] == "scan"]
    assert [c[1] for c in scans] == [ids[0:2], ids[2:4]]                   # one plan per segment, oldest first (storage.rs:343-358)
    assert scans[0][3] == [("pk1", "eq", 1)] and scans[0][5] is False      # predicates lowered, builtin columns stripped
    assert scans[0][2][0].endswith(f"/data/{ids[0]}.sst")
    eng.calls.clear()
    assert list(st.scan(ScanRequest(TimeRange(200, 300), [], None))) == [] and eng.calls == []   # no file overlaps: empty stream
    with pytest.raises(HoraeError):
        list(st.scan(ScanRequest(TimeRange(0, 10), ["pk1 LIKE 'x'"], None)))  # not lowerable: error, never a CPU fallback


def test_compaction_failure_releases_inputs(tmp_path):
    user, eng, st = _storage(tmp_path)
    b = record_batch(user, {"pk1": [1], "pk2": [0], "value": [1]})
    st.write(WriteRequest(b, TimeRange(0, 10)))
    st.write(WriteRequest(b, TimeRange(10, 20)))
    files = st.manifest.all_ssts()
    for f in files:
        f.mark_compaction()
    with pytest.raises(RuntimeError):
        st.do_compaction(Task(files))
    assert st.inused_memory == 0 and not any(f.is_compaction() for f in files)   # on_failure (executor.rs:123-137)
    assert len(st.manifest.all_ssts()) == 2                                      # manifest untouched
    st.config.scheduler.memory_limit = 1
    with pytest.raises(HoraeError) as ei:                                        # pre_check (executor.rs:93-114)
        st.do_compaction(Task(files))
    assert "Compaction memory usage too high" in str(ei.value)
