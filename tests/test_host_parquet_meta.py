"""Host logic without a GPU: the library's footer / page-header reader (csrc/parquet_meta.cpp, the part of S2 that stays on
the host — in the reference parquet-rs's metadata reader behind ParquetExec, read.rs:66-93, 442-465) against pyarrow's
reading of the same SST bytes, through the C ABI (hg_parquet_inspect / hg_parquet_chunk_info)."""
import io
import struct

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from horaedb_b200 import sstgen
from horaedb_b200._ffi import HgError, parquet_chunk_info, parquet_inspect
from horaedb_b200.config import ParquetCompression, WriteConfig

PHYS = {"BOOLEAN": 0, "INT32": 1, "INT64": 2, "FLOAT": 4, "DOUBLE": 5, "BYTE_ARRAY": 6}
CODEC = {"UNCOMPRESSED": 0, "SNAPPY": 1, "ZSTD": 6}


def _plain(v, phys):
    if v is None:
        return None
    return {"INT32": lambda x: struct.pack("<i", int(x)), "INT64": lambda x: struct.pack("<q", int(x)),
            "FLOAT": lambda x: struct.pack("<f", float(x)), "DOUBLE": lambda x: struct.pack("<d", float(x))}[phys](v)


def _check_against_pyarrow(data: bytes):
    md = pq.ParquetFile(io.BytesIO(data)).metadata
    s = parquet_inspect(data)
    assert s["num_rows"] == md.num_rows and s["num_row_groups"] == md.num_row_groups and s["num_columns"] == md.num_columns
    assert s["sum_page_values"] == md.num_rows * md.num_columns          # every data page of every chunk was walked
    codecs = 0
    for g in range(md.num_row_groups):
        rg = md.row_group(g)
        for c in range(md.num_columns):
            col = rg.column(c)
            codecs |= 1 << CODEC[col.compression]
            ci = parquet_chunk_info(data, g, c)
            assert ci["num_rows"] == rg.num_rows and ci["num_values"] == col.num_values
            assert ci["data_page_offset"] == col.data_page_offset and ci["total_compressed_size"] == col.total_compressed_size
            assert ci["physical_type"] == PHYS[col.physical_type] and ci["codec"] == CODEC[col.compression]
            assert ci["num_pages"] >= 1 and ci["first_page_payload_offset"] > col.data_page_offset
            st = col.statistics
            if st is not None and st.has_null_count:
                assert ci["null_count"] == st.null_count
            if st is not None and st.has_min_max and col.physical_type in ("INT32", "INT64", "FLOAT", "DOUBLE"):
                assert ci["has_min_max"] == 1
                w = 4 if col.physical_type in ("INT32", "FLOAT") else 8
                # pyarrow hands back logical values (e.g. uint32 above 2^31): compare modulo the physical width
                lo, hi = st.min, st.max
                if col.physical_type == "INT32":
                    lo, hi = [int(np.array(x).astype(np.int64)) & 0xffffffff for x in (lo, hi)]
                    got_lo, got_hi = struct.unpack("<I", ci["min"][:4])[0], struct.unpack("<I", ci["max"][:4])[0]
                    assert (got_lo, got_hi) == (lo, hi)
                elif col.physical_type == "INT64":
                    lo, hi = [int(x) & 0xffffffffffffffff for x in (lo, hi)]
                    assert (struct.unpack("<Q", ci["min"])[0], struct.unpack("<Q", ci["max"])[0]) == (lo, hi)
                else:
                    assert ci["min"][:w] == _plain(lo, col.physical_type) and ci["max"][:w] == _plain(hi, col.physical_type)
    assert s["codec_mask"] == codecs
    return s, md


@pytest.mark.parametrize("compression", [ParquetCompression.Snappy, ParquetCompression.Uncompressed])
def test_metric_sst_metadata_matches_pyarrow(compression):
    data, n = sstgen.synth_sst(0, 40, 700, 1000, seq=9, compression=compression)      # 28 000 rows: 4 row groups, the last ragged
    s, md = _check_against_pyarrow(data)
    assert s["num_rows"] == n and s["num_row_groups"] == 4 and s["max_pages_per_chunk"] == 1   # one V1 page per chunk
    first = parquet_chunk_info(data, 0, 0)
    assert first["first_page_type"] == 0 and first["first_page_num_values"] == 8192
    if compression == ParquetCompression.Uncompressed:
        assert s["sum_compressed_bytes"] == s["sum_uncompressed_bytes"]
    else:
        assert s["sum_compressed_bytes"] < s["sum_uncompressed_bytes"]


def test_reference_vector_ssts(golden, tmp_path):
    """The SSTs the golden tests write (UInt8 / Int64 / Binary-free tables of storage.rs:391-491) parse identically."""
    from horaedb_b200.types import StorageSchema
    user = pa.schema([pa.field("pk1", pa.uint8(), True), pa.field("pk2", pa.uint8(), True), pa.field("value", pa.int64(), True)])
    schema = StorageSchema.try_new(user, 2)
    batch = pa.RecordBatch.from_arrays([pa.array([11, 11, 9, 10, 5], pa.uint8()), pa.array([11, 10, 1, 2, 3], pa.uint8()),
                                        pa.array([2, 2, 4, 22, 22], pa.int64())], schema=user)
    data = sstgen.write_sst(schema, batch, seq=1)
    s, md = _check_against_pyarrow(data)
    assert s["num_rows"] == 5 and s["num_columns"] == 5                    # + __seq__, __reserved__ (types.rs:176-187)
    assert parquet_chunk_info(data, 0, 4)["null_count"] == 5               # __reserved__ is all null


def test_small_pages_and_v2_pages():
    rng = np.random.default_rng(3)
    n = 50_000
    t = pa.table({"a": pa.array(rng.integers(0, 1 << 40, n), pa.int64()), "b": pa.array(rng.random(n)),
                  "c": pa.array(rng.integers(0, 1000, n).astype(np.uint32))})
    for version, codec in (("1.0", "snappy"), ("2.0", "none")):
        buf = io.BytesIO()
        pq.write_table(t, buf, row_group_size=20_000, data_page_size=4096, use_dictionary=False, compression=codec,
                       data_page_version=version, write_statistics=True)
        s, md = _check_against_pyarrow(buf.getvalue())
        assert s["max_pages_per_chunk"] > 1 and s["num_row_groups"] == 3
        assert parquet_chunk_info(buf.getvalue(), 0, 0)["first_page_type"] == (0 if version == "1.0" else 3)


def test_malformed_inputs_are_errors():
    data, _ = sstgen.synth_sst(0, 2, 100, 1000, seq=1, compression=ParquetCompression.Uncompressed)
    for bad in (b"", b"PAR1", data[:-1], data[: len(data) // 2], b"\x00" * 64, data[:-8] + b"\xff\xff\xff\x7fPAR1"):
        with pytest.raises(HgError):
            parquet_inspect(bad if bad else b"\x00")
    with pytest.raises(HgError):
        parquet_chunk_info(data, 99, 0)


@pytest.mark.parametrize("compression", [ParquetCompression.Uncompressed, ParquetCompression.Snappy, ParquetCompression.Zstd])
def test_damaged_files_under_sanitizers(compression, tmp_path):
    """tests/c/fuzz_parquet_meta.cpp: the reader compiled with ASan + UBSan, 4000 damaged copies of an SST per codec; any out-of-bounds
    read, overflow or leak fails the run."""
    import os
    import subprocess
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    csrc = os.path.join(root, "horaedb_b200", "csrc")
    exe = str(tmp_path / "fuzz_parquet_meta")
    subprocess.check_call(["g++", "-g", "-O1", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", "-o", exe,
                           os.path.join(root, "tests", "c", "fuzz_parquet_meta.cpp"), os.path.join(csrc, "parquet_meta.cpp"), os.path.join(csrc, "inspect.cpp")])
    data, _ = sstgen.synth_sst(0, 4, 500, 1000, seq=1, compression=compression)
    sst = tmp_path / "in.sst"
    sst.write_bytes(data)
    r = subprocess.run([exe, str(sst), "4000", "11"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    accepted, rejected = [int(x) for x in r.stdout.split()[1::2]]
    assert accepted > 500 and rejected > 500                      # both outcomes are exercised
