"""The device decoders (snappy_core.h / zstd_core.h, run by the CPU warp emulator) on the PAGES of real SSTs: files written with the
reference's writer properties (sstgen.write_sst = build_write_props, storage.rs:258-298) in Snappy and Zstd, every data page's payload
cut out of the file with a few lines of Thrift-compact reading, decoded by the emulated warp and compared with libsnappy / libzstd
(pyarrow's codecs) decoding the same payload."""
import ctypes as C
import io

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from horaedb_b200 import sstgen
from horaedb_b200.config import ParquetCompression, WriteConfig

from test_snappy_emu import emu as snappy_emu  # noqa: F401  (fixtures: build + load the emulators)
from test_zstd_emu import emu as zstd_emu  # noqa: F401


def _varint(b, i):
    v = sh = 0
    while True:
        x = b[i]
        i += 1
        v |= (x & 0x7F) << sh
        sh += 7
        if not x & 0x80:
            return v, i


def _page_header(buf, pos):
    """PageHeader (parquet.thrift): fields 1 type, 2 uncompressed_page_size, 3 compressed_page_size are the first three, all i32 (compact
    type 5, field-id delta 1); everything after them up to the struct's end is skipped generically.  Returns (type, uncomp, comp, header length)."""
    i = pos
    vals = []
    for _ in range(3):
        assert buf[i] == 0x15, hex(buf[i])
        z, i = _varint(buf, i + 1)
        vals.append((z >> 1) ^ -(z & 1))

    def skip(t, i):
        if t in (1, 2):
            return i
        if t == 3:
            return i + 1
        if t in (4, 5, 6):
            return _varint(buf, i)[1]
        if t == 7:
            return i + 8
        if t == 8:
            n, i = _varint(buf, i)
            return i + n
        if t in (9, 10):
            h = buf[i]
            i += 1
            n = h >> 4
            if n == 15:
                n, i = _varint(buf, i)
            for _ in range(n):
                i = skip(h & 15, i) if (h & 15) not in (1, 2) else i + 1
            return i
        if t == 12:
            while True:
                h = buf[i]
                i += 1
                if h == 0:
                    return i
                if not h >> 4:
                    i = _varint(buf, i)[1]
                i = skip(h & 15, i)
        raise AssertionError(t)

    while True:
        h = buf[i]
        i += 1
        if h == 0:
            break
        if not h >> 4:
            i = _varint(buf, i)[1]
        i = skip(h & 15, i)
    return vals[0], vals[1], vals[2], i - pos


def _pages(data):
    md = pq.ParquetFile(io.BytesIO(data)).metadata
    for g in range(md.num_row_groups):
        for c in range(md.num_columns):
            col = md.row_group(g).column(c)
            pos, end = col.data_page_offset, col.data_page_offset + col.total_compressed_size
            if col.dictionary_page_offset:
                pos = min(pos, col.dictionary_page_offset)
            while pos < end:
                ptype, uncomp, comp, hl = _page_header(data, pos)
                yield g, c, ptype, uncomp, data[pos + hl: pos + hl + comp]
                pos += hl + comp


def _sst(compression, rng):
    n = 20_000
    sid = np.repeat(np.arange(20), 1000)
    ts = sstgen.T0_MS + np.tile(np.arange(1000) * 1000, 20) + rng.integers(0, 500, n)
    schema = sstgen.metric_storage_schema()
    batch = pa.RecordBatch.from_arrays([pa.array(sid.astype(np.uint64)), pa.array(ts.astype(np.int64)), pa.array(rng.random(n)),
                                        pa.array((sid % 16).astype(np.uint32))], schema=sstgen.METRIC_SCHEMA)
    return sstgen.write_sst(schema, batch, seq=9, cfg=WriteConfig(compression=compression), presorted=True)


def test_snappy_pages_of_an_sst(snappy_emu):  # noqa: F811
    data = _sst(ParquetCompression.Snappy, np.random.default_rng(1))
    codec = pa.Codec("snappy")
    npages = 0
    for g, c, ptype, uncomp, payload in _pages(data):
        if ptype != 0:
            continue
        want = codec.decompress(payload, uncomp, asbytes=True)
        out = np.full(uncomp + 320, 0xEE, dtype=np.uint8)
        n = C.c_long(0)
        err = snappy_emu.emu_snappy_page(payload, len(payload), out.ctypes.data, uncomp, 0xFFFFFFFF, C.byref(n))
        assert err == 0 and bytes(out[:uncomp]) == want, (g, c)
        npages += 1
    assert npages == 3 * 6                                       # 3 row groups x (4 user columns + __seq__ + __reserved__)


def test_zstd_pages_of_an_sst(zstd_emu):  # noqa: F811
    data = _sst(ParquetCompression.Zstd, np.random.default_rng(2))
    codec = pa.Codec("zstd")
    npages = 0
    for g, c, ptype, uncomp, payload in _pages(data):
        if ptype != 0:
            continue
        want = codec.decompress(payload, uncomp, asbytes=True)
        out = np.full(uncomp + 320, 0xEE, dtype=np.uint8)
        n = C.c_long(0)
        err = zstd_emu.emu_zstd_page(payload, len(payload), out.ctypes.data, uncomp, C.byref(n))
        assert err == 0 and bytes(out[:uncomp]) == want, (g, c)
        npages += 1
    assert npages == 3 * 6
