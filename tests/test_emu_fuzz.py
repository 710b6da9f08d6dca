"""Robustness of the device decoders (snappy_core.h / zstd_core.h under the CPU warp emulator) against damaged streams and against
the order the lanes of a warp run in between two collectives.

* damaged streams: valid streams with a few bytes overwritten / bits flipped / the tail cut off.  The decoder must terminate, stay inside
  its page's scratch, and agree with the library decoder (libsnappy exactly: same accept/reject decision and, when accepted, the same
  bytes; libzstd: never accept what libzstd rejects, identical bytes when both accept — the device decoder follows RFC 8878 3.1.1.3.1.6
  "consumed exactly" for Huffman streams, where libzstd's fast path only checks the output count, so it may reject more).
* lane order: the emulator's default runs lanes 0..31 in turn; `emu_set_order` runs them 31..0 or in a fresh random permutation per
  interval.  A missing __syncwarp() between a read and an overwrite of the shared ring shows up as wrong bytes under one of the orders —
  `test_pending_ring_words_survive_a_long_literal` is the stream that found one (the tail of a long literal was stored into ring slots
  other lanes had not flushed yet)."""
import ctypes as C

import numpy as np
import pyarrow as pa
import pytest

import test_snappy_emu as S
import test_zstd_emu as Z
from test_snappy_emu import emu as snappy_emu  # noqa: F401
from test_zstd_emu import emu as zstd_emu  # noqa: F401

GUARD = 0xEE


def _damage(rng, comp, lo=0):
    comp = bytearray(comp)
    for _ in range(int(rng.integers(1, 4))):
        p = int(rng.integers(min(lo, len(comp) - 1), len(comp)))
        if rng.random() < 0.5:
            comp[p] = int(rng.integers(0, 256))
        else:
            comp[p] ^= 1 << int(rng.integers(0, 8))
    if rng.random() < 0.15:
        comp = comp[: int(rng.integers(1, len(comp)))]
    return bytes(comp)


def _stream_length(comp):
    v = sh = 0
    for x in comp[:5]:
        v |= (x & 0x7F) << sh
        sh += 7
        if not x & 0x80:
            return v
    return -1


def _snappy(lib, comp, ulen, stop_at=0xFFFFFFFF):
    out = np.full(ulen + 320, GUARD, dtype=np.uint8)
    n = C.c_long(0)
    err = lib.emu_snappy_page(comp, len(comp), out.ctypes.data, ulen, stop_at, C.byref(n))
    assert (out[ulen + 64:] == GUARD).all()                   # scratch slack of a page: 32 bytes + alignment
    return err, bytes(out[:ulen])


def _zstd(lib, comp, ulen):
    out = np.full(ulen + 320, GUARD, dtype=np.uint8)
    n = C.c_long(0)
    err = lib.emu_zstd_page(comp, len(comp), out.ctypes.data, ulen, C.byref(n))
    assert (out[ulen:] == GUARD).all()
    return err, bytes(out[:ulen])


@pytest.mark.parametrize("order", [0, 1, 2])
def test_damaged_snappy_streams_agree_with_libsnappy(snappy_emu, order):  # noqa: F811
    snappy_emu.emu_set_order(order)
    try:
        codec = pa.Codec("snappy")
        rng = np.random.default_rng(100 + order)
        names = ["jitter_ts", "mixed", "sawtooth", "tag_u32", "period3", "few_values", "short_run", "f64_cumsum"]
        accepted = 0
        for it in range(240):
            raw = S.CASES[names[it % len(names)]]
            comp = _damage(rng, codec.compress(raw, asbytes=True))
            try:
                want = codec.decompress(comp, len(raw), asbytes=True)
            except Exception:
                want = None
            if _stream_length(comp) != len(raw):
                want = None                                       # (pyarrow does not compare the stream's length prefix with the size it was given)
            err, got = _snappy(snappy_emu, comp, len(raw))
            assert (err == 0) == (want is not None), (it, err)
            if want is not None:
                assert got == want, it
                accepted += 1
        assert 40 < accepted < 200                                # both outcomes are exercised
    finally:
        snappy_emu.emu_set_order(0)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_snappy_cases_under_other_lane_orders(snappy_emu, order):  # noqa: F811
    snappy_emu.emu_set_order(order)
    try:
        codec = pa.Codec("snappy")
        for name, raw in sorted(S.CASES.items()):
            comp = codec.compress(raw, asbytes=True)
            for stop_at in (0xFFFFFFFF, 5000):
                err, got = _snappy(snappy_emu, comp, len(raw), stop_at)
                upto = min(stop_at, len(raw))
                assert err == 0 and got[:upto] == raw[:upto], (name, stop_at)
    finally:
        snappy_emu.emu_set_order(0)


def _varint(v):
    out = bytearray()
    while v >= 128:
        out.append((v & 127) | 128)
        v >>= 7
    out.append(v)
    return bytes(out)


def test_pending_ring_words_survive_a_long_literal(snappy_emu):  # noqa: F811
    # 16 literal bytes + five 64-byte copies at offset 8 leave 336 bytes pending in the ring (less than the flush threshold); the literal
    # of 4196 bytes that follows keeps its last 2048 bytes in the ring, in slots that include the pending ones
    rng = np.random.default_rng(3)
    head = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
    big = rng.integers(0, 256, 4196, dtype=np.uint8).tobytes()
    s = bytes([(16 - 1) << 2]) + head
    s += bytes([((64 - 1) << 2) | 2, 8, 0]) * 5
    s += bytes([61 << 2]) + (len(big) - 1).to_bytes(2, "little") + big
    raw = head + head[8:] * 40 + big
    comp = _varint(len(raw)) + s
    assert pa.Codec("snappy").decompress(comp, len(raw), asbytes=True) == raw
    for order in (0, 1, 2, 3):
        snappy_emu.emu_set_order(order)
        try:
            err, got = _snappy(snappy_emu, comp, len(raw))
        finally:
            snappy_emu.emu_set_order(0)
        assert err == 0 and got == raw, order


@pytest.mark.parametrize("order", [0, 1, 2])
def test_damaged_zstd_frames(zstd_emu, order):  # noqa: F811
    zstd_emu.emu_set_order(order)
    try:
        dec = pa.Codec("zstd")
        rng = np.random.default_rng(200 + order)
        names = sorted(Z.CASES)
        both = 0
        for it in range(90):
            raw = Z.CASES[names[it % len(names)]]
            comp = pa.Codec("zstd", compression_level=[1, 3, 9][it % 3]).compress(raw, asbytes=True)
            # the frame header (<= 9 bytes for these frames) stays: its Frame_Content_Size is not read, the page header's size rules
            comp = _damage(rng, comp, lo=9) if len(comp) > 12 else comp
            try:
                want = dec.decompress(comp, len(raw), asbytes=True)
                want = want if len(want) == len(raw) else None
            except Exception:
                want = None
            err, got = _zstd(zstd_emu, comp, len(raw))
            if err == 0:
                assert want is not None and got == want, it
                both += 1
        assert both > 10
    finally:
        zstd_emu.emu_set_order(0)


@pytest.mark.parametrize("order", [1, 2])
def test_zstd_cases_under_other_lane_orders(zstd_emu, order):  # noqa: F811
    zstd_emu.emu_set_order(order)
    try:
        for name, raw in sorted(Z.CASES.items()):
            comp = pa.Codec("zstd", compression_level=3).compress(raw, asbytes=True)
            err, got = _zstd(zstd_emu, comp, len(raw))
            assert err == 0 and got == raw, name
    finally:
        zstd_emu.emu_set_order(0)


@pytest.mark.parametrize("order", [0, 2])
def test_partial_decodes_of_damaged_and_cut_snappy_streams(snappy_emu, order):  # noqa: F811
    """stop_at (the consumer needs only the first rows of a page) with damage behind or before the stop position, and compressed PREFIXES
    (transient loads ship only the bytes a partial decode is expected to need): the first stop_at bytes are right, or the call reports an
    error (the engine then repeats the load with whole streams) — never wrong bytes, never a write outside the page's scratch."""
    snappy_emu.emu_set_order(order)
    try:
        codec = pa.Codec("snappy")
        rng = np.random.default_rng(300 + order)
        names = ["jitter_ts", "mixed", "sawtooth", "tag_u32", "period3", "few_values", "random", "f64_cumsum"]
        delivered = short = 0
        for it in range(360):
            raw = S.CASES[names[it % len(names)]]
            full = codec.compress(raw, asbytes=True)
            stop_at = int(rng.integers(1, len(raw)))
            if it % 3 == 1:
                comp = full[: int(rng.integers(1, len(full)))]
                err, got = _snappy(snappy_emu, comp, len(raw), stop_at)
                if err == 0:
                    assert got[:stop_at] == raw[:stop_at], it
                    delivered += 1
                else:
                    short += 1
                continue
            comp = bytearray(full)
            if it % 3 == 0:
                for _ in range(int(rng.integers(1, 3))):
                    comp[int(rng.integers(0, len(comp)))] ^= 1 << int(rng.integers(0, 8))
            comp = bytes(comp)
            try:
                want = codec.decompress(comp, len(raw), asbytes=True)
            except Exception:
                want = None
            err, got = _snappy(snappy_emu, comp, len(raw), stop_at)
            if want is not None:
                assert err == 0 and got[:stop_at] == want[:stop_at], it
        assert delivered > 20 and short > 20
    finally:
        snappy_emu.emu_set_order(0)
