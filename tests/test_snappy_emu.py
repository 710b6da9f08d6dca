"""The product's warp-level Snappy decoder (horaedb_b200/csrc/snappy_core.h — the text nvcc compiles for sm_100a) executed on the
CPU: tests/emu/snappy_emu.cpp maps the 32 lanes onto coroutines and every warp collective onto a rendezvous.  Streams come from
pyarrow's Snappy (the compressor pyarrow's Parquet writer uses for the test SSTs); expected output = the bytes that were compressed.
Covers the shapes the decoder special-cases (cf. tests/test_gpu_snappy_fused.py::test_snappy_decoder_torture, which runs the same
shapes through the GPU): literal+copy value pairs with near / far sources, periodic runs of every small offset, long literals,
mixtures, tiny streams, partial decodes (stop_at) and malformed input."""
import ctypes as C
import os
import subprocess

import numpy as np
import pyarrow as pa
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "snappy_emu.cpp")
CORE = os.path.join(HERE, "..", "horaedb_b200", "csrc", "snappy_core.h")
OUT = os.path.join(HERE, "emu", "_build", "libsnappy_emu.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(CORE), os.path.getmtime(os.path.join(HERE, "emu", "warp_emu.h"))):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    lib.emu_snappy_page.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_long)]
    lib.emu_snappy_page.restype = C.c_int
    lib.emu_set_order.argtypes = [C.c_int]
    return lib


CODEC = pa.Codec("snappy")
GUARD = 0xEE


def decode(lib, comp, ulen, stop_at=0xFFFFFFFF):
    out = np.full(ulen + 320, GUARD, dtype=np.uint8)
    n = C.c_long(0)
    err = lib.emu_snappy_page(comp, len(comp), out.ctypes.data, ulen, stop_at, C.byref(n))
    return err, out


def check(lib, raw, stop_at=0xFFFFFFFF):
    comp = CODEC.compress(raw, asbytes=True)
    err, out = decode(lib, comp, len(raw), stop_at)
    assert err == 0
    upto = min(len(raw), stop_at)
    want = np.frombuffer(raw, dtype=np.uint8)
    assert np.array_equal(out[:upto], want[:upto])
    assert (out[len(raw) + 64:] == GUARD).all()          # the page's scratch slack (>= 32 bytes + alignment) is all it may touch
    if stop_at >= len(raw):
        assert np.array_equal(out[:len(raw)], want)


def _cases():
    rng = np.random.default_rng(5)
    n = 8192
    prefix = b"\x02\x00\x00\x00\x03\x10"                     # a V1 page's level prefix: values start unaligned
    ts = (1_700_000_000_000 + np.tile(np.arange(1000), 9)[:n].astype(np.int64) * 1000 + rng.integers(0, 500, n)).astype(np.int64)
    cases = {
        "jitter_ts": prefix + ts.tobytes(),                     # literal(1-2) + copy(6-7), a third of the copies 8000 bytes back
        "series_id": prefix + np.repeat(np.arange(9, dtype=np.uint64) + 77, 1000)[:n].tobytes(),
        "tag_u32": prefix + (np.repeat(np.arange(9), 1000)[:n] % 16).astype(np.uint32).tobytes(),
        "random": rng.integers(0, 2**63, n, dtype=np.uint64).tobytes(),       # incompressible: one long literal
        "sawtooth": ((np.arange(n, dtype=np.uint64) % 1000) * 37).tobytes(),
        "slow_counter": (np.arange(n // 2, dtype=np.uint64) // 3).tobytes(),
        "few_values": rng.choice(rng.integers(0, 2**60, 5, dtype=np.uint64), n // 2).tobytes(),
        "small_u32": rng.integers(0, 16, n // 2).astype(np.uint32).tobytes(),
        "f64_cumsum": np.cumsum(rng.integers(0, 1000, n // 2)).astype(np.float64).tobytes(),
        "one_byte": b"x", "three_bytes": b"abc", "short_run": b"a" * 70, "empty": b"",
    }
    for period in (1, 2, 3, 4, 5, 6, 7, 8, 12, 16, 24):
        cases[f"period{period}"] = np.resize(rng.integers(0, 256, period, dtype=np.uint8), 16384).tobytes()
    mix = rng.integers(0, 2**63, n, dtype=np.uint64)
    mix[1000:5000] = 7
    mix[6000:6100] = np.arange(100, dtype=np.uint64)
    cases["mixed"] = mix.tobytes()
    return cases


CASES = _cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_emulated_decoder_matches_input(emu, name):
    check(emu, CASES[name])


@pytest.mark.parametrize("name", ["jitter_ts", "mixed", "sawtooth", "series_id"])
@pytest.mark.parametrize("stop_at", [1, 100, 5000, 30000])
def test_emulated_partial_decode(emu, name, stop_at):
    """stop_at = the consumer needs only the page's first bytes (gate-first decompression): that prefix must be exact."""
    check(emu, CASES[name], stop_at)


def test_emulated_prefix_of_a_long_literal(emu):
    """A compressed PREFIX (transient loads) may end inside a long literal: fine when the bytes present reach stop_at, an error otherwise."""
    raw = CASES["random"]                                       # one 64 KiB literal
    comp = CODEC.compress(raw, asbytes=True)
    err, out = decode(emu, comp[:20_000], len(raw), stop_at=15_000)
    assert err == 0 and bytes(out[:15_000]) == raw[:15_000]
    err, _ = decode(emu, comp[:20_000], len(raw), stop_at=30_000)
    assert err != 0
    err, _ = decode(emu, comp[:20_000], len(raw))
    assert err != 0


def test_emulated_decoder_rejects_malformed(emu):
    raw = CASES["jitter_ts"]
    comp = CODEC.compress(raw, asbytes=True)
    err, _ = decode(emu, comp, len(raw) + 1)                   # length prefix disagrees with the page header
    assert err == 101
    err, _ = decode(emu, comp[: len(comp) // 2], len(raw))     # truncated stream
    assert err != 0
    bad = bytearray(comp)
    bad[3] = 0x01 | (7 << 2)                                   # first element becomes a copy: nothing to copy from yet
    bad[4] = 0x08
    err, _ = decode(emu, bytes(bad), len(raw))
    assert err != 0
