"""ctypes front-end of the C oracle (oracle/horae_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")

# column type codes shared with horae_oracle.c
OT = {pa.uint8(): 0, pa.int8(): 1, pa.uint16(): 2, pa.int16(): 3, pa.uint32(): 4, pa.int32(): 5,
      pa.uint64(): 6, pa.int64(): 7, pa.float32(): 8, pa.float64(): 9}
_NP = {0: np.uint8, 1: np.int8, 2: np.uint16, 3: np.int16, 4: np.uint32, 5: np.int32, 6: np.uint64, 7: np.int64,
       8: np.float32, 9: np.float64}
OPS = {"eq": 0, "ne": 1, "lt": 2, "le": 3, "gt": 4, "ge": 5, "in": 6}


def widened_bits(t: pa.DataType, v) -> int:
    """A literal in the column's widened 64-bit domain: i64 / u64 two's complement, f64 bit pattern."""
    if pa.types.is_floating(t):
        return int(np.array([float(v)], dtype=np.float64).view(np.uint64)[0])
    return int(v) & 0xFFFFFFFFFFFFFFFF


class OrcPred(C.Structure):
    _fields_ = [("col", C.c_int32), ("op", C.c_int32), ("i", C.c_int64), ("u", C.c_uint64), ("f", C.c_double),
                ("in_vals", C.POINTER(C.c_uint64)), ("in_count", C.c_int32), ("_pad", C.c_int32)]


def build() -> str:
    srcs = [os.path.join(_HERE, "horae_oracle.c"), os.path.join(_HERE, "zstd_oracle.h")]
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_last_error.restype = C.c_char_p
        L.orc_table_rows.restype = C.c_int64
        L.orc_table_col.restype = C.POINTER(C.c_uint64)
        L.orc_table_valid.restype = C.POINTER(C.c_uint8)
        L.orc_scan_rows.restype = C.c_void_p
        L.orc_scan_batch_end.restype = C.POINTER(C.c_int64)
        L.orc_scan_stat.restype = C.c_int64
        L.orc_agg_ngroups.restype = C.c_int64
        L.orc_agg_stat.restype = C.c_int64
        for f, t in (("gkey", C.c_uint64), ("bucket", C.c_int64), ("count", C.c_uint64), ("sum", C.c_double),
                     ("min", C.c_double), ("max", C.c_double)):
            getattr(L, "orc_agg_" + f).restype = C.POINTER(t)
        L.orc_truncate_by.restype = C.c_int64
        L.orc_truncate_by.argtypes = [C.c_int64, C.c_int64]
        _lib = L
    return _lib


def schema_types(schema: pa.Schema) -> List[int]:
    return [OT[f.type] for f in schema]


def make_preds(schema: pa.Schema, preds: Sequence[tuple]):
    """preds: [(column name or index, op, literal)] — a conjunction (read.rs:459)."""
    arr = (OrcPred * max(len(preds), 1))()
    keep = []
    for k, (col, op, lit) in enumerate(preds):
        idx = col if isinstance(col, int) else schema.get_field_index(col)
        t = schema.field(idx).type
        arr[k].col = idx
        arr[k].op = OPS[op]
        if op == "in":
            vals = (C.c_uint64 * max(len(lit), 1))(*[widened_bits(t, v) for v in lit])
            keep.append(vals)
            arr[k].in_vals = vals
            arr[k].in_count = len(lit)
        elif pa.types.is_floating(t):
            arr[k].f = float(lit)
        elif pa.types.is_signed_integer(t):
            arr[k].i = int(lit)
        else:
            arr[k].u = int(lit)
    arr._keep = keep          # the IN lists must outlive the call
    return arr


def _table_to_arrow(L, tptr, schema: pa.Schema, ncols: int) -> pa.Table:
    n = L.orc_table_rows(C.c_void_p(tptr))
    cols = []
    for c in range(ncols):
        t = OT[schema.field(c).type]
        vp = L.orc_table_col(C.c_void_p(tptr), c)
        bp = L.orc_table_valid(C.c_void_p(tptr), c)
        if n:
            raw = np.ctypeslib.as_array(vp, shape=(n,)).copy()
            valid = np.ctypeslib.as_array(bp, shape=(n,)).astype(bool)
        else:
            raw = np.zeros(0, np.uint64)
            valid = np.zeros(0, bool)
        if t in (8, 9):
            v = raw.view(np.float64).astype(_NP[t])
        else:
            v = raw.astype(_NP[t])  # truncation of the widened slot
        cols.append(pa.array(v, type=schema.field(c).type, mask=~valid))
    return pa.Table.from_arrays(cols, schema=pa.schema([schema.field(c) for c in range(ncols)]))


def _sst_args(ssts: Sequence[bytes]):
    k = len(ssts)
    bufs = [np.frombuffer(s, dtype=np.uint8) for s in ssts]
    ptrs = (C.c_void_p * max(k, 1))(*[b.ctypes.data for b in bufs])
    lens = (C.c_uint64 * max(k, 1))(*[len(s) for s in ssts])
    return bufs, ptrs, lens


def decode_sst(data: bytes, schema: pa.Schema, preds=(), prune: bool = False) -> pa.Table:
    L = lib()
    types = (C.c_int * len(schema))(*schema_types(schema))
    p = make_preds(schema, preds)
    out = C.c_void_p()
    buf = np.frombuffer(data, dtype=np.uint8)
    rc = L.orc_decode_sst(C.c_void_p(buf.ctypes.data), C.c_uint64(len(data)), len(schema), types, p, len(preds),
                          int(prune), C.byref(out))
    if rc:
        raise RuntimeError(L.orc_last_error().decode())
    try:
        return _table_to_arrow(L, out.value, schema, len(schema))
    finally:
        L.orc_table_free(out)


@dataclass
class ScanResult:
    batches: List[pa.RecordBatch]
    rows_in_files: int
    rows_decoded: int
    rows_filtered: int
    rows_merged: int


def scan(ssts: Sequence[bytes], schema: pa.Schema, num_pk: int, preds=(), keep_builtin: bool = False,
         batch_size: int = 8192, prune: bool = True, threads: int = 1, materialize: bool = True) -> ScanResult:
    """S2→S3→S4→S5→S6 (read.rs:429-494).  `schema` is the full storage schema (with builtin columns)."""
    L = lib()
    types = (C.c_int * len(schema))(*schema_types(schema))
    p = make_preds(schema, preds)
    bufs, ptrs, lens = _sst_args(ssts)
    out = C.c_void_p()
    rc = L.orc_scan(ptrs, lens, len(ssts), len(schema), types, num_pk, p, len(preds), int(keep_builtin), batch_size,
                    int(prune), threads, C.byref(out))
    if rc:
        raise RuntimeError(L.orc_last_error().decode())
    try:
        stats = [L.orc_scan_stat(out, i) for i in range(4)]
        batches = []
        if materialize:
            oc = len(schema) if keep_builtin else len(schema) - 2
            tbl = _table_to_arrow(L, L.orc_scan_rows(out), schema, oc)
            nb = L.orc_scan_nbatches(out)
            ends = [L.orc_scan_batch_end(out)[i] for i in range(nb)]
            lo = 0
            for e in ends:
                batches.append(tbl.slice(lo, e - lo).combine_chunks().to_batches()[0])
                lo = e
        return ScanResult(batches, *stats)
    finally:
        L.orc_scan_result_free(out)


@dataclass
class AggResult:
    gkey: np.ndarray
    bucket: np.ndarray
    count: np.ndarray
    sum: np.ndarray
    min: np.ndarray
    max: np.ndarray
    rows_in_files: int
    rows_decoded: int
    rows_filtered: int
    rows_merged: int
    rows_out: int


def scan_aggregate(ssts: Sequence[bytes], schema: pa.Schema, num_pk: int, preds=(), group_col: int = 0,
                   ts_col: int = -1, window_ms: int = 0, value_col: int = -1, prune: bool = True,
                   threads: int = 1, mode: int = 0) -> AggResult:
    """mode 0: groups are runs of equal (group value, bucket) in the sorted stream; mode 1 ("hash"): true GROUP BY for any
    key, rows accumulated in stream order, groups emitted sorted by (group value, bucket)."""
    L = lib()
    types = (C.c_int * len(schema))(*schema_types(schema))
    p = make_preds(schema, preds)
    bufs, ptrs, lens = _sst_args(ssts)
    out = C.c_void_p()
    fn = L.orc_scan_aggregate_hash if mode == 1 else L.orc_scan_aggregate
    rc = fn(ptrs, lens, len(ssts), len(schema), types, num_pk, p, len(preds), int(prune), threads,
                              group_col, ts_col, C.c_int64(window_ms), value_col, C.byref(out))
    if rc:
        raise RuntimeError(L.orc_last_error().decode())
    try:
        n = L.orc_agg_ngroups(out)

        def arr(name, dt):
            ptr = getattr(L, "orc_agg_" + name)(out)
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt).copy() if n else np.zeros(0, dt)

        return AggResult(arr("gkey", np.uint64), arr("bucket", np.int64), arr("count", np.uint64),
                         arr("sum", np.float64), arr("min", np.float64), arr("max", np.float64),
                         *[L.orc_agg_stat(out, i) for i in range(5)])
    finally:
        L.orc_agg_result_free(out)


def truncate_by(ts: int, duration_ms: int) -> int:
    return lib().orc_truncate_by(ts, duration_ms)
