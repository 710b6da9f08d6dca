"""Pure-Python restatement of `MergeStream` and the two `MergeOperator`s.  TEST INFRASTRUCTURE ONLY.

Follows the reference line by line so the literal golden vectors of `test_merge_stream` (read.rs:512-573),
`test_last_value_operator` / `test_bytes_merge_operator` (operator.rs:119-159) — which use Binary columns the C
oracle does not model — can be replayed.  Small inputs only.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import pyarrow as pa

BUILTIN_COLUMN_NUM = 2  # types.rs:35


class LastValueOperator:
    """operator.rs:37-44."""

    def merge(self, batch: pa.RecordBatch) -> pa.RecordBatch:
        return batch.slice(batch.num_rows - 1, 1)


class BytesMergeOperator:
    """operator.rs:47-111: concatenate the Binary value columns of the run, first row for the others."""

    def __init__(self, value_idxes: List[int]):
        self.value_idxes = value_idxes

    def merge(self, batch: pa.RecordBatch) -> pa.RecordBatch:
        assert batch.num_rows > 0
        for idx in self.value_idxes:
            if batch.column(idx).type != pa.binary():
                raise ValueError(f"MergeOperator is only used for binary column, current:{batch.column(idx).type}")
        cols = []
        for idx, col in enumerate(batch.columns):
            if idx in self.value_idxes:
                parts = [v.as_py() or b"" for v in col]
                joined = b"".join(parts)
                if len(col) == 0 or len(joined) == 0:   # operator.rs:83-92: empty -> column returned unchanged
                    cols.append(col)
                else:
                    cols.append(pa.array([joined], pa.binary()))
            else:
                cols.append(col.slice(0, 1))
        return pa.RecordBatch.from_arrays(cols, schema=batch.schema)


_PK_EQ_TYPES = (pa.uint8(), pa.int8(), pa.uint32(), pa.int32(), pa.uint64(), pa.int64(), pa.binary())


class MergeStream:
    """read.rs:196-391."""

    def __init__(self, batches: Iterable[pa.RecordBatch], num_primary_keys: int, value_operator, keep_builtin: bool):
        self.stream = iter(batches)
        self.num_primary_keys = num_primary_keys
        self.value_operator = value_operator
        self.keep_builtin = keep_builtin
        self.pending_batch: Optional[pa.RecordBatch] = None

    def _maybe_remove_builtin_columns(self, batch: pa.RecordBatch) -> pa.RecordBatch:  # read.rs:251-260
        if self.keep_builtin:
            return batch
        for _ in range(BUILTIN_COLUMN_NUM):
            batch = batch.remove_column(batch.num_columns - 1)
        return batch

    def _primary_key_eq(self, lhs, li, rhs, ri) -> bool:  # read.rs:262-287
        for k in range(self.num_primary_keys):
            lc, rc = lhs.column(k), rhs.column(k)
            if lc.type in _PK_EQ_TYPES:
                # `.value()` ignores the validity bitmap; test data has no null PKs
                if lc[li].as_py() != rc[ri].as_py():
                    return False
            # any other type falls through -> treated as equal
        return True

    def _merge_batch(self, batch: pa.RecordBatch) -> Optional[pa.RecordBatch]:  # read.rs:289-343
        if batch.num_rows == 0:
            return None
        groups = []
        start = 0
        while start < batch.num_rows:
            end = start + 1
            while end < batch.num_rows and self._primary_key_eq(batch, start, batch, end):
                end += 1
            groups.append(batch.slice(start, end - start))
            start = end
        outputs = []
        if self.pending_batch is not None:
            pending, self.pending_batch = self.pending_batch, None
            if self._primary_key_eq(pending, pending.num_rows - 1, groups[0], 0):
                groups[0] = pa.Table.from_batches([pending, groups[0]]).combine_chunks().to_batches()[0]
            else:
                outputs.append(self.value_operator.merge(pending))
        self.pending_batch = groups.pop()
        for g in groups:
            outputs.append(self.value_operator.merge(g))
        if not outputs:
            return None
        out = pa.Table.from_batches(outputs).combine_chunks().to_batches()[0]
        return self._maybe_remove_builtin_columns(out)

    def __iter__(self):  # poll_next, read.rs:349-384
        for batch in self.stream:
            out = self._merge_batch(batch)
            if out is not None:
                yield out
        if self.pending_batch is not None:
            pending, self.pending_batch = self.pending_batch, None
            pending = self._maybe_remove_builtin_columns(pending)
            yield self.value_operator.merge(pending)
