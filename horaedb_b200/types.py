"""Host-side mirror of the reference's `columnar_storage::types` (types.rs:35-240).

Names, argument meaning and error behaviour follow the reference so the parity tests read like
its own unit tests (`types.rs:246-302`).  Nothing here computes on the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import pyarrow as pa

# types.rs:35-41
BUILTIN_COLUMN_NUM = 2
SEQ_COLUMN_NAME = "__seq__"
RESERVED_COLUMN_NAME = "__reserved__"

I64_MAX = (1 << 63) - 1
I64_MIN = -(1 << 63)


class HoraeError(Exception):
    """Mirror of `common::Error` (common/src/error.rs:21-28)."""


def ensure(cond: bool, msg: str) -> None:
    """`ensure!` (macros.rs:36-52)."""
    if not cond:
        raise HoraeError(msg)


def _trunc_div(a: int, b: int) -> int:
    """Rust `i64 /` — truncation toward zero (types.rs:82-85 relies on it)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


@dataclass(frozen=True, order=True)
class Timestamp:
    """`Timestamp(i64)` (types.rs:46-86)."""

    value: int

    MAX = None  # filled below
    MIN = None

    def truncate_by(self, duration_ms: int) -> "Timestamp":
        # types.rs:82-85: self.0 / duration_millis * duration_millis (truncating division)
        return Timestamp(_trunc_div(self.value, duration_ms) * duration_ms)


Timestamp.MAX = Timestamp(I64_MAX)
Timestamp.MIN = Timestamp(I64_MIN)


@dataclass
class TimeRange:
    """Half-open `[start, end)` (types.rs:88-133)."""

    start: int
    end: int

    @staticmethod
    def new(start, end) -> "TimeRange":
        s = start.value if isinstance(start, Timestamp) else int(start)
        e = end.value if isinstance(end, Timestamp) else int(end)
        return TimeRange(s, e)

    def overlaps(self, other: "TimeRange") -> bool:
        # types.rs:125-127
        return self.start < other.end and other.start < self.end

    def merge(self, other: "TimeRange") -> None:
        # types.rs:129-132
        self.start = min(self.start, other.start)
        self.end = max(self.end, other.end)

    def __repr__(self) -> str:  # types.rs:91-95
        return f"[{self.start}, {self.end})"


class UpdateMode:
    """config.rs:166-172."""

    Overwrite = 0
    Append = 1


@dataclass
class StorageSchema:
    """`StorageSchema` (types.rs:143-240): user columns + `__seq__`, `__reserved__` (UInt64, nullable)."""

    arrow_schema: pa.Schema
    num_primary_keys: int
    seq_idx: int
    reserved_idx: int
    value_idxes: List[int]
    update_mode: int = UpdateMode.Overwrite

    @staticmethod
    def try_new(arrow_schema: pa.Schema, num_primary_keys: int, update_mode: int = UpdateMode.Overwrite):
        ensure(num_primary_keys > 0, "num_primary_keys should large than 0")
        ensure(
            not any(StorageSchema.is_builtin_field(f) for f in arrow_schema),
            "schema should not use builtin columns name",
        )
        value_idxes = list(range(num_primary_keys, len(arrow_schema)))
        ensure(len(value_idxes) > 0, "no value column found")
        fields = list(arrow_schema) + [
            pa.field(SEQ_COLUMN_NAME, pa.uint64(), True),
            pa.field(RESERVED_COLUMN_NAME, pa.uint64(), True),
        ]
        full = pa.schema(fields, metadata=arrow_schema.metadata)
        return StorageSchema(full, num_primary_keys, len(fields) - 2, len(fields) - 1, value_idxes, update_mode)

    @staticmethod
    def is_builtin_field(f: pa.Field) -> bool:
        return f.name in (SEQ_COLUMN_NAME, RESERVED_COLUMN_NAME)

    def fill_required_projections(self, projection: Optional[List[int]]) -> Optional[List[int]]:
        """types.rs:203-216 — appends PKs then `__seq__` (NOT `__reserved__`)."""
        if projection is None:
            return None
        for i in range(self.num_primary_keys):
            if i not in projection:
                projection.append(i)
        if self.seq_idx not in projection:
            projection.append(self.seq_idx)
        return projection

    def fill_builtin_columns(self, batch: pa.RecordBatch, sequence: int) -> pa.RecordBatch:
        """types.rs:219-239."""
        n = batch.num_rows
        if n == 0:
            return batch
        cols = list(batch.columns)
        cols.append(pa.array([sequence] * n, pa.uint64()))
        cols.append(pa.nulls(n, pa.uint64()))
        return pa.RecordBatch.from_arrays(cols, schema=self.arrow_schema)

    def user_schema(self) -> pa.Schema:
        return pa.schema(list(self.arrow_schema)[: -BUILTIN_COLUMN_NAME_COUNT], metadata=self.arrow_schema.metadata)


BUILTIN_COLUMN_NAME_COUNT = BUILTIN_COLUMN_NUM
