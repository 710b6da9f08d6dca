"""Mirror of `columnar_storage::sst` (sst.rs:35-205): file ids, `FileMeta`, `SstFile`, path scheme."""
from __future__ import annotations

import itertools
import threading
import time
from dataclasses import dataclass

from .types import TimeRange

PREFIX_PATH = "data"  # sst.rs:33

_next_id = itertools.count(time.time_ns())  # sst.rs:39-46: seeded from wall-clock ns
_id_lock = threading.Lock()


def allocate_id() -> int:
    """`SstFile::allocate_id` (sst.rs:120-122)."""
    with _id_lock:
        return next(_next_id)


@dataclass
class FileMeta:  # sst.rs:155-160 — num_rows/size are u32 in the reference
    max_sequence: int
    num_rows: int
    size: int
    time_range: TimeRange


class SstFile:  # sst.rs:51-123
    def __init__(self, id: int, meta: FileMeta):
        self._id = id
        self._meta = meta
        self._in_compaction = False

    def id(self) -> int:
        return self._id

    def meta(self) -> FileMeta:
        return self._meta

    def size(self) -> int:
        return self._meta.size

    def mark_compaction(self):
        self._in_compaction = True

    def unmark_compaction(self):
        self._in_compaction = False

    def is_compaction(self) -> bool:
        return self._in_compaction

    def is_expired(self, expire_time) -> bool:  # sst.rs:100-107
        return expire_time is not None and self._meta.time_range.end < expire_time

    def __repr__(self):
        return f"SstFile(id={self._id}, meta={self._meta}, in_compaction={self._in_compaction})"


class SstPathGenerator:  # sst.rs:193-205
    def __init__(self, prefix: str):
        self.prefix = prefix

    def generate(self, id: int) -> str:
        return f"{self.prefix}/{PREFIX_PATH}/{id}.sst"
