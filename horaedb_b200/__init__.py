"""horaedb_b200 — B200-native (sm_100a) implementation of HoraeDB's columnar hot path:
SST scan (Parquet page decode -> filter -> k-way merge -> LastValue dedup), time-bucket aggregation and
merge-compaction, behind the reference's `ColumnarStorage` surface (src/columnar_storage/src/storage.rs:76-87).

The data path lives in csrc/ (CUDA + C ABI, include/horae_gpu.h); this package is the host-side mirror used by the
parity tests and bench.py.  There is no CPU fallback: importing the engine without the built library raises.
"""
from .types import (BUILTIN_COLUMN_NUM, RESERVED_COLUMN_NAME, SEQ_COLUMN_NAME, HoraeError, StorageSchema, TimeRange,
                    Timestamp, UpdateMode)
from .config import (ColumnOptions, ParquetCompression, ParquetEncoding, SchedulerConfig, StorageConfig, WriteConfig)
from .sst import FileMeta, SstFile, SstPathGenerator, allocate_id

__all__ = ["BUILTIN_COLUMN_NUM", "RESERVED_COLUMN_NAME", "SEQ_COLUMN_NAME", "HoraeError", "StorageSchema", "TimeRange",
           "Timestamp", "UpdateMode", "ColumnOptions", "ParquetCompression", "ParquetEncoding", "SchedulerConfig",
           "StorageConfig", "WriteConfig", "FileMeta", "SstFile", "SstPathGenerator", "allocate_id"]
