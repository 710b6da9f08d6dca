"""Mirror of `columnar_storage::config` (config.rs:26-172) — only what defines the SST format (S1)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional


class ParquetEncoding:  # config.rs:54-75
    Plain = "PLAIN"
    Rle = "RLE"
    DeltaBinaryPacked = "DELTA_BINARY_PACKED"
    DeltaLengthByteArray = "DELTA_LENGTH_BYTE_ARRAY"
    DeltaByteArray = "DELTA_BYTE_ARRAY"
    RleDictionary = "RLE_DICTIONARY"


class ParquetCompression:  # config.rs:79-94
    Uncompressed = "none"
    Snappy = "snappy"
    Zstd = "zstd"


@dataclass
class ColumnOptions:  # config.rs:98-103
    enable_dict: Optional[bool] = None
    enable_bloom_filter: Optional[bool] = None
    encoding: Optional[str] = None
    compression: Optional[str] = None


@dataclass
class WriteConfig:  # config.rs:107-133 (defaults 120-133)
    max_row_group_size: int = 8192
    write_bacth_size: int = 1024  # sic — the reference's spelling
    enable_sorting_columns: bool = True
    enable_dict: bool = False
    enable_bloom_filter: bool = False
    encoding: str = ParquetEncoding.Plain
    compression: str = ParquetCompression.Snappy
    column_options: Optional[Dict[str, ColumnOptions]] = None


@dataclass
class SchedulerConfig:  # config.rs:26-50 (caller of the compaction path; kept for its limits)
    memory_limit: int = 2 << 30
    new_sst_max_size: int = 1 << 30
    input_sst_max_num: int = 30
    input_sst_min_num: int = 5


@dataclass
class StorageConfig:  # config.rs:157-164
    write: WriteConfig = field(default_factory=WriteConfig)
    scheduler: SchedulerConfig = field(default_factory=SchedulerConfig)
    update_mode: int = 0
