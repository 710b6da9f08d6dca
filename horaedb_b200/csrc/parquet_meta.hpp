// parquet_meta.hpp — host-side Parquet footer + page-header walk for the SST format the reference writes
// (build_write_props, storage.rs:258-298; WriteConfig::default, config.rs:120-133).
//
// In the reference this work happens inside parquet-rs (ParquetExec / DefaultParquetFileReaderFactory,
// read.rs:66-93, 456-465).  Here the host only produces a flat page table; every byte of page payload is decoded on
// the GPU (kernels.cu).  Thrift compact protocol and the FileMetaData / PageHeader field ids follow the Apache
// Parquet format specification (parquet.thrift).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace horae {

enum PhysType : int { PT_BOOLEAN = 0, PT_INT32 = 1, PT_INT64 = 2, PT_INT96 = 3, PT_FLOAT = 4, PT_DOUBLE = 5, PT_BYTE_ARRAY = 6, PT_FLBA = 7 };
enum Codec : int { CODEC_UNCOMPRESSED = 0, CODEC_SNAPPY = 1, CODEC_ZSTD = 6 };
enum Encoding : int { ENC_PLAIN = 0, ENC_PLAIN_DICT = 2, ENC_RLE = 3, ENC_DELTA_BINARY_PACKED = 5, ENC_DELTA_LENGTH_BYTE_ARRAY = 6, ENC_RLE_DICT = 8 };
enum PageType : int { PAGE_DATA = 0, PAGE_INDEX = 1, PAGE_DICT = 2, PAGE_DATA_V2 = 3 };

struct ColumnStats {
  bool has_min = false, has_max = false, has_null_count = false;
  uint8_t min[8] = {0}, max[8] = {0};
  int64_t null_count = 0;
};

struct PageMeta {
  uint64_t payload_off = 0;  // file offset of the first byte after the page header
  uint32_t comp_size = 0, uncomp_size = 0, num_values = 0;
  uint32_t v2_def_len = 0, v2_rep_len = 0;
  uint8_t page_type = 0, encoding = 0, v2_compressed = 1;
};

struct ChunkMeta {
  int phys_type = 0, codec = 0;
  int64_t num_values = 0, data_page_offset = 0, dict_page_offset = -1, total_compressed = 0;
  ColumnStats stats;
  uint32_t first_page = 0, num_pages = 0;   // into FileMetaData::pages (data pages only)
  uint64_t scratch_bytes = 0;               // bytes of decompression scratch this chunk needs
  bool has_dict_page = false;
  uint64_t dict_payload_off = 0;            // dictionary page (RLE_DICTIONARY chunks): PLAIN values
  uint32_t dict_comp_size = 0, dict_uncomp_size = 0, dict_num_values = 0;
};

struct RowGroupMeta {
  int64_t num_rows = 0, first_row = 0;
  std::vector<ChunkMeta> cols;
};

struct FileMetaData {
  int ncols = 0;
  std::vector<int> repetition;  // per leaf: 0 required, 1 optional
  std::vector<int> phys_types;
  std::vector<std::string> names;
  int64_t num_rows = 0;
  std::vector<RowGroupMeta> rgs;
  std::vector<PageMeta> pages;
};

// Parses the footer and walks every column chunk's page headers.  Returns false and fills *err on malformed input.
bool parse_parquet(const uint8_t* data, size_t len, FileMetaData* out, std::string* err);

inline uint64_t page_scratch_bytes(uint32_t uncomp) { return ((uint64_t)uncomp + 15u) / 16u * 16u + 32u; }

}  // namespace horae
