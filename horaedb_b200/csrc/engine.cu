// engine.cu — host side of libhorae_gpu.so: SST residency, scan planning, pipeline orchestration, Arrow C export
// and the C ABI declared in include/horae_gpu.h.
//
// The planner mirrors ParquetReader::build_df_plan (read.rs:429-494):
//   ParquetExec (row-group pruning by chunk statistics)  -> FilterExec -> SortPreservingMergeExec -> MergeExec
// but runs every data-touching step as CUDA kernels (kernels.cu / fused_scan.cu).  There is no CPU fallback: if the
// device library cannot do something it returns HG_ERR_UNSUPPORTED.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "engine_internal.h"
#include "fused_scan.h"
#include "sst_writer.h"

thread_local Arena* g_arena = nullptr;
static thread_local std::string g_last_error;
int set_error(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

// A Snappy stream that holds only literals is the page's bytes behind a few header bytes: incompressible columns (random
// f64 values) are written as one literal per 64 KiB block.  Returns true when the single V1 page of the chunk is such a
// stream with at most two literals, the level prefix lies inside the first one and the value bytes of both are whole
// values, i.e. row i can be addressed in place: the fused scan then never decompresses the column.
static bool classify_stored(const uint8_t* data, uint64_t size, const PageMeta& pm, bool optional, uint32_t width, uint64_t rows) {
  const uint8_t* p = data + pm.payload_off;
  const uint8_t* end = p + pm.comp_size;
  if (pm.payload_off + pm.comp_size > size) return false;
  uint64_t ulen = 0;
  int sh = 0;
  for (;;) {
    if (p >= end || sh > 28) return false;
    const uint8_t b = *p++;
    ulen |= uint64_t(b & 0x7f) << sh;
    sh += 7;
    if (!(b & 0x80)) break;
  }
  if (ulen != pm.uncomp_size) return false;
  uint64_t lens[2] = {0, 0};
  const uint8_t* lit[2] = {nullptr, nullptr};
  int n = 0;
  uint64_t total = 0;
  while (p < end) {
    if (n == 2) return false;
    const uint8_t t = *p;
    if (t & 3) return false;                       // a copy element: real compression
    uint64_t len = t >> 2;
    uint32_t hdr = 1;
    if (len >= 60) {
      const uint32_t nb = uint32_t(len) - 59;
      if (p + 1 + nb > end) return false;
      len = 0;
      for (uint32_t i = 0; i < nb; i++) len |= uint64_t(p[1 + i]) << (8 * i);
      hdr = 1 + nb;
    }
    len += 1;
    if (p + hdr + len > end) return false;
    lit[n] = p + hdr;
    lens[n] = len;
    n++;
    total += len;
    p += hdr + len;
  }
  if (n == 0 || total != ulen) return false;
  uint64_t prefix = 0;
  if (optional) {
    if (lens[0] < 4) return false;
    uint32_t dl;
    std::memcpy(&dl, lit[0], 4);
    prefix = 4 + uint64_t(dl);
    if (prefix > lens[0]) return false;
  }
  if ((lens[0] - prefix) % width != 0 || lens[1] % width != 0) return false;
  return lens[0] - prefix + lens[1] == rows * width;
}

// Host-only, thread-safe: footer + page walk, validation against the schema, device tables, planning facts.
static int prepare_sst(const hg_schema_desc* schema, uint64_t id, const uint8_t* data, uint64_t size, SstResident* r,
                       std::vector<PageDev>* pages_out, std::vector<ChunkDev>* chunks_out, std::string* errmsg) {
  auto fail = [&](int code, const std::string& msg) { *errmsg = msg; return code; };
  r->id = id;
  r->size = size;
  std::string err;
  if (!parse_parquet(data, size, &r->meta, &err)) return fail(HG_ERR_FORMAT, "sst " + std::to_string(id) + ": " + err);
  const FileMetaData& m = r->meta;
  if (uint32_t(m.ncols) != schema->num_columns)
    return fail(HG_ERR_INVALID, "sst has " + std::to_string(m.ncols) + " columns, schema has " + std::to_string(schema->num_columns));
  for (int c = 0; c < m.ncols; c++) {
    if (m.phys_types[c] != expected_phys(schema->types[c]))
      return fail(HG_ERR_INVALID, "column " + std::to_string(c) + ": parquet physical type does not match the schema");
    if (m.repetition[c] == 2) return fail(HG_ERR_UNSUPPORTED, "repeated columns");
  }
  std::vector<PageDev>& pages = *pages_out;
  pages.assign(m.pages.size(), PageDev());
  for (size_t i = 0; i < pages.size(); i++) {
    const PageMeta& pm = m.pages[i];
    if (pm.encoding != ENC_PLAIN && pm.encoding != ENC_DELTA_BINARY_PACKED && pm.encoding != ENC_DELTA_LENGTH_BYTE_ARRAY && pm.encoding != ENC_RLE_DICT &&
        pm.encoding != ENC_PLAIN_DICT)
      return fail(HG_ERR_UNSUPPORTED, "page encoding " + std::to_string(pm.encoding) +
                                      " (PLAIN, DELTA_BINARY_PACKED, DELTA_LENGTH_BYTE_ARRAY and RLE_DICTIONARY are implemented)");
    PageDev& pd = pages[i];
    pd.payload_off = pm.payload_off;
    pd.comp_size = pm.comp_size;
    pd.uncomp_size = pm.uncomp_size;
    pd.num_values = pm.num_values;
    pd.v2_def_len = pm.v2_def_len;
    pd.v2_rep_len = pm.v2_rep_len;
    pd.page_type = pm.page_type;
    pd.encoding = pm.encoding;
    pd.v2_compressed = pm.v2_compressed;
    pd._pad = 0;
  }
  std::vector<ChunkDev>& chunks = *chunks_out;
  chunks.assign(m.rgs.size() * size_t(m.ncols), ChunkDev());
  for (size_t g = 0; g < m.rgs.size(); g++)
    for (int c = 0; c < m.ncols; c++) {
      const ChunkMeta& cm = m.rgs[g].cols[c];
      if (cm.codec != CODEC_UNCOMPRESSED && cm.codec != CODEC_SNAPPY && cm.codec != CODEC_ZSTD)
        return fail(HG_ERR_UNSUPPORTED, "codec " + std::to_string(cm.codec) + " (UNCOMPRESSED, SNAPPY and ZSTD are implemented)");
      if (cm.scratch_bytes > 0xffffffffull) return fail(HG_ERR_UNSUPPORTED, "column chunk larger than 4 GiB");
      // the kernels decode by the CHUNK's physical type; the output buffers are laid out by the schema's
      if (cm.phys_type != m.phys_types[c])
        return fail(HG_ERR_FORMAT, "sst " + std::to_string(id) + ": column chunk type differs from the schema element's type");
      // every kernel indexes a chunk by the ROW GROUP's row count: the chunk must hold exactly that many values, and an
      // uncompressed page must really contain the bytes the decoders will read (compressed pages are bounded by their
      // scratch size on the device)
      if (cm.num_values != m.rgs[g].num_rows)
        return fail(HG_ERR_FORMAT, "sst " + std::to_string(id) + ": column chunk value count differs from the row group's row count");
      {
        const uint32_t pw = (cm.phys_type == PT_INT32 || cm.phys_type == PT_FLOAT) ? 4u : 8u;
        const bool optional = m.repetition[c] == 1;
        const bool no_nulls = cm.phys_type != PT_BYTE_ARRAY && (!optional || (cm.stats.has_null_count && cm.stats.null_count == 0));   // (byte arrays: variable width)
        for (uint32_t pi = cm.first_page; pi < cm.first_page + cm.num_pages; pi++) {
          const PageMeta& pm = m.pages[pi];
          if (pm.page_type == PAGE_DATA_V2) {
            if (uint64_t(pm.v2_def_len) + pm.v2_rep_len > pm.comp_size || uint64_t(pm.v2_def_len) + pm.v2_rep_len > pm.uncomp_size)
              return fail(HG_ERR_FORMAT, "sst " + std::to_string(id) + ": V2 level bytes exceed the page");
            if (pm.encoding == ENC_PLAIN && no_nulls && uint64_t(pm.v2_def_len) + pm.v2_rep_len + uint64_t(pm.num_values) * pw > pm.uncomp_size)
              return fail(HG_ERR_FORMAT, "sst " + std::to_string(id) + ": page smaller than its values");
          } else if (cm.codec == CODEC_UNCOMPRESSED) {
            if (pm.comp_size != pm.uncomp_size) return fail(HG_ERR_FORMAT, "sst " + std::to_string(id) + ": uncompressed page with differing sizes");
            uint64_t prefix = 0;
            if (optional) {
              if (pm.uncomp_size < 4) return fail(HG_ERR_FORMAT, "sst " + std::to_string(id) + ": page smaller than its level header");
              uint32_t dl;
              std::memcpy(&dl, data + pm.payload_off, 4);
              prefix = 4 + uint64_t(dl);
              if (prefix > pm.uncomp_size) return fail(HG_ERR_FORMAT, "sst " + std::to_string(id) + ": definition levels exceed the page");
            }
            if (pm.encoding == ENC_PLAIN && no_nulls && prefix + uint64_t(pm.num_values) * pw > pm.uncomp_size)
              return fail(HG_ERR_FORMAT, "sst " + std::to_string(id) + ": page smaller than its values");
          }
        }
      }
      ChunkDev& cd = chunks[g * m.ncols + c];
      cd.first_page = cm.first_page;
      cd.num_pages = cm.num_pages;
      cd.scratch_bytes = uint32_t(cm.scratch_bytes);
      cd.phys = uint8_t(cm.phys_type);
      cd.codec = uint8_t(cm.codec);
      cd.optional = uint8_t(m.repetition[c] == 1);
      cd.stored = 0;
      if (cm.phys_type == PT_BYTE_ARRAY)
        for (uint32_t pi = cm.first_page; pi < cm.first_page + cm.num_pages; pi++)
          if (m.pages[pi].encoding != ENC_PLAIN && m.pages[pi].encoding != ENC_DELTA_LENGTH_BYTE_ARRAY)
            return fail(HG_ERR_UNSUPPORTED, "byte-array column: PLAIN and DELTA_LENGTH_BYTE_ARRAY pages are implemented");
      cd.dict_payload_off = cm.has_dict_page ? cm.dict_payload_off : 0;
      cd.dict_comp = cm.has_dict_page ? cm.dict_comp_size : 0;
      cd.dict_uncomp = cm.has_dict_page ? cm.dict_uncomp_size : 0;
      for (uint32_t pi = cm.first_page; pi < cm.first_page + cm.num_pages; pi++) {
        if (m.pages[pi].encoding == ENC_DELTA_LENGTH_BYTE_ARRAY && cm.phys_type != PT_BYTE_ARRAY)
          return fail(HG_ERR_FORMAT, "DELTA_LENGTH_BYTE_ARRAY on a fixed-width column");
        if (m.pages[pi].encoding == ENC_DELTA_BINARY_PACKED && cm.phys_type != PT_INT32 && cm.phys_type != PT_INT64)
          return fail(HG_ERR_FORMAT, "DELTA_BINARY_PACKED on a non-integer column");
        if ((m.pages[pi].encoding == ENC_RLE_DICT || m.pages[pi].encoding == ENC_PLAIN_DICT) && !cm.has_dict_page)
          return fail(HG_ERR_FORMAT, "dictionary-encoded page without a dictionary page");
      }
      if (cm.phys_type != PT_BYTE_ARRAY && cm.codec == CODEC_SNAPPY && cm.num_pages == 1 && m.pages[cm.first_page].page_type == PAGE_DATA && m.rgs[g].num_rows > 0)
        cd.stored = classify_stored(data, size, m.pages[cm.first_page], cd.optional != 0,
                                    (cm.phys_type == PT_INT32 || cm.phys_type == PT_FLOAT) ? 4u : 8u, uint64_t(m.rgs[g].num_rows)) ? 1 : 0;
    }
  r->rgcol.resize(m.rgs.size() * size_t(m.ncols));
  r->rg_rows.resize(m.rgs.size());
  for (size_t g = 0; g < m.rgs.size(); g++) {
    r->rg_rows[g] = uint32_t(m.rgs[g].num_rows);
    for (int c = 0; c < m.ncols; c++) {
      const ChunkMeta& cm = m.rgs[g].cols[c];
      RgCol& rc = r->rgcol[g * m.ncols + c];
      const uint32_t t = schema->types[c];
      if (cm.stats.has_min && cm.stats.has_max && cm.phys_type != PT_BYTE_ARRAY) {
        rc.has_minmax = 1;
        rc.mn = widen_stat(cm.stats.min, cm.phys_type, t);
        rc.mx = widen_stat(cm.stats.max, cm.phys_type, t);
      }
      rc.null_all = cm.stats.has_null_count && cm.stats.null_count == m.rgs[g].num_rows;
      rc.null_none = cm.stats.has_null_count && cm.stats.null_count == 0;
      rc.snappy = cm.codec == CODEC_SNAPPY;
      rc.scratch = uint32_t(cm.scratch_bytes);
      const bool one_plain_v1 = cm.num_pages == 1 && m.pages[cm.first_page].page_type == PAGE_DATA && m.pages[cm.first_page].encoding == ENC_PLAIN &&
                                !cm.has_dict_page && cm.phys_type != PT_BYTE_ARRAY;
      rc.simple_page = cm.codec == CODEC_UNCOMPRESSED && one_plain_v1;
      rc.single_page = one_plain_v1;
      rc.stored = chunks[g * m.ncols + c].stored;
    }
  }
  {
    const uint32_t t0 = schema->types[0];
    for (int c = 0; c < m.ncols && c < MAX_COLS; c++) {
      r->col_all_simple[c] = true; r->col_null_none[c] = true; r->col_has_minmax[c] = true;
      r->col_all_single[c] = true; r->col_any_snappy[c] = false; r->col_snappy_all_stored[c] = true; r->col_snappy_any_stored[c] = false; r->col_any_zstd[c] = false;
    }
    bool first = true;
    r->pk0_range_ok = true;
    for (size_t g = 0; g < m.rgs.size(); g++) {
      const uint32_t rows = r->rg_rows[g];
      r->rows_total += rows;
      if (rows == 0) continue;
      const RgCol* rc = &r->rgcol[g * m.ncols];
      for (int c = 0; c < m.ncols && c < MAX_COLS; c++) {
        if (!rc[c].simple_page) r->col_all_simple[c] = false;
        if (!rc[c].single_page) r->col_all_single[c] = false;
        if (m.rgs[g].cols[c].codec == CODEC_ZSTD) { r->col_any_zstd[c] = true; r->any_zstd = true; }
        if (rc[c].snappy) { r->col_any_snappy[c] = true; if (!rc[c].stored) r->col_snappy_all_stored[c] = false; else r->col_snappy_any_stored[c] = true; }
        r->col_max_scratch[c] = std::max(r->col_max_scratch[c], rc[c].scratch);
        r->col_comp_bytes[c] += uint64_t(m.rgs[g].cols[c].total_compressed);
        if (!rc[c].null_none) r->col_null_none[c] = false;
        if (!rc[c].has_minmax) r->col_has_minmax[c] = false;
      }
      if (rc[0].has_minmax && rc[0].null_none) {
        if (first) { r->pk0_min = rc[0].mn; r->pk0_max = rc[0].mx; first = false; }
        else {
          if (cmp_host(rc[0].mn, r->pk0_min, t0) < 0) r->pk0_min = rc[0].mn;
          if (cmp_host(rc[0].mx, r->pk0_max, t0) > 0) r->pk0_max = rc[0].mx;
        }
        uint64_t span = rc[0].mx - rc[0].mn + 1;
        r->group_bound += std::min<uint64_t>(span == 0 ? rows : span, rows) + 1;
      } else { r->pk0_range_ok = false; r->group_bound += uint64_t(rows) + 1; }
    }
  }
  return HG_OK;
}

static int read_whole_file(const char* path, std::vector<uint8_t>* buf) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return set_error(HG_ERR_NOT_FOUND, std::string("cannot open ") + path);
  long n = -1;
  if (std::fseek(f, 0, SEEK_END) == 0) n = std::ftell(f);
  if (n < 0 || std::fseek(f, 0, SEEK_SET) != 0) { std::fclose(f); return set_error(HG_ERR_NOT_FOUND, std::string("cannot size ") + path); }
  buf->resize(size_t(n));
  size_t got = std::fread(buf->data(), 1, size_t(n), f);
  std::fclose(f);
  if (got != size_t(n)) return set_error(HG_ERR_NOT_FOUND, std::string("short read on ") + path);
  return HG_OK;
}

// hg_sst_load: the whole file becomes resident (cudaMalloc'd, cached until hg_sst_unload).
static int load_sst_locked(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* d) {
  if (e->ssts.count(d->id)) return HG_OK;
  std::vector<uint8_t> filebuf;
  const uint8_t* data = d->data;
  uint64_t size = d->size;
  if (!data) {
    if (!d->path) return set_error(HG_ERR_NOT_FOUND, "sst " + std::to_string(d->id) + " is not resident and no data/path given");
    int rc = read_whole_file(d->path, &filebuf);
    if (rc) return rc;
    data = filebuf.data();
    size = filebuf.size();
  }
  auto r = std::make_unique<SstResident>();
  std::vector<PageDev> pages;
  std::vector<ChunkDev> chunks;
  std::string err;
  int prc = prepare_sst(schema, d->id, data, size, r.get(), &pages, &chunks, &err);
  if (prc) return set_error(prc, err);
  uint64_t need = size + 64 + pages.size() * sizeof(PageDev) + chunks.size() * sizeof(ChunkDev) + r->rgcol.size() * sizeof(RgCol) +
                  r->rg_rows.size() * sizeof(uint32_t);
  if (e->budget && e->resident_bytes + need > e->budget)
    return set_error(HG_ERR_OOM, "HBM budget exceeded while loading sst " + std::to_string(d->id));
  CU_TRY(cudaMalloc(&r->d_bytes, size + 64));
  CU_TRY(cudaMalloc(&r->d_pages, std::max<size_t>(pages.size(), 1) * sizeof(PageDev)));
  CU_TRY(cudaMalloc(&r->d_chunks, std::max<size_t>(chunks.size(), 1) * sizeof(ChunkDev)));
  CU_TRY(cudaMalloc(&r->d_rgcol, std::max<size_t>(r->rgcol.size(), 1) * sizeof(RgCol)));
  CU_TRY(cudaMalloc(&r->d_rg_rows, std::max<size_t>(r->rg_rows.size(), 1) * sizeof(uint32_t)));
  if (!r->rgcol.empty()) CU_TRY(cudaMemcpyAsync(r->d_rgcol, r->rgcol.data(), r->rgcol.size() * sizeof(RgCol), cudaMemcpyHostToDevice, e->stream));
  if (!r->rg_rows.empty()) CU_TRY(cudaMemcpyAsync(r->d_rg_rows, r->rg_rows.data(), r->rg_rows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
  CU_TRY(cudaMemcpyAsync(r->d_bytes, data, size, cudaMemcpyHostToDevice, e->stream));
  CU_TRY(cudaMemsetAsync(r->d_bytes + size, 0, 64, e->stream));
  if (!pages.empty()) CU_TRY(cudaMemcpyAsync(r->d_pages, pages.data(), pages.size() * sizeof(PageDev), cudaMemcpyHostToDevice, e->stream));
  if (!chunks.empty()) CU_TRY(cudaMemcpyAsync(r->d_chunks, chunks.data(), chunks.size() * sizeof(ChunkDev), cudaMemcpyHostToDevice, e->stream));
  CU_TRY(cudaStreamSynchronize(e->stream));
  r->device_bytes = need;
  e->resident_bytes += need;
  e->stats.bytes_h2d += need;
  e->ssts[d->id] = std::move(r);
  return HG_OK;
}

// ------------------------------------------------------------------------------------------------------ scan planning
static bool rg_may_match(const RgCol* rc, uint32_t num_rows, const hg_schema_desc* schema, const hg_predicate* preds, const uint64_t* lits, size_t np) {
  // DataFusion PruningPredicate (pinned by the plan text at read.rs:613):
  //   CASE WHEN null_count = row_count THEN false ELSE <min/max rewrite of the comparison> END
  for (size_t i = 0; i < np; i++) {
    const RgCol& c = rc[preds[i].column];
    const uint32_t t = schema->types[preds[i].column];
    if (c.null_all) return false;
    if (!c.has_minmax) continue;
    const uint64_t mn = c.mn, mx = c.mx, lit = lits[i];
    bool ok = true;
    switch (preds[i].op) {
      case HG_OP_EQ: ok = cmp_host(mn, lit, t) <= 0 && cmp_host(lit, mx, t) <= 0; break;
      case HG_OP_NE: ok = cmp_host(mn, lit, t) != 0 || cmp_host(lit, mx, t) != 0; break;
      case HG_OP_LT: ok = cmp_host(mn, lit, t) < 0; break;
      case HG_OP_LE: ok = cmp_host(mn, lit, t) <= 0; break;
      case HG_OP_GT: ok = cmp_host(mx, lit, t) > 0; break;
      case HG_OP_GE: ok = cmp_host(mx, lit, t) >= 0; break;
      case HG_OP_IN:    // PruningPredicate expands a short IN list into `c = v1 OR c = v2 ..`
        ok = false;
        for (uint32_t j = 0; j < preds[i].in_count && !ok; j++) ok = cmp_host(mn, preds[i].in_values[j], t) <= 0 && cmp_host(preds[i].in_values[j], mx, t) <= 0;
        break;
    }
    if (!ok) return false;
  }
  return true;
}

// Small host -> device uploads go through one pinned staging buffer.  The cursor is per CALL (reset in begin_call, when the
// stream is idle): copies are asynchronous, so a region must not be reused before the stream has consumed it.
int stage_upload(hg_engine* e, void* dst, const void* src, size_t bytes, size_t* stage_off) {
  size_t off = (e->stage_cursor + 255) & ~size_t(255);
  if (off + bytes > e->h_stage_bytes) {
    // grow (rare): everything staged so far in this call must reach the device first
    CU_TRY(cudaStreamSynchronize(e->stream));
    size_t nb = std::max<size_t>((off + bytes) * 2, 1 << 20);
    void* p = nullptr;
    CU_TRY(cudaMallocHost(&p, nb));
    if (e->h_stage) cudaFreeHost(e->h_stage);
    e->h_stage = p;
    e->h_stage_bytes = nb;
    off = 0;
  }
  std::memcpy(static_cast<char*>(e->h_stage) + off, src, bytes);
  CU_TRY(cudaMemcpyAsync(dst, static_cast<char*>(e->h_stage) + off, bytes, cudaMemcpyHostToDevice, e->stream));
  e->stage_cursor = off + bytes;
  if (stage_off) *stage_off = e->stage_cursor;
  return HG_OK;
}


// ------------------------------------------------------------------------------------- transient, selective SST loads
// A scan called with host bytes for SSTs that are not resident does not cache them: it copies ONLY the byte ranges the
// query can touch — column chunks of the needed columns in row groups that survive statistics pruning — into arena
// memory laid out at the file's own offsets (so the page table stays valid), uploads the page tables, and forgets
// everything at the end of the call.  Footers are parsed on a small thread pool.  With pinned host buffers the ranges
// are fetched by one gather kernel reading host memory over PCIe (no per-range API call).
struct CopyRange { const uint8_t* src; uint8_t* dst; uint64_t bytes; };

__global__ void __launch_bounds__(256) gather_ranges_kernel(const CopyRange* __restrict__ ranges, uint32_t nranges) {
  for (uint32_t r = blockIdx.x; r < nranges; r += gridDim.x) {
    const CopyRange cr = ranges[r];
    const uintptr_t sa = reinterpret_cast<uintptr_t>(cr.src), da = reinterpret_cast<uintptr_t>(cr.dst);
    if (((sa ^ da) & 15) == 0) {
      uint64_t head = (16 - (sa & 15)) & 15;
      if (head > cr.bytes) head = cr.bytes;
      for (uint64_t i = threadIdx.x; i < head; i += 256) cr.dst[i] = cr.src[i];
      const uint64_t nvec = (cr.bytes - head) >> 4;
      const uint4* s = reinterpret_cast<const uint4*>(cr.src + head);
      uint4* d = reinterpret_cast<uint4*>(cr.dst + head);
      for (uint64_t i = threadIdx.x; i < nvec; i += 256) d[i] = s[i];
      for (uint64_t i = head + (nvec << 4) + threadIdx.x; i < cr.bytes; i += 256) cr.dst[i] = cr.src[i];
    } else {
      for (uint64_t i = threadIdx.x; i < cr.bytes; i += 256) cr.dst[i] = cr.src[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------- host worker pool
// A transient load parses footers / page headers and builds tens of thousands of byte ranges per call on the host while PCIe waits:
// per-file work, run on a small persistent pool (threads created per call cost as much as the work they would do).
namespace {
class WorkPool {
 public:
  explicit WorkPool(unsigned n) { for (unsigned i = 0; i < n; i++) th_.emplace_back([this] { loop(); }); }
  ~WorkPool() {
    { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  // fn(i) for every i in [0, n), on the workers and on the caller; returns when all have finished
  void parallel_for(size_t n, const std::function<void(size_t)>& fn) {
    if (n == 0) return;
    if (n == 1 || th_.empty()) { for (size_t i = 0; i < n; i++) fn(i); return; }
    std::lock_guard<std::mutex> serial(call_mu_);           // one parallel_for at a time
    {
      std::lock_guard<std::mutex> g(mu_);
      fn_ = &fn; n_ = n; next_.store(0); done_ = 0; gen_++;
    }
    cv_.notify_all();
    run();
    std::unique_lock<std::mutex> g(mu_);
    cv_done_.wait(g, [&] { return done_ == n_; });
    fn_ = nullptr;
  }

 private:
  void run() {
    size_t mine = 0;
    for (;;) {
      const size_t i = next_.fetch_add(1);
      if (i >= n_) break;
      (*fn_)(i);
      mine++;
    }
    if (mine) {
      std::lock_guard<std::mutex> g(mu_);
      done_ += mine;
      if (done_ == n_) cv_done_.notify_all();
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
      }
      run();
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_, call_mu_;
  std::condition_variable cv_, cv_done_;
  const std::function<void(size_t)>* fn_ = nullptr;
  size_t n_ = 0, done_ = 0;
  std::atomic<size_t> next_{0};
  uint64_t gen_ = 0;
  bool stop_ = false;
};
WorkPool& work_pool() {
  static WorkPool* p = new WorkPool(std::min(15u, std::max(1u, std::thread::hardware_concurrency() / 2)));   // + the caller = 16
  return *p;
}
}  // namespace

static int load_transient(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, const std::vector<size_t>& pending,
                          const hg_predicate* preds, size_t np, std::vector<uint32_t> need_cols, bool seq_if_overlap,
                          const std::vector<size_t>& resident_idx) {
  const size_t k = pending.size();
  static const bool trace = getenv("HORAE_TRACE") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
  const auto tt0 = now();
  std::vector<std::unique_ptr<SstResident>> rs(k);
  std::vector<std::vector<PageDev>> pages(k);
  std::vector<std::vector<ChunkDev>> chunks(k);
  std::vector<std::vector<uint8_t>> filebufs(k);
  std::vector<const uint8_t*> datas(k);
  std::vector<uint64_t> sizes(k);
  for (size_t j = 0; j < k; j++) {
    const hg_sst_desc& d = ssts[pending[j]];
    datas[j] = d.data;
    sizes[j] = d.size;
    if (!d.data) {
      if (!d.path) return set_error(HG_ERR_NOT_FOUND, "sst " + std::to_string(d.id) + " is not resident and no data/path given");
      int rc = read_whole_file(d.path, &filebufs[j]);
      if (rc) return rc;
      datas[j] = filebufs[j].data();
      sizes[j] = filebufs[j].size();
    }
    rs[j] = std::make_unique<SstResident>();
    rs[j]->owned = false;
  }
  // ---- parse in parallel
  std::vector<int> codes(k, 0);
  std::vector<std::string> errs(k);
  work_pool().parallel_for(k, [&](size_t j) {
    codes[j] = prepare_sst(schema, ssts[pending[j]].id, datas[j], sizes[j], rs[j].get(), &pages[j], &chunks[j], &errs[j]);
  });
  for (size_t j = 0; j < k; j++) if (codes[j]) return set_error(codes[j], errs[j]);
  const auto tt1 = now();
  // ---- __seq__ is only needed when the inputs are not provably PK-disjoint (a real merge will run)
  if (seq_if_overlap) {
    std::vector<const SstResident*> all;
    for (auto& r : rs) if (r->rows_total) all.push_back(r.get());
    for (size_t i : resident_idx) { auto it = e->ssts.find(ssts[i].id); if (it != e->ssts.end() && it->second->rows_total) all.push_back(it->second.get()); }
    bool disjoint = true;
    const uint32_t t0 = schema->types[0];
    if (all.size() > 1) {
      for (auto* f : all) if (!f->pk0_range_ok) disjoint = false;
      if (disjoint) {
        std::stable_sort(all.begin(), all.end(), [&](const SstResident* a, const SstResident* b) { return cmp_host(a->pk0_min, b->pk0_min, t0) < 0; });
        for (size_t j = 0; j + 1 < all.size() && disjoint; j++) disjoint = cmp_host(all[j]->pk0_max, all[j + 1]->pk0_min, t0) < 0;
      }
    }
    if (!disjoint) need_cols.push_back(schema->num_columns - 2);
  }
  std::sort(need_cols.begin(), need_cols.end());
  need_cols.erase(std::unique(need_cols.begin(), need_cols.end()), need_cols.end());
  // ---- arena allocations + byte ranges
  const bool prune = !(e->flags & HG_FLAG_NO_PRUNING);
  uint64_t lits[MAX_PREDS];
  for (size_t i = 0; i < np; i++) lits[i] = pred_literal(preds[i], schema->types[preds[i].column]);
  bool all_pinned = k > 0;
  for (size_t j = 0; j < k && all_pinned; j++) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, datas[j]) != cudaSuccess || at.type != cudaMemoryTypeHost) { all_pinned = false; cudaGetLastError(); }
  }
  size_t stage_off = 0;
  uint64_t copied = 0;
  size_t n_ranges = 0;
  // moves a batch of byte ranges host -> device on the engine's stream
  auto move_ranges = [&](std::vector<CopyRange>& ranges) -> int {
    n_ranges += ranges.size();
    for (auto& cr : ranges) copied += cr.bytes;
    if (ranges.empty()) return HG_OK;
    if (all_pinned) {
      CopyRange* d_ranges = static_cast<CopyRange*>(g_arena->alloc(ranges.size() * sizeof(CopyRange)));
      if (!d_ranges) return set_error(HG_ERR_OOM, "out of device memory");
      int rc = stage_upload(e, d_ranges, ranges.data(), ranges.size() * sizeof(CopyRange), &stage_off);
      if (rc) return rc;
      gather_ranges_kernel<<<int(std::min<size_t>(ranges.size(), 148 * 8)), 256, 0, e->stream>>>(d_ranges, uint32_t(ranges.size()));
      e->launches++;
    } else {
      for (auto& cr : ranges) CU_TRY(cudaMemcpyAsync(cr.dst, cr.src, cr.bytes, cudaMemcpyHostToDevice, e->stream));
    }
    return HG_OK;
  };
  auto add_range = [&](std::vector<CopyRange>& ranges, size_t j, uint32_t g, uint32_t c) {
    SstResident& r = *rs[j];
    const ChunkMeta& cm = r.meta.rgs[g].cols[c];
    uint64_t lo = uint64_t(cm.data_page_offset);
    if (cm.dict_page_offset > 0 && uint64_t(cm.dict_page_offset) < lo) lo = uint64_t(cm.dict_page_offset);   // the chunk starts at its dictionary page
    uint64_t hi = lo + uint64_t(cm.total_compressed);
    if (hi > r.size) hi = r.size;
    hi = std::min<uint64_t>(r.size, hi + 16);               // the unaligned 8-byte loads may touch one word past the values
    if (!ranges.empty() && ranges.back().src + ranges.back().bytes >= datas[j] + lo && ranges.back().src <= datas[j] + lo &&
        ranges.back().dst == r.d_bytes + (ranges.back().src - datas[j])) {
      uint64_t end = std::max<uint64_t>(uint64_t(ranges.back().src - datas[j]) + ranges.back().bytes, hi);
      ranges.back().bytes = end - uint64_t(ranges.back().src - datas[j]);
    } else ranges.push_back(CopyRange{datas[j] + lo, r.d_bytes + lo, hi - lo});
  };
  // row groups that survive statistics pruning, in file order
  struct KeptRg { uint32_t j, g; };
  std::vector<KeptRg> kept;
  for (size_t j = 0; j < k; j++) {
    SstResident& r = *rs[j];
    r.d_bytes = static_cast<uint8_t*>(g_arena->alloc(r.size + 64));
    if (!r.d_bytes) return set_error(HG_ERR_OOM, "out of device memory for transient SST");
    const size_t ncols = size_t(r.meta.ncols);
    for (size_t g = 0; g < r.meta.rgs.size(); g++) {
      if (r.rg_rows[g] == 0) continue;
      if (prune && np && !rg_may_match(&r.rgcol[g * ncols], r.rg_rows[g], schema, preds, lits, np)) continue;
      kept.push_back(KeptRg{uint32_t(j), uint32_t(g)});
    }
  }
  // ---- late materialisation across PCIe: when one predicate column is a single PLAIN page without NULLs in every file (Snappy or
  //      not), move ITS chunks first, let the device find — per row group — the first and the last row that pass its predicates,
  //      and move the other columns only for row groups that have one; of the non-key columns that can be addressed by row
  //      (uncompressed PLAIN pages, stored Snappy pages) only the rows between the first and the last passing row.
  //      The filter precedes merge and dedup (read.rs:459-480): rows that fail the gate take part in nothing downstream.
  int gate_col = -1;
  if (prune && np && !(e->flags & HG_FLAG_NO_LATE_MATERIALIZATION) && need_cols.size() > 1 && !kept.empty()) {
    uint64_t best_bytes = ~0ull;
    for (size_t i = 0; i < np; i++) {
      const uint32_t c = preds[i].column;
      bool ok = c < uint32_t(MAX_COLS);
      for (size_t i2 = 0; i2 < np; i2++) if (preds[i2].column == c && preds[i2].op == HG_OP_IN) ok = false;   // the gate kernel tests intervals
      uint64_t bytes = 0;
      for (size_t j = 0; j < k && ok; j++) {
        ok = rs[j]->rows_total == 0 || (rs[j]->col_all_single[c] && rs[j]->col_null_none[c] && !rs[j]->col_any_zstd[c]);
        bytes += rs[j]->col_comp_bytes[c];
      }
      if (ok && bytes < best_bytes) { best_bytes = bytes; gate_col = int(c); }
    }
    uint64_t need_bytes = 0;
    for (uint32_t c : need_cols) for (size_t j = 0; j < k; j++) need_bytes += rs[j]->col_comp_bytes[c];
    if (gate_col >= 0 && best_bytes * 2 > need_bytes) gate_col = -1;          // the gate would be most of the bytes anyway
  }
  std::vector<fused::GateOut> gate_out;
  if (gate_col >= 0) {
    std::vector<fused::GateRg> descs(kept.size());
    std::vector<k::RawPage> raw;
    const uint32_t gw = type_width_host(schema->types[gate_col]) <= 4 ? 4u : 8u;
    // per file on the worker pool: the gate column's byte ranges, the pages to decompress and the gate descriptors.  Scratch addresses
    // are offsets into the file's share until the (serial) arena allocation below
    std::vector<std::vector<CopyRange>> franges(k);
    std::vector<std::vector<k::RawPage>> fraw(k);
    std::vector<uint64_t> fneed(k, 0);
    std::vector<int> ferr(k, 0);
    std::vector<std::pair<size_t, size_t>> gseg(k, {0, 0});             // kept[] is ordered by file: [begin, end) per file
    for (size_t i = 0; i < kept.size(); i++) {
      if (gseg[kept[i].j].second == 0) gseg[kept[i].j].first = i;
      gseg[kept[i].j].second = i + 1;
    }
    work_pool().parallel_for(k, [&](size_t j) {
      SstResident& r = *rs[j];
      for (size_t i = gseg[j].first; i < gseg[j].second; i++) {
        const uint32_t g = kept[i].g;
        add_range(franges[j], j, g, uint32_t(gate_col));
        const ChunkDev& cd = chunks[j][size_t(g) * size_t(r.meta.ncols) + size_t(gate_col)];
        const PageDev& pg = pages[j][cd.first_page];
        uint64_t off = pg.payload_off;
        if (cd.codec == CODEC_SNAPPY) {
          // decompressed on the device into arena scratch; the level prefix is skipped there
          uint8_t* rel = reinterpret_cast<uint8_t*>(uintptr_t(fneed[j]));
          fneed[j] += (page_scratch_bytes(pg.uncomp_size) + 16 + 255) & ~uint64_t(255);
          fraw[j].push_back(k::RawPage{r.d_bytes + off, rel, pg.comp_size, pg.uncomp_size});
          if (uint64_t(pg.uncomp_size) < uint64_t(r.rg_rows[g]) * gw) { ferr[j] = 1; return; }
          descs[i] = fused::GateRg{rel, r.rg_rows[g], (cd.optional ? 1u : 0u) | 0x80000000u};      // bit 31: vals is still an offset
        } else {
          if (cd.optional) {                                  // [u32 len][RLE def levels] in front of the values (all valid here)
            uint32_t len = 0;
            if (off + 4 > r.size) { ferr[j] = 2; return; }
            std::memcpy(&len, datas[j] + off, 4);
            off += 4 + uint64_t(len);
          }
          if (off + uint64_t(r.rg_rows[g]) * gw > r.size) { ferr[j] = 3; return; }
          descs[i] = fused::GateRg{r.d_bytes + off, r.rg_rows[g], 0};
        }
      }
    });
    for (size_t j = 0; j < k; j++) {
      if (ferr[j] == 1) return set_error(HG_ERR_FORMAT, "column chunk smaller than its values");
      if (ferr[j] == 2) return set_error(HG_ERR_FORMAT, "page payload out of bounds");
      if (ferr[j] == 3) return set_error(HG_ERR_FORMAT, "column chunk out of bounds");
      uint8_t* base = nullptr;
      if (fneed[j]) {
        base = static_cast<uint8_t*>(g_arena->alloc(fneed[j]));
        if (!base) return set_error(HG_ERR_OOM, "out of device memory");
      }
      for (auto& rp : fraw[j]) { rp.dst = base + uintptr_t(rp.dst); raw.push_back(rp); }
      for (size_t i = gseg[j].first; i < gseg[j].second; i++)
        if (descs[i].prefixed & 0x80000000u) { descs[i].vals = base + uintptr_t(descs[i].vals); descs[i].prefixed &= 1u; }
      int rc = move_ranges(franges[j]);
      if (rc) return rc;
    }
    int rc = 0;
    fused::GateRg* d_descs = static_cast<fused::GateRg*>(g_arena->alloc(descs.size() * sizeof(fused::GateRg)));
    fused::GateOut* d_out = static_cast<fused::GateOut*>(g_arena->alloc(kept.size() * sizeof(fused::GateOut) + 16));
    uint32_t* d_tick = static_cast<uint32_t*>(g_arena->alloc(64));
    if (!d_descs || !d_out || !d_tick) return set_error(HG_ERR_OOM, "out of device memory");
    CU_TRY(cudaMemsetAsync(d_tick, 0, 64, e->stream));
    if (!raw.empty()) {
      k::RawPage* d_raw = static_cast<k::RawPage*>(g_arena->alloc(raw.size() * sizeof(k::RawPage)));
      if (!d_raw) return set_error(HG_ERR_OOM, "out of device memory");
      rc = stage_upload(e, d_raw, raw.data(), raw.size() * sizeof(k::RawPage), &stage_off);
      if (rc) return rc;
      k::snappy_raw_pages(e->L(), d_raw, uint32_t(raw.size()), d_tick, reinterpret_cast<int*>(d_tick + 1));
    }
    rc = stage_upload(e, d_descs, descs.data(), descs.size() * sizeof(fused::GateRg), &stage_off);
    if (rc) return rc;
    hg_predicate gp[MAX_PREDS];
    size_t ngp = 0;
    for (size_t i = 0; i < np; i++) if (int(preds[i].column) == gate_col) gp[ngp++] = preds[i];
    rc = fused::gate_row_groups(e, d_descs, uint32_t(kept.size()), schema->types[gate_col], gp, ngp, d_out);
    if (rc) return rc;
    gate_out.resize(kept.size());
    int herr = 0;
    CU_TRY(cudaMemcpyAsync(gate_out.data(), d_out, kept.size() * sizeof(fused::GateOut), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaMemcpyAsync(&herr, d_tick + 1, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    if (herr) return set_error(HG_ERR_FORMAT, "device decode error code " + std::to_string(herr) + " (gate column)");
    e->stats.bytes_d2h += kept.size() * sizeof(fused::GateOut);
    e->stage_cursor = 0;                                    // the stream is idle: the staging buffer can be reused
    for (size_t j = 0; j < k; j++) rs[j]->rg_dead.assign(rs[j]->rg_rows.size(), 0);
    std::vector<KeptRg> alive;
    std::vector<fused::GateOut> alive_out;
    for (size_t i = 0; i < kept.size(); i++) {
      if (gate_out[i].first <= gate_out[i].last) { alive.push_back(kept[i]); alive_out.push_back(gate_out[i]); }
      else rs[kept[i].j]->rg_dead[kept[i].g] = 1;
    }
    kept.swap(alive);
    gate_out.swap(alive_out);
  }
  const auto tt1b = now();
  // ---- the remaining columns of the row groups still in play: one task per file on the worker pool (a file's ranges only touch
  //      that file's tables), the files' range lists leave in file order
  {
    std::vector<std::vector<CopyRange>> file_ranges(k);
    std::vector<std::pair<size_t, size_t>> seg(k, {0, 0});              // kept[] is ordered by file: [begin, end) per file
    for (size_t i = 0; i < kept.size(); i++) {
      if (seg[kept[i].j].second == 0) seg[kept[i].j].first = i;
      seg[kept[i].j].second = i + 1;
    }
    std::vector<uint8_t> file_trunc(k, 0);
    work_pool().parallel_for(k, [&](size_t fj) {
      std::vector<CopyRange>& ranges = file_ranges[fj];
      auto add_bytes = [&](size_t j, uint64_t lo, uint64_t hi) {           // file byte range [lo, hi) (+ slack for the unaligned loads)
        SstResident& r = *rs[j];
        hi = std::min<uint64_t>(r.size, hi + 16);
        if (lo >= hi) return;
        if (!ranges.empty() && ranges.back().src + ranges.back().bytes >= datas[j] + lo && ranges.back().src <= datas[j] + lo &&
            ranges.back().dst == r.d_bytes + (ranges.back().src - datas[j])) {
          const uint64_t b0 = uint64_t(ranges.back().src - datas[j]);
          ranges.back().bytes = std::max<uint64_t>(b0 + ranges.back().bytes, hi) - b0;
        } else ranges.push_back(CopyRange{datas[j] + lo, r.d_bytes + lo, hi - lo});
      };
      for (size_t i = seg[fj].first; i < seg[fj].second; i++) {
      const KeptRg& kr = kept[i];
      SstResident& r = *rs[kr.j];
      for (uint32_t c : need_cols) {
        if (int(c) == gate_col) continue;
        const ChunkDev& cd = chunks[kr.j][size_t(kr.g) * size_t(r.meta.ncols) + c];
        const RgCol& rc = r.rgcol[size_t(kr.g) * size_t(r.meta.ncols) + c];
        const bool by_row = !gate_out.empty() && c >= schema->num_primary_keys && rc.single_page && rc.null_none &&
                            (cd.codec == CODEC_UNCOMPRESSED || cd.stored);
        if (!by_row) {
          // A Snappy page the device will decode only up to the last gate-passing row (fused scan, partial decode) travels as a
          // PREFIX of its compressed stream: the share of the stream that the needed share of the output takes, plus a margin.  The
          // page table tells the decoder where the prefix ends; a stream that turns out lopsided ends early, the decoder reports
          // it, and the entry point repeats the call without prefixes (e->trunc_used) — never a wrong result.
          if (!gate_out.empty() && gate_col == e->trunc_gate && c < 32 && ((e->trunc_mask >> c) & 1u) && cd.codec == CODEC_SNAPPY && rc.single_page && !cd.stored) {
            const PageDev& pg = pages[kr.j][cd.first_page];
            const uint32_t w = (cd.phys == PT_INT32 || cd.phys == PT_FLOAT) ? 4u : 8u;
            const uint64_t rows = r.rg_rows[kr.g];
            const uint64_t out_row = std::min<uint64_t>(uint64_t(gate_out[i].last) + 2, rows);            // gate_sel_kernel's RgSel::out_row
            const uint64_t need_uncomp = 16 + (rows + 7) / 8 + 8 + out_row * w + 2304;                    // stop_at + one batch of overshoot
            const uint64_t est = uint64_t(double(pg.comp_size) * double(need_uncomp) / double(std::max<uint32_t>(pg.uncomp_size, 1)) * 1.08) + 1024;
            if (est + 4096 < pg.comp_size) {
              const ChunkMeta& cm = r.meta.rgs[kr.g].cols[c];
              add_bytes(kr.j, uint64_t(cm.data_page_offset), pg.payload_off + est);
              pages[kr.j][cd.first_page].comp_size = uint32_t(est);
              file_trunc[fj] = 1;
              continue;
            }
          }
          add_range(ranges, kr.j, kr.g, c);
          continue;
        }
        // a page whose values can be addressed by row: only the blocks of rows that hold a passing row (GateOut::mask), cut to
        // [first, last]; adjacent blocks travel as one interval
        const PageDev& pg = pages[kr.j][cd.first_page];
        const uint32_t w = (cd.phys == PT_INT32 || cd.phys == PT_FLOAT) ? 4u : 8u;
        const uint8_t* base = datas[kr.j];
        const uint64_t body = pg.payload_off;
        // layout: PLAIN page = [prefix][values]; stored page (classify_stored) = [varint][literal 0 = prefix + n0 values][literal 1 = the rest]
        uint64_t v0 = body, v1 = 0, n0 = ~0ull, prefix = 0;                  // v0 / v1: file offsets of value 0 and of value n0
        if (cd.codec == CODEC_UNCOMPRESSED) {
          if (cd.optional) { uint32_t dl; std::memcpy(&dl, base + body, 4); prefix = 4 + uint64_t(dl); }
          add_bytes(kr.j, body, body + prefix);
          v0 = body + prefix;
        } else {
          uint64_t p = body;
          while (base[p] & 0x80) p++;
          p++;
          auto lit = [&](uint64_t at, uint64_t* len) -> uint64_t {
            uint64_t l = base[at] >> 2, hdr = 1;
            if (l >= 60) { const uint64_t nb = l - 59; l = 0; for (uint64_t q = 0; q < nb; q++) l |= uint64_t(base[at + 1 + q]) << (8 * q); hdr = 1 + nb; }
            *len = l + 1;
            return hdr;
          };
          uint64_t len0 = 0, len1 = 0;
          const uint64_t lit0 = p + lit(p, &len0);
          if (cd.optional) { uint32_t dl; std::memcpy(&dl, base + lit0, 4); prefix = 4 + uint64_t(dl); }
          n0 = (len0 - prefix) / w;
          add_bytes(kr.j, body, lit0 + prefix);
          v0 = lit0 + prefix;
          if (lit0 + len0 < body + pg.comp_size) {
            v1 = lit0 + len0 + lit(lit0 + len0, &len1);
            add_bytes(kr.j, lit0 + len0, v1);
          }
        }
        const uint32_t brows = fused::gate_block_rows(r.rg_rows[kr.g]);
        const uint32_t mask = gate_out[i].mask;
        for (uint32_t b = 0; b < 32u;) {
          if (!((mask >> b) & 1u)) { b++; continue; }
          uint32_t e2 = b;
          while (e2 + 1 < 32u && ((mask >> (e2 + 1)) & 1u)) e2++;
          const uint64_t first = std::max<uint64_t>(gate_out[i].first, uint64_t(b) * brows);
          const uint64_t last = std::min<uint64_t>(gate_out[i].last, uint64_t(e2 + 1) * brows - 1);
          b = e2 + 1;
          if (first > last) continue;
          if (first < n0) add_bytes(kr.j, v0 + first * w, v0 + (std::min<uint64_t>(last, n0 - 1) + 1) * w);
          if (v1 && last >= n0) add_bytes(kr.j, v1 + (std::max<uint64_t>(first, n0) - n0) * w, v1 + (last - n0 + 1) * w);
        }
      }
      }
    });
    for (size_t j = 0; j < k; j++) {
      if (file_trunc[j]) e->trunc_used = true;
      int rc = move_ranges(file_ranges[j]);
      if (rc) return rc;
    }
  }
  const auto tt1c = now();
  // ---- planning tables (device copies: dead row groups have zero rows => pruned by every device-side planner)
  for (size_t j = 0; j < k; j++) {
    SstResident& r = *rs[j];
    r.d_pages = static_cast<PageDev*>(g_arena->alloc(std::max<size_t>(pages[j].size(), 1) * sizeof(PageDev)));
    r.d_chunks = static_cast<ChunkDev*>(g_arena->alloc(std::max<size_t>(chunks[j].size(), 1) * sizeof(ChunkDev)));
    r.d_rgcol = static_cast<RgCol*>(g_arena->alloc(std::max<size_t>(r.rgcol.size(), 1) * sizeof(RgCol)));
    r.d_rg_rows = static_cast<uint32_t*>(g_arena->alloc(std::max<size_t>(r.rg_rows.size(), 1) * sizeof(uint32_t)));
    if (!r.d_pages || !r.d_chunks || !r.d_rgcol || !r.d_rg_rows) return set_error(HG_ERR_OOM, "out of device memory for transient SST");
    int rc = 0;
    if (!pages[j].empty()) rc = stage_upload(e, r.d_pages, pages[j].data(), pages[j].size() * sizeof(PageDev), &stage_off);
    if (!rc && !chunks[j].empty()) rc = stage_upload(e, r.d_chunks, chunks[j].data(), chunks[j].size() * sizeof(ChunkDev), &stage_off);
    if (!rc && !r.rgcol.empty()) rc = stage_upload(e, r.d_rgcol, r.rgcol.data(), r.rgcol.size() * sizeof(RgCol), &stage_off);
    if (!rc && !r.rg_rows.empty()) {
      if (r.rg_dead.empty()) rc = stage_upload(e, r.d_rg_rows, r.rg_rows.data(), r.rg_rows.size() * sizeof(uint32_t), &stage_off);
      else {
        std::vector<uint32_t> live(r.rg_rows);
        for (size_t g = 0; g < live.size(); g++) if (r.rg_dead[g]) live[g] = 0;
        rc = stage_upload(e, r.d_rg_rows, live.data(), live.size() * sizeof(uint32_t), &stage_off);
      }
    }
    if (rc) return rc;
    copied += pages[j].size() * sizeof(PageDev) + chunks[j].size() * sizeof(ChunkDev) + r.rgcol.size() * sizeof(RgCol);
  }
  e->stats.bytes_h2d += copied;
  if (trace) {
    const auto tt2 = now();
    cudaStreamSynchronize(e->stream);
    fprintf(stderr, "[transient] %zu files: parse %.0f us, gate phase %.0f us, main ranges %.0f us, tables %.0f us (%zu ranges, %.1f MB, gate column %d), copy wait %.0f us\n", k,
            us(tt0, tt1), us(tt1, tt1b), us(tt1b, tt1c), us(tt1c, tt2), n_ranges, copied / 1e6, gate_col, us(tt2, now()));
  }
  for (size_t j = 0; j < k; j++) {
    e->transient_ids.push_back(rs[j]->id);
    e->ssts[rs[j]->id] = std::move(rs[j]);
  }
  // host buffers read from files must outlive the async copies
  if (!all_pinned || std::any_of(filebufs.begin(), filebufs.end(), [](const std::vector<uint8_t>& b) { return !b.empty(); }))
    CU_TRY(cudaStreamSynchronize(e->stream));
  return HG_OK;
}

int build_plan(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n, const hg_predicate* preds,
                      size_t np, const std::vector<uint32_t>& need_cols, ScanPlan* plan) {
  const bool prune = !(e->flags & HG_FLAG_NO_PRUNING);
  struct FileSel { SstResident* f; std::vector<uint32_t> rgs; bool has_range = false; uint64_t mn = 0, mx = 0; size_t given_idx; };
  std::vector<FileSel> fs(n);
  const uint32_t t0 = schema->types[0];
  uint64_t lits[MAX_PREDS];
  for (size_t i = 0; i < np; i++) lits[i] = pred_literal(preds[i], schema->types[preds[i].column]);
  bool ranges_ok = true;
  size_t total_rgs = 0;
  for (size_t i = 0; i < n; i++) {
    auto it = e->ssts.find(ssts[i].id);
    if (it == e->ssts.end()) return set_error(HG_ERR_INTERNAL, "sst not resident after load");
    fs[i].f = it->second.get();
    fs[i].given_idx = i;
    const SstResident& f = *fs[i].f;
    const size_t ncols = size_t(f.meta.ncols), nrg = f.rg_rows.size();
    fs[i].rgs.reserve(nrg);
    for (size_t g = 0; g < nrg; g++) {
      const uint32_t rows = f.rg_rows[g];
      plan->rows_in_files += rows;
      if (rows == 0) continue;
      if (!f.rg_dead.empty() && f.rg_dead[g]) continue;          // transient load: no row of this row group passes the predicate
      const RgCol* rc = &f.rgcol[g * ncols];
      if (prune && np && !rg_may_match(rc, rows, schema, preds, lits, np)) continue;
      fs[i].rgs.push_back(uint32_t(g));
      const RgCol& c0 = rc[0];
      if (c0.has_minmax && c0.null_none) {
        if (!fs[i].has_range) { fs[i].mn = c0.mn; fs[i].mx = c0.mx; fs[i].has_range = true; }
        else {
          if (cmp_host(c0.mn, fs[i].mn, t0) < 0) fs[i].mn = c0.mn;
          if (cmp_host(c0.mx, fs[i].mx, t0) > 0) fs[i].mx = c0.mx;
        }
      } else ranges_ok = false;
    }
    total_rgs += fs[i].rgs.size();
  }
  // PK-disjointness from pk0 chunk statistics: order files by min(pk0); require max_f < min_{f+1} strictly.
  std::vector<size_t> order(n);
  for (size_t i = 0; i < n; i++) order[i] = i;
  plan->disjoint = false;
  if (n <= 1) plan->disjoint = true;
  else if (ranges_ok) {
    std::vector<size_t> nonempty;
    for (size_t i = 0; i < n; i++) if (!fs[i].rgs.empty()) nonempty.push_back(i);
    std::stable_sort(nonempty.begin(), nonempty.end(), [&](size_t a, size_t b) { return cmp_host(fs[a].mn, fs[b].mn, t0) < 0; });
    bool ok = true;
    for (size_t j = 0; j + 1 < nonempty.size() && ok; j++)
      ok = cmp_host(fs[nonempty[j]].mx, fs[nonempty[j + 1]].mn, t0) < 0;
    if (ok) {
      plan->disjoint = true;
      order.clear();
      for (size_t i : nonempty) order.push_back(i);
      for (size_t i = 0; i < n; i++) if (fs[i].rgs.empty()) order.push_back(i);
    }
  }
  plan->col_has_nulls.assign(schema->num_columns, false);
  plan->sel.reserve(total_rgs);
  std::vector<uint8_t> has_nulls(schema->num_columns, 0);
  uint64_t row = 0, scratch = 0;
  for (size_t oi = 0; oi < order.size(); oi++) {
    FileSel& f = fs[order[oi]];
    plan->files.push_back(f.f);
    plan->file_base.push_back(uint32_t(row));
    const size_t ncols = size_t(f.f->meta.ncols);
    for (uint32_t g : f.rgs) {
      const uint32_t rows = f.f->rg_rows[g];
      const RgCol* rc = &f.f->rgcol[size_t(g) * ncols];
      RgSel s;
      s.sst = uint32_t(oi);
      s.rg = g;
      s.out_row = uint32_t(row);
      s.num_rows = rows;
      s.scratch_off = scratch;
      for (uint32_t c : need_cols) {
        const RgCol& cc = rc[c];
        scratch += cc.scratch;             // 0 for uncompressed PLAIN chunks
        if (!cc.null_none) has_nulls[c] = 1;
        if (!cc.simple_page) plan->all_single_plain_page = false;
      }
      plan->sel.push_back(s);
      if (n == 1) {
        for (uint64_t b = e->batch_size; b < rows; b += e->batch_size) plan->piece_end.push_back(uint32_t(row + b));
        plan->piece_end.push_back(uint32_t(row + rows));
      }
      row += rows;
      if (row >= 0xfffffff0ull) return set_error(HG_ERR_UNSUPPORTED, "more than 2^32 rows in one scan call");
    }
  }
  for (uint32_t c = 0; c < schema->num_columns; c++) plan->col_has_nulls[c] = has_nulls[c] != 0;
  plan->file_base.push_back(uint32_t(row));
  plan->rows_decoded = row;
  plan->scratch_bytes = scratch;
  return HG_OK;
}

// ------------------------------------------------------------------------------------------- the general pipeline
struct DecodedCol {
  DevBuf vals, valid, lens;            // lens: Binary columns only (vals = one byte pointer per row)
  uint32_t type = 0, width = 0;
  bool present = false;
  ColView view() const { return ColView{vals.p, reinterpret_cast<const uint8_t*>(valid.p), type, width, reinterpret_cast<const uint32_t*>(lens.p)}; }
};

struct PipelineState {
  ScanPlan plan;
  std::vector<DecodedCol> cols;       // indexed by schema column
  uint32_t N = 0;                     // decoded rows (capacity of every row-indexed buffer)
  DevBuf d_ssts, d_sel, d_colsel, d_scratch, d_err, d_counters;   // counters: [0]=M survivors [1]=R outputs [2]=G groups
  DevBuf alive, surv, keep, out_pos, out_rows, tmp, run_start, file_base, recA, recB, order, chunk_end, piece_end, bound;
  uint32_t nchunks = 0;
  const uint32_t* surv_ptr = nullptr;    // nullptr = identity
  bool keep_order = false;               // Append mode: the export needs the merged order of ALL surviving rows
  const uint32_t* order_ptr = nullptr;   //   row id of the t-th surviving row in merged order (nullptr = identity), valid when keep_order
  const uint32_t* d_m = nullptr;
  const uint32_t* d_r = nullptr;
  const uint32_t* d_g = nullptr;
  uint32_t* counters() const { return d_counters.as<uint32_t>(); }
};

static int validate_schema(const hg_schema_desc* s) {
  if (!s || !s->types) return set_error(HG_ERR_INVALID, "null schema");
  if (s->num_primary_keys == 0) return set_error(HG_ERR_INVALID, "num_primary_keys should large than 0");  // types.rs:165
  if (s->num_columns < s->num_primary_keys + 3 || s->num_columns > uint32_t(MAX_COLS))
    return set_error(HG_ERR_INVALID, "schema needs pk columns, at least one value column and the two builtin columns");
  if (s->num_primary_keys > uint32_t(MAX_PK)) return set_error(HG_ERR_UNSUPPORTED, "more than 4 primary key columns");
  if (s->update_mode > HG_UPDATE_APPEND) return set_error(HG_ERR_INVALID, "bad update mode");
  uint32_t pk_bytes = 0;
  for (uint32_t c = 0; c < s->num_columns; c++) if (s->types[c] > T_BINARY) return set_error(HG_ERR_INVALID, "bad column type");
  if (s->update_mode == HG_UPDATE_APPEND)       // read.rs:485-490 + operator.rs:66-73: BytesMergeOperator over ALL value columns
    for (uint32_t c = s->num_primary_keys; c + 2 < s->num_columns; c++)
      if (s->types[c] != T_BINARY) return set_error(HG_ERR_INVALID, "MergeOperator is only used for binary column (UpdateMode::Append)");
  for (uint32_t c = 0; c < s->num_primary_keys; c++) {
    uint32_t t = s->types[c];
    // primary_key_eq supports exactly these (read.rs:269-286); other types silently compare equal there — fenced off here
    if (!(t == T_U8 || t == T_I8 || t == T_U32 || t == T_I32 || t == T_U64 || t == T_I64))
      return set_error(HG_ERR_UNSUPPORTED, "primary key type not supported by the reference's primary_key_eq");
    pk_bytes += type_width_host(t);
  }
  if (pk_bytes > 16) return set_error(HG_ERR_UNSUPPORTED, "primary key wider than 128 bits");
  if (s->types[s->num_columns - 2] != T_U64 || s->types[s->num_columns - 1] != T_U64)
    return set_error(HG_ERR_INVALID, "builtin columns must be UInt64");
  return HG_OK;
}

static int validate_preds(const hg_schema_desc* s, const hg_predicate* preds, size_t np) {
  if (np > size_t(MAX_PREDS)) return set_error(HG_ERR_UNSUPPORTED, "more than 8 predicates");
  for (size_t i = 0; i < np; i++) {
    if (preds[i].column >= s->num_columns) return set_error(HG_ERR_INVALID, "predicate column out of range");
    if (preds[i].op > HG_OP_IN) return set_error(HG_ERR_UNSUPPORTED, "predicate operator");
    if (s->types[preds[i].column] == T_BINARY) return set_error(HG_ERR_UNSUPPORTED, "predicates on Binary columns are not implemented on the GPU path");
    if (preds[i].op == HG_OP_IN && (preds[i].in_count > HG_MAX_IN_LIST || (preds[i].in_count && !preds[i].in_values)))
      return set_error(HG_ERR_INVALID, "IN list: null pointer or more than HG_MAX_IN_LIST values");
  }
  return HG_OK;
}

// Runs decode -> filter -> merge -> dedup.  On return st->out_rows holds the surviving row ids in stream order,
// counters()[0..1] = M, R (device), st->out_pos their merged positions.
static int run_pipeline(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n, const hg_predicate* preds,
                        size_t np, std::vector<uint32_t> need_cols, bool want_batches, PipelineState* st) {
  cudaStream_t s = e->stream;
  Launch L = e->L();
  // PKs are always needed (dedup); __seq__ whenever a real merge happens (decided after planning, so include it
  // only if more than one SST is given).
  for (uint32_t c = 0; c < schema->num_primary_keys; c++) need_cols.push_back(c);
  for (size_t i = 0; i < np; i++) need_cols.push_back(preds[i].column);
  std::sort(need_cols.begin(), need_cols.end());
  need_cols.erase(std::unique(need_cols.begin(), need_cols.end()), need_cols.end());
  const uint32_t seq_idx = schema->num_columns - 2;
  {
    std::vector<uint32_t> probe = need_cols;
    ScanPlan trial;
    // plan once without seq to learn disjointness, then add seq if a merge is required
    int rc = build_plan(e, schema, ssts, n, preds, np, probe, &trial);
    if (rc) return rc;
    if (!trial.disjoint && std::find(need_cols.begin(), need_cols.end(), seq_idx) == need_cols.end()) {
      need_cols.push_back(seq_idx);
      std::sort(need_cols.begin(), need_cols.end());
      ScanPlan again;
      rc = build_plan(e, schema, ssts, n, preds, np, need_cols, &again);
      if (rc) return rc;
      st->plan = std::move(again);
    } else st->plan = std::move(trial);
  }
  static const bool trace = getenv("HORAE_TRACE") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
  auto tp0 = now();
  ScanPlan& plan = st->plan;
  const uint32_t N = uint32_t(plan.rows_decoded);
  st->N = N;
  const int k = int(n);

  CU_TRY(st->d_counters.alloc(8 * sizeof(uint32_t), s));
  CU_TRY(cudaMemsetAsync(st->d_counters.p, 0, 8 * sizeof(uint32_t), s));
  CU_TRY(st->d_err.alloc(sizeof(int), s));
  CU_TRY(cudaMemsetAsync(st->d_err.p, 0, sizeof(int), s));
  st->d_m = st->counters() + 0;
  st->d_r = st->counters() + 1;
  st->d_g = st->counters() + 2;

  // --- S2: decode
  st->cols.resize(schema->num_columns);
  std::vector<ColSel> colsel;
  for (uint32_t c : need_cols) {
    DecodedCol& dc = st->cols[c];
    dc.type = schema->types[c];
    dc.width = type_width_host(dc.type);
    dc.present = true;
    CU_TRY(dc.vals.alloc(size_t(N) * dc.width + 16, s));
    if (plan.col_has_nulls[c] || dc.type == T_BINARY) CU_TRY(dc.valid.alloc(size_t(N) + 16, s));
    if (dc.type == T_BINARY) CU_TRY(dc.lens.alloc(size_t(N) * 4 + 16, s));
    ColSel cs;
    cs.col = c;
    cs.type = dc.type;
    cs.out_width = dc.width;
    cs._pad = 0;
    cs.out_vals = dc.vals.p;
    cs.out_valid = reinterpret_cast<uint8_t*>(dc.valid.p);
    cs.out_lens = reinterpret_cast<uint32_t*>(dc.lens.p);
    colsel.push_back(cs);
  }
  if (N > 0) {
    std::vector<SstDev> sd(plan.files.size());
    for (size_t i = 0; i < plan.files.size(); i++) {
      SstResident* f = plan.files[i];
      sd[i] = SstDev{f->d_bytes, f->d_pages, f->d_chunks, uint32_t(f->meta.ncols), uint32_t(f->meta.rgs.size())};
    }
    CU_TRY(st->d_ssts.alloc(sd.size() * sizeof(SstDev), s));
    CU_TRY(st->d_sel.alloc(plan.sel.size() * sizeof(RgSel), s));
    CU_TRY(st->d_colsel.alloc(colsel.size() * sizeof(ColSel), s));
    size_t stage_off = 0;
    int urc = stage_upload(e, st->d_ssts.p, sd.data(), sd.size() * sizeof(SstDev), &stage_off);
    if (!urc) urc = stage_upload(e, st->d_sel.p, plan.sel.data(), plan.sel.size() * sizeof(RgSel), &stage_off);
    if (!urc) urc = stage_upload(e, st->d_colsel.p, colsel.data(), colsel.size() * sizeof(ColSel), &stage_off);
    if (urc) return urc;
    // the host vectors must outlive the async copies: pageable memcpy is staged synchronously by the runtime
    auto tp1 = now();
    CU_TRY(cudaEventRecord(e->evk0, s));
    if (plan.scratch_bytes) {
      CU_TRY(st->d_scratch.alloc(plan.scratch_bytes + 64, s));
      static const bool snappy_v1 = getenv("HORAE_SNAPPY_V1") != nullptr;   // developer A/B switch
      if (snappy_v1)
        k::snappy_chunks(L, st->d_ssts.as<SstDev>(), st->d_sel.as<RgSel>(), uint32_t(plan.sel.size()), st->d_colsel.as<ColSel>(),
                         int(colsel.size()), st->d_scratch.as<uint8_t>(), st->d_err.as<int>());
      else
        k::snappy_chunks_v2(L, st->d_ssts.as<SstDev>(), st->d_sel.as<RgSel>(), uint32_t(plan.sel.size()), st->d_colsel.as<ColSel>(),
                            int(colsel.size()), st->d_scratch.as<uint8_t>(), st->counters() + 4, st->d_err.as<int>());
      bool any_zstd = false;
      for (const SstResident* f : plan.files) any_zstd = any_zstd || f->any_zstd;
      if (any_zstd)                                          // ParquetCompression::Zstd (config.rs:78-94): its own kernel, same scratch layout
        k::zstd_chunks(L, st->d_ssts.as<SstDev>(), st->d_sel.as<RgSel>(), uint32_t(plan.sel.size()), st->d_colsel.as<ColSel>(),
                       int(colsel.size()), st->d_scratch.as<uint8_t>(), st->counters() + 6, st->d_err.as<int>());
    }
    k::decode_chunks(L, st->d_ssts.as<SstDev>(), st->d_sel.as<RgSel>(), uint32_t(plan.sel.size()), st->d_colsel.as<ColSel>(),
                     int(colsel.size()), st->d_scratch.as<uint8_t>(), st->d_err.as<int>());
    CU_TRY(cudaEventRecord(e->evk1, s));
    bool rows_point_into_scratch = false;                   // Binary rows are pointers to their bytes in place (page or decompression scratch)
    for (uint32_t c : need_cols) if (schema->types[c] == T_BINARY) rows_point_into_scratch = true;
    if (!rows_point_into_scratch) st->d_scratch.reset();
    if (trace) fprintf(stderr, "[general] col alloc+upload %.0f us, scratch alloc (%.1f MB) + decode launches %.0f us\n", us(tp0, tp1), plan.scratch_bytes / 1e6, us(tp1, now()));
  }

  // --- S3: filter (before the merge, read.rs:459-470)
  CU_TRY(st->tmp.alloc(k::compact_tmp_elems(N) * sizeof(uint32_t), s));
  if (np > 0 && N > 0) {
    PredSet ps;
    ps.n = int(np);
    for (size_t i = 0; i < np; i++) {
      ps.p[i].col = st->cols[preds[i].column].view();
      ps.p[i].op = preds[i].op;
      ps.p[i].n_in = 0;
      ps.p[i].in_list = nullptr;
      ps.p[i].lit = pred_literal(preds[i], schema->types[preds[i].column]);
      if (preds[i].op == HG_OP_IN) {
        uint64_t* d_list = static_cast<uint64_t*>(g_arena->alloc(std::max<size_t>(preds[i].in_count, 1) * 8));
        if (!d_list) return set_error(HG_ERR_OOM, "out of device memory");
        if (preds[i].in_count) {
          int urc = stage_upload(e, d_list, preds[i].in_values, size_t(preds[i].in_count) * 8, nullptr);
          if (urc) return urc;
        }
        ps.p[i].n_in = preds[i].in_count;
        ps.p[i].in_list = d_list;
      }
    }
    CU_TRY(st->alive.alloc(size_t(N) + 16, s));
    CU_TRY(st->surv.alloc(size_t(N) * 4 + 16, s));
    k::eval_predicates(L, ps, N, st->alive.as<uint8_t>());
    k::compact_flags(L, st->alive.as<uint8_t>(), N, st->tmp.as<uint32_t>(), st->surv.as<uint32_t>(), st->counters() + 0);
    st->alive.reset();
    st->surv_ptr = st->surv.as<uint32_t>();
  } else {
    k::fill_u32(L, st->counters() + 0, N, 1);
    st->surv_ptr = nullptr;
  }

  // --- S4: merge on (pk..., __seq__) when the inputs are not provably PK-disjoint
  PkSet pk;
  pk.n = int(schema->num_primary_keys);
  for (int c = 0; c < pk.n; c++) {
    pk.c[c] = st->cols[c].view();
    if (plan.col_has_nulls[c]) return set_error(HG_ERR_UNSUPPORTED, "NULL primary keys are not supported on the GPU path");
  }
  const uint32_t* order = st->surv_ptr;
  CU_TRY(st->keep.alloc(size_t(N) + 16, s));
  CU_TRY(cudaEventRecord(e->evm0, s));
  if (!plan.disjoint && N > 0) {
    CU_TRY(st->file_base.alloc((k + 1) * sizeof(uint32_t), s));
    CU_TRY(st->run_start.alloc((k + 2) * sizeof(uint32_t), s));
    CU_TRY(cudaMemcpyAsync(st->file_base.p, plan.file_base.data(), (k + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    k::survivor_run_starts(L, st->surv_ptr, st->d_m, st->file_base.as<uint32_t>(), k, st->run_start.as<uint32_t>());
    // single pass over packed 64-bit keys when (pk..., __seq__, stream) fit in 52 bits after rebasing to the chunk statistics
    k::KeyPack kp;
    bool packed = k <= k::kMaxMergeRuns && !(e->flags & HG_FLAG_PAIRWISE_MERGE);
    if (packed) {
      std::memset(&kp, 0, sizeof(kp));
      const int npk = int(schema->num_primary_keys);
      uint64_t lo[MAX_PK + 1], hi[MAX_PK + 1];
      bool seen_c[MAX_PK + 1] = {false};
      bool seen = false, seq_nullable = false;
      for (int c = 0; c <= MAX_PK; c++) { lo[c] = 0; hi[c] = 0; }
      for (const RgSel& rs : plan.sel) {
        const SstResident* f = plan.files[rs.sst];
        const RgCol* rc = &f->rgcol[size_t(rs.rg) * size_t(f->meta.ncols)];
        for (int c = 0; c <= npk && packed; c++) {
          const uint32_t col = c < npk ? uint32_t(c) : seq_idx;
          const RgCol& x = rc[col];
          if (c == npk && x.null_all) { seq_nullable = true; continue; }
          if (!x.has_minmax) { packed = false; break; }
          const uint64_t flip = (c < npk && type_is_signed(schema->types[col])) ? (1ull << 63) : 0ull;
          const uint64_t a = x.mn ^ flip, b = x.mx ^ flip;
          if (c == npk && !x.null_none) seq_nullable = true;
          if (!seen_c[c] || a < lo[c]) lo[c] = a;
          if (!seen_c[c] || b > hi[c]) hi[c] = b;
          seen_c[c] = true;
        }
        seen = true;
      }
      if (!seen_c[npk]) { seq_nullable = true; lo[npk] = 0; hi[npk] = 0; }      // every __seq__ chunk is all-null
      if (packed && seen) {
        auto bits = [](uint64_t span) { int b = 0; while (span) { b++; span >>= 1; } return b; };
        int rb = 0;
        while ((1 << rb) < k) rb++;
        // seq lives in the (value + 1, NULL = 0) domain
        if (hi[npk] == ~0ull) packed = false;
        kp.seq_min = seq_nullable ? 0 : lo[npk] + 1;
        kp.seq_span = seen_c[npk] ? hi[npk] + 1 - kp.seq_min : 0;
        kp.seq_shift = uint32_t(rb);
        int used = rb + bits(kp.seq_span);
        kp.pk_shift = uint32_t(used);
        for (int c = npk - 1; c >= 0 && packed; c--) {
          kp.mn[c] = lo[c];
          kp.span[c] = hi[c] - lo[c];
          kp.shift[c] = uint32_t(used);
          used += bits(kp.span[c]);
        }
        if (used > 52) packed = false;
      } else packed = false;
    }
    if (packed) {
      uint32_t ranges = 1;
      DevBuf ktmp;
      CU_TRY(ktmp.alloc(k::kway_tmp_bytes(N, k, &ranges), s));
      CU_TRY(st->order.alloc(size_t(N) * 4 + 16, s));
      k::kway_merge(L, pk, st->cols[seq_idx].view(), st->surv_ptr, st->d_m, N, st->run_start.as<uint32_t>(), k, kp, ktmp.p,
                    st->counters() + 5, st->order.as<uint32_t>(), st->keep.as<uint8_t>(), st->d_err.as<int>());
      order = st->order.as<uint32_t>();
      st->surv_ptr = nullptr;
    } else {
    CU_TRY(st->recA.alloc(size_t(N) * sizeof(SortRec) + 32, s));
    CU_TRY(st->recB.alloc(size_t(N) * sizeof(SortRec) + 32, s));
    k::build_records(L, pk, st->cols[seq_idx].view(), st->surv_ptr, st->d_m, N, st->recA.as<SortRec>());
    DevBuf splits;
    CU_TRY(splits.alloc(k::merge_split_elems(N) * sizeof(uint32_t), s));
    SortRec* src = st->recA.as<SortRec>();
    SortRec* dst = st->recB.as<SortRec>();
    for (int level = 0; (1 << level) < k; level++) {
      k::merge_pass(L, src, dst, st->run_start.as<uint32_t>(), k, level, st->d_m, N, splits.as<uint32_t>());
      std::swap(src, dst);
    }
    CU_TRY(st->order.alloc(size_t(N) * 4 + 16, s));
    k::records_to_rows(L, src, st->d_m, N, st->order.as<uint32_t>());
    k::dedup_flags_recs(L, src, st->d_m, N, st->keep.as<uint8_t>());
    order = st->order.as<uint32_t>();
    st->recA.reset();
    st->recB.reset();
    st->surv_ptr = nullptr;  // (no longer valid)
    }
    st->surv.reset();
  } else if (N > 0) {
    k::dedup_flags_cols(L, pk, order, st->d_m, N, st->keep.as<uint8_t>());
  }

  // --- batch boundaries of MergeStream need the chunking of its INPUT (computed before surv is dropped)
  if (want_batches && N > 0) {
    if (k == 1) {
      st->nchunks = uint32_t(plan.piece_end.size());
      CU_TRY(st->piece_end.alloc(st->nchunks * sizeof(uint32_t) + 16, s));
      CU_TRY(st->chunk_end.alloc(st->nchunks * sizeof(uint32_t) + 16, s));
      CU_TRY(cudaMemcpyAsync(st->piece_end.p, plan.piece_end.data(), st->nchunks * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
      k::chunk_ends_from_rows(L, order /* == surv or identity */, st->d_m, st->piece_end.as<uint32_t>(), st->nchunks,
                              st->chunk_end.as<uint32_t>());
    } else {
      st->nchunks = (N + e->batch_size - 1) / e->batch_size;
      CU_TRY(st->chunk_end.alloc(st->nchunks * sizeof(uint32_t) + 16, s));
      k::uniform_chunk_ends(L, st->d_m, e->batch_size, st->nchunks, st->chunk_end.as<uint32_t>());
    }
  }

  // --- S5/S6: keep the last row of every PK run
  CU_TRY(st->out_pos.alloc(size_t(N) * 4 + 16, s));
  CU_TRY(st->out_rows.alloc(size_t(N) * 4 + 16, s));
  if (N > 0) {
    // compaction over the first M flags only: flags beyond M are stale, so clear the tail by bounding n on device
    k::clear_tail(L, st->keep.as<uint8_t>(), st->d_m, N);
    k::compact_flags(L, st->keep.as<uint8_t>(), N, st->tmp.as<uint32_t>(), st->out_pos.as<uint32_t>(), st->counters() + 1);
    k::gather_rows(L, order, st->out_pos.as<uint32_t>(), st->d_r, N, st->out_rows.as<uint32_t>());
    if (want_batches && st->nchunks) {
      CU_TRY(st->bound.alloc(st->nchunks * sizeof(uint32_t) + 16, s));
      k::batch_bounds(L, st->out_pos.as<uint32_t>(), st->d_r, st->chunk_end.as<uint32_t>(), st->nchunks, st->bound.as<uint32_t>());
    }
  }
  CU_TRY(cudaEventRecord(e->evm1, s));
  st->keep.reset();
  if (st->keep_order) { st->order_ptr = order; return HG_OK; }       // (order points into st->order or st->surv: both stay allocated)
  st->order.reset();
  st->surv.reset();
  return HG_OK;
}

static int check_device_error(hg_engine* e, PipelineState* st) {
  int herr = 0;
  CU_TRY(cudaMemcpyAsync(&herr, st->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(cudaStreamSynchronize(e->stream));
  if (herr == 130) return set_error(HG_ERR_INVALID, "failed to construct RecordBatch in BytesMergeOperator (a run of several rows whose Binary values are all empty)");
  if (herr) return set_error(HG_ERR_FORMAT, "device decode error code " + std::to_string(herr));
  return HG_OK;
}

// ----------------------------------------------------------------------------------------------- pinned host pool
// Result batches live in pinned host memory owned by the Arrow stream; cudaMallocHost is slow and synchronising, so
// released buffers go back to a process-wide free list (bounded) instead of being freed.
namespace {
struct PinnedPool {
  std::mutex mu;
  std::vector<std::pair<size_t, void*>> free_list;
  size_t pooled = 0;
  static constexpr size_t kMaxPooled = size_t(16) << 30;
  void* alloc(size_t bytes) {
    bytes = (bytes + 4095) & ~size_t(4095);
    {
      std::lock_guard<std::mutex> g(mu);
      size_t best = free_list.size();
      for (size_t i = 0; i < free_list.size(); i++)
        if (free_list[i].first >= bytes && free_list[i].first <= bytes * 2 + (1 << 20) && (best == free_list.size() || free_list[i].first < free_list[best].first)) best = i;
      if (best != free_list.size()) {
        void* p = free_list[best].second;
        pooled -= free_list[best].first;
        sizes[p] = free_list[best].first;
        free_list.erase(free_list.begin() + best);
        return p;
      }
    }
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> g(mu);
    sizes[p] = bytes;
    return p;
  }
  void release(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> g(mu);
    auto it = sizes.find(p);
    size_t b = it == sizes.end() ? 0 : it->second;
    if (it != sizes.end()) sizes.erase(it);
    if (b && pooled + b <= kMaxPooled) { free_list.emplace_back(b, p); pooled += b; }
    else cudaFreeHost(p);
  }
  std::unordered_map<void*, size_t> sizes;
};
PinnedPool& pinned_pool() { static PinnedPool* p = new PinnedPool(); return *p; }
}  // namespace

// --------------------------------------------------------------------------------------------- Arrow C stream export
struct HostColumn {
  std::string name;
  uint32_t type = 0, width = 0;
  void* vals = nullptr;          // pinned host: the values — for Binary columns the concatenated bytes
  uint8_t* bitmap = nullptr;     // pinned host, nullptr = no nulls
  int32_t* offsets = nullptr;    // pinned host, Binary columns only: Arrow offsets (rows + 1)
  int64_t null_count = 0;
};
struct StreamData {
  std::vector<HostColumn> cols;
  std::vector<uint32_t> batch_start;  // nb+1 row offsets
  size_t next = 0;
  std::string last_error;
  ~StreamData() {
    for (auto& c : cols) {
      pinned_pool().release(c.vals);
      pinned_pool().release(c.bitmap);
      pinned_pool().release(c.offsets);
    }
  }
};
struct StreamPriv { std::shared_ptr<StreamData> data; };

static void schema_release(struct ArrowSchema* s) {
  if (!s || !s->release) return;
  for (int64_t i = 0; i < s->n_children; i++) {
    if (s->children[i]->release) s->children[i]->release(s->children[i]);
    delete s->children[i];
  }
  delete[] s->children;
  delete reinterpret_cast<std::string*>(s->private_data);
  s->release = nullptr;
}
static void fill_schema(struct ArrowSchema* out, const std::shared_ptr<StreamData>& d) {
  std::memset(out, 0, sizeof(*out));
  out->format = "+s";
  out->name = "";
  out->n_children = int64_t(d->cols.size());
  out->children = new ArrowSchema*[d->cols.size() ? d->cols.size() : 1];
  for (size_t i = 0; i < d->cols.size(); i++) {
    ArrowSchema* c = new ArrowSchema();
    std::memset(c, 0, sizeof(*c));
    std::string* nm = new std::string(d->cols[i].name);
    c->format = arrow_format(d->cols[i].type);
    c->name = nm->c_str();
    c->flags = ARROW_FLAG_NULLABLE;
    c->private_data = nm;
    c->release = [](struct ArrowSchema* s) { delete reinterpret_cast<std::string*>(s->private_data); s->release = nullptr; };
    out->children[i] = c;
  }
  out->private_data = nullptr;
  out->release = schema_release;
}
struct ArrayPriv { std::shared_ptr<StreamData> keep; const void* bufs[3]; };
static void child_release(struct ArrowArray* a) {
  if (!a || !a->release) return;
  delete reinterpret_cast<ArrayPriv*>(a->private_data);
  a->release = nullptr;
}
static void batch_release(struct ArrowArray* a) {
  if (!a || !a->release) return;
  for (int64_t i = 0; i < a->n_children; i++) {
    if (a->children[i]->release) a->children[i]->release(a->children[i]);
    delete a->children[i];
  }
  delete[] a->children;
  delete reinterpret_cast<ArrayPriv*>(a->private_data);
  a->release = nullptr;
}
static int stream_get_schema(struct ArrowArrayStream* st, struct ArrowSchema* out) {
  fill_schema(out, reinterpret_cast<StreamPriv*>(st->private_data)->data);
  return 0;
}
static int stream_get_next(struct ArrowArrayStream* st, struct ArrowArray* out) {
  auto d = reinterpret_cast<StreamPriv*>(st->private_data)->data;
  std::memset(out, 0, sizeof(*out));
  if (d->next + 1 >= d->batch_start.size()) { out->release = nullptr; return 0; }  // end of stream
  uint32_t lo = d->batch_start[d->next], hi = d->batch_start[d->next + 1];
  d->next++;
  ArrayPriv* top = new ArrayPriv{d, {nullptr, nullptr, nullptr}};
  out->length = hi - lo;
  out->null_count = 0;
  out->offset = 0;
  out->n_buffers = 1;
  out->buffers = top->bufs;
  out->n_children = int64_t(d->cols.size());
  out->children = new ArrowArray*[d->cols.size() ? d->cols.size() : 1];
  for (size_t i = 0; i < d->cols.size(); i++) {
    ArrowArray* c = new ArrowArray();
    std::memset(c, 0, sizeof(*c));
    const bool bin = d->cols[i].type == T_BINARY;          // Binary: validity, int32 offsets, data
    ArrayPriv* p = bin ? new ArrayPriv{d, {d->cols[i].bitmap, d->cols[i].offsets, d->cols[i].vals}}
                       : new ArrayPriv{d, {d->cols[i].bitmap, d->cols[i].vals, nullptr}};
    c->length = hi - lo;
    c->offset = lo;
    c->null_count = d->cols[i].bitmap ? -1 : 0;
    c->n_buffers = bin ? 3 : 2;
    c->buffers = p->bufs;
    c->private_data = p;
    c->release = child_release;
    out->children[i] = c;
  }
  out->private_data = top;
  out->release = batch_release;
  return 0;
}
static const char* stream_last_error(struct ArrowArrayStream* st) {
  auto d = reinterpret_cast<StreamPriv*>(st->private_data)->data;
  return d->last_error.empty() ? nullptr : d->last_error.c_str();
}
static void stream_release(struct ArrowArrayStream* st) {
  if (!st || !st->release) return;
  delete reinterpret_cast<StreamPriv*>(st->private_data);
  st->release = nullptr;
}
static void make_stream(struct ArrowArrayStream* out, std::shared_ptr<StreamData> d) {
  out->get_schema = stream_get_schema;
  out->get_next = stream_get_next;
  out->get_last_error = stream_last_error;
  out->release = stream_release;
  out->private_data = new StreamPriv{std::move(d)};
}

static const char* col_name(const hg_schema_desc* s, uint32_t c, std::string* tmp) {
  if (s->names && s->names[c]) return s->names[c];
  *tmp = "c" + std::to_string(c);
  return tmp->c_str();
}

// ------------------------------------------------------------------------------------------------------------ C ABI
extern "C" {

uint32_t hg_abi_version(void) { return HG_ABI_VERSION; }
const char* hg_last_error(void) { return g_last_error.c_str(); }

int hg_engine_create(const hg_config* cfg, hg_engine** out) {
  HG_GUARD_BEGIN
  if (!cfg || !out) return set_error(HG_ERR_INVALID, "null argument");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return set_error(HG_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(ce) + " (this library has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return set_error(HG_ERR_INVALID, "device ordinal out of range");
  CU_TRY(cudaSetDevice(cfg->device));
  auto e = std::make_unique<hg_engine>();
  e->device = cfg->device;
  e->batch_size = cfg->batch_size ? cfg->batch_size : 8192;
  e->flags = cfg->flags;
  e->budget = cfg->hbm_budget_bytes;
  CU_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreate(&e->ev0));
  CU_TRY(cudaEventCreate(&e->ev1));
  CU_TRY(cudaEventCreate(&e->evk0));
  CU_TRY(cudaEventCreate(&e->evk1));
  CU_TRY(cudaEventCreate(&e->evm0));
  CU_TRY(cudaEventCreate(&e->evm1));
  CU_TRY(cudaEventCreate(&e->evd0));
  CU_TRY(cudaEventCreate(&e->evd1));
  cudaMemPool_t pool;
  CU_TRY(cudaDeviceGetDefaultMemPool(&pool, cfg->device));
  uint64_t thresh = UINT64_MAX;
  CU_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  *out = e.release();
  return HG_OK;
  HG_GUARD_END
}

void hg_engine_destroy(hg_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  hg_comm_free(e->comm);
  e->comm = nullptr;
  g_arena = nullptr;
  e->arena.destroy();
  if (e->h_stage) cudaFreeHost(e->h_stage);
  if (e->h_small) cudaFreeHost(e->h_small);
  e->ssts.clear();
  cudaEventDestroy(e->ev0);
  cudaEventDestroy(e->ev1);
  cudaEventDestroy(e->evk0);
  cudaEventDestroy(e->evk1);
  cudaEventDestroy(e->evm0);
  cudaEventDestroy(e->evm1);
  cudaEventDestroy(e->evd0);
  cudaEventDestroy(e->evd1);
  cudaStreamDestroy(e->stream);
  delete e;
}

void* hg_engine_stream(hg_engine* e) { return e ? reinterpret_cast<void*>(e->stream) : nullptr; }
int hg_engine_set_flags(hg_engine* e, uint32_t flags) {
  HG_GUARD_BEGIN
  if (!e) return set_error(HG_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> g(e->mu);
  e->flags = flags;
  return HG_OK;
  HG_GUARD_END
}

int hg_sst_load(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* sst) {
  HG_GUARD_BEGIN
  if (!e || !schema || !sst) return set_error(HG_ERR_INVALID, "null argument");
  int rc = validate_schema(schema);
  if (rc) return rc;
  std::lock_guard<std::mutex> g(e->mu);
  CU_TRY(cudaSetDevice(e->device));
  return load_sst_locked(e, schema, sst);
  HG_GUARD_END
}

int hg_sst_unload(hg_engine* e, uint64_t id) {
  HG_GUARD_BEGIN
  if (!e) return set_error(HG_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> g(e->mu);
  auto it = e->ssts.find(id);
  if (it == e->ssts.end()) return set_error(HG_ERR_NOT_FOUND, "sst not resident");
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  e->resident_bytes -= it->second->device_bytes;
  e->ssts.erase(it);
  return HG_OK;
  HG_GUARD_END
}

int hg_sst_resident_bytes(hg_engine* e, uint64_t* out) {
  HG_GUARD_BEGIN
  if (!e || !out) return set_error(HG_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> g(e->mu);
  *out = e->resident_bytes;
  return HG_OK;
  HG_GUARD_END
}

int hg_agg_export_packed(hg_engine* e, void* d_dst, uint64_t cap) {
  HG_GUARD_BEGIN
  if (!e || !d_dst) return set_error(HG_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> g(e->mu);
  if (cap < e->last_agg.num_groups) return set_error(HG_ERR_INVALID, "capacity smaller than the number of groups");
  CU_TRY(cudaSetDevice(e->device));
  AggOut in{const_cast<void*>(e->last_agg.d_gkey), const_cast<int64_t*>(e->last_agg.d_bucket), const_cast<uint64_t*>(e->last_agg.d_count),
            const_cast<double*>(e->last_agg.d_sum), const_cast<double*>(e->last_agg.d_min), const_cast<double*>(e->last_agg.d_max)};
  Launch L = e->L();
  k::pack_agg(L, in, e->last_gwidth, e->last_agg.num_groups, cap, static_cast<long long*>(d_dst));
  return HG_OK;
  HG_GUARD_END
}

int hg_plan_row_groups(const hg_schema_desc* schema, const uint8_t* data, uint64_t size, const hg_predicate* preds, size_t n_preds,
                       uint8_t* keep, uint32_t cap, uint32_t* num_row_groups) {
  HG_GUARD_BEGIN
  if (!data || !keep || !num_row_groups || (n_preds && !preds)) return set_error(HG_ERR_INVALID, "null argument");
  int rc = validate_schema(schema);
  if (rc) return rc;
  rc = validate_preds(schema, preds, n_preds);
  if (rc) return rc;
  SstResident r;
  std::vector<PageDev> pages;
  std::vector<ChunkDev> chunks;
  std::string err;
  rc = prepare_sst(schema, 0, data, size, &r, &pages, &chunks, &err);
  if (rc) return set_error(rc, err);
  const size_t nrg = r.rg_rows.size(), ncols = size_t(r.meta.ncols);
  *num_row_groups = uint32_t(nrg);
  if (nrg > cap) return set_error(HG_ERR_INVALID, "keep[] is smaller than the number of row groups");
  uint64_t lits[MAX_PREDS];
  for (size_t i = 0; i < n_preds; i++) lits[i] = pred_literal(preds[i], schema->types[preds[i].column]);
  for (size_t g = 0; g < nrg; g++)
    keep[g] = r.rg_rows[g] > 0 && (n_preds == 0 || rg_may_match(&r.rgcol[g * ncols], r.rg_rows[g], schema, preds, lits, n_preds)) ? 1 : 0;
  return HG_OK;
  HG_GUARD_END
}

int hg_last_stats(hg_engine* e, hg_scan_stats* out) {
  HG_GUARD_BEGIN
  if (!e || !out) return set_error(HG_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> g(e->mu);
  *out = e->stats;
  return HG_OK;
  HG_GUARD_END
}

static void end_call(hg_engine* e) {
  for (uint64_t id : e->transient_ids) e->ssts.erase(id);     // transient SSTs live in the arena: nothing to free
  e->transient_ids.clear();
}
struct CallGuard { hg_engine* e; ~CallGuard() { end_call(e); } };

// need_cols: the columns this call can touch (only used to select the byte ranges of non-resident SSTs)
static int begin_call(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n, const hg_predicate* preds, size_t np,
                      std::vector<uint32_t> need_cols, bool seq_if_overlap, uint32_t trunc_mask = 0, int trunc_gate = -1) {
  int rc = validate_schema(schema);
  if (rc) return rc;
  rc = validate_preds(schema, preds, np);
  if (rc) return rc;
  if (n && !ssts) return set_error(HG_ERR_INVALID, "null sst list");
  CU_TRY(cudaSetDevice(e->device));
  std::memset(&e->stats, 0, sizeof(e->stats));
  e->launches = 0;
  e->stage_cursor = 0;
  e->last_agg = hg_agg_device{};
  e->trunc_mask = trunc_mask;
  e->trunc_gate = trunc_gate;
  e->trunc_used = false;
  e->arena.reset();
  g_arena = &e->arena;
  CU_TRY(cudaEventRecord(e->ev0, e->stream));
  std::vector<size_t> pending, resident;
  for (size_t i = 0; i < n; i++) {
    if (e->ssts.count(ssts[i].id)) resident.push_back(i);
    else if (std::find_if(pending.begin(), pending.end(), [&](size_t j) { return ssts[j].id == ssts[i].id; }) == pending.end()) pending.push_back(i);
  }
  if (!pending.empty()) {
    for (uint32_t c = 0; c < schema->num_primary_keys; c++) need_cols.push_back(c);
    for (size_t i = 0; i < np; i++) need_cols.push_back(preds[i].column);
    rc = load_transient(e, schema, ssts, pending, preds, np, need_cols, seq_if_overlap && n > 1, resident);
    if (rc) { end_call(e); return rc; }
  }
  return HG_OK;
}

static int scan_impl(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n, const hg_predicate* preds,
                     size_t np, const uint32_t* projection, size_t nproj, int keep_builtin, struct ArrowArrayStream* out) {
  if (!e || !out) return set_error(HG_ERR_INVALID, "null argument");
  {
    int vrc = validate_schema(schema);     // before anything reads schema->num_columns
    if (vrc) return vrc;
  }
  std::lock_guard<std::mutex> g(e->mu);
  std::vector<uint32_t> touch;
  if (projection) for (size_t i = 0; i < nproj; i++) { if (projection[i] < schema->num_columns) touch.push_back(projection[i]); }
  else for (uint32_t c = 0; c < (keep_builtin ? schema->num_columns : schema->num_columns - 2); c++) touch.push_back(c);
  int rc = begin_call(e, schema, ssts, n, preds, np, touch, true);
  if (rc) return rc;
  CallGuard guard{e};
  cudaStream_t s = e->stream;
  Launch L = e->L();
  // output columns: all user columns (+ builtin when keep_builtin) or the projection (treated as a
  // post-merge column selection over the user columns)
  std::vector<uint32_t> out_cols;
  const uint32_t user_cols = schema->num_columns - 2;
  if (projection) {
    for (size_t i = 0; i < nproj; i++) {
      if (projection[i] >= schema->num_columns) return set_error(HG_ERR_INVALID, "projection index out of range");
      out_cols.push_back(projection[i]);
    }
  } else {
    for (uint32_t c = 0; c < (keep_builtin ? schema->num_columns : user_cols); c++) out_cols.push_back(c);
  }
  auto data = std::make_shared<StreamData>();
  std::string tmpname;
  for (uint32_t c : out_cols) {
    HostColumn hc;
    hc.name = col_name(schema, c, &tmpname);
    hc.type = schema->types[c];
    hc.width = type_width_host(hc.type);
    data->cols.push_back(hc);
  }
  data->batch_start.push_back(0);
  if (n == 0) { make_stream(out, data); return HG_OK; }   // EmptyRecordBatchStream (storage.rs:337-341)

  const bool append = schema->update_mode == HG_UPDATE_APPEND;
  PipelineState st;
  st.keep_order = append;
  rc = run_pipeline(e, schema, ssts, n, preds, np, out_cols, /*want_batches=*/true, &st);
  if (rc) return rc;
  const uint32_t N = st.N;
  uint32_t hc[8] = {0};
  CU_TRY(cudaMemcpyAsync(hc, st.d_counters.p, sizeof(hc), cudaMemcpyDeviceToHost, s));
  rc = check_device_error(e, &st);   // synchronises
  if (rc) return rc;
  const uint32_t M = hc[0], R = hc[1];
  // gather + export
  DevBuf d_null;
  CU_TRY(d_null.alloc(sizeof(unsigned long long) * out_cols.size() + 16, s));
  CU_TRY(cudaMemsetAsync(d_null.p, 0, sizeof(unsigned long long) * out_cols.size() + 16, s));
  std::vector<DevBuf> gv(out_cols.size()), gb(out_cols.size()), gm(out_cols.size());
  uint64_t d2h = 0;
  // which row stands for an output row in the fixed-width columns: the run's LAST row (LastValueOperator, operator.rs:39-44) or,
  // in Append mode, its FIRST row (BytesMergeOperator takes column.slice(0, 1), operator.rs:96-100)
  DevBuf first_rows;
  const uint32_t* rep_rows = st.out_rows.as<uint32_t>();
  if (append && R > 0) {
    CU_TRY(first_rows.alloc(size_t(R) * 4 + 16, s));
    k::first_rows(L, st.order_ptr, st.out_pos.as<uint32_t>(), st.d_r, R, first_rows.as<uint32_t>());
    rep_rows = first_rows.as<uint32_t>();
  }
  for (size_t i = 0; i < out_cols.size() && R > 0; i++) {
    DecodedCol& dc = st.cols[out_cols[i]];
    HostColumn& hcx = data->cols[i];
    if (dc.type == T_BINARY) {
      // Binary: Arrow offsets by an exclusive scan of the byte lengths, then one warp per value copies the bytes.
      //   Overwrite: one value per output row (its run's last row).   Append (BytesMergeOperator, operator.rs:75-95): the values of
      //   ALL rows of a run concatenated in merged order = every surviving row's bytes laid out in merged order, offsets taken at
      //   the runs' first rows; the result is never NULL.
      const uint32_t cnt_cap = append ? N : R;                     // elements scanned (device counts: M resp. R)
      const uint32_t* src_rows = append ? st.order_ptr : st.out_rows.as<uint32_t>();
      const uint32_t* d_cnt = append ? st.d_m : st.d_r;
      DevBuf lens_scan, offs_dev, d_total;
      CU_TRY(lens_scan.alloc((size_t(cnt_cap) + 1) * 4 + 16, s));
      CU_TRY(d_total.alloc(16, s));
      k::gather_lens(L, dc.view(), src_rows, d_cnt, cnt_cap + 1, lens_scan.as<uint32_t>());
      k::exclusive_scan_u32(L, lens_scan.as<uint32_t>(), cnt_cap + 1, d_total.as<uint32_t>());
      uint32_t total = 0;
      CU_TRY(cudaMemcpyAsync(&total, d_total.p, 4, cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaStreamSynchronize(s));
      if (total >= 0x7fffffffu) return set_error(HG_ERR_UNSUPPORTED, "Binary column larger than 2 GiB in one call (Arrow int32 offsets)");
      CU_TRY(gv[i].alloc(size_t(total) + 16, s));
      k::copy_var(L, dc.view(), src_rows, d_cnt, cnt_cap, lens_scan.as<uint32_t>(), gv[i].as<uint8_t>());
      const uint32_t* offs_src = lens_scan.as<uint32_t>();
      if (append) {
        CU_TRY(offs_dev.alloc((size_t(R) + 1) * 4 + 16, s));
        k::run_offsets(L, lens_scan.as<uint32_t>(), st.out_pos.as<uint32_t>(), st.d_r, st.d_m, R, offs_dev.as<uint32_t>());
        offs_src = offs_dev.as<uint32_t>();
      }
      hcx.vals = pinned_pool().alloc(size_t(total) + 16);
      hcx.offsets = static_cast<int32_t*>(pinned_pool().alloc((size_t(R) + 1) * 4 + 16));
      if (!hcx.vals || !hcx.offsets) return set_error(HG_ERR_OOM, "pinned host memory");
      if (total) CU_TRY(cudaMemcpyAsync(hcx.vals, gv[i].p, total, cudaMemcpyDeviceToHost, s));
      CU_TRY(cudaMemcpyAsync(hcx.offsets, offs_src, (size_t(R) + 1) * 4, cudaMemcpyDeviceToHost, s));
      d2h += size_t(total) + (size_t(R) + 1) * 4;
      // validity: Overwrite = the representative row's; Append = valid unless the run is ONE row whose value is NULL (a run with no
      // bytes returns its column unchanged, operator.rs:80-92; more than one such row is the reference's RecordBatch error)
      CU_TRY(gb[i].alloc(size_t(R) + 16, s));
      CU_TRY(gm[i].alloc((size_t(R) + 7) / 8 + 16, s));
      DevBuf scratch_ptrs;
      if (append) k::append_validity(L, dc.view(), st.order_ptr, st.out_pos.as<uint32_t>(), st.d_r, offs_src, R, gb[i].as<uint8_t>(), st.d_err.as<int>());
      else {
        CU_TRY(scratch_ptrs.alloc(size_t(R) * 8 + 16, s));
        k::gather_column(L, dc.view(), st.out_rows.as<uint32_t>(), st.d_r, R, scratch_ptrs.p, gb[i].as<uint8_t>());
      }
      k::pack_validity(L, gb[i].as<uint8_t>(), R, gm[i].as<uint8_t>(), d_null.as<unsigned long long>() + i);
      hcx.bitmap = static_cast<uint8_t*>(pinned_pool().alloc((size_t(R) + 7) / 8 + 16));
      if (!hcx.bitmap) return set_error(HG_ERR_OOM, "pinned host memory");
      CU_TRY(cudaMemcpyAsync(hcx.bitmap, gm[i].p, (size_t(R) + 7) / 8, cudaMemcpyDeviceToHost, s));
      d2h += (size_t(R) + 7) / 8;
      CU_TRY(cudaStreamSynchronize(s));                              // the temporaries above are popped off the arena on scope exit
      continue;
    }
    CU_TRY(gv[i].alloc(size_t(R) * dc.width + 16, s));
    bool nulls = dc.valid.p != nullptr;
    if (nulls) { CU_TRY(gb[i].alloc(size_t(R) + 16, s)); CU_TRY(gm[i].alloc((size_t(R) + 7) / 8 + 16, s)); }
    k::gather_column(L, dc.view(), rep_rows, st.d_r, R, gv[i].p, gb[i].as<uint8_t>());
    hcx.vals = pinned_pool().alloc(size_t(R) * dc.width + 16);
    if (!hcx.vals) return set_error(HG_ERR_OOM, "pinned host memory");
    CU_TRY(cudaMemcpyAsync(hcx.vals, gv[i].p, size_t(R) * dc.width, cudaMemcpyDeviceToHost, s));
    d2h += size_t(R) * dc.width;
    if (nulls) {
      k::pack_validity(L, gb[i].as<uint8_t>(), R, gm[i].as<uint8_t>(), d_null.as<unsigned long long>() + i);
      hcx.bitmap = static_cast<uint8_t*>(pinned_pool().alloc((size_t(R) + 7) / 8 + 16));
      if (!hcx.bitmap) return set_error(HG_ERR_OOM, "pinned host memory");
      CU_TRY(cudaMemcpyAsync(hcx.bitmap, gm[i].p, (size_t(R) + 7) / 8, cudaMemcpyDeviceToHost, s));
      d2h += (size_t(R) + 7) / 8;
    }
  }
  std::vector<uint32_t> bound(st.nchunks);
  if (st.nchunks && N > 0) CU_TRY(cudaMemcpyAsync(bound.data(), st.bound.p, st.nchunks * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaEventRecord(e->ev1, s));
  if (append) { rc = check_device_error(e, &st); if (rc) return rc; }       // BytesMergeOperator's own failure mode (see append_validity)
  CU_TRY(cudaStreamSynchronize(s));
  // MergeStream batch boundaries (read.rs:289-343, 349-384): batch c = outputs [bound[c-1], bound[c]); final flush = the rest
  uint32_t prev = 0;
  for (uint32_t c = 0; c < st.nchunks; c++) {
    uint32_t b = std::min(bound[c], R);
    if (b > prev) { data->batch_start.push_back(b); prev = b; }
  }
  if (R > prev) data->batch_start.push_back(R);
  float ms = 0;
  cudaEventElapsedTime(&ms, e->ev0, e->ev1);
  e->stats.rows_in_files = st.plan.rows_in_files;
  e->stats.rows_decoded = st.plan.rows_decoded;
  e->stats.rows_materialized = st.plan.rows_decoded;
  e->stats.rows_filtered = M;
  e->stats.rows_out = R;
  e->stats.bytes_d2h = d2h;
  e->stats.kernel_launches = e->launches;
  e->stats.gpu_ms = ms;
  e->stats.path = 0;
  if (N > 0) { float kms = 0; cudaEventElapsedTime(&kms, e->evk0, e->evk1); e->stats.kernel_ms = kms; cudaEventElapsedTime(&kms, e->evm0, e->evm1); e->stats.merge_ms = kms; }
  make_stream(out, data);
  return HG_OK;
}

int hg_scan_open(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n_ssts, const hg_predicate* preds,
                 size_t n_preds, const uint32_t* projection, size_t n_projection, int keep_builtin, struct ArrowArrayStream* out) {
  HG_GUARD_BEGIN
  return scan_impl(e, schema, ssts, n_ssts, preds, n_preds, projection, n_projection, keep_builtin, out);
  HG_GUARD_END
}

int hg_plan_pk_splitters(const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n, uint32_t parts, uint64_t* splitters) {
  HG_GUARD_BEGIN
  if (!ssts || !splitters || parts == 0) return set_error(HG_ERR_INVALID, "null argument");
  int rc = validate_schema(schema);
  if (rc) return rc;
  const uint32_t t0 = schema->types[0];
  struct Iv { uint64_t mn, mx, rows; };
  std::vector<Iv> ivs;
  uint64_t total = 0;
  for (size_t i = 0; i < n; i++) {
    std::vector<uint8_t> filebuf;
    const uint8_t* data = ssts[i].data;
    uint64_t size = ssts[i].size;
    if (!data) {
      if (!ssts[i].path) return set_error(HG_ERR_NOT_FOUND, "hg_plan_pk_splitters needs the SST bytes or a path");
      rc = read_whole_file(ssts[i].path, &filebuf);
      if (rc) return rc;
      data = filebuf.data();
      size = filebuf.size();
    }
    SstResident r;
    std::vector<PageDev> pages;
    std::vector<ChunkDev> chunks;
    std::string err;
    rc = prepare_sst(schema, ssts[i].id, data, size, &r, &pages, &chunks, &err);
    if (rc) return set_error(rc, err);
    const size_t ncols = size_t(r.meta.ncols);
    for (size_t g = 0; g < r.rg_rows.size(); g++) {
      if (!r.rg_rows[g]) continue;
      const RgCol& c0 = r.rgcol[g * ncols];
      if (!c0.has_minmax) return set_error(HG_ERR_UNSUPPORTED, "pk0 statistics missing: cannot range-partition");
      ivs.push_back(Iv{c0.mn, c0.mx, r.rg_rows[g]});
      total += r.rg_rows[g];
    }
  }
  // rows of a row group are taken as evenly spread over its [min, max] of pk0 (SSTs are PK-sorted, so they nearly are);
  // splitter q = the smallest pk0 value below which at least q / parts of all rows lie under that model (bisection in the
  // order-preserving unsigned image of the column).  A balance heuristic: any splitters give a correct partition.
  const uint64_t flip = type_is_signed(t0) ? (1ull << 63) : 0ull;
  uint64_t lo_all = ~0ull, hi_all = 0;
  for (Iv& iv : ivs) { iv.mn ^= flip; iv.mx ^= flip; lo_all = std::min(lo_all, iv.mn); hi_all = std::max(hi_all, iv.mx); }
  auto below = [&](uint64_t x) {       // modelled number of rows with pk0 < x
    long double acc = 0;
    for (const Iv& iv : ivs) {
      if (x <= iv.mn) continue;
      if (x > iv.mx) { acc += iv.rows; continue; }
      acc += (long double)iv.rows * ((long double)(x - iv.mn) / ((long double)(iv.mx - iv.mn) + 1.0L));
    }
    return acc;
  };
  for (uint32_t q = 1; q < parts; q++) {
    uint64_t sp = hi_all;
    if (!ivs.empty()) {
      const long double want = (long double)total * q / parts;
      uint64_t a = lo_all, b = hi_all;
      while (a < b) { const uint64_t mid = a + (b - a) / 2; if (below(mid) >= want) b = mid; else a = mid + 1; }
      sp = a;
    } else sp = 0;
    splitters[q - 1] = sp ^ flip;
  }
  return HG_OK;
  HG_GUARD_END
}

int hg_compact_to_sst(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n, const hg_predicate* shard_preds,
                      size_t n_shard_preds, const hg_write_props* props, const char* out_path, hg_file_meta* out) {
  HG_GUARD_BEGIN
  if (!e || !props || !out_path || !out || (n_shard_preds && !shard_preds)) return set_error(HG_ERR_INVALID, "null argument");
  for (size_t i = 0; i < n_shard_preds; i++)
    if (shard_preds[i].column != 0) return set_error(HG_ERR_INVALID, "compaction shards are ranges of the first primary-key column");
  if (schema && schema->types)
    for (uint32_t c = 0; c < schema->num_columns; c++)
      if (schema->types[c] == T_BINARY) return set_error(HG_ERR_UNSUPPORTED, "GPU SST writer: Binary columns are not implemented (use hg_compact_open + the host writer)");
  {
    int vrc = validate_schema(schema);
    if (vrc) return vrc;
  }
  std::lock_guard<std::mutex> g(e->mu);
  std::vector<uint32_t> touch;
  for (uint32_t c = 0; c < schema->num_columns; c++) touch.push_back(c);
  int rc = begin_call(e, schema, ssts, n, shard_preds, n_shard_preds, touch, true);
  if (rc) return rc;
  CallGuard guard{e};
  cudaStream_t s = e->stream;
  Launch L = e->L();
  std::memset(out, 0, sizeof(*out));
  for (size_t i = 0; i < n; i++) {
    if (i == 0 || ssts[i].time_start < out->time_start) out->time_start = ssts[i].time_start;
    if (i == 0 || ssts[i].time_end > out->time_end) out->time_end = ssts[i].time_end;
    out->max_sequence = std::max(out->max_sequence, ssts[i].max_sequence);
  }
  PipelineState st;
  uint32_t R = 0;
  std::vector<DevBuf> gv(schema->num_columns), gb(schema->num_columns);
  std::vector<writer::ColIn> cols(schema->num_columns);
  if (n > 0) {
    rc = run_pipeline(e, schema, ssts, n, shard_preds, n_shard_preds, touch, /*want_batches=*/false, &st);
    if (rc) return rc;
    uint32_t hc[8] = {0};
    CU_TRY(cudaMemcpyAsync(hc, st.d_counters.p, sizeof(hc), cudaMemcpyDeviceToHost, s));
    rc = check_device_error(e, &st);
    if (rc) return rc;
    R = hc[1];
    e->stats.rows_in_files = st.plan.rows_in_files;
    e->stats.rows_decoded = st.plan.rows_decoded;
    e->stats.rows_materialized = st.plan.rows_decoded;
    e->stats.rows_filtered = hc[0];
    e->stats.rows_out = R;
  }
  for (uint32_t c = 0; c < schema->num_columns; c++) {
    const uint32_t width = type_width_host(schema->types[c]);
    cols[c] = writer::ColIn{nullptr, nullptr, schema->types[c], width};
    if (R == 0) continue;
    DecodedCol& dc = st.cols[c];
    CU_TRY(gv[c].alloc(size_t(R) * width + 16, s));
    const bool nulls = dc.valid.p != nullptr;
    if (nulls) CU_TRY(gb[c].alloc(size_t(R) + 16, s));
    k::gather_column(L, dc.view(), st.out_rows.as<uint32_t>(), st.d_r, R, gv[c].p, gb[c].as<uint8_t>());
    cols[c].vals = gv[c].p;
    cols[c].valid = nulls ? gb[c].as<uint8_t>() : nullptr;
  }
  uint8_t* host = nullptr;
  uint64_t size = 0;
  rc = writer::write_sst(e, schema, cols.data(), schema->num_columns, R, props, &host, &size);
  if (rc) return rc;
  CU_TRY(cudaEventRecord(e->ev1, s));
  CU_TRY(cudaStreamSynchronize(s));
  FILE* f = std::fopen(out_path, "wb");
  bool ok = f != nullptr;
  if (ok) ok = std::fwrite(host, 1, size_t(size), f) == size_t(size);
  if (f) ok = std::fclose(f) == 0 && ok;
  cudaFreeHost(host);
  if (!ok) return set_error(HG_ERR_NOT_FOUND, std::string("cannot write ") + out_path);
  out->size = size;
  out->num_rows = R;
  float ms = 0;
  cudaEventElapsedTime(&ms, e->ev0, e->ev1);
  e->stats.gpu_ms = ms;
  e->stats.kernel_launches = e->launches;
  if (st.N > 0) { float kms = 0; cudaEventElapsedTime(&kms, e->evm0, e->evm1); e->stats.merge_ms = kms; }
  return HG_OK;
  HG_GUARD_END
}

int hg_write_batch(hg_engine* e, const hg_schema_desc* schema, const struct ArrowArray* batch, uint64_t sequence, const hg_write_props* props,
                   const char* out_path, hg_file_meta* out) {
  HG_GUARD_BEGIN
  if (!e || !batch || !props || !out_path || !out) return set_error(HG_ERR_INVALID, "null argument");
  int rc = validate_schema(schema);
  if (rc) return rc;
  const uint32_t ncols = schema->num_columns, user = ncols - 2, npk = schema->num_primary_keys;
  for (uint32_t c = 0; c < ncols; c++)
    if (schema->types[c] == T_BINARY) return set_error(HG_ERR_UNSUPPORTED, "GPU SST writer: Binary columns are not implemented");
  if (batch->n_children != int64_t(user)) return set_error(HG_ERR_INVALID, "batch must hold the user columns of the schema");
  if (batch->length < 0 || batch->length >= 0xfffffff0ll) return set_error(HG_ERR_UNSUPPORTED, "batch larger than 2^32 rows");
  if (batch->null_count > 0) return set_error(HG_ERR_UNSUPPORTED, "NULL rows (struct-level validity) are not supported");
  const uint32_t n = uint32_t(batch->length);
  std::lock_guard<std::mutex> g(e->mu);
  CU_TRY(cudaSetDevice(e->device));
  std::memset(&e->stats, 0, sizeof(e->stats));
  e->launches = 0;
  e->stage_cursor = 0;
  e->arena.reset();
  g_arena = &e->arena;
  cudaStream_t s = e->stream;
  Launch L = e->L();
  CU_TRY(cudaEventRecord(e->ev0, s));
  // ---- upload the user columns (values + validity expanded to one byte per row)
  std::vector<DevBuf> vals(ncols), valid(ncols), sorted(ncols), svalid(ncols);
  std::vector<ColView> views(ncols);
  uint64_t h2d = 0;
  for (uint32_t c = 0; c < user; c++) {
    const struct ArrowArray* col = batch->children[c];
    if (!col || col->n_buffers < 2 || col->length != batch->length) return set_error(HG_ERR_INVALID, "column " + std::to_string(c) + ": not a primitive array of the batch's length");
    const uint32_t w = type_width_host(schema->types[c]);
    CU_TRY(vals[c].alloc(size_t(n) * w + 16, s));
    if (n) {
      if (!col->buffers[1]) return set_error(HG_ERR_INVALID, "column without a data buffer");
      CU_TRY(cudaMemcpyAsync(vals[c].p, static_cast<const uint8_t*>(col->buffers[1]) + size_t(col->offset) * w, size_t(n) * w, cudaMemcpyHostToDevice, s));
      h2d += size_t(n) * w;
    }
    const bool has_nulls = col->null_count != 0 && col->buffers[0] != nullptr;
    if (has_nulls) {
      if (c < npk) return set_error(HG_ERR_UNSUPPORTED, "NULL primary keys are not supported on the GPU path");
      const size_t nbytes = size_t((col->offset + col->length + 7) / 8);
      DevBuf bm;
      CU_TRY(bm.alloc(nbytes + 16, s));
      CU_TRY(cudaMemcpyAsync(bm.p, col->buffers[0], nbytes, cudaMemcpyHostToDevice, s));
      CU_TRY(valid[c].alloc(size_t(n) + 16, s));
      k::unpack_bitmap(L, bm.as<uint8_t>(), uint64_t(col->offset), n, valid[c].as<uint8_t>());
      bm.release();                          // arena memory: stays valid until the next call
      h2d += nbytes;
    }
    views[c] = ColView{vals[c].p, has_nulls ? valid[c].as<uint8_t>() : nullptr, schema->types[c], w, nullptr};
  }
  // ---- sort by (pk0, .., pkN-1): LSD over the key columns, last key first; every pass is a stable radix sort
  DevBuf perm, perm2, keys, keys2, counts, d_n;
  CU_TRY(perm.alloc(size_t(n) * 4 + 16, s));
  CU_TRY(perm2.alloc(size_t(n) * 4 + 16, s));
  CU_TRY(keys.alloc(size_t(n) * 8 + 16, s));
  CU_TRY(keys2.alloc(size_t(n) * 8 + 16, s));
  CU_TRY(counts.alloc(k::radix_tmp_elems(n) * 4, s));
  CU_TRY(d_n.alloc(16, s));
  k::fill_u32(L, d_n.as<uint32_t>(), n, 4);
  k::iota_u32(L, perm.as<uint32_t>(), n);
  uint32_t* pcur = perm.as<uint32_t>();
  uint32_t* palt = perm2.as<uint32_t>();
  for (int c = int(npk) - 1; c >= 0 && n > 1; c--) {
    k::column_sort_keys(L, views[c], pcur, n, keys.as<uint64_t>());
    const uint32_t t = schema->types[c];
    const int bits = (type_is_signed(t) || type_is_float(t)) ? 64 : 8 * int(type_width_host(t));
    if (k::radix_sort_pairs(L, keys.as<uint64_t>(), pcur, keys2.as<uint64_t>(), palt, d_n.as<uint32_t>(), n, bits, counts.as<uint32_t>())) std::swap(pcur, palt);
  }
  // ---- gather into sorted columns, append the builtin columns
  std::vector<writer::ColIn> cols(ncols);
  for (uint32_t c = 0; c < user; c++) {
    const uint32_t w = views[c].width;
    CU_TRY(sorted[c].alloc(size_t(n) * w + 16, s));
    if (views[c].valid) CU_TRY(svalid[c].alloc(size_t(n) + 16, s));
    k::gather_column(L, views[c], pcur, d_n.as<uint32_t>(), n, sorted[c].p, svalid[c].as<uint8_t>());
    cols[c] = writer::ColIn{sorted[c].p, views[c].valid ? svalid[c].as<uint8_t>() : nullptr, schema->types[c], w};
  }
  CU_TRY(sorted[user].alloc(size_t(n) * 8 + 16, s));
  k::fill_u64(L, sorted[user].as<uint64_t>(), sequence, n);
  cols[user] = writer::ColIn{sorted[user].p, nullptr, T_U64, 8};
  CU_TRY(sorted[user + 1].alloc(size_t(n) * 8 + 16, s));
  CU_TRY(svalid[user + 1].alloc(size_t(n) + 16, s));
  CU_TRY(cudaMemsetAsync(sorted[user + 1].p, 0, size_t(n) * 8 + 16, s));
  CU_TRY(cudaMemsetAsync(svalid[user + 1].p, 0, size_t(n) + 16, s));
  cols[user + 1] = writer::ColIn{sorted[user + 1].p, svalid[user + 1].as<uint8_t>(), T_U64, 8};
  uint8_t* host = nullptr;
  uint64_t size = 0;
  rc = writer::write_sst(e, schema, cols.data(), ncols, n, props, &host, &size);
  if (rc) return rc;
  CU_TRY(cudaEventRecord(e->ev1, s));
  CU_TRY(cudaStreamSynchronize(s));
  FILE* f = std::fopen(out_path, "wb");
  bool ok = f != nullptr;
  if (ok) ok = std::fwrite(host, 1, size_t(size), f) == size_t(size);
  if (f) ok = std::fclose(f) == 0 && ok;
  cudaFreeHost(host);
  if (!ok) return set_error(HG_ERR_NOT_FOUND, std::string("cannot write ") + out_path);
  std::memset(out, 0, sizeof(*out));
  out->size = size;
  out->num_rows = n;
  out->max_sequence = sequence;
  float ms = 0;
  cudaEventElapsedTime(&ms, e->ev0, e->ev1);
  e->stats.gpu_ms = ms;
  e->stats.bytes_h2d = h2d;
  e->stats.rows_out = n;
  e->stats.kernel_launches = e->launches;
  return HG_OK;
  HG_GUARD_END
}

int hg_compact_open(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n_ssts, struct ArrowArrayStream* out) {
  HG_GUARD_BEGIN
  // Executor::do_compaction builds the same plan with no predicate and keep_builtin = true (executor.rs:164-169)
  return scan_impl(e, schema, ssts, n_ssts, nullptr, 0, nullptr, 0, 1, out);
  HG_GUARD_END
}


static int aggregate_core(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n, const hg_predicate* preds,
                          size_t np, const hg_agg_spec* agg, AggBuffers* ab) {
  cudaStream_t s = e->stream;
  Launch L = e->L();
  if (!agg) return set_error(HG_ERR_INVALID, "null aggregation spec");
  auto col_ok = [&](int32_t c) { return c < 0 || uint32_t(c) < schema->num_columns; };
  if (!col_ok(agg->group_col) || !col_ok(agg->ts_col) || !col_ok(agg->value_col)) return set_error(HG_ERR_INVALID, "aggregation column out of range");
  const bool has_ts = agg->ts_col >= 0 && agg->window_ms > 0;
  if (has_ts && type_is_float(schema->types[agg->ts_col])) return set_error(HG_ERR_INVALID, "time column must be an integer column");
  if (agg->group_col >= 0) { ab->gtype = schema->types[agg->group_col]; ab->gwidth = type_width_host(ab->gtype); }
  if (n == 0) { ab->G = 0; return HG_OK; }

  if (agg->mode > HG_AGG_HASH) return set_error(HG_ERR_INVALID, "aggregation mode");
  if (schema->update_mode != HG_UPDATE_OVERWRITE) return set_error(HG_ERR_UNSUPPORTED, "aggregation over an Append-mode (BytesMergeOperator) table");
  for (int32_t c : {agg->group_col, agg->ts_col, agg->value_col})
    if (c >= 0 && schema->types[c] == T_BINARY) return set_error(HG_ERR_INVALID, "Binary columns cannot be grouped or aggregated");
  // HASH mode only differs from RUNS when the key is not a prefix of the sort order (pk0 [, bucket of pk1]) / not global
  const bool prefix_key = (agg->group_col < 0 && !has_ts) || (agg->group_col == 0 && (!has_ts || agg->ts_col == 1));
  const bool hash_sort = agg->mode == HG_AGG_HASH && !prefix_key;
  // fused fast path: sorted PK-disjoint inputs, one PLAIN page per chunk, group = pk0, time = pk1
  if (!(e->flags & HG_FLAG_NO_FUSED) && !hash_sort) {
    int frc = fused::try_scan_aggregate(e, schema, ssts, n, preds, np, agg, ab);
    if (frc != fused::NOT_APPLICABLE) return frc;
  }

  std::vector<uint32_t> need;
  if (agg->group_col >= 0) need.push_back(uint32_t(agg->group_col));
  if (has_ts) need.push_back(uint32_t(agg->ts_col));
  if (agg->value_col >= 0) need.push_back(uint32_t(agg->value_col));
  PipelineState st;
  int rc = run_pipeline(e, schema, ssts, n, preds, np, need, /*want_batches=*/false, &st);
  if (rc) return rc;
  const uint32_t N = st.N;
  AggSpecDev spec;
  std::memset(&spec, 0, sizeof(spec));
  spec.has_group = agg->group_col >= 0;
  spec.has_ts = has_ts;
  spec.has_value = agg->value_col >= 0;
  spec.window_ms = has_ts ? agg->window_ms : 1;
  if (spec.has_group) spec.group = st.cols[agg->group_col].view();
  if (spec.has_ts) spec.ts = st.cols[agg->ts_col].view();
  if (spec.has_value) spec.value = st.cols[agg->value_col].view();
  DevBuf head, seg, gk, gk2, vals, vals2, rcounts;
  const uint32_t* agg_rows = st.out_rows.as<uint32_t>();
  if (hash_sort && N > 0) {
    // radix partition: stable sort of the surviving rows by (group value, bucket) — bucket first, then the group value
    CU_TRY(gk.alloc(size_t(N) * 8 + 16, s));
    CU_TRY(gk2.alloc(size_t(N) * 8 + 16, s));
    CU_TRY(vals.alloc(size_t(N) * 4 + 16, s));
    CU_TRY(vals2.alloc(size_t(N) * 4 + 16, s));
    CU_TRY(rcounts.alloc(k::radix_tmp_elems(N) * sizeof(uint32_t), s));
    const uint32_t* cur = st.out_rows.as<uint32_t>();
    if (spec.has_ts) {
      k::group_sort_keys(L, spec, cur, st.d_r, N, nullptr, gk.as<uint64_t>(), vals.as<uint32_t>());
      int w = k::radix_sort_pairs(L, gk.as<uint64_t>(), vals.as<uint32_t>(), gk2.as<uint64_t>(), vals2.as<uint32_t>(), st.d_r, N, 64, rcounts.as<uint32_t>());
      if (w) std::swap(vals, vals2);
      cur = vals.as<uint32_t>();
    }
    if (spec.has_group) {
      // (vals2 receives the row ids again: the keys are recomputed in the order the first sort produced)
      k::group_sort_keys(L, spec, cur, st.d_r, N, gk.as<uint64_t>(), nullptr, vals2.as<uint32_t>());
      int w = k::radix_sort_pairs(L, gk.as<uint64_t>(), vals2.as<uint32_t>(), gk2.as<uint64_t>(), vals.as<uint32_t>(), st.d_r, N,
                                  // unsigned values occupy their native width; signed / float keys are 64-bit images
                                  (type_is_signed(schema->types[agg->group_col]) || type_is_float(schema->types[agg->group_col]))
                                      ? 64 : 8 * int(type_width_host(schema->types[agg->group_col])),
                                  rcounts.as<uint32_t>());
      if (!w) std::swap(vals, vals2);
      cur = vals.as<uint32_t>();
    }
    agg_rows = cur;
    gk.reset();
  }
  CU_TRY(head.alloc(size_t(N) + 16, s));
  CU_TRY(seg.alloc(size_t(N) * 4 + 16, s));
  uint32_t hc[8] = {0};
  if (N > 0) {
    k::group_flags(L, spec, agg_rows, st.d_r, N, head.as<uint8_t>());
    k::clear_tail(L, head.as<uint8_t>(), st.d_r, N);
    k::compact_flags(L, head.as<uint8_t>(), N, st.tmp.as<uint32_t>(), seg.as<uint32_t>(), st.counters() + 2);
  }
  CU_TRY(cudaMemcpyAsync(hc, st.d_counters.p, sizeof(hc), cudaMemcpyDeviceToHost, s));
  rc = check_device_error(e, &st);
  if (rc) return rc;
  const uint32_t G = hc[2];
  ab->G = G;
  CU_TRY(ab->gkey.alloc(size_t(G) * 8 + 16, s));
  CU_TRY(ab->bucket.alloc(size_t(G) * 8 + 16, s));
  CU_TRY(ab->count.alloc(size_t(G) * 8 + 16, s));
  CU_TRY(ab->sum.alloc(size_t(G) * 8 + 16, s));
  CU_TRY(ab->mn.alloc(size_t(G) * 8 + 16, s));
  CU_TRY(ab->mx.alloc(size_t(G) * 8 + 16, s));
  AggOut ao{ab->gkey.p, ab->bucket.as<int64_t>(), ab->count.as<uint64_t>(), ab->sum.as<double>(), ab->mn.as<double>(), ab->mx.as<double>()};
  if (G > 0) k::reduce_groups(L, spec, agg_rows, st.d_r, seg.as<uint32_t>(), st.d_g, G, ao);
  e->stats.rows_in_files = st.plan.rows_in_files;
  e->stats.rows_decoded = st.plan.rows_decoded;
  e->stats.rows_materialized = st.plan.rows_decoded;
  e->stats.rows_filtered = hc[0];
  e->stats.rows_out = hc[1];
  e->stats.groups_out = G;
  e->stats.path = 0;
  if (N > 0) { float kms = 0; cudaEventElapsedTime(&kms, e->evk0, e->evk1); e->stats.kernel_ms = kms; }
  return HG_OK;
}


// Which columns may a transient load ship as compressed prefixes for this aggregate?  Only when the call has the fused scan's shape
// with late materialisation and no time buckets: that kernel reads every column but pk0 only up to the row group's last row passing
// the gate column (fused_scan.cu: gate_sel_kernel, SnappyJob::partial).  *gate = the column that will be its gate.
static uint32_t aggregate_trunc_mask(const hg_engine* e, const hg_schema_desc* schema, const hg_predicate* preds, size_t np, const hg_agg_spec* agg, int* gate) {
  *gate = -1;
  if (!schema || !agg || np == 0 || schema->num_primary_keys < 2) return 0;
  if (e->flags & (HG_FLAG_NO_FUSED | HG_FLAG_NO_LATE_MATERIALIZATION | HG_FLAG_NO_PRUNING)) return 0;
  if (agg->ts_col >= 0 && agg->window_ms > 0) return 0;
  if (!(agg->group_col == 0 || (agg->group_col < 0 && agg->value_col < 0))) return 0;
  int extra = -1;
  bool on_pk1 = false;
  for (size_t i = 0; i < np; i++) {
    const uint32_t c = preds[i].column;
    if (c >= schema->num_columns || c >= 32) return 0;
    if (type_is_float(schema->types[c]) || preds[i].op == HG_OP_NE || preds[i].op == HG_OP_IN) return 0;
    if (c == 1) on_pk1 = true;
    if (c >= 2) { if (extra >= 0 && extra != int(c)) return 0; extra = int(c); }
  }
  *gate = extra >= 0 ? extra : (on_pk1 ? 1 : -1);
  if (*gate < 0) return 0;
  return ~1u;                                        // everything but pk0
}

static int aggregate_device_once(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n_ssts, const hg_predicate* preds,
                                 size_t n_preds, const hg_agg_spec* agg, uint32_t trunc_mask, int trunc_gate, hg_agg_device* out) {
  std::vector<uint32_t> touch;
  if (agg) for (int32_t c : {agg->group_col, agg->ts_col, agg->value_col}) if (c >= 0 && uint32_t(c) < (schema ? schema->num_columns : 0)) touch.push_back(uint32_t(c));
  int rc = begin_call(e, schema, ssts, n_ssts, preds, n_preds, touch, true, trunc_mask, trunc_gate);
  if (rc) return rc;
  CallGuard guard{e};
  AggBuffers ab;
  rc = aggregate_core(e, schema, ssts, n_ssts, preds, n_preds, agg, &ab);
  if (rc) return rc;
  CU_TRY(cudaEventRecord(e->ev1, e->stream));
  CU_TRY(cudaStreamSynchronize(e->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, e->ev0, e->ev1);
  e->stats.gpu_ms = ms;
  e->stats.kernel_launches = e->launches;
  out->num_groups = ab.G;
  out->d_gkey = ab.gkey.p;
  out->d_bucket = ab.bucket.as<int64_t>();
  out->d_count = ab.count.as<uint64_t>();
  out->d_sum = ab.sum.as<double>();
  out->d_min = ab.mn.as<double>();
  out->d_max = ab.mx.as<double>();
  e->last_agg = *out;
  e->last_gwidth = ab.gwidth;
  e->last_gtype = ab.gtype;
  for (DevBuf* b : {&ab.gkey, &ab.bucket, &ab.count, &ab.sum, &ab.mn, &ab.mx}) b->release();   // arena memory: valid until the next call
  return HG_OK;
}

int hg_scan_aggregate_device(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n_ssts,
                             const hg_predicate* preds, size_t n_preds, const hg_agg_spec* agg, hg_agg_device* out) {
  HG_GUARD_BEGIN
  if (!e || !out) return set_error(HG_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> g(e->mu);
  int gate = -1;
  const uint32_t mask = aggregate_trunc_mask(e, schema, preds, n_preds, agg, &gate);
  int rc = aggregate_device_once(e, schema, ssts, n_ssts, preds, n_preds, agg, mask, gate, out);
  if (rc && e->trunc_used) {                          // a compressed prefix ran out (lopsided page): repeat with whole pages
    static const bool trace = getenv("HORAE_TRACE") != nullptr;
    if (trace) fprintf(stderr, "[transient] a compressed prefix ended before the last needed row: repeating the call with whole pages\n");
    const uint64_t wasted = e->stats.bytes_h2d;
    rc = aggregate_device_once(e, schema, ssts, n_ssts, preds, n_preds, agg, 0, -1, out);
    e->stats.bytes_h2d += wasted;
    e->stats.path |= 2u;
  }
  return rc;
  HG_GUARD_END
}

static int aggregate_host_once(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n_ssts, const hg_predicate* preds,
                               size_t n_preds, const hg_agg_spec* agg, uint32_t trunc_mask, int trunc_gate, struct ArrowArrayStream* out) {
  std::vector<uint32_t> touch;
  if (agg) for (int32_t c : {agg->group_col, agg->ts_col, agg->value_col}) if (c >= 0 && uint32_t(c) < (schema ? schema->num_columns : 0)) touch.push_back(uint32_t(c));
  int rc = begin_call(e, schema, ssts, n_ssts, preds, n_preds, touch, true, trunc_mask, trunc_gate);
  if (rc) return rc;
  CallGuard guard{e};
  cudaStream_t s = e->stream;
  AggBuffers ab;
  rc = aggregate_core(e, schema, ssts, n_ssts, preds, n_preds, agg, &ab);
  if (rc) return rc;
  const uint32_t G = ab.G;
  auto data = std::make_shared<StreamData>();
  std::string tmp;
  struct Src { const char* name; uint32_t type; void* dev; uint32_t width; };
  std::vector<Src> srcs;
  if (agg->group_col >= 0) srcs.push_back({col_name(schema, uint32_t(agg->group_col), &tmp), ab.gtype, ab.gkey.p, ab.gwidth});
  if (agg->ts_col >= 0 && agg->window_ms > 0) srcs.push_back({"bucket", T_I64, ab.bucket.p, 8});
  srcs.push_back({"count", T_U64, ab.count.p, 8});
  if (agg->value_col >= 0) {
    srcs.push_back({"sum", T_F64, ab.sum.p, 8});
    srcs.push_back({"min", T_F64, ab.mn.p, 8});
    srcs.push_back({"max", T_F64, ab.mx.p, 8});
  }
  uint64_t d2h = 0;
  for (auto& sc : srcs) {
    HostColumn hc;
    hc.name = sc.name;
    hc.type = sc.type;
    hc.width = sc.width;
    data->cols.push_back(hc);                 // owned by the stream from here on: an early return releases the pinned buffer
    if (G) {
      HostColumn& col = data->cols.back();
      col.vals = pinned_pool().alloc(size_t(G) * sc.width + 16);
      if (!col.vals) return set_error(HG_ERR_OOM, "pinned host memory");
      CU_TRY(cudaMemcpyAsync(col.vals, sc.dev, size_t(G) * sc.width, cudaMemcpyDeviceToHost, s));
      d2h += size_t(G) * sc.width;
    }
  }
  CU_TRY(cudaEventRecord(e->ev1, s));
  CU_TRY(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, e->ev0, e->ev1);
  e->stats.gpu_ms = ms;
  e->stats.bytes_d2h = d2h;
  e->stats.kernel_launches = e->launches;
  data->batch_start.push_back(0);
  if (G) data->batch_start.push_back(G);
  make_stream(out, data);
  return HG_OK;
}

int hg_scan_aggregate(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n_ssts, const hg_predicate* preds,
                      size_t n_preds, const hg_agg_spec* agg, struct ArrowArrayStream* out) {
  HG_GUARD_BEGIN
  if (!e || !out) return set_error(HG_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> g(e->mu);
  int gate = -1;
  const uint32_t mask = aggregate_trunc_mask(e, schema, preds, n_preds, agg, &gate);
  int rc = aggregate_host_once(e, schema, ssts, n_ssts, preds, n_preds, agg, mask, gate, out);
  if (rc && e->trunc_used) {                          // a compressed prefix ran out (lopsided page): repeat with whole pages
    static const bool trace = getenv("HORAE_TRACE") != nullptr;
    if (trace) fprintf(stderr, "[transient] a compressed prefix ended before the last needed row: repeating the call with whole pages\n");
    const uint64_t wasted = e->stats.bytes_h2d;
    rc = aggregate_host_once(e, schema, ssts, n_ssts, preds, n_preds, agg, 0, -1, out);
    e->stats.bytes_h2d += wasted;
    e->stats.path |= 2u;
  }
  return rc;
  HG_GUARD_END
}

}  // extern "C"
