// engine_internal.h — internals shared by engine.cu and fused_scan.cu (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/horae_gpu.h"
#include "device_types.h"
#include "kernels.h"
#include "parquet_meta.hpp"

using namespace horae;

// --------------------------------------------------------------------------------------------------- error plumbing
int set_error(int code, const std::string& msg);   // defined in engine.cu (thread-local message)
#define CU_TRY(expr)                                                                                         \
  do {                                                                                                       \
    cudaError_t _e = (expr);                                                                                 \
    if (_e != cudaSuccess)                                                                                   \
      return set_error(HG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
  } while (0)

inline uint32_t type_width_host(uint32_t t) {
  switch (t) {
    case T_U8: case T_I8: return 1;
    case T_U16: case T_I16: return 2;
    case T_U32: case T_I32: case T_F32: return 4;
    default: return 8;
  }
}
inline bool type_is_signed(uint32_t t) { return t == T_I8 || t == T_I16 || t == T_I32 || t == T_I64; }
inline bool type_is_float(uint32_t t) { return t == T_F32 || t == T_F64; }
inline int expected_phys(uint32_t t) {
  switch (t) {
    case T_U64: case T_I64: return PT_INT64;
    case T_F32: return PT_FLOAT;
    case T_F64: return PT_DOUBLE;
    case T_BINARY: return PT_BYTE_ARRAY;
    default: return PT_INT32;
  }
}
inline const char* arrow_format(uint32_t t) {
  static const char* f[] = {"C", "c", "S", "s", "I", "i", "L", "l", "f", "g", "z"};
  return f[t];
}

// Exception barrier of the C ABI: nothing may unwind across an extern "C" frame (the caller is Rust / C).  Host containers
// (std::vector / std::string / make_unique) can throw bad_alloc; everything else is reported as an internal error.
#define HG_GUARD_BEGIN try {
#define HG_GUARD_END                                                                                              \
  }                                                                                                               \
  catch (const std::bad_alloc&) { return set_error(HG_ERR_OOM, "host allocation failed"); }                      \
  catch (const std::exception& ex) { return set_error(HG_ERR_INTERNAL, std::string("exception: ") + ex.what()); } \
  catch (...) { return set_error(HG_ERR_INTERNAL, "unknown exception"); }

// widen a PLAIN-encoded statistics value to the comparison domain (i64 / u64 / f64 bits)
inline uint64_t widen_stat(const uint8_t raw[8], int phys, uint32_t t) {
  if (phys == PT_INT32) {
    int32_t v;
    std::memcpy(&v, raw, 4);
    switch (t) {
      case T_U8: return uint8_t(v);
      case T_U16: return uint16_t(v);
      case T_U32: return uint32_t(v);
      default: return uint64_t(int64_t(v));
    }
  }
  if (phys == PT_FLOAT) {
    float f;
    std::memcpy(&f, raw, 4);
    double d = f;
    uint64_t v;
    std::memcpy(&v, &d, 8);
    return v;
  }
  uint64_t v;
  std::memcpy(&v, raw, 8);
  return v;
}
inline int cmp_host(uint64_t a, uint64_t b, uint32_t t) {
  if (type_is_float(t)) return cmp_f64_total(a, b);   // IEEE totalOrder, like arrow-rs (device_types.h)
  if (type_is_signed(t)) {
    int64_t x = int64_t(a), y = int64_t(b);
    return x < y ? -1 : (x > y ? 1 : 0);
  }
  return a < b ? -1 : (a > b ? 1 : 0);
}
inline uint64_t pred_literal(const hg_predicate& p, uint32_t t) {
  if (type_is_float(t)) {
    uint64_t v;
    std::memcpy(&v, &p.f64, 8);
    return v;
  }
  return type_is_signed(t) ? uint64_t(p.i64) : p.u64;
}

// ------------------------------------------------------------------------------------------------------ SST residency
// Per (row group, column) planning facts, precomputed at load time: statistics widened to the comparison domain
// (i64 / u64 / f64 bits) so that planning a scan is a tight loop over plain arrays.
struct RgCol {
  uint64_t mn = 0, mx = 0;
  uint32_t scratch = 0;      // decompression scratch of the chunk
  uint8_t has_minmax = 0, null_all = 0, null_none = 0, snappy = 0, simple_page = 0;   // simple = 1 uncompressed V1 page
  uint8_t single_page = 0;   // exactly one V1 PLAIN data page, UNCOMPRESSED or SNAPPY (what the fused scan can address by row)
  uint8_t stored = 0;        // Snappy page whose stream is one or two literals (incompressible data): readable in place
  uint8_t _pad = 0;
};

struct SstResident {
  uint64_t id = 0, size = 0;
  FileMetaData meta;
  std::vector<RgCol> rgcol;      // [rg * ncols + col]
  std::vector<uint32_t> rg_rows;
  std::vector<uint8_t> rg_dead;        // transient loads: row groups proven (on the device) to hold no row passing the predicate
  RgCol* d_rgcol = nullptr;      // the same two tables in HBM (device-side pruning of the fused path)
  uint32_t* d_rg_rows = nullptr;
  // per-file planning facts (over ALL row groups of the file)
  uint64_t rows_total = 0;
  bool col_all_simple[MAX_COLS] = {false}, col_null_none[MAX_COLS] = {false}, col_has_minmax[MAX_COLS] = {false};
  bool col_all_single[MAX_COLS] = {false};       // every chunk: one V1 PLAIN page (any supported codec)
  bool col_any_snappy[MAX_COLS] = {false};
  bool col_snappy_all_stored[MAX_COLS] = {false};  // every Snappy chunk of the column is a stored (literal-only) page
  bool col_snappy_any_stored[MAX_COLS] = {false};  // some Snappy chunk of the column is one
  bool col_any_zstd[MAX_COLS] = {false};           // some chunk of the column is Zstandard-compressed (general pipeline only)
  bool any_zstd = false;
  uint32_t col_max_scratch[MAX_COLS] = {0};      // largest decompression scratch of one chunk of the column
  uint64_t col_comp_bytes[MAX_COLS] = {0};       // compressed bytes of the column (work estimate for the decompressor)
  uint64_t pk0_min = 0, pk0_max = 0;
  bool pk0_range_ok = false;
  uint64_t group_bound = 0;      // sum over row groups of min(#distinct pk0 possible, rows) + 1
  uint8_t* d_bytes = nullptr;
  PageDev* d_pages = nullptr;
  ChunkDev* d_chunks = nullptr;
  uint64_t device_bytes = 0;
  bool owned = true;             // false: transient copy living in the engine arena
  ~SstResident() {
    if (!owned) return;
    if (d_bytes) cudaFree(d_bytes);
    if (d_pages) cudaFree(d_pages);
    if (d_chunks) cudaFree(d_chunks);
    if (d_rgcol) cudaFree(d_rgcol);
    if (d_rg_rows) cudaFree(d_rg_rows);
  }
};

// Device workspace of one engine: a grow-only arena.  Every per-call temporary is a bump allocation; the most recent
// allocation can be popped (LIFO) so large short-lived scratch does not raise the peak; the arena is reset at the start
// of the next call (results handed out as device pointers stay valid until then).  Steady state = no cudaMalloc at all.
struct Arena {
  struct Chunk { char* base; size_t cap, used; };
#ifdef HORAE_EMULATED_BUILD
  static bool chunk_per_alloc() { static const bool on = getenv("HORAE_EMU_GUARD") != nullptr; return on; }
#endif
  std::vector<Chunk> chunks;
  size_t high_water = 0, in_call = 0;
  void* alloc(size_t bytes) {
#ifdef HORAE_EMULATED_BUILD
    const size_t exact = (bytes + 15) & ~size_t(15);
#endif
    bytes = (bytes + 255) & ~size_t(255);
    if (chunks.empty() || chunks.back().used + bytes > chunks.back().cap) {
      size_t cap = std::max<size_t>(bytes, chunks.empty() ? (size_t(64) << 20) : chunks.back().cap * 2);
#ifdef HORAE_EMULATED_BUILD          // the test-suite's CPU emulation (tests/emu): with HORAE_EMU_GUARD every allocation is its own
      if (chunk_per_alloc()) cap = bytes = exact;   // guarded mapping (16-byte granularity): an overrun between arena neighbours is a fault
#endif
      void* p = nullptr;
      if (cudaMalloc(&p, cap) != cudaSuccess) return nullptr;
      chunks.push_back(Chunk{static_cast<char*>(p), cap, 0});
    }
    Chunk& c = chunks.back();
    void* p = c.base + c.used;
    c.used += bytes;
    in_call += bytes;
    if (in_call > high_water) high_water = in_call;
    return p;
  }
  void free_if_top(void* p, size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    if (chunks.empty()) return;
    Chunk& c = chunks.back();
    if (c.used >= bytes && c.base + c.used - bytes == static_cast<char*>(p)) { c.used -= bytes; in_call -= bytes; }
  }
  // start of a call (stream idle): one chunk big enough for everything seen so far
  void reset() {
    in_call = 0;
#ifdef HORAE_EMULATED_BUILD
    if (chunk_per_alloc()) { destroy(); return; }
#endif
    if (chunks.size() > 1) {
      size_t total = 0;
      for (auto& c : chunks) { total += c.cap; cudaFree(c.base); }
      chunks.clear();
      void* p = nullptr;
      if (cudaMalloc(&p, total) == cudaSuccess) chunks.push_back(Chunk{static_cast<char*>(p), total, 0});
    } else if (!chunks.empty()) chunks[0].used = 0;
  }
  void destroy() {
    for (auto& c : chunks) cudaFree(c.base);
    chunks.clear();
  }
};
extern thread_local Arena* g_arena;   // arena of the engine whose call is running on this thread (engine.cu)

// per-call device buffer (arena-backed)
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; }
  DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { reset(); p = o.p; bytes = o.bytes; o.p = nullptr; } return *this; }
  ~DevBuf() { reset(); }
  void reset() {
    if (p && g_arena) g_arena->free_if_top(p, bytes);
    p = nullptr;
  }
  cudaError_t alloc(size_t nbytes, cudaStream_t) {
    reset();
    bytes = nbytes ? nbytes : 16;
    p = g_arena ? g_arena->alloc(bytes) : nullptr;
    return p ? cudaSuccess : cudaErrorMemoryAllocation;
  }
  void* release() { void* q = p; p = nullptr; return q; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct hg_comm;
void hg_comm_free(hg_comm* c);   // comm.cu

struct hg_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint32_t batch_size = 8192;
  uint32_t flags = 0;
  uint64_t budget = 0;
  std::mutex mu;
  std::unordered_map<uint64_t, std::unique_ptr<SstResident>> ssts;
  uint64_t resident_bytes = 0;
  hg_scan_stats stats{};
  uint32_t launches = 0;
  size_t stage_cursor = 0;             // next free byte of h_stage in the current call
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr, evm0 = nullptr, evm1 = nullptr, evd0 = nullptr, evd1 = nullptr;  // call / dominant-kernel / merge brackets
  void* h_stage = nullptr;       // pinned staging for per-scan descriptor uploads
  size_t h_stage_bytes = 0;
  void* h_small = nullptr;       // 256 pinned bytes: the per-call counter block comes back here (one small D2H)
  Arena arena;
  hg_agg_device last_agg{};      // device pointers of the last aggregate (arena memory, valid until the next call)
  uint32_t last_gwidth = 8, last_gtype = T_U64;
  struct hg_comm* comm = nullptr;  // NCCL communicator + combine stream (comm.cu)
  std::vector<uint64_t> transient_ids;   // SSTs loaded only for the running call
  uint32_t trunc_mask = 0;               // bit c: the running call reads column c of a row group only up to its last gate-passing row
  int trunc_gate = -1;                    // ... and the column whose predicates define that row (the device's gate column)
  bool trunc_used = false;               // the transient load shipped a compressed PREFIX of some page (see load_transient)
  Launch L() { return Launch{stream, &launches}; }
};



struct AggBuffers {
  DevBuf gkey, bucket, count, sum, mn, mx;
  uint32_t G = 0;
  uint32_t gwidth = 8, gtype = T_U64;
};


struct ScanPlan {
  std::vector<SstResident*> files;     // in decode order
  std::vector<RgSel> sel;
  std::vector<uint32_t> file_base;     // decoded-row base per file (k+1)
  std::vector<uint32_t> piece_end;     // single-SST pass-through: reader batch boundaries (decoded rows)
  uint64_t rows_in_files = 0, rows_decoded = 0, scratch_bytes = 0;
  bool disjoint = false;               // concatenation in decode order is sorted by PK with no cross-file equal PKs
  std::vector<bool> col_has_nulls;     // per schema column: may any selected chunk contain nulls?
  bool all_single_plain_page = true;   // every selected chunk is one uncompressed V1 page (fused path precondition)
};


// Row-group selection (statistics pruning), decode order, PK-disjointness.  Defined in engine.cu.
int build_plan(hg_engine* e, const hg_schema_desc* schema, const hg_sst_desc* ssts, size_t n, const hg_predicate* preds,
               size_t np, const std::vector<uint32_t>& need_cols, ScanPlan* plan);

// Copies `bytes` of host data to the device through the engine's pinned staging area (async on the engine stream;
// the staging area is reused by the next call, which is safe because every call ends with a stream synchronise).
int stage_upload(hg_engine* e, void* dst, const void* src, size_t bytes, size_t* stage_off);
