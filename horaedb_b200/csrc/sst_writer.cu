// sst_writer.cu — GPU Parquet page encoder + SST assembly: the second half of Executor::do_compaction
// (compaction/executor.rs:173-203: AsyncArrowWriter over the merged stream) and of write_batch (storage.rs:189-225), with
// the writer properties of build_write_props (storage.rs:258-298) and WriteConfig::default (config.rs:120-133):
// row groups of max_row_group_size rows, one DataPage V1 per column chunk, PLAIN values, RLE/bit-packed definition levels
// (every field is nullable), dictionary off, bloom filters off, chunk statistics (min / max / null_count), Snappy or
// uncompressed pages, sorting_columns = primary keys ascending nulls first, Thrift-compact footer.
//
// Device work: page bodies (level prefix + compacted non-null values), chunk statistics, page compression and the final
// gather into one contiguous file image.  Host work: the few KB of Thrift (page headers, footer) and the offsets.
//
// The Snappy compressor is written for what these pages hold — fixed-width numbers: value i is compared with value i-1
// (8-byte columns: how many HIGH bytes agree; 4-byte columns: equal or not) and the page becomes literal runs, 2-byte
// copies of the agreeing high bytes (offset = value width) and 64-byte run-length copies.  Every value computes its own
// emitted size, one prefix sum gives all positions, every value writes its own bytes: no serial parse, any Snappy decoder
// reads the result.  (On the synthetic metric data it lands within a few percent of the reference compressor's ratio.)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "engine_internal.h"
#include "sst_writer.h"

namespace horae {
namespace writer {

namespace {

constexpr int kThreads = 256;

struct PageMetaDev {
  uint32_t uncomp_size, comp_size, null_count, has_minmax;
  uint64_t mn, mx;             // PLAIN bytes of min / max (little endian, low `width` bytes)
};

struct PageJob {
  const void* vals;            // dense column
  const uint8_t* valid;        // one byte per row or nullptr
  uint32_t type, width;        // hg_type, value width in the column array
  uint32_t pwidth;             // physical width in the page (4 or 8)
};

__device__ __forceinline__ uint32_t varint_put(uint8_t* p, uint32_t v) {
  uint32_t n = 0;
  while (v >= 0x80) { p[n++] = uint8_t(v | 0x80); v >>= 7; }
  p[n++] = uint8_t(v);
  return n;
}

// block-wide exclusive scan of one value per thread (256 threads); *total = block sum
__device__ __forceinline__ uint32_t block_scan(uint32_t v, uint32_t* total, uint32_t* s_w /*[9]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
  if (lane == 31) s_w[w] = inc;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t run = 0; for (int x = 0; x < kThreads / 32; x++) { uint32_t c = s_w[x]; s_w[x] = run; run += c; } s_w[8] = run; }
  __syncthreads();
  const uint32_t r = s_w[w] + inc - v;
  *total = s_w[8];
  __syncthreads();
  return r;
}

__device__ __forceinline__ uint64_t load_phys(const PageJob& j, uint32_t row) {
  // the value as PLAIN physical bytes: 1/2-byte integers widen to INT32 (sign- or zero-extended by their type)
  switch (j.type) {
    case T_U8: return reinterpret_cast<const uint8_t*>(j.vals)[row];
    case T_I8: return uint32_t(int32_t(reinterpret_cast<const int8_t*>(j.vals)[row]));
    case T_U16: return reinterpret_cast<const uint16_t*>(j.vals)[row];
    case T_I16: return uint32_t(int32_t(reinterpret_cast<const int16_t*>(j.vals)[row]));
    case T_U32: case T_I32: case T_F32: return reinterpret_cast<const uint32_t*>(j.vals)[row];
    default: return reinterpret_cast<const uint64_t*>(j.vals)[row];
  }
}

// order key of a physical value for the chunk statistics (unsigned compare of the key == typed compare of the value)
__device__ __forceinline__ uint64_t stat_key(uint64_t phys, uint32_t type, bool* is_nan) {
  *is_nan = false;
  switch (type) {
    case T_I8: case T_I16: case T_I32: return uint64_t(int64_t(int32_t(uint32_t(phys)))) ^ (1ull << 63);
    case T_I64: return phys ^ (1ull << 63);
    case T_F32: {
      const float f = __uint_as_float(uint32_t(phys));
      *is_nan = f != f;
      const uint32_t b = uint32_t(phys);
      return uint64_t(b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u));
    }
    case T_F64: {
      const double d = __longlong_as_double((long long)phys);
      *is_nan = d != d;
      return phys ^ ((phys >> 63) ? ~0ull : (1ull << 63));
    }
    default: return phys;
  }
}
__device__ __forceinline__ uint64_t stat_unkey(uint64_t key, uint32_t type) {
  switch (type) {
    case T_I8: case T_I16: case T_I32: return uint64_t(uint32_t(key ^ (1ull << 63)));
    case T_I64: return key ^ (1ull << 63);
    case T_F32: { const uint32_t k = uint32_t(key); return uint64_t(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu)); }
    case T_F64: return key ^ ((key >> 63) ? (1ull << 63) : ~0ull);
    default: return key;
  }
}

// One block per page (row group g, column c): [u32 level bytes][definition levels][PLAIN values of the non-null rows]
__global__ void __launch_bounds__(kThreads) page_body_kernel(const PageJob* __restrict__ jobs, uint32_t ncols, uint32_t R, uint32_t rg_rows,
                                                            uint8_t* __restrict__ body, uint64_t bstride, PageMetaDev* __restrict__ meta) {
  __shared__ uint32_t s_w[9];
  __shared__ uint32_t s_nulls, s_prefix;
  __shared__ unsigned long long s_mn, s_mx;
  __shared__ uint32_t s_seen;
  const uint32_t page = blockIdx.x, g = page / ncols, c = page % ncols;
  const PageJob j = jobs[c];
  const uint32_t row0 = g * rg_rows, rows = (R - row0) < rg_rows ? (R - row0) : rg_rows;
  uint8_t* out = body + uint64_t(page) * bstride;
  const int tid = threadIdx.x;
  if (tid == 0) { s_nulls = 0; s_mn = ~0ull; s_mx = 0; s_seen = 0; }
  __syncthreads();
  // ---- null count
  uint32_t nulls = 0;
  if (j.valid) for (uint32_t i = tid; i < rows; i += kThreads) nulls += j.valid[row0 + i] == 0;
  for (int d = 16; d > 0; d >>= 1) nulls += __shfl_down_sync(0xffffffffu, nulls, d);
  if ((tid & 31) == 0 && nulls) atomicAdd(&s_nulls, nulls);
  __syncthreads();
  const uint32_t nnull = s_nulls;
  // ---- definition levels (bit width 1): one RLE run when uniform, else one bit-packed run of ceil(rows/8) groups
  if (tid == 0) {
    uint32_t n = 0;
    uint8_t* lv = out + 4;
    if (nnull == 0 || nnull == rows) { n = varint_put(lv, rows << 1); lv[n++] = nnull == 0 ? 1 : 0; }
    else { const uint32_t groups = (rows + 7) / 8; n = varint_put(lv, (groups << 1) | 1u); n += groups; }
    out[0] = uint8_t(n); out[1] = uint8_t(n >> 8); out[2] = uint8_t(n >> 16); out[3] = uint8_t(n >> 24);
    s_prefix = 4 + n;
  }
  __syncthreads();
  const uint32_t prefix = s_prefix;
  if (nnull != 0 && nnull != rows) {
    const uint32_t groups = (rows + 7) / 8;
    uint8_t* bits = out + prefix - groups;
    for (uint32_t b = tid; b < groups; b += kThreads) {
      uint32_t v = 0;
#pragma unroll
      for (int k2 = 0; k2 < 8; k2++) { const uint32_t i = b * 8 + k2; if (i < rows && j.valid[row0 + i]) v |= 1u << k2; }
      bits[b] = uint8_t(v);
    }
  }
  // ---- values of the non-null rows, compacted in order; statistics
  uint8_t* vout = out + prefix;
  uint32_t running = 0;
  unsigned long long mn = ~0ull, mx = 0;
  bool seen = false;
  for (uint32_t base = 0; base < rows; base += kThreads) {
    const uint32_t i = base + tid;
    const uint32_t v = (i < rows && (!j.valid || j.valid[row0 + i])) ? 1u : 0u;
    uint32_t total;
    const uint32_t k2 = running + block_scan(v, &total, s_w);
    if (v) {
      const uint64_t x = load_phys(j, row0 + i);
      // (the page body starts 64-byte aligned; the level prefix is usually 8 bytes, so the values are naturally aligned)
      if (j.pwidth == 8) {
        uint8_t* q = vout + size_t(k2) * 8;
        if ((prefix & 7) == 0) *reinterpret_cast<uint64_t*>(q) = x;
        else for (int b = 0; b < 8; b++) q[b] = uint8_t(x >> (8 * b));
      } else {
        uint8_t* q = vout + size_t(k2) * 4;
        if ((prefix & 3) == 0) *reinterpret_cast<uint32_t*>(q) = uint32_t(x);
        else for (int b = 0; b < 4; b++) q[b] = uint8_t(x >> (8 * b));
      }
      bool nan;
      const uint64_t key = stat_key(x, j.type, &nan);
      if (!nan) { mn = key < mn ? key : mn; mx = key > mx ? key : mx; seen = true; }
    }
    running += total;
  }
  if (seen) { atomicMin(&s_mn, mn); atomicMax(&s_mx, mx); atomicOr(&s_seen, 1u); }
  __syncthreads();
  if (tid == 0) {
    PageMetaDev m;
    m.uncomp_size = prefix + (rows - nnull) * j.pwidth;
    m.comp_size = m.uncomp_size;
    m.null_count = nnull;
    m.has_minmax = s_seen;
    m.mn = s_seen ? stat_unkey(s_mn, j.type) : 0;
    m.mx = s_seen ? stat_unkey(s_mx, j.type) : 0;
    meta[page] = m;
  }
}

// ------------------------------------------------------------------------------------------------ Snappy compression
__device__ __forceinline__ uint32_t lit_header_len(uint32_t len) { return len <= 60 ? 1u : (len <= 0x100 ? 2u : (len <= 0x10000 ? 3u : (len <= 0x1000000 ? 4u : 5u))); }
__device__ __forceinline__ uint32_t lit_header_put(uint8_t* p, uint32_t len) {
  const uint32_t n = len - 1;
  if (len <= 60) { p[0] = uint8_t(n << 2); return 1; }
  const uint32_t nb = lit_header_len(len) - 1;
  p[0] = uint8_t((59 + nb) << 2);
  for (uint32_t i = 0; i < nb; i++) p[1 + i] = uint8_t(n >> (8 * i));
  return 1 + nb;
}
// copy of len (4..64) bytes at offset off (< 2048): 2-byte form when len <= 11, else the 3-byte form
__device__ __forceinline__ uint32_t copy_len(uint32_t len) { return len <= 11 ? 2u : 3u; }
__device__ __forceinline__ uint32_t copy_put(uint8_t* p, uint32_t len, uint32_t off) {
  if (len <= 11) { p[0] = uint8_t(1 | ((len - 4) << 2) | ((off >> 8) << 5)); p[1] = uint8_t(off); return 2; }
  p[0] = uint8_t(2 | ((len - 1) << 2)); p[1] = uint8_t(off); p[2] = uint8_t(off >> 8);
  return 3;
}

// class of value i relative to value i-1:  0 = no usable match (literal bytes), 1..4 = k low bytes differ, the 8-k high bytes
// are copied (8-byte values only), 9 = equal (run-length copy)
constexpr uint8_t kClsN = 0, kClsF = 9;

// One block per page.  scratch: per page `sstride` bytes = cls[nv] (u8) + run id / offsets (u32 x 2 per value).
__global__ void __launch_bounds__(kThreads) snappy_encode_kernel(const uint8_t* __restrict__ body, uint64_t bstride, PageMetaDev* __restrict__ meta,
                                                                const PageJob* __restrict__ jobs, uint32_t ncols, uint8_t* __restrict__ comp, uint64_t cstride,
                                                                uint8_t* __restrict__ scratch, uint64_t sstride, uint32_t max_vals) {
  __shared__ uint32_t s_w[9];
  __shared__ uint32_t s_total;
  const uint32_t page = blockIdx.x;
  const PageJob j = jobs[page % ncols];
  const uint32_t w = j.pwidth;
  const uint8_t* in = body + uint64_t(page) * bstride;
  uint8_t* out = comp + uint64_t(page) * cstride;
  const uint32_t ulen = meta[page].uncomp_size;
  const uint32_t prefix = 4 + (uint32_t(in[0]) | (uint32_t(in[1]) << 8) | (uint32_t(in[2]) << 16) | (uint32_t(in[3]) << 24));
  const uint32_t nv = (ulen - prefix) / w;
  uint8_t* cls = scratch + uint64_t(page) * sstride;
  uint32_t* run_of = reinterpret_cast<uint32_t*>(cls + ((max_vals + 15u) & ~15u));   // run index of value i
  uint32_t* run_pos = run_of + max_vals;                                            // first value of run r, then: output offset of run r
  uint32_t* run_out = run_pos + max_vals + 1;
  const int tid = threadIdx.x;
  const uint8_t* v = in + prefix;
  const bool val_aligned = (prefix & (w - 1)) == 0;
  auto val_at = [&](uint32_t i) -> uint64_t {
    if (val_aligned) return w == 8 ? *reinterpret_cast<const uint64_t*>(v + size_t(i) * 8) : uint64_t(*reinterpret_cast<const uint32_t*>(v + size_t(i) * 4));
    uint64_t x = 0;
    for (uint32_t b = 0; b < w; b++) x |= uint64_t(v[size_t(i) * w + b]) << (8 * b);
    return x;
  };
  // ---- classes
  for (uint32_t i = tid; i < nv; i += kThreads) {
    uint8_t c = kClsN;
    if (i > 0) {
      const uint64_t x = val_at(i) ^ val_at(i - 1);
      if (x == 0) c = kClsF;
      else if (w == 8) { const uint32_t k = (71u - uint32_t(__clzll((long long)x))) / 8u; if (k <= 4) c = uint8_t(k); }   // k = differing low bytes
    }
    cls[i] = c;
  }
  __syncthreads();
  // ---- runs: a new run starts where the class changes; partial-match values are runs of their own
  uint32_t running = 0;
  for (uint32_t base = 0; base < nv; base += kThreads) {
    const uint32_t i = base + tid;
    uint32_t st = 0;
    if (i < nv) { const uint8_t c = cls[i]; st = (i == 0 || c != cls[i - 1] || (c >= 1 && c <= 4)) ? 1u : 0u; }
    uint32_t total;
    const uint32_t r = running + block_scan(st, &total, s_w) + st - 1;      // inclusive - 1 = run index
    if (i < nv) { run_of[i] = r; if (st) run_pos[r] = i; }
    running += total;
  }
  const uint32_t nruns = running;
  if (tid == 0) run_pos[nruns] = nv;
  __syncthreads();
  // ---- emitted size of every run; the level prefix joins the first literal run
  uint32_t pre = 0;
  { uint32_t u = ulen; while (u >= 0x80) { pre++; u >>= 7; } pre++; }       // varint(uncompressed length)
  running = 0;
  for (uint32_t base = 0; base < nruns + (nv == 0 ? 1u : 0u); base += kThreads) {
    const uint32_t r = base + tid;
    uint32_t sz = 0;
    if (nv == 0) { if (r == 0) sz = lit_header_len(prefix) + prefix; }
    else if (r < nruns) {
      const uint32_t a = run_pos[r], b = run_pos[r + 1], c = cls[a];
      if (c == kClsN) { const uint32_t len = (b - a) * w + (r == 0 ? prefix : 0); sz = lit_header_len(len) + len; }
      else if (c == kClsF) { const uint32_t bytes = (b - a) * w, full = bytes / 64, rem = bytes % 64; sz = full * 3 + (rem ? copy_len(rem) : 0); }
      else sz = 1 + c + 2;                                                   // literal(c) + copy(8 - c, offset 8)
    }
    uint32_t total;
    const uint32_t o = running + block_scan(sz, &total, s_w);
    if (r < nruns) run_out[r] = o;
    running += total;
  }
  if (tid == 0) s_total = running;
  __syncthreads();
  uint8_t* body_out = out + pre;
  if (tid == 0) {
    varint_put(out, ulen);
    meta[page].comp_size = pre + s_total;
    if (nv == 0) { const uint32_t h = lit_header_put(body_out, prefix); for (uint32_t i = 0; i < prefix; i++) body_out[h + i] = in[i]; }
  }
  if (nv == 0) return;
  // ---- emission: every value writes its own share
  {
    // the first run is a literal run (value 0 has no predecessor): header + level prefix
    const uint32_t len0 = (run_pos[1] - run_pos[0]) * w + prefix;
    const uint32_t h0 = lit_header_len(len0);
    if (tid == 0) lit_header_put(body_out, len0);
    for (uint32_t i = tid; i < prefix; i += kThreads) body_out[h0 + i] = in[i];
  }
  for (uint32_t i = tid; i < nv; i += kThreads) {
    const uint32_t r = run_of[i], a = run_pos[r], b = run_pos[r + 1];
    const uint8_t c = cls[i];
    uint8_t* o = body_out + run_out[r];
    if (c == kClsN) {
      const uint32_t len = (b - a) * w + (r == 0 ? prefix : 0);
      const uint32_t h = lit_header_len(len);
      if (i == a && r != 0) lit_header_put(o, len);
      uint8_t* q = o + h + (r == 0 ? prefix : 0) + (i - a) * w;
      for (uint32_t bb = 0; bb < w; bb++) q[bb] = v[size_t(i) * w + bb];
    } else if (c == kClsF) {
      const uint32_t per = 64 / w, qn = i - a;
      if (qn % per == 0) {
        const uint32_t left = (b - i) * w;
        copy_put(o + (qn / per) * 3, left < 64 ? left : 64, w);
      }
    } else {
      o[0] = uint8_t((uint32_t(c) - 1) << 2);
      for (uint32_t bb = 0; bb < c; bb++) o[1 + bb] = v[size_t(i) * w + bb];
      copy_put(o + 1 + c, 8 - c, 8);
    }
  }
}

struct GatherDesc { uint64_t src_off, dst_off; uint32_t bytes, _pad; };
__global__ void __launch_bounds__(kThreads) gather_pages_kernel(const uint8_t* __restrict__ src, const GatherDesc* __restrict__ d, uint8_t* __restrict__ file) {
  const GatherDesc g = d[blockIdx.x];
  const uint8_t* s = src + g.src_off;
  uint8_t* t = file + g.dst_off;
  // destination-aligned 8-byte stores; the source word comes from two aligned loads + a funnel shift (any relative alignment)
  uint32_t head = uint32_t((8 - (reinterpret_cast<uintptr_t>(t) & 7)) & 7);
  if (head > g.bytes) head = g.bytes;
  if (threadIdx.x < head) t[threadIdx.x] = s[threadIdx.x];
  const uint32_t nwords = (g.bytes - head) >> 3;
  uint64_t* t8 = reinterpret_cast<uint64_t*>(t + head);
  const uint8_t* s0 = s + head;
  for (uint32_t wi = threadIdx.x; wi < nwords; wi += kThreads) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(s0 + (size_t(wi) << 3));
    const uint64_t* q = reinterpret_cast<const uint64_t*>(a & ~uintptr_t(7));
    const uint32_t sh = uint32_t(a & 7) * 8;
    const uint64_t lo = q[0];
    t8[wi] = sh ? ((lo >> sh) | (q[1] << (64 - sh))) : lo;
  }
  for (uint32_t i = head + (nwords << 3) + threadIdx.x; i < g.bytes; i += kThreads) t[i] = s[i];
}

// ------------------------------------------------------------------------------------------------ Thrift compact writer
class TOut {
 public:
  std::vector<uint8_t> b;
  std::vector<int> last;
  void uvar(uint64_t v) { while (v >= 0x80) { b.push_back(uint8_t(v | 0x80)); v >>= 7; } b.push_back(uint8_t(v)); }
  void svar(int64_t v) { uvar((uint64_t(v) << 1) ^ uint64_t(v >> 63)); }
  void begin() { last.push_back(0); }
  void end() { b.push_back(0); last.pop_back(); }
  void field(int id, int type) {
    const int delta = id - last.back();
    if (delta > 0 && delta <= 15) b.push_back(uint8_t((delta << 4) | type));
    else { b.push_back(uint8_t(type)); svar(id); }
    last.back() = id;
  }
  void i32(int id, int64_t v) { field(id, 5); svar(v); }
  void i64(int id, int64_t v) { field(id, 6); svar(v); }
  void boolean(int id, bool v) { field(id, v ? 1 : 2); }
  void binary(int id, const void* p, size_t n) { field(id, 8); uvar(n); const uint8_t* q = static_cast<const uint8_t*>(p); b.insert(b.end(), q, q + n); }
  void str(int id, const std::string& s) { binary(id, s.data(), s.size()); }
  void list(int id, int etype, size_t n) { field(id, 9); if (n < 15) b.push_back(uint8_t((n << 4) | etype)); else { b.push_back(uint8_t(0xf0 | etype)); uvar(n); } }
  void struct_field(int id) { field(id, 12); begin(); }
  void list_str(const std::string& s) { uvar(s.size()); b.insert(b.end(), s.begin(), s.end()); }
};

int phys_of(uint32_t t) { return t == T_U64 || t == T_I64 ? 2 : (t == T_F32 ? 4 : (t == T_F64 ? 5 : 1)); }
int converted_of(uint32_t t) {      // parquet ConvertedType for the integer types that need one (-1: none)
  switch (t) {
    case T_U8: return 11; case T_U16: return 12; case T_U32: return 13; case T_U64: return 14;
    case T_I8: return 15; case T_I16: return 16;
    default: return -1;
  }
}

}  // namespace

int write_sst(hg_engine* e, const hg_schema_desc* schema, const ColIn* cols, uint32_t ncols, uint32_t R, const hg_write_props* props,
              uint8_t** host_out, uint64_t* size_out) {
  cudaStream_t s = e->stream;
  const uint32_t rg_rows = props->max_row_group_size ? props->max_row_group_size : 8192;
  const bool snappy = props->compression == 1;
  if (props->compression > 1) return set_error(HG_ERR_UNSUPPORTED, "write: only UNCOMPRESSED and SNAPPY pages are implemented");
  const uint32_t nrg = (R + rg_rows - 1) / rg_rows;
  const uint64_t npages = uint64_t(nrg) * ncols;
  std::vector<PageJob> jobs(ncols);
  for (uint32_t c = 0; c < ncols; c++) {
    jobs[c] = PageJob{cols[c].vals, cols[c].valid, cols[c].type, cols[c].width, (cols[c].type == T_U64 || cols[c].type == T_I64 || cols[c].type == T_F64) ? 8u : 4u};
  }
  const uint32_t max_vals = std::min<uint32_t>(rg_rows, R ? R : 1);
  const uint64_t bstride = (uint64_t(16) + (max_vals + 7) / 8 + 8 + uint64_t(max_vals) * 8 + 63) & ~uint64_t(63);
  const uint64_t cstride = bstride + 64;
  const uint64_t sstride = ((uint64_t(max_vals) + 15) & ~uint64_t(15)) + (uint64_t(max_vals) * 3 + 4) * 4;
  DevBuf d_jobs, d_body, d_comp, d_meta, d_scratch;
  std::vector<PageMetaDev> meta(npages);
  if (npages) {
    CU_TRY(d_jobs.alloc(jobs.size() * sizeof(PageJob), s));
    CU_TRY(d_body.alloc(npages * bstride, s));
    CU_TRY(d_meta.alloc(npages * sizeof(PageMetaDev), s));
    int rc = stage_upload(e, d_jobs.p, jobs.data(), jobs.size() * sizeof(PageJob), nullptr);
    if (rc) return rc;
    page_body_kernel<<<uint32_t(npages), kThreads, 0, s>>>(d_jobs.as<PageJob>(), ncols, R, rg_rows, d_body.as<uint8_t>(), bstride, d_meta.as<PageMetaDev>());
    e->launches++;
    if (snappy) {
      CU_TRY(d_comp.alloc(npages * cstride, s));
      CU_TRY(d_scratch.alloc(npages * sstride, s));
      snappy_encode_kernel<<<uint32_t(npages), kThreads, 0, s>>>(d_body.as<uint8_t>(), bstride, d_meta.as<PageMetaDev>(), d_jobs.as<PageJob>(), ncols,
                                                                d_comp.as<uint8_t>(), cstride, d_scratch.as<uint8_t>(), sstride, max_vals);
      e->launches++;
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(meta.data(), d_meta.p, npages * sizeof(PageMetaDev), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
  }
  // ---- host: page headers, offsets, footer
  std::vector<std::vector<uint8_t>> headers(npages);
  std::vector<GatherDesc> gd(npages);
  std::vector<uint64_t> hdr_off(npages);
  uint64_t pos = 4;
  for (uint64_t p = 0; p < npages; p++) {
    const uint32_t g = uint32_t(p / ncols);
    const uint32_t rows = std::min<uint32_t>(rg_rows, R - g * rg_rows);
    TOut t;
    t.begin();
    t.i32(1, 0);                           // DATA_PAGE
    t.i32(2, meta[p].uncomp_size);
    t.i32(3, meta[p].comp_size);
    t.struct_field(5);                     // DataPageHeader
    t.i32(1, rows);
    t.i32(2, 0);                           // PLAIN
    t.i32(3, 3);                           // definition levels: RLE
    t.i32(4, 3);                           // repetition levels: RLE
    t.end();
    t.end();
    headers[p] = std::move(t.b);
    hdr_off[p] = pos;
    pos += headers[p].size();
    gd[p] = GatherDesc{p * (snappy ? cstride : bstride), pos, meta[p].comp_size, 0};
    pos += meta[p].comp_size;
  }
  TOut f;
  f.begin();
  f.i32(1, 1);                             // version (WriterVersion::PARQUET_1_0)
  f.list(2, 12, size_t(ncols) + 1);        // schema
  {
    f.begin();
    f.str(4, "arrow_schema");
    f.i32(5, ncols);
    f.end();
    for (uint32_t c = 0; c < ncols; c++) {
      f.begin();
      f.i32(1, phys_of(cols[c].type));
      f.i32(3, 1);                         // OPTIONAL: every field of the reference's schemas is nullable
      std::string tmp;
      f.str(4, schema->names && schema->names[c] ? std::string(schema->names[c]) : "c" + std::to_string(c));
      const int cv = converted_of(cols[c].type);
      if (cv >= 0) f.i32(6, cv);
      f.end();
    }
  }
  f.i64(3, R);
  f.list(4, 12, nrg);
  for (uint32_t g = 0; g < nrg; g++) {
    const uint32_t rows = std::min<uint32_t>(rg_rows, R - g * rg_rows);
    f.begin();
    f.list(1, 12, ncols);
    uint64_t rg_uncomp = 0, rg_comp = 0;
    for (uint32_t c = 0; c < ncols; c++) {
      const uint64_t p = uint64_t(g) * ncols + c;
      const uint64_t hsz = headers[p].size();
      f.begin();                           // ColumnChunk
      f.i64(2, int64_t(hdr_off[p]));       // file_offset
      f.struct_field(3);                   // ColumnMetaData
      f.i32(1, phys_of(cols[c].type));
      f.list(2, 5, 2); f.svar(0); f.svar(3);      // encodings: PLAIN, RLE
      f.list(3, 8, 1); f.list_str(schema->names && schema->names[c] ? std::string(schema->names[c]) : "c" + std::to_string(c));
      f.i32(4, snappy ? 1 : 0);
      f.i64(5, rows);
      f.i64(6, int64_t(meta[p].uncomp_size + hsz));
      f.i64(7, int64_t(meta[p].comp_size + hsz));
      f.i64(9, int64_t(hdr_off[p]));       // data_page_offset
      f.struct_field(12);                  // Statistics
      f.i64(3, meta[p].null_count);
      if (meta[p].has_minmax) {
        const uint32_t pw = jobs[c].pwidth;
        f.binary(5, &meta[p].mx, pw);      // max_value
        f.binary(6, &meta[p].mn, pw);      // min_value
      }
      f.end();
      f.end();
      f.end();
      rg_uncomp += meta[p].uncomp_size + hsz;
      rg_comp += meta[p].comp_size + hsz;
    }
    f.i64(2, int64_t(rg_uncomp));
    f.i64(3, rows);
    if (props->enable_sorting_columns) {
      f.list(4, 12, schema->num_primary_keys);
      for (uint32_t c = 0; c < schema->num_primary_keys; c++) { f.begin(); f.i32(1, c); f.boolean(2, false); f.boolean(3, true); f.end(); }
    }
    f.i64(5, int64_t(hdr_off[uint64_t(g) * ncols]));
    f.i64(6, int64_t(rg_comp));
    f.field(7, 4); f.svar(g);              // ordinal (i16)
    f.end();
  }
  f.str(6, "horaedb_b200 GPU SST writer (PLAIN, RLE levels, " + std::string(snappy ? "SNAPPY" : "UNCOMPRESSED") + ")");
  f.list(7, 12, ncols);                    // column_orders: TYPE_ORDER for every column (makes min_value / max_value usable)
  for (uint32_t c = 0; c < ncols; c++) { f.begin(); f.struct_field(1); f.end(); f.end(); }
  f.end();
  const uint64_t footer_off = pos;
  const uint64_t total = footer_off + f.b.size() + 8;
  if (total > 0xffffffffull) return set_error(HG_ERR_UNSUPPORTED, "output SST larger than 4 GiB (FileMeta.size is u32, sst.rs:155-160)");
  // ---- assemble on the device, one copy back
  uint8_t* host = nullptr;
  CU_TRY(cudaMallocHost(&host, total + 16));
  DevBuf d_file, d_gd;
  CU_TRY(d_file.alloc(total + 16, s));
  std::memcpy(host, "PAR1", 4);
  if (npages) {
    CU_TRY(d_gd.alloc(gd.size() * sizeof(GatherDesc), s));
    CU_TRY(cudaMemcpyAsync(d_gd.p, gd.data(), gd.size() * sizeof(GatherDesc), cudaMemcpyHostToDevice, s));
    gather_pages_kernel<<<uint32_t(npages), kThreads, 0, s>>>(snappy ? d_comp.as<uint8_t>() : d_body.as<uint8_t>(), d_gd.as<GatherDesc>(), d_file.as<uint8_t>());
    e->launches++;
    CU_TRY(cudaMemcpyAsync(host + 4, d_file.as<uint8_t>() + 4, footer_off - 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    for (uint64_t p = 0; p < npages; p++) std::memcpy(host + hdr_off[p], headers[p].data(), headers[p].size());   // a few dozen bytes each
  }
  std::memcpy(host + footer_off, f.b.data(), f.b.size());
  const uint32_t flen = uint32_t(f.b.size());
  std::memcpy(host + footer_off + f.b.size(), &flen, 4);
  std::memcpy(host + footer_off + f.b.size() + 4, "PAR1", 4);
  *host_out = host;
  *size_out = total;
  e->stats.bytes_d2h += total;
  return HG_OK;
}

}  // namespace writer
}  // namespace horae
