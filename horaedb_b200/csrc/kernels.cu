// kernels.cu — general (materialising) pipeline of the columnar hot path, hand-written for sm_100a.
//
//   S2  snappy_chunks / decode_chunks : page decompress, RLE def levels, PLAIN values  (ParquetExec, read.rs:456-465)
//   S3  eval_predicates               : conjunction -> alive bytes                      (FilterExec, read.rs:467-469)
//   S4  build_records / merge_pass    : k-way merge on (pk.., __seq__)                  (SortPreservingMergeExec, read.rs:479-480)
//   S5  dedup_flags_*                 : PK-run boundaries                               (MergeStream::merge_batch, read.rs:289-343)
//   S6  keep last row of each run                                                       (LastValueOperator, operator.rs:39-44)
//   A1/A2 group_flags / reduce_groups : (group, ts/window) runs, sequential f64 sums    (types.rs:82-85 for the window)
//
// All of this is integer / byte work bounded by HBM bandwidth: kernels are grid-stride over 148 SMs, loads are
// coalesced and vectorised where the layout allows, no tensor cores.  Row counts that later kernels depend on stay on
// the device (d_m / d_r / d_g) so the pipeline never synchronises with the host between stages.
#include "kernels.h"

namespace horae {
namespace k {

namespace {

constexpr int kThreads = 256;
constexpr int kSMs = 148;

inline int grid_for(uint64_t n, int per_block = kThreads, int max_blocks = kSMs * 16) {
  uint64_t b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return int(b > uint64_t(max_blocks) ? max_blocks : b);
}

// ---------------------------------------------------------------------------------------------- small device helpers
template <bool kCoherent>
__device__ __forceinline__ uint64_t ld64_any(const uint8_t* p) {
  uintptr_t a = reinterpret_cast<uintptr_t>(p);
  uint32_t sh = uint32_t(a & 7) * 8;
  const uint64_t* q = reinterpret_cast<const uint64_t*>(a & ~uintptr_t(7));
  uint64_t lo = kCoherent ? *reinterpret_cast<const volatile uint64_t*>(q) : __ldg(q);
  if (sh == 0) return lo;
  uint64_t hi = kCoherent ? *reinterpret_cast<const volatile uint64_t*>(q + 1) : __ldg(q + 1);
  return (lo >> sh) | (hi << (64 - sh));
}
__device__ __forceinline__ uint32_t ld32_any(const uint8_t* p) {
  uintptr_t a = reinterpret_cast<uintptr_t>(p);
  uint32_t sh = uint32_t(a & 3) * 8;
  const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
  uint32_t lo = __ldg(q);
  if (sh == 0) return lo;
  uint32_t hi = __ldg(q + 1);
  return (lo >> sh) | (hi << (32 - sh));
}

__device__ __forceinline__ bool type_signed(uint32_t t) { return t == T_I8 || t == T_I16 || t == T_I32 || t == T_I64; }
__device__ __forceinline__ bool type_float(uint32_t t) { return t == T_F32 || t == T_F64; }

// raw value, zero-extended to 64 bits (bit pattern at native width)
__device__ __forceinline__ uint64_t col_raw(const ColView& c, uint32_t row) {
  switch (c.width) {
    case 1: return reinterpret_cast<const uint8_t*>(c.vals)[row];
    case 2: return reinterpret_cast<const uint16_t*>(c.vals)[row];
    case 4: return reinterpret_cast<const uint32_t*>(c.vals)[row];
    default: return reinterpret_cast<const uint64_t*>(c.vals)[row];
  }
}
// widened domain: signed -> sign-extended i64 bits, unsigned -> u64, floats -> f64 bits
__device__ __forceinline__ uint64_t col_widened(const ColView& c, uint32_t row) {
  uint64_t r = col_raw(c, row);
  switch (c.type) {
    case T_I8: return uint64_t(int64_t(int8_t(r)));
    case T_I16: return uint64_t(int64_t(int16_t(r)));
    case T_I32: return uint64_t(int64_t(int32_t(r)));
    case T_F32: return uint64_t(__double_as_longlong(double(__uint_as_float(uint32_t(r)))));
    default: return r;
  }
}
__device__ __forceinline__ bool col_valid(const ColView& c, uint32_t row) { return c.valid == nullptr || c.valid[row] != 0; }

__device__ __forceinline__ int cmp_widened(uint64_t a, uint64_t b, uint32_t t) {
  if (type_float(t)) return cmp_f64_total(a, b);
  if (type_signed(t)) {
    int64_t x = int64_t(a), y = int64_t(b);
    return x < y ? -1 : (x > y ? 1 : 0);
  }
  return a < b ? -1 : (a > b ? 1 : 0);
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across a 256-thread block; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total, uint32_t* s_warp /*[9]*/) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = warp_incl_scan(v, lane);
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t x = lane < (kThreads / 32) ? s_warp[lane] : 0;
    uint32_t xi = warp_incl_scan(x, lane);
    if (lane < (kThreads / 32)) s_warp[lane] = xi - x;
    if (lane == (kThreads / 32) - 1) s_warp[8] = xi;
  }
  __syncthreads();
  uint32_t r = s_warp[w] + inc - v;
  *total = s_warp[8];
  __syncthreads();
  return r;
}

__host__ __device__ __forceinline__ uint64_t page_scratch(uint32_t uncomp) { return (uint64_t(uncomp) + 15u) / 16u * 16u + 32u; }

// ------------------------------------------------------------------------------------------------ Snappy (raw format)
// One warp per column chunk; elements are processed in stream order, every copy is spread over the 32 lanes.
template <bool kCoherent>
__device__ __forceinline__ void warp_copy(uint8_t* dst, const uint8_t* src, uint32_t len, int lane) {
  if (len < 32) {
    if (uint32_t(lane) < len) dst[lane] = kCoherent ? *reinterpret_cast<const volatile uint8_t*>(src + lane) : __ldg(src + lane);
    return;
  }
  uint32_t head = uint32_t((8 - (reinterpret_cast<uintptr_t>(dst) & 7)) & 7);
  if (uint32_t(lane) < head) dst[lane] = kCoherent ? *reinterpret_cast<const volatile uint8_t*>(src + lane) : __ldg(src + lane);
  uint32_t nwords = (len - head) >> 3;
  uint64_t* d8 = reinterpret_cast<uint64_t*>(dst + head);
  const uint8_t* s = src + head;
  for (uint32_t w = lane; w < nwords; w += 32) d8[w] = ld64_any<kCoherent>(s + (size_t(w) << 3));
  uint32_t done = head + (nwords << 3);
  uint32_t rem = len - done;
  if (uint32_t(lane) < rem)
    dst[done + lane] = kCoherent ? *reinterpret_cast<const volatile uint8_t*>(src + done + lane) : __ldg(src + done + lane);
}

__device__ void snappy_warp(const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t ulen_expected, int lane, int* err) {
  uint32_t pos = 0;
  uint32_t ulen = 0;
  for (int sh = 0; pos < n && sh < 35; sh += 7) {
    uint32_t b = __ldg(src + pos++);
    ulen |= (b & 0x7f) << sh;
    if (!(b & 0x80)) break;
  }
  if (ulen != ulen_expected) { if (lane == 0) atomicExch(err, 101); return; }
  uint32_t o = 0;
  while (pos < n) {
    __syncwarp();
    uint32_t tag = __ldg(src + pos++);
    uint32_t len, off = 0;
    uint32_t kind = tag & 3;
    if (kind == 0) {
      len = (tag >> 2) + 1;
      if (len > 60) {
        uint32_t nb = len - 60;
        len = 0;
        for (uint32_t i = 0; i < nb; i++) len |= uint32_t(__ldg(src + pos + i)) << (8 * i);
        len += 1;
        pos += nb;
      }
      if (pos + len > n || o + len > ulen) { if (lane == 0) atomicExch(err, 102); return; }
      warp_copy<false>(dst + o, src + pos, len, lane);
      pos += len;
      o += len;
      continue;
    }
    if (kind == 1) {
      len = ((tag >> 2) & 7) + 4;
      off = ((tag >> 5) << 8) | __ldg(src + pos);
      pos += 1;
    } else if (kind == 2) {
      len = (tag >> 2) + 1;
      off = uint32_t(__ldg(src + pos)) | (uint32_t(__ldg(src + pos + 1)) << 8);
      pos += 2;
    } else {
      len = (tag >> 2) + 1;
      off = uint32_t(__ldg(src + pos)) | (uint32_t(__ldg(src + pos + 1)) << 8) | (uint32_t(__ldg(src + pos + 2)) << 16) |
            (uint32_t(__ldg(src + pos + 3)) << 24);
      pos += 4;
    }
    if (off == 0 || off > o || o + len > ulen) { if (lane == 0) atomicExch(err, 103); return; }
    if (off >= len) {
      warp_copy<true>(dst + o, dst + o - off, len, lane);
    } else {
      // overlapping copy = the last `off` bytes repeated: every source byte already exists, so lanes are independent
      const volatile uint8_t* base = dst + o - off;
      for (uint32_t i = lane; i < len; i += 32) dst[o + i] = base[i % off];
    }
    o += len;
  }
  if (o != ulen) { if (lane == 0) atomicExch(err, 104); }
}

__device__ __forceinline__ uint64_t chunk_scratch_off(const RgSel& rs, const ChunkDev* chunks, const ColSel* cols, int ci) {
  uint64_t off = rs.scratch_off;
  for (int j = 0; j < ci; j++) {
    off += chunks[cols[j].col].scratch_bytes;      // 0 for uncompressed PLAIN chunks
  }
  return off;
}

__global__ void __launch_bounds__(32) snappy_chunks_kernel(const SstDev* __restrict__ ssts, const RgSel* __restrict__ sel,
                                                          const ColSel* __restrict__ cols, int ncolsel,
                                                          uint8_t* __restrict__ scratch, int* err) {
  int lane = threadIdx.x;
  uint32_t si = blockIdx.x / ncolsel;
  int ci = blockIdx.x % ncolsel;
  RgSel rs = sel[si];
  SstDev sst = ssts[rs.sst];
  const ChunkDev* chunks = sst.chunks + size_t(rs.rg) * sst.ncols;
  ChunkDev ch = chunks[cols[ci].col];
  if (ch.codec != 1) return;
  uint8_t* dst = scratch + chunk_scratch_off(rs, chunks, cols, ci);
  if (ch.dict_uncomp) {
    snappy_warp(sst.bytes + ch.dict_payload_off, ch.dict_comp, dst, ch.dict_uncomp, lane, err);
    dst += page_scratch(ch.dict_uncomp);
  }
  for (uint32_t p = 0; p < ch.num_pages; p++) {
    PageDev pg = sst.pages[ch.first_page + p];
    const uint8_t* src = sst.bytes + pg.payload_off;
    uint32_t n = pg.comp_size, ulen = pg.uncomp_size;
    bool compressed = true;
    if (pg.page_type == 3) {
      uint32_t skip = pg.v2_def_len + pg.v2_rep_len;
      src += skip; n -= skip; ulen -= skip;
      compressed = pg.v2_compressed != 0;
    }
    if (compressed) snappy_warp(src, n, dst, ulen, lane, err);
    dst += page_scratch(pg.uncomp_size);
    if (pg.encoding == 5 || pg.encoding == 6 || pg.encoding == 8 || pg.encoding == 2) dst += page_scratch(pg.num_values * 8u);
  }
}

// ------------------------------------------------------------------------------- def levels + PLAIN values -> columns
template <int OW>
__device__ __forceinline__ void store_val(void* out, uint32_t row, uint64_t v) {
  if (OW == 1) reinterpret_cast<uint8_t*>(out)[row] = uint8_t(v);
  else if (OW == 2) reinterpret_cast<uint16_t*>(out)[row] = uint16_t(v);
  else if (OW == 4) reinterpret_cast<uint32_t*>(out)[row] = uint32_t(v);
  else reinterpret_cast<uint64_t*>(out)[row] = v;
}
__device__ __forceinline__ void store_val_dyn(void* out, uint32_t ow, uint32_t row, uint64_t v) {
  switch (ow) {
    case 1: store_val<1>(out, row, v); break;
    case 2: store_val<2>(out, row, v); break;
    case 4: store_val<4>(out, row, v); break;
    default: store_val<8>(out, row, v);
  }
}

// ------------------------------------------------------------------------------------ DELTA_BINARY_PACKED -> PLAIN
// (Apache Parquet Encodings.md; a per-column option of the reference's writer, config.rs:54-75.)  One block per page:
// thread 0 walks the block headers (min delta, one bit width per miniblock), all threads unpack one block's deltas in
// parallel and a block-wide prefix sum turns them into values.  Sums wrap in the physical width, like the reference decoder.
__device__ __forceinline__ uint64_t block_incl_scan64(uint64_t v, uint64_t* total, uint64_t* s_w64 /*[9]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t lo = __shfl_up_sync(0xffffffffu, uint32_t(inc), d), hi = __shfl_up_sync(0xffffffffu, uint32_t(inc >> 32), d);
    if (lane >= d) inc += (uint64_t(hi) << 32) | lo;
  }
  if (lane == 31) s_w64[w] = inc;
  __syncthreads();
  if (threadIdx.x == 0) { uint64_t run = 0; for (int x = 0; x < kThreads / 32; x++) { const uint64_t c = s_w64[x]; s_w64[x] = run; run += c; } s_w64[8] = run; }
  __syncthreads();
  const uint64_t r = s_w64[w] + inc;
  *total = s_w64[8];
  __syncthreads();
  return r;
}

// *end_pos (optional): bytes of [p, end) the encoded values occupy (what follows is the caller's: DELTA_LENGTH_BYTE_ARRAY data)
__device__ bool delta_decode_page(const uint8_t* p, const uint8_t* end, uint32_t pw, uint32_t max_out, uint8_t* out, uint32_t* count_out,
                                  uint32_t* end_pos = nullptr) {
  __shared__ uint64_t s_w64[9];
  __shared__ uint64_t s_min, s_cur;
  __shared__ uint32_t s_block, s_nmini, s_total, s_pos, s_ok;
  __shared__ uint32_t s_bw[64], s_moff[64];
  const int tid = threadIdx.x;
  auto varint = [&](uint32_t& pos, uint64_t* v) -> bool {
    uint64_t r = 0;
    for (int sh = 0; sh < 70; sh += 7) {
      if (p + pos >= end) return false;
      const uint8_t b = __ldg(p + pos++);
      r |= uint64_t(b & 0x7f) << sh;
      if (!(b & 0x80)) { *v = r; return true; }
    }
    return false;
  };
  if (tid == 0) {
    uint32_t pos = 0;
    uint64_t block = 0, nmini = 0, total = 0, zz = 0;
    bool ok = varint(pos, &block) && varint(pos, &nmini) && varint(pos, &total) && varint(pos, &zz);
    ok = ok && nmini > 0 && nmini <= 64 && block > 0 && block <= 65536 && block % nmini == 0 && (block / nmini) % 32 == 0 && total <= max_out;
    s_ok = ok;
    s_block = uint32_t(block); s_nmini = uint32_t(nmini); s_total = uint32_t(total); s_pos = pos;
    s_cur = (zz >> 1) ^ (0 - (zz & 1));
    if (ok && total > 0) {
      if (pw == 4) *reinterpret_cast<uint32_t*>(out) = uint32_t(s_cur); else *reinterpret_cast<uint64_t*>(out) = s_cur;
    }
  }
  __syncthreads();
  if (!s_ok) return false;
  const uint32_t total = s_total, block = s_block, nmini = s_nmini, per = block / nmini;
  *count_out = total;
  uint32_t done = total ? 1u : 0u;
  while (done < total) {
    const uint32_t nvals = (total - done) < block ? (total - done) : block;
    if (tid == 0) {
      uint32_t pos = s_pos;
      uint64_t mz = 0;
      bool ok = varint(pos, &mz) && p + pos + nmini <= end;
      s_min = (mz >> 1) ^ (0 - (mz & 1));
      if (ok) {
        const uint32_t bwpos = pos;
        pos += nmini;
        for (uint32_t m = 0; m < nmini; m++) {
          const uint32_t b = __ldg(p + bwpos + m);
          s_bw[m] = b;
          s_moff[m] = pos;
          if (m * per < nvals) { if (b > 64) ok = false; pos += per * b / 8; }      // miniblocks past the last value are not stored
        }
        if (p + pos > end) ok = false;
      }
      s_pos = pos;
      s_ok = ok;
    }
    __syncthreads();
    if (!s_ok) return false;
    for (uint32_t base = 0; base < nvals; base += kThreads) {
      const uint32_t v = base + tid;
      uint64_t delta = 0;
      if (v < nvals) {
        const uint32_t m = v / per, j = v % per, b = s_bw[m];
        if (b) {
          const uint64_t bit = uint64_t(j) * b;
          const uint8_t* q = p + s_moff[m] + (bit >> 3);
          const uint32_t sh = uint32_t(bit & 7);
          const uint64_t lo = ld64_any<false>(q);
          uint64_t x = lo >> sh;
          if (sh && b + sh > 64) x |= uint64_t(__ldg(q + 8)) << (64 - sh);
          delta = b == 64 ? x : (x & ((1ull << b) - 1));
        }
        delta += s_min;
      }
      uint64_t tile_total;
      const uint64_t incl = block_incl_scan64(delta, &tile_total, s_w64);
      const uint64_t cur = s_cur;
      if (v < nvals) {
        const uint64_t val = cur + incl;
        if (pw == 4) reinterpret_cast<uint32_t*>(out)[done + v] = uint32_t(val); else reinterpret_cast<uint64_t*>(out)[done + v] = val;
      }
      __syncthreads();
      if (tid == 0) s_cur = cur + tile_total;
      __syncthreads();
    }
    done += nvals;
  }
  if (end_pos) *end_pos = s_pos;
  return true;
}

// ------------------------------------------------------------------------------------ RLE_DICTIONARY -> PLAIN
// Data page = [bit width][RLE / bit-packed hybrid runs of dictionary indices] (Parquet Encodings.md; enable_dict, config.rs:98-103).
// Thread 0 walks the run headers, all threads expand a run: out[i] = dict[index_i] as PLAIN values of width pw.
__device__ bool dict_decode_page(const uint8_t* p, const uint8_t* end, uint32_t pw, uint32_t max_out, const uint8_t* dict, uint32_t dict_n,
                                 uint8_t* out, uint32_t* count_out) {
  __shared__ uint32_t s_kind, s_cnt, s_idx, s_pos, s_ok, s_bad_idx;
  const int tid = threadIdx.x;
  if (tid == 0) { s_pos = 1; s_ok = p < end && __ldg(p) <= 32; s_bad_idx = 0; }
  __syncthreads();
  if (!s_ok) return false;
  const uint32_t bw = __ldg(p);
  uint32_t done = 0;
  while (done < max_out) {
    if (tid == 0) {
      uint32_t pos = s_pos;
      bool ok = true;
      if (p + pos >= end) { s_cnt = 0; }
      else {
        uint64_t h = 0;
        int sh = 0;
        for (;;) {
          if (p + pos >= end || sh > 35) { ok = false; break; }
          const uint8_t b = __ldg(p + pos++);
          h |= uint64_t(b & 0x7f) << sh;
          sh += 7;
          if (!(b & 0x80)) break;
        }
        if (ok && (h & 1)) {                           // bit-packed run: (h >> 1) groups of 8 indices
          const uint64_t groups = h >> 1, bytes = groups * bw;
          if (groups == 0 || p + pos + bytes > end) ok = false;
          s_kind = 1; s_cnt = uint32_t(groups * 8 > 0xffffffffull ? 0xffffffffu : groups * 8); s_idx = pos;
          pos += uint32_t(bytes);
        } else if (ok) {                               // RLE run: count, then the index in ceil(bw / 8) bytes
          const uint32_t nb = (bw + 7) / 8;
          if ((h >> 1) == 0 || p + pos + nb > end) ok = false;
          uint32_t idx = 0;
          if (ok) for (uint32_t b = 0; b < nb; b++) idx |= uint32_t(__ldg(p + pos + b)) << (8 * b);
          s_kind = 0; s_cnt = uint32_t((h >> 1) > 0xffffffffull ? 0xffffffffu : (h >> 1)); s_idx = idx;
          pos += nb;
        }
      }
      s_pos = pos;
      s_ok = ok;
    }
    __syncthreads();
    if (!s_ok) return false;
    uint32_t cnt = s_cnt;
    if (cnt == 0) break;                               // index bytes exhausted
    if (cnt > max_out - done) cnt = max_out - done;
    const uint32_t kind = s_kind, ref = s_idx;
    for (uint32_t j = tid; j < cnt; j += kThreads) {
      uint32_t idx = ref;
      if (kind) {
        const uint64_t bit = uint64_t(j) * bw;
        const uint8_t* q = p + ref + (bit >> 3);
        const uint64_t x = ld64_any<false>(q) >> (bit & 7);
        idx = bw == 32 ? uint32_t(x) : uint32_t(x & ((1ull << bw) - 1));
      }
      if (idx >= dict_n) { s_bad_idx = 1; idx = 0; }
      if (dict_n) {
        if (pw == 4) reinterpret_cast<uint32_t*>(out)[done + j] = ld32_any(dict + size_t(idx) * 4);
        else reinterpret_cast<uint64_t*>(out)[done + j] = ld64_any<false>(dict + size_t(idx) * 8);
      }
    }
    done += cnt;
    __syncthreads();
  }
  __syncthreads();
  *count_out = done;
  return s_bad_idx == 0;
}

__global__ void __launch_bounds__(kThreads) decode_chunks_kernel(const SstDev* __restrict__ ssts, const RgSel* __restrict__ sel,
                                                                const ColSel* __restrict__ cols, int ncolsel,
                                                                uint8_t* __restrict__ scratch, int* err) {
  __shared__ uint32_t s_warp[9];
  __shared__ uint64_t s_w64b[9];
  __shared__ uint32_t s_kind, s_count, s_val, s_bad;
  __shared__ const uint8_t* s_ptr;
  const int tid = threadIdx.x;
  uint32_t si = blockIdx.x / ncolsel;
  int ci = blockIdx.x % ncolsel;
  RgSel rs = sel[si];
  SstDev sst = ssts[rs.sst];
  const ChunkDev* chunks = sst.chunks + size_t(rs.rg) * sst.ncols;
  ColSel cs = cols[ci];
  ChunkDev ch = chunks[cs.col];
  const uint32_t pw = (ch.phys == 1 || ch.phys == 4) ? 4u : 8u;   // INT32/FLOAT : INT64/DOUBLE (BYTE_ARRAY: variable, handled apart)
  uint8_t* sc = scratch + (ch.scratch_bytes ? chunk_scratch_off(rs, chunks, cols, ci) : 0);
  const uint8_t* dict = sst.bytes + ch.dict_payload_off;             // dictionary values (PLAIN): in place, or decompressed first in the scratch
  if (ch.dict_uncomp && ch.codec != 0) { dict = sc; sc += page_scratch(ch.dict_uncomp); }
  const uint32_t dict_n = ch.dict_uncomp / pw;
  uint32_t row = rs.out_row;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  for (uint32_t p = 0; p < ch.num_pages; p++) {
    PageDev pg = sst.pages[ch.first_page + p];
    const uint32_t nv = pg.num_values;
    const uint8_t* payload = sst.bytes + pg.payload_off;
    const uint8_t* lv_ptr = nullptr;
    uint32_t lv_len = 0;
    const uint8_t* val_ptr;
    if (pg.page_type == 3) {            // V2: levels uncompressed in front, values optionally compressed
      lv_ptr = payload + pg.v2_rep_len;
      lv_len = pg.v2_def_len;
      val_ptr = (ch.codec != 0 && pg.v2_compressed) ? sc : payload + pg.v2_rep_len + pg.v2_def_len;
    } else {                            // V1: [u32 len][levels][values], compressed as a whole
      const uint8_t* body = ch.codec != 0 ? sc : payload;        // 1 Snappy, 6 Zstandard: decompressed into the scratch before this kernel
      if (ch.optional) {
        lv_len = ld32_any(body);
        lv_ptr = body + 4;
        val_ptr = body + 4 + lv_len;
      } else val_ptr = body;
    }
    // values available in this page (malformed pages must not make the decoder read past the page: ABI = HG_ERR_FORMAT)
    const uint8_t* page_end = (ch.codec != 0 && !(pg.page_type == 3 && !pg.v2_compressed)) ? sc + (pg.page_type == 3 ? pg.uncomp_size - pg.v2_def_len - pg.v2_rep_len : pg.uncomp_size)
                                                                                       : payload + pg.comp_size;
    uint32_t max_vals = val_ptr <= page_end ? uint32_t(size_t(page_end - val_ptr) / pw) : 0u;
    if (ch.codec != 0) sc += page_scratch(pg.uncomp_size);
    if (pg.encoding == 5) {
      // DELTA_BINARY_PACKED: expand the values into the page's PLAIN image in scratch, then decode that like any PLAIN page
      uint8_t* img = sc;
      sc += page_scratch(nv * 8u);
      uint32_t cnt = 0;
      const bool ok = val_ptr <= page_end && delta_decode_page(val_ptr, page_end, pw, nv, img, &cnt);
      __syncthreads();
      if (!ok) { if (tid == 0) s_bad = 4; cnt = 0; }
      val_ptr = img;
      max_vals = cnt;
      __syncthreads();
    } else if (pg.encoding == 8 || pg.encoding == 2) {
      uint8_t* img = sc;
      sc += page_scratch(nv * 8u);
      uint32_t cnt = 0;
      const bool ok = val_ptr <= page_end && dict_decode_page(val_ptr, page_end, pw, nv, dict, dict_n, img, &cnt);
      __syncthreads();
      if (!ok) { if (tid == 0) s_bad = 5; cnt = 0; }
      val_ptr = img;
      max_vals = cnt;
      __syncthreads();
    }

    bool all_valid = true;
    if (ch.optional) {
      // RLE / bit-packed hybrid, bit width 1.  Fast path: a single RLE run of 1s covering the page.
      const uint8_t* lp = lv_ptr;
      const uint8_t* lend = lv_ptr + lv_len;
      uint32_t i = 0;
      bool first = true;
      while (i < nv) {
        if (tid == 0) {
          uint32_t h = 0;
          int sh = 0;
          bool ok = false;
          while (lp < lend && sh < 35) {
            uint32_t b = __ldg(lp++);
            h |= (b & 0x7f) << sh;
            sh += 7;
            if (!(b & 0x80)) { ok = true; break; }
          }
          uint32_t kind = h & 1, cnt = h >> 1;
          if (kind) { cnt *= 8; s_ptr = lp; lp += (h >> 1); }
          else { s_val = (lp < lend) ? (__ldg(lp) & 1u) : 0u; lp += 1; }
          if (!ok || cnt == 0 || lp > lend) { s_bad = 1; cnt = nv; kind = 0; s_val = 1; }
          s_kind = kind;
          s_count = cnt;
        }
        __syncthreads();
        uint32_t kind = s_kind, cnt = s_count, val = s_val;
        const uint8_t* bp = s_ptr;
        if (cnt > nv - i) cnt = nv - i;
        if (first && kind == 0 && val == 1 && cnt == nv) { i = nv; __syncthreads(); break; }
        first = false;
        all_valid = false;
        if (cs.out_valid == nullptr) { if (tid == 0) s_bad = 2; }
        else {
          for (uint32_t j = tid; j < cnt; j += kThreads)
            cs.out_valid[row + i + j] = kind ? uint8_t((__ldg(bp + (j >> 3)) >> (j & 7)) & 1u) : uint8_t(val);
        }
        i += cnt;
        __syncthreads();
      }
    }
    if (ch.phys == 6 && pg.encoding == 6) {
      // BYTE_ARRAY, DELTA_LENGTH_BYTE_ARRAY (config.rs:54-75): [DELTA_BINARY_PACKED lengths of the non-null values][their bytes, back to
      // back].  The lengths expand into the page's scratch image, an exclusive scan turns them into offsets, and every row points at
      // its bytes in place.
      const uint8_t** optr = reinterpret_cast<const uint8_t**>(cs.out_vals);
      uint32_t* lens = reinterpret_cast<uint32_t*>(sc);
      sc += page_scratch(nv * 8u);
      uint32_t cnt = 0, used = 0;
      bool ok = val_ptr <= page_end && delta_decode_page(val_ptr, page_end, 4, nv, reinterpret_cast<uint8_t*>(lens), &cnt, &used);
      __syncthreads();
      const uint8_t* data = val_ptr + used;
      const uint64_t room = ok && data <= page_end ? uint64_t(page_end - data) : 0;
      if (!ok || data > page_end) { if (tid == 0) s_bad = 7; cnt = 0; }
      if (all_valid && cs.out_valid) for (uint32_t j = tid; j < nv; j += kThreads) cs.out_valid[row + j] = 1;
      uint32_t run_vals = 0;                                  // non-null values before the current tile
      uint64_t run_off = 0;                                   // ... and their bytes
      for (uint32_t base = 0; base < nv; base += kThreads) {
        const uint32_t j = base + tid;
        const uint32_t v = (j < nv) ? (all_valid ? 1u : uint32_t(cs.out_valid[row + j] != 0)) : 0u;
        uint32_t tile_vals;
        const uint32_t kidx = run_vals + block_excl_scan(v, &tile_vals, s_warp);
        const uint64_t len = (v && kidx < cnt) ? lens[kidx] : 0;
        uint64_t tile_bytes;
        const uint64_t incl = block_incl_scan64(len, &tile_bytes, s_w64b);
        if (j < nv) {
          const uint64_t off = run_off + incl - len;
          if (v && (kidx >= cnt || off + len > room)) { s_bad = 7; optr[row + j] = nullptr; cs.out_lens[row + j] = 0; }
          else if (v) { optr[row + j] = data + off; cs.out_lens[row + j] = uint32_t(len); }
          else { optr[row + j] = nullptr; cs.out_lens[row + j] = 0; }
        }
        run_vals += tile_vals;
        run_off += tile_bytes;
        __syncthreads();
      }
    } else if (ch.phys == 6) {
      // BYTE_ARRAY, PLAIN: [u32 length][bytes] per non-null value — a serial walk (a value's position depends on every length
      // before it); Binary values of this engine's tables are few and large (batched payloads), one thread does it.  Rows point
      // at their bytes in place (page payload or decompression scratch): nothing is copied here.
      const uint8_t** optr = reinterpret_cast<const uint8_t**>(cs.out_vals);
      if (tid == 0) {
        const uint8_t* p = val_ptr;
        bool bad = p > page_end;
        for (uint32_t j = 0; j < nv && !bad; j++) {
          const bool v = all_valid || cs.out_valid[row + j] != 0;
          if (v) {
            if (p + 4 > page_end) { bad = true; break; }
            const uint32_t len = ld32_any(p);
            if (len > uint32_t(page_end - p - 4)) { bad = true; break; }
            optr[row + j] = p + 4;
            cs.out_lens[row + j] = len;
            p += 4 + size_t(len);
          } else { optr[row + j] = nullptr; cs.out_lens[row + j] = 0; }
        }
        if (bad) s_bad = 6;
      }
      if (all_valid && cs.out_valid) for (uint32_t j = tid; j < nv; j += kThreads) cs.out_valid[row + j] = 1;
    } else if (all_valid && nv > max_vals) {
      if (tid == 0) s_bad = 3;
    } else if (all_valid) {
      if (pw == 8) {
        for (uint32_t j = tid; j < nv; j += kThreads) store_val<8>(cs.out_vals, row + j, ld64_any<false>(val_ptr + size_t(j) * 8));
      } else {
        for (uint32_t j = tid; j < nv; j += kThreads) store_val_dyn(cs.out_vals, cs.out_width, row + j, ld32_any(val_ptr + size_t(j) * 4));
      }
      if (cs.out_valid) for (uint32_t j = tid; j < nv; j += kThreads) cs.out_valid[row + j] = 1;
    } else if (cs.out_valid) {
      uint32_t running = 0;
      for (uint32_t base = 0; base < nv; base += kThreads) {
        uint32_t j = base + tid;
        uint32_t v = (j < nv) ? cs.out_valid[row + j] : 0;
        uint32_t total;
        uint32_t kidx = running + block_excl_scan(v, &total, s_warp);
        if (j < nv) {
          uint64_t x = 0;
          if (v && kidx >= max_vals) { s_bad = 3; v = 0; }
          if (v) x = pw == 8 ? ld64_any<false>(val_ptr + size_t(kidx) * 8) : uint64_t(ld32_any(val_ptr + size_t(kidx) * 4));
          store_val_dyn(cs.out_vals, cs.out_width, row + j, x);
        }
        running += total;
      }
    }
    row += nv;
    __syncthreads();
  }
  if (tid == 0 && s_bad) atomicExch(err, 110 + int(s_bad));
}

// ---------------------------------------------------------------------------------------------------- S3: predicates
__global__ void __launch_bounds__(kThreads) eval_predicates_kernel(PredSet preds, uint32_t n, uint8_t* __restrict__ alive) {
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    bool keep = true;
    for (int p = 0; p < preds.n && keep; p++) {
      const PredDev& pd = preds.p[p];
      if (!col_valid(pd.col, i)) { keep = false; break; }      // NULL => false
      if (pd.op == OP_IN) {
        const uint64_t v = col_widened(pd.col, i);
        bool any = false;
        for (uint32_t j = 0; j < pd.n_in && !any; j++) any = cmp_widened(v, pd.in_list[j], pd.col.type) == 0;
        keep = any;
        continue;
      }
      int c = cmp_widened(col_widened(pd.col, i), pd.lit, pd.col.type);
      switch (pd.op) {
        case OP_EQ: keep = c == 0; break;
        case OP_NE: keep = c != 0; break;
        case OP_LT: keep = c < 0; break;
        case OP_LE: keep = c <= 0; break;
        case OP_GT: keep = c > 0; break;
        default: keep = c >= 0;
      }
    }
    alive[i] = keep ? 1 : 0;
  }
}

// --------------------------------------------------------------------------------------------- stream compaction
constexpr int kCompactPerThread = 8;
constexpr int kCompactTile = kThreads * kCompactPerThread;  // 2048 flags per block

__device__ __forceinline__ uint32_t load_flags8(const uint8_t* flags, uint32_t base, uint32_t n, uint8_t f[8]) {
  uint32_t cnt = 0;
  if (base + 8 <= n) {
    uint2 v = *reinterpret_cast<const uint2*>(flags + base);   // base is a multiple of 8 and flags is 256B-aligned
    uint32_t w0 = v.x, w1 = v.y;
#pragma unroll
    for (int i = 0; i < 4; i++) { f[i] = (w0 >> (8 * i)) & 0xff ? 1 : 0; f[4 + i] = (w1 >> (8 * i)) & 0xff ? 1 : 0; }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) f[i] = (base + i < n && flags[base + i]) ? 1 : 0;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) cnt += f[i];
  return cnt;
}

__global__ void __launch_bounds__(kThreads) compact_count_kernel(const uint8_t* __restrict__ flags, uint32_t n, uint32_t* __restrict__ sums) {
  __shared__ uint32_t s_warp[9];
  uint32_t nblocks = (n + kCompactTile - 1) / kCompactTile;
  for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    uint8_t f[8];
    uint32_t base = b * kCompactTile + threadIdx.x * kCompactPerThread;
    uint32_t c = base < n ? load_flags8(flags, base, n, f) : 0;
    uint32_t total;
    (void)block_excl_scan(c, &total, s_warp);
    if (threadIdx.x == 0) sums[b] = total;
  }
}

// single block: exclusive scan of sums[0..nb) in place; total -> *d_total
__global__ void __launch_bounds__(1024) compact_scan_sums_kernel(uint32_t* sums, uint32_t nb, uint32_t* d_total) {
  __shared__ uint32_t s_w[33];
  __shared__ uint32_t s_carry;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nb; base += 1024) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < nb ? sums[i] : 0;
    uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
      uint32_t x = s_w[lane];
      uint32_t xi = warp_incl_scan(x, lane);
      s_w[lane] = xi - x;
      if (lane == 31) s_w[32] = xi;
    }
    __syncthreads();
    uint32_t carry = s_carry;
    if (i < nb) sums[i] = carry + s_w[w] + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + s_w[32];
    __syncthreads();
  }
  if (threadIdx.x == 0) *d_total = s_carry;
}

__global__ void __launch_bounds__(kThreads) compact_write_kernel(const uint8_t* __restrict__ flags, uint32_t n,
                                                                const uint32_t* __restrict__ sums, uint32_t* __restrict__ out_idx) {
  __shared__ uint32_t s_warp[9];
  uint32_t nblocks = (n + kCompactTile - 1) / kCompactTile;
  for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
    uint8_t f[8];
    uint32_t base = b * kCompactTile + threadIdx.x * kCompactPerThread;
    uint32_t c = base < n ? load_flags8(flags, base, n, f) : 0;
    uint32_t total;
    uint32_t o = sums[b] + block_excl_scan(c, &total, s_warp);
    if (c) {
#pragma unroll
      for (int i = 0; i < 8; i++) if (f[i]) out_idx[o++] = base + i;
    }
  }
}

// ------------------------------------------------------------------------------------------------ S4: merge records
__global__ void survivor_run_starts_kernel(const uint32_t* __restrict__ surv, const uint32_t* d_m,
                                           const uint32_t* __restrict__ file_base, int k, uint32_t* run_start) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f > k) return;
  uint32_t m = *d_m;
  if (f == k) { run_start[k] = m; return; }
  uint32_t target = file_base[f];
  uint32_t lo = 0, hi = m;             // first survivor index with row >= file_base[f]
  if (surv == nullptr) lo = target < m ? target : m;
  else while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (surv[mid] < target) lo = mid + 1; else hi = mid; }
  run_start[f] = lo;
}

__device__ __forceinline__ void pk_key128(const PkSet& pk, uint32_t row, uint64_t* hi, uint64_t* lo) {
  unsigned __int128 key = 0;
  for (int c = 0; c < pk.n; c++) {
    uint32_t w = pk.c[c].width;
    uint64_t v = col_raw(pk.c[c], row);
    if (type_signed(pk.c[c].type)) v ^= (uint64_t(1) << (8 * w - 1));   // order-preserving map to unsigned
    key = (key << (8 * w)) | v;
  }
  *hi = uint64_t(key >> 64);
  *lo = uint64_t(key);
}

__global__ void __launch_bounds__(kThreads) build_records_kernel(PkSet pk, ColView seq, const uint32_t* __restrict__ surv,
                                                                const uint32_t* d_m, SortRec* __restrict__ rec) {
  uint32_t m = *d_m;
  for (uint32_t j = blockIdx.x * kThreads + threadIdx.x; j < m; j += gridDim.x * kThreads) {
    uint32_t row = surv ? surv[j] : j;
    SortRec r;
    pk_key128(pk, row, &r.k0, &r.k1);
    r.seq = col_valid(seq, row) ? col_raw(seq, row) + 1 : 0;   // ASC NULLS FIRST: null sorts before every value
    r.row = row;
    rec[j] = r;
  }
}

__device__ __forceinline__ bool rec_less(const SortRec& a, const SortRec& b) {
  if (a.k0 != b.k0) return a.k0 < b.k0;
  if (a.k1 != b.k1) return a.k1 < b.k1;
  if (a.seq != b.seq) return a.seq < b.seq;
  return a.row < b.row;        // ties -> lower stream index (rows are numbered file by file)
}

template <class GetA, class GetB>
__device__ __forceinline__ uint32_t merge_path(uint32_t diag, uint32_t na, uint32_t nb, GetA A, GetB B) {
  uint32_t lo = diag > nb ? diag - nb : 0, hi = diag < na ? diag : na;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    SortRec a = A(mid), b = B(diag - 1 - mid);
    if (!rec_less(b, a)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

constexpr int kMergeVT = 4;
constexpr int kMergeTile = kThreads * kMergeVT;  // 1024 records = 32 KB of shared memory

struct PairView { uint32_t a0, a1, b1; };
__device__ __forceinline__ PairView pair_of(const uint32_t* __restrict__ run_start, int k, int level, int nruns, uint32_t pos) {
  auto rstart = [&](int r) -> uint32_t { long idx = long(r) << level; return run_start[idx > k ? k : idx]; };
  int lo = 0, hi = nruns;        // upper_bound over rstart: pair containing pos = largest even r with rstart(r) <= pos
  while (lo < hi) { int mid = (lo + hi) >> 1; if (rstart(mid) <= pos) lo = mid + 1; else hi = mid; }
  int r = (lo - 1) & ~1;
  PairView p;
  p.a0 = rstart(r);
  p.a1 = rstart(r + 1 > nruns ? nruns : r + 1);
  p.b1 = rstart(r + 2 > nruns ? nruns : r + 2);
  return p;
}

// One thread per output tile: merge-path split (records taken from run A) at the tile's first output position.
// Thousands of independent binary searches overlap their latency instead of stalling every merge CTA.
__global__ void __launch_bounds__(kThreads) merge_partition_kernel(const SortRec* __restrict__ src, const uint32_t* __restrict__ run_start, int k,
                                                                  int level, const uint32_t* d_m, uint32_t* __restrict__ splits) {
  const uint32_t total = *d_m;
  const int nruns = (k + (1 << level) - 1) >> level;
  const uint32_t ntiles = (total + kMergeTile - 1) / kMergeTile;
  for (uint32_t t = blockIdx.x * kThreads + threadIdx.x; t < ntiles; t += gridDim.x * kThreads) {
    const uint32_t pos = t * kMergeTile;
    PairView p = pair_of(run_start, k, level, nruns, pos);
    const SortRec* A = src + p.a0;
    const SortRec* B = src + p.a1;
    splits[t] = merge_path(pos - p.a0, p.a1 - p.a0, p.b1 - p.a1, [&](uint32_t i) { return A[i]; }, [&](uint32_t i) { return B[i]; });
  }
}

__global__ void __launch_bounds__(kThreads) merge_pass_kernel(const SortRec* __restrict__ src, SortRec* __restrict__ dst,
                                                             const uint32_t* __restrict__ run_start, int k, int level,
                                                             const uint32_t* d_m, const uint32_t* __restrict__ splits) {
  __shared__ SortRec s_rec[kMergeTile];
  const uint32_t total = *d_m;
  const int tid = threadIdx.x;
  const int nruns = (k + (1 << level) - 1) >> level;   // runs at this level
  const uint32_t ntiles = (total + kMergeTile - 1) / kMergeTile;
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    uint32_t pos = tile * kMergeTile;
    const uint32_t tile_hi = pos + kMergeTile < total ? pos + kMergeTile : total;
    bool first = true;
    while (pos < tile_hi) {
      const PairView p = pair_of(run_start, k, level, nruns, pos);
      const uint32_t na = p.a1 - p.a0;
      const uint32_t seg_hi = tile_hi < p.b1 ? tile_hi : p.b1;
      const uint32_t d0 = pos - p.a0, d1 = seg_hi - p.a0;
      const SortRec* A = src + p.a0;
      const SortRec* B = src + p.a1;
      // splits: the tile's own start comes from the partition kernel; later segments start at a pair start (0);
      // a segment ends either at the pair end (na) or at the next tile's start (same pair)
      const uint32_t ia0 = first ? splits[tile] : 0u;
      const uint32_t ia1 = seg_hi == p.b1 ? na : splits[tile + 1];
      first = false;
      const uint32_t ib0 = d0 - ia0, ib1 = d1 - ia1;
      const uint32_t la = ia1 - ia0, lb = ib1 - ib0, n = la + lb;
      for (uint32_t i = tid; i < n; i += kThreads) s_rec[i] = i < la ? A[ia0 + i] : B[ib0 + (i - la)];
      __syncthreads();
      const uint32_t t0 = uint32_t(tid) * kMergeVT;
      if (t0 < n) {
        const SortRec* sA = s_rec;
        const SortRec* sB = s_rec + la;
        uint32_t ai = merge_path(t0, la, lb, [&](uint32_t i) { return sA[i]; }, [&](uint32_t i) { return sB[i]; });
        uint32_t bi = t0 - ai;
        SortRec out[kMergeVT];
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < kMergeVT; i++) {
          if (t0 + i >= n) break;
          bool takeA;
          if (ai >= la) takeA = false;
          else if (bi >= lb) takeA = true;
          else takeA = !rec_less(sB[bi], sA[ai]);
          out[i] = takeA ? sA[ai++] : sB[bi++];
          cnt++;
        }
        for (int i = 0; i < cnt; i++) dst[p.a0 + d0 + t0 + i] = out[i];
      }
      __syncthreads();
      pos = seg_hi;
    }
  }
}

__global__ void __launch_bounds__(kThreads) records_to_rows_kernel(const SortRec* __restrict__ rec, const uint32_t* d_m,
                                                                  uint32_t* __restrict__ order) {
  uint32_t m = *d_m;
  for (uint32_t j = blockIdx.x * kThreads + threadIdx.x; j < m; j += gridDim.x * kThreads) order[j] = uint32_t(rec[j].row);
}

// ------------------------------------------------------------------------------------------- S5/S6: PK-run ends
// primary_key_eq (read.rs:262-287) compares VALUES only (null bitmap ignored)
__device__ __forceinline__ bool pk_equal(const PkSet& pk, uint32_t a, uint32_t b) {
  for (int c = 0; c < pk.n; c++)
    if (col_raw(pk.c[c], a) != col_raw(pk.c[c], b)) return false;
  return true;
}

__global__ void __launch_bounds__(kThreads) dedup_flags_cols_kernel(PkSet pk, const uint32_t* __restrict__ order, const uint32_t* d_m,
                                                                   uint8_t* __restrict__ keep) {
  uint32_t m = *d_m;
  for (uint32_t j = blockIdx.x * kThreads + threadIdx.x; j < m; j += gridDim.x * kThreads) {
    bool last = true;
    if (j + 1 < m) {
      uint32_t a = order ? order[j] : j, b = order ? order[j + 1] : j + 1;
      last = !pk_equal(pk, a, b);
    }
    keep[j] = last ? 1 : 0;
  }
}

__global__ void __launch_bounds__(kThreads) dedup_flags_recs_kernel(const SortRec* __restrict__ rec, const uint32_t* d_m,
                                                                   uint8_t* __restrict__ keep) {
  uint32_t m = *d_m;
  for (uint32_t j = blockIdx.x * kThreads + threadIdx.x; j < m; j += gridDim.x * kThreads) {
    bool last = true;
    if (j + 1 < m) last = rec[j].k0 != rec[j + 1].k0 || rec[j].k1 != rec[j + 1].k1;
    keep[j] = last ? 1 : 0;
  }
}

__global__ void __launch_bounds__(kThreads) gather_rows_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ out_pos,
                                                              const uint32_t* d_r, uint32_t* __restrict__ out_rows) {
  uint32_t r = *d_r;
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < r; i += gridDim.x * kThreads)
    out_rows[i] = order ? order[out_pos[i]] : out_pos[i];
}

__global__ void batch_bounds_kernel(const uint32_t* __restrict__ out_pos, const uint32_t* d_r, const uint32_t* __restrict__ chunk_end,
                                    uint32_t nchunks, uint32_t* bound) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  uint32_t r = *d_r;
  uint32_t ce = chunk_end[c];
  uint32_t target = ce == 0 ? 0 : ce - 1;     // outputs with merged position < ce-1 belong to batches <= c
  uint32_t lo = 0, hi = r;
  while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (out_pos[mid] < target) lo = mid + 1; else hi = mid; }
  bound[c] = lo;
}

__global__ void chunk_ends_kernel(const uint32_t* __restrict__ surv, const uint32_t* d_m, const uint32_t* __restrict__ piece_end_row,
                                  uint32_t npieces, uint32_t* chunk_end) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= npieces) return;
  uint32_t m = *d_m, target = piece_end_row[c];
  uint32_t lo = 0, hi = m;
  if (surv == nullptr) lo = target < m ? target : m;
  else while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (surv[mid] < target) lo = mid + 1; else hi = mid; }
  chunk_end[c] = lo;
}

// ------------------------------------------------------------------------------------------- output materialisation
template <typename T>
__global__ void __launch_bounds__(kThreads) gather_column_kernel(const T* __restrict__ src, const uint8_t* __restrict__ src_valid,
                                                                const uint32_t* __restrict__ rows, const uint32_t* d_r,
                                                                T* __restrict__ dst, uint8_t* __restrict__ dst_valid) {
  uint32_t r = *d_r;
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < r; i += gridDim.x * kThreads) {
    uint32_t row = rows ? rows[i] : i;
    dst[i] = src[row];
    if (dst_valid) dst_valid[i] = src_valid ? src_valid[row] : 1;
  }
}

__global__ void __launch_bounds__(kThreads) pack_validity_kernel(const uint8_t* __restrict__ valid, uint32_t n, uint8_t* __restrict__ bitmap,
                                                                unsigned long long* null_count) {
  uint32_t nbytes = (n + 7) / 8;
  uint32_t nulls = 0;
  for (uint32_t b = blockIdx.x * kThreads + threadIdx.x; b < nbytes; b += gridDim.x * kThreads) {
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint32_t idx = b * 8 + i;
      if (idx < n) { if (valid[idx]) bits |= 1u << i; else nulls++; }
    }
    bitmap[b] = uint8_t(bits);
  }
  for (int d = 16; d > 0; d >>= 1) nulls += __shfl_down_sync(0xffffffffu, nulls, d);
  if ((threadIdx.x & 31) == 0 && nulls) atomicAdd(null_count, (unsigned long long)nulls);
}

// ------------------------------------------------------------------------------------------------ A1/A2: aggregation
__device__ __forceinline__ int64_t bucket_of(const AggSpecDev& s, uint32_t row) {
  int64_t ts = int64_t(col_widened(s.ts, row));
  return ts / s.window_ms * s.window_ms;          // truncating division == Timestamp::truncate_by (types.rs:82-85)
}

__global__ void __launch_bounds__(kThreads) group_flags_kernel(AggSpecDev spec, const uint32_t* __restrict__ rows, const uint32_t* d_r,
                                                              uint8_t* __restrict__ head) {
  uint32_t r = *d_r;
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < r; i += gridDim.x * kThreads) {
    bool h = i == 0;
    if (!h) {
      uint32_t a = rows ? rows[i - 1] : i - 1, b = rows ? rows[i] : i;
      if (spec.has_group && col_raw(spec.group, a) != col_raw(spec.group, b)) h = true;
      if (!h && spec.has_ts && bucket_of(spec, a) != bucket_of(spec, b)) h = true;
    }
    head[i] = h ? 1 : 0;
  }
}

__device__ __forceinline__ double value_as_double(const ColView& c, uint32_t row) {
  uint64_t w = col_widened(c, row);
  if (type_float(c.type)) return __longlong_as_double((long long)w);
  if (type_signed(c.type)) return double(int64_t(w));
  return double(w);
}

// One thread per group walks its rows in stream order: the f64 sum is a strictly sequential chain.
__global__ void __launch_bounds__(kThreads) reduce_groups_kernel(AggSpecDev spec, const uint32_t* __restrict__ rows, const uint32_t* d_r,
                                                                const uint32_t* __restrict__ seg_start, const uint32_t* d_g, AggOut out) {
  uint32_t g_total = *d_g, r_total = *d_r;
  for (uint32_t g = blockIdx.x * kThreads + threadIdx.x; g < g_total; g += gridDim.x * kThreads) {
    uint32_t lo = seg_start[g], hi = g + 1 < g_total ? seg_start[g + 1] : r_total;
    uint32_t first = rows ? rows[lo] : lo;
    if (out.gkey) {
      uint64_t kv = spec.has_group ? col_raw(spec.group, first) : 0;
      store_val_dyn(out.gkey, spec.has_group ? spec.group.width : 8, g, kv);
    }
    out.bucket[g] = spec.has_ts ? bucket_of(spec, first) : 0;
    out.count[g] = hi - lo;
    double sum = 0.0, mn = __longlong_as_double(0x7ff0000000000000LL), mx = __longlong_as_double((long long)0xfff0000000000000ULL);
    if (spec.has_value) {
      bool seen = false;
      for (uint32_t i = lo; i < hi; i++) {
        uint32_t row = rows ? rows[i] : i;
        if (!col_valid(spec.value, row)) continue;
        double v = value_as_double(spec.value, row);
        sum += v;
        if (!seen || v < mn) mn = v;
        if (!seen || v > mx) mx = v;
        seen = true;
      }
    }
    out.sum[g] = sum;
    out.min[g] = mn;
    out.max[g] = mx;
  }
}

__global__ void pack_agg_kernel(AggOut in, uint32_t gwidth, uint64_t g, uint64_t cap, long long* __restrict__ dst) {
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < cap; i += uint64_t(gridDim.x) * blockDim.x) {
    const bool v = i < g;
    long long key = 0;
    if (v) {
      switch (gwidth) {
        case 1: key = reinterpret_cast<const uint8_t*>(in.gkey)[i]; break;
        case 2: key = reinterpret_cast<const uint16_t*>(in.gkey)[i]; break;
        case 4: key = reinterpret_cast<const uint32_t*>(in.gkey)[i]; break;
        default: key = (long long)reinterpret_cast<const uint64_t*>(in.gkey)[i];
      }
    }
    dst[i] = key;
    dst[cap + i] = v ? in.bucket[i] : 0;
    dst[2 * cap + i] = v ? (long long)in.count[i] : 0;
    dst[3 * cap + i] = v ? __double_as_longlong(in.sum[i]) : 0;
    dst[4 * cap + i] = v ? __double_as_longlong(in.min[i]) : 0;
    dst[5 * cap + i] = v ? __double_as_longlong(in.max[i]) : 0;
  }
}

// ------------------------------------------------------------------------------------------- Binary columns: export
// byte length of the i-th output row (0 for NULL and beyond the count), for an exclusive scan -> Arrow offsets
__global__ void __launch_bounds__(kThreads) gather_lens_kernel(ColView col, const uint32_t* __restrict__ rows, const uint32_t* d_n, uint32_t cap,
                                                              uint32_t* __restrict__ out) {
  const uint32_t n = *d_n;
  for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < cap; i += gridDim.x * kThreads) {
    uint32_t len = 0;
    if (i < n) { const uint32_t row = rows ? rows[i] : i; if (col.valid == nullptr || col.valid[row]) len = col.lens[row]; }
    out[i] = len;
  }
}
// one warp per row: bytes of row rows[i] -> dst + offs[i]
__global__ void __launch_bounds__(kThreads) copy_var_kernel(ColView col, const uint32_t* __restrict__ rows, const uint32_t* d_n,
                                                           const uint32_t* __restrict__ offs, uint8_t* __restrict__ dst) {
  const uint32_t n = *d_n;
  const int lane = threadIdx.x & 31;
  const uint32_t nwarps = gridDim.x * (kThreads / 32);
  for (uint32_t i = (blockIdx.x * kThreads + threadIdx.x) >> 5; i < n; i += nwarps) {
    const uint32_t row = rows ? rows[i] : i;
    if (col.valid && !col.valid[row]) continue;
    const uint8_t* src = reinterpret_cast<const uint8_t* const*>(col.vals)[row];
    const uint32_t len = col.lens[row];
    uint8_t* d = dst + offs[i];
    for (uint32_t b = lane; b < len; b += 32) d[b] = src[b];
  }
}
// Append mode: first row of the j-th primary-key run in merged order (the run ends at merged position out_pos[j])
__global__ void __launch_bounds__(kThreads) first_rows_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ out_pos, const uint32_t* d_r,
                                                             uint32_t* __restrict__ out) {
  const uint32_t r = *d_r;
  for (uint32_t j = blockIdx.x * kThreads + threadIdx.x; j < r; j += gridDim.x * kThreads) {
    const uint32_t pos = j ? out_pos[j - 1] + 1 : 0;
    out[j] = order ? order[pos] : pos;
  }
}
// Append mode: Arrow offsets of the concatenated values = the running byte count sampled at the runs' first rows (+ the total)
__global__ void __launch_bounds__(kThreads) run_offsets_kernel(const uint32_t* __restrict__ cum, const uint32_t* __restrict__ out_pos, const uint32_t* d_r,
                                                              const uint32_t* d_m, uint32_t* __restrict__ out) {
  const uint32_t r = *d_r, m = *d_m;
  for (uint32_t j = blockIdx.x * kThreads + threadIdx.x; j <= r; j += gridDim.x * kThreads)
    out[j] = j == r ? cum[m] : cum[j ? out_pos[j - 1] + 1 : 0];
}

__global__ void __launch_bounds__(kThreads) append_validity_kernel(ColView col, const uint32_t* __restrict__ order, const uint32_t* __restrict__ out_pos,
                                                                  const uint32_t* d_r, const uint32_t* __restrict__ run_offs, uint8_t* __restrict__ valid, int* err) {
  const uint32_t r = *d_r;
  for (uint32_t j = blockIdx.x * kThreads + threadIdx.x; j < r; j += gridDim.x * kThreads) {
    uint8_t v = 1;
    if (run_offs[j + 1] == run_offs[j]) {
      const uint32_t first = j ? out_pos[j - 1] + 1 : 0, last = out_pos[j];
      if (first == last) { const uint32_t row = order ? order[first] : first; v = (col.valid == nullptr || col.valid[row]) ? 1 : 0; }
      else atomicExch(err, 130);
    }
    valid[j] = v;
  }
}

__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}

__global__ void uniform_chunk_ends_kernel(const uint32_t* d_m, uint32_t batch, uint32_t nchunks, uint32_t* chunk_end) {
  uint32_t m = *d_m;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += gridDim.x * blockDim.x) {
    uint64_t e = uint64_t(c + 1) * batch;
    chunk_end[c] = e < m ? uint32_t(e) : m;
  }
}
__global__ void clear_tail_kernel(uint8_t* flags, const uint32_t* d_n, uint32_t cap) {
  uint32_t n = *d_n;
  for (uint64_t i = uint64_t(n) + blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += uint64_t(gridDim.x) * blockDim.x) flags[i] = 0;
}

}  // namespace

// =================================================================================================== launch wrappers
void uniform_chunk_ends(const Launch& L, const uint32_t* d_m, uint32_t batch, uint32_t nchunks, uint32_t* chunk_end) {
  if (!nchunks) return;
  uniform_chunk_ends_kernel<<<grid_for(nchunks), kThreads, 0, L.stream>>>(d_m, batch, nchunks, chunk_end);
  L.tick();
}
void clear_tail(const Launch& L, uint8_t* flags, const uint32_t* d_n, uint32_t cap) {
  if (!cap) return;
  clear_tail_kernel<<<grid_for(cap), kThreads, 0, L.stream>>>(flags, d_n, cap);
  L.tick();
}
void snappy_chunks(const Launch& L, const SstDev* ssts, const RgSel* sel, uint32_t nsel, const ColSel* cols, int ncolsel,
                   uint8_t* scratch, int* err) {
  if (!nsel || !ncolsel) return;
  snappy_chunks_kernel<<<nsel * ncolsel, 32, 0, L.stream>>>(ssts, sel, cols, ncolsel, scratch, err);
  L.tick();
}
void decode_chunks(const Launch& L, const SstDev* ssts, const RgSel* sel, uint32_t nsel, const ColSel* cols, int ncolsel,
                   uint8_t* scratch, int* err) {
  if (!nsel || !ncolsel) return;
  decode_chunks_kernel<<<nsel * ncolsel, kThreads, 0, L.stream>>>(ssts, sel, cols, ncolsel, scratch, err);
  L.tick();
}
void eval_predicates(const Launch& L, const PredSet& preds, uint32_t n, uint8_t* alive) {
  if (!n) return;
  eval_predicates_kernel<<<grid_for(n), kThreads, 0, L.stream>>>(preds, n, alive);
  L.tick();
}
size_t compact_tmp_elems(uint32_t n) { return size_t(n) / kCompactTile + 2; }
void compact_flags(const Launch& L, const uint8_t* flags, uint32_t n, uint32_t* tmp, uint32_t* out_idx, uint32_t* d_total) {
  uint32_t nb = (n + kCompactTile - 1) / kCompactTile;
  if (n) { compact_count_kernel<<<grid_for(nb, 1), kThreads, 0, L.stream>>>(flags, n, tmp); L.tick(); }
  compact_scan_sums_kernel<<<1, 1024, 0, L.stream>>>(tmp, nb, d_total);
  L.tick();
  if (n) { compact_write_kernel<<<grid_for(nb, 1), kThreads, 0, L.stream>>>(flags, n, tmp, out_idx); L.tick(); }
}
void survivor_run_starts(const Launch& L, const uint32_t* surv, const uint32_t* d_m, const uint32_t* file_base, int k,
                         uint32_t* run_start) {
  survivor_run_starts_kernel<<<(k + 1 + 127) / 128, 128, 0, L.stream>>>(surv, d_m, file_base, k, run_start);
  L.tick();
}
void build_records(const Launch& L, const PkSet& pk, ColView seq, const uint32_t* surv, const uint32_t* d_m, uint32_t cap,
                   SortRec* rec) {
  if (!cap) return;
  build_records_kernel<<<grid_for(cap), kThreads, 0, L.stream>>>(pk, seq, surv, d_m, rec);
  L.tick();
}
size_t merge_split_elems(uint32_t cap) { return size_t(cap) / kMergeTile + 2; }
void merge_pass(const Launch& L, const SortRec* src, SortRec* dst, const uint32_t* run_start, int k, int level,
                const uint32_t* d_m, uint32_t cap, uint32_t* splits) {
  if (!cap) return;
  const uint32_t ntiles = (cap + kMergeTile - 1) / kMergeTile;
  merge_partition_kernel<<<grid_for(ntiles), kThreads, 0, L.stream>>>(src, run_start, k, level, d_m, splits);
  L.tick();
  merge_pass_kernel<<<grid_for(cap, kMergeTile, kSMs * 8), kThreads, 0, L.stream>>>(src, dst, run_start, k, level, d_m, splits);
  L.tick();
}
void records_to_rows(const Launch& L, const SortRec* rec, const uint32_t* d_m, uint32_t cap, uint32_t* order) {
  if (!cap) return;
  records_to_rows_kernel<<<grid_for(cap), kThreads, 0, L.stream>>>(rec, d_m, order);
  L.tick();
}
void dedup_flags_cols(const Launch& L, const PkSet& pk, const uint32_t* order, const uint32_t* d_m, uint32_t cap, uint8_t* keep) {
  if (!cap) return;
  dedup_flags_cols_kernel<<<grid_for(cap), kThreads, 0, L.stream>>>(pk, order, d_m, keep);
  L.tick();
}
void dedup_flags_recs(const Launch& L, const SortRec* rec, const uint32_t* d_m, uint32_t cap, uint8_t* keep) {
  if (!cap) return;
  dedup_flags_recs_kernel<<<grid_for(cap), kThreads, 0, L.stream>>>(rec, d_m, keep);
  L.tick();
}
void gather_rows(const Launch& L, const uint32_t* order, const uint32_t* out_pos, const uint32_t* d_r, uint32_t cap,
                 uint32_t* out_rows) {
  if (!cap) return;
  gather_rows_kernel<<<grid_for(cap), kThreads, 0, L.stream>>>(order, out_pos, d_r, out_rows);
  L.tick();
}
void batch_bounds(const Launch& L, const uint32_t* out_pos, const uint32_t* d_r, const uint32_t* chunk_end, uint32_t nchunks,
                  uint32_t* bound) {
  if (!nchunks) return;
  batch_bounds_kernel<<<(nchunks + 127) / 128, 128, 0, L.stream>>>(out_pos, d_r, chunk_end, nchunks, bound);
  L.tick();
}
void chunk_ends_from_rows(const Launch& L, const uint32_t* surv, const uint32_t* d_m, const uint32_t* piece_end_row,
                          uint32_t npieces, uint32_t* chunk_end) {
  if (!npieces) return;
  chunk_ends_kernel<<<(npieces + 127) / 128, 128, 0, L.stream>>>(surv, d_m, piece_end_row, npieces, chunk_end);
  L.tick();
}
void gather_column(const Launch& L, ColView src, const uint32_t* rows, const uint32_t* d_r, uint32_t cap, void* dst_vals,
                   uint8_t* dst_valid) {
  if (!cap) return;
  int g = grid_for(cap);
  switch (src.width) {
    case 1: gather_column_kernel<uint8_t><<<g, kThreads, 0, L.stream>>>((const uint8_t*)src.vals, src.valid, rows, d_r, (uint8_t*)dst_vals, dst_valid); break;
    case 2: gather_column_kernel<uint16_t><<<g, kThreads, 0, L.stream>>>((const uint16_t*)src.vals, src.valid, rows, d_r, (uint16_t*)dst_vals, dst_valid); break;
    case 4: gather_column_kernel<uint32_t><<<g, kThreads, 0, L.stream>>>((const uint32_t*)src.vals, src.valid, rows, d_r, (uint32_t*)dst_vals, dst_valid); break;
    default: gather_column_kernel<uint64_t><<<g, kThreads, 0, L.stream>>>((const uint64_t*)src.vals, src.valid, rows, d_r, (uint64_t*)dst_vals, dst_valid);
  }
  L.tick();
}
void pack_validity(const Launch& L, const uint8_t* valid_bytes, uint32_t n, uint8_t* bitmap, unsigned long long* null_count) {
  if (!n) return;
  pack_validity_kernel<<<grid_for((n + 7) / 8), kThreads, 0, L.stream>>>(valid_bytes, n, bitmap, null_count);
  L.tick();
}
void group_flags(const Launch& L, const AggSpecDev& spec, const uint32_t* rows, const uint32_t* d_r, uint32_t cap, uint8_t* head) {
  if (!cap) return;
  group_flags_kernel<<<grid_for(cap), kThreads, 0, L.stream>>>(spec, rows, d_r, head);
  L.tick();
}
void reduce_groups(const Launch& L, const AggSpecDev& spec, const uint32_t* rows, const uint32_t* d_r, const uint32_t* seg_start,
                   const uint32_t* d_g, uint32_t cap, AggOut out) {
  if (!cap) return;
  reduce_groups_kernel<<<grid_for(cap), kThreads, 0, L.stream>>>(spec, rows, d_r, seg_start, d_g, out);
  L.tick();
}
void pack_agg(const Launch& L, AggOut in, uint32_t gwidth, uint64_t g, uint64_t cap, long long* dst) {
  if (!cap) return;
  pack_agg_kernel<<<grid_for(cap), kThreads, 0, L.stream>>>(in, gwidth, g, cap, dst);
  L.tick();
}
void gather_lens(const Launch& L, ColView col, const uint32_t* rows, const uint32_t* d_n, uint32_t cap, uint32_t* out) {
  if (!cap) return;
  gather_lens_kernel<<<grid_for(cap), kThreads, 0, L.stream>>>(col, rows, d_n, cap, out);
  L.tick();
}
void copy_var(const Launch& L, ColView col, const uint32_t* rows, const uint32_t* d_n, uint32_t cap, const uint32_t* offs, uint8_t* dst) {
  if (!cap) return;
  copy_var_kernel<<<grid_for(uint64_t(cap) * 32), kThreads, 0, L.stream>>>(col, rows, d_n, offs, dst);
  L.tick();
}
void first_rows(const Launch& L, const uint32_t* order, const uint32_t* out_pos, const uint32_t* d_r, uint32_t cap, uint32_t* out) {
  if (!cap) return;
  first_rows_kernel<<<grid_for(cap), kThreads, 0, L.stream>>>(order, out_pos, d_r, out);
  L.tick();
}
void run_offsets(const Launch& L, const uint32_t* cum, const uint32_t* out_pos, const uint32_t* d_r, const uint32_t* d_m, uint32_t cap, uint32_t* out) {
  run_offsets_kernel<<<grid_for(uint64_t(cap) + 1), kThreads, 0, L.stream>>>(cum, out_pos, d_r, d_m, out);
  L.tick();
}
void append_validity(const Launch& L, ColView col, const uint32_t* order, const uint32_t* out_pos, const uint32_t* d_r, const uint32_t* run_offs,
                     uint32_t cap, uint8_t* valid, int* err) {
  if (!cap) return;
  append_validity_kernel<<<grid_for(cap), kThreads, 0, L.stream>>>(col, order, out_pos, d_r, run_offs, valid, err);
  L.tick();
}
void exclusive_scan_u32(const Launch& L, uint32_t* data, uint32_t n, uint32_t* d_total) {
  compact_scan_sums_kernel<<<1, 1024, 0, L.stream>>>(data, n, d_total);
  L.tick();
}
void fill_u32(const Launch& L, uint32_t* p, uint32_t v, uint32_t n) {
  if (!n) return;
  fill_u32_kernel<<<grid_for(n), kThreads, 0, L.stream>>>(p, v, n);
  L.tick();
}

}  // namespace k
}  // namespace horae
