// snappy.cu — raw-Snappy page decompression (Parquet SNAPPY codec: parquet 53.2 -> snap 1.1.1 in the reference,
// Cargo.lock:3145; format restated from the published Snappy format description).
//
// Snappy is byte-serial by definition: an element's position depends on all element lengths before it, and a copy may
// read bytes produced by the element just before it.  One warp owns one page and breaks both dependencies:
//
//   parse    a 256-byte window of the compressed stream is staged in shared memory; every byte position computes
//            "where would the next element start if one started here" (J1, a 256-entry LUT on the tag byte), four
//            doubling steps give J2..J16, and lane k finds the start of the k-th element with 5 dependent lookups
//            (binary lifting).  32 elements are decoded per step instead of one.
//   execute  tiny elements (columns of 8-byte values compress to ~2.5-byte elements: literal(2) + copy(6, offset 8)):
//            every output byte of the batch gets a source pointer (literal byte / earlier output byte); chains are
//            collapsed by pointer jumping in shared memory (log2 of the chain length rounds), then the batch is
//            written out.  Long elements (run-length-like columns: 64-byte copies at offset 8): one element per step,
//            spread over the 32 lanes, reading the last kHist bytes of output from shared memory instead of waiting
//            for the previous element's global stores.
//   literals longer than 60 bytes (incompressible columns are one literal per 64 KiB block) are plain warp copies.
#include "kernels.h"

#include <cstring>

namespace horae {
namespace k {

namespace {

constexpr int kWin = 256;          // compressed-stream window (bytes)
constexpr int kRing = 4096;        // ring buffer of the most recent output (power of two)
constexpr int kOut = 2048;         // max output of one batch: 32 elements x 64 bytes
constexpr int kTiny = 512;         // max output of a tiny-element batch (byte-level resolve)
constexpr int kHist = kRing - kOut;  // bytes before the current batch that are guaranteed to still be in the ring
constexpr int kLevels = 6;         // J1, J2, J4, J8, J16, J32
constexpr uint16_t kExit = 0xffff;
constexpr uint16_t kDone = 0xffff;
constexpr int kWarpsPerCta = 4;

struct alignas(16) WarpSmem {
  uint8_t win[kWin + 16];
  uint16_t J[kLevels][kWin];
  uint8_t ring[kRing];             // output byte at absolute position x lives at ring[x & (kRing-1)]
  uint16_t ptr[kTiny];             // tiny-element batches only (T <= 32*16): batch-relative source index, or kDone
};

__host__ __device__ __forceinline__ uint64_t page_scratch2(uint32_t uncomp) { return (uint64_t(uncomp) + 15u) / 16u * 16u + 32u; }

__device__ __forceinline__ uint64_t ld8_any(const uint8_t* p) {
  uintptr_t a = reinterpret_cast<uintptr_t>(p);
  uint32_t sh = uint32_t(a & 7) * 8;
  const uint64_t* q = reinterpret_cast<const uint64_t*>(a & ~uintptr_t(7));
  uint64_t lo = __ldg(q), hi = __ldg(q + 1);
  return (lo >> sh) | ((hi << 1) << (63 - sh));
}
__device__ __forceinline__ uint8_t ldcg_u8(const uint8_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return uint8_t(v);
}

// compressed size of the element whose tag byte is t; 0 = literal with a multi-byte length (handled separately)
__device__ __forceinline__ uint32_t elem_csize(uint32_t t) {
  uint32_t kind = t & 3;
  if (kind == 0) { uint32_t l = t >> 2; return l < 60 ? l + 2 : 0; }
  return kind == 1 ? 2u : (kind == 2 ? 3u : 5u);
}

// plain copy global->global spread over the warp (source is read-only input)
__device__ __forceinline__ void warp_copy_in(uint8_t* dst, const uint8_t* src, uint32_t len, int lane) {
  uint32_t head = uint32_t((8 - (reinterpret_cast<uintptr_t>(dst) & 7)) & 7);
  if (head > len) head = len;
  if (uint32_t(lane) < head) dst[lane] = __ldg(src + lane);
  uint32_t nwords = (len - head) >> 3;
  uint64_t* d8 = reinterpret_cast<uint64_t*>(dst + head);
  const uint8_t* s = src + head;
  for (uint32_t w = lane; w < nwords; w += 32) d8[w] = ld8_any(s + (size_t(w) << 3));
  uint32_t done = head + (nwords << 3);
  for (uint32_t i = done + lane; i < len; i += 32) dst[i] = __ldg(src + i);
}

// byte at absolute output position x (< o, i.e. produced by an earlier batch): ring if recent enough, else global
__device__ __forceinline__ uint8_t old_byte(const WarpSmem& sm, const uint8_t* dst, uint32_t o, uint32_t x) {
  return (o - x <= uint32_t(kHist)) ? sm.ring[x & (kRing - 1)] : ldcg_u8(dst + x);
}

// pointer jumping over at most kTrips*32 batch bytes, all state of a round staged in registers
template <int kTrips>
__device__ __forceinline__ void resolve_small(WarpSmem& sm, uint32_t o, uint32_t T, int lane) {
  for (;;) {
    uint16_t pr[kTrips], pq[kTrips];
    uint8_t vq[kTrips];
#pragma unroll
    for (int j = 0; j < kTrips; j++) {
      const uint32_t r = j * 32 + lane;
      pr[j] = kDone; pq[j] = kDone; vq[j] = 0;
      if (r < T) {
        pr[j] = sm.ptr[r];
        if (pr[j] != kDone) { pq[j] = sm.ptr[pr[j]]; vq[j] = sm.ring[(o + pr[j]) & (kRing - 1)]; }
      }
    }
    __syncwarp();
    bool pending = false;
#pragma unroll
    for (int j = 0; j < kTrips; j++) {
      const uint32_t r = j * 32 + lane;
      if (r < T && pr[j] != kDone) {
        if (pq[j] == kDone) { sm.ring[(o + r) & (kRing - 1)] = vq[j]; sm.ptr[r] = kDone; }
        else { sm.ptr[r] = pq[j]; pending = true; }
      }
    }
    __syncwarp();
    if (!__any_sync(0xffffffffu, pending)) break;
  }
}

__device__ void snappy_page(const uint8_t* __restrict__ src, uint32_t n, uint8_t* __restrict__ dst, uint32_t ulen_expected, WarpSmem& sm,
                            int lane, int* err) {
  uint32_t pos = 0, ulen = 0;
  for (int sh = 0; pos < n && sh < 35; sh += 7) {
    uint32_t b = __ldg(src + pos++);
    ulen |= (b & 0x7f) << sh;
    if (!(b & 0x80)) break;
  }
  if (ulen != ulen_expected) { if (lane == 0) atomicExch(err, 101); return; }
  uint32_t o = 0;                 // bytes produced so far
  while (pos < n) {
    const uint32_t avail = n - pos;
    const uint32_t tag0 = __ldg(src + pos);
    // ---- literal with an explicit length field: straight copy
    if ((tag0 & 3) == 0 && (tag0 >> 2) >= 60) {
      uint32_t nb = (tag0 >> 2) - 59, len = 0;
      for (uint32_t i = 0; i < nb; i++) len |= uint32_t(__ldg(src + pos + 1 + i)) << (8 * i);
      len += 1;
      if (1 + nb + len > avail || o + len > ulen) { if (lane == 0) atomicExch(err, 102); return; }
      const uint8_t* lsrc = src + pos + 1 + nb;
      warp_copy_in(dst + o, lsrc, len, lane);
      __syncwarp();
      // the ring keeps the tail of the literal
      const uint32_t keep = len < uint32_t(kRing) ? len : uint32_t(kRing);
      for (uint32_t i = lane; i < keep; i += 32) sm.ring[(o + len - keep + i) & (kRing - 1)] = __ldg(lsrc + (len - keep) + i);
      __syncwarp();
      pos += 1 + nb + len;
      o += len;
      continue;
    }
    // ---- stage the window and build the jump tables
    __syncwarp();
    {
      uint64_t w = (uint32_t(lane) * 8 < avail + 8) ? ld8_any(src + pos + lane * 8) : 0ull;
      *reinterpret_cast<uint64_t*>(&sm.win[lane * 8]) = w;
      if (lane < 2) *reinterpret_cast<uint64_t*>(&sm.win[kWin + lane * 8]) = 0ull;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint32_t p = lane * 8 + j;
      const uint32_t sz = elem_csize(sm.win[p]);
      const uint32_t nx = p + sz;
      // the NEXT element must start inside the stream and have its (<= 5 byte) header inside the window
      sm.J[0][p] = (sz == 0 || p >= avail || nx + 5 > uint32_t(kWin) || nx >= avail) ? kExit : uint16_t(nx);
    }
    __syncwarp();
#pragma unroll
    for (int lv = 1; lv < kLevels; lv++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t p = lane * 8 + j;
        const uint16_t a = sm.J[lv - 1][p];
        sm.J[lv][p] = a == kExit ? kExit : sm.J[lv - 1][a];
      }
      __syncwarp();
    }
    uint32_t q = 0;                                  // window-relative start of element `lane` of the current batch
#pragma unroll
    for (int lv = 0; lv < 5; lv++)
      if ((lane >> lv) & 1) q = q == kExit ? uint32_t(kExit) : sm.J[lv][q];
    uint32_t next_pos = pos;
    // up to three batches of 32 elements from one parsed window
    for (int b = 0; b < 3; b++) {
      if (b > 0) q = q == kExit ? uint32_t(kExit) : sm.J[5][q];
      bool valid = q != kExit;
      uint32_t len = 0, off = 0, hdr = 0, csz = 0;
      bool is_lit = false;
      if (valid) {
        const uint32_t t = sm.win[q];
        const uint32_t kind = t & 3;
        if (kind == 0) {
          if ((t >> 2) >= 60) valid = false;         // long literal: ends the batch, handled by the straight-copy path
          else { is_lit = true; len = (t >> 2) + 1; hdr = 1; }
        } else if (kind == 1) { len = ((t >> 2) & 7) + 4; off = ((t >> 5) << 8) | sm.win[q + 1]; hdr = 2; }
        else if (kind == 2) { len = (t >> 2) + 1; off = uint32_t(sm.win[q + 1]) | (uint32_t(sm.win[q + 2]) << 8); hdr = 3; }
        else { len = (t >> 2) + 1; off = uint32_t(sm.win[q + 1]) | (uint32_t(sm.win[q + 2]) << 8) | (uint32_t(sm.win[q + 3]) << 16) | (uint32_t(sm.win[q + 4]) << 24); hdr = 5; }
        csz = hdr + (is_lit ? len : 0);
        if (valid && q + csz > avail) valid = false; // truncated stream: caught by the final size check / m == 0
      }
      const unsigned vm = __ballot_sync(0xffffffffu, valid);
      const int m = (vm == 0xffffffffu) ? 32 : (__ffs(~vm) - 1);   // valid lanes form a prefix
      if (m == 0) {
        if (b == 0) { if (lane == 0) atomicExch(err, 105); return; }
        break;
      }
      if (lane >= m) { len = 0; csz = 0; }
      uint32_t inc = len;                                          // exclusive prefix sum of the output lengths
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
      const uint32_t doff = inc - len;
      const uint32_t T = __shfl_sync(0xffffffffu, inc, 31);
      next_pos = pos + __shfl_sync(0xffffffffu, q + csz, m - 1);
      bool bad = lane < m && !is_lit && (off == 0 || off > o + doff);
      if (__any_sync(0xffffffffu, bad) || o + T > ulen) { if (lane == 0) atomicExch(err, 103); return; }

      if (T <= uint32_t(m) * 16) {
        // ---------------- tiny elements.
        // Fast path: every copy either reads bytes older than the batch or reproduces exactly one earlier element of
        // the batch (same start, same length) — the shape of fixed-width numeric columns.  Then the dependency graph is
        // over ELEMENTS: parent links are collapsed with five register shuffles and each lane copies its own bytes.
        bool elem_done = false;
        {
          uint8_t* tbl = reinterpret_cast<uint8_t*>(sm.ptr);            // start offset -> lane (no init: verified below)
          if (lane < m) tbl[doff] = uint8_t(lane);
          __syncwarp();
          int parent = lane;
          bool fail = false, need_parent = false;
          uint32_t dkind = 0, dpos = 0;                                  // 0: bytes at win[dpos..], 1: output bytes at abs dpos..
          if (lane < m) {
            if (is_lit) { dpos = q + 1; }
            else {
              const int32_t s0 = int32_t(doff) - int32_t(off);
              if (s0 < 0) { if (s0 + int32_t(len) <= 0) { dkind = 1; dpos = o + doff - off; } else fail = true; }
              else { parent = tbl[s0] & 31; need_parent = true; }
            }
          }
          const uint32_t pd = __shfl_sync(0xffffffffu, doff, parent), pl = __shfl_sync(0xffffffffu, len, parent);
          if (need_parent && (parent >= lane || pd + off != doff || pl != len)) fail = true;   // (a stale table entry may even name this lane)
          if (!__any_sync(0xffffffffu, fail)) {
#pragma unroll
            for (int it = 0; it < 5; it++) parent = __shfl_sync(0xffffffffu, parent, parent);
            const uint32_t rk = __shfl_sync(0xffffffffu, dkind, parent), rp = __shfl_sync(0xffffffffu, dpos, parent);
            if (lane < m) {
              for (uint32_t i = 0; i < len; i++) {
                uint8_t v;
                if (rk == 0) { const uint32_t wp = rp + i; v = wp < uint32_t(kWin) ? sm.win[wp] : __ldg(src + pos + wp); }
                else v = old_byte(sm, dst, o, rp + i);
                sm.ring[(o + doff + i) & (kRing - 1)] = v;
                dst[o + doff + i] = v;
              }
            }
            elem_done = true;
          }
          __syncwarp();
        }
        if (!elem_done) {
        // general case: byte-level source pointers + pointer jumping
        if (lane < m) {
          if (is_lit) {
            const uint8_t* ls = src + pos + q + 1;
            for (uint32_t i = 0; i < len; i++) {
              const uint32_t wp = q + 1 + i;                           // literal bytes usually sit in the staged window
              sm.ring[(o + doff + i) & (kRing - 1)] = wp < uint32_t(kWin) ? sm.win[wp] : __ldg(ls + i);
              sm.ptr[doff + i] = kDone;
            }
          } else {
            for (uint32_t i = 0; i < len; i++) {
              const uint32_t r = doff + i;
              if (off <= r) sm.ptr[r] = uint16_t(r - off);             // produced by this batch: resolved below
              else { sm.ring[(o + r) & (kRing - 1)] = old_byte(sm, dst, o, o + r - off); sm.ptr[r] = kDone; }
            }
          }
        }
        __syncwarp();
        if (T <= 256) {
          // register-staged rounds: one barrier between the read and the write phase
          if (T <= 64) resolve_small<2>(sm, o, T, lane);
          else if (T <= 128) resolve_small<4>(sm, o, T, lane);
          else resolve_small<8>(sm, o, T, lane);
        } else {
          const uint32_t trips = (T + 31) / 32;
          for (;;) {
            bool pending = false;
            for (uint32_t j = 0; j < trips; j++) {
              const uint32_t r = j * 32 + lane;
              uint16_t pr = kDone, pq = kDone;
              uint8_t vq = 0;
              if (r < T) {
                pr = sm.ptr[r];
                if (pr != kDone) { pq = sm.ptr[pr]; vq = sm.ring[(o + pr) & (kRing - 1)]; }
              }
              __syncwarp();
              if (r < T && pr != kDone) {
                if (pq == kDone) { sm.ring[(o + r) & (kRing - 1)] = vq; sm.ptr[r] = kDone; }
                else { sm.ptr[r] = pq; pending = true; }
              }
              __syncwarp();
            }
            if (!__any_sync(0xffffffffu, pending)) break;
          }
        }
        for (uint32_t r = lane; r < T; r += 32) dst[o + r] = sm.ring[(o + r) & (kRing - 1)];
        }
      } else {
        // ---------------- long elements.  Adjacent copies with the same offset continue one periodic pattern
        // (out[x] = out[x - off] over the union), so they are merged into a single run and spread over the warp.
        const uint32_t poff = __shfl_up_sync(0xffffffffu, off, 1);
        const bool plit = __shfl_up_sync(0xffffffffu, int(is_lit), 1) != 0;
        const bool head = lane < m && (lane == 0 || is_lit || plit || poff != off);
        unsigned heads = __ballot_sync(0xffffffffu, head);
        while (heads) {
          const int e = __ffs(heads) - 1;
          heads &= heads - 1;
          const int enext = heads ? (__ffs(heads) - 1) : m;            // first lane of the next run
          const uint32_t s0 = __shfl_sync(0xffffffffu, doff, e);
          const uint32_t s1 = enext < 32 ? __shfl_sync(0xffffffffu, doff, enext & 31) : T;
          const uint32_t rlen = (enext == m ? T : s1) - s0;
          const uint32_t eoff = __shfl_sync(0xffffffffu, off, e);
          const uint32_t eq = __shfl_sync(0xffffffffu, q, e);
          const bool elit = __shfl_sync(0xffffffffu, int(is_lit), e) != 0;
          __syncwarp();
          if (elit) {
            const uint8_t* ls = src + pos + eq + 1;
            for (uint32_t i = lane; i < rlen; i += 32) sm.ring[(o + s0 + i) & (kRing - 1)] = __ldg(ls + i);
          } else {
            // every source byte precedes the run: x - off taken modulo the pattern length (kept incrementally: no
            // integer division per byte)
            const uint32_t start = o + s0;
            const uint32_t stride = 32u % eoff;
            uint32_t r = uint32_t(lane) % eoff;
            const bool all_ring = eoff <= uint32_t(kHist);                 // the whole pattern is inside the ring window
            for (uint32_t i = lane; i < rlen; i += 32) {
              const uint32_t x = start - eoff + r;
              sm.ring[(start + i) & (kRing - 1)] = (all_ring || start - x <= uint32_t(kHist) + s0) ? sm.ring[x & (kRing - 1)] : ldcg_u8(dst + x);
              r += stride;
              if (r >= eoff) r -= eoff;
            }
          }
          __syncwarp();
          for (uint32_t i = lane; i < rlen; i += 32) dst[o + s0 + i] = sm.ring[(o + s0 + i) & (kRing - 1)];
        }
      }
      o += T;
      if (m < 32) break;
    }
    pos = next_pos;
  }
  if (o != ulen) { if (lane == 0) atomicExch(err, 104); }
}

__device__ __forceinline__ uint64_t chunk_scratch_off2(const RgSel& rs, const ChunkDev* chunks, const ColSel* cols, int ci) {
  uint64_t off = rs.scratch_off;
  for (int j = 0; j < ci; j++) {
    ChunkDev cj = chunks[cols[j].col];
    if (cj.codec == 1) off += cj.scratch_bytes;
  }
  return off;
}

// One warp per column chunk, chunks handed out by an atomic ticket in the order (column order[0] of every row group,
// then order[1], ...): the host lists the columns with the most compressed bytes first, so the long pages start early
// and the short ones fill the tail.
__global__ void __launch_bounds__(kWarpsPerCta * 32, 6) snappy_pages_kernel(const __grid_constant__ SnappyJob J) {
  __shared__ WarpSmem s_w[kWarpsPerCta];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  WarpSmem& sm = s_w[wid];
  const uint32_t nsel = J.d_nsel ? *J.d_nsel : J.nsel;
  const uint32_t nchunks = nsel * uint32_t(J.ncols);
  for (;;) {
    uint32_t c = 0;
    if (lane == 0) c = atomicAdd(J.ticket, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= nchunks) return;
    const uint32_t si = c % nsel;
    const int ci = J.order[c / nsel];
    RgSel rs = J.sel[si];
    SstDev sst = J.ssts[rs.sst];
    const ChunkDev* chunks = sst.chunks + size_t(rs.rg) * sst.ncols;
    ChunkDev ch = chunks[J.col_from_cols ? J.cols[ci].col : J.col[ci]];
    if (ch.codec != 1) continue;
    if (ch.stored && J.skip_stored[ci]) continue;                 // read in place by the consumer
    uint8_t* dst = J.scratch + (J.fixed_stride ? rs.scratch_off + uint64_t(J.region[ci]) * J.fixed_stride
                                               : chunk_scratch_off2(rs, chunks, J.cols, ci));
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
      if (compressed) snappy_page(src, n, dst, ulen, sm, lane, J.err);
      dst += page_scratch2(pg.uncomp_size);
    }
  }
}

}  // namespace

void snappy_pages(const Launch& L, const SnappyJob& job, uint32_t max_chunks) {
  if (!max_chunks) return;
  uint32_t ctas = (max_chunks + kWarpsPerCta - 1) / kWarpsPerCta;
  if (ctas > 148u * 6) ctas = 148u * 6;
  snappy_pages_kernel<<<ctas, kWarpsPerCta * 32, 0, L.stream>>>(job);
  L.tick();
}

void snappy_chunks_v2(const Launch& L, const SstDev* ssts, const RgSel* sel, uint32_t nsel, const ColSel* cols, int ncolsel,
                      uint8_t* scratch, unsigned int* ticket, int* err) {
  if (!nsel || !ncolsel) return;
  SnappyJob J;
  std::memset(&J, 0, sizeof(J));
  J.ssts = ssts; J.sel = sel; J.d_nsel = nullptr; J.nsel = nsel; J.cols = cols; J.ncols = ncolsel;
  J.scratch = scratch; J.ticket = ticket; J.err = err; J.fixed_stride = 0;
  for (int i = 0; i < ncolsel && i < kSnappyMaxCols; i++) { J.order[i] = i; J.col[i] = 0; }
  // general pipeline: the column ids live in the device-side ColSel array; copy them lazily inside the kernel
  J.col_from_cols = 1;
  snappy_pages(L, J, nsel * uint32_t(ncolsel));
}

}  // namespace k
}  // namespace horae
