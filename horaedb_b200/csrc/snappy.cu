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

namespace horae {
namespace k {

namespace {

constexpr int kWin = 256;          // compressed-stream window (bytes)
constexpr int kHist = 1024;        // bytes of most recent output kept in shared memory
constexpr int kOut = 2048;         // max output of one batch: 32 elements x 64 bytes
constexpr int kLevels = 5;         // J1, J2, J4, J8, J16
constexpr uint16_t kExit = 0xffff;
constexpr uint16_t kDone = 0xffff;
constexpr int kWarpsPerCta = 4;

struct alignas(16) WarpSmem {
  uint8_t win[kWin + 16];
  uint16_t J[kLevels][kWin];
  uint8_t buf[kHist + kOut];       // [0, kHist) = tail of the output produced so far, [kHist, ..) = current batch
  uint16_t ptr[kOut];              // per batch byte: index into buf of its source, or kDone
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

// keep the last kHist bytes of [hist | batch of T bytes] as the new history
__device__ __forceinline__ void shift_history(WarpSmem& sm, uint32_t T, int lane) {
  if (T == 0) return;
  __syncwarp();
  if (T >= uint32_t(kHist)) {
    for (uint32_t i = lane; i < uint32_t(kHist); i += 32) sm.buf[i] = sm.buf[T + i];   // src index > dst index, ascending i per lane
  } else {
    // overlapping forward move by T: do it in two synchronised phases through registers
    uint8_t tmp[kHist / 32];
#pragma unroll
    for (int j = 0; j < kHist / 32; j++) tmp[j] = sm.buf[T + j * 32 + lane];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kHist / 32; j++) sm.buf[j * 32 + lane] = tmp[j];
  }
  __syncwarp();
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
  uint32_t hist_valid = 0;        // how many bytes of history in sm.buf[kHist - hist_valid, kHist) are valid
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
      // refresh history with the tail of the literal
      __syncwarp();
      if (len >= uint32_t(kHist)) {
        for (uint32_t i = lane; i < uint32_t(kHist); i += 32) sm.buf[i] = __ldg(lsrc + (len - kHist) + i);
        hist_valid = kHist;
      } else {
        for (uint32_t i = lane; i < len; i += 32) sm.buf[kHist + i] = __ldg(lsrc + i);
        shift_history(sm, len, lane);
        hist_valid = hist_valid + len < uint32_t(kHist) ? hist_valid + len : uint32_t(kHist);
      }
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
    uint32_t q = 0;                                  // window-relative start of element `lane`
#pragma unroll
    for (int lv = 0; lv < kLevels; lv++)
      if ((lane >> lv) & 1) q = q == kExit ? uint32_t(kExit) : sm.J[lv][q];
    bool valid = q != kExit;
    uint32_t len = 0, off = 0, hdr = 0, csz = 0;
    bool is_lit = false;
    if (valid) {
      const uint32_t t = sm.win[q];
      const uint32_t kind = t & 3;
      if (kind == 0) {
        if ((t >> 2) >= 60) valid = false;           // long literal: ends the batch, handled by the straight-copy path
        else { is_lit = true; len = (t >> 2) + 1; hdr = 1; }
      } else if (kind == 1) { len = ((t >> 2) & 7) + 4; off = ((t >> 5) << 8) | sm.win[q + 1]; hdr = 2; }
      else if (kind == 2) { len = (t >> 2) + 1; off = uint32_t(sm.win[q + 1]) | (uint32_t(sm.win[q + 2]) << 8); hdr = 3; }
      else { len = (t >> 2) + 1; off = uint32_t(sm.win[q + 1]) | (uint32_t(sm.win[q + 2]) << 8) | (uint32_t(sm.win[q + 3]) << 16) | (uint32_t(sm.win[q + 4]) << 24); hdr = 5; }
      csz = hdr + (is_lit ? len : 0);
      if (valid && q + csz > avail) valid = false;   // truncated stream: reported below through the size check
    }
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    const int m = (vm == 0xffffffffu) ? 32 : (__ffs(~vm) - 1);   // valid lanes form a prefix
    if (m == 0) { if (lane == 0) atomicExch(err, 105); return; }
    if (lane >= m) { len = 0; csz = 0; }
    // exclusive prefix sum of the output lengths
    uint32_t inc = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    const uint32_t doff = inc - len;
    const uint32_t T = __shfl_sync(0xffffffffu, inc, 31);
    const uint32_t next_pos = pos + __shfl_sync(0xffffffffu, q + csz, m - 1);
    // sanity: copies must not reach before the start of the output
    bool bad = lane < m && !is_lit && (off == 0 || off > o + doff);
    if (__any_sync(0xffffffffu, bad) || o + T > ulen) { if (lane == 0) atomicExch(err, 103); return; }

    if (T <= uint32_t(m) * 16) {
      // ---------------- tiny elements: byte-level source pointers + pointer jumping
      if (lane < m) {
        if (is_lit) {
          const uint8_t* ls = src + pos + q + 1;
          for (uint32_t i = 0; i < len; i++) {
            const uint32_t wp = q + 1 + i;                               // literal bytes usually sit in the staged window
            sm.buf[kHist + doff + i] = wp < uint32_t(kWin) ? sm.win[wp] : __ldg(ls + i);
            sm.ptr[doff + i] = kDone;
          }
        } else {
          for (uint32_t i = 0; i < len; i++) {
            const uint32_t r = doff + i;
            const int32_t sr = int32_t(kHist + r) - int32_t(off);       // index into buf of the source byte
            if (sr >= int32_t(kHist)) sm.ptr[r] = uint16_t(sr - kHist);   // produced by this batch: resolve below
            else {
              uint8_t v;
              if (sr >= int32_t(kHist - hist_valid) && sr >= 0) v = sm.buf[sr];
              else v = ldcg_u8(dst + (o + r - off));                     // older than the history window
              sm.buf[kHist + r] = v;
              sm.ptr[r] = kDone;
            }
          }
        }
      }
      __syncwarp();
      const uint32_t trips = (T + 31) / 32;
      for (;;) {
        bool pending = false;
        for (uint32_t j = 0; j < trips; j++) {
          const uint32_t r = j * 32 + lane;
          uint16_t pr = kDone, pq = kDone;
          uint8_t vq = 0;
          if (r < T) {
            pr = sm.ptr[r];
            if (pr != kDone) { pq = sm.ptr[pr]; vq = sm.buf[kHist + pr]; }
          }
          __syncwarp();
          if (r < T && pr != kDone) {
            if (pq == kDone) { sm.buf[kHist + r] = vq; sm.ptr[r] = kDone; }
            else { sm.ptr[r] = pq; pending = true; }
          }
          __syncwarp();
        }
        if (!__any_sync(0xffffffffu, pending)) break;
      }
      for (uint32_t r = lane; r < T; r += 32) dst[o + r] = sm.buf[kHist + r];
      shift_history(sm, T, lane);
    } else {
      // ---------------- long elements: one element per step, history served from shared memory
      uint32_t produced = 0;                          // bytes of this batch already placed in buf[kHist..)
      for (int e = 0; e < m; e++) {
        const uint32_t elen = __shfl_sync(0xffffffffu, len, e);
        const uint32_t eoff = __shfl_sync(0xffffffffu, off, e);
        const uint32_t eq = __shfl_sync(0xffffffffu, q, e);
        const bool elit = __shfl_sync(0xffffffffu, int(is_lit), e) != 0;
        __syncwarp();
        if (elit) {
          const uint8_t* ls = src + pos + eq + 1;
          for (uint32_t i = lane; i < elen; i += 32) sm.buf[kHist + produced + i] = __ldg(ls + i);
        } else {
          // source bytes all precede this element: out[x] = out[x - off] with x - off taken modulo the pattern length
          for (uint32_t i = lane; i < elen; i += 32) {
            const uint32_t back = eoff - (i % eoff);                   // distance from the element start to the source byte
            const int32_t sr = int32_t(kHist + produced) - int32_t(back);
            uint8_t v;
            if (sr >= int32_t(kHist - hist_valid) && sr >= 0) v = sm.buf[sr];
            else v = ldcg_u8(dst + (o + produced - back));
            sm.buf[kHist + produced + i] = v;
          }
        }
        __syncwarp();
        for (uint32_t i = lane; i < elen; i += 32) dst[o + produced + i] = sm.buf[kHist + produced + i];
        produced += elen;
      }
      shift_history(sm, T, lane);
    }
    hist_valid = hist_valid + T < uint32_t(kHist) ? hist_valid + T : uint32_t(kHist);
    o += T;
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

__global__ void __launch_bounds__(kWarpsPerCta * 32) snappy_chunks_v2_kernel(const SstDev* __restrict__ ssts, const RgSel* __restrict__ sel,
                                                                             const ColSel* __restrict__ cols, int ncolsel, uint32_t nchunks,
                                                                             uint8_t* __restrict__ scratch, unsigned int* ticket, int* err) {
  __shared__ WarpSmem s_w[kWarpsPerCta];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  WarpSmem& sm = s_w[wid];
  for (;;) {
    uint32_t c = 0;
    if (lane == 0) c = atomicAdd(ticket, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= nchunks) return;
    const uint32_t si = c / ncolsel;
    const int ci = int(c % ncolsel);
    RgSel rs = sel[si];
    SstDev sst = ssts[rs.sst];
    const ChunkDev* chunks = sst.chunks + size_t(rs.rg) * sst.ncols;
    ChunkDev ch = chunks[cols[ci].col];
    if (ch.codec != 1) continue;
    uint8_t* dst = scratch + chunk_scratch_off2(rs, chunks, cols, ci);
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
      if (compressed) snappy_page(src, n, dst, ulen, sm, lane, err);
      dst += page_scratch2(pg.uncomp_size);
    }
  }
}

}  // namespace

void snappy_chunks_v2(const Launch& L, const SstDev* ssts, const RgSel* sel, uint32_t nsel, const ColSel* cols, int ncolsel,
                      uint8_t* scratch, unsigned int* ticket, int* err) {
  if (!nsel || !ncolsel) return;
  const uint32_t nchunks = nsel * uint32_t(ncolsel);
  uint32_t ctas = (nchunks + kWarpsPerCta - 1) / kWarpsPerCta;
  if (ctas > 148u * 5) ctas = 148u * 5;
  snappy_chunks_v2_kernel<<<ctas, kWarpsPerCta * 32, 0, L.stream>>>(ssts, sel, cols, ncolsel, nchunks, scratch, ticket, err);
  L.tick();
}

}  // namespace k
}  // namespace horae
