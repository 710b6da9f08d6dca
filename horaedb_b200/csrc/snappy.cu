// snappy.cu — raw-Snappy page decompression (Parquet SNAPPY codec: parquet 53.2 -> snap 1.1.1 in the reference,
// Cargo.lock:3145; format restated from the published Snappy format description).
//
// Snappy is byte-serial by definition: an element's position depends on all element lengths before it, and a copy may
// read bytes produced by the element just before it.  One warp owns one page and breaks both dependencies:
//
//   parse    a 256-byte window of the compressed stream is staged in shared memory; every byte position computes
//            "where would the next element start if one started here" (J1, from the tag byte), five doubling steps give
//            J2..J32, and lane k finds the start of the k-th element after ANY start position with 5 dependent lookups
//            (binary lifting): 32 elements are decoded per step instead of one, and a batch may end after any element.
//   execute  the longest prefix of the batch that one of two modes can take:
//     word mode   elements of <= 8 bytes (fixed-width numeric columns compress to literal(1-2) + copy(6-7) pairs): every
//                 lane builds its element's bytes in ONE 64-bit register — from the staged literal, from the ring / the
//                 page's earlier output, or from an earlier element of the same batch (parent links collapsed with five
//                 register shuffles) — and drops them into a shared-memory ring.  An element whose source straddles
//                 two elements of the batch simply ends the prefix: it starts the next batch, where its source is old.
//     run mode    a long element, or a run of copies with one offset (RLE-like columns: 64-byte copies at offset 4/8):
//                 out[x] = out[x - off] over the union, i.e. one periodic pattern; for off in {1,2,4,8} that is a single
//                 64-bit word stored to every aligned word of the run.
//   flush    the ring is written to global memory in aligned 8-byte words, all lanes at once.
//   literals longer than 60 bytes (incompressible columns are one literal per 64 KiB block) are plain warp copies.
#include "kernels.h"

#include <cstring>

namespace horae {
namespace k {

namespace {

constexpr int kWin = 256;          // compressed-stream window covered by the jump tables (bytes)
constexpr int kWinPad = 16;        // staged beyond the window: the payload of a <= 8-byte literal that starts near its end
constexpr int kRing = 4096;        // ring buffer of the most recent output (power of two)
constexpr int kHist = 2048;        // bytes before the current batch that are guaranteed to still be in the ring
constexpr int kLevels = 5;         // J1, J2, J4, J8, J16 (the next batch starts right after the last executed element)
constexpr uint32_t kExit = 0xff;      // positions inside the window are <= kWin - 5: one byte per table entry
constexpr int kWarpsPerCta = 4;
constexpr uint32_t kRestage = kWin - 96;   // start a new window when a batch would begin beyond this position

struct alignas(16) WarpSmem {
  uint64_t ring64[kRing / 8];      // output byte at absolute position x lives at byte x & (kRing-1)
  uint8_t J[kLevels][kWin];
  uint8_t win[kWin + kWinPad];
};

__host__ __device__ __forceinline__ uint64_t page_scratch2(uint32_t uncomp) { return (uint64_t(uncomp) + 15u) / 16u * 16u + 32u; }

__device__ __forceinline__ uint64_t funnel64(uint64_t lo, uint64_t hi, uint32_t sh_bits) {   // sh_bits in {0, 8, .., 56}
  return (lo >> sh_bits) | ((hi << 1) << (63 - sh_bits));
}
// 8 bytes of read-only input at any alignment
__device__ __forceinline__ uint64_t ld8_any(const uint8_t* p) {
  uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint64_t* q = reinterpret_cast<const uint64_t*>(a & ~uintptr_t(7));
  return funnel64(__ldg(q), __ldg(q + 1), uint32_t(a & 7) * 8);
}
// 8 bytes of this page's earlier OUTPUT at any alignment (written by this warp: coherent loads, never the read-only path)
__device__ __forceinline__ uint64_t ld8_out(const uint8_t* p) {
  uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint64_t* q = reinterpret_cast<const uint64_t*>(a & ~uintptr_t(7));
  uint64_t lo, hi;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(lo) : "l"(q));
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(hi) : "l"(q + 1));
  return funnel64(lo, hi, uint32_t(a & 7) * 8);
}
__device__ __forceinline__ uint8_t ldcg_u8(const uint8_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return uint8_t(v);
}
__device__ __forceinline__ uint8_t* ring_bytes(WarpSmem& sm) { return reinterpret_cast<uint8_t*>(sm.ring64); }
// 8 ring bytes starting at absolute output position x (any alignment)
__device__ __forceinline__ uint64_t ring_ld8(const WarpSmem& sm, uint32_t x) {
  const uint32_t w = (x >> 3) & (kRing / 8 - 1);
  return funnel64(sm.ring64[w], sm.ring64[(w + 1) & (kRing / 8 - 1)], (x & 7) * 8);
}
__device__ __forceinline__ uint64_t win_ld8(const WarpSmem& sm, uint32_t p) {          // p + 8 <= kWin + kWinPad
  const uint32_t* w = reinterpret_cast<const uint32_t*>(sm.win) + (p >> 2);
  const uint32_t sh = (p & 3) * 8;
  const uint32_t a = w[0], b = w[1], c = w[2];
  return (uint64_t(__funnelshift_r(b, c, sh)) << 32) | __funnelshift_r(a, b, sh);
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, uint32_t(v), src);
  uint32_t hi = __shfl_sync(0xffffffffu, uint32_t(v >> 32), src);
  return (uint64_t(hi) << 32) | lo;
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
#pragma unroll 4
  for (uint32_t w = lane; w < nwords; w += 32) d8[w] = ld8_any(s + (size_t(w) << 3));
  uint32_t done = head + (nwords << 3);
  for (uint32_t i = done + lane; i < len; i += 32) dst[i] = __ldg(src + i);
}

// byte at absolute output position x (< o, i.e. produced by an earlier batch): ring if recent enough, else global
__device__ __forceinline__ uint8_t old_byte(WarpSmem& sm, const uint8_t* dst, uint32_t o, uint32_t x) {
  return (o - x <= uint32_t(kHist)) ? ring_bytes(sm)[x & (kRing - 1)] : ldcg_u8(dst + x);
}

// ring -> global in whole 32-byte sectors [fl, align_down(upto, 32)), one 8-byte word per lane and trip; returns the new flush
// position.  (Partial sectors would make L2 fetch the rest of the sector from DRAM before the write-back.)
__device__ __forceinline__ uint32_t flush_words(const WarpSmem& sm, uint8_t* dst, uint32_t fl, uint32_t upto, int lane) {
  const uint32_t w0 = fl >> 3, w1 = (upto >> 5) << 2;
  uint64_t* d8 = reinterpret_cast<uint64_t*>(dst);
  for (uint32_t w = w0 + lane; w < w1; w += 32) d8[w] = sm.ring64[w & (kRing / 8 - 1)];
  return w1 << 3;
}

// stop_at: the consumer only needs the first stop_at bytes of the page (>= ulen: all of it).  Decoding may overshoot by one batch.
// csz: CTA-shared table, compressed size of an element by its tag byte (elem_csize)
__device__ void snappy_page(const uint8_t* __restrict__ src, uint32_t n, uint8_t* __restrict__ dst, uint32_t ulen_expected, uint32_t stop_at,
                            WarpSmem& sm, const uint8_t* __restrict__ csz, int lane, int* err) {
  uint32_t pos = 0, ulen = 0;
  for (int sh = 0; pos < n && sh < 35; sh += 7) {
    uint32_t b = __ldg(src + pos++);
    ulen |= (b & 0x7f) << sh;
    if (!(b & 0x80)) break;
  }
  if (ulen != ulen_expected) { if (lane == 0) atomicExch(err, 101); return; }
  uint8_t* const ring = ring_bytes(sm);
  uint32_t o = 0;                 // bytes produced so far
  uint32_t fl = 0;                // output bytes [0, fl) are in global memory (fl is a multiple of 32, fl <= o)
  while (pos < n && o < stop_at) {
    const uint32_t avail = n - pos;
    const uint32_t tag0 = __ldg(src + pos);
    // ---- literal with an explicit length field: straight copy
    if ((tag0 & 3) == 0 && (tag0 >> 2) >= 60) {
      uint32_t nb = (tag0 >> 2) - 59, len = 0;
      for (uint32_t i = 0; i < nb; i++) len |= uint32_t(__ldg(src + pos + 1 + i)) << (8 * i);
      len += 1;
      if (1 + nb + len > avail || o + len > ulen || len < 1) { if (lane == 0) atomicExch(err, 102); return; }
      const uint8_t* lsrc = src + pos + 1 + nb;
      __syncwarp();
      if (uint32_t(lane) < o - fl) dst[fl + lane] = ring[(fl + lane) & (kRing - 1)];     // pending partial sector (< 32 bytes)
      warp_copy_in(dst + o, lsrc, len, lane);
      // the ring keeps the tail of the literal (whole words where possible)
      const uint32_t keep = len < uint32_t(kHist) ? len : uint32_t(kHist);
      const uint32_t k0 = o + len - keep, k1 = o + len;
      const uint32_t a0 = (k0 + 7) & ~7u, a1 = k1 & ~7u;
      if (a0 < a1) {
        for (uint32_t w = (a0 >> 3) + lane; w < (a1 >> 3); w += 32) sm.ring64[w & (kRing / 8 - 1)] = ld8_any(lsrc + ((w << 3) - o));
        if (k0 + lane < a0) ring[(k0 + lane) & (kRing - 1)] = __ldg(lsrc + (k0 + lane - o));
        if (a1 + lane < k1) ring[(a1 + lane) & (kRing - 1)] = __ldg(lsrc + (a1 + lane - o));
      } else {
        for (uint32_t i = k0 + lane; i < k1; i += 32) ring[i & (kRing - 1)] = __ldg(lsrc + (i - o));
      }
      __syncwarp();
      pos += 1 + nb + len;
      o += len;
      fl = o & ~31u;
      continue;
    }
    // ---- stage the window and build the jump tables.  Lane l owns the 8 positions [8l, 8l+8): their tag bytes are the window word it
    //      just loaded, a table row is one 64-bit store per lane and level.  J[lv][p] = start of the 2^lv-th element after the one at
    //      p, kExit when that leaves the window; position kWin-1 can never start an element with a staged successor, so
    //      J[lv][kExit] == kExit on every level and the lookups need no test for kExit.
    __syncwarp();
    uint32_t jlo = 0, jhi = 0;
    {
      const uint64_t w = (uint32_t(lane) * 8 < avail + 8) ? ld8_any(src + pos + lane * 8) : 0ull;
      reinterpret_cast<uint64_t*>(sm.win)[lane] = w;
      if (lane < kWinPad / 8) reinterpret_cast<uint64_t*>(sm.win)[32 + lane] = (uint32_t(kWin + lane * 8) < avail + 8) ? ld8_any(src + pos + kWin + lane * 8) : 0ull;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const uint32_t sz = csz[uint32_t(w >> (8 * i)) & 0xffu];
        const uint32_t nx = uint32_t(lane) * 8 + i + sz;
        // the NEXT element must start inside the stream and have its (<= 5 byte) header inside the window
        const uint32_t v = (sz == 0 || nx + 5 > uint32_t(kWin) || nx >= avail) ? kExit : nx;
        if (i < 4) jlo |= v << (8 * i); else jhi |= v << (8 * (i - 4));
      }
      reinterpret_cast<uint2*>(sm.J[0])[lane] = make_uint2(jlo, jhi);
    }
    __syncwarp();
#pragma unroll
    for (int lv = 1; lv < kLevels; lv++) {
      uint32_t nlo = 0, nhi = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        nlo |= uint32_t(sm.J[lv - 1][(jlo >> (8 * i)) & 0xffu]) << (8 * i);
        nhi |= uint32_t(sm.J[lv - 1][(jhi >> (8 * i)) & 0xffu]) << (8 * i);
      }
      jlo = nlo; jhi = nhi;
      reinterpret_cast<uint2*>(sm.J[lv])[lane] = make_uint2(jlo, jhi);
      __syncwarp();
    }
    uint32_t qs = 0;                                   // window-relative start of the next batch
    bool first = true;
    for (;;) {
      uint32_t q = qs;
#pragma unroll
      for (int lv = 0; lv < 5; lv++)
        if ((lane >> lv) & 1) q = sm.J[lv][q];
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
        if (q + csz > avail) valid = false;          // truncated stream: caught by m == 0 / the final size check
      }
      const unsigned vm = __ballot_sync(0xffffffffu, valid);
      const int m = (vm == 0xffffffffu) ? 32 : (__ffs(~vm) - 1);   // valid lanes form a prefix
      if (m == 0) {
        if (first) { if (lane == 0) atomicExch(err, 105); return; }
        pos += qs;                                                 // a long literal (or the window's end) starts here: restage
        break;
      }
      first = false;
      if (lane >= m) { len = 0; csz = 0; off = 1; is_lit = true; }
      uint32_t inc = len;                                          // inclusive prefix sum of the output lengths
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
      const uint32_t doff = inc - len;
      {
        const bool bad = lane < m && !is_lit && (off == 0 || off > o + doff);
        if (__any_sync(0xffffffffu, bad)) { if (lane == 0) atomicExch(err, 103); return; }
      }
      const uint32_t len0 = __shfl_sync(0xffffffffu, len, 0);
      int cnt;                                                     // elements executed by this step (a prefix of the batch)
      uint32_t T;                                                  // their output bytes
      if (len0 <= 8) {
        // ---------------- word mode
        // where does this element's data come from?  0 literal bytes in the window, 1 earlier output (ring / global),
        // 2 one earlier element of this batch (parent, byte delta).  Anything else ends the prefix.
        uint32_t skind = 0, spos = q + 1;
        int parent = lane;
        uint32_t delta = 0;
        bool fail = len > 8;
        if (!is_lit && !fail) {
          const int32_t s0 = int32_t(doff) - int32_t(off);
          if (off >= len) {
            if (s0 + int32_t(len) <= 0) { skind = 1; spos = o + doff - off; }
            else if (s0 < 0) fail = true;                          // straddles the batch start
            else skind = 2;
          } else {                                                 // self-overlapping (periodic) copy: fine if its pattern is old
            if (doff == 0) { skind = 1; spos = o - off; } else fail = true;
          }
        }
        // parent = the element that contains byte s0 (binary search over the element starts, register shuffles only)
        {
          const uint32_t s0 = doff - off;
          int lo = 0;
#pragma unroll
          for (int step = 16; step > 0; step >>= 1) {
            const int cand = lo + step;
            const uint32_t d = __shfl_sync(0xffffffffu, doff, cand & 31);
            if (skind == 2 && cand < lane && d <= s0) lo = cand;
          }
          const uint32_t pd = __shfl_sync(0xffffffffu, doff, lo), pl = __shfl_sync(0xffffffffu, len, lo);
          if (skind == 2) {
            if (lo >= lane || s0 < pd || s0 + len > pd + pl) fail = true;     // not inside ONE earlier element
            else { parent = lo; delta = s0 - pd; }
          }
        }
        const unsigned fm = __ballot_sync(0xffffffffu, fail || lane >= m);
        cnt = fm ? __ffs(fm) - 1 : 32;                             // >= 1: element 0 has len <= 8 and an old / literal source
        if (lane >= cnt) { parent = lane; delta = 0; }
        // collapse parent chains (parents are always earlier lanes inside the prefix)
#pragma unroll
        for (int it = 0; it < 5; it++) {
          const uint32_t d2 = __shfl_sync(0xffffffffu, delta, parent);
          const int p2 = __shfl_sync(0xffffffffu, parent, parent);
          delta += d2;
          parent = p2;
        }
        uint64_t w = 0;
        if (lane < cnt && skind != 2) {
          if (skind == 0) w = win_ld8(sm, spos);
          else {
            w = (o - spos <= uint32_t(kHist)) ? ring_ld8(sm, spos) : ld8_out(dst + spos);
            if (off < len) {                                       // periodic: repeat the first `off` bytes
              w &= (off >= 8) ? ~0ull : ((1ull << (8 * off)) - 1);
              for (uint32_t f = off; f < 8; f <<= 1) w |= w << (8 * f);
            }
          }
        }
        {
          const uint64_t wr = shfl64(w, parent);
          if (skind == 2) w = wr >> (8 * delta);
        }
        T = __shfl_sync(0xffffffffu, inc, cnt - 1);
        if (o + T > ulen) { if (lane == 0) atomicExch(err, 103); return; }
        if (lane < cnt) {
          const uint32_t rb = (o + doff) & (kRing - 1);
          uint8_t* rp = ring + rb;
          if (rb <= uint32_t(kRing) - 8) {             // no wrap inside the element: fixed offsets from one pointer
#pragma unroll
            for (int i = 0; i < 8; i++)
              if (uint32_t(i) < len) rp[i] = uint8_t(w >> (8 * i));
          } else {
#pragma unroll
            for (int i = 0; i < 8; i++)
              if (uint32_t(i) < len) ring[(rb + i) & (kRing - 1)] = uint8_t(w >> (8 * i));
          }
        }
      } else {
        // ---------------- run mode: element 0 is long.  A literal goes alone; a copy takes every following copy with the
        // same offset along (one periodic pattern over the union).
        const uint32_t off0 = __shfl_sync(0xffffffffu, off, 0);
        const bool lit0 = __shfl_sync(0xffffffffu, int(is_lit), 0) != 0;
        if (lit0) cnt = 1;
        else {
          const unsigned brk = __ballot_sync(0xffffffffu, lane >= m || is_lit || off != off0);
          cnt = brk ? __ffs(brk) - 1 : 32;
        }
        T = __shfl_sync(0xffffffffu, inc, cnt - 1);
        if (o + T > ulen) { if (lane == 0) atomicExch(err, 103); return; }
        __syncwarp();
        if (lit0) {
          const uint32_t q0 = __shfl_sync(0xffffffffu, q, 0);
          const uint8_t* ls = src + pos + q0 + 1;
          for (uint32_t i = lane; i < T; i += 32) ring[(o + i) & (kRing - 1)] = __ldg(ls + i);
        } else if (off0 == 8 || off0 == 4 || off0 == 2 || off0 == 1) {
          // the pattern as one 64-bit word, phased for 8-aligned absolute positions (off0 divides 8)
          uint64_t pw = ring_ld8(sm, o - off0);
          pw &= (off0 >= 8) ? ~0ull : ((1ull << (8 * off0)) - 1);
          for (uint32_t f = off0; f < 8; f <<= 1) pw |= pw << (8 * f);
          const uint32_t c = (off0 - (o % off0)) % off0;         // (aligned address - o) mod off0
          const uint64_t W = c ? ((pw >> (8 * c)) | (pw << (64 - 8 * c))) : pw;
          const uint32_t a0 = (o + 7) & ~7u, a1 = (o + T) & ~7u;
          if (a0 < a1) {
            for (uint32_t wd = (a0 >> 3) + lane; wd < (a1 >> 3); wd += 32) sm.ring64[wd & (kRing / 8 - 1)] = W;
            if (o + lane < a0) ring[(o + lane) & (kRing - 1)] = uint8_t(W >> (8 * ((o + lane) & 7)));
            if (a1 + lane < o + T) ring[(a1 + lane) & (kRing - 1)] = uint8_t(W >> (8 * lane));
          } else {
            for (uint32_t i = o + lane; i < o + T; i += 32) ring[i & (kRing - 1)] = uint8_t(W >> (8 * (i & 7)));
          }
        } else {
          // any other offset: byte i of the run = old byte (i mod off0) of the pattern (kept incrementally: no division per byte)
          const uint32_t stride = 32u % off0;
          uint32_t r = uint32_t(lane) % off0;
          for (uint32_t i = lane; i < T; i += 32) {
            const uint32_t x = o - off0 + r;
            ring[(o + i) & (kRing - 1)] = old_byte(sm, dst, o, x);
            r += stride;
            if (r >= off0) r -= off0;
          }
        }
      }
      __syncwarp();
      o += T;
      fl = flush_words(sm, dst, fl, o, lane);
      // where the next batch starts: right after the last executed element
      const uint32_t adv = __shfl_sync(0xffffffffu, q + csz, cnt - 1);
      if (adv > kRestage || adv >= avail || o >= stop_at) { pos += adv; break; }      // (kRestage + 5 <= kWin: the next header is staged)
      qs = adv;
      __syncwarp();
    }
  }
  __syncwarp();
  if (fl + lane < o) dst[fl + lane] = ring[(fl + lane) & (kRing - 1)];
  if (o != ulen && stop_at >= ulen) { if (lane == 0) atomicExch(err, 104); }
}

__device__ __forceinline__ uint64_t chunk_scratch_off2(const RgSel& rs, const ChunkDev* chunks, const ColSel* cols, int ci) {
  uint64_t off = rs.scratch_off;
  for (int j = 0; j < ci; j++) {
    off += chunks[cols[j].col].scratch_bytes;      // 0 for uncompressed PLAIN chunks
  }
  return off;
}

// One warp per column chunk, chunks handed out by an atomic ticket in the order (column order[0] of every row group,
// then order[1], ...): the host lists the columns with the most compressed bytes first, so the long pages start early
// and the short ones fill the tail.
__global__ void __launch_bounds__(kWarpsPerCta * 32, 8) snappy_pages_kernel(const __grid_constant__ SnappyJob J) {
  __shared__ WarpSmem s_w[kWarpsPerCta];
  __shared__ uint8_t s_csz[256];
  for (uint32_t t = threadIdx.x; t < 256; t += kWarpsPerCta * 32) s_csz[t] = uint8_t(elem_csize(t));
  __syncthreads();
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
    if (ch.dict_uncomp) {                                        // compressed dictionary page: first in the chunk's scratch
      const uint8_t* dsrc = sst.bytes + ch.dict_payload_off;
      snappy_page(dsrc, ch.dict_comp, dst, ch.dict_uncomp, 0xffffffffu, sm, s_csz, lane, J.err);
      dst += page_scratch2(ch.dict_uncomp);
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
      uint32_t stop_at = 0xffffffffu;
      if (J.partial[ci]) {
        // the consumer reads rows [0, rs.out_row) only (gate-first: nothing behind the last row that passes the gate column can
        // survive the filter): level prefix (<= 16 + rows / 8 bytes) + that many values
        const uint32_t w = (ch.phys == 1 || ch.phys == 4) ? 4u : 8u;
        stop_at = 16u + (rs.num_rows + 7u) / 8u + 8u + rs.out_row * w;
      }
      if (compressed) snappy_page(src, n, dst, ulen, stop_at, sm, s_csz, lane, J.err);
      dst += page_scratch2(pg.uncomp_size);
      if (pg.encoding == 5 || pg.encoding == 8 || pg.encoding == 2) dst += page_scratch2(pg.num_values * 8u);   // PLAIN image of a DELTA / dictionary page (decode_chunks)
    }
  }
}

// pages given by pointer (transient loads decompress the gate column before the SST's tables exist on the device)
__global__ void __launch_bounds__(kWarpsPerCta * 32, 8) snappy_raw_kernel(const RawPage* __restrict__ pages, uint32_t n, unsigned int* ticket, int* err) {
  __shared__ WarpSmem s_w[kWarpsPerCta];
  __shared__ uint8_t s_csz[256];
  for (uint32_t t = threadIdx.x; t < 256; t += kWarpsPerCta * 32) s_csz[t] = uint8_t(elem_csize(t));
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (;;) {
    uint32_t c = 0;
    if (lane == 0) c = atomicAdd(ticket, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= n) return;
    const RawPage pg = pages[c];
    snappy_page(pg.src, pg.comp_size, pg.dst, pg.uncomp_size, 0xffffffffu, s_w[wid], s_csz, lane, err);
  }
}

}  // namespace

void snappy_raw_pages(const Launch& L, const RawPage* d_pages, uint32_t n, unsigned int* ticket, int* err) {
  if (!n) return;
  uint32_t ctas = (n + kWarpsPerCta - 1) / kWarpsPerCta;
  if (ctas > 148u * 8) ctas = 148u * 8;
  snappy_raw_kernel<<<ctas, kWarpsPerCta * 32, 0, L.stream>>>(d_pages, n, ticket, err);
  L.tick();
}

void snappy_pages(const Launch& L, const SnappyJob& job, uint32_t max_chunks) {
  if (!max_chunks) return;
  uint32_t ctas = (max_chunks + kWarpsPerCta - 1) / kWarpsPerCta;
  if (ctas > 148u * 8) ctas = 148u * 8;
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
  for (int i = 0; i < ncolsel && i < kSnappyMaxCols; i++) J.order[i] = uint8_t(i);
  J.col_from_cols = 1;      // general pipeline: column ids and the variable scratch layout come from the ColSel table
  snappy_pages(L, J, nsel * uint32_t(ncolsel));
}

}  // namespace k
}  // namespace horae
