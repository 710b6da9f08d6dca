// snappy_core.h — the warp-level raw-Snappy page decoder (Parquet SNAPPY codec: parquet 53.2 -> snap 1.1.1 in the reference,
// Cargo.lock:3145; format restated from the published Snappy format description).
//
// This header is the decoder's single source: snappy.cu compiles it for sm_100a (the product), and the CPU test-suite compiles
// the SAME text with the warp primitives below mapped onto 32 coroutines (tests/emu/snappy_emu.cpp), so the lane-level logic is
// checked without a GPU.  The includer provides, before including:
//   SNP_FN                          function qualifiers (__device__ __forceinline__ / inline)
//   snp_shfl(v, src)  snp_shfl_up(v, d)  snp_ballot(pred)  snp_any(pred)  snp_syncwarp()      full-warp collectives (32-bit values)
//   snp_ldg8(p)  snp_ldg64(p)       read-only input loads (uint8_t / aligned uint64_t)
//   snp_ldcg8(p) snp_ldcg32(p)      coherent loads of this page's earlier OUTPUT (written by other lanes of the warp)
//   snp_funnel_r(lo, hi, sh)        32-bit funnel shift right (sh < 32),  snp_byte_perm(a, b, sel),  snp_ffs(x)
//   snp_set_err(err, code)
//
// Snappy is byte-serial by definition: an element's position depends on all element lengths before it, and a copy may
// read bytes produced by the element just before it.  One warp owns one page and breaks both dependencies:
//
//   parse    a 256-byte window of the compressed stream is staged in shared memory; every byte position computes
//            "where would the next element start if one started here" (J1, from a tag-byte table), four doubling steps give
//            J2..J16, and lane k finds the start of the k-th element after ANY start position with 5 dependent lookups
//            (binary lifting): 32 elements are decoded per step instead of one, and a batch may end after any element.
//            Lane l owns the 8 positions [8l, 8l+8): a table row is one 64-bit store per lane and level, built in registers.
//   execute  the longest prefix of the batch that one of two modes can take:
//     word mode   elements of <= 8 bytes (fixed-width numeric columns compress to literal(1-2) + copy(6-7) pairs): every
//                 lane builds its element's bytes in ONE 64-bit register — from the staged literal, from the ring / the
//                 page's earlier output, or from an earlier element of the same batch (parent links collapsed with five
//                 register shuffles) — and drops them into a shared-memory ring.  An element whose source straddles
//                 two elements of the batch simply ends the prefix: it starts the next batch, where its source is old.
//   staging  the compressed bytes travel global -> shared by cp.async.bulk + mbarrier, one region ahead of the parser.
//     run mode    a long element, or a run of copies with one offset (RLE-like columns: 64-byte copies at offset 4/8):
//                 out[x] = out[x - off] over the union, i.e. one periodic pattern; for off in {1,2,4,8} that is a single
//                 64-bit word stored to every aligned word of the run.
//   flush    the ring is written to global memory in whole 32-byte sectors, 256 bytes per warp instruction.
//   literals longer than 60 bytes (incompressible columns are one literal per 64 KiB block) are plain warp copies.
#pragma once
#include <cstdint>

#ifndef SNP_STAT
#define SNP_STAT(counter, amount)      // the emulator counts windows / steps / elements here
#endif

namespace horae {
namespace snp {

constexpr int kWin = 256;          // compressed-stream window covered by the jump tables (bytes)
constexpr int kWinPad = 16;        // staged beyond the window: header + payload of a <= 8-byte element that starts near its end
constexpr int kRing = 4096;        // ring buffer of the most recent output (power of two)
constexpr int kHist = 2048;        // bytes before the current batch that are guaranteed to still be in the ring
constexpr int kLevels = 5;         // J1, J2, J4, J8, J16 (the next batch starts right after the last executed element)
constexpr uint32_t kExit = 0xff;   // "leaves the window"; window positions are one byte per table entry
constexpr uint32_t kRestage = kWin - 80;   // start a new window when a batch would begin beyond this position (176: 14 % fewer windows than 160 on ts pages, same number of steps)
constexpr uint32_t kFlushAt = 512;         // ring -> global once this many bytes are pending (two 8-byte words per lane; 256 measured 0.8 % slower)
constexpr uint64_t kFill = 0xfcfcfcfcfcfcfcfcull;   // tag of a long literal: what positions behind the stream's end are staged as

// 256-byte aligned, every jump table on a 256-byte boundary: a table address is the block's base with the index as its low
// byte, i.e. ONE byte-permute (index extraction and address formation together) in front of the load.
constexpr int kStage = 448;        // bytes per staging buffer of the compressed stream (multiple of 16: bulk-copy granularity)
constexpr uint32_t kAhead = kRestage;      // the next window starts more than this many bytes behind the current one
struct alignas(256) WarpSmem {
  uint64_t ring64[kRing / 8];      // output byte at absolute position x lives at byte x & (kRing-1)
  uint8_t J[kLevels][kWin];
  // The compressed stream reaches shared memory by asynchronous bulk copies (cp.async.bulk + mbarrier: one instruction moves the
  // whole region, no registers, no per-lane address arithmetic), double-buffered: while the batches of one window execute, the
  // region the next window must lie in is already on its way.  A window is a byte offset into one of the two buffers.
  alignas(16) uint8_t stage[2][kStage];
  uint64_t bar[2];                 // one mbarrier per buffer
};
constexpr uint32_t kJOff = kRing;                          // byte offsets inside WarpSmem
constexpr uint32_t kStageOff = kRing + kLevels * kWin;
constexpr uint32_t kBarOff = kStageOff + 2 * kStage;
// where the staging buffers stand: which one holds the current window, what the other one was asked to fetch, barrier phases
struct StageState {
  uint32_t phase;                  // bit b = parity the next wait on bar[b] uses
  int cur;                         // buffer of the current window
  bool pf;                         // a copy into buffer cur ^ 1 was issued and not yet waited for
  int pf_start;                    // stream position (relative to the page's first byte, may be < 0) of that buffer's byte 0
  uint32_t pf_bytes;
};

// Tag-byte tables (256 entries each, shared by the CTA).
//   csz : compressed size of the element, 255 = literal with a multi-byte length field (never part of a batch)
//   lut : len (bits 0-6) | offset bits 8-10 of a 1-byte-offset copy, in place (bits 8-10) | is_lit << 16 | long_lit << 17 |
//         (32 - 8 * offset bytes) << 18 (5 bits) | csz << 24
SNP_FN uint32_t elem_csize(uint32_t t) {
  const uint32_t kind = t & 3;
  if (kind == 0) { const uint32_t l = t >> 2; return l < 60 ? l + 2 : 255u; }
  return kind == 1 ? 2u : (kind == 2 ? 3u : 5u);
}
SNP_FN uint32_t elem_lut(uint32_t t) {
  const uint32_t kind = t & 3;
  uint32_t len, hdr, offhi = 0, is_lit = 0, long_lit = 0, msh = 0;
  if (kind == 0) {
    const uint32_t l = t >> 2;
    is_lit = 1; hdr = 1;
    if (l < 60) len = l + 1; else { len = 0; long_lit = 1; }
  } else if (kind == 1) { len = ((t >> 2) & 7) + 4; hdr = 2; offhi = t >> 5; msh = 24; }
  else if (kind == 2) { len = (t >> 2) + 1; hdr = 3; msh = 16; }
  else { len = (t >> 2) + 1; hdr = 5; msh = 0; }
  const uint32_t csz = long_lit ? 0u : hdr + (is_lit ? len : 0u);
  return len | (offhi << 8) | (is_lit << 16) | (long_lit << 17) | (msh << 18) | (csz << 24);
}

SNP_FN uint64_t page_scratch2(uint32_t uncomp) { return (uint64_t(uncomp) + 15u) / 16u * 16u + 32u; }

SNP_FN uint64_t funnel64(uint64_t lo, uint64_t hi, uint32_t sh_bits) {   // sh_bits in {0, 8, .., 56}
  return (lo >> sh_bits) | ((hi << 1) << (63 - sh_bits));
}
// 8 bytes of read-only input at any alignment
SNP_FN uint64_t ld8_any(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint64_t* q = reinterpret_cast<const uint64_t*>(a & ~uintptr_t(7));
  return funnel64(snp_ldg64(q), snp_ldg64(q + 1), uint32_t(a & 7) * 8);
}
// 8 bytes of this page's earlier OUTPUT at any alignment (written by this warp: coherent loads, never the read-only path)
SNP_FN uint64_t ld8_out(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
  const uint32_t sh = uint32_t(a & 3) * 8;
  const uint32_t x = snp_ldcg32(q), y = snp_ldcg32(q + 1), z = snp_ldcg32(q + 2);
  return (uint64_t(snp_funnel_r(y, z, sh)) << 32) | snp_funnel_r(x, y, sh);
}
SNP_FN uint8_t* ring_bytes(WarpSmem& sm) { return reinterpret_cast<uint8_t*>(sm.ring64); }
// 8 ring bytes starting at absolute output position x (any alignment, wraps)
SNP_FN uint64_t ring_ld8(const WarpSmem& sm, uint32_t x) {
  const uint32_t w = (x >> 3) & (kRing / 8 - 1);
  return funnel64(sm.ring64[w], sm.ring64[(w + 1) & (kRing / 8 - 1)], (x & 7) * 8);
}
// 8 bytes at byte address a of the warp's shared block (ring ... window), no ring wrap: a + 8 <= kRing, or inside the staged window
SNP_FN uint64_t sm_ld8(const WarpSmem& sm, uint32_t a) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&sm) + (a >> 2);
  const uint32_t sh = (a & 3) * 8;
  const uint32_t x = w[0], y = w[1], z = w[2];
  return (uint64_t(snp_funnel_r(y, z, sh)) << 32) | snp_funnel_r(x, y, sh);
}
SNP_FN uint64_t shfl64(uint64_t v, int src) {
  const uint32_t lo = snp_shfl(uint32_t(v), src);
  const uint32_t hi = snp_shfl(uint32_t(v >> 32), src);
  return (uint64_t(hi) << 32) | lo;
}
// byte i (compile-time) of a 32-bit word, zero-extended / the low byte of v placed into byte i of acc
template <int I> SNP_FN uint32_t byte_of(uint32_t x) { return snp_byte_perm(x, 0u, 0x4440u + I); }
template <int I> SNP_FN uint32_t put_byte(uint32_t acc, uint32_t v) {
  return snp_byte_perm(acc, v, I == 0 ? 0x3214u : (I == 1 ? 0x3240u : (I == 2 ? 0x3410u : 0x4210u)));
}

// J[LV][byte I of packed]; smbase = shared-space address of the warp's block (device only)
// store the low `len` (1..8) bytes of w at ring byte rb (no wrap inside the element)
#ifdef __CUDACC__
template <int I, int LV> SNP_FN uint32_t jt_get(const WarpSmem&, uint32_t smbase, uint32_t packed) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(__byte_perm(packed, smbase, 0x7650u + I)), "n"(kJOff + LV * kWin));
  return v;
}
SNP_FN uint32_t sm_base(const WarpSmem& sm) { return uint32_t(__cvta_generic_to_shared(&sm)); }
#define SNP_ST_BYTE(i, v) asm volatile("{ .reg .pred p; setp.gt.u32 p, %2, " #i "; @p st.shared.u8 [%0+" #i "], %1; }" ::"r"(a), "r"(v), "r"(len) : "memory")
SNP_FN void store_elem(WarpSmem&, uint32_t smbase, uint32_t rb, uint64_t w, uint32_t len) {
  const uint32_t a = smbase + rb, lo = uint32_t(w), hi = uint32_t(w >> 32);
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(lo) : "memory");
  SNP_ST_BYTE(1, lo >> 8); SNP_ST_BYTE(2, lo >> 16); SNP_ST_BYTE(3, lo >> 24);
  SNP_ST_BYTE(4, hi); SNP_ST_BYTE(5, hi >> 8); SNP_ST_BYTE(6, hi >> 16); SNP_ST_BYTE(7, hi >> 24);
}
#undef SNP_ST_BYTE
#else
template <int I, int LV> SNP_FN uint32_t jt_get(const WarpSmem& sm, uint32_t, uint32_t packed) { return sm.J[LV][(packed >> (8 * I)) & 0xffu]; }
SNP_FN uint32_t sm_base(const WarpSmem&) { return 0; }
SNP_FN void store_elem(WarpSmem& sm, uint32_t, uint32_t rb, uint64_t w, uint32_t len) {
  for (uint32_t i = 0; i < len; i++) reinterpret_cast<uint8_t*>(sm.ring64)[rb + i] = uint8_t(w >> (8 * i));
}
#endif
// bulk_init: once per warp (lane 0 initialises both barriers);  bulk_issue: lane 0 starts the copy of `bytes` (multiple of 16) from the
// 16-byte aligned global address g into buffer b;  bulk_wait: every lane blocks until the copy into buffer b has landed.
#ifdef __CUDACC__
SNP_FN void bulk_init(WarpSmem& sm, int lane) {
  if (lane == 0) {
    const uint32_t a = sm_base(sm) + kBarOff;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  snp_syncwarp();
}
SNP_FN void bulk_issue(WarpSmem& sm, int b, const uint8_t* g, uint32_t bytes, int lane) {
  if (lane == 0) {
    const uint32_t bar = sm_base(sm) + kBarOff + 8u * uint32_t(b), dst = sm_base(sm) + kStageOff + uint32_t(kStage) * uint32_t(b);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // the lanes' earlier reads of this buffer are done (syncwarp before)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(g), "r"(bytes), "r"(bar) : "memory");
  }
}
SNP_FN void bulk_wait(WarpSmem& sm, int b, uint32_t parity) {
  const uint32_t bar = sm_base(sm) + kBarOff + 8u * uint32_t(b);
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
#else
SNP_FN void bulk_init(WarpSmem&, int) {}
SNP_FN void bulk_issue(WarpSmem& sm, int b, const uint8_t* g, uint32_t bytes, int lane) {
  if (lane == 0) for (uint32_t i = 0; i < bytes; i++) sm.stage[b][i] = g[i];
}
SNP_FN void bulk_wait(WarpSmem&, int, uint32_t) { snp_syncwarp(); }
#endif
// start fetching the region that holds stream positions [from, from + kStage) (clamped to the stream's end n) into buffer b
SNP_FN void stage_fetch(WarpSmem& sm, StageState& st, int b, const uint8_t* src, uint32_t n, uint32_t from, int lane) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(src + from) & ~uintptr_t(15);
  const uintptr_t end = (reinterpret_cast<uintptr_t>(src + n) + 15) & ~uintptr_t(15);
  uint32_t bytes = uint32_t(end - a);
  if (bytes > uint32_t(kStage)) bytes = uint32_t(kStage);
  st.pf_start = int(intptr_t(a) - intptr_t(reinterpret_cast<uintptr_t>(src)));
  st.pf_bytes = bytes;
  bulk_issue(sm, b, reinterpret_cast<const uint8_t*>(a), bytes, lane);
}

template <int LV> SNP_FN void jt_level(WarpSmem& sm, uint32_t smbase, uint32_t& jlo, uint32_t& jhi, int lane) {
  uint32_t nlo = 0, nhi = 0;
  nlo = put_byte<0>(nlo, jt_get<0, LV - 1>(sm, smbase, jlo)); nhi = put_byte<0>(nhi, jt_get<0, LV - 1>(sm, smbase, jhi));
  nlo = put_byte<1>(nlo, jt_get<1, LV - 1>(sm, smbase, jlo)); nhi = put_byte<1>(nhi, jt_get<1, LV - 1>(sm, smbase, jhi));
  nlo = put_byte<2>(nlo, jt_get<2, LV - 1>(sm, smbase, jlo)); nhi = put_byte<2>(nhi, jt_get<2, LV - 1>(sm, smbase, jhi));
  nlo = put_byte<3>(nlo, jt_get<3, LV - 1>(sm, smbase, jlo)); nhi = put_byte<3>(nhi, jt_get<3, LV - 1>(sm, smbase, jhi));
  jlo = nlo; jhi = nhi;
  reinterpret_cast<uint2*>(sm.J[LV])[lane] = make_uint2(jlo, jhi);
  snp_syncwarp();
}

// plain copy global->global spread over the warp (source is read-only input)
SNP_FN void warp_copy_in(uint8_t* dst, const uint8_t* src, uint32_t len, int lane) {
  uint32_t head = uint32_t((8 - (reinterpret_cast<uintptr_t>(dst) & 7)) & 7);
  if (head > len) head = len;
  if (uint32_t(lane) < head) dst[lane] = snp_ldg8(src + lane);
  const uint32_t nwords = (len - head) >> 3;
  uint64_t* d8 = reinterpret_cast<uint64_t*>(dst + head);
  const uint8_t* s = src + head;
#pragma unroll 4
  for (uint32_t w = lane; w < nwords; w += 32) d8[w] = ld8_any(s + (size_t(w) << 3));
  const uint32_t done = head + (nwords << 3);
  for (uint32_t i = done + lane; i < len; i += 32) dst[i] = snp_ldg8(src + i);
}

// byte at absolute output position x (< o, i.e. produced by an earlier batch): ring if recent enough, else global
SNP_FN uint8_t old_byte(WarpSmem& sm, const uint8_t* dst, uint32_t o, uint32_t x) {
  return (o - x <= uint32_t(kHist)) ? ring_bytes(sm)[x & (kRing - 1)] : snp_ldcg8(dst + x);
}

// ring -> global in whole 32-byte sectors [fl, align_down(upto, 32)), one 8-byte word per lane and trip; returns the new flush
// position.  (Partial sectors would make L2 fetch the rest of the sector from DRAM before the write-back.)
SNP_FN uint32_t flush_words(const WarpSmem& sm, uint8_t* dst, uint32_t fl, uint32_t upto, int lane) {
  const uint32_t w0 = fl >> 3, w1 = (upto >> 5) << 2;
  uint64_t* d8 = reinterpret_cast<uint64_t*>(dst);
  for (uint32_t w = w0 + lane; w < w1; w += 32) d8[w] = sm.ring64[w & (kRing / 8 - 1)];
  return w1 << 3;
}

// stop_at: the consumer only needs the first stop_at bytes of the page (>= ulen: all of it).  Decoding may overshoot by one batch.
// csz / lut: the CTA-shared tag tables (elem_csize / elem_lut)
SNP_FN void snappy_page_body(const uint8_t* __restrict__ src, uint32_t n, uint8_t* __restrict__ dst, uint32_t ulen_expected, uint32_t stop_at,
                             WarpSmem& sm, StageState& st, const uint8_t* __restrict__ csz, const uint32_t* __restrict__ lut, int lane, int* err) {
  uint32_t pos = 0, ulen = 0;
  for (int sh = 0; pos < n && sh < 35; sh += 7) {
    const uint32_t b = snp_ldg8(src + pos++);
    ulen |= (b & 0x7f) << sh;
    if (!(b & 0x80)) break;
  }
  if (ulen != ulen_expected) { if (lane == 0) snp_set_err(err, 101); return; }
  uint8_t* const ring = ring_bytes(sm);
  const uint32_t* const sm32 = reinterpret_cast<const uint32_t*>(&sm);
  const uint8_t* const sm8 = reinterpret_cast<const uint8_t*>(&sm);
  const uint32_t smbase = sm_base(sm);
  uint32_t o = 0;                 // bytes produced so far
  uint32_t fl = 0;                // output bytes [0, fl) are in global memory (fl is a multiple of 32, fl <= o)
  while (pos < n && o < stop_at) {
    const uint32_t avail = n - pos;
    // ---- stage the window (every look at the compressed stream goes through the staged bytes, never a dependent global load)
    snp_syncwarp();
    SNP_STAT(windows, 1);
    // the window's bytes: already fetched (or on their way) if the last window's look-ahead covers [pos, pos + need), else fetched now
    uint32_t wbase;                                    // byte offset of window position 0 inside the warp's shared block
    {
      const uint32_t need = avail < uint32_t(kWin + kWinPad) ? avail : uint32_t(kWin + kWinPad);
      const int nb = st.cur ^ 1;
      bool hit = false;
      if (st.pf) {
        bulk_wait(sm, nb, (st.phase >> nb) & 1u);
        st.phase ^= 1u << nb;
        st.pf = false;
        hit = int(pos) >= st.pf_start && pos + need <= uint32_t(st.pf_start + int(st.pf_bytes));
        SNP_STAT(stage_hits, hit ? 1 : 0);
      }
      if (!hit) {
        stage_fetch(sm, st, nb, src, n, pos, lane);
        bulk_wait(sm, nb, (st.phase >> nb) & 1u);
        st.phase ^= 1u << nb;
      }
      wbase = kStageOff + uint32_t(kStage) * uint32_t(nb) + uint32_t(int(pos) - st.pf_start);
      st.cur = nb;
      // look ahead: the next window starts in (pos + kAhead, pos + kWin + 61]; its region goes into the buffer just left
      if (pos + kAhead < n) { stage_fetch(sm, st, nb ^ 1, src, n, pos + kAhead, lane); st.pf = true; }
    }
    const uint32_t tag0 = sm8[wbase];
    // ---- literal with an explicit length field: straight copy
    if ((tag0 & 3) == 0 && (tag0 >> 2) >= 60) {
      const uint32_t nb = (tag0 >> 2) - 59;
      uint32_t len = 0;
      for (uint32_t i = 0; i < nb && i + 1 < avail; i++) len |= uint32_t(sm8[wbase + 1 + i]) << (8 * i);
      len += 1;
      if (1 + nb > avail || o + len > ulen || len < 1) { if (lane == 0) snp_set_err(err, 102); return; }
      if (len > avail - 1 - nb) {
        // the literal runs past the end of the stream: an error, unless the stream is a compressed PREFIX (transient loads ship only
        // what a partial decode needs) and the bytes that are there reach the position the consumer stops at
        if (o + (avail - 1 - nb) < stop_at) { if (lane == 0) snp_set_err(err, 102); return; }
        len = avail - 1 - nb;
      }
      const uint8_t* lsrc = src + pos + 1 + nb;
      snp_syncwarp();
      fl = flush_words(sm, dst, fl, o, lane);
      if (uint32_t(lane) < o - fl) dst[fl + lane] = ring[(fl + lane) & (kRing - 1)];     // pending partial sector (< 32 bytes)
      warp_copy_in(dst + o, lsrc, len, lane);
      snp_syncwarp();                                  // every lane has read its pending ring words before the tail below overwrites them
      // the ring keeps the tail of the literal (whole words where possible)
      const uint32_t keep = len < uint32_t(kHist) ? len : uint32_t(kHist);
      const uint32_t k0 = o + len - keep, k1 = o + len;
      const uint32_t a0 = (k0 + 7) & ~7u, a1 = k1 & ~7u;
      if (a0 < a1) {
        for (uint32_t w = (a0 >> 3) + lane; w < (a1 >> 3); w += 32) sm.ring64[w & (kRing / 8 - 1)] = ld8_any(lsrc + ((w << 3) - o));
        if (k0 + lane < a0) ring[(k0 + lane) & (kRing - 1)] = snp_ldg8(lsrc + (k0 + lane - o));
        if (a1 + lane < k1) ring[(a1 + lane) & (kRing - 1)] = snp_ldg8(lsrc + (a1 + lane - o));
      } else {
        for (uint32_t i = k0 + lane; i < k1; i += 32) ring[i & (kRing - 1)] = snp_ldg8(lsrc + (i - o));
      }
      snp_syncwarp();
      pos += 1 + nb + len;
      o += len;
      fl = o & ~31u;
      continue;
    }
    // ---- build the jump tables.  Lane l owns the 8 positions [8l, 8l+8): their tag bytes are the window word it
    //      just loaded.  J[lv][p] = start of the 2^lv-th element after the one at p, kExit when that leaves the window.  Positions
    //      behind the end of the stream are staged as long-literal tags (csz 255): every chain ends there, and a lookup that lands
    //      on one finds an element that can never be part of a batch.  csz[...] + p saturates at kExit, so J[lv][kExit] == kExit on
    //      every level and the lookups need no test.
    uint32_t jlo, jhi;
    {
      uint64_t w = kFill;
      const int nv = int(avail) - lane * 8;                       // stream bytes in this lane's word
      if (nv > 0) {
        w = sm_ld8(sm, wbase + uint32_t(lane) * 8);
        if (nv < 8) w = (w & ((1ull << (8 * nv)) - 1)) | (kFill << (8 * nv));
      }
      const uint32_t wl = uint32_t(w), wh = uint32_t(w >> 32), p0 = uint32_t(lane) * 8;
      uint32_t a;
      jlo = 0; jhi = 0;
      a = p0 + 0 + csz[byte_of<0>(wl)]; jlo = put_byte<0>(jlo, a < kExit ? a : kExit);
      a = p0 + 1 + csz[byte_of<1>(wl)]; jlo = put_byte<1>(jlo, a < kExit ? a : kExit);
      a = p0 + 2 + csz[byte_of<2>(wl)]; jlo = put_byte<2>(jlo, a < kExit ? a : kExit);
      a = p0 + 3 + csz[byte_of<3>(wl)]; jlo = put_byte<3>(jlo, a < kExit ? a : kExit);
      a = p0 + 4 + csz[byte_of<0>(wh)]; jhi = put_byte<0>(jhi, a < kExit ? a : kExit);
      a = p0 + 5 + csz[byte_of<1>(wh)]; jhi = put_byte<1>(jhi, a < kExit ? a : kExit);
      a = p0 + 6 + csz[byte_of<2>(wh)]; jhi = put_byte<2>(jhi, a < kExit ? a : kExit);
      a = p0 + 7 + csz[byte_of<3>(wh)]; jhi = put_byte<3>(jhi, a < kExit ? a : kExit);
      reinterpret_cast<uint2*>(sm.J[0])[lane] = make_uint2(jlo, jhi);
    }
    snp_syncwarp();
    jt_level<1>(sm, smbase, jlo, jhi, lane);
    jt_level<2>(sm, smbase, jlo, jhi, lane);
    jt_level<3>(sm, smbase, jlo, jhi, lane);
    jt_level<4>(sm, smbase, jlo, jhi, lane);
    uint32_t qs = 0;                                   // window-relative start of the next batch
    bool first = true;
    for (;;) {
      uint32_t q = qs;
#pragma unroll
      for (int lv = 0; lv < 5; lv++)
        if ((lane >> lv) & 1) q = sm.J[lv][q];
      // ---- decode the element at q (branch-free: tag table + the 4 bytes behind the tag)
      const uint32_t e = lut[sm8[wbase + q]];
      uint32_t len = e & 0x7fu;
      uint32_t ecsz = e >> 24;
      bool is_lit = (e >> 16) & 1u;
      uint32_t off;
      {
        const uint32_t p1 = wbase + q + 1;
        const uint32_t raw = snp_funnel_r(sm32[p1 >> 2], sm32[(p1 >> 2) + 1], (p1 & 3) * 8);
        off = is_lit ? 0u : ((raw & (0xffffffffu >> ((e >> 18) & 31u))) | (e & 0x700u));
      }
      // a long literal ends the batch (straight-copy path); a truncated element is caught by m == 0 / the final size check
      const bool valid = q != kExit && !((e >> 17) & 1u) && q + ecsz <= avail;
      const unsigned vm = snp_ballot(valid);
      const int m = (vm == 0xffffffffu) ? 32 : (snp_ffs(~vm) - 1);   // valid lanes form a prefix
      if (m == 0) {
        if (first) { if (lane == 0) snp_set_err(err, 105); return; }
        pos += qs;                                                 // a long literal (or the window's end) starts here: restage
        break;
      }
      first = false;
      if (lane >= m) { len = 0; ecsz = 0; off = 1; is_lit = true; }
      uint32_t inc = len;                                          // inclusive prefix sum of the output lengths
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t t = snp_shfl_up(inc, d); if (lane >= d) inc += t; }
      const uint32_t doff = inc - len;
      {
        const bool bad = lane < m && !is_lit && (off == 0 || off > o + doff);
        if (snp_any(bad)) { if (lane == 0) snp_set_err(err, 103); return; }
      }
      const uint32_t len0 = snp_shfl(len, 0);
      int cnt;                                                     // elements executed by this step (a prefix of the batch)
      uint32_t T;                                                  // their output bytes
      if (len0 <= 8) {
        // ---------------- word mode
        // where does this element's data come from?  0 literal bytes in the window, 1 earlier output (ring / global),
        // 2 one earlier element of this batch (parent, byte delta).  Anything else ends the prefix.
        uint32_t skind = 0, spos = 0;
        bool fail = len > 8;
        const uint32_t s0 = doff - off;                            // source start relative to the batch (meaningful for skind 2)
        if (!is_lit && !fail) {
          if (off >= len) {
            if (s0 + len - 1 >= 0x80000000u) { skind = 1; spos = o + s0; }      // ends before the batch starts
            else if (s0 >= 0x80000000u) fail = true;                          // straddles the batch start
            else skind = 2;
          } else {                                                 // self-overlapping (periodic) copy: fine if its pattern is old
            if (doff == 0) { skind = 1; spos = o - off; } else fail = true;
          }
        }
        // issue the loads of the old / literal sources now: the parent search below hides their latency
        uint64_t w = 0;
        if (lane < m && !fail && skind != 2) {
          const uint32_t r = spos & (kRing - 1);
          if (skind == 1 && o - spos > uint32_t(kHist)) w = ld8_out(dst + spos);
          else if (skind == 1 && r > uint32_t(kRing) - 8) w = ring_ld8(sm, spos);
          else w = sm_ld8(sm, skind == 0 ? wbase + q + 1 : r);
          if (skind == 1 && off < len) {                           // periodic: repeat the first `off` bytes
            w &= (off >= 8) ? ~0ull : ((1ull << (8 * off)) - 1);
            for (uint32_t f = off; f < 8; f <<= 1) w |= w << (8 * f);
          }
        }
        // parent = the element that contains byte s0.  Fixed-width columns compress to elements of ~4 bytes (literal + copy per
        // value, offsets that are multiples of the width), so the element off/4 places back is the first guess; pp = parent lane |
        // byte delta << 8
        uint32_t pp = uint32_t(lane);
        bool need = false;
        {
          const int cand = lane - int(off >> 2);
          const uint32_t x = snp_shfl(doff | (len << 16), cand);
          if (skind == 2) {
            const uint32_t d = s0 - (x & 0xffffu);
            if (cand >= 0 && cand < lane && d < 0x10000u && d + len <= (x >> 16)) pp = uint32_t(cand) | (d << 8);
            else need = true;
          }
        }
        if (snp_any(need)) {
          SNP_STAT(parent_searches, 1);
          // general case: binary search over the element starts (register shuffles only)
          int lo = 0;
#pragma unroll
          for (int step = 16; step > 0; step >>= 1) {
            const int cand = lo + step;
            const uint32_t d = snp_shfl(doff, cand & 31);
            if (need && cand < lane && d <= s0) lo = cand;
          }
          const uint32_t pd = snp_shfl(doff, lo), pl = snp_shfl(len, lo);
          if (need) {
            if (lo >= lane || s0 < pd || s0 + len > pd + pl) fail = true;     // not inside ONE earlier element
            else pp = uint32_t(lo) | ((s0 - pd) << 8);
          }
        }
        const unsigned fm = snp_ballot(fail || lane >= m);
        cnt = fm ? snp_ffs(fm) - 1 : 32;                           // >= 1: element 0 has len <= 8 and an old / literal source
        if (lane >= cnt) pp = uint32_t(lane);
        // collapse parent chains (parents are always earlier lanes inside the prefix): parent' = parent's parent, delta' = sum
#pragma unroll
        for (int it = 0; it < 5; it++) pp = snp_shfl(pp, int(pp)) + (pp & ~0xffu);      // (the shuffle takes the lane modulo 32)
        {
          const uint64_t wr = shfl64(w, int(pp));
          if (skind == 2) w = wr >> (8 * (pp >> 8));
        }
        T = snp_shfl(inc, cnt - 1);
        if (o + T > ulen) { if (lane == 0) snp_set_err(err, 103); return; }
        if (lane < cnt) {
          const uint32_t rb = (o + doff) & (kRing - 1);
          if (rb <= uint32_t(kRing) - 8) store_elem(sm, smbase, rb, w, len);   // no wrap inside the element
          else {
#pragma unroll
            for (int i = 0; i < 8; i++)
              if (uint32_t(i) < len) ring[(rb + i) & (kRing - 1)] = uint8_t(w >> (8 * i));
          }
        }
      } else {
        // ---------------- run mode: element 0 is long.  A literal goes alone; a copy takes every following copy with the
        // same offset along (one periodic pattern over the union).
        const uint32_t off0 = snp_shfl(off, 0);
        const bool lit0 = snp_shfl(uint32_t(is_lit), 0) != 0;
        if (lit0) cnt = 1;
        else {
          const unsigned brk = snp_ballot(lane >= m || is_lit || off != off0);
          cnt = brk ? snp_ffs(brk) - 1 : 32;
        }
        T = snp_shfl(inc, cnt - 1);
        if (o + T > ulen) { if (lane == 0) snp_set_err(err, 103); return; }
        snp_syncwarp();
        if (lit0) {
          const uint32_t q0 = snp_shfl(q, 0);
          const uint8_t* ls = src + pos + q0 + 1;
          for (uint32_t i = lane; i < T; i += 32) ring[(o + i) & (kRing - 1)] = snp_ldg8(ls + i);
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
      snp_syncwarp();
      SNP_STAT(steps, 1); SNP_STAT(elements, cnt); SNP_STAT(word_steps, len0 <= 8 ? 1 : 0); SNP_STAT(bytes, T);
      o += T;
      if (o - fl >= kFlushAt) fl = flush_words(sm, dst, fl, o, lane);
      // where the next batch starts: right after the last executed element
      const uint32_t adv = snp_shfl(q + ecsz, cnt - 1);
      if (adv > kRestage || adv >= avail || o >= stop_at) { pos += adv; break; }      // (kRestage + 5 <= kWin: the next header is staged)
      qs = adv;
      snp_syncwarp();
    }
  }
  snp_syncwarp();
  fl = flush_words(sm, dst, fl, o, lane);
  if (fl + lane < o) dst[fl + lane] = ring[(fl + lane) & (kRing - 1)];
  if (o != ulen && o < stop_at) { if (lane == 0) snp_set_err(err, 104); }      // the stream ended before the bytes the consumer needs
}

// One page.  st.phase carries the barriers' phases from page to page; no copy is in flight on return.
SNP_FN void snappy_page(const uint8_t* __restrict__ src, uint32_t n, uint8_t* __restrict__ dst, uint32_t ulen_expected, uint32_t stop_at,
                        WarpSmem& sm, uint32_t& phase, const uint8_t* __restrict__ csz, const uint32_t* __restrict__ lut, int lane, int* err) {
  StageState st;
  st.phase = phase; st.cur = 0; st.pf = false; st.pf_start = 0; st.pf_bytes = 0;
  snappy_page_body(src, n, dst, ulen_expected, stop_at, sm, st, csz, lut, lane, err);
  snp_syncwarp();
  if (st.pf) {                                         // drain the look-ahead nobody came to use
    const int nb = st.cur ^ 1;
    bulk_wait(sm, nb, (st.phase >> nb) & 1u);
    st.phase ^= 1u << nb;
  }
  phase = st.phase;
}

}  // namespace snp
}  // namespace horae
