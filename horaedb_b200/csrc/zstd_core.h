// zstd_core.h — warp-level Zstandard frame decoder for Parquet pages written with `ParquetCompression::Zstd`
// (config.rs:78-94 -> parquet 53.2 -> zstd 0.13.2 / libzstd 1.5, Cargo.lock:3893; the format is restated from RFC 8878).
//
// Like snappy_core.h this header is ONE source: zstd.cu compiles it for sm_100a, tests/emu/zstd_emu.cpp compiles the same text with the
// 32 lanes as coroutines, so the CPU test-suite runs the device decoder on real libzstd streams.  The includer provides the same
// primitives as for snappy_core.h: SNP_FN, snp_syncwarp(), snp_any(pred), snp_ldg8(p), snp_ldg64u(p) (8 input bytes at any alignment),
// snp_ldcg8(p), snp_set_err(err, code), and SNP_CONST (qualifier of a namespace-scope constant table).  Input buffers carry >= 8 readable bytes behind their last byte (SST bytes: +64, transient ranges: +16).
//
// One warp owns one page (= one frame).  Zstandard is two serial entropy decoders per block — Huffman literals, then FSE-coded
// (literal length, match length, offset) sequences whose copies may overlap their own output — so the warp splits the work by
// what is serial and what is not:
//   tables     lane 0 reads the table descriptions and builds the FSE / Huffman decoding tables in shared memory;
//   literals   the (up to) four Huffman streams of a block are independent: lanes 0..3 decode one each into the page's literal
//              buffer (global scratch behind the output);
//   sequences  every lane runs the same scalar FSE state machine on the same bits (broadcast loads, no shuffles), and the two copies
//              of a sequence — literals, then the match, which may overlap itself — are spread over the 32 lanes.
// Raw and RLE blocks / literal sections are plain warp copies / fills.  Throughput is far below the Snappy path's (one sequence per
// pass instead of 32 elements): the codec is supported on the general pipeline, it is not the bench's path.
#pragma once
#include <cstdint>

namespace horae {
namespace zst {

constexpr int kMaxLL = 36, kMaxML = 53, kMaxOF = 32;       // symbols of the three sequence alphabets
constexpr int kLLLog = 9, kMLLog = 9, kOFLog = 8;          // maximum accuracy logs
constexpr int kHufLog = 11;                                // maximum Huffman code length
constexpr uint32_t kBlockMax = 128u << 10;
constexpr uint32_t kRingZ = 4096;      // output history kept in shared memory: a match that starts within it never waits for L2

struct alignas(16) WarpSmem {
  uint32_t ll[1 << kLLLog], ml[1 << kMLLog], of[1 << kOFLog];   // FSE cells: symbol | nbits << 8 | new-state base << 16
  uint16_t huf[1 << kHufLog];                                  // Huffman cells: symbol | nbits << 8
  int16_t norm[256];                                           // scratch: normalised counts / Huffman weights
  uint8_t sym_of_cell[1 << kLLLog];                            // scratch of the FSE table build
  uint16_t state_desc[64];
  uint32_t ll_log, ml_log, of_log, huf_log;                    // accuracy logs of the current tables (they persist across blocks)
  uint32_t rep[3];                                             // repeat offsets (persist across the blocks of a frame)
  uint32_t huf_ok;
  uint8_t per[32];                                             // the period of a short-offset match
  alignas(8) uint8_t ring[kRingZ];                                        // the most recent output bytes (byte at output position x: x & (kRingZ-1))
  uint32_t wt[128];                                            // FSE table of a compressed Huffman tree description
  uint32_t xfer[2];                                            // lane 0 -> warp: bytes consumed by a table description
};

// Constant tables live at namespace scope (SNP_CONST: __constant__ memory on the device — every lane reads the same entry, a broadcast):
// as function-local arrays the compiler rebuilt them on the thread's stack at every call.
// literal length / match length codes -> (baseline, extra bits)
SNP_CONST uint32_t kLLBase[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
SNP_CONST uint8_t kLLBits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
SNP_CONST uint32_t kMLBase[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
                                  35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
SNP_CONST uint8_t kMLBits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
// predefined distributions (RFC 8878 section 3.1.1.3.2.2)
SNP_CONST int8_t kLLDefault[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
SNP_CONST int8_t kMLDefault[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                   1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
SNP_CONST int8_t kOFDefault[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
SNP_FN uint32_t ll_base(uint32_t c) { return kLLBase[c]; }
SNP_FN uint32_t ll_bits(uint32_t c) { return kLLBits[c]; }
SNP_FN uint32_t ml_base(uint32_t c) { return kMLBase[c]; }
SNP_FN uint32_t ml_bits(uint32_t c) { return kMLBits[c]; }
SNP_FN int ll_default(int s) { return kLLDefault[s]; }
SNP_FN int ml_default(int s) { return kMLDefault[s]; }
SNP_FN int of_default(int s) { return kOFDefault[s]; }

SNP_FN int highbit(uint32_t v) { int r = -1; while (v) { v >>= 1; r++; } return r; }      // index of the highest set bit, -1 for 0

// 8 little-endian bytes starting at p.  Bytes at or behind `end` may be anything: every caller masks what it takes to bits that lie
// inside its stream (the readable slack behind the input makes the load itself safe).
SNP_FN uint64_t ld_le(const uint8_t* p, const uint8_t* end) { (void)end; return snp_ldg64u(p); }

// ---- forward bit reader (FSE table descriptions): LSB first
struct FwdBits { const uint8_t* p; const uint8_t* end; uint32_t bit; };
SNP_FN uint32_t fwd_read(FwdBits& b, int n) {                  // n <= 24
  const uint8_t* q = b.p + (b.bit >> 3);
  const uint64_t v = ld_le(q, b.end) >> (b.bit & 7);
  b.bit += uint32_t(n);
  return uint32_t(v) & ((1u << n) - 1u);
}

// ---- backward bit reader (Huffman and FSE streams): the stream is read from its last byte down; `off` = bits not yet consumed, may go
//      negative (reads below the stream's start deliver zero bits, RFC 8878 4.1 / 4.2.1)
struct BwdBits { const uint8_t* p; const uint8_t* end; int64_t off; uint64_t cache; int64_t cache_bit; };
SNP_FN bool bwd_init(BwdBits& b, const uint8_t* p, uint32_t len) {
  b.p = p; b.end = p + len; b.off = 0; b.cache = 0; b.cache_bit = int64_t(1) << 40;      // (empty cache)
  if (len == 0) return false;
  const uint32_t last = snp_ldg8(p + len - 1);
  if (last == 0) return false;                                 // the final byte carries the end mark
  b.off = int64_t(len) * 8 - (8 - highbit(last));
  return true;
}
// The reader keeps the 8 stream bytes that END at the current position in a register: reads walk down through them, one load serves
// ~50 bits (two sequences' worth), and the dependent chain of a sequence no longer contains a memory access per field.
SNP_FN uint32_t bwd_read(BwdBits& b, int n) {                  // n <= 32
  if (n == 0) return 0;
  b.off -= n;
  if (b.off >= b.cache_bit && b.off + n <= b.cache_bit + 64) return uint32_t(b.cache >> (b.off - b.cache_bit)) & uint32_t((1ull << n) - 1ull);
  if (b.off >= 0) {
    int64_t byte = ((b.off + n + 7) >> 3) - 8;                 // the word ends at the byte holding the top bit of this read
    if (byte < 0) byte = 0;
    b.cache = ld_le(b.p + byte, b.end);
    b.cache_bit = byte * 8;
    return uint32_t(b.cache >> (b.off - b.cache_bit)) & uint32_t((1ull << n) - 1ull);
  }
  // below the start of the stream: the missing low bits read as zero
  int64_t o = 0;
  int bits = n + int(b.off);
  uint64_t v = 0;
  if (bits > 0) {
    v = ld_le(b.p + (o >> 3), b.end) >> (o & 7);
    v &= (1ull << bits) - 1ull;
  }
  v = (-b.off >= 32) ? 0 : (v << (-b.off));
  return uint32_t(v);
}

// ---- FSE: read a table description (RFC 8878 4.1.1) into sm.norm; returns the bytes consumed, 0 on error
SNP_FN uint32_t fse_read_norm(WarpSmem& sm, const uint8_t* p, const uint8_t* end, int max_log, int max_syms, uint32_t* log_out, int* nsym_out) {
  FwdBits b{p, end, 0};
  const int al = int(fwd_read(b, 4)) + 5;
  if (al > max_log) return 0;
  int remaining = 1 << al, s = 0;
  while (remaining > 0 && s < max_syms) {
    const int bits = highbit(uint32_t(remaining + 1)) + 1;
    uint32_t val = fwd_read(b, bits);
    const uint32_t lower = (1u << (bits - 1)) - 1u;
    const uint32_t threshold = (1u << bits) - 1u - uint32_t(remaining + 1);
    if ((val & lower) < threshold) { b.bit -= 1; val &= lower; }
    else if (val > lower) val -= threshold;
    const int proba = int(val) - 1;
    remaining -= proba < 0 ? -proba : proba;
    sm.norm[s++] = int16_t(proba);
    if (proba == 0) {
      uint32_t rep = fwd_read(b, 2);
      for (;;) {
        for (uint32_t i = 0; i < rep && s < max_syms; i++) sm.norm[s++] = 0;
        if (rep == 3) rep = fwd_read(b, 2); else break;
      }
    }
    if (b.p + (b.bit >> 3) > end) return 0;
  }
  if (remaining != 0) return 0;
  *log_out = uint32_t(al);
  *nsym_out = s;
  return (b.bit + 7) >> 3;
}
// build the decoding table (cells: symbol | nbits << 8 | base << 16) of `nsym` symbols from sm.norm (RFC 8878 4.1.1)
SNP_FN bool fse_build(WarpSmem& sm, uint32_t* tbl, int al, int nsym) {
  const int size = 1 << al;
  int high = size;
  if (nsym > 64) return false;
  for (int s = 0; s < nsym; s++) if (sm.norm[s] == -1) { sm.sym_of_cell[--high] = uint8_t(s); sm.state_desc[s] = 1; }
  const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
  int pos = 0;
  for (int s = 0; s < nsym; s++) {
    if (sm.norm[s] <= 0) continue;
    sm.state_desc[s] = uint16_t(sm.norm[s]);
    for (int i = 0; i < sm.norm[s]; i++) {
      sm.sym_of_cell[pos] = uint8_t(s);
      do { pos = (pos + step) & mask; } while (pos >= high);
    }
  }
  if (pos != 0) return false;
  for (int i = 0; i < size; i++) {
    const uint32_t s = sm.sym_of_cell[i];
    const uint32_t d = sm.state_desc[s]++;
    const uint32_t nb = uint32_t(al - highbit(d));
    const uint32_t base = (d << nb) - uint32_t(size);
    tbl[i] = s | (nb << 8) | (base << 16);
  }
  return true;
}
SNP_FN void fse_rle(uint32_t* tbl, uint32_t sym) { tbl[0] = sym; }      // one cell: nbits 0, base 0

// ---- Huffman: weights -> decoding table.  nw explicit weights in sm.norm; the last weight is implied (RFC 8878 4.2.1)
SNP_FN bool huf_build(WarpSmem& sm, int nw) {
  uint32_t sum = 0;
  for (int i = 0; i < nw; i++) { const int w = sm.norm[i]; if (w < 0 || w > kHufLog) return false; if (w) sum += 1u << (w - 1); }
  if (sum == 0) return false;
  const int max_bits = highbit(sum) + 1;
  if (max_bits > kHufLog) return false;
  const uint32_t left = (1u << max_bits) - sum;
  if (left & (left - 1)) return false;                         // must be a power of two
  sm.norm[nw] = int16_t(highbit(left) + 1);
  const int n = nw + 1;
  // codes of equal length are consecutive, the longest codes come first in the table
  uint32_t rank_cnt[kHufLog + 2];
  for (int i = 0; i <= kHufLog + 1; i++) rank_cnt[i] = 0;
  for (int i = 0; i < n; i++) { const int w = sm.norm[i]; if (w) rank_cnt[max_bits + 1 - w]++; }      // by code length
  uint32_t rank_idx[kHufLog + 2];
  rank_idx[max_bits] = 0;
  for (int i = max_bits; i >= 1; i--) rank_idx[i - 1] = rank_idx[i] + rank_cnt[i] * (1u << (max_bits - i));
  if (rank_idx[0] != (1u << max_bits)) return false;
  for (int i = 0; i < n; i++) {
    const int w = sm.norm[i];
    if (!w) continue;
    const int bits = max_bits + 1 - w;
    const uint32_t len = 1u << (max_bits - bits);
    const uint32_t code = rank_idx[bits];
    for (uint32_t j = 0; j < len; j++) sm.huf[code + j] = uint16_t(uint32_t(i) | (uint32_t(bits) << 8));
    rank_idx[bits] += len;
  }
  sm.huf_log = uint32_t(max_bits);
  return true;
}
// tree description at p: returns the bytes consumed (0 on error) and leaves the table in sm.huf
SNP_FN uint32_t huf_read_tree(WarpSmem& sm, const uint8_t* p, const uint8_t* end) {
  if (p >= end) return 0;
  const uint32_t hb = snp_ldg8(p);
  int nw = 0;
  uint32_t used;
  if (hb >= 128) {                                             // direct: 4-bit weights
    nw = int(hb) - 127;
    const uint32_t nbytes = uint32_t(nw + 1) / 2;
    if (p + 1 + nbytes > end) return 0;
    for (int i = 0; i < nw; i++) {
      const uint32_t b = snp_ldg8(p + 1 + i / 2);
      sm.norm[i] = int16_t((i & 1) ? (b & 15) : (b >> 4));
    }
    used = 1 + nbytes;
  } else {                                                     // FSE-compressed weights: two interleaved states
    if (hb == 0 || p + 1 + hb > end) return 0;
    uint32_t al; int nsym;
    const uint32_t hdr = fse_read_norm(sm, p + 1, p + 1 + hb, 7, 32, &al, &nsym);
    if (!hdr || hdr > hb) return 0;
    if (!fse_build(sm, sm.wt, int(al), nsym)) return 0;
    BwdBits b;
    if (!bwd_init(b, p + 1 + hdr, hb - hdr)) return 0;
    uint32_t s1 = bwd_read(b, int(al)), s2 = bwd_read(b, int(al));
    // the weights land in a local array first: sm.norm holds the FSE counts the table was built from
    uint8_t wts[256];
    for (;;) {
      if (nw >= 254) return 0;
      uint32_t c = sm.wt[s1];
      wts[nw++] = uint8_t(c);
      s1 = (c >> 16) + bwd_read(b, int((c >> 8) & 0xff));
      if (b.off < 0) { wts[nw++] = uint8_t(sm.wt[s2]); break; }
      c = sm.wt[s2];
      wts[nw++] = uint8_t(c);
      s2 = (c >> 16) + bwd_read(b, int((c >> 8) & 0xff));
      if (b.off < 0) { wts[nw++] = uint8_t(sm.wt[s1]); break; }
    }
    for (int i = 0; i < nw; i++) sm.norm[i] = wts[i];
    used = 1 + hb;
  }
  if (nw < 1 || nw > 255) return 0;
  if (!huf_build(sm, nw)) return 0;
  return used;
}
// one Huffman stream: `count` symbols to out
SNP_FN bool huf_stream(const WarpSmem& sm, const uint8_t* p, uint32_t len, uint8_t* out, uint32_t count) {
  BwdBits b;
  if (!bwd_init(b, p, len)) return false;
  const int mb = int(sm.huf_log);
  const uint32_t mask = (1u << mb) - 1u;
  uint32_t st = bwd_read(b, mb);
  for (uint32_t i = 0; i < count; i++) {
    const uint32_t c = sm.huf[st];
    out[i] = uint8_t(c);
    const int nb = int(c >> 8);
    st = ((st << nb) + bwd_read(b, nb)) & mask;
  }
  return b.off == -int64_t(mb);                                // the stream is consumed exactly
}

// warp copies to output position `pos` (every output byte also lands in the ring).  in: read-only input; buf: the literal buffer
SNP_FN void copy_in(WarpSmem& sm, uint8_t* out, uint32_t pos, const uint8_t* src, uint32_t n, int lane) {
  uint32_t done = 0;
  if (n >= 256 && (reinterpret_cast<uintptr_t>(out) & 7) == 0) {          // (output position and address share their alignment)
    // long runs (raw blocks, raw literal sections of incompressible columns): bytes up to the first 8-aligned output address, then 8 bytes per lane
    const uint32_t head = uint32_t((8 - (reinterpret_cast<uintptr_t>(out + pos) & 7)) & 7);
    if (uint32_t(lane) < head) { const uint8_t v = snp_ldg8(src + lane); out[pos + lane] = v; if (n - uint32_t(lane) <= kRingZ) sm.ring[(pos + lane) & (kRingZ - 1)] = v; }
    const uint32_t nwords = (n - head) >> 3;
    uint64_t* d8 = reinterpret_cast<uint64_t*>(out + pos + head);
    for (uint32_t w = lane; w < nwords; w += 32) {
      const uint64_t v = snp_ldg64u(src + head + (size_t(w) << 3));
      d8[w] = v;
      const uint32_t at = head + (w << 3);
      if (n - at <= kRingZ) *reinterpret_cast<uint64_t*>(&sm.ring[(pos + at) & (kRingZ - 1)]) = v;       // (pos + at is 8-aligned, like the ring)
    }
    done = head + (nwords << 3);
  }
  for (uint32_t i = done + lane; i < n; i += 32) { const uint8_t v = snp_ldg8(src + i); out[pos + i] = v; if (n - i <= kRingZ) sm.ring[(pos + i) & (kRingZ - 1)] = v; }
}
SNP_FN void copy_buf(WarpSmem& sm, uint8_t* out, uint32_t pos, const uint8_t* src, uint32_t n, int lane) {
  for (uint32_t i = lane; i < n; i += 32) { const uint8_t v = snp_ldcg8(src + i); out[pos + i] = v; if (n - i <= kRingZ) sm.ring[(pos + i) & (kRingZ - 1)] = v; }
}
SNP_FN void fill(WarpSmem& sm, uint8_t* out, uint32_t pos, uint8_t v, uint32_t n, int lane) {
  for (uint32_t i = lane; i < n; i += 32) { out[pos + i] = v; if (n - i <= kRingZ) sm.ring[(pos + i) & (kRingZ - 1)] = v; }
}
// byte at output position x < `written`: from the ring if it is among the last kRingZ / 2 bytes (a pass writes at most kRingZ / 2 new
// bytes, so those slots are not overwritten while it reads), else from global memory (coherent load)
SNP_FN uint8_t out_byte(const WarpSmem& sm, const uint8_t* out, uint32_t written, uint32_t x) {
  return (written - x <= kRingZ / 2) ? sm.ring[x & (kRingZ - 1)] : snp_ldcg8(out + x);
}
// out[pos + i] = out[pos + i - off] for i < n, overlapping allowed: byte i comes from the period out[pos - off .. pos)
SNP_FN void copy_match(WarpSmem& sm, uint8_t* out, uint32_t pos, uint32_t off, uint32_t n, int lane) {
  if (off >= n || off >= 32) {
    // passes of at most `off` bytes (a multiple of 32) never read what the same pass writes
    const uint32_t chunk = off < n ? (off / 32u) * 32u : n;      // off >= 32 here when off < n
    uint32_t done = 0;
    while (done < n) {
      uint32_t m = n - done < chunk ? n - done : chunk;
      if (m > kRingZ / 2) m = kRingZ / 2;                        // a pass must not overwrite ring bytes it still has to read
      for (uint32_t i = lane; i < m; i += 32) {
        const uint8_t v = out_byte(sm, out, pos + done, pos + done + i - off);
        out[pos + done + i] = v;
        sm.ring[(pos + done + i) & (kRingZ - 1)] = v;
      }
      done += m;
      snp_syncwarp();
    }
  } else {
    // short period: every byte of the match is a byte of out[pos - off .. pos): park the period, then spread it
    if (uint32_t(lane) < off) sm.per[lane] = out_byte(sm, out, pos, pos - off + uint32_t(lane));
    snp_syncwarp();
    for (uint32_t i = lane; i < n; i += 32) {
      const uint8_t v = sm.per[i % off];
      out[pos + i] = v;
      if (n - i <= kRingZ) sm.ring[(pos + i) & (kRingZ - 1)] = v;
    }
  }
}

// Decodes the frame(s) in [src, src + n) to dst (ulen bytes expected).  lit = literal buffer (global scratch, >= min(ulen, 128 KiB) + 32).
SNP_FN void zstd_page(const uint8_t* __restrict__ src, uint32_t n, uint8_t* __restrict__ dst, uint32_t ulen, uint8_t* __restrict__ lit,
                      WarpSmem& sm, int lane, int* err) {
#define ZFAIL(code) do { if (lane == 0) snp_set_err(err, (code)); return; } while (0)
  const uint8_t* const end = src + n;
  const uint8_t* p = src;
  uint32_t o = 0;
  while (p < end) {
    // ---- frame header
    if (p + 5 > end) ZFAIL(201);
    const uint32_t magic = uint32_t(ld_le(p, end));
    if ((magic & 0xfffffff0u) == 0x184d2a50u) {                 // skippable frame
      const uint32_t sz = uint32_t(ld_le(p + 4, end));
      if (p + 8 + sz > end) ZFAIL(201);
      p += 8 + sz;
      continue;
    }
    if (magic != 0xfd2fb528u) ZFAIL(202);
    const uint32_t fhd = snp_ldg8(p + 4);
    p += 5;
    const uint32_t fcs_flag = fhd >> 6, single = (fhd >> 5) & 1u, checksum = (fhd >> 2) & 1u, did = fhd & 3u;
    if (fhd & 0x08u) ZFAIL(203);                               // reserved bit
    if (!single) p += 1;                                       // window descriptor (the output buffer is the window)
    p += did == 3 ? 4 : did;
    if (did) ZFAIL(204);                                       // dictionaries are not used by Parquet pages
    p += fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (p > end) ZFAIL(201);
    if (lane == 0) { sm.rep[0] = 1; sm.rep[1] = 4; sm.rep[2] = 8; sm.huf_ok = 0; sm.ll_log = sm.ml_log = sm.of_log = 0xffu; }
    snp_syncwarp();
    // ---- blocks
    for (;;) {
      if (p + 3 > end) ZFAIL(205);
      const uint32_t bh = uint32_t(ld_le(p, end)) & 0xffffffu;
      p += 3;
      const uint32_t last = bh & 1u, type = (bh >> 1) & 3u, bsize = bh >> 3;
      if (type == 0) {                                         // raw
        if (p + bsize > end || o + bsize > ulen) ZFAIL(206);
        copy_in(sm, dst, o, p, bsize, lane);
        p += bsize; o += bsize;
      } else if (type == 1) {                                  // RLE
        if (p + 1 > end || o + bsize > ulen) ZFAIL(206);
        fill(sm, dst, o, snp_ldg8(p), bsize, lane);
        p += 1; o += bsize;
      } else if (type == 2) {
        if (bsize > kBlockMax || p + bsize > end || bsize < 2) ZFAIL(207);
        const uint8_t* const bend = p + bsize;
        // ---- literals section
        const uint32_t b0 = snp_ldg8(p);
        const uint32_t ltype = b0 & 3u, sf = (b0 >> 2) & 3u;
        uint32_t regen, csize = 0, nstreams = 1, hdr;
        if (ltype < 2) {
          if ((sf & 1u) == 0) { regen = b0 >> 3; hdr = 1; }
          else if (sf == 1) { regen = (uint32_t(ld_le(p, bend)) & 0xffffu) >> 4; hdr = 2; }
          else { regen = (uint32_t(ld_le(p, bend)) & 0xffffffu) >> 4; hdr = 3; }
        } else {
          const uint64_t h = ld_le(p, bend);
          if (sf == 0 || sf == 1) { regen = uint32_t(h >> 4) & 0x3ffu; csize = uint32_t(h >> 14) & 0x3ffu; hdr = 3; nstreams = sf == 0 ? 1 : 4; }
          else if (sf == 2) { regen = uint32_t(h >> 4) & 0x3fffu; csize = uint32_t(h >> 18) & 0x3fffu; hdr = 4; nstreams = 4; }
          else { regen = uint32_t(h >> 4) & 0x3ffffu; csize = uint32_t(h >> 22) & 0x3ffffu; hdr = 5; nstreams = 4; }
        }
        if (regen > kBlockMax || regen > ulen - o) ZFAIL(208);   // literals all end up in the output; also bounds the literal buffer (min(page, 128 KB))
        const uint8_t* lp = p + hdr;                           // literal payload
        const uint8_t* lits;                                   // where the block's literals are read from
        bool lits_in_input = false;
        snp_syncwarp();
        if (ltype == 0) {
          if (lp + regen > bend) ZFAIL(208);
          lits = lp; lits_in_input = true;
          p = lp + regen;
        } else if (ltype == 1) {
          if (lp + 1 > bend) ZFAIL(208);
          for (uint32_t i = lane; i < regen; i += 32) lit[i] = snp_ldg8(lp);
          lits = lit;
          p = lp + 1;
        } else {
          if (lp + csize > bend) ZFAIL(208);
          const uint8_t* hp = lp;
          uint32_t tree = 0;
          if (ltype == 2) {
            if (lane == 0) { tree = huf_read_tree(sm, hp, lp + csize); sm.huf_ok = tree ? 1u : 0u; sm.xfer[0] = tree; }
            snp_syncwarp();
            if (!sm.huf_ok) ZFAIL(209);
            tree = sm.xfer[0];
            hp += tree;
          } else if (!sm.huf_ok) ZFAIL(210);                   // treeless without an earlier tree
          const uint32_t total = csize - tree;
          bool ok = true;
          if (nstreams == 1) {
            if (lane == 0) ok = huf_stream(sm, hp, total, lit, regen);
          } else {
            if (total < 6) ZFAIL(211);
            const uint32_t s1 = uint32_t(ld_le(hp, bend)) & 0xffffu, s2 = uint32_t(ld_le(hp + 2, bend)) & 0xffffu, s3 = uint32_t(ld_le(hp + 4, bend)) & 0xffffu;
            if (6 + s1 + s2 + s3 > total) ZFAIL(211);
            const uint32_t s4 = total - 6 - s1 - s2 - s3;
            const uint32_t per = (regen + 3) / 4;
            if (3 * per > regen) ZFAIL(211);
            if (lane < 4) {
              const uint32_t so = lane == 0 ? 0 : (lane == 1 ? s1 : (lane == 2 ? s1 + s2 : s1 + s2 + s3));
              const uint32_t sl = lane == 0 ? s1 : (lane == 1 ? s2 : (lane == 2 ? s3 : s4));
              const uint32_t cnt = lane < 3 ? per : regen - 3 * per;
              ok = huf_stream(sm, hp + 6 + so, sl, lit + uint32_t(lane) * per, cnt);
            }
          }
          if (snp_any(!ok)) ZFAIL(212);
          lits = lit;
          p = lp + csize;
        }
        snp_syncwarp();
        // ---- sequences section
        if (p >= bend) ZFAIL(213);
        uint32_t nseq = snp_ldg8(p);
        if (nseq == 0) p += 1;
        else if (nseq < 128) p += 1;
        else if (nseq < 255) { if (p + 2 > bend) ZFAIL(213); nseq = ((nseq - 128) << 8) + snp_ldg8(p + 1); p += 2; }
        else { if (p + 3 > bend) ZFAIL(213); nseq = uint32_t(snp_ldg8(p + 1)) + (uint32_t(snp_ldg8(p + 2)) << 8) + 0x7f00u; p += 3; }
        uint32_t lpos = 0;                                     // literals consumed
        if (nseq) {
          if (p >= bend) ZFAIL(213);
          const uint32_t modes = snp_ldg8(p);
          p += 1;
          if (modes & 3u) ZFAIL(214);
          // tables: literal lengths, offsets, match lengths — lane 0 builds, everybody reads
          uint32_t used_total = 0;
          if (lane == 0) {
            bool ok = true;
            const uint8_t* q = p;
            for (int t = 0; t < 3 && ok; t++) {
              const uint32_t mode = (modes >> (6 - 2 * t)) & 3u;
              uint32_t* tbl = t == 0 ? sm.ll : (t == 1 ? sm.of : sm.ml);
              uint32_t* lg = t == 0 ? &sm.ll_log : (t == 1 ? &sm.of_log : &sm.ml_log);
              const int max_log = t == 0 ? kLLLog : (t == 1 ? kOFLog : kMLLog);
              const int max_sym = t == 0 ? kMaxLL : (t == 1 ? kMaxOF : kMaxML);
              if (mode == 0) {
                const int al = t == 1 ? 5 : 6, ns = t == 0 ? 36 : (t == 1 ? 29 : 53);
                for (int s = 0; s < ns; s++) sm.norm[s] = int16_t(t == 0 ? ll_default(s) : (t == 1 ? of_default(s) : ml_default(s)));
                ok = fse_build(sm, tbl, al, ns);
                *lg = uint32_t(al);
              } else if (mode == 1) {
                if (q >= bend) { ok = false; break; }
                const uint32_t sym = snp_ldg8(q);
                if (int(sym) >= max_sym) { ok = false; break; }
                fse_rle(tbl, sym);
                *lg = 0;
                q += 1;
              } else if (mode == 2) {
                uint32_t al; int ns;
                const uint32_t used = fse_read_norm(sm, q, bend, max_log, max_sym, &al, &ns);
                if (!used) { ok = false; break; }
                ok = fse_build(sm, tbl, int(al), ns);
                *lg = al;
                q += used;
              } else if (*lg == 0xffu) ok = false;             // repeat without an earlier table
            }
            used_total = ok ? uint32_t(q - p) + 1u : 0u;
            sm.xfer[1] = used_total;
          }
          snp_syncwarp();
          used_total = sm.xfer[1];
          if (!used_total) ZFAIL(215);
          p += used_total - 1;
          if (p >= bend) ZFAIL(215);
          // ---- the sequences: every lane runs the same state machine on the same bits
          BwdBits b;
          if (!bwd_init(b, p, uint32_t(bend - p))) ZFAIL(216);
          const uint32_t ll_log = sm.ll_log, of_log = sm.of_log, ml_log = sm.ml_log;
          uint32_t sl = bwd_read(b, int(ll_log)), so = bwd_read(b, int(of_log)), sml = bwd_read(b, int(ml_log));
          uint32_t r0 = sm.rep[0], r1 = sm.rep[1], r2 = sm.rep[2];
          for (uint32_t i = 0; i < nseq; i++) {
            const uint32_t cl = sm.ll[sl], co = sm.of[so], cm = sm.ml[sml];
            const uint32_t ofc = co & 0xffu, mlc = cm & 0xffu, llc = cl & 0xffu;
            if (ofc > 31 || mlc >= uint32_t(kMaxML) || llc >= uint32_t(kMaxLL)) ZFAIL(217);
            const uint32_t ofv = (1u << ofc) + bwd_read(b, int(ofc));
            const uint32_t mlen = ml_base(mlc) + bwd_read(b, int(ml_bits(mlc)));
            const uint32_t llen = ll_base(llc) + bwd_read(b, int(ll_bits(llc)));
            if (i + 1 < nseq) {
              sl = (cl >> 16) + bwd_read(b, int((cl >> 8) & 0xffu));
              sml = (cm >> 16) + bwd_read(b, int((cm >> 8) & 0xffu));
              so = (co >> 16) + bwd_read(b, int((co >> 8) & 0xffu));
            }
            if (b.off < 0) ZFAIL(218);
            uint32_t off;
            if (ofv > 3) { off = ofv - 3; r2 = r1; r1 = r0; r0 = off; }
            else {
              uint32_t idx = ofv - 1 + (llen == 0 ? 1u : 0u);
              if (idx == 0) off = r0;
              else {
                off = idx == 1 ? r1 : (idx == 2 ? r2 : r0 - 1);
                if (idx > 1) r2 = r1;
                r1 = r0;
                r0 = off;
              }
            }
            if (lpos + llen > regen || o + llen + mlen > ulen || off == 0 || off > o + llen) ZFAIL(219);
            if (lits_in_input) copy_in(sm, dst, o, lits + lpos, llen, lane); else copy_buf(sm, dst, o, lits + lpos, llen, lane);
            lpos += llen; o += llen;
            snp_syncwarp();
            copy_match(sm, dst, o, off, mlen, lane);
            o += mlen;
            snp_syncwarp();
          }
          if (b.off != 0) ZFAIL(220);
          snp_syncwarp();
          if (lane == 0) { sm.rep[0] = r0; sm.rep[1] = r1; sm.rep[2] = r2; }
        }
        // ---- the literals behind the last sequence
        {
          const uint32_t rest = regen - lpos;
          if (o + rest > ulen) ZFAIL(221);
          if (lits_in_input) copy_in(sm, dst, o, lits + lpos, rest, lane); else copy_buf(sm, dst, o, lits + lpos, rest, lane);
          o += rest;
        }
        p = bend;
        snp_syncwarp();
      } else ZFAIL(222);
      if (last) break;
    }
    if (checksum) p += 4;                                      // xxh64 of the content: not verified (the Parquet page has its own size check)
  }
  snp_syncwarp();
  if (o != ulen) ZFAIL(223);
#undef ZFAIL
}

}  // namespace zst
}  // namespace horae
