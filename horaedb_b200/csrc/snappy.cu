// snappy.cu — raw-Snappy page decompression kernels.  The decoder itself (one warp per page: binary-lifting parse of a staged
// window, word mode / run mode execution, ring -> global flushes) lives in snappy_core.h, which the CPU test-suite compiles too.
#include "kernels.h"

#include <cstdlib>
#include <cstring>

#define SNP_FN __device__ __forceinline__
#define snp_shfl(v, src) __shfl_sync(0xffffffffu, (v), (src))
#define snp_shfl_up(v, d) __shfl_up_sync(0xffffffffu, (v), (d))
#define snp_ballot(p) __ballot_sync(0xffffffffu, (p))
#define snp_any(p) __any_sync(0xffffffffu, (p))
#define snp_syncwarp() __syncwarp()
#define snp_ldg8(p) __ldg(p)
#define snp_ldg64(p) __ldg(p)
#define snp_funnel_r(lo, hi, sh) __funnelshift_r((lo), (hi), (sh))
#define snp_byte_perm(a, b, s) __byte_perm((a), (b), (s))
#define snp_ffs(x) __ffs(x)
#define snp_set_err(err, code) atomicExch((err), (code))
namespace horae {
namespace snp {
__device__ __forceinline__ uint32_t snp_ldcg32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint8_t snp_ldcg8(const uint8_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return uint8_t(v);
}
}  // namespace snp
}  // namespace horae
#include "snappy_core.h"

namespace horae {
namespace k {

namespace {

using namespace horae::snp;
constexpr int kWarpsPerCta = 4;

__device__ __forceinline__ uint64_t chunk_scratch_off2(const RgSel& rs, const ChunkDev* chunks, const ColSel* cols, int ci) {
  uint64_t off = rs.scratch_off;
  for (int j = 0; j < ci; j++) {
    off += chunks[cols[j].col].scratch_bytes;      // 0 for uncompressed PLAIN chunks
  }
  return off;
}

__device__ __forceinline__ void init_tag_tables(uint8_t* s_csz, uint32_t* s_lut) {
  for (uint32_t t = threadIdx.x; t < 256; t += kWarpsPerCta * 32) { s_csz[t] = uint8_t(elem_csize(t)); s_lut[t] = elem_lut(t); }
  __syncthreads();
}

// One warp per column chunk, chunks handed out by an atomic ticket in the order (column order[0] of every row group,
// then order[1], ...): the host lists the columns with the most compressed bytes first, and inside a column J.lpt lists
// the row groups with the most bytes to decode first, so the long pages start early and the short ones fill the tail.
__global__ void __launch_bounds__(kWarpsPerCta * 32, 8) snappy_pages_kernel(const __grid_constant__ SnappyJob J) {
  __shared__ WarpSmem s_w[kWarpsPerCta];
  __shared__ uint32_t s_lut[256];
  __shared__ uint8_t s_csz[256];
  init_tag_tables(s_csz, s_lut);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  WarpSmem& sm = s_w[wid];
  bulk_init(sm, lane);
  uint32_t phase = 0;
  const uint32_t nsel = J.d_nsel ? *J.d_nsel : J.nsel;
  const uint32_t nchunks = nsel * uint32_t(J.ncols);
  for (;;) {
    uint32_t c = 0;
    if (lane == 0) c = atomicAdd(J.ticket, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= nchunks) return;
    const uint32_t si = J.lpt ? J.lpt[c % nsel] : c % nsel;
    const int ci = J.order[c / nsel];
    RgSel rs = J.sel[si];
    SstDev sst = J.ssts[rs.sst];
    const ChunkDev* chunks = sst.chunks + size_t(rs.rg) * sst.ncols;
    ChunkDev ch = chunks[J.col_from_cols ? J.cols[ci].col : J.col[ci]];
    if (ch.codec != 1) continue;
    if (ch.stored && J.skip_stored[ci]) continue;                 // read in place by the consumer
    uint8_t* dst = J.scratch + (J.fixed_stride ? rs.scratch_off + uint64_t(J.region[ci]) * J.fixed_stride
                                               : chunk_scratch_off2(rs, chunks, J.cols, ci));
    // the chunk's streams in scratch order: a compressed dictionary page first (p == -1), then the data pages.  ONE call site
    // of the decoder keeps the kernel's code (and its instruction-cache footprint) at one copy.
    for (int p = ch.dict_uncomp ? -1 : 0; p < int(ch.num_pages); p++) {
      const uint8_t* src;
      uint32_t n, ulen, stop_at = 0xffffffffu;
      uint64_t advance;
      bool compressed = true;
      if (p < 0) {
        src = sst.bytes + ch.dict_payload_off; n = ch.dict_comp; ulen = ch.dict_uncomp;
        advance = page_scratch2(ch.dict_uncomp);
      } else {
        const PageDev pg = sst.pages[ch.first_page + p];
        src = sst.bytes + pg.payload_off; n = pg.comp_size; ulen = pg.uncomp_size;
        if (pg.page_type == 3) {
          const uint32_t skip = pg.v2_def_len + pg.v2_rep_len;
          src += skip; n -= skip; ulen -= skip;
          compressed = pg.v2_compressed != 0;
        }
        if (J.partial[ci]) {
          // the consumer reads rows [0, rs.out_row) only (gate-first: nothing behind the last row that passes the gate column can
          // survive the filter): level prefix (<= 16 + rows / 8 bytes) + that many values
          const uint32_t w = (ch.phys == 1 || ch.phys == 4) ? 4u : 8u;
          stop_at = 16u + (rs.num_rows + 7u) / 8u + 8u + rs.out_row * w;
        }
        advance = page_scratch2(pg.uncomp_size);
        if (pg.encoding == 5 || pg.encoding == 6 || pg.encoding == 8 || pg.encoding == 2) advance += page_scratch2(pg.num_values * 8u);   // PLAIN image of a DELTA / dictionary page (decode_chunks)
      }
      if (compressed) snappy_page(src, n, dst, ulen, stop_at, sm, phase, s_csz, s_lut, lane, J.err);
      dst += advance;
    }
  }
}

// pages given by pointer (transient loads decompress the gate column before the SST's tables exist on the device)
__global__ void __launch_bounds__(kWarpsPerCta * 32, 8) snappy_raw_kernel(const RawPage* __restrict__ pages, uint32_t n, unsigned int* ticket, int* err) {
  __shared__ WarpSmem s_w[kWarpsPerCta];
  __shared__ uint32_t s_lut[256];
  __shared__ uint8_t s_csz[256];
  init_tag_tables(s_csz, s_lut);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  bulk_init(s_w[wid], lane);
  uint32_t phase = 0;
  for (;;) {
    uint32_t c = 0;
    if (lane == 0) c = atomicAdd(ticket, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= n) return;
    const RawPage pg = pages[c];
    snappy_page(pg.src, pg.comp_size, pg.dst, pg.uncomp_size, 0xffffffffu, s_w[wid], phase, s_csz, s_lut, lane, err);
  }
}

}  // namespace

// Resident CTAs per SM the decompression kernels ask for.  8 would fill an SM's shared memory (8 x (26.9 + 1) KB of 228 KB); 7 measured
// 2.6 % faster on the bench (config 2: 8 -> 1.997, 7 -> 1.944, 6 -> 1.983, 5 -> 2.141 ms per step: the stage is bound by instruction
// issue and shared-memory traffic, not by the number of warps; compiling for 7 with 72 registers was no faster than 64), and 28 KB per
// SM stay free for the library's own NCCL all-gather of the previous step's partials, which runs NEXT to the decompression of the
// current step (hg_agg_combine).  Also measured without effect: a persisting-L2 carve-out (32 / 64 MB) with evict_last stores of the
// word-mode output, meant to keep a page's far-copy sources in L2.
static int g_ctas_per_sm = 0;
void snappy_set_ctas_per_sm(int n) { g_ctas_per_sm = n; }
static uint32_t snappy_max_ctas() {
  static const int env = getenv("HORAE_SNAPPY_CTAS_PER_SM") ? atoi(getenv("HORAE_SNAPPY_CTAS_PER_SM")) : 0;
  int n = env > 0 ? env : (g_ctas_per_sm > 0 ? g_ctas_per_sm : 7);
  if (n > 8) n = 8;
  return 148u * uint32_t(n);
}

void snappy_raw_pages(const Launch& L, const RawPage* d_pages, uint32_t n, unsigned int* ticket, int* err) {
  if (!n) return;
  uint32_t ctas = (n + kWarpsPerCta - 1) / kWarpsPerCta;
  if (ctas > snappy_max_ctas()) ctas = snappy_max_ctas();
  snappy_raw_kernel<<<ctas, kWarpsPerCta * 32, 0, L.stream>>>(d_pages, n, ticket, err);
  L.tick();
}

void snappy_pages(const Launch& L, const SnappyJob& job, uint32_t max_chunks) {
  if (!max_chunks) return;
  uint32_t ctas = (max_chunks + kWarpsPerCta - 1) / kWarpsPerCta;
  if (ctas > snappy_max_ctas()) ctas = snappy_max_ctas();
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
