// zstd_emu.cpp — TEST INFRASTRUCTURE.  Runs the product's warp-level Zstandard decoder (horaedb_b200/csrc/zstd_core.h, the text nvcc
// compiles for sm_100a) on the CPU with the 32 lanes as coroutines (warp_emu.h).  Built by tests/test_zstd_emu.py; nothing in the
// product links it.
#include "warp_emu.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define SNP_FN static inline
#define SNP_CONST static const
#define snp_any(p) (emu::ballot(bool(p)) != 0)
#define snp_syncwarp() ((void)emu::rendezvous(0))
#define snp_ldg8(p) (*(p))
#define snp_ldcg8(p) (*(p))
static inline uint64_t snp_ldg64u(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
#define snp_set_err(err, code) (*(err) = (code))
#include "../../horaedb_b200/csrc/zstd_core.h"

namespace {
struct Job { const uint8_t* src; uint32_t n; uint8_t* dst; uint32_t ulen; uint8_t* lit; horae::zst::WarpSmem* sm; int* err; };
Job g_job;
void lane_main() {
  horae::zst::zstd_page(g_job.src, g_job.n, g_job.dst, g_job.ulen, g_job.lit, *g_job.sm, emu::lane_id(), g_job.err);
  emu::lane_exit();
}
}  // namespace

// Decode the Zstandard frame(s) in src to dst (ulen bytes expected, ulen + 64 bytes of room).  Returns the decoder's error word.
extern "C" void emu_set_order(int order) { emu::g_order = order; }
extern "C" int emu_zstd_page(const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t ulen, long* collectives) {
  std::vector<uint8_t> in(size_t(n) + 128, 0xA5);
  std::memcpy(in.data() + 32, src, n);
  std::vector<uint8_t> lit(size_t(ulen < (128u << 10) ? ulen : (128u << 10)) + 64, 0xEE);
  horae::zst::WarpSmem* sm = static_cast<horae::zst::WarpSmem*>(aligned_alloc(256, (sizeof(horae::zst::WarpSmem) + 255) / 256 * 256));
  std::memset(sm, 0xa5, sizeof(horae::zst::WarpSmem));
  int err = 0;
  g_job = Job{in.data() + 32, n, dst, ulen, lit.data(), sm, &err};
  const int werr = emu::run_warp(lane_main, collectives);
  if (werr) err = werr;
  // the literal buffer holds min(page, 128 KB) bytes (+ the slack every scratch region has): nothing may be written behind it
  for (size_t i = lit.size() - 32; i < lit.size(); i++) if (lit[i] != 0xEE) err = 9003;
  free(sm);
  return err;
}
