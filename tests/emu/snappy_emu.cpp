// snappy_emu.cpp — TEST INFRASTRUCTURE.  Runs the product's warp-level Snappy decoder (horaedb_b2../../horaedb_b200/csrc/snappy_core.h, the very text
// nvcc compiles for sm_100a) on the CPU: the 32 lanes of the warp are 32 coroutines (ucontext), every warp collective
// (shuffle / ballot / any / syncwarp) is a rendezvous.  Lanes run one after another between collectives, so the emulation checks the
// lane-level LOGIC (tables, source classification, parent links, ring arithmetic, flush positions), not instruction timing.
// Built by tests/test_snappy_emu.py with g++; nothing in the product links it.
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace emu {
constexpr int kLanes = 32;
struct Warp {
  ucontext_t sched, lane_ctx[kLanes];
  std::vector<char> stacks[kLanes];
  bool done[kLanes];
  uint32_t slot[2][kLanes];
  int parity[kLanes];     // per lane: which buffer its next collective uses (all lanes stay in step)
  int cur = 0;
  long collectives = 0;
};
static Warp* W;
static inline int lane_id() { return W->cur; }
// publish v, wait for everyone, return the buffer all lanes published into
static inline const uint32_t* rendezvous(uint32_t v) {
  const int l = W->cur, p = W->parity[l];
  W->slot[p][l] = v;
  W->parity[l] = p ^ 1;
  swapcontext(&W->lane_ctx[l], &W->sched);
  return W->slot[p];
}
static inline uint32_t shfl(uint32_t v, int src) { return rendezvous(v)[src & 31]; }
static inline uint32_t shfl_up(uint32_t v, int d) { const int l = W->cur; const uint32_t* s = rendezvous(v); return l >= d ? s[l - d] : v; }
static inline uint32_t ballot(bool p) { const uint32_t* s = rendezvous(p ? 1u : 0u); uint32_t m = 0; for (int i = 0; i < kLanes; i++) m |= (s[i] & 1u) << i; return m; }
static inline uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) {
  const uint64_t pool = (uint64_t(b) << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= uint32_t((pool >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
}  // namespace emu

struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

// dynamic counters of the decoder's loop (lane 0 counts): windows, steps, elements, word_steps, bytes, parent_searches
struct EmuStats { long windows, steps, elements, word_steps, bytes, parent_searches, stage_hits; };
static EmuStats g_stats;
#define SNP_STAT(counter, amount) do { if (emu::lane_id() == 0) g_stats.counter += long(amount); } while (0)
#define SNP_FN static inline
#define snp_shfl(v, src) emu::shfl(uint32_t(v), int(src))
#define snp_shfl_up(v, d) emu::shfl_up(uint32_t(v), int(d))
#define snp_ballot(p) emu::ballot(bool(p))
#define snp_any(p) (emu::ballot(bool(p)) != 0)
#define snp_syncwarp() ((void)emu::rendezvous(0))
#define snp_ldg8(p) (*(p))
#define snp_ldg64(p) (*(p))
#define snp_ldcg8(p) (*(p))
#define snp_ldcg32(p) (*(p))
#define snp_funnel_r(lo, hi, sh) emu::funnel_r((lo), (hi), (sh))
#define snp_byte_perm(a, b, s) emu::byte_perm((a), (b), (s))
#define snp_ffs(x) __builtin_ffs(int(x))
#define snp_set_err(err, code) (*(err) = (code))
#include "../../horaedb_b200/csrc/snappy_core.h"

namespace {
struct Job {
  const uint8_t* src; uint32_t n; uint8_t* dst; uint32_t ulen, stop_at;
  horae::snp::WarpSmem* sm; const uint8_t* csz; const uint32_t* lut; int* err;
};
Job g_job;
void lane_main() {
  const int lane = emu::W->cur;
  uint32_t phase = 0;
  horae::snp::bulk_init(*g_job.sm, lane);
  horae::snp::snappy_page(g_job.src, g_job.n, g_job.dst, g_job.ulen, g_job.stop_at, *g_job.sm, phase, g_job.csz, g_job.lut, lane, g_job.err);
  emu::W->done[lane] = true;
  swapcontext(&emu::W->lane_ctx[lane], &emu::W->sched);
}
}  // namespace

// Decode one raw Snappy stream.  dst must have ulen + 64 bytes of room (the decoder may overshoot stop_at by one batch and reads
// whole words).  Returns the decoder's error word (0 = ok); *collectives = warp collectives executed (a proxy for steps).
extern "C" void emu_stats(long* out7) { std::memcpy(out7, &g_stats, sizeof(g_stats)); std::memset(&g_stats, 0, sizeof(g_stats)); }
extern "C" int emu_snappy_page(const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t ulen, uint32_t stop_at, long* collectives) {
  using namespace horae::snp;
  static uint8_t csz[256];
  static uint32_t lut[256];
  for (uint32_t t = 0; t < 256; t++) { csz[t] = uint8_t(elem_csize(t)); lut[t] = elem_lut(t); }
  // input copy with slack on both sides: the decoder loads aligned words around unaligned addresses
  std::vector<uint8_t> in(size_t(n) + 128, 0);
  std::memcpy(in.data() + 32, src, n);
  WarpSmem* sm = static_cast<WarpSmem*>(aligned_alloc(256, sizeof(WarpSmem)));
  std::memset(sm, 0xa5, sizeof(WarpSmem));
  int err = 0;
  g_job = Job{in.data() + 32, n, dst, ulen, stop_at, sm, csz, lut, &err};
  emu::Warp warp;
  emu::W = &warp;
  for (int l = 0; l < emu::kLanes; l++) {
    warp.stacks[l].resize(256 * 1024);
    warp.done[l] = false;
    warp.parity[l] = 0;
    getcontext(&warp.lane_ctx[l]);
    warp.lane_ctx[l].uc_stack.ss_sp = warp.stacks[l].data();
    warp.lane_ctx[l].uc_stack.ss_size = warp.stacks[l].size();
    warp.lane_ctx[l].uc_link = &warp.sched;
    makecontext(&warp.lane_ctx[l], lane_main, 0);
  }
  for (;;) {
    int live = 0;
    for (int l = 0; l < emu::kLanes; l++) {
      if (warp.done[l]) continue;
      warp.cur = l;
      swapcontext(&warp.sched, &warp.lane_ctx[l]);
      live++;
    }
    if (!live) break;
    warp.collectives++;
    // lanes must stay in step: a collective reached by some lanes only is a bug in the decoder (undefined on the GPU too)
    int p = -1;
    bool any_done = false, any_live = false;
    for (int l = 0; l < emu::kLanes; l++) {
      if (warp.done[l]) { any_done = true; continue; }
      any_live = true;
      if (p < 0) p = warp.parity[l];
      else if (p != warp.parity[l]) { err = 9001; goto out; }
    }
    if (any_done && any_live) { err = 9002; goto out; }      // some lanes returned while others wait in a collective
  }
out:
  if (collectives) *collectives = warp.collectives;
  free(sm);
  return err;
}
