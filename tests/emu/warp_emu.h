// warp_emu.h — TEST INFRASTRUCTURE shared by snappy_emu.cpp / zstd_emu.cpp: a 32-lane "warp" on the CPU.  Lanes are coroutines (ucontext),
// every warp collective (shuffle / ballot / any / syncwarp) is a rendezvous; lanes run one after another between collectives, so the
// emulation checks lane-level LOGIC, not instruction timing.  run_warp() returns 9001 / 9002 when lanes fall out of step (a collective
// reached by some lanes only: undefined on the GPU too).  Between two collectives the hardware may run the lanes in ANY order (independent
// thread scheduling), so the order is a knob: g_order 0 ascending, 1 descending, >= 2 a fresh pseudo-random permutation per interval
// (seeded by the value).  A read-after-write or write-after-read between lanes that lacks a __syncwarp() shows up as a wrong result under
// one of them.
#pragma once
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
static int g_order = 0;
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


namespace emu {
// runs lane_main() on 32 lanes until all have returned; *collectives = warp collectives executed
static inline int run_warp(void (*lane_main)(), long* collectives) {
  Warp warp;
  W = &warp;      // (valid until run_warp returns: nothing runs on the lanes afterwards)
  int err = 0;
  uint64_t rng = uint64_t(g_order) * 0x9e3779b97f4a7c15ull + 1;
  for (int l = 0; l < kLanes; l++) {
    warp.stacks[l].resize(512 * 1024);
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
    int perm[kLanes];
    for (int i = 0; i < kLanes; i++) perm[i] = g_order == 1 ? kLanes - 1 - i : i;
    if (g_order >= 2)
      for (int i = kLanes - 1; i > 0; i--) {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        const int j = int((rng >> 33) % uint64_t(i + 1));
        const int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
      }
    for (int i = 0; i < kLanes; i++) {
      const int l = perm[i];
      if (warp.done[l]) continue;
      warp.cur = l;
      swapcontext(&warp.sched, &warp.lane_ctx[l]);
      live++;
    }
    if (!live) break;
    warp.collectives++;
    int p = -1;
    bool any_done = false, any_live = false;
    for (int l = 0; l < kLanes; l++) {
      if (warp.done[l]) { any_done = true; continue; }
      any_live = true;
      if (p < 0) p = warp.parity[l];
      else if (p != warp.parity[l]) { err = 9001; goto out; }
    }
    if (any_done && any_live) { err = 9002; goto out; }      // some lanes returned while others wait in a collective
  }
out:
  if (collectives) *collectives = warp.collectives;
  return err;
}
// a lane's last act: mark itself done and hand control back for good
static inline void lane_exit() {
  const int lane = W->cur;
  W->done[lane] = true;
  swapcontext(&W->lane_ctx[lane], &W->sched);
}
}  // namespace emu
