// cuda_emu.h — TEST INFRASTRUCTURE.  Just enough of CUDA on the CPU to run the product's kernels UNCHANGED under `pytest -m "not gpu"`:
// tests/emu/build_engine_emu.py rewrites the `<<<...>>>` launches of horaedb_b200/csrc/*.cu into EMU_LAUNCH(...) (and nothing else of
// substance), and g++ compiles the result against this header into tests/emu/_build/libhorae_emu.so.  Nothing in the product loads that
// library; the product still needs a GPU and fails loudly without one.
//
//   threads     every thread of a block is a coroutine (own stack, hand-written register switch); blocks run one after another, in order
//   barriers    __syncthreads() and the warp collectives are rendezvous points: a thread runs until it reaches one, the scheduler
//               releases a warp / the block when every expected thread has arrived.  Every collective carries its source line, so lanes
//               that meet in DIFFERENT collectives (undefined on the GPU) are reported, as are barriers that can never complete.
//   order       between two rendezvous the hardware may run threads in any order; emu_set_order(0 ascending | 1 descending | >= 2 a fresh
//               random permutation per scheduling pass) makes a missing barrier show up as a wrong result
//   memory      cudaMalloc = malloc filled with 0xCD, "device" pointers are host pointers, copies are memcpy, streams and events are
//               no-ops (launches run synchronously)
//   knobs       HORAE_EMU_ORDER (thread order), HORAE_EMU_GUARD (an inaccessible page right behind every allocation), HORAE_EMU_PINNED (host
//               buffers count as pinned: the zero-copy gather path of transient loads), HORAE_EMU_CRASH_REPORT (kernel / block / thread of a
//               fault), HORAE_EMU_TRACE_ALLOC; HORAE_EMU_DROP_BARRIER at build time (mutants, build_engine_emu.py)
//   ranks       engines of several host threads take turns kernel by kernel (launch_mutex); tests/emu/nccl_emu.cpp is the all-gather between them
//   not modelled: timing, caches, memory-model effects beyond "a write is visible after the next rendezvous or kernel end"
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <type_traits>
#include <vector>

// ------------------------------------------------------------------------------------------------ language
#define __global__
#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// CUDA's min / max accept mixed integer types
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
static inline std::common_type_t<A, B> min(A a, B b) { using T = std::common_type_t<A, B>; return T(a) < T(b) ? T(a) : T(b); }
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
static inline std::common_type_t<A, B> max(A a, B b) { using T = std::common_type_t<A, B>; return T(a) > T(b) ? T(a) : T(b); }

// ------------------------------------------------------------------------------------------------ scheduler
namespace emu {
constexpr int kMaxThreads = 1024;
constexpr size_t kStackBytes = 256 * 1024;
enum State : uint8_t { RUNNABLE, WAIT_WARP, WAIT_BLOCK, DONE };
// a coroutine = a stack pointer; emu_switch (engine_emu_glue.cpp, 14 instructions) saves the callee-saved registers on the old stack and
// pops them from the new one.  (ucontext's swapcontext makes a sigprocmask system call per switch: ~10x slower.)
struct Ctx { void* sp; };
extern "C" void emu_switch(Ctx* from, Ctx* to);
struct Block {
  Ctx sched;
  Ctx ctx[kMaxThreads];
  State state[kMaxThreads];
  uint64_t slot[kMaxThreads];           // value a lane published for the collective it waits in
  uint64_t delivered[kMaxThreads];      // copy made when the group is released: every released lane runs (and reads) in the next pass,
                                        // before anything is released again, while new values go to `slot`
  uint32_t want[kMaxThreads];           // the mask a waiting lane named
  int site[kMaxThreads];                // source line of the rendezvous a thread waits in
  int nthreads = 0;
  int cur = 0;
};
struct Globals {
  Block* blk = nullptr;
  char* stacks = nullptr;
  int order = 0;
  uint64_t rng = 1;
  int error = 0;                        // first scheduling error of the process (sticky until read)
  int error_line = 0;
  long launches = 0, rendezvous = 0;
  std::function<void()> body;
  alignas(16) unsigned char* dyn_smem = nullptr;
  size_t dyn_cap = 0;
  const char* kernel_name = "";        // the launch in progress (for the crash report of engine_emu_glue.cpp)
};
inline Globals& G() { static Globals g; return g; }

struct Idx { unsigned x, y, z; };
inline Idx& tidx() { static Idx v{0, 0, 0}; return v; }
inline Idx& bidx() { static Idx v{0, 0, 0}; return v; }
inline Idx& bdim() { static Idx v{1, 1, 1}; return v; }
inline Idx& gdim() { static Idx v{1, 1, 1}; return v; }

inline void set_error(int code, int line) { Globals& g = G(); if (!g.error) { g.error = code; g.error_line = line; } }

// leave the running thread for the scheduler
inline void yield_to_sched() { Block* b = G().blk; emu_switch(&b->ctx[b->cur], &b->sched); }

inline void syncthreads(int line) {
  Block* b = G().blk;
  if (!b) return;
  const int t = b->cur;
  b->state[t] = WAIT_BLOCK;
  b->site[t] = line;
  yield_to_sched();
}

// publish v, wait for the lanes of `mask`, return the buffer they published into (indexed by thread id: base of the warp + lane)
inline const uint64_t* warp_rendezvous(uint32_t mask, uint64_t v, int line) {
  Block* b = G().blk;
  const int t = b->cur, w = t >> 5;
  b->slot[t] = v;
  b->want[t] = mask;
  b->site[t] = line;
  b->state[t] = WAIT_WARP;
  yield_to_sched();
  return b->delivered + (w << 5);
}

inline void thread_main() {
  G().body();
  Block* b = G().blk;
  b->state[b->cur] = DONE;
  yield_to_sched();
  abort();                              // a finished thread is never resumed
}

inline void run_block(int nthreads) {
  Globals& g = G();
  if (!g.blk) g.blk = new Block();
  if (!g.stacks) g.stacks = static_cast<char*>(malloc(kStackBytes * kMaxThreads));
  Block* b = g.blk;
  b->nthreads = nthreads;
  for (int t = 0; t < nthreads; t++) {
    b->state[t] = RUNNABLE;
    // fresh stack: six zeroed register slots, the entry point as the return address emu_switch's `ret` takes, and one pad word so that
    // the entry point sees the stack alignment of a called function
    void** top = reinterpret_cast<void**>(g.stacks + size_t(t + 1) * kStackBytes);
    top -= 2;
    top[0] = reinterpret_cast<void*>(&thread_main);
    top[1] = nullptr;
    top -= 6;
    for (int r = 0; r < 6; r++) top[r] = nullptr;
    b->ctx[t].sp = top;
  }
  std::vector<int> perm(static_cast<size_t>(nthreads));
  for (;;) {
    for (int i = 0; i < nthreads; i++) perm[size_t(i)] = g.order == 1 ? nthreads - 1 - i : i;
    if (g.order >= 2)
      for (int i = nthreads - 1; i > 0; i--) {
        g.rng = g.rng * 6364136223846793005ull + 1442695040888963407ull;
        std::swap(perm[size_t(i)], perm[size_t((g.rng >> 33) % uint64_t(i + 1))]);
      }
    bool ran = false;
    for (int i = 0; i < nthreads; i++) {
      const int t = perm[size_t(i)];
      if (b->state[t] != RUNNABLE) continue;
      b->cur = t;
      tidx().x = unsigned(t);
      emu_switch(&b->sched, &b->ctx[t]);
      ran = true;
    }
    // release what is complete
    bool released = false;
    int live = 0, at_block = 0, block_site = -1;
    bool block_site_mixed = false;
    for (int t = 0; t < nthreads; t++) {
      if (b->state[t] == DONE) continue;
      live++;
      if (b->state[t] == WAIT_BLOCK) {
        at_block++;
        if (block_site < 0) block_site = b->site[t]; else if (block_site != b->site[t]) block_site_mixed = true;
      }
    }
    if (!live) break;
    for (int w = 0; w < (nthreads + 31) / 32; w++) {
      const int base = w << 5;
      uint32_t waiting = 0;
      for (int l = 0; l < 32 && base + l < nthreads; l++) if (b->state[base + l] == WAIT_WARP) waiting |= 1u << l;
      if (!waiting) continue;
      // groups of lanes that named the same mask (e.g. the two halves of a diverged warp): a group goes when all its lanes are here
      uint32_t todo = waiting;
      while (todo) {
        const int l0 = __builtin_ctz(todo);
        uint32_t lanes_in_block = (nthreads - base >= 32) ? 0xffffffffu : ((1u << (nthreads - base)) - 1u);
        const uint32_t m = b->want[base + l0] & lanes_in_block;
        todo &= ~m;
        if ((waiting & m) != m) {
          // someone in the mask is not here: fine while it can still arrive; an exited lane never will
          for (int l = 0; l < 32; l++) if (((m >> l) & 1u) && b->state[base + l] == DONE) set_error(9002, b->site[base + l0]);
          continue;
        }
        for (int l = 0; l < 32; l++)
          if ((m >> l) & 1u) {
            if (b->site[base + l] != b->site[base + l0]) set_error(9001, b->site[base + l0]);      // lanes met in different collectives
            b->state[base + l] = RUNNABLE;
            b->delivered[base + l] = b->slot[base + l];
          }
        released = true;
        g.rendezvous++;
      }
    }
    if (at_block && at_block == live) {
      if (block_site_mixed) set_error(9003, block_site);          // threads wait in different __syncthreads(): legal only if counts match; flagged
      for (int t = 0; t < nthreads; t++) if (b->state[t] == WAIT_BLOCK) b->state[t] = RUNNABLE;
      released = true;
      g.rendezvous++;
    }
    if (g.error == 9001 || g.error == 9002) break;
    if (!ran && !released) { set_error(9004, block_site); break; }   // nobody can move: a barrier that cannot complete
  }
}

// kernels of different host threads (ranks of an emulated multi-GPU run) take turns: the scheduler's state is global
inline std::mutex& launch_mutex() { static std::mutex m; return m; }
template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F&& body) {
  std::lock_guard<std::mutex> turn(launch_mutex());
  Globals& g = G();
  g.launches++;
  if (smem > g.dyn_cap) { free(g.dyn_smem); g.dyn_smem = static_cast<unsigned char*>(aligned_alloc(256, (smem + 255) / 256 * 256)); g.dyn_cap = smem; }
  g.body = std::function<void()>(body);
  bdim() = Idx{block.x, 1, 1};
  gdim() = Idx{grid.x, 1, 1};
  for (unsigned bx = 0; bx < grid.x; bx++) {
    bidx() = Idx{bx, 0, 0};
    run_block(int(block.x));
    if (g.error == 9001 || g.error == 9002 || g.error == 9004) break;
  }
  g.blk->nthreads = 0;
}
// the arguments are evaluated once, at the launch (as on the GPU), and copied into every thread's call
template <class K, class... A>
inline void launch_kernel(dim3 grid, dim3 block, size_t smem, K kernel, A... args) { launch(grid, block, smem, [=]() { kernel(args...); }); }
}  // namespace emu

#define threadIdx (emu::tidx())
#define blockIdx (emu::bidx())
#define blockDim (emu::bdim())
#define gridDim (emu::gdim())
#define warpSize 32
#define EMU_LAUNCH(kernel, grid, block, smem, stream, ...) (emu::G().kernel_name = #kernel, emu::launch_kernel(dim3(grid), dim3(block), size_t(smem), kernel, ##__VA_ARGS__))
#define EMU_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::G().dyn_smem)

// ------------------------------------------------------------------------------------------------ intrinsics
#define __syncthreads() emu::syncthreads(__LINE__)
#define __syncwarp(...) ((void)emu::warp_rendezvous(emu::mask_of(__VA_ARGS__), 0, __LINE__))
#define __shfl_sync(m, v, src) emu::shfl((m), (v), int(src), __LINE__)
#define __shfl_up_sync(m, v, d) emu::shfl_up((m), (v), int(d), __LINE__)
#define __shfl_down_sync(m, v, d) emu::shfl_down((m), (v), int(d), __LINE__)
#define __shfl_xor_sync(m, v, x) emu::shfl_xor((m), (v), int(x), __LINE__)
#define __ballot_sync(m, p) emu::ballot((m), bool(p), __LINE__)
#define __any_sync(m, p) (emu::ballot((m), bool(p), __LINE__) != 0)
#define __all_sync(m, p) (emu::ballot((m), !bool(p), __LINE__) == 0)
#define __match_any_sync(m, v) emu::match_any((m), uint64_t(v), __LINE__)

namespace emu {
inline uint32_t mask_of() { return 0xffffffffu; }
inline uint32_t mask_of(uint32_t m) { return m; }
inline int lane() { return G().blk->cur & 31; }
template <class T> inline uint64_t bits_of(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "shuffle of a wide type"); std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
template <class T> inline T shfl(uint32_t m, T v, int src, int line) { const uint64_t* s = warp_rendezvous(m, bits_of(v), line); return from_bits<T>(s[src & 31]); }
template <class T> inline T shfl_up(uint32_t m, T v, int d, int line) { const int l = lane(); const uint64_t* s = warp_rendezvous(m, bits_of(v), line); return l >= d ? from_bits<T>(s[l - d]) : v; }
template <class T> inline T shfl_down(uint32_t m, T v, int d, int line) { const int l = lane(); const uint64_t* s = warp_rendezvous(m, bits_of(v), line); return l + d < 32 ? from_bits<T>(s[l + d]) : v; }
template <class T> inline T shfl_xor(uint32_t m, T v, int x, int line) { const int l = lane(); const uint64_t* s = warp_rendezvous(m, bits_of(v), line); return from_bits<T>(s[(l ^ x) & 31]); }
inline uint32_t ballot(uint32_t m, bool p, int line) {
  const uint64_t* s = warp_rendezvous(m, p ? 1u : 0u, line);
  uint32_t r = 0;
  for (int i = 0; i < 32; i++) if ((m >> i) & 1u) r |= uint32_t(s[i] & 1u) << i;
  return r;
}
inline uint32_t match_any(uint32_t m, uint64_t v, int line) {
  const uint64_t* s = warp_rendezvous(m, v, line);
  uint32_t r = 0;
  for (int i = 0; i < 32; i++) if (((m >> i) & 1u) && s[i] == v) r |= 1u << i;
  return r;
}
}  // namespace emu

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __popcll(uint64_t v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz(unsigned(v)) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll(static_cast<unsigned long long>(v)) : 64; }
static inline uint32_t __brev(uint32_t v) { uint32_t r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) { sh &= 31; return sh ? (hi << sh) | (lo >> (32 - sh)) : hi; }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
  const uint64_t pool = (uint64_t(b) << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= uint32_t((pool >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline float __uint_as_float(uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t v; std::memcpy(&v, &f, 4); return v; }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
static inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) {}

// atomics (threads never run concurrently)
template <class T, class U> static inline T atomicAdd(T* p, U v) { const T old = *p; *p = T(old + T(v)); return old; }
template <class T, class U> static inline T atomicSub(T* p, U v) { const T old = *p; *p = T(old - T(v)); return old; }
template <class T, class U> static inline T atomicExch(T* p, U v) { const T old = *p; *p = T(v); return old; }
template <class T, class U> static inline T atomicMax(T* p, U v) { const T old = *p; if (T(v) > old) *p = T(v); return old; }
template <class T, class U> static inline T atomicMin(T* p, U v) { const T old = *p; if (T(v) < old) *p = T(v); return old; }
template <class T, class U> static inline T atomicOr(T* p, U v) { const T old = *p; *p = T(old | T(v)); return old; }
template <class T, class U> static inline T atomicAnd(T* p, U v) { const T old = *p; *p = T(old & T(v)); return old; }
template <class T, class U, class V> static inline T atomicCAS(T* p, U cmp, V v) { const T old = *p; if (old == T(cmp)) *p = T(v); return old; }

// ------------------------------------------------------------------------------------------------ runtime API
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef struct EmuStream* cudaStream_t;
typedef struct EmuEvent* cudaEvent_t;
typedef struct EmuPool* cudaMemPool_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; size_t totalGlobalMem; int major, minor; size_t sharedMemPerBlockOptin; };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory (emulated)" : "emulated CUDA error"); }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { std::memset(p, 0, sizeof(*p)); std::strcpy(p->name, "emulated"); p->multiProcessorCount = 148; p->totalGlobalMem = size_t(8) << 30; p->major = 10; p->sharedMemPerBlockOptin = 227 * 1024; return cudaSuccess; }
// HORAE_EMU_GUARD=1: every allocation ends (up to 16-byte alignment) right before an inaccessible page, so a read or write past the end
// of a cudaMalloc'ed buffer is a segfault (reported with kernel / block / thread by engine_emu_glue.cpp) instead of a silent neighbour access
namespace emu {
void* guarded_alloc(size_t n);
bool guarded_free(void* p);
}  // namespace emu
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) {
  void* q = nullptr;
  static const bool guard = getenv("HORAE_EMU_GUARD") != nullptr;
  if (guard) q = emu::guarded_alloc(n ? n : 1);
  else if (posix_memalign(&q, 512, n ? (n + 511) / 512 * 512 : 512) != 0) q = nullptr;
  if (!q) return cudaErrorMemoryAllocation;
  std::memset(q, 0xCD, n);
  *p = static_cast<T*>(q);
  if (getenv("HORAE_EMU_TRACE_ALLOC")) fprintf(stderr, "[emu] alloc %p .. %p (%zu bytes)\n", q, static_cast<void*>(static_cast<char*>(q) + n), n);
  return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) { if (p && !emu::guarded_free(p)) free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n, unsigned = 0) { return cudaMalloc(p, n); }
template <class T> static inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(uintptr_t(0x10)); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { return cudaStreamCreateWithFlags(s, 0); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = reinterpret_cast<cudaEvent_t>(uintptr_t(0x20)); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = nullptr; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }
// HORAE_EMU_PINNED=1: every host buffer counts as pinned, so transient loads take the zero-copy gather kernel instead of one memcpy per range
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { static const bool pinned = getenv("HORAE_EMU_PINNED") != nullptr; a->type = pinned ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered; a->device = 0; a->devicePointer = nullptr; a->hostPointer = nullptr; return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }

extern "C" {
void emu_set_order(int order);
int emu_take_error(int* line);
void emu_counters(long* launches, long* rendezvous);
}
