// engine_emu_glue.cpp — TEST INFRASTRUCTURE: the knobs of the emulated build (see cuda_emu.h)
#include "cuda_emu.h"

#include <execinfo.h>
#include <signal.h>
#include <sys/mman.h>
#include <unistd.h>

#include <map>
#include <mutex>

extern "C" void emu_set_order(int order) { emu::G().order = order; emu::G().rng = uint64_t(order) * 0x9e3779b97f4a7c15ull + 1; }
// first scheduling error since the last call (0 = none): 9001 lanes met in different collectives, 9002 a collective names an exited lane,
// 9003 threads wait in different __syncthreads(), 9004 a barrier that cannot complete; *line = source line of the rendezvous
extern "C" int emu_take_error(int* line) { emu::Globals& g = emu::G(); const int e = g.error; if (line) *line = g.error_line; g.error = 0; g.error_line = 0; return e; }
extern "C" void emu_counters(long* launches, long* rendezvous) { *launches = emu::G().launches; *rendezvous = emu::G().rendezvous; }

// emu_switch(from, to): x86-64 System V.  Callee-saved registers go onto the current stack, the stack pointer into *from; then the same
// in reverse from *to.  A fresh coroutine's stack is laid out by run_block() so that the final `ret` enters thread_main().
#if !defined(__x86_64__)
#error "the emulated build's context switch is written for x86-64"
#endif
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq (%rsi), %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

// a crash inside an emulated kernel: say which launch, block and thread before dying (an out-of-bounds access of a kernel is a segfault here)
namespace {
void on_segv(int sig, siginfo_t* info, void*) {
  char buf[512];
  const emu::Globals& g = emu::G();
  const int n = snprintf(buf, sizeof buf, "\n[emu] signal %d at address %p in kernel %s (launch %ld), block %u of %u, thread %u of %u\n", sig, info ? info->si_addr : nullptr, g.kernel_name, g.launches,
                         emu::bidx().x, emu::gdim().x, emu::tidx().x, emu::bdim().x);
  if (n > 0) (void)!write(2, buf, size_t(n));
  void* frames[48];
  backtrace_symbols_fd(frames, backtrace(frames, 48), 2);
  signal(sig, SIG_DFL);
  raise(sig);
}
struct Install {
  Install() {
    static char alt[1 << 16];
    stack_t ss; ss.ss_sp = alt; ss.ss_size = sizeof alt; ss.ss_flags = 0;
    sigaltstack(&ss, nullptr);
    struct sigaction sa; memset(&sa, 0, sizeof sa); sa.sa_sigaction = on_segv; sa.sa_flags = SA_ONSTACK | SA_SIGINFO;
    if (getenv("HORAE_EMU_CRASH_REPORT")) { sigaction(SIGSEGV, &sa, nullptr); sigaction(SIGBUS, &sa, nullptr); }
    if (const char* o = getenv("HORAE_EMU_ORDER")) emu_set_order(atoi(o));      // (the pytest plugin sets it too; the tools rely on this)
  }
} install;
}  // namespace

namespace emu {
namespace {
std::mutex g_guard_mu;
std::map<void*, std::pair<void*, size_t>> g_guarded;          // user pointer -> (mapping, length)
}  // namespace
void* guarded_alloc(size_t n) {
  const size_t page = 4096, body = (n + 15) / 16 * 16, len = (body + page - 1) / page * page + page;
  char* base = static_cast<char*>(mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
  if (base == MAP_FAILED) return nullptr;
  mprotect(base + len - page, page, PROT_NONE);
  void* user = base + len - page - body;
  std::lock_guard<std::mutex> lk(g_guard_mu);
  g_guarded[user] = {base, len};
  return user;
}
bool guarded_free(void* p) {
  std::lock_guard<std::mutex> lk(g_guard_mu);
  auto it = g_guarded.find(p);
  if (it == g_guarded.end()) return false;
  munmap(it->second.first, it->second.second);
  g_guarded.erase(it);
  return true;
}
}  // namespace emu
