// nccl_emu.cpp — TEST INFRASTRUCTURE.  The five NCCL entry points csrc/comm.cu binds at run time, for ranks that are THREADS of one test
// process over the emulated build of the library (build_engine_emu.py points comm.cu's dlopen at this file's library).  An all-gather is
// a barrier: every rank publishes its send buffer, waits for the others, copies all blocks into its own receive buffer, waits again (so no
// rank reuses its send buffer while another still reads it).
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace {
struct Group {
  int world = 0;
  std::vector<const void*> send;
  int arrived = 0;
  uint64_t generation = 0;
  int refs = 0;
};
struct Comm { Group* g; int rank; };
std::mutex g_mu;
std::condition_variable g_cv;
std::map<uint64_t, Group*> g_groups;
uint64_t g_next_id = 1;

void barrier(Group* g, std::unique_lock<std::mutex>& lk) {
  const uint64_t gen = g->generation;
  if (++g->arrived == g->world) { g->arrived = 0; g->generation++; g_cv.notify_all(); }
  else g_cv.wait(lk, [&] { return g->generation != gen; });
}
}  // namespace

extern "C" {
typedef struct { char internal[128]; } ncclUniqueId;
int ncclGetUniqueId(ncclUniqueId* id) {
  std::lock_guard<std::mutex> lk(g_mu);
  std::memset(id, 0, sizeof(*id));
  const uint64_t v = g_next_id++;
  std::memcpy(id->internal, &v, 8);
  return 0;
}
int ncclCommInitRank(void** comm, int world, ncclUniqueId id, int rank) {
  uint64_t v;
  std::memcpy(&v, id.internal, 8);
  std::lock_guard<std::mutex> lk(g_mu);
  Group*& g = g_groups[v];
  if (!g) { g = new Group(); g->world = world; g->send.assign(size_t(world), nullptr); }
  if (g->world != world || rank < 0 || rank >= world) return 4;
  g->refs++;
  *comm = new Comm{g, rank};
  return 0;
}
int ncclCommDestroy(void* comm) {
  delete static_cast<Comm*>(comm);
  return 0;
}
int ncclAllGather(const void* send, void* recv, size_t count, int dtype, void* comm, void* /*stream*/) {
  Comm* c = static_cast<Comm*>(comm);
  const size_t bytes = count * (dtype == 4 || dtype == 5 || dtype == 8 ? 8u : (dtype == 2 || dtype == 3 || dtype == 7 ? 4u : 1u));   // int64 / uint64 / f64 : int32 / uint32 / f32
  std::unique_lock<std::mutex> lk(g_mu);
  Group* g = c->g;
  g->send[size_t(c->rank)] = send;
  barrier(g, lk);
  for (int r = 0; r < g->world; r++) std::memcpy(static_cast<char*>(recv) + size_t(r) * bytes, g->send[size_t(r)], bytes);
  barrier(g, lk);
  return 0;
}
const char* ncclGetErrorString(int) { return "emulated NCCL error"; }
}
