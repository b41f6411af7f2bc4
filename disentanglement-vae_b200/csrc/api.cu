// Error / bookkeeping entry points of the C ABI (see include/dvae_b200.h).
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace dvae {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool force_simt_gemm() {
  const char* e = getenv("DVAE_GEMM_IMPL");
  return e && e[0] == 's';
}

// ---- fork / join side streams ---------------------------------------------------------------------
int join_side_streams(cudaStream_t main);
void set_defer_joins(bool on);
namespace {
struct SideStreams {
  cudaStream_t s[3];
  cudaEvent_t fork_ev, join_ev[3];
  bool ok = false;
  bool pending[3] = {false, false, false};     // detached work (deferred joins)
};
thread_local bool g_defer_joins = false;
SideStreams* side_streams() {
  static thread_local SideStreams* per_dev[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!per_dev[dev]) {
    SideStreams* ss = new SideStreams();
    ss->ok = cudaEventCreateWithFlags(&ss->fork_ev, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 3 && ss->ok; ++i)
      ss->ok = cudaStreamCreateWithFlags(&ss->s[i], cudaStreamNonBlocking) == cudaSuccess &&
               cudaEventCreateWithFlags(&ss->join_ev[i], cudaEventDisableTiming) == cudaSuccess;
    per_dev[dev] = ss;
  }
  return per_dev[dev]->ok ? per_dev[dev] : nullptr;
}
bool fork_enabled() {
  const char* e = getenv("DVAE_FORK");
  return !(e && e[0] == '0');
}
}  // namespace

Fork::Fork(cudaStream_t main) : main_(main), ok_(false) {
  if (!fork_enabled()) return;
  SideStreams* ss = side_streams();
  if (!ss) return;
  ok_ = cudaEventRecord(ss->fork_ev, main_) == cudaSuccess;
}
cudaStream_t Fork::side(int i) {
  if (!ok_) return main_;
  SideStreams* ss = side_streams();
  if (!used_[i]) {
    if (cudaStreamWaitEvent(ss->s[i], ss->fork_ev, 0) != cudaSuccess) return main_;
    used_[i] = true;
  }
  return ss->s[i];
}
int Fork::join() {
  if (!ok_) return DVAE_OK;
  SideStreams* ss = side_streams();
  for (int i = 0; i < 3; ++i)
    if (used_[i]) {
      DVAE_CUDA(cudaEventRecord(ss->join_ev[i], ss->s[i]));
      DVAE_CUDA(cudaStreamWaitEvent(main_, ss->join_ev[i], 0));
      used_[i] = false;
      ss->pending[i] = false;      // stream order: everything detached earlier on this side stream is covered too
    }
  return DVAE_OK;
}
int Fork::join_or_defer() {
  if (!ok_ || !g_defer_joins) return join();
  SideStreams* ss = side_streams();
  for (int i = 0; i < 3; ++i)
    if (used_[i]) {
      ss->pending[i] = true;
      used_[i] = false;
    }
  return DVAE_OK;
}
int join_side_streams(cudaStream_t main) {
  SideStreams* ss = side_streams();
  if (!ss) return DVAE_OK;
  for (int i = 0; i < 3; ++i)
    if (ss->pending[i]) {
      DVAE_CUDA(cudaEventRecord(ss->join_ev[i], ss->s[i]));
      DVAE_CUDA(cudaStreamWaitEvent(main, ss->join_ev[i], 0));
      ss->pending[i] = false;
    }
  return DVAE_OK;
}
void set_defer_joins(bool on) { g_defer_joins = on; }
}  // namespace dvae

extern "C" int dvae_defer_joins(int on) {
  dvae::set_defer_joins(on != 0);
  return DVAE_OK;
}
extern "C" int dvae_join_side_streams(void* stream) {
  dvae::set_defer_joins(false);
  return dvae::join_side_streams((cudaStream_t)stream);
}
extern "C" const char* dvae_last_error_string(void) { return dvae::g_err; }
extern "C" int dvae_version(void) { return 100; }
extern "C" int64_t dvae_launch_count(void) { return dvae::g_launches.load(); }
