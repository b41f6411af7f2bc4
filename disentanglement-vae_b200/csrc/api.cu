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
}  // namespace dvae

extern "C" const char* dvae_last_error_string(void) { return dvae::g_err; }
extern "C" int dvae_version(void) { return 100; }
extern "C" int64_t dvae_launch_count(void) { return dvae::g_launches.load(); }
