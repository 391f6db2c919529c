// Error reporting, launch accounting and ABI version of librnerf_b200.so.
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace rnerf {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

}  // namespace rnerf

extern "C" {

int rnerf_abi_version(void) { return RNERF_ABI_VERSION; }
const char* rnerf_last_error(void) { return rnerf::g_err; }
uint64_t rnerf_launch_count(void) { return rnerf::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
