// Library-wide plumbing of the C-ABI: version, error string, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace gsn {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return GSN_E_CUDA;
  }
  return GSN_OK;
}

}  // namespace gsn

extern "C" int gsn_version(void) { return 100; }
extern "C" const char *gsn_last_error(void) { return gsn::g_err; }
extern "C" unsigned long long gsn_launch_count(void) { return gsn::g_launches.load(); }
