// Library-level entry points: version, thread-local error text, launch counter.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace unit {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

}  // namespace unit

extern "C" {

int unit_version(void) { return 100; }  // 0.1.0

const char* unit_last_error(void) { return unit::g_err; }

unsigned long long unit_launch_count(void) { return unit::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
