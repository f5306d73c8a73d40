// Library-level entry points: version, thread-local error text, launch counter.
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <cuda.h>

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

int ensure_driver_context(const void* dev_ptr) {
  typedef CUresult (*CtxGetCurrentFn)(CUcontext*);
  static CtxGetCurrentFn get_current = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return (CtxGetCurrentFn)ptr;
  }();
  CUcontext ctx = nullptr;
  if (get_current && get_current(&ctx) == CUDA_SUCCESS && ctx != nullptr) return UNIT_OK;
  int dev = -1;
  cudaPointerAttributes attr;
  if (dev_ptr && cudaPointerGetAttributes(&attr, dev_ptr) == cudaSuccess && attr.type == cudaMemoryTypeDevice)
    dev = attr.device;
  (void)cudaGetLastError();  // a host pointer is not an error here
  if (dev < 0) UNIT_CUDA(cudaGetDevice(&dev));
  UNIT_CUDA(cudaSetDevice(dev));  // CUDA 12: initialises the runtime and makes the primary context current
  return UNIT_OK;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

const Switches& switches() {
  static const Switches s = [] {
    Switches v;
    v.filter_single = getenv("UNIT_FILTER_SINGLE") != nullptr;
    v.nms_single = getenv("UNIT_NMS_SINGLE") != nullptr;
    v.fwd_v3 = getenv("UNIT_ROI_FWD_V3") != nullptr;
    v.bwd_v4 = getenv("UNIT_ROI_BWD_V4") != nullptr;
    v.bwd_cl1 = getenv("UNIT_ROI_BWD_CL1") != nullptr;
    v.paste_flat = getenv("UNIT_PASTE_FLAT") != nullptr;
    v.roi_debug = env_int("UNIT_ROI_DEBUG", 0);
    v.bwd_promo = env_int("UNIT_ROI_BWD_PROMO", 2);
    v.bwd_evict_first = env_int("UNIT_ROI_BWD_EVICT_FIRST", 1);
    v.bwd2_evict_first = env_int("UNIT_ROI_BWD2_EVICT_FIRST", 0);
    v.bwd_sweep3 = env_int("UNIT_ROI_BWD_SWEEP3", 1);
    return v;
  }();
  return s;
}

}  // namespace unit

extern "C" {

int unit_version(void) { return 100; }  // 0.1.0

const char* unit_last_error(void) { return unit::g_err; }

#ifndef UNIT_SOURCE_DIGEST
#define UNIT_SOURCE_DIGEST "unknown"
#endif
static const char g_digest[] = "UNIT_SOURCE_DIGEST=" UNIT_SOURCE_DIGEST;  // marker read by build.embedded_digest
const char* unit_source_digest(void) { return g_digest + 19; }

unsigned long long unit_launch_count(void) { return unit::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
