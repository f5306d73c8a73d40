// Shared helpers for libunit_b200 (sm_100a).  Error plumbing, launch counting, exact-rounding float ops.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/unit_b200.h"

namespace unit {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define UNIT_REQUIRE(cond, ...)                \
  do {                                         \
    if (!(cond)) {                             \
      ::unit::set_error(__VA_ARGS__);          \
      return UNIT_EINVAL;                      \
    }                                          \
  } while (0)

#define UNIT_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::unit::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                        \
      return UNIT_ECUDA;                                                                  \
    }                                                                                     \
  } while (0)

#define UNIT_CHECK_LAUNCH(name)                                                        \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      ::unit::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));      \
      return UNIT_ECUDA;                                                               \
    }                                                                                  \
    ::unit::count_launch();                                                            \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int sm_count();

// The driver entry points this library calls (cuTensorMapEncodeTiled) need a context current on the CALLING thread.  A
// thread that has made no runtime call yet has none -- e.g. a PyTorch autograd worker whose first CUDA action is a
// call into this library -- and the driver answers CUDA_ERROR_INVALID_CONTEXT.  When no context is current, bind the
// primary context of the device that owns `dev_ptr` (the device the call is about to launch on).  No effect otherwise.
int ensure_driver_context(const void* dev_ptr);

// Developer switches (kernel-variant experiments), read from the environment ONCE per process -- never on a launch path.
struct Switches {
  bool filter_single, nms_single, fwd_v3, bwd_v4, bwd_cl1, paste_flat;
  int roi_debug, bwd_promo, bwd_evict_first, bwd2_evict_first, bwd_sweep3;
};
const Switches& switches();

// IoU with every operation individually rounded to fp32 (the CPU oracle has no FMA contraction):
//   inter / ((area_a + area_b) - inter), guarded by inter > 0  ([D2] pairwise_iou / [TV] nms).
__device__ __forceinline__ float box_area_rn(float x1, float y1, float x2, float y2) {
  return __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
}
__device__ __forceinline__ float box_inter_rn(float ax1, float ay1, float ax2, float ay2, float bx1, float by1,
                                              float bx2, float by2) {
  float w = fmaxf(__fsub_rn(fminf(ax2, bx2), fmaxf(ax1, bx1)), 0.f);
  float h = fmaxf(__fsub_rn(fminf(ay2, by2), fmaxf(ay1, by1)), 0.f);
  return __fmul_rn(w, h);
}
__device__ __forceinline__ float iou_from_rn(float inter, float area_a, float area_b) {
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

}  // namespace unit
