// Geometry shared by every ROIAlign kernel: the exact fp32 operation order of torchvision's roi_align kernels,
// so that sampling-grid sizes, floor() and validity decisions are identical to the reference's.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace unit {
namespace roi {

struct Geom {
  float start_w, start_h, bin_w, bin_h;
  int gw, gh;
  float count;
};

// torchvision roi_align_kernel: same fp32 operation order (separate roundings).
__device__ __forceinline__ Geom roi_geom(const float* roi, float scale, int ph, int pw, int sampling_ratio,
                                         int aligned) {
  Geom g;
  const float offset = aligned ? 0.5f : 0.f;
  g.start_w = __fsub_rn(__fmul_rn(roi[1], scale), offset);
  g.start_h = __fsub_rn(__fmul_rn(roi[2], scale), offset);
  const float end_w = __fsub_rn(__fmul_rn(roi[3], scale), offset);
  const float end_h = __fsub_rn(__fmul_rn(roi[4], scale), offset);
  float rw = __fsub_rn(end_w, g.start_w);
  float rh = __fsub_rn(end_h, g.start_h);
  if (!aligned) {
    rw = fmaxf(rw, 1.f);
    rh = fmaxf(rh, 1.f);
  }
  g.bin_w = __fdiv_rn(rw, (float)pw);
  g.bin_h = __fdiv_rn(rh, (float)ph);
  g.gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)pw));
  g.gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)ph));
  const int c = g.gw * g.gh;
  g.count = (float)(c > 1 ? c : 1);
  return g;
}

// coordinate of sample `i` of bin `p`:  start + p*bin + (i + .5)*bin/grid   (left to right, fp32 each)
__device__ __forceinline__ float sample_coord(float start, float bin, int p, int i, int grid) {
  return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                   __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)grid));
}

// One axis of pre_calc_for_bilinear_interpolate: returns validity, low index and weights.
__device__ __forceinline__ bool axis_tap(float v, int size, int& lo, int& hi, float& l, float& h) {
  if (v < -1.0f || v > (float)size) {
    lo = v < -1.0f ? 0 : size - 1;
    hi = lo;
    l = 0.f;
    h = 0.f;
    return false;
  }
  if (v <= 0.f) v = 0.f;
  lo = (int)v;
  if (lo >= size - 1) {
    hi = lo = size - 1;
    v = (float)lo;
  } else {
    hi = lo + 1;
  }
  l = __fsub_rn(v, (float)lo);
  h = __fsub_rn(1.f, l);
  return true;
}

// ---------------------------------------------------------------------------------------------- generic path
// One thread per output element, taps read through L1/L2 (any P, C, roi order; used when the slab kernel's
// preconditions do not hold).
template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }


}  // namespace roi
}  // namespace unit
