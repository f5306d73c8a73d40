// Shared pieces of the slab-resident ROIAlign kernels (forward: roi_align_fwd.cu, backward: roi_align_bwd.cu):
// per-warp work areas, sampling-table construction, channel-pair interleaved slab loader, TMA bulk helpers.
#pragma once
#include <stdlib.h>

#include <algorithm>

#include "roi_common.cuh"

namespace unit {
namespace roi {
namespace v2 {

constexpr int P = 14;
constexpr int CS = 8;           // channels per slab
constexpr int NPAIR = CS / 2;   // channel pairs
constexpr int MAXG = 5;         // sampling grid per bin handled with tables (RoI side <= 70 feature px)
constexpr int MAXS = P * MAXG;
constexpr int NWARPS = 12;
constexpr int NTHREADS = NWARPS * 32;

struct __align__(16) YTap {
  int lo;    // row * W of the lower tap
  float h;   // its weight (hy)
  float l;   // weight of the upper tap (ly)
  int hi;    // row * W of the upper tap
};

struct __align__(16) Header {
  int gw, gh;
  float inv_count;
  int mode;   // 0 zero output, 1 tables, 2 direct evaluation
  int x0;     // first column of the sliding window
  int nsamp;  // 14 * gw
  float start_w, start_h;
  float bin_w, bin_h;
  int pad0, pad1;
};

// per-warp shared-memory area
template <typename T>
struct WarpArea {
  Header hdr;
  float2 xtab[MAXS + 2];  // (hx | ADV sign, lx | END sign)
  YTap ytab[MAXS];
  __align__(16) T stage[CS * P * P];
};

struct Params {
  const void* feat;
  const float* rois;
  void* out;
  const int* img_off;
  int N, C, H, W, R;
  float scale;
  int sampling_ratio, aligned;
  int pair_stride;  // float2 elements per channel-pair plane, == 1 (mod 16)
  long long units_total;
  int debug;
};

__device__ __forceinline__ float2 ffma2(float s, float2 v, float2 acc) {
  unsigned long long a, b, c, d;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(v.x), "f"(v.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.x), "f"(acc.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(d));
  return r;
}

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Build the tables of one RoI with the whole warp.
//   x entry s: (hx * inv_count | LASTCOL sign, lx * inv_count | END sign); LASTCOL = last sample whose lower tap is
//   this column, END = last sample of its bin.  Samples clamped to the last column (x >= W-1, value F[W-1]) are
//   re-expressed as lo = W-2 with weights (0, 1): identical value, and the window never has to step past W-1.
template <typename T>
__device__ __forceinline__ void build_tables(const float* __restrict__ roi, const Params& p, WarpArea<T>* wa,
                                             int lane) {
  const Geom g = roi_geom(roi, p.scale, P, P, p.sampling_ratio, p.aligned);
  int mode = 1;
  if (g.gw <= 0 || g.gh <= 0) mode = 0;
  else if (g.gw > MAXG || g.gh > MAXG || p.W < 2) mode = 2;
  bool jump = false;
  if (mode == 1) {
    const int ns = P * g.gw;
    const float inv = 1.f / g.count;
    const float inv_gw = 1.f / (float)g.gw;
    // 31 samples per round; lane 31 evaluates the look-ahead sample of the next round's lane 0
    for (int s0 = 0; s0 < ns; s0 += 31) {
      const int s = s0 + lane;
      int lo = 0x7fffffff, hi;
      float l = 0.f, h = 0.f;
      if (s < ns) {
        const int pw = (int)(((float)s + 0.5f) * inv_gw);  // s / gw (exact for s < 2^20)
        axis_tap(sample_coord(g.start_w, g.bin_w, pw, s - pw * g.gw, g.gw), p.W, lo, hi, l, h);
        if (lo >= p.W - 1) {  // clamped sample: 0 * F[W-2] + 1 * F[W-1]  (or 0 * .. + 0 * .. when invalid)
          lo = p.W - 2;
          l = h;  // valid: h == 1, l == 0 -> (0, 1); invalid: (0, 0)
          h = 0.f;
        }
      }
      const int nlo = __shfl_down_sync(0xffffffffu, lo, 1);
      if (s < ns && lane < 31) {
        const bool last = (s == ns - 1) || (nlo != lo);
        jump |= (s < ns - 1) && (nlo - lo > 1 || nlo < lo);
        const int pw = (int)(((float)s + 0.5f) * inv_gw);
        const bool end = (s - pw * g.gw) == g.gw - 1;
        float2 e;
        e.x = __uint_as_float(__float_as_uint(h * inv) | (last ? 0x80000000u : 0u));
        e.y = __uint_as_float(__float_as_uint(l * inv) | (end ? 0x80000000u : 0u));
        wa->xtab[s] = e;
        if (s == 0) wa->hdr.x0 = lo;
      }
    }
    if (lane < 2) wa->xtab[ns + lane] = make_float2(0.f, 0.f);
    const float inv_gh = 1.f / (float)g.gh;
    for (int s = lane; s < P * g.gh; s += 32) {
      YTap t;
      const int ph = (int)(((float)s + 0.5f) * inv_gh);
      axis_tap(sample_coord(g.start_h, g.bin_h, ph, s - ph * g.gh, g.gh), p.H, t.lo, t.hi, t.l, t.h);
      t.lo *= p.W;
      t.hi *= p.W;
      wa->ytab[s] = t;
    }
  }
  // a sample step larger than one column can only come from fp32 rounding of bin/grid; evaluate such RoIs directly
  if (__any_sync(0xffffffffu, jump)) mode = 2;
  if (lane == 0) {
    wa->hdr.gw = g.gw;
    wa->hdr.gh = g.gh;
    wa->hdr.inv_count = 1.f / g.count;
    wa->hdr.mode = mode;
    wa->hdr.nsamp = P * g.gw;
    wa->hdr.start_w = g.start_w;
    wa->hdr.start_h = g.start_h;
    wa->hdr.bin_w = g.bin_w;
    wa->hdr.bin_h = g.bin_h;
  }
}

// slab loader: global planar [c][HW] -> shared [c/2][HW(+pad)][2]
template <typename T>
__device__ __forceinline__ void load_slab(const T* __restrict__ src, float* __restrict__ slab, int HW, int ps, int tid);
template <>
__device__ __forceinline__ void load_slab<float>(const float* __restrict__ src, float* __restrict__ slab, int HW,
                                                 int ps, int tid) {
  const int total = CS * HW;
  if ((((uintptr_t)src) & 15) == 0 && (HW & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int i = tid; i < (total >> 2); i += NTHREADS) {
      const float4 v = __ldg(s4 + i);
      const int e = i << 2;
      const int c = e / HW, o = e - c * HW;
      float* d = slab + ((size_t)(c >> 1) * ps + o) * 2 + (c & 1);
      d[0] = v.x;
      d[2] = v.y;
      d[4] = v.z;
      d[6] = v.w;
    }
  } else {
    for (int e = tid; e < total; e += NTHREADS) {
      const int c = e / HW, o = e - c * HW;
      slab[((size_t)(c >> 1) * ps + o) * 2 + (c & 1)] = __ldg(src + e);
    }
  }
}
template <>
__device__ __forceinline__ void load_slab<__nv_bfloat16>(const __nv_bfloat16* __restrict__ src,
                                                         float* __restrict__ slab, int HW, int ps, int tid) {
  const int total = CS * HW;
  if ((((uintptr_t)src) & 15) == 0 && (HW & 7) == 0) {
    const uint4* s8 = reinterpret_cast<const uint4*>(src);
    for (int i = tid; i < (total >> 3); i += NTHREADS) {
      const uint4 v = __ldg(s8 + i);
      const int e = i << 3;
      const int c = e / HW, o = e - c * HW;
      float* d = slab + ((size_t)(c >> 1) * ps + o) * 2 + (c & 1);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        d[4 * k] = __uint_as_float(w[k] << 16);
        d[4 * k + 2] = __uint_as_float(w[k] & 0xffff0000u);
      }
    }
  } else {
    for (int e = tid; e < total; e += NTHREADS) {
      const int c = e / HW, o = e - c * HW;
      slab[((size_t)(c >> 1) * ps + o) * 2 + (c & 1)] = __bfloat162float(src[e]);
    }
  }
}

inline int pair_stride_host(int HW) {
  int s = HW + 3;  // the window reads up to 3 elements past a row end (never used; see build_tables)
  while ((s & 15) != 1) ++s;
  return s;
}


}  // namespace v2
}  // namespace roi
}  // namespace unit
