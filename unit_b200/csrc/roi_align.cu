// ROIAlign forward / backward for the C4 RoI stage.
//
// Replaces [D2] ROIPooler -> ROIAlign(aligned=True) -> [TV] roi_align / _roi_align_backward
// (reference call sites: modeling/roi_heads/roi_heads.py:356,364,499,511,598,610,708,715,729,829,843,911).
//
// Slab-resident kernel (the hot path): one persistent CTA per SM walks a contiguous range of
// (image, 8-channel slab, RoI) units.  The slab -- 8 full H x W channel planes of one image, 134 KB for the
// 50x84 res4 map -- is staged once in shared memory, so every feature byte is read from HBM/L2 once per slab and
// all bilinear taps are shared-memory reads.  Per RoI a warp builds the separable sampling tables (x: 14*gw
// samples, y: 14*gh samples) with the exact fp32 operation order of the torchvision kernel, so floor / validity
// decisions are identical.  A thread owns (channel, bin-row pair): it slides a two-column window of vertically
// interpolated values across the RoI (each feature column is combined once per bin row instead of once per
// sample) and keeps its 2 x 14 outputs in registers.  Outputs are staged in shared memory as the exact
// [8 ch][14][14] block of the NCHW output and leave the SM as one 6272-byte TMA bulk store per RoI
// (cp.async.bulk.global.shared::cta), i.e. fully coalesced 16-byte-aligned writes, which are 96 % of the bytes.
//
// Algorithmic bytes per image (fp32): C*H*W*4 + R*20 + R*C*196*4  (SURVEY.md section 8d).
#include <cuda_bf16.h>
#include <stdlib.h>
#include <algorithm>

#include "roi_common.cuh"

namespace unit {
namespace roi {

constexpr int P = 14;            // pooled size of the fast path
constexpr int CS = 8;            // channels per slab
constexpr int MAXG = 6;          // max sampling grid per bin handled with tables (RoI side <= 84 feature px)
constexpr int MAXS = P * MAXG;   // table entries per axis
constexpr int NB = 8;            // RoIs per batch
constexpr int NTHREADS = 512;    // 16 warps: 2 warp tasks per RoI of the batch

struct __align__(16) Tap {       // one sample along one axis
  int lo;                        // x: column index; y: row offset (row * W)
  float h;                       // weight of lo   (hy / hx)
  float l;                       // weight of lo+1 (ly / lx)
  int hi;                        // x: column index of the upper tap; y: row offset of it
};

struct RoiHeader {
  int gw, gh;
  float inv_count;
  int mode;  // 0 = zero output, 1 = tables, 2 = direct (grid larger than MAXG)
  float start_w, start_h, bin_w, bin_h;
};

template <typename T>
__global__ void roi_align_fwd_generic(const T* __restrict__ feat, const float* __restrict__ rois, T* __restrict__ out,
                                      int N, int C, int H, int W, long long total, int PH, int PW, float scale,
                                      int sampling_ratio, int aligned) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(idx % PW);
    const int ph = (int)((idx / PW) % PH);
    const int c = (int)((idx / ((long long)PW * PH)) % C);
    const int r = (int)(idx / ((long long)PW * PH * C));
    const float* roi = rois + (long long)r * 5;
    const int n = (int)roi[0];
    float acc = 0.f;
    Geom g = roi_geom(roi, scale, PH, PW, sampling_ratio, aligned);
    if (n >= 0 && n < N) {
      const T* plane = feat + ((long long)n * C + c) * H * W;
      for (int iy = 0; iy < g.gh; ++iy) {
        int ylo, yhi;
        float ly, hy;
        const bool vy = axis_tap(sample_coord(g.start_h, g.bin_h, ph, iy, g.gh), H, ylo, yhi, ly, hy);
        for (int ix = 0; ix < g.gw; ++ix) {
          int xlo, xhi;
          float lx, hx;
          const bool vx = axis_tap(sample_coord(g.start_w, g.bin_w, pw, ix, g.gw), W, xlo, xhi, lx, hx);
          if (vy && vx) {
            const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx), w3 = __fmul_rn(ly, hx),
                        w4 = __fmul_rn(ly, lx);
            const float v = __fadd_rn(
                __fadd_rn(__fadd_rn(__fmul_rn(w1, ldf(plane + ylo * W + xlo)), __fmul_rn(w2, ldf(plane + ylo * W + xhi))),
                          __fmul_rn(w3, ldf(plane + yhi * W + xlo))),
                __fmul_rn(w4, ldf(plane + yhi * W + xhi)));
            acc = __fadd_rn(acc, v);
          }
        }
      }
    }
    stf(out + idx, __fdiv_rn(acc, g.count));
  }
}

template <typename T>
__global__ void roi_align_bwd_generic(const T* __restrict__ gout, const float* __restrict__ rois,
                                      float* __restrict__ gfeat, int N, int C, int H, int W, long long total, int PH,
                                      int PW, float scale, int sampling_ratio, int aligned) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(idx % PW);
    const int ph = (int)((idx / PW) % PH);
    const int c = (int)((idx / ((long long)PW * PH)) % C);
    const int r = (int)(idx / ((long long)PW * PH * C));
    const float* roi = rois + (long long)r * 5;
    const int n = (int)roi[0];
    if (n < 0 || n >= N) continue;
    Geom g = roi_geom(roi, scale, PH, PW, sampling_ratio, aligned);
    const float go = ldf(gout + idx);
    float* plane = gfeat + ((long long)n * C + c) * H * W;
    for (int iy = 0; iy < g.gh; ++iy) {
      int ylo, yhi;
      float ly, hy;
      const bool vy = axis_tap(sample_coord(g.start_h, g.bin_h, ph, iy, g.gh), H, ylo, yhi, ly, hy);
      for (int ix = 0; ix < g.gw; ++ix) {
        int xlo, xhi;
        float lx, hx;
        const bool vx = axis_tap(sample_coord(g.start_w, g.bin_w, pw, ix, g.gw), W, xlo, xhi, lx, hx);
        if (vy && vx) {
          atomicAdd(plane + ylo * W + xlo, go * hy * hx / g.count);
          atomicAdd(plane + ylo * W + xhi, go * hy * lx / g.count);
          atomicAdd(plane + yhi * W + xlo, go * ly * hx / g.count);
          atomicAdd(plane + yhi * W + xhi, go * ly * lx / g.count);
        }
      }
    }
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

// per-image RoI offsets from the (sorted) batch-index column: off[n] = first r with batch_idx >= n
__global__ void roi_offsets_kernel(const float* __restrict__ rois, int R, int N, int* __restrict__ off) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > R) return;
  const int cur = r < R ? min(max((int)rois[(long long)r * 5], 0), N) : N;
  const int prev = r > 0 ? min(max((int)rois[(long long)(r - 1) * 5], 0), N) : -1;
  for (int n = prev + 1; n <= cur; ++n) off[n] = r;
}

// ---------------------------------------------------------------------------------------------- slab kernel
struct SlabSmem {
  // dynamic shared memory layout (offsets in bytes computed on the host and passed in)
  int plane_stride;  // floats; == 1 (mod 32) so the 8 channel lanes of a row hit 8 consecutive banks
};

__device__ __forceinline__ int plane_stride_for(int HW) {
  int s = HW;
  while ((s & 31) != 1) ++s;
  return s;
}

// Build the sampling tables of one RoI (one warp).
__device__ __forceinline__ void build_tables(const float* roi, float scale, int sampling_ratio, int aligned, int H,
                                             int W, RoiHeader* hdr, Tap* xtab, Tap* ytab, int lane) {
  const Geom g = roi_geom(roi, scale, P, P, sampling_ratio, aligned);
  int mode = 1;
  if (g.gw <= 0 || g.gh <= 0) mode = 0;
  else if (g.gw > MAXG || g.gh > MAXG) mode = 2;
  if (lane == 0) {
    hdr->gw = g.gw;
    hdr->gh = g.gh;
    hdr->inv_count = 1.f / g.count;
    hdr->mode = mode;
    hdr->start_w = g.start_w;
    hdr->start_h = g.start_h;
    hdr->bin_w = g.bin_w;
    hdr->bin_h = g.bin_h;
  }
  if (mode != 1) return;
  for (int s = lane; s < P * g.gw; s += 32) {
    Tap t;
    axis_tap(sample_coord(g.start_w, g.bin_w, s / g.gw, s % g.gw, g.gw), W, t.lo, t.hi, t.l, t.h);
    xtab[s] = t;
  }
  for (int s = lane; s < P * g.gh; s += 32) {
    Tap t;
    axis_tap(sample_coord(g.start_h, g.bin_h, s / g.gh, s % g.gh, g.gh), H, t.lo, t.hi, t.l, t.h);
    t.lo *= W;
    t.hi *= W;
    ytab[s] = t;
  }
}

// Vertically interpolated value of column `col` for the two bin rows of this thread.
template <int GH>
struct VTaps {
  Tap a[GH > 0 ? GH : 1], b[GH > 0 ? GH : 1];
};

template <typename T>
__device__ __forceinline__ void load_slab(const T* __restrict__ src, float* __restrict__ slab, int HW, int ps,
                                          int tid, int nthreads);
template <>
__device__ __forceinline__ void load_slab<float>(const float* __restrict__ src, float* __restrict__ slab, int HW,
                                                 int ps, int tid, int nthreads) {
  // src: CS contiguous planes of HW floats.  16-byte vector loads when the slab base is aligned.
  const long long total = (long long)CS * HW;
  if ((((uintptr_t)src) & 15) == 0 && (HW & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    const int n4 = (int)(total >> 2);
    for (int i = tid; i < n4; i += nthreads) {
      const float4 v = __ldg(s4 + i);
      const int e = i << 2;
      const int c = e / HW, o = e - c * HW;
      float* d = slab + c * ps + o;
      d[0] = v.x;
      d[1] = v.y;
      d[2] = v.z;
      d[3] = v.w;
    }
  } else {
    for (int e = tid; e < (int)total; e += nthreads) {
      const int c = e / HW, o = e - c * HW;
      slab[c * ps + o] = __ldg(src + e);
    }
  }
}
template <>
__device__ __forceinline__ void load_slab<__nv_bfloat16>(const __nv_bfloat16* __restrict__ src,
                                                         float* __restrict__ slab, int HW, int ps, int tid,
                                                         int nthreads) {
  const long long total = (long long)CS * HW;
  if ((((uintptr_t)src) & 15) == 0 && (HW & 7) == 0) {
    const uint4* s8 = reinterpret_cast<const uint4*>(src);
    const int n8 = (int)(total >> 3);
    for (int i = tid; i < n8; i += nthreads) {
      const uint4 v = __ldg(s8 + i);
      const int e = i << 3;
      const int c = e / HW, o = e - c * HW;
      float* d = slab + c * ps + o;
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        d[2 * k] = __uint_as_float(w[k] << 16);
        d[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
      }
    }
  } else {
    for (int e = tid; e < (int)total; e += nthreads) {
      const int c = e / HW, o = e - c * HW;
      slab[c * ps + o] = __bfloat162float(src[e]);
    }
  }
}

struct FwdParams {
  const void* feat;
  const float* rois;
  void* out;
  const int* img_off;  // [N+1]
  int N, C, H, W, R;
  float scale;
  int sampling_ratio, aligned;
  int plane_stride;
  long long units_total;  // R * (C / CS)
  int debug;              // experiments only (UNIT_ROI_DEBUG): bit0 skip compute, bit1 skip stores
};

// shared memory carve-up (bytes): slab | staging | xtab | ytab | headers
template <typename T>
__host__ __device__ inline size_t smem_stage_bytes() { return (size_t)NB * CS * P * P * sizeof(T); }
inline size_t smem_bytes_total(int plane_stride, size_t stage_bytes) {
  return (size_t)CS * plane_stride * sizeof(float) + stage_bytes + 2 * (size_t)NB * MAXS * sizeof(Tap) +
         (size_t)NB * sizeof(RoiHeader) + 64;
}

// ---------------------------------------------------------------------------------------------- slab backward
// One CTA per (image, 8-channel slab).  The gradient tile of the slab is accumulated in shared memory over all
// RoIs of the image (transposed sliding window: horizontal spread in registers, vertical spread with shared-memory
// atomics) and written to HBM once with plain coalesced stores: no global atomics, no pre-zeroed output.
template <int GH>
__device__ __forceinline__ void bwd_flush(float* __restrict__ plane, const VTaps<GH>& vt, const Tap* ya,
                                          const Tap* yb, int gh, int col, float da, float db) {
  if (da == 0.f && db == 0.f) return;
  if (GH > 0) {
#pragma unroll
    for (int i = 0; i < GH; ++i) {
      atomicAdd(plane + vt.a[i].lo + col, vt.a[i].h * da);
      atomicAdd(plane + vt.a[i].hi + col, vt.a[i].l * da);
      atomicAdd(plane + vt.b[i].lo + col, vt.b[i].h * db);
      atomicAdd(plane + vt.b[i].hi + col, vt.b[i].l * db);
    }
  } else {
    for (int i = 0; i < gh; ++i) {
      const Tap a = ya[i], b = yb[i];
      atomicAdd(plane + a.lo + col, a.h * da);
      atomicAdd(plane + a.hi + col, a.l * da);
      atomicAdd(plane + b.lo + col, b.h * db);
      atomicAdd(plane + b.hi + col, b.l * db);
    }
  }
}

template <int GH>
__device__ __forceinline__ void bwd_task(float* __restrict__ plane, int W, const RoiHeader& hdr,
                                         const Tap* __restrict__ xtab, const Tap* __restrict__ ytab, int pp,
                                         const float* __restrict__ g) {
  const int gw = hdr.gw, gh = hdr.gh;
  const Tap* ya = ytab + (2 * pp) * gh;
  const Tap* yb = ytab + (2 * pp + 1) * gh;
  VTaps<GH> vt;
  if (GH > 0) {
#pragma unroll
    for (int i = 0; i < GH; ++i) {
      vt.a[i] = ya[i];
      vt.b[i] = yb[i];
    }
  }
  int cur = xtab[0].lo;
  float dlo_a = 0.f, dlo_b = 0.f, dhi_a = 0.f, dhi_b = 0.f;
  const float inv = hdr.inv_count;
#pragma unroll
  for (int pw = 0; pw < P; ++pw) {
    const float ga = g[pw] * inv, gb = g[P + pw] * inv;
    const Tap* xs = xtab + pw * gw;
    for (int ix = 0; ix < gw; ++ix) {
      const Tap e = xs[ix];
      while (e.lo > cur) {
        bwd_flush<GH>(plane, vt, ya, yb, gh, cur, dlo_a, dlo_b);
        ++cur;
        dlo_a = dhi_a;
        dlo_b = dhi_b;
        dhi_a = 0.f;
        dhi_b = 0.f;
      }
      dlo_a = fmaf(e.h, ga, dlo_a);
      dhi_a = fmaf(e.l, ga, dhi_a);
      dlo_b = fmaf(e.h, gb, dlo_b);
      dhi_b = fmaf(e.l, gb, dhi_b);
    }
  }
  bwd_flush<GH>(plane, vt, ya, yb, gh, cur, dlo_a, dlo_b);
  bwd_flush<GH>(plane, vt, ya, yb, gh, min(cur + 1, W - 1), dhi_a, dhi_b);
}

template <typename T>
__device__ __noinline__ void bwd_task_direct(float* __restrict__ plane, int H, int W, const RoiHeader& hdr, int pp,
                                             const T* __restrict__ g) {
  for (int half = 0; half < 2; ++half) {
    const int ph = 2 * pp + half;
    for (int pw = 0; pw < P; ++pw) {
      const float go = ldf(g + half * P + pw) * hdr.inv_count;
      for (int iy = 0; iy < hdr.gh; ++iy) {
        int ylo, yhi;
        float ly, hy;
        const bool vy = axis_tap(sample_coord(hdr.start_h, hdr.bin_h, ph, iy, hdr.gh), H, ylo, yhi, ly, hy);
        for (int ix = 0; ix < hdr.gw; ++ix) {
          int xlo, xhi;
          float lx, hx;
          const bool vx = axis_tap(sample_coord(hdr.start_w, hdr.bin_w, pw, ix, hdr.gw), W, xlo, xhi, lx, hx);
          if (vy && vx) {
            atomicAdd(plane + ylo * W + xlo, go * hy * hx);
            atomicAdd(plane + ylo * W + xhi, go * hy * lx);
            atomicAdd(plane + yhi * W + xlo, go * ly * hx);
            atomicAdd(plane + yhi * W + xhi, go * ly * lx);
          }
        }
      }
    }
  }
}

template <typename T>
__device__ __forceinline__ void stage_read(const T* src, float* v);  // 28 values
template <>
__device__ __forceinline__ void stage_read<float>(const float* src, float* v) {
  const float4* s = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const float4 q = s[i];
    v[4 * i] = q.x;
    v[4 * i + 1] = q.y;
    v[4 * i + 2] = q.z;
    v[4 * i + 3] = q.w;
  }
}
template <>
__device__ __forceinline__ void stage_read<__nv_bfloat16>(const __nv_bfloat16* src, float* v) {
  const uint2* s = reinterpret_cast<const uint2*>(src);
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const uint2 q = s[i];
    v[4 * i] = __uint_as_float(q.x << 16);
    v[4 * i + 1] = __uint_as_float(q.x & 0xffff0000u);
    v[4 * i + 2] = __uint_as_float(q.y << 16);
    v[4 * i + 3] = __uint_as_float(q.y & 0xffff0000u);
  }
}

template <typename T>
__global__ void __launch_bounds__(NTHREADS, 1) roi_align_bwd_slab(const FwdParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* slab = reinterpret_cast<float*>(smem_raw);
  T* stage = reinterpret_cast<T*>(smem_raw + (size_t)CS * p.plane_stride * sizeof(float));
  Tap* xtabs = reinterpret_cast<Tap*>(reinterpret_cast<unsigned char*>(stage) + smem_stage_bytes<T>());
  Tap* ytabs = xtabs + NB * MAXS;
  RoiHeader* hdrs = reinterpret_cast<RoiHeader*>(ytabs + NB * MAXS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nslab = p.C / CS;
  const int HW = p.H * p.W;
  const int n = blockIdx.x / nslab, k = blockIdx.x % nslab;
  const T* gout = reinterpret_cast<const T*>(p.feat);  // grad_out [R,C,14,14]
  T* gfeat = reinterpret_cast<T*>(p.out);              // grad_feat [N,C,H,W]
  const int r_base = p.img_off[n];
  const int Rn = p.img_off[n + 1] - r_base;

  for (int i = tid; i < CS * p.plane_stride; i += NTHREADS) slab[i] = 0.f;

  constexpr int VEC = 16 / sizeof(T);                    // elements per 16-byte vector
  constexpr int ROI_VECS = CS * P * P / VEC;             // 16-byte vectors per RoI block
  for (int rb = 0; rb < Rn; rb += NB) {
    const int nb = min(NB, Rn - rb);
    __syncthreads();  // previous batch fully consumed (also orders the zero fill)
    if (warp < nb)
      build_tables(p.rois + (long long)(r_base + rb + warp) * 5, p.scale, p.sampling_ratio, p.aligned, p.H, p.W,
                   hdrs + warp, xtabs + warp * MAXS, ytabs + warp * MAXS, lane);
    // stage grad_out blocks of the batch: nb x (8 ch x 196) contiguous elements each
    for (int i = tid; i < nb * ROI_VECS; i += NTHREADS) {
      const int b = i / ROI_VECS, o = i - b * ROI_VECS;
      const uint4* src = reinterpret_cast<const uint4*>(gout + ((long long)(r_base + rb + b) * p.C + (long long)k * CS) * (P * P));
      reinterpret_cast<uint4*>(stage + (size_t)b * CS * P * P)[o] = __ldg(src + o);
    }
    __syncthreads();
    const int b = warp >> 1;
    const int pp = ((warp & 1) << 2) + (lane >> 3);
    const int c = lane & 7;
    if (b < nb && pp < P / 2) {
      const RoiHeader hdr = hdrs[b];
      float* plane = slab + c * p.plane_stride;
      const T* src = stage + ((size_t)b * CS + c) * (P * P) + pp * 2 * P;
      if (hdr.mode == 1) {
        float g[2 * P];
        stage_read<T>(src, g);
        if (hdr.gh == 1) bwd_task<1>(plane, p.W, hdr, xtabs + b * MAXS, ytabs + b * MAXS, pp, g);
        else if (hdr.gh == 2) bwd_task<2>(plane, p.W, hdr, xtabs + b * MAXS, ytabs + b * MAXS, pp, g);
        else bwd_task<0>(plane, p.W, hdr, xtabs + b * MAXS, ytabs + b * MAXS, pp, g);
      } else if (hdr.mode == 2) {
        bwd_task_direct<T>(plane, p.H, p.W, hdr, pp, src);
      }
    }
  }
  __syncthreads();
  T* dst = gfeat + ((long long)n * p.C + (long long)k * CS) * HW;
  for (int e = tid; e < CS * HW; e += NTHREADS) {
    const int c = e / HW, o = e - c * HW;
    stf(dst + e, slab[c * p.plane_stride + o]);
  }
}

static int plane_stride_host(int HW) {
  int s = HW;
  while ((s & 31) != 1) ++s;
  return s;
}

template <typename T>
static int launch_slab(bool backward, const void* a, const float* rois, void* b, int N, int C, int H, int W, int R,
                       float scale, int sr, int aligned, int* img_off, cudaStream_t st) {
  FwdParams p;
  p.feat = a;
  p.rois = rois;
  p.out = b;
  p.img_off = img_off;
  p.N = N;
  p.C = C;
  p.H = H;
  p.W = W;
  p.R = R;
  p.scale = scale;
  p.sampling_ratio = sr;
  p.aligned = aligned;
  p.plane_stride = plane_stride_host(H * W);
  p.units_total = (long long)R * (C / CS);
  {
    const char* dbg = getenv("UNIT_ROI_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  const size_t smem = smem_bytes_total(p.plane_stride, smem_stage_bytes<T>());
  roi_offsets_kernel<<<cdiv(R + 1, 256), 256, 0, st>>>(rois, R, N, img_off);
  UNIT_CHECK_LAUNCH("roi_offsets_kernel");
  if (!backward) {
    set_error("internal: forward goes through launch_fwd_slab2");
    return UNIT_EINVAL;
  } else {
    UNIT_CUDA(cudaFuncSetAttribute(roi_align_bwd_slab<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_align_bwd_slab<T><<<N * (C / CS), NTHREADS, smem, st>>>(p);
    UNIT_CHECK_LAUNCH("roi_align_bwd_slab");
  }
  return UNIT_OK;
}

static bool slab_ok(int C, int H, int W, int PH, int PW, int rois_sorted, size_t stage_bytes) {
  if (!rois_sorted || PH != P || PW != P || (C % CS) != 0) return false;
  return smem_bytes_total(plane_stride_host(H * W), stage_bytes) <= 227 * 1024;
}

bool fwd_slab2_fits(int C, int H, int W, int dtype);
int launch_fwd_slab2(const void* feat, const float* rois, void* out, int N, int C, int H, int W, int R, float scale,
                     int sr, int aligned, int dtype, const int* img_off, cudaStream_t st);

}  // namespace roi
}  // namespace unit

using namespace unit;
using namespace unit::roi;

extern "C" {

size_t unit_roi_align_workspace_bytes(int N) { return ((size_t)(N + 2) * sizeof(int) + 255) / 256 * 256; }

int unit_roi_align_fwd(const void* feat, const float* rois, void* out, int N, int C, int H, int W, int R, int PH,
                       int PW, float spatial_scale, int sampling_ratio, int aligned, int dtype, int rois_sorted,
                       void* workspace, size_t workspace_bytes, unit_stream_t stream) {
  UNIT_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0 && R >= 0 && PH > 0 && PW > 0, "roi_align_fwd: bad shape");
  UNIT_REQUIRE(dtype == UNIT_F32 || dtype == UNIT_BF16, "roi_align_fwd: dtype must be f32 or bf16");
  if (R == 0) return UNIT_OK;
  UNIT_REQUIRE(feat && rois && out, "roi_align_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (rois_sorted && PH == P && PW == P && N > 0 && fwd_slab2_fits(C, H, W, dtype)) {
    if (!workspace || workspace_bytes < unit_roi_align_workspace_bytes(N)) {
      set_error("roi_align_fwd: workspace too small (%zu < %zu)", workspace_bytes, unit_roi_align_workspace_bytes(N));
      return UNIT_EWORKSPACE;
    }
    UNIT_REQUIRE((((uintptr_t)out) & 15) == 0, "roi_align_fwd: out must be 16-byte aligned");
    roi_offsets_kernel<<<cdiv(R + 1, 256), 256, 0, st>>>(rois, R, N, (int*)workspace);
    UNIT_CHECK_LAUNCH("roi_offsets_kernel");
    return launch_fwd_slab2(feat, rois, out, N, C, H, W, R, spatial_scale, sampling_ratio, aligned, dtype,
                            (const int*)workspace, st);
  }
  const long long total = (long long)R * C * PH * PW;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 16);
  if (dtype == UNIT_F32)
    roi_align_fwd_generic<float><<<grid, 256, 0, st>>>((const float*)feat, rois, (float*)out, N, C, H, W, total, PH,
                                                        PW, spatial_scale, sampling_ratio, aligned);
  else
    roi_align_fwd_generic<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)feat, rois, (__nv_bfloat16*)out,
                                                                N, C, H, W, total, PH, PW, spatial_scale,
                                                                sampling_ratio, aligned);
  UNIT_CHECK_LAUNCH("roi_align_fwd_generic");
  return UNIT_OK;
}

int unit_roi_align_bwd(const void* grad_out, const float* rois, void* grad_feat, int N, int C, int H, int W, int R,
                       int PH, int PW, float spatial_scale, int sampling_ratio, int aligned, int dtype,
                       int rois_sorted, void* workspace, size_t workspace_bytes, unit_stream_t stream) {
  UNIT_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0 && R >= 0 && PH > 0 && PW > 0, "roi_align_bwd: bad shape");
  UNIT_REQUIRE(dtype == UNIT_F32 || dtype == UNIT_BF16, "roi_align_bwd: dtype must be f32 or bf16");
  if (N == 0) return UNIT_OK;
  UNIT_REQUIRE(grad_feat && (R == 0 || (grad_out && rois)), "roi_align_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esz = dtype == UNIT_F32 ? 4 : 2;
  if (R == 0) {
    UNIT_CUDA(cudaMemsetAsync(grad_feat, 0, (size_t)N * C * H * W * esz, st));
    return UNIT_OK;
  }
  const size_t stage = dtype == UNIT_F32 ? smem_stage_bytes<float>() : smem_stage_bytes<__nv_bfloat16>();
  if (slab_ok(C, H, W, PH, PW, rois_sorted, stage) && (((uintptr_t)grad_out) & 15) == 0) {
    if (!workspace || workspace_bytes < unit_roi_align_workspace_bytes(N)) {
      set_error("roi_align_bwd: workspace too small");
      return UNIT_EWORKSPACE;
    }
    if (dtype == UNIT_F32)
      return launch_slab<float>(true, grad_out, rois, grad_feat, N, C, H, W, R, spatial_scale, sampling_ratio,
                                aligned, (int*)workspace, st);
    return launch_slab<__nv_bfloat16>(true, grad_out, rois, grad_feat, N, C, H, W, R, spatial_scale, sampling_ratio,
                                      aligned, (int*)workspace, st);
  }
  UNIT_REQUIRE(dtype == UNIT_F32, "roi_align_bwd: the generic (unsorted / non-14x14) path supports f32 only");
  UNIT_CUDA(cudaMemsetAsync(grad_feat, 0, (size_t)N * C * H * W * 4, st));
  const long long total = (long long)R * C * PH * PW;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 16);
  roi_align_bwd_generic<float><<<grid, 256, 0, st>>>((const float*)grad_out, rois, (float*)grad_feat, N, C, H, W,
                                                      total, PH, PW, spatial_scale, sampling_ratio, aligned);
  UNIT_CHECK_LAUNCH("roi_align_bwd_generic");
  return UNIT_OK;
}

}  // extern "C"
