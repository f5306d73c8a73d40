// ROIAlign C-ABI entry points + the generic (any pooled size / channel count / RoI order) kernels.
//
// Replaces [D2] ROIPooler -> ROIAlign(aligned=True) -> [TV] roi_align / _roi_align_backward
// (reference call sites: modeling/roi_heads/roi_heads.py:356,364,499,511,598,610,708,715,729,829,843,911).
// The hot path (14x14, RoIs grouped by image, C % 8 == 0, slab fits in shared memory) is the slab-resident pair of
// kernels in roi_align_fwd.cu / roi_align_bwd.cu; everything else falls to the one-thread-per-output kernels here,
// which follow the torchvision arithmetic literally.
//
// Algorithmic bytes per image (fp32): C*H*W*4 + R*20 + R*C*196*4  (SURVEY.md section 8d), forward and backward.
#include <stdlib.h>

#include <algorithm>

#include "roi_common.cuh"

namespace unit {
namespace roi {

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

// per-image RoI offsets from the (sorted) batch-index column: off[n] = first r with batch_idx >= n
__global__ void roi_offsets_kernel(const float* __restrict__ rois, int R, int N, int* __restrict__ off) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > R) return;
  const int cur = r < R ? min(max((int)rois[(long long)r * 5], 0), N) : N;
  const int prev = r > 0 ? min(max((int)rois[(long long)(r - 1) * 5], 0), N) : -1;
  for (int n = prev + 1; n <= cur; ++n) off[n] = r;
}

// [D2] convert_boxes_to_pooler_format: rois[r] = (image of r, x1, y1, x2, y2) for boxes concatenated in image order
__global__ void boxes_to_rois_kernel(const float4* __restrict__ boxes, const int* __restrict__ off, int n_img, int R,
                                     float* __restrict__ rois) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int lo = 0, hi = n_img;  // largest i with off[i] <= r
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= r) lo = mid; else hi = mid;
  }
  const float4 b = boxes[r];
  float* o = rois + (long long)r * 5;
  o[0] = (float)lo;
  o[1] = b.x;
  o[2] = b.y;
  o[3] = b.z;
  o[4] = b.w;
}

bool fwd_slab2_fits(int C, int H, int W, int dtype);
int launch_fwd_slab2(const void* feat, const float* rois, void* out, int N, int C, int H, int W, int R, float scale,
                     int sr, int aligned, int dtype, const int* img_off, cudaStream_t st);
bool fwd_band_fits(int C, int H, int W, int dtype);
size_t fwd_band_workspace_bytes(int R);
int launch_fwd_band(const void* feat, const float* rois, void* out, void* tabs_ws, int N, int C, int H, int W, int R,
                    float scale, int sr, int aligned, int dtype, int* img_off, cudaStream_t st);
bool bwd_slab2_fits(int C, int H, int W, int dtype);
size_t bwd_slab2_workspace_bytes(int N, int C, int H, int W, int dtype);
int launch_bwd_slab2(const void* gout, const float* rois, void* gfeat, void* f32_scratch, int N, int C, int H, int W,
                     int R, float scale, int sr, int aligned, int dtype, const int* img_off, cudaStream_t st);

bool bwd_cl_fits(int C, int H, int W, int R, int dtype, const void* gout);
int launch_bwd_cl(const void* gout, const float* rois, void* gfeat, void* ws, int N, int C, int H, int W, int R,
                  float scale, int sr, int aligned, int dtype, cudaStream_t st);

size_t bwd_cl2_table_bytes(int R);
int launch_bwd_cl2(const void* gout, const float* rois, void* gfeat, void* ws, int N, int C, int H, int W, int R,
                   float scale, int sr, int aligned, cudaStream_t st);

static size_t offsets_bytes(int N) { return ((size_t)(N + 2) * sizeof(int) + 255) / 256 * 256; }

}  // namespace roi
}  // namespace unit

using namespace unit;
using namespace unit::roi;

extern "C" {

size_t unit_roi_align_workspace_bytes(int N, int C, int H, int W, int R, int dtype) {
  const size_t fwd = fwd_band_workspace_bytes(R);
  const size_t bwd = bwd_slab2_workspace_bytes(N, C, H, W, dtype) + 512 + bwd_cl2_table_bytes(R);
  return offsets_bytes(N) + (fwd > bwd ? fwd : bwd) + 256;
}

int unit_roi_align_fwd(const void* feat, const float* rois, void* out, int N, int C, int H, int W, int R, int PH,
                       int PW, float spatial_scale, int sampling_ratio, int aligned, int dtype, int rois_sorted,
                       void* workspace, size_t workspace_bytes, unit_stream_t stream) {
  UNIT_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0 && R >= 0 && PH > 0 && PW > 0, "roi_align_fwd: bad shape");
  UNIT_REQUIRE(dtype == UNIT_F32 || dtype == UNIT_BF16, "roi_align_fwd: dtype must be f32 or bf16");
  if (R == 0) return UNIT_OK;
  UNIT_REQUIRE(feat && rois && out, "roi_align_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (rois_sorted && PH == 14 && PW == 14 && N > 0 && fwd_band_fits(C, H, W, dtype) &&
      !switches().fwd_v3) {  // fp32 and bf16 I/O take the same band-gather kernel (the slab is fp32 either way)
    const size_t need = offsets_bytes(N) + fwd_band_workspace_bytes(R);
    if (!workspace || workspace_bytes < need) {
      set_error("roi_align_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
      return UNIT_EWORKSPACE;
    }
    UNIT_REQUIRE((((uintptr_t)out) & 15) == 0, "roi_align_fwd: out must be 16-byte aligned");
    // the per-image offsets are written by the table kernel (first words of the workspace)
    return launch_fwd_band(feat, rois, out, (char*)workspace + offsets_bytes(N), N, C, H, W, R, spatial_scale,
                           sampling_ratio, aligned, dtype, (int*)workspace, st);
  }
  if (rois_sorted && PH == 14 && PW == 14 && N > 0 && fwd_slab2_fits(C, H, W, dtype)) {
    if (!workspace || workspace_bytes < offsets_bytes(N)) {
      set_error("roi_align_fwd: workspace too small (%zu < %zu)", workspace_bytes, offsets_bytes(N));
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

int unit_boxes_to_rois(const float* boxes, const int* offsets, int n_img, int R, float* rois, unit_stream_t stream) {
  UNIT_REQUIRE(n_img >= 0 && R >= 0, "boxes_to_rois: bad shape");
  if (R == 0) return UNIT_OK;
  UNIT_REQUIRE(boxes && offsets && rois && n_img > 0, "boxes_to_rois: null pointer");
  UNIT_REQUIRE(((uintptr_t)boxes & 15) == 0, "boxes_to_rois: boxes must be 16-byte aligned");
  boxes_to_rois_kernel<<<cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)boxes, offsets, n_img, R, rois);
  UNIT_CHECK_LAUNCH("boxes_to_rois_kernel");
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
  if (PH == 14 && PW == 14 && bwd_cl_fits(C, H, W, R, dtype, grad_out) && !switches().bwd_v4) {
    const size_t need = offsets_bytes(N) + bwd_slab2_workspace_bytes(N, C, H, W, dtype) + 512 + bwd_cl2_table_bytes(R);
    if (!workspace || workspace_bytes < need) {
      set_error("roi_align_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
      return UNIT_EWORKSPACE;
    }
    if (dtype == UNIT_F32 && !switches().bwd_cl1)  // table-driven channel-lane kernel (fp32 I/O)
      return launch_bwd_cl2(grad_out, rois, grad_feat, (char*)workspace + offsets_bytes(N), N, C, H, W, R,
                            spatial_scale, sampling_ratio, aligned, st);
    return launch_bwd_cl(grad_out, rois, grad_feat, (char*)workspace + offsets_bytes(N), N, C, H, W, R, spatial_scale,
                         sampling_ratio, aligned, dtype, st);
  }
  if (rois_sorted && PH == 14 && PW == 14 && bwd_slab2_fits(C, H, W, dtype) && (((uintptr_t)grad_out) & 15) == 0) {
    const size_t need = offsets_bytes(N) + bwd_slab2_workspace_bytes(N, C, H, W, dtype);
    if (!workspace || workspace_bytes < need) {
      set_error("roi_align_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
      return UNIT_EWORKSPACE;
    }
    roi_offsets_kernel<<<cdiv(R + 1, 256), 256, 0, st>>>(rois, R, N, (int*)workspace);
    UNIT_CHECK_LAUNCH("roi_offsets_kernel");
    return launch_bwd_slab2(grad_out, rois, grad_feat, (char*)workspace + offsets_bytes(N), N, C, H, W, R,
                            spatial_scale, sampling_ratio, aligned, dtype, (const int*)workspace, st);
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
