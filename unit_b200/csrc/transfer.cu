// Base -> novel knowledge transfer.
//
//   lingual similarity     fast_rcnn.py:376-382 (+ softmax of roi_heads.py:272)
//   visual similarity      roi_heads.py:246-257
//   combination            roi_heads.py:266-322 ('Sum' mode; class-level terms arrive pre-reduced in static_*)
//   transfer               fast_rcnn.py:403-426 / 503-528, weak scores :360-368, fine-tune terms :527-528
//   mask transfer          mask_head.py:16-37 / 72-94 + [D2] mask_rcnn_inference
//   mask paste             [D2] paste_masks_in_image (meta_arch/rcnn.py:423 via detector_postprocess)
//
// One 128-thread CTA per RoI fuses what the reference does with ~30 ATen launches (softmax, index_select,
// renormalise, threshold, broadcast add, renormalise, 2 x bmm with 5x15 / 20x60 matrices, index_copy, adds).
// The dense predictor GEMMs feeding this epilogue are K=2048 contractions and run on tensor cores (gemm.cu /
// cuBLAS); this file is the memory-bound part and never leaves fp32.
#include <math.h>

#include <stdlib.h>

#include "common.cuh"

namespace unit {
namespace transfer {

constexpr int NOVEL_TAG = 1000000;

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------- lingual similarity
__global__ void lingual_kernel(const float* __restrict__ emb, const int64_t* __restrict__ indexer,
                               const int64_t* __restrict__ base, const int64_t* __restrict__ novel, int D, int B,
                               float* __restrict__ raw, float* __restrict__ soft) {
  extern __shared__ float s_raw[];  // [B]
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float* en = emb + indexer[novel[n]] * D;
  for (int b = warp; b < B; b += nwarp) {
    const float* eb = emb + indexer[base[b]] * D;
    // fp64 accumulation: exp() in the softmax amplifies the rounding of these |L| ~ 30 dot products to ~1e-5 relative
    double acc = 0.0;
    for (int d = lane; d < D; d += 32) acc += (double)en[d] * (double)eb[d];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_raw[b] = (float)acc;
  }
  __syncthreads();
  if (warp == 0) {
    float m = -INFINITY;
    for (int b = lane; b < B; b += 32) m = fmaxf(m, s_raw[b]);
    m = warp_max(m);
    float sum = 0.f;
    for (int b = lane; b < B; b += 32) sum += expf(s_raw[b] - m);
    sum = warp_sum(sum);
    for (int b = lane; b < B; b += 32) {
      if (raw) raw[n * B + b] = s_raw[b];
      if (soft) soft[n * B + b] = expf(s_raw[b] - m) / sum;
    }
  }
}

// ---------------------------------------------------------------------------------------- fused transfer
constexpr int TT = 128;

struct TransferArgs {
  unit_transfer_params p;
  const float* vis_logits;
  const float* stat[3];  // cls, bbox, seg
  const int* base;
  const int* novel;
  const int* class_kind;
  const float* delta_scores;
  const float* proposal_deltas;
  const float* weak_scores;
  const float* ft_scores;
  const float* ft_deltas;
  float* out_scores;
  float* out_bbox;
  float* out_s[3];
};

__global__ void __launch_bounds__(TT) similarity_transfer_kernel(const TransferArgs a) {
  extern __shared__ float sm[];
  const int K = a.p.K, B = a.p.B, Nn = a.p.Nn, K1 = K + 1;
  float* s_v = sm;              // [B]   visual similarity
  float* s_delta = s_v + B;     // [K1]
  float* s_pd = s_delta + K1;   // [4K]
  float* s_tc = s_pd + 4 * K;   // [Nn]  transferred class logits
  float* s_tb = s_tc + Nn;      // [4Nn] transferred box deltas
  __shared__ float s_red[2];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = TT / 32;

  // row strides of the (possibly packed) GEMM outputs; 0 = dense
  const long long ld_d = a.p.ld_delta_scores ? a.p.ld_delta_scores : K1;
  const long long ld_p = a.p.ld_proposal_deltas ? a.p.ld_proposal_deltas : 4 * K;
  const long long ld_fs = a.p.ld_ft_scores ? a.p.ld_ft_scores : K1;
  const long long ld_fd = a.p.ld_ft_deltas ? a.p.ld_ft_deltas : 4 * K;
  for (int k = tid; k < K1; k += TT) s_delta[k] = a.delta_scores[r * ld_d + k];
  for (int k = tid; k < 4 * K; k += TT) s_pd[k] = a.proposal_deltas[r * ld_p + k];

  const bool need_v = a.p.do_transfer && (a.p.wv_cls != 0.f || a.p.wv_bbox != 0.f || a.p.wv_seg != 0.f);
  if (need_v) {
    // softmax over all K+1 OICR logits, base columns, renormalise, threshold (roi_heads.py:255-257)
    const float* vl = a.vis_logits + (long long)r * (a.p.ld_vis_logits ? a.p.ld_vis_logits : K1);
    if (warp == 0) {
      float m = -INFINITY;
      for (int k = lane; k < K1; k += 32) m = fmaxf(m, vl[k]);
      m = warp_max(m);
      float sum = 0.f;
      for (int k = lane; k < K1; k += 32) sum += expf(vl[k] - m);
      sum = warp_sum(sum);
      float bs = 0.f;
      for (int b = lane; b < B; b += 32) {
        const float pv = expf(vl[a.base[b]] - m) / sum;
        s_v[b] = pv;
        bs += pv;
      }
      bs = warp_sum(bs);
      const float den = fmaxf(bs, 1e-9f);
      for (int b = lane; b < B; b += 32) {
        float v = s_v[b] / den;
        if (v < a.p.vis_threshold) v = 0.f;
        s_v[b] = v;
      }
    }
  } else {
    for (int b = tid; b < B; b += TT) s_v[b] = 0.f;
  }
  __syncthreads();
  (void)s_red;

  if (a.p.do_transfer) {
    const float wv[3] = {a.p.wv_cls, a.p.wv_bbox, a.p.wv_seg};
    const int nrm[3] = {a.p.norm_cls, a.p.norm_bbox, a.p.norm_seg};
    for (int h = 0; h < 3; ++h) {
      if (h == 2 && !a.out_s[2]) continue;  // the seg similarity is only materialised for the mask head
      for (int n = warp; n < Nn; n += nwarp) {
        const float* st = a.stat[h] ? a.stat[h] + (((a.p.static_per_roi >> h) & 1) ? (long long)r * Nn * B : 0) + n * B : nullptr;
        float sum = 0.f;
        for (int b = lane; b < B; b += 32) sum += (st ? st[b] : 0.f) + wv[h] * s_v[b];
        sum = warp_sum(sum);
        const float inv_den = nrm[h] ? 1.f / fmaxf(sum, 1e-9f) : 1.f;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        for (int b = lane; b < B; b += 32) {
          float s = (st ? st[b] : 0.f) + wv[h] * s_v[b];
          s = nrm[h] ? s / fmaxf(sum, 1e-9f) : s;
          if (a.out_s[h]) a.out_s[h][((long long)r * Nn + n) * B + b] = s;
          const int kb = a.base[b];
          if (h == 0) {
            acc0 = fmaf(s, s_delta[kb], acc0);
          } else if (h == 1) {
            acc0 = fmaf(s, s_pd[4 * kb + 0], acc0);
            acc1 = fmaf(s, s_pd[4 * kb + 1], acc1);
            acc2 = fmaf(s, s_pd[4 * kb + 2], acc2);
            acc3 = fmaf(s, s_pd[4 * kb + 3], acc3);
          }
        }
        (void)inv_den;
        if (h == 0) {
          acc0 = warp_sum(acc0);
          if (lane == 0) s_tc[n] = acc0;
        } else if (h == 1) {
          acc0 = warp_sum(acc0);
          acc1 = warp_sum(acc1);
          acc2 = warp_sum(acc2);
          acc3 = warp_sum(acc3);
          if (lane == 0) {
            s_tb[4 * n + 0] = acc0;
            s_tb[4 * n + 1] = acc1;
            s_tb[4 * n + 2] = acc2;
            s_tb[4 * n + 3] = acc3;
          }
        }
      }
    }
  }
  __syncthreads();

  for (int k = tid; k < K1; k += TT) {
    float v = s_delta[k];
    const int kind = k < K ? a.class_kind[k] : -1;
    if (a.p.do_transfer && kind >= NOVEL_TAG) v = v + s_tc[kind - NOVEL_TAG];
    if (a.weak_scores) v = v + a.weak_scores[(long long)r * (a.p.ld_weak_scores ? a.p.ld_weak_scores : K1) + k];
    if (a.ft_scores) v = v + a.ft_scores[r * ld_fs + k];
    if (a.p.novel_neg_inf && kind >= NOVEL_TAG) v = -INFINITY;
    a.out_scores[(long long)r * K1 + k] = v;
  }
  for (int i = tid; i < 4 * K; i += TT) {
    const int k = i >> 2, j = i & 3;
    float v = s_pd[i];
    if (a.p.do_transfer) {
      const int kind = a.class_kind[k];
      if (kind >= NOVEL_TAG) v = s_tb[4 * (kind - NOVEL_TAG) + j];
      else if (kind < 0) v = 0.f;
    }
    if (a.ft_deltas) v = v + a.ft_deltas[r * ld_fd + i];
    a.out_bbox[(long long)r * 4 * K + i] = v;
  }
}

struct TransferBwdArgs {
  int R, K, B, Nn;
  const float* s_cls;
  const float* s_bbox;
  const int* base;
  const int* novel;
  const int* class_kind;
  const float* g_scores;
  const float* g_bbox;
  int detach;
  float* g_delta;
  float* g_pd;
};

__global__ void __launch_bounds__(TT) similarity_transfer_bwd_kernel(const TransferBwdArgs a) {
  const int K = a.K, B = a.B, Nn = a.Nn, K1 = K + 1;
  const int r = blockIdx.x, tid = threadIdx.x;
  // scores: identity everywhere; base columns additionally collect sum_n S_cls[n,b] * g[novel n]
  for (int k = tid; k < K1; k += TT) {
    float g = a.g_scores[(long long)r * K1 + k];
    const int kind = k < K ? a.class_kind[k] : -1;
    if (!a.detach && kind >= 0 && kind < NOVEL_TAG) {
      const int b = kind;
      for (int n = 0; n < Nn; ++n)
        g = fmaf(a.s_cls[((long long)r * Nn + n) * B + b], a.g_scores[(long long)r * K1 + a.novel[n]], g);
    }
    a.g_delta[(long long)r * K1 + k] = g;
  }
  for (int i = tid; i < 4 * K; i += TT) {
    const int k = i >> 2, j = i & 3;
    const int kind = a.class_kind[k];
    float g = 0.f;
    if (kind >= 0 && kind < NOVEL_TAG) {
      g = a.g_bbox[(long long)r * 4 * K + i];
      if (!a.detach)
        for (int n = 0; n < Nn; ++n)
          g = fmaf(a.s_bbox[((long long)r * Nn + n) * B + kind], a.g_bbox[(long long)r * 4 * K + 4 * a.novel[n] + j], g);
    }
    a.g_pd[(long long)r * 4 * K + i] = g;
  }
}


// Gradient of the fused similarity + transfer w.r.t. the visual logits (roi_heads.py:245-257 is NOT under no_grad in
// the reference: with a trainable box head the fine-tune loss reaches box_features through softmax -> renormalise ->
// threshold -> S -> bmm).  One CTA per RoI recomputes the forward pieces (p, u, t, S) and applies the chain rule:
//   gS_cls[n,b]  = g_scores[novel n] * delta[base b]          gS_bbox[n,b] = sum_j g_bbox[novel n, j] * pd[base b, j]
//   gt = (gS - <gS, S>) / sum(t)      (normalised heads; gt = gS otherwise)
//   gv[b] = sum_h wv_h sum_n gt_h[n,b];   thresholded entries pass no gradient;   u = pb / sum(pb);   p = softmax(vl)
struct TransferBwdVisArgs {
  unit_transfer_params p;
  const float* vis_logits;
  const float* stat[2];
  const int* base;
  const int* novel;
  const float* delta_scores;
  const float* proposal_deltas;
  const float* g_scores;
  const float* g_bbox;
  float* g_vis;
};

__global__ void __launch_bounds__(TT) similarity_transfer_bwd_vis_kernel(const TransferBwdVisArgs a) {
  extern __shared__ float sm[];
  const int K = a.p.K, B = a.p.B, Nn = a.p.Nn, K1 = K + 1;
  float* s_p = sm;            // [K1] softmax of the visual logits
  float* s_u = s_p + K1;      // [B]  renormalised base probabilities (before the threshold)
  float* s_gv = s_u + B;      // [B]  gradient w.r.t. the thresholded visual similarity
  __shared__ float s_scal[2];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = TT / 32;
  const long long ld_d = a.p.ld_delta_scores ? a.p.ld_delta_scores : K1;
  const long long ld_p = a.p.ld_proposal_deltas ? a.p.ld_proposal_deltas : 4 * K;
  const float* vl = a.vis_logits + (long long)r * (a.p.ld_vis_logits ? a.p.ld_vis_logits : K1);
  for (int b = tid; b < B; b += TT) s_gv[b] = 0.f;
  if (warp == 0) {
    float m = -INFINITY;
    for (int k = lane; k < K1; k += 32) m = fmaxf(m, vl[k]);
    m = warp_max(m);
    float sum = 0.f;
    for (int k = lane; k < K1; k += 32) sum += expf(vl[k] - m);
    sum = warp_sum(sum);
    for (int k = lane; k < K1; k += 32) s_p[k] = expf(vl[k] - m) / sum;
    __syncwarp();
    float bs = 0.f;
    for (int b = lane; b < B; b += 32) bs += s_p[a.base[b]];
    bs = warp_sum(bs);
    const float den = fmaxf(bs, 1e-9f);
    for (int b = lane; b < B; b += 32) s_u[b] = s_p[a.base[b]] / den;
    if (lane == 0) {
      s_scal[0] = den;
      s_scal[1] = bs >= 1e-9f ? 1.f : 0.f;  // the clamp is inactive: the renormalisation has a gradient of its own
    }
  }
  __syncthreads();
  const float wv[2] = {a.p.wv_cls, a.p.wv_bbox};
  const int nrm[2] = {a.p.norm_cls, a.p.norm_bbox};
  const float thr = a.p.vis_threshold;
  for (int h = 0; h < 2; ++h) {
    if (wv[h] == 0.f) continue;
    for (int n = warp; n < Nn; n += nwarp) {
      const float* st = a.stat[h] ? a.stat[h] + (((a.p.static_per_roi >> h) & 1) ? (long long)r * Nn * B : 0) + n * B : nullptr;
      const int kn = a.novel[n];
      float g4[4] = {0.f, 0.f, 0.f, 0.f};
      if (h == 0) {
        g4[0] = a.g_scores[(long long)r * K1 + kn];
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) g4[j] = a.g_bbox[(long long)r * 4 * K + 4 * kn + j];
      }
      float sum = 0.f, dot = 0.f;
      for (int b = lane; b < B; b += 32) {
        const float v = s_u[b] < thr ? 0.f : s_u[b];
        const float t = (st ? st[b] : 0.f) + wv[h] * v;
        const int kb = a.base[b];
        float gs;
        if (h == 0) {
          gs = g4[0] * a.delta_scores[r * ld_d + kb];
        } else {
          const float* pd = a.proposal_deltas + r * ld_p + 4 * kb;
          gs = g4[0] * pd[0] + g4[1] * pd[1] + g4[2] * pd[2] + g4[3] * pd[3];
        }
        sum += t;
        dot += gs * t;
      }
      sum = warp_sum(sum);
      dot = warp_sum(dot);
      const float den = fmaxf(sum, 1e-9f);
      const float mean = (nrm[h] && sum >= 1e-9f) ? dot / den : 0.f;  // <gS, S>
      for (int b = lane; b < B; b += 32) {
        const int kb = a.base[b];
        float gs;
        if (h == 0) {
          gs = g4[0] * a.delta_scores[r * ld_d + kb];
        } else {
          const float* pd = a.proposal_deltas + r * ld_p + 4 * kb;
          gs = g4[0] * pd[0] + g4[1] * pd[1] + g4[2] * pd[2] + g4[3] * pd[3];
        }
        const float gt = nrm[h] ? (gs - mean) / den : gs;
        atomicAdd(&s_gv[b], wv[h] * gt);  // <= Nn adds per entry, shared memory
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    const float den = s_scal[0], live = s_scal[1];
    float dot = 0.f;
    for (int b = lane; b < B; b += 32) {
      const float gu = s_u[b] < thr ? 0.f : s_gv[b];
      s_gv[b] = gu;
      dot += gu * s_u[b];
    }
    dot = warp_sum(dot) * live;
    // gradient w.r.t. the softmax probabilities: non-zero on base columns only
    float pdot = 0.f;
    for (int b = lane; b < B; b += 32) {
      const float gpb = (s_gv[b] - dot) / den;
      s_gv[b] = gpb;
      pdot += gpb * s_p[a.base[b]];
    }
    pdot = warp_sum(pdot);
    __syncwarp();
    for (int k = lane; k < K1; k += 32) a.g_vis[(long long)r * K1 + k] = -s_p[k] * pdot;
    __syncwarp();
    for (int b = lane; b < B; b += 32) {
      const int kb = a.base[b];
      a.g_vis[(long long)r * K1 + kb] = s_p[kb] * (s_gv[b] - pdot);
    }
  }
}

// ---------------------------------------------------------------------------------------- mask transfer
__global__ void mask_transfer_kernel(const float* __restrict__ logits, const float* __restrict__ s_seg, int s_is_2d,
                                     const int* __restrict__ base, const int* __restrict__ class_kind,
                                     const float* __restrict__ x_delta, const int64_t* __restrict__ pred_classes,
                                     float* __restrict__ out_logits, float* __restrict__ out_probs, int K, int B,
                                     int Nn, int MM) {
  // grid = (detections, chunks): the chunks of one detection split its pixels (and, for the full-logits output, its
  // (class, pixel) pairs), so that 100 detections fill the GPU instead of 100 SMs running 60-long dependent chains
  const int d = blockIdx.x;
  const int nchunk = gridDim.y, chunk = blockIdx.y;
  const float* lg = logits + (long long)d * K * MM;
  const float* S = s_seg ? s_seg + (s_is_2d ? 0 : (long long)d * Nn * B) : nullptr;
  // base-class combination of one pixel: four independent accumulators (the loads of a round are all in flight)
  auto combine = [&](const float* sr, int m) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int b = 0;
    for (; b + 4 <= B; b += 4) {
      a0 = fmaf(sr[b], lg[base[b] * MM + m], a0);
      a1 = fmaf(sr[b + 1], lg[base[b + 1] * MM + m], a1);
      a2 = fmaf(sr[b + 2], lg[base[b + 2] * MM + m], a2);
      a3 = fmaf(sr[b + 3], lg[base[b + 3] * MM + m], a3);
    }
    for (; b < B; ++b) a0 = fmaf(sr[b], lg[base[b] * MM + m], a0);
    return (a0 + a1) + (a2 + a3);
  };
  if (out_logits) {
    for (int i = chunk * blockDim.x + threadIdx.x; i < K * MM; i += nchunk * blockDim.x) {
      const int k = i / MM, m = i - k * MM;
      float v = lg[i];
      if (S) {
        const int kind = class_kind[k];
        if (kind >= NOVEL_TAG) {
          const float* sr = S + (kind - NOVEL_TAG) * B;
          float acc = 0.f;
          for (int b = 0; b < B; ++b) acc = fmaf(sr[b], lg[base[b] * MM + m], acc);  // reference summation order
          v = acc;
        } else if (kind < 0) {
          v = 0.f;
        }
      }
      if (x_delta) v = v + x_delta[(long long)d * K * MM + i];
      out_logits[(long long)d * K * MM + i] = v;
    }
  }
  if (out_probs) {
    const int k = (int)pred_classes[d];
    const int kind = class_kind[k];
    for (int m = chunk * blockDim.x + threadIdx.x; m < MM; m += nchunk * blockDim.x) {
      float v = lg[k * MM + m];
      if (S) {
        if (kind >= NOVEL_TAG) {
          v = combine(S + (kind - NOVEL_TAG) * B, m);
        } else if (kind < 0) {
          v = 0.f;
        }
      }
      if (x_delta) v = v + x_delta[((long long)d * K + k) * MM + m];
      out_probs[(long long)d * MM + m] = 1.f / (1.f + expf(-v));
    }
  }
}

// ---------------------------------------------------------------------------------------- mask paste
// out[d,y,x] = bilinear(mask_d, grid(x,y)) >= thr with F.grid_sample(align_corners=False, zeros) semantics.
// Thread = 16 consecutive output bytes (one 16-byte store); runs entirely outside the box window are zero-filled
// without sampling (about 95 % of an 800x1333 canvas).
__device__ __forceinline__ float paste_sample(const float* __restrict__ mk, int M, float gx, float gy) {
  const float half = (float)M * 0.5f;
  const float ix = __fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), half), 0.5f);
  const float iy = __fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), half), 0.5f);
  const float xw = floorf(ix), yn = floorf(iy);
  const float w = __fsub_rn(ix, xw), e = __fsub_rn(1.f, w);
  const float n = __fsub_rn(iy, yn), s = __fsub_rn(1.f, n);
  const int x0 = (int)xw, y0 = (int)yn;
  auto at = [&](int y, int x) -> float { return (x >= 0 && x < M && y >= 0 && y < M) ? mk[y * M + x] : 0.f; };
  const float nw = __fmul_rn(at(y0, x0), __fmul_rn(s, e));
  const float ne = __fmul_rn(at(y0, x0 + 1), __fmul_rn(s, w));
  const float sw = __fmul_rn(at(y0 + 1, x0), __fmul_rn(n, e));
  const float se = __fmul_rn(at(y0 + 1, x0 + 1), __fmul_rn(n, w));
  return __fadd_rn(__fadd_rn(__fadd_rn(nw, ne), sw), se);
}

// PX pixels per thread, one PX-byte store.  4 (not 16): a warp then spans 128 pixels of a row, so far fewer warps
// straddle a box window and drag all their lanes through the sampling path.
#ifndef UNIT_PASTE_PX
#define UNIT_PASTE_PX 4
#endif
constexpr int PX = UNIT_PASTE_PX;
static_assert(PX == 4 || PX == 16, "mask_paste: 4 or 16 pixels per thread");
__device__ __forceinline__ void paste_store(uint8_t* dst, const unsigned char* v) {
  if (PX == 16) {
    uint4 q;
    memcpy(&q, v, 16);
    *reinterpret_cast<uint4*>(dst) = q;
  } else {
    uint32_t q;
    memcpy(&q, v, 4);
    *reinterpret_cast<uint32_t*>(dst) = q;
  }
}
__global__ void mask_paste_kernel(const float* __restrict__ masks, const float4* __restrict__ boxes, int D, int M,
                                  int H, int W, float thr, uint8_t* __restrict__ out) {
  const long long total = (long long)D * H * W;
  const long long start = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * PX;
  if (start >= total) return;
  const long long hw = (long long)H * W;
  int d, y, x;
  if (total < (1ll << 31)) {  // 32-bit index arithmetic (64-bit divisions cost more than the rest of the thread)
    const unsigned s32 = (unsigned)start, hw32 = (unsigned)hw;
    d = (int)(s32 / hw32);
    const unsigned rem = s32 - (unsigned)d * hw32;
    y = (int)(rem / (unsigned)W);
    x = (int)(rem - (unsigned)y * (unsigned)W);
  } else {
    d = (int)(start / hw);
    const long long rem = start - (long long)d * hw;
    y = (int)(rem / W);
    x = (int)(rem - (long long)y * W);
  }
  unsigned char v[PX];
  float4 bx = __ldg(boxes + d);
  {
    // Fast reject of a whole PX-pixel run (most of the image lies outside the box): a pixel can only be sampled when
    // its centre is within bw / M resp. bh / M of the box; one extra pixel of margin covers every rounding of the
    // exact expressions below, which still decide every pixel that is not rejected here.
    const float bw = bx.z - bx.x, bh = bx.w - bx.y;
    if (x + PX - 1 < W && start + PX <= total && bw > 0.f && bh > 0.f && bw < 1e6f && bh < 1e6f &&
        fabsf(bx.x) < 1e6f && fabsf(bx.y) < 1e6f && ((((uintptr_t)out) + start) & (PX - 1)) == 0) {
      const float mx = bw / (float)M + 1.f, my = bh / (float)M + 1.f;
      const float py = (float)y + 0.5f;
      if (py < bx.y - my || py > bx.w + my || (float)(x + PX - 1) + 0.5f < bx.x - mx || (float)x + 0.5f > bx.z + mx) {
        const unsigned char zb = (0.f >= thr) ? 1 : 0;
#pragma unroll
        for (int i = 0; i < PX; ++i) v[i] = zb;
        paste_store(out + start, v);
        return;
      }
    }
  }
  for (int i = 0; i < PX; ++i) {
    unsigned char o = 0;
    if (start + i < total) {
      const float bw = __fsub_rn(bx.z, bx.x), bh = __fsub_rn(bx.w, bx.y);
      // quick reject: sample index must lie in (-1, M) on both axes
      const float px = (float)x + 0.5f, py = (float)y + 0.5f;
      const float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(px, bx.x), bw), 2.f), 1.f);
      const float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(py, bx.y), bh), 2.f), 1.f);
      const float lim = 1.f + 2.f / (float)M;
      if (fabsf(gx) < lim && fabsf(gy) < lim) {
        const float val = paste_sample(masks + (long long)d * M * M, M, gx, gy);
        o = val >= thr ? 1 : 0;
      } else if (!(gx == gx) || !(gy == gy)) {
        const float val = paste_sample(masks + (long long)d * M * M, M, gx, gy);
        o = val >= thr ? 1 : 0;
      } else {
        o = 0.f >= thr ? 1 : 0;
      }
      if (++x == W) {
        x = 0;
        if (++y == H) {
          y = 0;
          ++d;
          if (d < D) bx = __ldg(boxes + d);
        }
      }
    }
    v[i] = o;
  }
  if (start + PX <= total && ((((uintptr_t)out) + start) & (PX - 1)) == 0) {
    paste_store(out + start, v);
  } else {
    for (int i = 0; i < PX && start + i < total; ++i) out[start + i] = v[i];
  }
}

// One pixel, the exact expression chain shared by both paste kernels.
__device__ __forceinline__ unsigned char paste_pixel(const float* __restrict__ mk, int M, float4 bx, int x, int y,
                                                     float thr) {
  const float bw = __fsub_rn(bx.z, bx.x), bh = __fsub_rn(bx.w, bx.y);
  const float px = (float)x + 0.5f, py = (float)y + 0.5f;
  const float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(px, bx.x), bw), 2.f), 1.f);
  const float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(py, bx.y), bh), 2.f), 1.f);
  const float lim = 1.f + 2.f / (float)M;
  if ((fabsf(gx) < lim && fabsf(gy) < lim) || !(gx == gx) || !(gy == gy)) return paste_sample(mk, M, gx, gy) >= thr ? 1 : 0;
  return 0.f >= thr ? 1 : 0;
}

// Box-centric paste for thr > 0 (pixels outside the window are 0): the canvas is cleared by a memset and one CTA per
// (detection, strip of WIN_ROWS rows) visits only the pixels of its strip that lie in the box window; strips outside
// the window exit at once.  Irregular boxes (empty, non-finite, huge) get the whole canvas as their window.
constexpr int WIN_ROWS = 8;
__global__ void __launch_bounds__(256) mask_paste_window_kernel(const float* __restrict__ masks,
                                                                const float4* __restrict__ boxes, int M, int H, int W,
                                                                float thr, uint8_t* __restrict__ out) {
  const int d = blockIdx.y;
  const float4 bx = __ldg(boxes + d);
  const float bw = bx.z - bx.x, bh = bx.w - bx.y;
  int x_lo = 0, x_hi = W - 1, y_lo = 0, y_hi = H - 1;
  if (bw > 0.f && bh > 0.f && bw < 1e6f && bh < 1e6f && fabsf(bx.x) < 1e6f && fabsf(bx.y) < 1e6f) {
    // a pixel is sampled only when its centre is within bw / M (bh / M) of the box; + 1 pixel covers all rounding
    const float mx = bw / (float)M + 1.f, my = bh / (float)M + 1.f;
    x_lo = max(0, (int)floorf(bx.x - mx - 0.5f));
    x_hi = min(W - 1, (int)ceilf(bx.z + mx - 0.5f));
    y_lo = max(0, (int)floorf(bx.y - my - 0.5f));
    y_hi = min(H - 1, (int)ceilf(bx.w + my - 0.5f));
  }
  const int r0 = max(y_lo, (int)blockIdx.x * WIN_ROWS), r1 = min(y_hi, (int)blockIdx.x * WIN_ROWS + WIN_ROWS - 1);
  if (r0 > r1 || x_lo > x_hi) return;
  const float* mk = masks + (long long)d * M * M;
  uint8_t* dst = out + (long long)d * H * W;
  const int ww = x_hi - x_lo + 1, n = (r1 - r0 + 1) * ww;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int y = r0 + i / ww, x = x_lo + i % ww;
    dst[(long long)y * W + x] = paste_pixel(mk, M, bx, x, y, thr);
  }
}

// Separable single-pass paste for thr > 0: ONE launch writes the whole canvas with 8-byte stores (no memset node).
// grid_sample's coordinate chain depends on x only (column) or y only (row), so a CTA = (detection, strip of rows)
// evaluates it once per window column / strip row into shared-memory tables -- (x0, w, e, inside) and (y0, n, s,
// inside), the very values paste_sample computes per pixel, from the same rounded operations -- and a pixel costs one
// table read, four reads of the zero-bordered mask copy and the reference's product / sum order.  Pixels outside the
// window are written as zeros by the same threads.  Irregular boxes (empty, non-finite, huge) take paste_pixel per byte.
struct __align__(16) PasteCol {
  int x0;  // padded column index of the west tap (x0 + 2, clamped into the bordered copy)
  float w, e;
  int ok;  // |gx| < 1 + 2 / M
};
#ifndef UNIT_PASTE_STRIP
#define UNIT_PASTE_STRIP 16
#endif
constexpr int PASTE_STRIP = UNIT_PASTE_STRIP;  // canvas rows per CTA
constexpr int PASTE_PXT = 8;     // pixels per thread and store
// Column x lives in slot x ^ ((x >> 3) & 7): neighbouring lanes read columns 8 apart (16-byte entries, 128 bytes = all
// 32 banks apart -- an 8-way conflict in every quarter-warp phase of the LDS.128); the swizzle spreads them over the 8
// bank groups.  A slot stays inside its aligned group of 8, so the table is W rounded up to 8 entries.
__device__ __forceinline__ int paste_slot(int x) { return x ^ ((x >> 3) & 7); }
__host__ __device__ __forceinline__ int paste_cols(int W) { return (W + 7) & ~7; }

__global__ void __launch_bounds__(256) mask_paste_rows_kernel(const float* __restrict__ masks,
                                                              const float4* __restrict__ boxes, int M, int H, int W,
                                                              float thr, uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char paste_smem[];
  PasteCol* col = reinterpret_cast<PasteCol*>(paste_smem);       // [paste_cols(W)] swizzled, window columns only
  PasteCol* rowt = col + paste_cols(W);                          // [PASTE_STRIP]: (y0, n, s, ok) of the strip's rows
  float* mb = reinterpret_cast<float*>(rowt + PASTE_STRIP);      // [(M + 4)^2] mask with a 2-pixel zero border
  const int MP = M + 4;
  const int d = blockIdx.y;
  const int r0 = blockIdx.x * PASTE_STRIP, r1 = min(H, r0 + PASTE_STRIP);  // rows [r0, r1)
  const float4 bx = __ldg(boxes + d);
  const float bw = __fsub_rn(bx.z, bx.x), bh = __fsub_rn(bx.w, bx.y);
  const float* mk = masks + (long long)d * M * M;
  uint8_t* dst = out + (long long)d * H * W;
  const int a = r0 * W, b = r1 * W;  // this CTA's bytes of the detection's canvas
  const bool regular = bw > 0.f && bh > 0.f && bw < 1e6f && bh < 1e6f && fabsf(bx.x) < 1e6f && fabsf(bx.y) < 1e6f;
  if (!regular) {
    for (int f = a + (int)threadIdx.x; f < b; f += blockDim.x) {
      const int y = f / W;
      dst[f] = paste_pixel(mk, M, bx, f - y * W, y, thr);
    }
    return;
  }
  // a pixel is sampled only when its centre is within bw / M (bh / M) of the box; + 1 pixel covers all rounding
  const float mx = bw / (float)M + 1.f, my = bh / (float)M + 1.f;
  const int x_lo = max(0, (int)floorf(bx.x - mx - 0.5f)), x_hi = min(W - 1, (int)ceilf(bx.z + mx - 0.5f));
  const int y_lo = max(r0, (int)floorf(bx.y - my - 0.5f)), y_hi = min(r1 - 1, (int)ceilf(bx.w + my - 0.5f));
  const bool touch = y_lo <= y_hi && x_lo <= x_hi;  // warp-uniform (CTA-uniform)
  if (touch) {
    const float lim = 1.f + 2.f / (float)M, half = (float)M * 0.5f;
    for (int i = threadIdx.x; i < MP * MP; i += blockDim.x) {
      const int y = i / MP - 2, x = i % MP - 2;
      mb[i] = (x >= 0 && x < M && y >= 0 && y < M) ? __ldg(mk + y * M + x) : 0.f;
    }
    auto entry = [&](float p, float lo, float size) {  // the per-axis part of paste_pixel + paste_sample
      const float g = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(p, lo), size), 2.f), 1.f);
      const float i = __fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), half), 0.5f);
      const float fl = floorf(i);
      PasteCol c;
      c.w = __fsub_rn(i, fl);
      c.e = __fsub_rn(1.f, c.w);
      c.ok = fabsf(g) < lim;
      c.x0 = min(max((int)fl + 2, 0), M + 2);  // inside the border whenever ok
      return c;
    };
    for (int x = x_lo + (int)threadIdx.x; x <= x_hi; x += blockDim.x) col[paste_slot(x)] = entry((float)x + 0.5f, bx.x, bw);
    for (int y = y_lo + (int)threadIdx.x; y <= y_hi; y += blockDim.x) {
      PasteCol c = entry((float)y + 0.5f, bx.y, bh);
      c.x0 *= MP;
      rowt[y - r0] = c;
    }
    __syncthreads();
  }
  auto pixel = [&](int y, int x) -> uint32_t {
    if (!touch || y < y_lo || y > y_hi || x < x_lo || x > x_hi) return 0u;
    const PasteCol c = col[paste_slot(x)], r = rowt[y - r0];  // r: w = north fraction n, e = s
    const float* p = mb + r.x0 + c.x0;
    const float nw = __fmul_rn(p[0], __fmul_rn(r.e, c.e));
    const float ne = __fmul_rn(p[1], __fmul_rn(r.e, c.w));
    const float sw = __fmul_rn(p[MP], __fmul_rn(r.w, c.e));
    const float se = __fmul_rn(p[MP + 1], __fmul_rn(r.w, c.w));
    const float val = __fadd_rn(__fadd_rn(__fadd_rn(nw, ne), sw), se);
    return (c.ok && r.ok && val >= thr) ? 1u : 0u;
  };
  // bytes [a, b): an unaligned head, 8-byte words, a tail
  const int head = min(b - a, (int)((8 - ((uintptr_t)(dst + a) & 7)) & 7));
  const int nwords = (b - a - head) / PASTE_PXT;
  const int w0 = a + head;  // canvas byte of word 0
  const int tail0 = w0 + nwords * PASTE_PXT;
  uint2* words = reinterpret_cast<uint2*>(dst + w0);
  // phase 1: clear the strip (dense 8-byte stores)
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) words[i] = make_uint2(0u, 0u);
  if ((int)threadIdx.x < head) {
    const int f = a + threadIdx.x, y = f / W;
    dst[f] = (uint8_t)pixel(y, f - y * W);
  }
  if (tail0 + (int)threadIdx.x < b) {
    const int f = tail0 + threadIdx.x, y = f / W;
    dst[f] = (uint8_t)pixel(y, f - y * W);
  }
  if (!touch) return;
  __syncthreads();  // the window words below overwrite zeros written by other threads of this CTA
  // phase 2: the words that hold window pixels, threads packed densely over (window row, word of that row)
  const int nwr = ((x_hi - x_lo) >> 3) + 2;  // words per window row, upper bound
  const int total = (y_hi - y_lo + 1) * nwr;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const int yr = t / nwr, y = y_lo + yr;
    const int f_lo = y * W + x_lo - w0, f_hi = y * W + x_hi - w0;  // the row's window bytes relative to word 0
    const int wi = max(f_lo, 0) / PASTE_PXT + (t - yr * nwr);
    if (f_hi < 0 || wi > min(f_hi / PASTE_PXT, nwords - 1)) continue;
    int x = wi * PASTE_PXT + w0 - y * W;  // column of the word's first pixel on row y (may lie before the row)
    uint2 q = make_uint2(0u, 0u);
    if (x >= 0 && x + PASTE_PXT <= W) {
      // one row: one row entry; column entries clamped into the window, pixels outside it masked
      const PasteCol r = rowt[y - r0];
      const float* mrow = mb + r.x0;
#pragma unroll
      for (int j = 0; j < PASTE_PXT; ++j) {
        const int xj = x + j;
        const PasteCol c = col[paste_slot(min(max(xj, x_lo), x_hi))];
        const float* p = mrow + c.x0;
        const float nw = __fmul_rn(p[0], __fmul_rn(r.e, c.e));
        const float ne = __fmul_rn(p[1], __fmul_rn(r.e, c.w));
        const float sw = __fmul_rn(p[MP], __fmul_rn(r.w, c.e));
        const float se = __fmul_rn(p[MP + 1], __fmul_rn(r.w, c.w));
        const float val = __fadd_rn(__fadd_rn(__fadd_rn(nw, ne), sw), se);
        const bool on = xj >= x_lo && xj <= x_hi && (c.ok & r.ok) != 0 && val >= thr;
        const uint32_t o = on ? (1u << (8 * (j & 3))) : 0u;
        if (j < 4) q.x |= o;
        else q.y |= o;
      }
    } else {
      // the word straddles two canvas rows (a window that touches the left / right border): every pixel at its own
      // (row, column); when both rows are window rows the word is written twice with identical contents
      int yy = y;
      if (x < 0) {
        x += W;
        --yy;
      }
#pragma unroll 1
      for (int j = 0; j < PASTE_PXT; ++j) {
        const uint32_t o = pixel(yy, x) << (8 * (j & 3));
        if (j < 4) q.x |= o;
        else q.y |= o;
        if (++x == W) {
          x = 0;
          ++yy;
        }
      }
    }
    words[wi] = q;
  }
}

}  // namespace transfer
}  // namespace unit

using namespace unit;
using namespace unit::transfer;

extern "C" {

int unit_lingual_similarity(const float* emb, const int64_t* indexer, const int64_t* base, const int64_t* novel,
                            int D, int B, int Nn, float* raw, float* soft, unit_stream_t stream) {
  UNIT_REQUIRE(D > 0 && B > 0 && Nn >= 0, "lingual_similarity: bad shape");
  if (Nn == 0) return UNIT_OK;
  UNIT_REQUIRE(emb && indexer && base && novel && (raw || soft), "lingual_similarity: null pointer");
  lingual_kernel<<<Nn, 128, B * sizeof(float), (cudaStream_t)stream>>>(emb, indexer, base, novel, D, B, raw, soft);
  UNIT_CHECK_LAUNCH("lingual_kernel");
  return UNIT_OK;
}

int unit_similarity_transfer(const unit_transfer_params* p, const float* vis_logits, const float* static_cls,
                             const float* static_bbox, const float* static_seg, const int* base, const int* novel,
                             const int* class_kind, const float* delta_scores, const float* proposal_deltas,
                             const float* weak_scores, const float* ft_scores, const float* ft_deltas,
                             float* out_scores, float* out_bbox, float* out_s_cls, float* out_s_bbox,
                             float* out_s_seg, unit_stream_t stream) {
  UNIT_REQUIRE(p, "similarity_transfer: null params");
  UNIT_REQUIRE(p->R >= 0 && p->K > 0 && p->B >= 0 && p->Nn >= 0, "similarity_transfer: bad shape");
  if (p->R == 0) return UNIT_OK;
  UNIT_REQUIRE(delta_scores && proposal_deltas && out_scores && out_bbox && class_kind,
               "similarity_transfer: null pointer");
  const bool need_v = p->do_transfer && (p->wv_cls != 0.f || p->wv_bbox != 0.f || p->wv_seg != 0.f);
  UNIT_REQUIRE(!need_v || vis_logits, "similarity_transfer: a head uses the visual term but vis_logits is NULL");
  UNIT_REQUIRE(!p->do_transfer || (base && novel) || (p->B == 0 && p->Nn == 0),
               "similarity_transfer: base/novel index arrays missing");
  UNIT_REQUIRE((p->ld_delta_scores == 0 || p->ld_delta_scores >= p->K + 1) &&
                   (p->ld_proposal_deltas == 0 || p->ld_proposal_deltas >= 4 * p->K) &&
                   (p->ld_ft_scores == 0 || p->ld_ft_scores >= p->K + 1) &&
                   (p->ld_ft_deltas == 0 || p->ld_ft_deltas >= 4 * p->K) &&
                   (p->ld_vis_logits == 0 || p->ld_vis_logits >= p->K + 1) &&
                   (p->ld_weak_scores == 0 || p->ld_weak_scores >= p->K + 1),
               "similarity_transfer: a row stride is smaller than its row");
  TransferArgs a;
  a.p = *p;
  a.vis_logits = vis_logits;
  a.stat[0] = static_cls;
  a.stat[1] = static_bbox;
  a.stat[2] = static_seg;
  a.base = base;
  a.novel = novel;
  a.class_kind = class_kind;
  a.delta_scores = delta_scores;
  a.proposal_deltas = proposal_deltas;
  a.weak_scores = weak_scores;
  a.ft_scores = ft_scores;
  a.ft_deltas = ft_deltas;
  a.out_scores = out_scores;
  a.out_bbox = out_bbox;
  a.out_s[0] = out_s_cls;
  a.out_s[1] = out_s_bbox;
  a.out_s[2] = out_s_seg;
  const size_t smem = (size_t)(p->B + (p->K + 1) + 4 * p->K + 5 * p->Nn + 8) * sizeof(float);
  similarity_transfer_kernel<<<p->R, TT, smem, (cudaStream_t)stream>>>(a);
  UNIT_CHECK_LAUNCH("similarity_transfer_kernel");
  return UNIT_OK;
}

int unit_similarity_transfer_bwd(const unit_transfer_params* p, const float* s_cls, const float* s_bbox,
                                 const int* base, const int* novel, const int* class_kind, const float* g_scores,
                                 const float* g_bbox, int detach_transfer, float* g_delta_scores,
                                 float* g_proposal_deltas, unit_stream_t stream) {
  UNIT_REQUIRE(p, "similarity_transfer_bwd: null params");
  if (p->R == 0) return UNIT_OK;
  UNIT_REQUIRE(g_scores && g_bbox && g_delta_scores && g_proposal_deltas && class_kind,
               "similarity_transfer_bwd: null pointer");
  UNIT_REQUIRE(detach_transfer || (s_cls && s_bbox && base && novel), "similarity_transfer_bwd: similarity missing");
  TransferBwdArgs a = {p->R, p->K, p->B, p->Nn, s_cls, s_bbox, base, novel, class_kind,
                       g_scores, g_bbox, detach_transfer, g_delta_scores, g_proposal_deltas};
  similarity_transfer_bwd_kernel<<<p->R, TT, 0, (cudaStream_t)stream>>>(a);
  UNIT_CHECK_LAUNCH("similarity_transfer_bwd_kernel");
  return UNIT_OK;
}

int unit_similarity_transfer_bwd_vis(const unit_transfer_params* p, const float* vis_logits, const float* static_cls,
                                     const float* static_bbox, const int* base, const int* novel,
                                     const float* delta_scores, const float* proposal_deltas, const float* g_scores,
                                     const float* g_bbox, float* g_vis_logits, unit_stream_t stream) {
  UNIT_REQUIRE(p, "similarity_transfer_bwd_vis: null params");
  UNIT_REQUIRE(p->R >= 0 && p->K > 0 && p->B > 0 && p->Nn > 0, "similarity_transfer_bwd_vis: bad shape");
  if (p->R == 0) return UNIT_OK;
  UNIT_REQUIRE(vis_logits && base && novel && delta_scores && proposal_deltas && g_scores && g_bbox && g_vis_logits,
               "similarity_transfer_bwd_vis: null pointer");
  TransferBwdVisArgs a;
  a.p = *p;
  a.vis_logits = vis_logits;
  a.stat[0] = static_cls;
  a.stat[1] = static_bbox;
  a.base = base;
  a.novel = novel;
  a.delta_scores = delta_scores;
  a.proposal_deltas = proposal_deltas;
  a.g_scores = g_scores;
  a.g_bbox = g_bbox;
  a.g_vis = g_vis_logits;
  const size_t smem = (size_t)((p->K + 1) + 2 * p->B + 8) * sizeof(float);
  similarity_transfer_bwd_vis_kernel<<<p->R, TT, smem, (cudaStream_t)stream>>>(a);
  UNIT_CHECK_LAUNCH("similarity_transfer_bwd_vis_kernel");
  return UNIT_OK;
}

int unit_mask_transfer(const float* logits, const float* s_seg, int s_is_2d, const int* base, const int* novel,
                       const int* class_kind, const float* x_delta, const int64_t* pred_classes, float* out_logits,
                       float* out_probs, int D, int K, int B, int Nn, int MM, unit_stream_t stream) {
  (void)novel;
  UNIT_REQUIRE(D >= 0 && K > 0 && MM > 0, "mask_transfer: bad shape");
  if (D == 0) return UNIT_OK;
  UNIT_REQUIRE(logits && class_kind && (out_logits || out_probs), "mask_transfer: null pointer");
  UNIT_REQUIRE(!out_probs || pred_classes, "mask_transfer: out_probs needs pred_classes");
  UNIT_REQUIRE(!s_seg || base, "mask_transfer: similarity given without base indices");
  // probabilities only (inference): one thread per pixel; full logits: a few chunks of (class, pixel) pairs
  const int threads = 128;
  const int chunks = out_logits ? 8 : (MM + threads - 1) / threads;
  dim3 grid(D, chunks);
  mask_transfer_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(logits, s_seg, s_is_2d, base, class_kind, x_delta,
                                                                  pred_classes, out_logits, out_probs, K, B, Nn, MM);
  UNIT_CHECK_LAUNCH("mask_transfer_kernel");
  return UNIT_OK;
}

int unit_mask_paste(const float* masks, const float* boxes, int D, int M, int img_h, int img_w, float threshold,
                    uint8_t* out, unit_stream_t stream) {
  UNIT_REQUIRE(D >= 0 && M > 0 && img_h > 0 && img_w > 0, "mask_paste: bad shape");
  if (D == 0) return UNIT_OK;
  UNIT_REQUIRE(masks && boxes && out, "mask_paste: null pointer");
  UNIT_REQUIRE((((uintptr_t)boxes) & 15) == 0, "mask_paste: boxes must be 16-byte aligned");
  const long long total = (long long)D * img_h * img_w;
  if (threshold > 0.f && D <= 65535 && !switches().paste_flat) {  // outside value is 0: only the box windows are sampled
    using namespace unit::transfer;
    const size_t smem = ((size_t)paste_cols(img_w) + PASTE_STRIP) * sizeof(PasteCol) + (size_t)(M + 4) * (M + 4) * sizeof(float);
    if (smem <= 200 * 1024 && (long long)img_h * img_w < (1ll << 31)) {  // one launch: zero fill + separable sampling
      if (smem > 48 * 1024)
        UNIT_CUDA(cudaFuncSetAttribute(mask_paste_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      mask_paste_rows_kernel<<<dim3(cdiv(img_h, PASTE_STRIP), D), 256, smem, (cudaStream_t)stream>>>(
          masks, (const float4*)boxes, M, img_h, img_w, threshold, out);
      UNIT_CHECK_LAUNCH("mask_paste_rows_kernel");
      return UNIT_OK;
    }
    UNIT_CUDA(cudaMemsetAsync(out, 0, (size_t)total, (cudaStream_t)stream));  // very wide canvases: memset + windows
    mask_paste_window_kernel<<<dim3(cdiv(img_h, WIN_ROWS), D), 256, 0, (cudaStream_t)stream>>>(
        masks, (const float4*)boxes, M, img_h, img_w, threshold, out);
    UNIT_CHECK_LAUNCH("mask_paste_window_kernel");
    return UNIT_OK;
  }
  const long long threads = (total + unit::transfer::PX - 1) / unit::transfer::PX;
  mask_paste_kernel<<<cdiv(threads, 256), 256, 0, (cudaStream_t)stream>>>(masks, (const float4*)boxes, D, M, img_h,
                                                                         img_w, threshold, out);
  UNIT_CHECK_LAUNCH("mask_paste_kernel");
  return UNIT_OK;
}

}  // extern "C"
