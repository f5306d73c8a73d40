// Weak-image training losses of the base stage.
//
//   mil_*_kernel         weak_detector_fast_rcnn.py:189-214  x = softmax_classes(cls) * softmax_proposals(det), the
//                        image-level BCE on sum_r x (clamped to [eps, 1-eps]) AND its gradients, in three grid-wide
//                        phases over 128-row chunks (a single CTA per image took 0.28 ms for 2 x 2000 proposals)
//   oicr_targets_kernel  :353-408 + :308-351  get_proposal_clusters (per present class, ascending: the proposal with
//                        the highest score, whose row is then zeroed) -> pairwise_iou + UniT Matcher against those
//                        pseudo boxes -> refinement labels and per-proposal loss weights, one CTA per image
//   weighted_ce_kernel   :220-227  mean(cross_entropy(score, label, 'none') * weight) and its gradient
//
// Labels are bit-exact (argmax ties resolve to the first proposal, IoU ops individually rounded as in match.cu);
// losses / gradients differ from the reference only in fp32 summation order.
#include "match_common.cuh"

namespace unit {
namespace weak {

using match::Thresholds;

constexpr int NT = 1024;
constexpr int NWARP = NT / 32;
constexpr int KMAX = 128;  // classes (VOC 20, COCO 80)
constexpr int KL = KMAX / 32;

__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// p[j] = softmax over the K classes of one row (lane owns classes lane + 32 j)
__device__ __forceinline__ void row_softmax(const float* __restrict__ row, int K, int lane, float (&p)[KL]) {
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < KL; ++j) {
    p[j] = lane + 32 * j < K ? row[lane + 32 * j] : -INFINITY;
    m = fmaxf(m, p[j]);
  }
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < KL; ++j) {
    p[j] = lane + 32 * j < K ? expf(p[j] - m) : 0.f;
    s += p[j];
  }
  s = warp_sum(s);
#pragma unroll
  for (int j = 0; j < KL; ++j) p[j] = __fdiv_rn(p[j], s);
}

// ---- MIL loss in three grid-wide phases (one CTA per 128-row chunk of an image; an image's softmax over its
// proposals needs column statistics of ALL its rows, so the phases are separate launches):
//   1 mil_colstats_kernel  per chunk and class: max of the detection logits and sum of exp(d - max)
//   2 mil_scores_kernel    combine the chunk statistics, x = p * q for the chunk's rows, per-chunk class sums
//   3 mil_grads_kernel     class vector, clamped BCE term, g = dL/dv, both gradients for the chunk's rows
constexpr int MT = 256;           // threads per CTA
constexpr int MW = MT / 32;       // warps
constexpr int MROWS = 128;        // rows per chunk

template <bool MAX>
__device__ __forceinline__ void chunk_reduce(float (*part)[KMAX], const float (&acc)[KL], float* dst, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < KL; ++j)
    if (lane + 32 * j < K) part[warp][lane + 32 * j] = acc[j];
  __syncthreads();
  if ((int)threadIdx.x < K) {
    float v = part[0][threadIdx.x];
    for (int w = 1; w < MW; ++w) v = MAX ? fmaxf(v, part[w][threadIdx.x]) : v + part[w][threadIdx.x];
    dst[threadIdx.x] = v;
  }
  __syncthreads();
}

// stats [n_img][nch][K][2] = (chunk max, chunk sum of exp(d - chunk max)); an empty chunk writes (-inf, 0)
__global__ void __launch_bounds__(MT) mil_colstats_kernel(const float* __restrict__ det, const int* __restrict__ img_off,
                                                          int K, int nch, float* __restrict__ stats) {
  __shared__ float part[MW][KMAX];
  __shared__ float cmax[KMAX], csum[KMAX];
  const int img = blockIdx.y, ch = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = img_off[img] + ch * MROWS, r1 = min(img_off[img + 1], r0 + MROWS);
  float acc[KL];
#pragma unroll
  for (int j = 0; j < KL; ++j) acc[j] = -INFINITY;
  for (int r = r0 + warp; r < r1; r += MW) {
#pragma unroll
    for (int j = 0; j < KL; ++j)
      if (lane + 32 * j < K) acc[j] = fmaxf(acc[j], det[(long long)r * K + lane + 32 * j]);
  }
  chunk_reduce<true>(part, acc, cmax, K);
#pragma unroll
  for (int j = 0; j < KL; ++j) acc[j] = 0.f;
  for (int r = r0 + warp; r < r1; r += MW) {
#pragma unroll
    for (int j = 0; j < KL; ++j)
      if (lane + 32 * j < K) acc[j] += expf(det[(long long)r * K + lane + 32 * j] - cmax[lane + 32 * j]);
  }
  chunk_reduce<false>(part, acc, csum, K);
  if ((int)threadIdx.x < K) {
    float* o = stats + (((long long)img * nch + ch) * K + threadIdx.x) * 2;
    o[0] = cmax[threadIdx.x];
    o[1] = csum[threadIdx.x];
  }
}

// column max / sum of an image from its chunk statistics, in chunk order (deterministic)
__device__ __forceinline__ void combine_stats(const float* __restrict__ stats, int img, int nch, int K, float* colmax,
                                              float* colsum) {
  if ((int)threadIdx.x < K) {
    const float* s = stats + ((long long)img * nch * K + threadIdx.x) * 2;
    float m = -INFINITY;
    for (int c = 0; c < nch; ++c) m = fmaxf(m, s[(long long)c * K * 2]);
    float t = 0.f;
    for (int c = 0; c < nch; ++c) {
      const float cm = s[(long long)c * K * 2], cs = s[(long long)c * K * 2 + 1];
      if (cs != 0.f) t += cs * expf(cm - m);
    }
    colmax[threadIdx.x] = m;
    colsum[threadIdx.x] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(MT) mil_scores_kernel(const float* __restrict__ cls, const float* __restrict__ det,
                                                        const int* __restrict__ img_off, const float* __restrict__ stats,
                                                        int K, int nch, float* __restrict__ mil,
                                                        float* __restrict__ vpart) {
  __shared__ float part[MW][KMAX];
  __shared__ float colmax[KMAX], colsum[KMAX], vsum[KMAX];
  const int img = blockIdx.y, ch = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = img_off[img] + ch * MROWS, r1 = min(img_off[img + 1], r0 + MROWS);
  combine_stats(stats, img, nch, K, colmax, colsum);
  float acc[KL];
#pragma unroll
  for (int j = 0; j < KL; ++j) acc[j] = 0.f;
  for (int r = r0 + warp; r < r1; r += MW) {
    float p[KL];
    row_softmax(cls + (long long)r * K, K, lane, p);
#pragma unroll
    for (int j = 0; j < KL; ++j) {
      const int k = lane + 32 * j;
      if (k < K) {
        const float q = __fdiv_rn(expf(det[(long long)r * K + k] - colmax[k]), colsum[k]);
        const float x = p[j] * q;
        mil[(long long)r * K + k] = x;
        acc[j] += x;
      }
    }
  }
  chunk_reduce<false>(part, acc, vsum, K);
  if ((int)threadIdx.x < K) vpart[((long long)img * nch + ch) * K + threadIdx.x] = vsum[threadIdx.x];
}

__global__ void __launch_bounds__(MT) mil_grads_kernel(const float* __restrict__ cls, const float* __restrict__ det,
                                                       const int* __restrict__ img_off, const float* __restrict__ gt_vec,
                                                       const float* __restrict__ stats, const float* __restrict__ vpart,
                                                       int K, int nch, float scale, float eps,
                                                       float* __restrict__ class_vec, float* __restrict__ img_loss,
                                                       float* __restrict__ d_cls, float* __restrict__ d_det) {
  __shared__ float colmax[KMAX], colsum[KMAX], vcol[KMAX], gcol[KMAX], term[KMAX];
  const int img = blockIdx.y, ch = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = img_off[img] + ch * MROWS, r1 = min(img_off[img + 1], r0 + MROWS);
  if (ch > 0 && r0 >= r1) return;  // chunk 0 always runs: it owns the image's class vector and loss term
  if ((int)threadIdx.x < K) {
    const int k = threadIdx.x;
    float v = 0.f;
    for (int c = 0; c < nch; ++c) v += vpart[((long long)img * nch + c) * K + k];
    const float y = gt_vec[(long long)img * K + k];
    const float vc = fminf(fmaxf(v, eps), 1.f - eps);
    vcol[k] = v;
    term[k] = -(y * fmaxf(logf(vc), -100.f) + (1.f - y) * fmaxf(log1pf(-vc), -100.f));
    gcol[k] = (v >= eps && v <= 1.f - eps) ? scale * ((1.f - y) / (1.f - vc) - y / vc) : 0.f;
    if (ch == 0) class_vec[(long long)img * K + k] = v;
  }
  combine_stats(stats, img, nch, K, colmax, colsum);  // ends with a barrier: vcol / gcol / term are visible too
  if (ch == 0 && threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < K; ++k) t += term[k];
    img_loss[img] = t * scale;
  }
  // gradients: d_det = g q (p - v),  d_cls = p (g q - sum_k g_k x_k)
  for (int r = r0 + warp; r < r1; r += MW) {
    float p[KL], q[KL];
    row_softmax(cls + (long long)r * K, K, lane, p);
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < KL; ++j) {
      const int k = lane + 32 * j;
      q[j] = 0.f;
      if (k < K) {
        q[j] = __fdiv_rn(expf(det[(long long)r * K + k] - colmax[k]), colsum[k]);
        t += gcol[k] * p[j] * q[j];
      }
    }
    t = warp_sum(t);
#pragma unroll
    for (int j = 0; j < KL; ++j) {
      const int k = lane + 32 * j;
      if (k < K) {
        d_cls[(long long)r * K + k] = p[j] * (gcol[k] * q[j] - t);
        d_det[(long long)r * K + k] = gcol[k] * q[j] * (p[j] - vcol[k]);
      }
    }
  }
}

// out[0] = sum of n values, fixed order
__global__ void sum_kernel(const float* __restrict__ v, int n, float mul, float* __restrict__ out) {
  __shared__ float s[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += v[i];
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s[w];
    out[0] = t * mul;
  }
}

__global__ void __launch_bounds__(NT) oicr_targets_kernel(const float* __restrict__ probs, int ld,
                                                          const float4* __restrict__ props,
                                                          const int* __restrict__ prop_off,
                                                          const float* __restrict__ gt_vec, Thresholds t, int K,
                                                          float bg_thresh, int64_t* __restrict__ labels,
                                                          float* __restrict__ weights, int64_t* __restrict__ pgt_index,
                                                          float* __restrict__ pgt_scores) {
  extern __shared__ uint32_t used[];  // one bit per proposal of the image: its score row has been zeroed
  __shared__ float4 gbox[KMAX];
  __shared__ float gscore[KMAX];
  __shared__ int gcls[KMAX];
  __shared__ float red_v[NWARP];
  __shared__ int red_i[NWARP];
  __shared__ int s_G;
  const int img = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = prop_off[img], R = prop_off[img + 1] - r0;
  for (int i = tid; i < (R + 31) / 32; i += NT) used[i] = 0u;
  if (tid == 0) s_G = 0;
  __syncthreads();
  // ---- get_proposal_clusters: classes present in the image, ascending (== torch.unique order)
  for (int c = 0; c < K; ++c) {
    const bool present = gt_vec[(long long)img * K + c] != 0.f;  // block-uniform
    if (tid == 0) {
      pgt_index[(long long)img * K + c] = -1;
      pgt_scores[(long long)img * K + c] = 0.f;
    }
    if (!present || R == 0) continue;
    float bv = -1.f;
    int bi = 0x7fffffff;
    for (int r = tid; r < R; r += NT) {
      const float v = (used[r >> 5] >> (r & 31)) & 1u ? 0.f : probs[(long long)(r0 + r) * ld + c];
      if (v > bv) {
        bv = v;
        bi = r;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      red_v[warp] = bv;
      red_i[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < NWARP; ++w)
        if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) {
          bv = red_v[w];
          bi = red_i[w];
        }
      if (bi == 0x7fffffff) bi = 0;  // every score is NaN
      used[bi >> 5] |= 1u << (bi & 31);
      const int g = s_G++;
      gbox[g] = props[r0 + bi];
      gscore[g] = bv;
      gcls[g] = c;
      pgt_index[(long long)img * K + c] = bi;
      pgt_scores[(long long)img * K + c] = bv;
    }
    __syncthreads();
  }
  // ---- pairwise_iou(pseudo boxes, proposals) + Matcher + labels / weights
  const int G = s_G;
  for (int r = tid; r < R; r += NT) {
    int64_t label = K;
    float w = 0.f;
    if (G > 0) {
      const float4 b = __ldg(props + r0 + r);
      const float area_b = box_area_rn(b.x, b.y, b.z, b.w);
      float best = 0.f;
      int arg = 0;
      for (int g = 0; g < G; ++g) {
        const float4 a = gbox[g];
        const float inter = box_inter_rn(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w);
        float v = 0.f;
        if (inter > 0.f) v = iou_from_rn(inter, box_area_rn(a.x, a.y, a.z, a.w), area_b);
        if (g == 0 || v > best || (v != v && best == best)) {
          best = v;
          arg = g;
        }
      }
      const int lab = match::bin_label(best, t);
      label = lab == 1 ? gcls[arg] : (lab == 0 ? K : -1);
      w = gscore[arg];
      if (bg_thresh > 0.f && best < bg_thresh) w = 0.f;
    }
    labels[r0 + r] = label;
    weights[r0 + r] = w;
  }
}

__global__ void weighted_ce_kernel(const float* __restrict__ scores, const int64_t* __restrict__ labels,
                                   const float* __restrict__ weights, int R, int K1, float* __restrict__ row_loss,
                                   float* __restrict__ d_scores) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= R) return;
  const float invR = 1.f / (float)R;
  const int cls = (int)labels[r];
  const bool ok = cls >= 0 && cls < K1;
  const float w = ok ? weights[r] : 0.f;
  const float* s = scores + (long long)r * K1;
  float m = -INFINITY;
  for (int k = lane; k < K1; k += 32) m = fmaxf(m, s[k]);
  m = warp_max(m);
  float sum = 0.f;
  for (int k = lane; k < K1; k += 32) sum += expf(s[k] - m);
  sum = warp_sum(sum);
  const float lse = m + logf(sum);
  for (int k = lane; k < K1; k += 32)
    d_scores[(long long)r * K1 + k] = (expf(s[k] - lse) - (k == cls ? 1.f : 0.f)) * (w * invR);
  if (lane == 0) row_loss[r] = ok ? (lse - s[cls]) * w : 0.f;
}

}  // namespace weak
}  // namespace unit

using namespace unit;

extern "C" {

size_t unit_mil_loss_workspace_bytes(int n_img, int max_rows, int K) {
  const size_t nch = (size_t)((max_rows > 0 ? max_rows : 1) + weak::MROWS - 1) / weak::MROWS;
  return ((size_t)n_img * nch * K * 3 + (size_t)n_img) * sizeof(float);
}

int unit_mil_loss(const float* cls_logits, const float* det_logits, const int* img_offsets, const float* gt_vector,
                  int n_img, int R, int max_rows, int K, float multiplier, float* mil_scores, float* class_vector,
                  float* loss, float* d_cls, float* d_det, void* workspace, size_t workspace_bytes,
                  unit_stream_t stream) {
  UNIT_REQUIRE(n_img > 0 && R >= 0 && K > 0 && K <= weak::KMAX, "mil_loss: bad shape (K <= %d)", weak::KMAX);
  UNIT_REQUIRE(max_rows >= 0 && max_rows <= R, "mil_loss: max_rows must be the largest per-image row count");
  UNIT_REQUIRE(img_offsets && gt_vector && class_vector && loss, "mil_loss: null pointer");
  UNIT_REQUIRE(R == 0 || (cls_logits && det_logits && mil_scores && d_cls && d_det), "mil_loss: null pointer");
  if (!workspace || workspace_bytes < unit_mil_loss_workspace_bytes(n_img, max_rows, K)) {
    set_error("mil_loss: workspace too small");
    return UNIT_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int nch = ((max_rows > 0 ? max_rows : 1) + weak::MROWS - 1) / weak::MROWS;
  UNIT_REQUIRE(n_img <= 65535, "mil_loss: at most 65535 images per call");
  float* stats = (float*)workspace;                       // [n_img][nch][K][2]
  float* vpart = stats + (size_t)n_img * nch * K * 2;      // [n_img][nch][K]
  float* img_loss = vpart + (size_t)n_img * nch * K;       // [n_img]
  // F.binary_cross_entropy 'mean' over the [n_img, K] class vectors, times MIL_MULTIPLIER
  const float scale = multiplier / ((float)n_img * (float)K);
  dim3 grid(nch, n_img);
  weak::mil_colstats_kernel<<<grid, weak::MT, 0, st>>>(det_logits, img_offsets, K, nch, stats);
  UNIT_CHECK_LAUNCH("mil_colstats_kernel");
  weak::mil_scores_kernel<<<grid, weak::MT, 0, st>>>(cls_logits, det_logits, img_offsets, stats, K, nch, mil_scores,
                                                      vpart);
  UNIT_CHECK_LAUNCH("mil_scores_kernel");
  weak::mil_grads_kernel<<<grid, weak::MT, 0, st>>>(cls_logits, det_logits, img_offsets, gt_vector, stats, vpart, K,
                                                     nch, scale, 1e-6f, class_vector, img_loss, d_cls, d_det);
  UNIT_CHECK_LAUNCH("mil_grads_kernel");
  weak::sum_kernel<<<1, 256, 0, st>>>(img_loss, n_img, 1.f, loss);
  UNIT_CHECK_LAUNCH("sum_kernel");
  return UNIT_OK;
}

int unit_oicr_targets(const float* probs, int ld, const float* prop_boxes, const int* prop_offsets,
                      const float* gt_vector, int n_img, int P_total, int K, const float* thresholds_host,
                      const int* labels_host, int T, float bg_threshold, int64_t* labels, float* weights,
                      int64_t* pgt_index, float* pgt_scores, unit_stream_t stream) {
  UNIT_REQUIRE(n_img > 0 && P_total >= 0 && K > 0 && K <= weak::KMAX && ld >= K, "oicr_targets: bad shape");
  UNIT_REQUIRE(prop_offsets && gt_vector && pgt_index && pgt_scores, "oicr_targets: null pointer");
  UNIT_REQUIRE(P_total == 0 || (probs && prop_boxes && labels && weights), "oicr_targets: null pointer");
  UNIT_REQUIRE(((uintptr_t)prop_boxes & 15) == 0, "oicr_targets: boxes must be 16-byte aligned");
  match::Thresholds t;
  if (int rc = match::make_thresholds(thresholds_host, labels_host, T, &t)) return rc;
  const size_t smem = ((size_t)P_total / 32 + 1) * sizeof(uint32_t);
  UNIT_REQUIRE(smem <= 40 * 1024, "oicr_targets: more than %d proposals in one call", 40 * 1024 * 8);
  cudaStream_t st = (cudaStream_t)stream;
  weak::oicr_targets_kernel<<<n_img, weak::NT, smem, st>>>(probs, ld, (const float4*)prop_boxes, prop_offsets,
                                                            gt_vector, t, K, bg_threshold, labels, weights, pgt_index,
                                                            pgt_scores);
  UNIT_CHECK_LAUNCH("oicr_targets_kernel");
  return UNIT_OK;
}

int unit_weighted_ce_loss(const float* scores, const int64_t* labels, const float* weights, int R, int K1,
                          float* loss, float* d_scores, void* workspace, size_t workspace_bytes,
                          unit_stream_t stream) {
  UNIT_REQUIRE(R >= 0 && K1 > 0, "weighted_ce_loss: bad shape");
  UNIT_REQUIRE(loss, "weighted_ce_loss: null loss");
  cudaStream_t st = (cudaStream_t)stream;
  if (R == 0) {  // weak_detector_fast_rcnn.py:225-227: 0.0 * score.sum()
    UNIT_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    return UNIT_OK;
  }
  UNIT_REQUIRE(scores && labels && weights && d_scores, "weighted_ce_loss: null pointer");
  if (!workspace || workspace_bytes < (size_t)R * sizeof(float)) {
    set_error("weighted_ce_loss: workspace too small");
    return UNIT_EWORKSPACE;
  }
  weak::weighted_ce_kernel<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(scores, labels, weights, R, K1,
                                                                          (float*)workspace, d_scores);
  UNIT_CHECK_LAUNCH("weighted_ce_kernel");
  weak::sum_kernel<<<1, 1024, 0, st>>>((const float*)workspace, R, 1.f / (float)R, loss);
  UNIT_CHECK_LAUNCH("sum_kernel");
  return UNIT_OK;
}

}  // extern "C"
