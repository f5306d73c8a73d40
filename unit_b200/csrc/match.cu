// pairwise IoU, Matcher, fg/bg labelling and sampling.
//
// Integer outputs (matches, labels, sampled indices) are bit-exact against the reference path: every fp32
// operation of the IoU is rounded separately (no FMA contraction) in the reference's evaluation order, argmax ties
// resolve to the first GT, threshold bins are half-open [low, high) compared in fp32.
#include <algorithm>

#include "match_common.cuh"

namespace unit {
namespace match {

__global__ void pairwise_iou_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2,
                                    float* __restrict__ iou, int G, int P) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  if (p >= P || g >= G) return;
  const float4 a = __ldg(b1 + g), b = __ldg(b2 + p);
  const float inter = box_inter_rn(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w);
  float v = 0.f;
  if (inter > 0.f)
    v = iou_from_rn(inter, box_area_rn(a.x, a.y, a.z, a.w), box_area_rn(b.x, b.y, b.z, b.w));
  iou[(long long)g * P + p] = v;
}

// column max / first argmax over the G rows of a [G,P] matrix
__global__ void matcher_kernel(const float* __restrict__ iou, int G, int P, Thresholds t, int64_t* __restrict__ matches,
                               int8_t* __restrict__ labels, float* __restrict__ vals) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  if (G == 0) {
    matches[p] = 0;
    labels[p] = (int8_t)t.labels[0];
    if (vals) vals[p] = 0.f;
    return;
  }
  float best = __ldg(iou + p);
  int arg = 0;
  for (int g = 1; g < G; ++g) {
    const float v = __ldg(iou + (long long)g * P + p);
    if (v > best || (v != v && best == best)) {  // torch.max propagates NaN
      best = v;
      arg = g;
    }
  }
  matches[p] = arg;
  labels[p] = bin_label(best, t);
  if (vals) vals[p] = best;
}

// set_low_quality_matches_ (modeling/matcher.py:100-119): row max, then every prediction attaining it -> label 1
__global__ void row_max_kernel(const float* __restrict__ iou, int G, int P, float* __restrict__ rowmax) {
  const int g = blockIdx.x;
  float m = -INFINITY;
  for (int p = threadIdx.x; p < P; p += blockDim.x) m = fmaxf(m, __ldg(iou + (long long)g * P + p));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) rowmax[g] = m;
  }
}
__global__ void low_quality_kernel(const float* __restrict__ iou, const float* __restrict__ rowmax, int G, int P,
                                   int8_t* __restrict__ labels) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  bool hit = false;
  for (int g = 0; g < G; ++g) hit |= (__ldg(iou + (long long)g * P + p) == __ldg(rowmax + g));
  if (hit) labels[p] = 1;
}

// fused pairwise_iou + Matcher, all images in one launch: thread = proposal, loops its image's GT boxes
__global__ void iou_match_kernel(const float4* __restrict__ gt, const int* __restrict__ gt_off,
                                 const float4* __restrict__ props, const int* __restrict__ prop_off, int n_img,
                                 int P_total, Thresholds t, int64_t* __restrict__ matches, int8_t* __restrict__ labels,
                                 float* __restrict__ vals) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P_total) return;
  const int img = find_segment(prop_off, n_img, p);
  const int g0 = gt_off[img], g1 = gt_off[img + 1];
  if (g1 == g0) {
    matches[p] = 0;
    labels[p] = (int8_t)t.labels[0];
    if (vals) vals[p] = 0.f;
    return;
  }
  const float4 b = __ldg(props + p);
  const float area_b = box_area_rn(b.x, b.y, b.z, b.w);
  float best = 0.f;
  int arg = 0;
  for (int g = g0; g < g1; ++g) {
    const float4 a = __ldg(gt + g);
    const float inter = box_inter_rn(a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w);
    float v = 0.f;
    if (inter > 0.f) v = iou_from_rn(inter, box_area_rn(a.x, a.y, a.z, a.w), area_b);
    if (g == g0 || v > best || (v != v && best == best)) {
      best = v;
      arg = g - g0;
    }
  }
  matches[p] = arg;
  labels[p] = bin_label(best, t);
  if (vals) vals[p] = best;
}

// One CTA per image: class assignment + ordered compaction of foreground / background indices.
__global__ void __launch_bounds__(1024) label_kernel(const int64_t* __restrict__ matches,
                                                      const int8_t* __restrict__ mlabels,
                                                      const int64_t* __restrict__ gt_classes,
                                                      const int* __restrict__ gt_off, const int* __restrict__ prop_off,
                                                      int num_classes, int64_t* __restrict__ prop_classes,
                                                      int64_t* __restrict__ pos_idx, int64_t* __restrict__ neg_idx,
                                                      int* __restrict__ counts) {
  const int img = blockIdx.x;
  const int p0 = prop_off[img], p1 = prop_off[img + 1];
  const int g0 = gt_off[img], ng = gt_off[img + 1] - g0;
  __shared__ int warp_pos[32], warp_neg[32];
  __shared__ int base_pos, base_neg, round_pos, round_neg;
  if (threadIdx.x == 0) {
    base_pos = 0;
    base_neg = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int start = p0; start < p1; start += blockDim.x) {
    const int p = start + threadIdx.x;
    long long cls = 0;
    bool is_pos = false, is_neg = false;
    if (p < p1) {
      if (ng > 0) {
        cls = gt_classes[g0 + matches[p]];
        const int8_t l = mlabels[p];
        if (l == 0) cls = num_classes;
        if (l == -1) cls = -1;
      } else {
        cls = num_classes;
      }
      prop_classes[p] = cls;
      is_pos = (cls != -1) && (cls != num_classes);
      is_neg = (cls == num_classes);
    }
    const unsigned bp = __ballot_sync(0xffffffffu, is_pos), bn = __ballot_sync(0xffffffffu, is_neg);
    if (lane == 0) {
      warp_pos[warp] = __popc(bp);
      warp_neg[warp] = __popc(bn);
    }
    __syncthreads();
    if (warp == 0) {
      const int vp = lane < nwarp ? warp_pos[lane] : 0, vn = lane < nwarp ? warp_neg[lane] : 0;
      int sp = vp, sn = vn;
      for (int o = 1; o < 32; o <<= 1) {
        const int tp = __shfl_up_sync(0xffffffffu, sp, o), tn = __shfl_up_sync(0xffffffffu, sn, o);
        if (lane >= o) {
          sp += tp;
          sn += tn;
        }
      }
      warp_pos[lane] = sp - vp;  // exclusive prefix over warps
      warp_neg[lane] = sn - vn;
      if (lane == 31) {
        round_pos = sp;
        round_neg = sn;
      }
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1u;
    if (is_pos) pos_idx[p0 + base_pos + warp_pos[warp] + __popc(bp & lt)] = p - p0;
    if (is_neg) neg_idx[p0 + base_neg + warp_neg[warp] + __popc(bn & lt)] = p - p0;
    __syncthreads();
    if (threadIdx.x == 0) {
      base_pos += round_pos;
      base_neg += round_neg;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counts[2 * img] = base_pos;
    counts[2 * img + 1] = base_neg;
  }
}

__global__ void sample_gather_kernel(const int64_t* __restrict__ pos_idx, const int64_t* __restrict__ neg_idx,
                                     const int64_t* __restrict__ perm_pos, const int* __restrict__ perm_pos_off,
                                     const int64_t* __restrict__ perm_neg, const int* __restrict__ perm_neg_off,
                                     const int* __restrict__ pos_sel_off, const int* __restrict__ neg_sel_off,
                                     const int* __restrict__ prop_off, const int* __restrict__ gt_off, int n_img,
                                     int S_total, const float4* __restrict__ prop_boxes,
                                     const int64_t* __restrict__ prop_classes, const int64_t* __restrict__ matches,
                                     const float4* __restrict__ gt_boxes, int64_t* __restrict__ sampled_idx,
                                     float4* __restrict__ out_boxes, int64_t* __restrict__ out_classes,
                                     int64_t* __restrict__ out_matched, float4* __restrict__ out_gt_boxes,
                                     const float* __restrict__ prop_field, float* __restrict__ out_field) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= S_total) return;
  // output rows of image i start at pos_sel_off[i] + neg_sel_off[i]
  int lo = 0, hi = n_img;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pos_sel_off[mid] + neg_sel_off[mid] <= j) lo = mid; else hi = mid;
  }
  const int img = lo;
  const int local = j - (pos_sel_off[img] + neg_sel_off[img]);
  const int npos = pos_sel_off[img + 1] - pos_sel_off[img];
  const int p0 = prop_off[img];
  long long src;
  if (local < npos) src = pos_idx[p0 + perm_pos[perm_pos_off[img] + local]];
  else src = neg_idx[p0 + perm_neg[perm_neg_off[img] + (local - npos)]];
  const long long gp = p0 + src;
  if (sampled_idx) sampled_idx[j] = src;
  if (out_boxes) out_boxes[j] = __ldg(prop_boxes + gp);
  if (out_classes) out_classes[j] = prop_classes[gp];
  if (out_field) out_field[j] = __ldg(prop_field + gp);  // pass-through proposal field (objectness_logits)
  const long long m = matches[gp];
  if (out_matched) out_matched[j] = m;
  if (out_gt_boxes) {
    const int g0 = gt_off[img], ng = gt_off[img + 1] - g0;
    out_gt_boxes[j] = ng > 0 ? __ldg(gt_boxes + g0 + m) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// add_ground_truth_to_proposals for a batch of images in ONE launch: image i's output rows are its proposals followed
// by its ground-truth boxes (objectness logit of a GT box = gt_logit), images back to back -- the concatenated layout
// the labelling kernels read, so the per-image results are views of one buffer and no further cat is needed.
constexpr int APPEND_MAX_IMG = 32;
struct AppendGtArgs {
  const float4* prop[APPEND_MAX_IMG];
  const float* logit[APPEND_MAX_IMG];
  const float4* gt[APPEND_MAX_IMG];
  int off[APPEND_MAX_IMG + 1];  // output row offsets of this launch's images
  int np[APPEND_MAX_IMG];       // proposals per image
  int n_img;
  float gt_logit;
};
__global__ void append_gt_kernel(const AppendGtArgs a, float4* __restrict__ out_boxes, float* __restrict__ out_logits) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.off[a.n_img]) return;
  int lo = 0, hi = a.n_img;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (a.off[mid] <= j) lo = mid; else hi = mid;
  }
  const int local = j - a.off[lo];
  if (local < a.np[lo]) {
    out_boxes[j] = __ldg(a.prop[lo] + local);
    out_logits[j] = __ldg(a.logit[lo] + local);
  } else {
    out_boxes[j] = __ldg(a.gt[lo] + (local - a.np[lo]));
    out_logits[j] = a.gt_logit;
  }
}

}  // namespace match
}  // namespace unit

using namespace unit;
using namespace unit::match;

extern "C" {

int unit_pairwise_iou(const float* boxes1, const float* boxes2, float* iou, int G, int P, unit_stream_t stream) {
  UNIT_REQUIRE(G >= 0 && P >= 0, "pairwise_iou: bad shape");
  if (G == 0 || P == 0) return UNIT_OK;
  UNIT_REQUIRE(boxes1 && boxes2 && iou, "pairwise_iou: null pointer");
  UNIT_REQUIRE((((uintptr_t)boxes1 | (uintptr_t)boxes2) & 15) == 0, "pairwise_iou: boxes must be 16-byte aligned");
  UNIT_REQUIRE(G <= 65535, "pairwise_iou: G too large");
  dim3 grid(cdiv(P, 256), G);
  pairwise_iou_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)boxes1, (const float4*)boxes2, iou, G, P);
  UNIT_CHECK_LAUNCH("pairwise_iou_kernel");
  return UNIT_OK;
}

int unit_matcher(const float* iou, int G, int P, const float* thresholds_host, const int* labels_host, int T,
                 int allow_low_quality_matches, int64_t* matches, int8_t* match_labels, float* matched_vals,
                 void* workspace, size_t workspace_bytes, unit_stream_t stream) {
  UNIT_REQUIRE(G >= 0 && P >= 0, "matcher: bad shape");
  Thresholds t;
  int rc = make_thresholds(thresholds_host, labels_host, T, &t);
  if (rc) return rc;
  if (P == 0) return UNIT_OK;
  UNIT_REQUIRE(matches && match_labels && (G == 0 || iou), "matcher: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  matcher_kernel<<<cdiv(P, 256), 256, 0, st>>>(iou, G, P, t, matches, match_labels, matched_vals);
  UNIT_CHECK_LAUNCH("matcher_kernel");
  if (allow_low_quality_matches && G > 0) {
    if (!workspace || workspace_bytes < (size_t)G * sizeof(float)) {
      set_error("matcher: workspace too small for allow_low_quality_matches (%d floats)", G);
      return UNIT_EWORKSPACE;
    }
    row_max_kernel<<<G, 256, 0, st>>>(iou, G, P, (float*)workspace);
    UNIT_CHECK_LAUNCH("row_max_kernel");
    low_quality_kernel<<<cdiv(P, 256), 256, 0, st>>>(iou, (const float*)workspace, G, P, match_labels);
    UNIT_CHECK_LAUNCH("low_quality_kernel");
  }
  return UNIT_OK;
}

int unit_iou_match(const float* gt_boxes, const int* gt_offsets, const float* prop_boxes, const int* prop_offsets,
                   int n_img, int P_total, const float* thresholds_host, const int* labels_host, int T,
                   int64_t* matches, int8_t* match_labels, float* matched_vals, unit_stream_t stream) {
  UNIT_REQUIRE(n_img >= 0 && P_total >= 0, "iou_match: bad shape");
  Thresholds t;
  int rc = make_thresholds(thresholds_host, labels_host, T, &t);
  if (rc) return rc;
  if (P_total == 0 || n_img == 0) return UNIT_OK;
  UNIT_REQUIRE(gt_offsets && prop_boxes && prop_offsets && matches && match_labels, "iou_match: null pointer");
  UNIT_REQUIRE((((uintptr_t)gt_boxes | (uintptr_t)prop_boxes) & 15) == 0, "iou_match: boxes must be 16-byte aligned");
  iou_match_kernel<<<cdiv(P_total, 128), 128, 0, (cudaStream_t)stream>>>(
      (const float4*)gt_boxes, gt_offsets, (const float4*)prop_boxes, prop_offsets, n_img, P_total, t, matches,
      match_labels, matched_vals);
  UNIT_CHECK_LAUNCH("iou_match_kernel");
  return UNIT_OK;
}

int unit_label_proposals(const int64_t* matches, const int8_t* match_labels, const int64_t* gt_classes,
                         const int* gt_offsets, const int* prop_offsets, int n_img, int P_total, int num_classes,
                         int64_t* prop_classes, int64_t* pos_idx, int64_t* neg_idx, int* counts,
                         unit_stream_t stream) {
  UNIT_REQUIRE(n_img >= 0 && P_total >= 0, "label_proposals: bad shape");
  if (n_img == 0) return UNIT_OK;
  UNIT_REQUIRE(matches && match_labels && gt_offsets && prop_offsets && prop_classes && pos_idx && neg_idx && counts,
               "label_proposals: null pointer");
  label_kernel<<<n_img, 1024, 0, (cudaStream_t)stream>>>(matches, match_labels, gt_classes, gt_offsets, prop_offsets,
                                                         num_classes, prop_classes, pos_idx, neg_idx, counts);
  UNIT_CHECK_LAUNCH("label_kernel");
  return UNIT_OK;
}

int unit_sample_gather(const int64_t* pos_idx, const int64_t* neg_idx, const int64_t* perm_pos,
                       const int* perm_pos_offsets, const int64_t* perm_neg, const int* perm_neg_offsets,
                       const int* pos_sel_offsets, const int* neg_sel_offsets, const int* prop_offsets,
                       const int* gt_offsets, int n_img, int S_total, const float* prop_boxes,
                       const int64_t* prop_classes, const int64_t* matches, const float* gt_boxes,
                       int64_t* sampled_idx, float* out_boxes, int64_t* out_classes, int64_t* out_matched,
                       float* out_gt_boxes, const float* prop_field, float* out_field, unit_stream_t stream) {
  UNIT_REQUIRE(n_img >= 0 && S_total >= 0, "sample_gather: bad shape");
  UNIT_REQUIRE(!out_field || prop_field, "sample_gather: out_field without prop_field");
  if (S_total == 0 || n_img == 0) return UNIT_OK;
  UNIT_REQUIRE(pos_idx && neg_idx && perm_pos_offsets && perm_neg_offsets && pos_sel_offsets && neg_sel_offsets &&
                   prop_offsets && gt_offsets && prop_boxes && prop_classes && matches,
               "sample_gather: null pointer");
  sample_gather_kernel<<<cdiv(S_total, 128), 128, 0, (cudaStream_t)stream>>>(
      pos_idx, neg_idx, perm_pos, perm_pos_offsets, perm_neg, perm_neg_offsets, pos_sel_offsets, neg_sel_offsets,
      prop_offsets, gt_offsets, n_img, S_total, (const float4*)prop_boxes, prop_classes, matches,
      (const float4*)gt_boxes, sampled_idx, (float4*)out_boxes, out_classes, out_matched, (float4*)out_gt_boxes,
      prop_field, out_field);
  UNIT_CHECK_LAUNCH("sample_gather_kernel");
  return UNIT_OK;
}

int unit_append_gt(const float* const* prop_boxes, const float* const* prop_logits, const float* const* gt_boxes,
                   const int* prop_counts, const int* gt_counts, int n_img, float gt_logit, float* out_boxes,
                   float* out_logits, unit_stream_t stream) {
  UNIT_REQUIRE(n_img >= 0, "append_gt: bad shape");
  if (n_img == 0) return UNIT_OK;
  UNIT_REQUIRE(prop_boxes && prop_logits && gt_boxes && prop_counts && gt_counts && out_boxes && out_logits,
               "append_gt: null pointer");
  UNIT_REQUIRE((((uintptr_t)out_boxes) & 15) == 0, "append_gt: out_boxes must be 16-byte aligned");
  long long done = 0;
  for (int i0 = 0; i0 < n_img; i0 += APPEND_MAX_IMG) {
    AppendGtArgs a;
    a.n_img = n_img - i0 < APPEND_MAX_IMG ? n_img - i0 : APPEND_MAX_IMG;
    a.gt_logit = gt_logit;
    a.off[0] = 0;
    for (int k = 0; k < a.n_img; ++k) {
      const int i = i0 + k;
      UNIT_REQUIRE(prop_counts[i] >= 0 && gt_counts[i] >= 0, "append_gt: negative count");
      UNIT_REQUIRE((prop_counts[i] == 0 || (prop_boxes[i] && prop_logits[i])) && (gt_counts[i] == 0 || gt_boxes[i]),
                   "append_gt: null image pointer");
      UNIT_REQUIRE((((uintptr_t)prop_boxes[i] | (uintptr_t)gt_boxes[i]) & 15) == 0,
                   "append_gt: boxes must be 16-byte aligned");
      a.prop[k] = (const float4*)prop_boxes[i];
      a.logit[k] = prop_logits[i];
      a.gt[k] = (const float4*)gt_boxes[i];
      a.np[k] = prop_counts[i];
      a.off[k + 1] = a.off[k] + prop_counts[i] + gt_counts[i];
    }
    const int rows = a.off[a.n_img];
    if (rows > 0) {
      append_gt_kernel<<<cdiv(rows, 256), 256, 0, (cudaStream_t)stream>>>(a, (float4*)out_boxes + done, out_logits + done);
      UNIT_CHECK_LAUNCH("append_gt_kernel");
    }
    done += rows;
  }
  return UNIT_OK;
}

}  // extern "C"
