/* Plain-C restatement of the integer/byte-exact kernels of the UniT RoI stage.  TEST INFRASTRUCTURE (oracle).
 *
 * Independent of PyTorch: used by tests/test_oracle_c.py to cross-check oracle.d2 (torch) and, through it, the
 * CUDA path.  Each function cites what it restates.  Compile: `make -C oracle/c` (gcc -O2 -ffp-contract=off, so
 * every fp32 operation is rounded separately, like the eager PyTorch / torchvision CPU reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* [D2] structures/boxes.py pairwise_iou: inter > 0 ? inter / ((a1 + a2) - inter) : 0 */
void oracle_pairwise_iou(const float* b1, int G, const float* b2, int P, float* iou) {
  for (int g = 0; g < G; ++g) {
    const float* a = b1 + 4 * g;
    const float area1 = (a[2] - a[0]) * (a[3] - a[1]);
    for (int p = 0; p < P; ++p) {
      const float* b = b2 + 4 * p;
      float w = fminf(a[2], b[2]) - fmaxf(a[0], b[0]);
      float h = fminf(a[3], b[3]) - fmaxf(a[1], b[1]);
      if (w < 0) w = 0;
      if (h < 0) h = 0;
      const float inter = w * h;
      const float area2 = (b[2] - b[0]) * (b[3] - b[1]);
      iou[(size_t)g * P + p] = inter > 0 ? inter / ((area1 + area2) - inter) : 0.f;
    }
  }
}

/* modeling/matcher.py:54-98 (allow_low_quality_matches=False): column max, first index on ties, half-open bins */
void oracle_matcher(const float* iou, int G, int P, const float* thr, const int* labels, int T, int64_t* matches,
                    int8_t* out_labels, float* vals) {
  for (int p = 0; p < P; ++p) {
    if (G == 0) {
      matches[p] = 0;
      out_labels[p] = (int8_t)labels[0];
      vals[p] = 0.f;
      continue;
    }
    float best = iou[p];
    int arg = 0;
    for (int g = 1; g < G; ++g)
      if (iou[(size_t)g * P + p] > best) {
        best = iou[(size_t)g * P + p];
        arg = g;
      }
    int8_t lab = 1;
    float low = -INFINITY;
    for (int i = 0; i <= T; ++i) {
      const float high = i < T ? thr[i] : INFINITY;
      if (best >= low && best < high) lab = (int8_t)labels[i];
      low = high;
    }
    matches[p] = arg;
    out_labels[p] = lab;
    vals[p] = best;
  }
}

/* [TV] ops/cpu/nms_kernel.cpp: stable descending sort, greedy, inter / (iarea + area_j - inter) > thr */
static const float* g_scores;
static int cmp_desc_stable(const void* a, const void* b) {
  const int64_t i = *(const int64_t*)a, j = *(const int64_t*)b;
  if (g_scores[i] > g_scores[j]) return -1;
  if (g_scores[i] < g_scores[j]) return 1;
  return i < j ? -1 : (i > j);
}
int oracle_nms(const float* boxes, const float* scores, int N, float thr, int64_t* keep) {
  int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (N > 0 ? N : 1));
  uint8_t* sup = (uint8_t*)calloc(N > 0 ? N : 1, 1);
  float* areas = (float*)malloc(sizeof(float) * (N > 0 ? N : 1));
  for (int i = 0; i < N; ++i) {
    order[i] = i;
    areas[i] = (boxes[4 * i + 2] - boxes[4 * i]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
  }
  g_scores = scores;
  qsort(order, N, sizeof(int64_t), cmp_desc_stable);
  int n_keep = 0;
  for (int _i = 0; _i < N; ++_i) {
    const int64_t i = order[_i];
    if (sup[i]) continue;
    keep[n_keep++] = i;
    const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
    const float iarea = areas[i];
    for (int _j = _i + 1; _j < N; ++_j) {
      const int64_t j = order[_j];
      if (sup[j]) continue;
      const float xx1 = fmaxf(ix1, boxes[4 * j]), yy1 = fmaxf(iy1, boxes[4 * j + 1]);
      const float xx2 = fminf(ix2, boxes[4 * j + 2]), yy2 = fminf(iy2, boxes[4 * j + 3]);
      const float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
      const float inter = w * h;
      const float ovr = inter / (iarea + areas[j] - inter);
      if (ovr > thr) sup[j] = 1;
    }
  }
  free(order);
  free(sup);
  free(areas);
  return n_keep;
}

/* [TV] ops/cpu/roi_align_kernel.cpp (aligned / sampling_ratio semantics of SURVEY.md section 8 row a1), NCHW fp32 */
void oracle_roi_align_fwd(const float* feat, int N, int C, int H, int W, const float* rois, int R, int PH, int PW,
                          float scale, int sampling_ratio, int aligned, float* out) {
  for (int r = 0; r < R; ++r) {
    const float* roi = rois + 5 * r;
    const int n = (int)roi[0];
    const float off = aligned ? 0.5f : 0.f;
    const float sw = roi[1] * scale - off, sh = roi[2] * scale - off;
    const float ew = roi[3] * scale - off, eh = roi[4] * scale - off;
    float rw = ew - sw, rh = eh - sh;
    if (!aligned) {
      rw = fmaxf(rw, 1.f);
      rh = fmaxf(rh, 1.f);
    }
    const float bh = rh / (float)PH, bw = rw / (float)PW;
    const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / PH);
    const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / PW);
    const float count = (float)(gh * gw > 1 ? gh * gw : 1);
    for (int c = 0; c < C; ++c) {
      const float* plane = feat + ((size_t)n * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float acc = 0.f;
          for (int iy = 0; iy < gh; ++iy) {
            float y = sh + ph * bh + (iy + .5f) * bh / (float)gh;
            for (int ix = 0; ix < gw; ++ix) {
              float x = sw + pw * bw + (ix + .5f) * bw / (float)gw;
              float yy = y;
              if (yy < -1.0f || yy > H || x < -1.0f || x > W) continue;
              if (yy <= 0) yy = 0;
              if (x <= 0) x = 0;
              int yl = (int)yy, xl = (int)x, yh, xh;
              if (yl >= H - 1) { yh = yl = H - 1; yy = (float)yl; } else yh = yl + 1;
              if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
              const float ly = yy - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
              const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
              acc += w1 * plane[yl * W + xl] + w2 * plane[yl * W + xh] + w3 * plane[yh * W + xl] +
                     w4 * plane[yh * W + xh];
            }
          }
          out[(((size_t)r * C + c) * PH + ph) * PW + pw] = acc / count;
        }
    }
  }
}

/* weak_detector_fast_rcnn.py:353-396 for ONE image: get_proposal_clusters (for each class of `classes` -- sorted,
 * unique -- the proposal with the largest probs[r][c] among rows not picked yet, first row on ties; a picked row
 * counts as all-zero afterwards), then pairwise_iou(picked boxes, proposals) + Matcher + the label / loss-weight
 * rules of label_and_sample_proposals and compute_loss_inputs.  probs [R][ld]; out: labels [R], weights [R],
 * picked [G] (row index per class). */
void oracle_oicr_targets(const float* probs, int ld, const float* boxes, int R, const int64_t* classes, int G,
                         const float* thr, const int* match_labels, int T, float bg_threshold, int num_classes,
                         int64_t* labels, float* weights, int64_t* picked) {
  if (R == 0) return;
  unsigned char* used = (unsigned char*)calloc((size_t)R, 1);
  float* score = (float*)malloc(sizeof(float) * (size_t)(G > 0 ? G : 1));
  float* gbox = (float*)malloc(sizeof(float) * 4 * (size_t)(G > 0 ? G : 1));
  for (int g = 0; g < G; ++g) {
    const int c = (int)classes[g];
    float best = used[0] ? 0.f : probs[c];
    int arg = 0;
    for (int r = 1; r < R; ++r) {
      const float v = used[r] ? 0.f : probs[(size_t)r * ld + c];
      if (v > best) {
        best = v;
        arg = r;
      }
    }
    used[arg] = 1;
    picked[g] = arg;
    score[g] = best;
    memcpy(gbox + 4 * g, boxes + 4 * arg, 4 * sizeof(float));
  }
  float* iou = (float*)malloc(sizeof(float) * (size_t)(G > 0 ? G : 1) * (size_t)R);
  int64_t* m = (int64_t*)malloc(sizeof(int64_t) * (size_t)R);
  int8_t* l = (int8_t*)malloc((size_t)R);
  float* v = (float*)malloc(sizeof(float) * (size_t)R);
  oracle_pairwise_iou(gbox, G, boxes, R, iou);
  oracle_matcher(iou, G, R, thr, match_labels, T, m, l, v);
  for (int r = 0; r < R; ++r) {
    if (G == 0) {
      labels[r] = num_classes;
      weights[r] = 0.f;
      continue;
    }
    labels[r] = l[r] == 1 ? classes[m[r]] : (l[r] == 0 ? num_classes : -1);
    weights[r] = (bg_threshold > 0.f && v[r] < bg_threshold) ? 0.f : score[m[r]];
  }
  free(used); free(score); free(gbox); free(iou); free(m); free(l); free(v);
}
