// Matcher pieces shared by match.cu and weak.cu: threshold bins (modeling/matcher.py:54-119) and segment lookup.
#pragma once
#include "common.cuh"

namespace unit {
namespace match {

constexpr int MAX_T = 8;

struct Thresholds {
  float thr[MAX_T];       // user thresholds, ascending
  int labels[MAX_T + 1];  // label of each bin
  int T;
};

__device__ __forceinline__ int8_t bin_label(float v, const Thresholds& t) {
  // modeling/matcher.py:89-91: for (l, low, high): labels[(v >= low) & (v < high)] = l ; default 1 (NaN)
  int8_t lab = 1;
  float low = -INFINITY;
  for (int i = 0; i <= t.T; ++i) {
    const float high = i < t.T ? t.thr[i] : INFINITY;
    if (v >= low && v < high) lab = (int8_t)t.labels[i];
    low = high;
  }
  return lab;
}

__device__ __forceinline__ int find_segment(const int* __restrict__ off, int n, int i) {
  int lo = 0, hi = n;  // largest s with off[s] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

static inline int make_thresholds(const float* thr, const int* labels, int T, Thresholds* out) {
  UNIT_REQUIRE(T >= 1 && T <= MAX_T, "matcher: 1..%d thresholds supported, got %d", MAX_T, T);
  UNIT_REQUIRE(thr && labels, "matcher: null thresholds/labels");
  out->T = T;
  for (int i = 0; i < T; ++i) out->thr[i] = thr[i];
  for (int i = 0; i <= T; ++i) {
    UNIT_REQUIRE(labels[i] >= -1 && labels[i] <= 1, "matcher: labels must be in {-1,0,1}");
    out->labels[i] = labels[i];
  }
  for (int i = 1; i < T; ++i) UNIT_REQUIRE(thr[i - 1] <= thr[i], "matcher: thresholds must be ascending");
  UNIT_REQUIRE(thr[0] > 0, "matcher: thresholds[0] must be > 0");
  return UNIT_OK;
}

}  // namespace match
}  // namespace unit
