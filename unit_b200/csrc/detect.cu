// softmax + box decode, score filter, class-wise batched NMS.
//
// fast_rcnn_inference for ALL images of a batch takes three launches (decode, filter, NMS) instead of the
// reference's per-image Python loop of ~40 ATen kernels with host syncs.  The keep list is bit-exact: candidates
// are emitted in the reference's row-major (roi, class) order, the sort is (score desc, index asc), IoUs are
// computed with separately rounded fp32 operations in torchvision's order and compared with a strict `>`.
//
// NMS layout: one CTA per image.  (1) bitonic sort of 64-bit keys (score | index) in shared memory, (2) a second
// bitonic sort groups the ranked boxes by class, (3) every class segment is reduced greedily by one warp: 32 boxes
// per step, suppression against already kept boxes, then the 32x32 intra-step matrix resolved with
// __ballot_sync / __shfl_sync, (4) kept boxes are compacted back into score order by rank.  The suppression
// bitmask never leaves the SM (torchvision writes an N x N/64 mask to HBM and finishes on a single block).
#include <algorithm>
#include <math.h>

#include "common.cuh"

namespace unit {
namespace detect {

// ---------------------------------------------------------------------------------------- softmax + decode
__global__ void softmax_decode_kernel(const float* __restrict__ scores, const float* __restrict__ deltas,
                                      const float4* __restrict__ proposals, float* __restrict__ probs,
                                      float* __restrict__ boxes, int R, int K1, int KB, float wx, float wy, float ww,
                                      float wh, float clampv) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const int r = warp;
  if (probs) {
    const float* s = scores + (long long)r * K1;
    float m = -INFINITY;
    for (int k = lane; k < K1; k += 32) m = fmaxf(m, s[k]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int k = lane; k < K1; k += 32) sum += expf(s[k] - m);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    for (int k = lane; k < K1; k += 32) probs[(long long)r * K1 + k] = __fdiv_rn(expf(s[k] - m), sum);
  }
  if (boxes) {
    // [D2] Box2BoxTransform.apply_deltas, each op rounded separately like the eager PyTorch reference
    const float4 p = __ldg(proposals + r);
    const float w = __fsub_rn(p.z, p.x), h = __fsub_rn(p.w, p.y);
    const float cx = __fadd_rn(p.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(p.y, __fmul_rn(0.5f, h));
    const float4* d4 = reinterpret_cast<const float4*>(deltas + (long long)r * KB * 4);
    float4* o4 = reinterpret_cast<float4*>(boxes + (long long)r * KB * 4);
    for (int k = lane; k < KB; k += 32) {
      const float4 d = d4[k];
      const float dx = __fdiv_rn(d.x, wx), dy = __fdiv_rn(d.y, wy);
      const float dw = fminf(__fdiv_rn(d.z, ww), clampv), dh = fminf(__fdiv_rn(d.w, wh), clampv);
      const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
      const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
      float4 o;
      o.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
      o.y = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
      o.z = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
      o.w = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
      o4[k] = o;
    }
  }
}

__global__ void get_deltas_kernel(const float4* __restrict__ src, const float4* __restrict__ tgt,
                                  float4* __restrict__ out, int R, float wx, float wy, float ww, float wh) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float4 s = __ldg(src + r), t = __ldg(tgt + r);
  const float sw = __fsub_rn(s.z, s.x), sh = __fsub_rn(s.w, s.y);
  const float scx = __fadd_rn(s.x, __fmul_rn(0.5f, sw)), scy = __fadd_rn(s.y, __fmul_rn(0.5f, sh));
  const float tw = __fsub_rn(t.z, t.x), th = __fsub_rn(t.w, t.y);
  const float tcx = __fadd_rn(t.x, __fmul_rn(0.5f, tw)), tcy = __fadd_rn(t.y, __fmul_rn(0.5f, th));
  float4 o;
  o.x = __fdiv_rn(__fmul_rn(wx, __fsub_rn(tcx, scx)), sw);
  o.y = __fdiv_rn(__fmul_rn(wy, __fsub_rn(tcy, scy)), sh);
  o.z = __fmul_rn(ww, logf(__fdiv_rn(tw, sw)));
  o.w = __fmul_rn(wh, logf(__fdiv_rn(th, sh)));
  out[r] = o;
}

// ---------------------------------------------------------------------------------------- candidate filter
constexpr int FT = 1024;  // threads per image CTA

__device__ __forceinline__ bool finite_f(float v) { return fabsf(v) <= 3.402823466e+38f; }  // false for inf / NaN

__global__ void __launch_bounds__(FT) detect_filter_kernel(const float* __restrict__ boxes,
                                                           const float* __restrict__ probs,
                                                           const int* __restrict__ roi_off,
                                                           const float* __restrict__ image_hw, int K, int KB,
                                                           float thresh, float4* __restrict__ cand_boxes,
                                                           float* __restrict__ cand_scores, int* __restrict__ cand_roi,
                                                           int* __restrict__ cand_cls, int* __restrict__ cand_counts) {
  const int img = blockIdx.x;
  const int r0 = roi_off[img], r1 = roi_off[img + 1];
  const float img_h = image_hw[2 * img], img_w = image_hw[2 * img + 1];
  const long long out0 = (long long)r0 * K;
  __shared__ int s_valid[FT], s_cnt[FT], s_wsum_v[32], s_wsum_c[32];
  __shared__ int base_rank, base_cnt, round_v, round_c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    base_rank = 0;
    base_cnt = 0;
  }
  __syncthreads();
  const int K1 = K + 1;
  for (int chunk = r0; chunk < r1; chunk += FT) {
    const int rows = min(FT, r1 - chunk);
    // (a) per-row validity and candidate count: warp w takes rows w, w+32, ...
    for (int rl = warp; rl < rows; rl += 32) {
      const int r = chunk + rl;
      bool ok = true;
      const float* b = boxes + (long long)r * KB * 4;
      const float* s = probs + (long long)r * K1;
      for (int i = lane; i < KB * 4; i += 32) ok &= finite_f(b[i]);
      for (int i = lane; i < K1; i += 32) ok &= finite_f(s[i]);
      ok = __all_sync(0xffffffffu, ok);
      int cnt = 0;
      if (ok)
        for (int k0 = 0; k0 < K; k0 += 32) {
          const int k = k0 + lane;
          cnt += __popc(__ballot_sync(0xffffffffu, k < K && s[k] > thresh));
        }
      if (lane == 0) {
        s_valid[rl] = ok ? 1 : 0;
        s_cnt[rl] = cnt;
      }
    }
    __syncthreads();
    // (b) block-wide exclusive scan of (valid, cnt) over the rows of the chunk (thread t <-> row t)
    const int v = tid < rows ? s_valid[tid] : 0, c = tid < rows ? s_cnt[tid] : 0;
    int sv = v, sc = c;
    for (int o = 1; o < 32; o <<= 1) {
      const int tv = __shfl_up_sync(0xffffffffu, sv, o), tc = __shfl_up_sync(0xffffffffu, sc, o);
      if (lane >= o) {
        sv += tv;
        sc += tc;
      }
    }
    if (lane == 31) {
      s_wsum_v[warp] = sv;
      s_wsum_c[warp] = sc;
    }
    __syncthreads();
    if (warp == 0) {
      const int wv = s_wsum_v[lane], wc = s_wsum_c[lane];
      int av = wv, ac = wc;
      for (int o = 1; o < 32; o <<= 1) {
        const int tv = __shfl_up_sync(0xffffffffu, av, o), tc = __shfl_up_sync(0xffffffffu, ac, o);
        if (lane >= o) {
          av += tv;
          ac += tc;
        }
      }
      s_wsum_v[lane] = av - wv;
      s_wsum_c[lane] = ac - wc;
      if (lane == 31) {
        round_v = av;
        round_c = ac;
      }
    }
    __syncthreads();
    const int ex_rank = base_rank + s_wsum_v[warp] + sv - v;  // rank of row among valid rows
    const int ex_cnt = base_cnt + s_wsum_c[warp] + sc - c;    // first candidate slot of row
    __syncthreads();
    if (tid < rows) {
      s_valid[tid] = v ? ex_rank : -1;
      s_cnt[tid] = ex_cnt;
    }
    __syncthreads();
    // (c) emit candidates in (row, class) order
    for (int rl = warp; rl < rows; rl += 32) {
      const int rank = s_valid[rl];
      if (rank < 0) continue;
      const int r = chunk + rl;
      const float* s = probs + (long long)r * K1;
      int pos = s_cnt[rl];
      for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        const float sc_k = k < K ? s[k] : 0.f;
        const bool take = k < K && sc_k > thresh;
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        if (take) {
          const long long o = out0 + pos + __popc(bal & ((1u << lane) - 1u));
          const float4 bx = reinterpret_cast<const float4*>(boxes + (long long)r * KB * 4)[KB == 1 ? 0 : k];
          float4 cb;  // [D2] Boxes.clip: x in [0,w], y in [0,h]
          cb.x = fminf(fmaxf(bx.x, 0.f), img_w);
          cb.y = fminf(fmaxf(bx.y, 0.f), img_h);
          cb.z = fminf(fmaxf(bx.z, 0.f), img_w);
          cb.w = fminf(fmaxf(bx.w, 0.f), img_h);
          cand_boxes[o] = cb;
          cand_scores[o] = sc_k;
          cand_roi[o] = rank;
          cand_cls[o] = k;
        }
        pos += __popc(bal);
      }
    }
    __syncthreads();
    if (tid == 0) {
      base_rank += round_v;
      base_cnt += round_c;
    }
    __syncthreads();
  }
  if (tid == 0) cand_counts[img] = base_cnt;
}

// ---------------------------------------------------------------------------------------- parallel candidate filter
// Same result as detect_filter_kernel (candidates in row-major (roi, class) order, roi = rank among the image's finite
// rows), but spread over the GPU: (1) one warp per row -> validity and candidate count, (2) one CTA per image -> exclusive
// scans over its rows, (3) one warp per row -> emit at the scanned position.  Row arrays live in the caller's workspace.
constexpr int FW = 8;  // warps (rows) per CTA

__global__ void __launch_bounds__(FW * 32)
filter_count_kernel(const float* __restrict__ boxes, const float* __restrict__ probs, int R, int K, int KB, float thresh,
                    int* __restrict__ row_valid, int* __restrict__ row_cnt) {
  const int r = blockIdx.x * FW + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  const int K1 = K + 1;
  const float* b = boxes + (long long)r * KB * 4;
  const float* s = probs + (long long)r * K1;
  bool ok = true;
  for (int i = lane; i < KB * 4; i += 32) ok &= finite_f(b[i]);
  for (int i = lane; i < K1; i += 32) ok &= finite_f(s[i]);
  ok = __all_sync(0xffffffffu, ok);
  int cnt = 0;
  if (ok)
    for (int k0 = 0; k0 < K; k0 += 32) {
      const int k = k0 + lane;
      cnt += __popc(__ballot_sync(0xffffffffu, k < K && s[k] > thresh));
    }
  if (lane == 0) {
    row_valid[r] = ok ? 1 : 0;
    row_cnt[r] = cnt;
  }
}

// per image: row_valid -> rank among valid rows (or -1), row_cnt -> first candidate slot; cand_counts[img] = total
__global__ void __launch_bounds__(1024)
filter_scan_kernel(const int* __restrict__ roi_off, int* __restrict__ row_valid, int* __restrict__ row_cnt,
                   int* __restrict__ cand_counts) {
  __shared__ int s_wv[32], s_wc[32], base_v, base_c, round_v, round_c;
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = roi_off[img], r1 = roi_off[img + 1];
  if (tid == 0) {
    base_v = 0;
    base_c = 0;
  }
  __syncthreads();
  for (int chunk = r0; chunk < r1; chunk += 1024) {
    const int r = chunk + tid;
    const int v = r < r1 ? row_valid[r] : 0, c = r < r1 ? row_cnt[r] : 0;
    int sv = v, sc = c;
    for (int o = 1; o < 32; o <<= 1) {
      const int tv = __shfl_up_sync(0xffffffffu, sv, o), tc = __shfl_up_sync(0xffffffffu, sc, o);
      if (lane >= o) {
        sv += tv;
        sc += tc;
      }
    }
    if (lane == 31) {
      s_wv[warp] = sv;
      s_wc[warp] = sc;
    }
    __syncthreads();
    if (warp == 0) {
      const int wv = s_wv[lane], wc = s_wc[lane];
      int av = wv, ac = wc;
      for (int o = 1; o < 32; o <<= 1) {
        const int tv = __shfl_up_sync(0xffffffffu, av, o), tc = __shfl_up_sync(0xffffffffu, ac, o);
        if (lane >= o) {
          av += tv;
          ac += tc;
        }
      }
      s_wv[lane] = av - wv;
      s_wc[lane] = ac - wc;
      if (lane == 31) {
        round_v = av;
        round_c = ac;
      }
    }
    __syncthreads();
    if (r < r1) {
      row_valid[r] = v ? base_v + s_wv[warp] + sv - v : -1;
      row_cnt[r] = base_c + s_wc[warp] + sc - c;
    }
    __syncthreads();
    if (tid == 0) {
      base_v += round_v;
      base_c += round_c;
    }
    __syncthreads();
  }
  if (tid == 0) cand_counts[img] = base_c;
}

__global__ void __launch_bounds__(FW * 32)
filter_emit_kernel(const float* __restrict__ boxes, const float* __restrict__ probs, const int* __restrict__ roi_off,
                   const float* __restrict__ image_hw, int n_img, int R, int K, int KB, float thresh,
                   const int* __restrict__ row_rank, const int* __restrict__ row_pos, float4* __restrict__ cand_boxes,
                   float* __restrict__ cand_scores, int* __restrict__ cand_roi, int* __restrict__ cand_cls) {
  const int r = blockIdx.x * FW + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  const int rank = row_rank[r];
  if (rank < 0) return;
  int img = 0;  // the image of row r (n_img is small: a linear walk of the offsets)
  while (img + 1 < n_img && roi_off[img + 1] <= r) ++img;
  const float img_h = image_hw[2 * img], img_w = image_hw[2 * img + 1];
  const long long out0 = (long long)roi_off[img] * K;
  const float* s = probs + (long long)r * (K + 1);
  int pos = row_pos[r];
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    const float sc_k = k < K ? s[k] : 0.f;
    const bool take = k < K && sc_k > thresh;
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    if (take) {
      const long long o = out0 + pos + __popc(bal & ((1u << lane) - 1u));
      const float4 bx = reinterpret_cast<const float4*>(boxes + (long long)r * KB * 4)[KB == 1 ? 0 : k];
      float4 cb;  // [D2] Boxes.clip: x in [0,w], y in [0,h]
      cb.x = fminf(fmaxf(bx.x, 0.f), img_w);
      cb.y = fminf(fmaxf(bx.y, 0.f), img_h);
      cb.z = fminf(fmaxf(bx.z, 0.f), img_w);
      cb.w = fminf(fmaxf(bx.w, 0.f), img_h);
      cand_boxes[o] = cb;
      cand_scores[o] = sc_k;
      cand_roi[o] = rank;
      cand_cls[o] = k;
    }
    pos += __popc(bal);
  }
}

// ---------------------------------------------------------------------------------------- segmented NMS
constexpr int NT = 1024;
constexpr int SMEM_CAP = 4096;  // elements handled entirely in shared memory

struct NmsParams {
  const float4* boxes;
  const float* scores;
  const void* cls;  // int32 or int64 class ids, NULL = single class
  int cls_i64;
  const int* seg_counts;  // [n_seg] or NULL (then n_single)
  const int* seg_base;    // element offset of segment s = seg_base[s] * seg_mul (NULL -> 0)
  int seg_mul;
  int n_single;
  float thr;
  int mode;  // 0 class-wise raw, 1 coordinate trick, 2 torchvision CUDA rule, 3 torchvision CPU rule, 4 plain
  int max_keep;
  int64_t* keep;  // generic API output (segment-local indices), stride keep_stride per segment
  int keep_stride;
  int* keep_counts;
  float4* det_boxes;  // detection outputs, stride det_stride per segment
  float* det_scores;
  int64_t* det_classes;
  int64_t* det_roi;
  const int* cand_roi;
  int det_stride;
  const unsigned char* seg_enable;  // optional: segments whose byte is 0 are skipped (handled by the grouped path)
  unsigned long long* ws_key1;  // global fallback arrays, indexed from 2 * segment offset
  unsigned long long* ws_key2;
  float4* ws_box;
  int* ws_flag;
};

__device__ __forceinline__ unsigned ordered_u32(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone: larger float -> larger unsigned
}

__device__ void bitonic_sort(unsigned long long* key, int n_pad) {
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pad; i += NT) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = key[i], b = key[ixj];
          const bool asc = (i & k) == 0;
          if ((a > b) == asc) {
            key[i] = b;
            key[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
  // all NT threads participate; returns exclusive prefix of v, total = sum
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int s = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, s, o);
    if (lane >= o) s += t;
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = s;
  __syncthreads();
  if (warp == 0) {
    const int w = s_warp[lane];
    int a = w;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, a, o);
      if (lane >= o) a += t;
    }
    s_warp[lane] = a - w;
    if (lane == 31) s_warp[32] = a;
  }
  __syncthreads();
  total = s_warp[32];
  return s_warp[warp] + s - v;
}

__device__ __forceinline__ int cls_at(const NmsParams& p, long long i) {
  if (!p.cls) return 0;
  return p.cls_i64 ? (int)reinterpret_cast<const long long*>(p.cls)[i] : reinterpret_cast<const int*>(p.cls)[i];
}

__global__ void __launch_bounds__(NT, 1) segmented_nms_kernel(const NmsParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_warp[33];
  __shared__ float s_red[64];

  const int seg = blockIdx.x;
  if (p.seg_enable && !p.seg_enable[seg]) return;
  const int n = p.seg_counts ? p.seg_counts[seg] : p.n_single;
  const long long base = p.seg_base ? (long long)p.seg_base[seg] * p.seg_mul : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (n <= 0) {
    if (tid == 0 && p.keep_counts) p.keep_counts[seg] = 0;
    return;
  }
  int n_pad = 1;
  while (n_pad < n) n_pad <<= 1;

  unsigned long long *key1, *key2;
  float4* sbox;
  int* flag;
  if (n_pad <= SMEM_CAP) {
    key1 = reinterpret_cast<unsigned long long*>(smem_raw);
    key2 = key1 + SMEM_CAP;
    sbox = reinterpret_cast<float4*>(key2 + SMEM_CAP);
    flag = reinterpret_cast<int*>(sbox + SMEM_CAP);
  } else {
    key1 = p.ws_key1 + 2 * base;
    key2 = p.ws_key2 + 2 * base;
    sbox = p.ws_box + 2 * base;
    flag = p.ws_flag + 2 * base;
  }
  const float4* boxes = p.boxes + base;
  const float* scores = p.scores + base;

  // mode resolution (torchvision ops/boxes.py:80)
  int mode = p.mode;
  if (mode == 2) mode = (4LL * n > 100000) ? 0 : 1;
  if (mode == 3) mode = (4LL * n > 4000) ? 0 : 1;
  if (!p.cls) mode = 4;

  // coordinate trick: offsets = cls * (max_coordinate + 1)
  float off_unit = 0.f;
  bool by_class = (mode == 0 || mode == 1);
  if (mode == 1) {
    float mx = -INFINITY, mn = INFINITY;
    for (int i = tid; i < n; i += NT) {
      const float4 b = boxes[i];
      mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
      mn = fminf(mn, fminf(fminf(b.x, b.y), fminf(b.z, b.w)));
    }
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) {
      s_red[warp] = mx;
      s_red[32 + warp] = mn;
    }
    __syncthreads();
    mx = s_red[0];
    mn = s_red[32];
    for (int w = 1; w < 32; ++w) {
      mx = fmaxf(mx, s_red[w]);
      mn = fminf(mn, s_red[32 + w]);
    }
    off_unit = __fadd_rn(mx, 1.f);
    // boxes of different classes cannot overlap after the shift iff min >= -1 (gap between class ranges = min+1);
    // otherwise fall back to the literal all-pairs formulation.
    if (!(mn >= -1.f) || !(p.thr >= 0.f)) by_class = false;
  }

  // (1) sort by (score desc, index asc)
  for (int i = tid; i < n_pad; i += NT) {
    unsigned long long k = ~0ull;
    if (i < n) k = ((unsigned long long)(~ordered_u32(scores[i])) << 32) | (unsigned)i;
    key1[i] = k;
  }
  __syncthreads();
  bitonic_sort(key1, n_pad);

  // (2) group by class keeping score order: key2 = cls << 32 | rank
  for (int i = tid; i < n_pad; i += NT) {
    unsigned long long k = ~0ull;
    if (i < n) {
      const unsigned idx = (unsigned)(key1[i] & 0xffffffffu);
      const unsigned c = by_class ? (unsigned)cls_at(p, base + idx) : 0u;
      k = ((unsigned long long)c << 32) | (unsigned)i;
    }
    key2[i] = k;
  }
  __syncthreads();
  if (by_class) bitonic_sort(key2, n_pad);

  // gather (shifted) boxes into class-grouped order
  for (int j = tid; j < n; j += NT) {
    const unsigned rank = (unsigned)(key2[j] & 0xffffffffu);
    const unsigned idx = (unsigned)(key1[rank] & 0xffffffffu);
    float4 b = boxes[idx];
    if (mode == 1) {
      const float o = __fmul_rn((float)cls_at(p, base + idx), off_unit);
      b.x = __fadd_rn(b.x, o);
      b.y = __fadd_rn(b.y, o);
      b.z = __fadd_rn(b.z, o);
      b.w = __fadd_rn(b.w, o);
    }
    sbox[j] = b;
    flag[j] = 0;
  }
  __syncthreads();

  // (3) greedy NMS per class segment.  Every warp walks the segment-start flags of the whole array (cheap) and
  // processes the segments whose ordinal is congruent to its warp index.
  {
    int seg_ord = 0;  // running ordinal (identical in all lanes/warps since every warp scans the same flags)
    for (int j0 = 0; j0 < n; j0 += 32) {
      const int j = j0 + lane;
      bool start = false;
      if (j < n) start = (j == 0) || ((key2[j] >> 32) != (key2[j - 1] >> 32));
      unsigned sm = __ballot_sync(0xffffffffu, start);
      while (sm) {
        const int l = __ffs(sm) - 1;
        sm &= sm - 1;
        const int s = j0 + l;
        if ((seg_ord & 31) == warp) {
          // find the end of this class segment
          const unsigned long long c = key2[s] >> 32;
          int e = s + 1;
          // gallop: the segment is contiguous
          {
            int lo = s, hi = n;  // last index with class c is in [lo, hi)
            while (hi - lo > 1) {
              const int mid = (lo + hi) >> 1;
              if ((key2[mid] >> 32) == c) lo = mid; else hi = mid;
            }
            e = lo + 1;
          }
          // greedy NMS of [s, e) by this warp
          for (int c0 = s; c0 < e; c0 += 32) {
            const int q = c0 + lane;
            const bool valid = q < e;
            float4 my = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) my = sbox[q];
            const float my_area = box_area_rn(my.x, my.y, my.z, my.w);
            bool alive = valid;
            // against boxes kept in earlier steps of this segment
            for (int t0 = s; t0 < c0; t0 += 32) {
              unsigned km = (unsigned)flag[t0 + lane];  // every earlier step is a full 32-wide step
              km = __ballot_sync(0xffffffffu, km != 0);
              while (km) {
                const int kl = __ffs(km) - 1;
                km &= km - 1;
                const float4 kb = sbox[t0 + kl];
                const float inter = box_inter_rn(kb.x, kb.y, kb.z, kb.w, my.x, my.y, my.z, my.w);
                const float iou = iou_from_rn(inter, box_area_rn(kb.x, kb.y, kb.z, kb.w), my_area);
                if (iou > p.thr) alive = false;
              }
              if (!__any_sync(0xffffffffu, alive)) break;
            }
            // inside the step
            for (int l2 = 0; l2 < 32; ++l2) {
              const unsigned am = __ballot_sync(0xffffffffu, alive);
              if (!((am >> l2) & 1u)) continue;
              if ((am >> l2) <= 1u) break;  // no alive lane above l2
              const float kx = __shfl_sync(0xffffffffu, my.x, l2), ky = __shfl_sync(0xffffffffu, my.y, l2);
              const float kz = __shfl_sync(0xffffffffu, my.z, l2), kw = __shfl_sync(0xffffffffu, my.w, l2);
              const float ka = __shfl_sync(0xffffffffu, my_area, l2);
              if (alive && lane > l2) {
                const float inter = box_inter_rn(kx, ky, kz, kw, my.x, my.y, my.z, my.w);
                if (iou_from_rn(inter, ka, my_area) > p.thr) alive = false;
              }
            }
            if (valid) flag[q] = alive ? 1 : 0;
            __syncwarp();
          }
        }
        ++seg_ord;
      }
    }
  }
  __syncthreads();

  // (4) back to score order: mark kept ranks (key2's low word is the rank), then compact in rank order
  int* mark = reinterpret_cast<int*>(sbox);  // the shifted boxes are dead from here on; n ints fit in n float4
  for (int j = tid; j < n; j += NT) mark[j] = 0;
  __syncthreads();
  for (int j = tid; j < n; j += NT)
    if (flag[j]) mark[(unsigned)(key2[j] & 0xffffffffu)] = 1;
  __syncthreads();

  int kept_before = 0;
  const int limit = p.max_keep >= 0 ? p.max_keep : n;
  for (int r0 = 0; r0 < n && kept_before < limit; r0 += NT) {
    const int r = r0 + tid;
    const int m = r < n ? mark[r] : 0;
    int total;
    const int ex = block_exclusive_scan(m, s_warp, total);
    const int pos = kept_before + ex;
    if (m && pos < limit) {
      const unsigned idx = (unsigned)(key1[r] & 0xffffffffu);
      if (p.keep) p.keep[(long long)seg * p.keep_stride + pos] = idx;
      if (p.det_boxes) {
        const long long o = (long long)seg * p.det_stride + pos;
        p.det_boxes[o] = boxes[idx];
        p.det_scores[o] = scores[idx];
        p.det_classes[o] = cls_at(p, base + idx);
        p.det_roi[o] = p.cand_roi[base + idx];
      }
    }
    kept_before += total;
    __syncthreads();
  }
  if (tid == 0 && p.keep_counts) p.keep_counts[seg] = min(kept_before, limit);
}

// ---------------------------------------------------------------------------------------- grouped NMS (detections)
// fast_rcnn_inference runs NMS class by class, so one image does not have to be one CTA: NG CTAs per image each take
// the classes c with c % NG == g, sort only their candidates by (class, score desc, index asc), run the same greedy
// reduction (same fp32 IoU arithmetic, same coordinate-trick shift, so the keep set is bit-identical) and mark the
// survivors; a second kernel per image sorts the survivors by (score desc, index asc) and writes the top-k.  Images
// that do not fit the grouped path (all-pairs fallback of the coordinate trick, a group over GCAP, > TOPCAP
// candidates) are flagged and handled by segmented_nms_kernel.
constexpr int NG = 8;
constexpr int GT = 512;
constexpr int GCAP = 4096;
constexpr int TOPCAP = 16384;

struct GroupParams {
  const float4* boxes;
  const float* scores;
  const int* cls;
  const int* counts;
  const int* seg_base;
  int seg_mul;
  float thr;
  int mode;
  unsigned char* keep_flag;  // [total candidates]
  unsigned char* use_old;    // [n_img]
  // top-k outputs
  int max_keep, det_stride;
  float4* det_boxes;
  float* det_scores;
  int64_t* det_classes;
  int64_t* det_roi;
  const int* cand_roi;
  int* det_counts;
};

template <int THREADS>
__device__ void bitonic_sort_t(unsigned long long* key, int n_pad) {
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pad; i += THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = key[i], b = key[ixj];
          const bool asc = (i & k) == 0;
          if ((a > b) == asc) {
            key[i] = b;
            key[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(GT) nms_group_kernel(const GroupParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  float4* sbox = reinterpret_cast<float4*>(keys + GCAP);
  int* flag = reinterpret_cast<int*>(sbox + GCAP);
  __shared__ int s_cnt[NG];
  __shared__ int s_m;
  __shared__ float s_red[2 * (GT / 32)];
  const int g = blockIdx.x, img = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.counts[img];
  if (n <= 0) return;
  const long long base = (long long)p.seg_base[img] * p.seg_mul;
  const float4* boxes = p.boxes + base;
  const float* scores = p.scores + base;
  const int* cls = p.cls + base;
  if (tid < NG) s_cnt[tid] = 0;
  if (tid == 0) s_m = 0;
  __syncthreads();

  // group sizes (every CTA of the image computes all of them: one consistent decision) and the coordinate range
  float mx = -INFINITY, mn = INFINITY;
  for (int i = tid; i < n; i += GT) {
    atomicAdd(&s_cnt[cls[i] % NG], 1);
    const float4 b = boxes[i];
    mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
    mn = fminf(mn, fminf(fminf(b.x, b.y), fminf(b.z, b.w)));
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if (lane == 0) {
    s_red[warp] = mx;
    s_red[GT / 32 + warp] = mn;
  }
  __syncthreads();
  mx = s_red[0];
  mn = s_red[GT / 32];
  for (int w = 1; w < GT / 32; ++w) {
    mx = fmaxf(mx, s_red[w]);
    mn = fminf(mn, s_red[GT / 32 + w]);
  }
  int mode = p.mode;  // torchvision ops/boxes.py:80
  if (mode == 2) mode = (4LL * n > 100000) ? 0 : 1;
  if (mode == 3) mode = (4LL * n > 4000) ? 0 : 1;
  const float off_unit = __fadd_rn(mx, 1.f);
  bool old = n > TOPCAP;
  if (mode == 1 && (!(mn >= -1.f) || !(p.thr >= 0.f))) old = true;  // literal all-pairs formulation
  for (int q = 0; q < NG; ++q) old |= s_cnt[q] > GCAP;
  if (tid == 0) p.use_old[img] = old ? 1 : 0;
  if (old) return;
  const int m = s_cnt[g];
  if (m == 0) return;
  int m_pad = 1;
  while (m_pad < m) m_pad <<= 1;

  // this group's candidates: key = class | score descending | index ascending
  for (int i = tid; i < n; i += GT) {
    const int c = cls[i];
    if (c % NG == g) {
      const int pos = atomicAdd(&s_m, 1);
      keys[pos] = ((unsigned long long)(unsigned)c << 49) | ((unsigned long long)(~ordered_u32(scores[i])) << 17) |
                  (unsigned long long)i;
    }
  }
  for (int i = m + tid; i < m_pad; i += GT) keys[i] = ~0ull;
  __syncthreads();
  bitonic_sort_t<GT>(keys, m_pad);
  for (int j = tid; j < m; j += GT) {
    const unsigned idx = (unsigned)(keys[j] & 0x1ffffu);
    float4 b = boxes[idx];
    if (mode == 1) {
      const float o = __fmul_rn((float)cls[idx], off_unit);
      b.x = __fadd_rn(b.x, o);
      b.y = __fadd_rn(b.y, o);
      b.z = __fadd_rn(b.z, o);
      b.w = __fadd_rn(b.w, o);
    }
    sbox[j] = b;
    flag[j] = 0;
  }
  __syncthreads();

  // greedy NMS per class segment (same reduction as segmented_nms_kernel)
  {
    int seg_ord = 0;
    for (int j0 = 0; j0 < m; j0 += 32) {
      const int j = j0 + lane;
      bool start = false;
      if (j < m) start = (j == 0) || ((keys[j] >> 49) != (keys[j - 1] >> 49));
      unsigned sm = __ballot_sync(0xffffffffu, start);
      while (sm) {
        const int l = __ffs(sm) - 1;
        sm &= sm - 1;
        const int s = j0 + l;
        if ((seg_ord % (GT / 32)) == warp) {
          const unsigned long long c = keys[s] >> 49;
          int e;
          {
            int lo = s, hi = m;
            while (hi - lo > 1) {
              const int mid = (lo + hi) >> 1;
              if ((keys[mid] >> 49) == c) lo = mid; else hi = mid;
            }
            e = lo + 1;
          }
          for (int c0 = s; c0 < e; c0 += 32) {
            const int q = c0 + lane;
            const bool valid = q < e;
            float4 my = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) my = sbox[q];
            const float my_area = box_area_rn(my.x, my.y, my.z, my.w);
            bool alive = valid;
            for (int t0 = s; t0 < c0; t0 += 32) {
              unsigned km = (unsigned)flag[t0 + lane];
              km = __ballot_sync(0xffffffffu, km != 0);
              while (km) {
                const int kl = __ffs(km) - 1;
                km &= km - 1;
                const float4 kb = sbox[t0 + kl];
                const float inter = box_inter_rn(kb.x, kb.y, kb.z, kb.w, my.x, my.y, my.z, my.w);
                const float iou = iou_from_rn(inter, box_area_rn(kb.x, kb.y, kb.z, kb.w), my_area);
                if (iou > p.thr) alive = false;
              }
              if (!__any_sync(0xffffffffu, alive)) break;
            }
            for (int l2 = 0; l2 < 32; ++l2) {
              const unsigned am = __ballot_sync(0xffffffffu, alive);
              if (!((am >> l2) & 1u)) continue;
              if ((am >> l2) <= 1u) break;
              const float kx = __shfl_sync(0xffffffffu, my.x, l2), ky = __shfl_sync(0xffffffffu, my.y, l2);
              const float kz = __shfl_sync(0xffffffffu, my.z, l2), kw = __shfl_sync(0xffffffffu, my.w, l2);
              const float ka = __shfl_sync(0xffffffffu, my_area, l2);
              if (alive && lane > l2) {
                const float inter = box_inter_rn(kx, ky, kz, kw, my.x, my.y, my.z, my.w);
                if (iou_from_rn(inter, ka, my_area) > p.thr) alive = false;
              }
            }
            if (valid) flag[q] = alive ? 1 : 0;
            __syncwarp();
          }
        }
        ++seg_ord;
      }
    }
  }
  __syncthreads();
  for (int j = tid; j < m; j += GT) p.keep_flag[base + (keys[j] & 0x1ffffu)] = (unsigned char)flag[j];
}

__global__ void __launch_bounds__(NT) nms_topk_kernel(const GroupParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  __shared__ int s_k;
  const int img = blockIdx.x, tid = threadIdx.x;
  const int n = p.counts[img];
  if (n <= 0) {
    if (tid == 0) p.det_counts[img] = 0;
    return;
  }
  if (p.use_old[img]) return;
  const long long base = (long long)p.seg_base[img] * p.seg_mul;
  if (tid == 0) s_k = 0;
  __syncthreads();
  for (int i = tid; i < n; i += NT) {
    if (p.keep_flag[base + i]) {
      const int pos = atomicAdd(&s_k, 1);
      keys[pos] = ((unsigned long long)(~ordered_u32(p.scores[base + i])) << 32) | (unsigned)i;
    }
  }
  __syncthreads();
  const int kk = s_k;
  int k_pad = 1;
  while (k_pad < kk) k_pad <<= 1;
  for (int i = kk + tid; i < k_pad; i += NT) keys[i] = ~0ull;
  __syncthreads();
  bitonic_sort_t<NT>(keys, k_pad);
  const int limit = p.max_keep >= 0 ? p.max_keep : n;
  const int cnt = min(kk, limit);
  for (int j = tid; j < cnt; j += NT) {
    const unsigned idx = (unsigned)(keys[j] & 0xffffffffu);
    const long long o = (long long)img * p.det_stride + j;
    p.det_boxes[o] = p.boxes[base + idx];
    p.det_scores[o] = p.scores[base + idx];
    p.det_classes[o] = p.cls[base + idx];
    p.det_roi[o] = p.cand_roi[base + idx];
  }
  if (tid == 0) p.det_counts[img] = cnt;
}

static size_t nms_smem_bytes() { return (size_t)SMEM_CAP * (8 + 8 + 16 + 4); }

static int launch_nms(NmsParams& p, int n_seg, long long total, void* ws, size_t ws_bytes, cudaStream_t st) {
  const size_t need = unit_nms_workspace_bytes(n_seg, (int)total);
  if (!ws || ws_bytes < need) {
    set_error("nms: workspace too small (%zu < %zu)", ws_bytes, need);
    return UNIT_EWORKSPACE;
  }
  const size_t M = 2 * (size_t)total + 64;
  unsigned char* w = (unsigned char*)ws;
  p.ws_key1 = (unsigned long long*)w;
  p.ws_key2 = (unsigned long long*)(w + 8 * M);
  p.ws_box = (float4*)(w + 16 * M);
  p.ws_flag = (int*)(w + 32 * M);
  UNIT_CUDA(cudaFuncSetAttribute(segmented_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)nms_smem_bytes()));
  segmented_nms_kernel<<<n_seg, NT, nms_smem_bytes(), st>>>(p);
  UNIT_CHECK_LAUNCH("segmented_nms_kernel");
  return UNIT_OK;
}

}  // namespace detect
}  // namespace unit

using namespace unit;
using namespace unit::detect;

extern "C" {

int unit_softmax_decode(const float* scores, const float* deltas, const float* proposals, float* probs, float* boxes,
                        int R, int K1, int KB, float wx, float wy, float ww, float wh, float scale_clamp,
                        unit_stream_t stream) {
  UNIT_REQUIRE(R >= 0 && K1 > 0 && KB >= 0, "softmax_decode: bad shape");
  if (R == 0) return UNIT_OK;
  UNIT_REQUIRE(!probs || scores, "softmax_decode: probs requested without scores");
  UNIT_REQUIRE(!boxes || (deltas && proposals && KB > 0), "softmax_decode: boxes requested without deltas/proposals");
  UNIT_REQUIRE((((uintptr_t)deltas | (uintptr_t)proposals | (uintptr_t)boxes) & 15) == 0,
               "softmax_decode: deltas/proposals/boxes must be 16-byte aligned");
  softmax_decode_kernel<<<cdiv((long long)R * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      scores, deltas, (const float4*)proposals, probs, boxes, R, K1, KB, wx, wy, ww, wh, scale_clamp);
  UNIT_CHECK_LAUNCH("softmax_decode_kernel");
  return UNIT_OK;
}

int unit_box_get_deltas(const float* src, const float* tgt, float* deltas, int R, float wx, float wy, float ww,
                        float wh, unit_stream_t stream) {
  UNIT_REQUIRE(R >= 0, "box_get_deltas: bad shape");
  if (R == 0) return UNIT_OK;
  UNIT_REQUIRE(src && tgt && deltas, "box_get_deltas: null pointer");
  UNIT_REQUIRE((((uintptr_t)src | (uintptr_t)tgt | (uintptr_t)deltas) & 15) == 0,
               "box_get_deltas: pointers must be 16-byte aligned");
  get_deltas_kernel<<<cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)src, (const float4*)tgt,
                                                                    (float4*)deltas, R, wx, wy, ww, wh);
  UNIT_CHECK_LAUNCH("get_deltas_kernel");
  return UNIT_OK;
}

int unit_detect_filter(const float* boxes, const float* probs, const int* roi_offsets, const float* image_hw,
                       int n_img, int R, int K, int KB, float score_thresh, float* cand_boxes, float* cand_scores,
                       int* cand_roi, int* cand_cls, int* cand_counts, void* workspace, size_t workspace_bytes,
                       unit_stream_t stream) {
  UNIT_REQUIRE(n_img >= 0 && R >= 0 && K > 0 && (KB == K || KB == 1), "detect_filter: bad shape");
  if (n_img == 0) return UNIT_OK;
  UNIT_REQUIRE(roi_offsets && image_hw && cand_counts && (R == 0 || (boxes && probs && cand_boxes && cand_scores &&
                                                                      cand_roi && cand_cls)),
               "detect_filter: null pointer");
  UNIT_REQUIRE((((uintptr_t)boxes | (uintptr_t)cand_boxes) & 15) == 0, "detect_filter: boxes must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (workspace && workspace_bytes >= (size_t)R * 8 + 256 && R > 0 && !switches().filter_single) {
    // three small launches spread over the GPU (rows -> counts, per-image scans, rows -> candidates)
    int* row_valid = (int*)workspace;
    int* row_cnt = row_valid + R;
    filter_count_kernel<<<cdiv(R, FW), FW * 32, 0, st>>>(boxes, probs, R, K, KB, score_thresh, row_valid, row_cnt);
    UNIT_CHECK_LAUNCH("filter_count_kernel");
    filter_scan_kernel<<<n_img, 1024, 0, st>>>(roi_offsets, row_valid, row_cnt, cand_counts);
    UNIT_CHECK_LAUNCH("filter_scan_kernel");
    filter_emit_kernel<<<cdiv(R, FW), FW * 32, 0, st>>>(boxes, probs, roi_offsets, image_hw, n_img, R, K, KB,
                                                        score_thresh, row_valid, row_cnt, (float4*)cand_boxes,
                                                        cand_scores, cand_roi, cand_cls);
    UNIT_CHECK_LAUNCH("filter_emit_kernel");
    return UNIT_OK;
  }
  detect_filter_kernel<<<n_img, FT, 0, st>>>(boxes, probs, roi_offsets, image_hw, K, KB, score_thresh,
                                             (float4*)cand_boxes, cand_scores, cand_roi, cand_cls, cand_counts);
  UNIT_CHECK_LAUNCH("detect_filter_kernel");
  return UNIT_OK;
}

size_t unit_nms_workspace_bytes(int n_seg, int total_candidates) {
  (void)n_seg;
  const size_t M = 2 * (size_t)(total_candidates > 0 ? total_candidates : 0) + 64;
  // segmented kernel arrays + the grouped path's keep flags (1 B / candidate) and per-image switch
  return 36 * M + 256 + (size_t)(total_candidates > 0 ? total_candidates : 0) + (size_t)(n_seg > 0 ? n_seg : 0) + 512;
}

int unit_detect_nms(const float* cand_boxes, const float* cand_scores, const int* cand_roi, const int* cand_cls,
                    const int* cand_counts, const int* roi_offsets, int n_img, int R, int K, float nms_thresh, int nms_mode,
                    int topk, float* det_boxes, float* det_scores, int64_t* det_classes, int64_t* det_roi,
                    int* det_counts, void* workspace, size_t workspace_bytes, unit_stream_t stream) {
  UNIT_REQUIRE(n_img >= 0 && K > 0 && nms_mode >= 0 && nms_mode <= 3, "detect_nms: bad arguments");
  if (n_img == 0) return UNIT_OK;
  UNIT_REQUIRE(cand_counts && roi_offsets && det_counts && det_boxes && det_scores && det_classes && det_roi,
               "detect_nms: null pointer");
  UNIT_REQUIRE(topk >= 0, "detect_nms: topk must be >= 0 (the per-image output stride)");
  NmsParams p = {};
  p.boxes = (const float4*)cand_boxes;
  p.scores = cand_scores;
  p.cls = cand_cls;
  p.cls_i64 = 0;
  p.seg_counts = cand_counts;
  p.seg_base = roi_offsets;
  p.seg_mul = K;
  p.thr = nms_thresh;
  p.mode = nms_mode;
  p.max_keep = topk;
  p.keep = nullptr;
  p.keep_counts = det_counts;
  p.det_boxes = (float4*)det_boxes;
  p.det_scores = det_scores;
  p.det_classes = det_classes;
  p.det_roi = det_roi;
  p.cand_roi = cand_roi;
  p.det_stride = topk;
  const long long total = (long long)R * K;
  if (K >= 2 && cand_cls && total < (1ll << 17) * 64 && !switches().nms_single) {
    // grouped path: NG CTAs per image + a top-k kernel; images it cannot take fall through to the segmented kernel
    const size_t need = unit_nms_workspace_bytes(n_img, (int)total);
    if (!workspace || workspace_bytes < need) {
      set_error("nms: workspace too small (%zu < %zu)", workspace_bytes, need);
      return UNIT_EWORKSPACE;
    }
    const size_t M = 2 * (size_t)total + 64;
    unsigned char* flags = (unsigned char*)workspace + 36 * M + 256;
    GroupParams gp = {};
    gp.boxes = (const float4*)cand_boxes;
    gp.scores = cand_scores;
    gp.cls = cand_cls;
    gp.counts = cand_counts;
    gp.seg_base = roi_offsets;
    gp.seg_mul = K;
    gp.thr = nms_thresh;
    gp.mode = nms_mode;
    gp.keep_flag = flags;
    gp.use_old = flags + total + 128;
    gp.max_keep = topk;
    gp.det_stride = topk;
    gp.det_boxes = (float4*)det_boxes;
    gp.det_scores = det_scores;
    gp.det_classes = det_classes;
    gp.det_roi = det_roi;
    gp.cand_roi = cand_roi;
    gp.det_counts = det_counts;
    cudaStream_t st = (cudaStream_t)stream;
    UNIT_CUDA(cudaMemsetAsync(gp.use_old, 1, (size_t)n_img, st));  // images with no candidates stay "old": harmless
    const size_t gsm = (size_t)GCAP * (8 + 16 + 4);
    UNIT_CUDA(cudaFuncSetAttribute(nms_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
    nms_group_kernel<<<dim3(NG, n_img), GT, gsm, st>>>(gp);
    UNIT_CHECK_LAUNCH("nms_group_kernel");
    const size_t tsm = (size_t)TOPCAP * 8;
    UNIT_CUDA(cudaFuncSetAttribute(nms_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
    nms_topk_kernel<<<n_img, NT, tsm, st>>>(gp);
    UNIT_CHECK_LAUNCH("nms_topk_kernel");
    p.seg_enable = gp.use_old;
  }
  return launch_nms(p, n_img, total, workspace, workspace_bytes, (cudaStream_t)stream);
}

int unit_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int N, float iou_thresh,
                     int nms_mode, int max_keep, int64_t* keep, int* keep_count, void* workspace,
                     size_t workspace_bytes, unit_stream_t stream) {
  UNIT_REQUIRE(N >= 0 && nms_mode >= 0 && nms_mode <= 3, "batched_nms: bad arguments");
  UNIT_REQUIRE(keep_count, "batched_nms: null keep_count");
  if (N == 0) {
    UNIT_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int), (cudaStream_t)stream));
    return UNIT_OK;
  }
  UNIT_REQUIRE(boxes && scores && keep, "batched_nms: null pointer");
  UNIT_REQUIRE((((uintptr_t)boxes) & 15) == 0, "batched_nms: boxes must be 16-byte aligned");
  NmsParams p = {};
  p.boxes = (const float4*)boxes;
  p.scores = scores;
  p.cls = idxs;
  p.cls_i64 = 1;
  p.n_single = N;
  p.thr = iou_thresh;
  p.mode = nms_mode;
  p.max_keep = max_keep;
  p.keep = keep;
  p.keep_stride = N;
  p.keep_counts = keep_count;
  return launch_nms(p, 1, N, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------- fused Fast R-CNN loss
// [D2] FastRCNNOutputs.losses (reached at fast_rcnn.py:438-445): softmax cross-entropy (mean over all RoIs) +
// smooth-L1 between the predicted deltas of the GT class and Box2BoxTransform.get_deltas(proposal, gt_box), summed
// over foreground RoIs and divided by the number of RoIs.  One launch produces both losses AND both gradients
// (SURVEY.md section 8f rank 2); a second single-block launch reduces the per-row terms in a fixed order.
namespace unit {
namespace detect {

__global__ void fastrcnn_loss_kernel(const float* __restrict__ scores, const float* __restrict__ deltas,
                                     const float4* __restrict__ proposals, const float4* __restrict__ gt_boxes,
                                     const int64_t* __restrict__ gt_classes, int R, int K, float wx, float wy,
                                     float ww, float wh, float beta, float* __restrict__ row_loss,
                                     float* __restrict__ d_scores, float* __restrict__ d_deltas, int ld_ds, int ld_dd,
                                     int pad_cols) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= R) return;
  const int K1 = K + 1;
  const float invR = 1.f / (float)R;
  const int cls = (int)gt_classes[r];
  const float* s = scores + (long long)r * K1;
  float m = -INFINITY;
  for (int k = lane; k < K1; k += 32) m = fmaxf(m, s[k]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int k = lane; k < K1; k += 32) sum += expf(s[k] - m);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float lse = m + logf(sum);
  for (int k = lane; k < K1; k += 32)
    d_scores[(long long)r * ld_ds + k] = (expf(s[k] - lse) - (k == cls ? 1.f : 0.f)) * invR;
  float box_loss = 0.f;
  const bool fg = cls >= 0 && cls < K;
  float dv = 0.f;
  if (fg && lane < 4) {
    const float4 p = proposals[r], g = gt_boxes[r];
    const float sw = __fsub_rn(p.z, p.x), sh = __fsub_rn(p.w, p.y);
    const float scx = __fadd_rn(p.x, __fmul_rn(0.5f, sw)), scy = __fadd_rn(p.y, __fmul_rn(0.5f, sh));
    const float tw = __fsub_rn(g.z, g.x), th = __fsub_rn(g.w, g.y);
    const float tcx = __fadd_rn(g.x, __fmul_rn(0.5f, tw)), tcy = __fadd_rn(g.y, __fmul_rn(0.5f, th));
    float t;
    if (lane == 0) t = __fdiv_rn(__fmul_rn(wx, __fsub_rn(tcx, scx)), sw);
    else if (lane == 1) t = __fdiv_rn(__fmul_rn(wy, __fsub_rn(tcy, scy)), sh);
    else if (lane == 2) t = __fmul_rn(ww, logf(__fdiv_rn(tw, sw)));
    else t = __fmul_rn(wh, logf(__fdiv_rn(th, sh)));
    const float diff = deltas[(long long)r * 4 * K + 4 * cls + lane] - t;
    const float n = fabsf(diff);
    if (beta < 1e-5f) {
      box_loss = n;
      dv = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
    } else if (n < beta) {
      box_loss = 0.5f * n * n / beta;
      dv = diff / beta;
    } else {
      box_loss = n - 0.5f * beta;
      dv = diff > 0.f ? 1.f : -1.f;
    }
  }
  const float dv0 = __shfl_sync(0xffffffffu, dv, 0), dv1 = __shfl_sync(0xffffffffu, dv, 1);
  const float dv2 = __shfl_sync(0xffffffffu, dv, 2), dv3 = __shfl_sync(0xffffffffu, dv, 3);
  for (int i = lane; i < 4 * K; i += 32) {
    float v = 0.f;
    const int j = i - 4 * cls;
    if (fg && j >= 0 && j < 4) v = (j == 0 ? dv0 : j == 1 ? dv1 : j == 2 ? dv2 : dv3) * invR;
    d_deltas[(long long)r * ld_dd + i] = v;
  }
  for (int i = 4 * K + lane; i < 4 * K + pad_cols; i += 32) d_deltas[(long long)r * ld_dd + i] = 0.f;  // row padding
  box_loss += __shfl_xor_sync(0xffffffffu, box_loss, 1);
  box_loss += __shfl_xor_sync(0xffffffffu, box_loss, 2);
  if (lane == 0) {
    row_loss[r] = lse - s[cls >= 0 && cls < K1 ? cls : 0];
    row_loss[R + r] = box_loss;
  }
}


// Per-RoI (unreduced) Fast R-CNN losses and their per-row gradients -- FastRCNNOutputsReduction / NLL / Regression
// (fast_rcnn.py:24-130, weak_detector_fast_rcnn.py:23-37).  One warp per RoI:
//   row_ce[r]    = cross_entropy(scores[r], cls)            (nll != 0: -scores[r][cls], the input already is log-probs)
//   row_box[r,j] = smooth_l1(deltas[r, 4 cls + j] - get_deltas(proposal, gt)[j])   for foreground rows, else 0
//   d_scores[r]  = d row_ce[r] / d scores[r]      d_box[r,j] = d row_box[r,j] / d deltas[r, 4 cls + j]
__global__ void row_losses_kernel(const float* __restrict__ scores, const float* __restrict__ deltas,
                                  const float4* __restrict__ proposals, const float4* __restrict__ gt_boxes,
                                  const int64_t* __restrict__ gt_classes, int R, int K, float wx, float wy, float ww,
                                  float wh, float beta, int nll, float* __restrict__ row_ce, float* __restrict__ row_box,
                                  float* __restrict__ d_scores, float* __restrict__ d_box) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= R) return;
  const int K1 = K + 1;
  const int cls = (int)gt_classes[r];
  const float* s = scores + (long long)r * K1;
  if (nll) {
    for (int k = lane; k < K1; k += 32) d_scores[(long long)r * K1 + k] = k == cls ? -1.f : 0.f;
    if (lane == 0) row_ce[r] = -s[cls >= 0 && cls < K1 ? cls : 0];
  } else {
    float m = -INFINITY;
    for (int k = lane; k < K1; k += 32) m = fmaxf(m, s[k]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int k = lane; k < K1; k += 32) sum += expf(s[k] - m);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float lse = m + logf(sum);
    for (int k = lane; k < K1; k += 32) d_scores[(long long)r * K1 + k] = expf(s[k] - lse) - (k == cls ? 1.f : 0.f);
    if (lane == 0) row_ce[r] = lse - s[cls >= 0 && cls < K1 ? cls : 0];
  }
  if (!deltas || lane >= 4) return;
  float loss = 0.f, dv = 0.f;
  if (cls >= 0 && cls < K) {
    const float4 p = proposals[r], g = gt_boxes[r];
    const float sw = __fsub_rn(p.z, p.x), sh = __fsub_rn(p.w, p.y);
    const float scx = __fadd_rn(p.x, __fmul_rn(0.5f, sw)), scy = __fadd_rn(p.y, __fmul_rn(0.5f, sh));
    const float tw = __fsub_rn(g.z, g.x), th = __fsub_rn(g.w, g.y);
    const float tcx = __fadd_rn(g.x, __fmul_rn(0.5f, tw)), tcy = __fadd_rn(g.y, __fmul_rn(0.5f, th));
    float t;
    if (lane == 0) t = __fdiv_rn(__fmul_rn(wx, __fsub_rn(tcx, scx)), sw);
    else if (lane == 1) t = __fdiv_rn(__fmul_rn(wy, __fsub_rn(tcy, scy)), sh);
    else if (lane == 2) t = __fmul_rn(ww, logf(__fdiv_rn(tw, sw)));
    else t = __fmul_rn(wh, logf(__fdiv_rn(th, sh)));
    const float diff = deltas[(long long)r * 4 * K + 4 * cls + lane] - t;
    const float n = fabsf(diff);
    if (beta < 1e-5f) {
      loss = n;
      dv = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
    } else if (n < beta) {
      loss = 0.5f * n * n / beta;
      dv = diff / beta;
    } else {
      loss = n - 0.5f * beta;
      dv = diff > 0.f ? 1.f : -1.f;
    }
  }
  row_box[(long long)r * 4 + lane] = loss;
  d_box[(long long)r * 4 + lane] = dv;
}

__global__ void loss_reduce_kernel(const float* __restrict__ row_loss, int R, float* __restrict__ out, int with_total) {
  __shared__ float s0[32], s1[32];
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    a += row_loss[i];
    b += row_loss[R + i];
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s0[threadIdx.x >> 5] = a;
    s1[threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ta = 0.f, tb = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      ta += s0[w];
      tb += s1[w];
    }
    out[0] = ta / (float)R;
    out[1] = tb / (float)R;
    if (with_total) out[2] = out[0] + out[1];  // what the trainer sums on the host side of the reference
  }
}

}  // namespace detect
}  // namespace unit

extern "C" int unit_fastrcnn_loss_packed(const float* scores, const float* deltas, const float* proposals,
                                         const float* gt_boxes, const int64_t* gt_classes, int R, int K, float wx,
                                         float wy, float ww, float wh, float smooth_l1_beta, float* losses,
                                         float* d_packed, int ld_packed, void* workspace, size_t workspace_bytes,
                                         unit_stream_t stream) {
  UNIT_REQUIRE(R >= 0 && K > 0, "fastrcnn_loss: bad shape");
  UNIT_REQUIRE(losses, "fastrcnn_loss: null losses");
  UNIT_REQUIRE(ld_packed >= 5 * K + 1, "fastrcnn_loss_packed: row stride smaller than (K+1) + 4K");
  cudaStream_t st = (cudaStream_t)stream;
  if (R == 0) {
    UNIT_CUDA(cudaMemsetAsync(losses, 0, 3 * sizeof(float), st));
    return UNIT_OK;
  }
  UNIT_REQUIRE(scores && deltas && proposals && gt_boxes && gt_classes && d_packed, "fastrcnn_loss: null pointer");
  UNIT_REQUIRE((((uintptr_t)proposals | (uintptr_t)gt_boxes) & 15) == 0, "fastrcnn_loss: boxes must be 16-byte aligned");
  if (!workspace || workspace_bytes < (size_t)2 * R * sizeof(float)) {
    set_error("fastrcnn_loss: workspace too small");
    return UNIT_EWORKSPACE;
  }
  unit::detect::fastrcnn_loss_kernel<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(
      scores, deltas, (const float4*)proposals, (const float4*)gt_boxes, gt_classes, R, K, wx, wy, ww, wh,
      smooth_l1_beta, (float*)workspace, d_packed, d_packed + (K + 1), ld_packed, ld_packed, ld_packed - (5 * K + 1));
  UNIT_CHECK_LAUNCH("fastrcnn_loss_kernel");
  unit::detect::loss_reduce_kernel<<<1, 1024, 0, st>>>((const float*)workspace, R, losses, 1);
  UNIT_CHECK_LAUNCH("loss_reduce_kernel");
  return UNIT_OK;
}

extern "C" int unit_fastrcnn_loss(const float* scores, const float* deltas, const float* proposals,
                                  const float* gt_boxes, const int64_t* gt_classes, int R, int K, float wx, float wy,
                                  float ww, float wh, float smooth_l1_beta, float* losses, float* d_scores,
                                  float* d_deltas, void* workspace, size_t workspace_bytes, unit_stream_t stream) {
  UNIT_REQUIRE(R >= 0 && K > 0, "fastrcnn_loss: bad shape");
  UNIT_REQUIRE(losses, "fastrcnn_loss: null losses");
  cudaStream_t st = (cudaStream_t)stream;
  if (R == 0) {
    UNIT_CUDA(cudaMemsetAsync(losses, 0, 2 * sizeof(float), st));
    return UNIT_OK;
  }
  UNIT_REQUIRE(scores && deltas && proposals && gt_boxes && gt_classes && d_scores && d_deltas,
               "fastrcnn_loss: null pointer");
  UNIT_REQUIRE((((uintptr_t)proposals | (uintptr_t)gt_boxes) & 15) == 0, "fastrcnn_loss: boxes must be 16-byte aligned");
  if (!workspace || workspace_bytes < (size_t)2 * R * sizeof(float)) {
    set_error("fastrcnn_loss: workspace too small");
    return UNIT_EWORKSPACE;
  }
  unit::detect::fastrcnn_loss_kernel<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(
      scores, deltas, (const float4*)proposals, (const float4*)gt_boxes, gt_classes, R, K, wx, wy, ww, wh,
      smooth_l1_beta, (float*)workspace, d_scores, d_deltas, K + 1, 4 * K, 0);
  UNIT_CHECK_LAUNCH("fastrcnn_loss_kernel");
  unit::detect::loss_reduce_kernel<<<1, 1024, 0, st>>>((const float*)workspace, R, losses, 0);
  UNIT_CHECK_LAUNCH("loss_reduce_kernel");
  return UNIT_OK;
}

extern "C" int unit_fastrcnn_row_losses(const float* scores, const float* deltas, const float* proposals,
                                        const float* gt_boxes, const int64_t* gt_classes, int R, int K, float wx,
                                        float wy, float ww, float wh, float smooth_l1_beta, int nll, float* row_ce,
                                        float* row_box, float* d_scores, float* d_box, unit_stream_t stream) {
  UNIT_REQUIRE(R >= 0 && K > 0, "fastrcnn_row_losses: bad shape");
  if (R == 0) return UNIT_OK;
  UNIT_REQUIRE(scores && gt_classes && row_ce && d_scores, "fastrcnn_row_losses: null pointer");
  UNIT_REQUIRE(!deltas || (proposals && gt_boxes && row_box && d_box), "fastrcnn_row_losses: box inputs missing");
  UNIT_REQUIRE((((uintptr_t)proposals | (uintptr_t)gt_boxes) & 15) == 0,
               "fastrcnn_row_losses: boxes must be 16-byte aligned");
  unit::detect::row_losses_kernel<<<cdiv((long long)R * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      scores, deltas, (const float4*)proposals, (const float4*)gt_boxes, gt_classes, R, K, wx, wy, ww, wh,
      smooth_l1_beta, nll, row_ce, row_box, d_scores, d_box);
  UNIT_CHECK_LAUNCH("row_losses_kernel");
  return UNIT_OK;
}
