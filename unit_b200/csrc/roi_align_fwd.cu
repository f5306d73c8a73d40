// ROIAlign forward, slab-resident kernel v2.
//
// Work unit = (image, 8-channel slab, RoI); a persistent CTA (one per SM) owns a contiguous range of units, keeps
// the slab -- 8 full H x W channel planes, stored channel-PAIR interleaved as float2 -- resident in shared
// memory, and its 12 warps pull RoIs of that slab from a shared counter.  A warp is a self-contained pipeline
// (no CTA barrier in steady state):
//   1. build the RoI's separable sampling tables in its private shared-memory area (fp32 operation order of the
//      torchvision kernel, so floor / validity decisions are identical to the reference's);
//   2. lane (q, cp), q = 0..6, cp = 0..3 computes output rows {q, q+7} of channels {2cp, 2cp+1}: it slides a window
//      of vertically interpolated float2 values (one LDS.64 + one packed FFMA2 per tap serves both channels) along
//      x, consuming the x-table in one rolled loop (advance / end-of-bin flags ride in the sign bits of the
//      weights), and drops each finished bin into the warp's staging block;
//   3. the [8 ch][14][14] block -- 6272 contiguous bytes of the NCHW output -- leaves as one TMA bulk store
//      (cp.async.bulk.global.shared::cta); the next RoI's table build overlaps the store.
// Every feature byte is read from HBM/L2 once per slab, every output byte is written once, fully coalesced.
#include "roi_slab.cuh"

namespace unit {
namespace roi {
namespace v2 {

template <int GH>
struct Taps {
  const float2* alo[GH > 0 ? GH : 1];
  const float2* ahi[GH > 0 ? GH : 1];
  const float2* blo[GH > 0 ? GH : 1];
  const float2* bhi[GH > 0 ? GH : 1];
  float ah[GH > 0 ? GH : 1], al[GH > 0 ? GH : 1], bh[GH > 0 ? GH : 1], bl[GH > 0 ? GH : 1];
};

// vertically interpolated float2 (channel pair) of the column `off` elements past the tap pointers, rows a and b
template <int GH>
__device__ __forceinline__ void column(const Taps<GH>& t, const float2* __restrict__ pl, const YTap* ya,
                                       const YTap* yb, int gh, int col, int off, float2& va, float2& vb) {
  va = make_float2(0.f, 0.f);
  vb = make_float2(0.f, 0.f);
  if (GH > 0) {
#pragma unroll
    for (int i = 0; i < GH; ++i) {
      va = ffma2(t.ah[i], t.alo[i][off], va);
      va = ffma2(t.al[i], t.ahi[i][off], va);
      vb = ffma2(t.bh[i], t.blo[i][off], vb);
      vb = ffma2(t.bl[i], t.bhi[i][off], vb);
    }
  } else {
    for (int i = 0; i < gh; ++i) {
      const YTap a = ya[i], b = yb[i];
      va = ffma2(a.h, pl[a.lo + col + off], va);
      va = ffma2(a.l, pl[a.hi + col + off], va);
      vb = ffma2(b.h, pl[b.lo + col + off], vb);
      vb = ffma2(b.l, pl[b.hi + col + off], vb);
    }
  }
}

// Lane task: rows {q, q+7} of channels {2cp, 2cp+1}.  The window (lo, hi, next) rotates through three register
// sets by unrolling the column loop three times, so advancing costs no register moves; the inner loop consumes the
// samples whose lower tap is the current column.
template <typename T, int GH>
__device__ __forceinline__ void task(const float2* __restrict__ pl, int W, const WarpArea<T>* wa, int q, int cp,
                                     T* __restrict__ stage) {
  const int gh = wa->hdr.gh;
  const YTap* ya = wa->ytab + q * gh;
  const YTap* yb = wa->ytab + (q + 7) * gh;
  int col = wa->hdr.x0;
  Taps<GH> t;
  if (GH > 0) {
#pragma unroll
    for (int i = 0; i < GH; ++i) {
      const YTap a = ya[i], b = yb[i];
      t.alo[i] = pl + a.lo + col;
      t.ahi[i] = pl + a.hi + col;
      t.blo[i] = pl + b.lo + col;
      t.bhi[i] = pl + b.hi + col;
      t.ah[i] = a.h;
      t.al[i] = a.l;
      t.bh[i] = b.h;
      t.bl[i] = b.l;
    }
  }
  float2 w0a, w0b, w1a, w1b, w2a, w2b;
  column<GH>(t, pl, ya, yb, gh, col, 0, w0a, w0b);
  column<GH>(t, pl, ya, yb, gh, col, 1, w1a, w1b);
  column<GH>(t, pl, ya, yb, gh, col, 2, w2a, w2b);
  int remaining = wa->hdr.nsamp;
  float2 acc_a = make_float2(0.f, 0.f), acc_b = make_float2(0.f, 0.f);
  T* da = stage + (2 * cp) * (P * P) + q * P;  // row q of channel 2cp; channel 2cp+1 is P*P further
  T* db = da + 7 * P;                          // row q+7
  const float2* xt = wa->xtab;
  float2 e = xt[0];

#define UNIT_CONSUME(LA, LB, HA, HB)                              \
  {                                                               \
    bool lastcol;                                                 \
    do {                                                          \
      const float2 en = xt[1];                                    \
      ++xt;                                                       \
      lastcol = (__float_as_uint(e.x) >> 31) != 0;                \
      const bool end = (__float_as_uint(e.y) >> 31) != 0;         \
      const float hx = fabsf(e.x), lx = fabsf(e.y);               \
      acc_a = ffma2(hx, LA, acc_a);                               \
      acc_b = ffma2(hx, LB, acc_b);                               \
      acc_a = ffma2(lx, HA, acc_a);                               \
      acc_b = ffma2(lx, HB, acc_b);                               \
      if (end) {                                                  \
        stf(da, acc_a.x);                                         \
        stf(da + P * P, acc_a.y);                                 \
        stf(db, acc_b.x);                                         \
        stf(db + P * P, acc_b.y);                                 \
        ++da;                                                     \
        ++db;                                                     \
        acc_a = make_float2(0.f, 0.f);                            \
        acc_b = make_float2(0.f, 0.f);                            \
      }                                                           \
      e = en;                                                     \
      --remaining;                                                \
    } while (!lastcol);                                           \
  }

  while (true) {
    UNIT_CONSUME(w0a, w0b, w1a, w1b);
    if (remaining <= 0) break;
    column<GH>(t, pl, ya, yb, gh, col, 3, w0a, w0b);
    UNIT_CONSUME(w1a, w1b, w2a, w2b);
    if (remaining <= 0) break;
    column<GH>(t, pl, ya, yb, gh, col, 4, w1a, w1b);
    UNIT_CONSUME(w2a, w2b, w0a, w0b);
    if (remaining <= 0) break;
    column<GH>(t, pl, ya, yb, gh, col, 5, w2a, w2b);
    col += 3;
    if (GH > 0) {
#pragma unroll
      for (int i = 0; i < GH; ++i) {
        t.alo[i] += 3;
        t.ahi[i] += 3;
        t.blo[i] += 3;
        t.bhi[i] += 3;
      }
    }
  }
#undef UNIT_CONSUME
}

// direct evaluation (grid larger than MAXG or irregular sample steps): every lane fills its rows sample by sample
template <typename T>
__device__ __noinline__ void task_direct(const float2* __restrict__ pl, int H, int W, const Header& hdr, int q, int cp,
                                         T* __restrict__ stage) {
  for (int half = 0; half < 2; ++half) {
    const int ph = q + 7 * half;
    for (int pw = 0; pw < P; ++pw) {
      float2 acc = make_float2(0.f, 0.f);
      for (int iy = 0; iy < hdr.gh; ++iy) {
        int ylo, yhi;
        float ly, hy;
        const bool vy = axis_tap(sample_coord(hdr.start_h, hdr.bin_h, ph, iy, hdr.gh), H, ylo, yhi, ly, hy);
        for (int ix = 0; ix < hdr.gw; ++ix) {
          int xlo, xhi;
          float lx, hx;
          const bool vx = axis_tap(sample_coord(hdr.start_w, hdr.bin_w, pw, ix, hdr.gw), W, xlo, xhi, lx, hx);
          if (vy && vx) {
            const float2 v1 = pl[ylo * W + xlo], v2 = pl[ylo * W + xhi], v3 = pl[yhi * W + xlo], v4 = pl[yhi * W + xhi];
            acc.x += hy * (hx * v1.x + lx * v2.x) + ly * (hx * v3.x + lx * v4.x);
            acc.y += hy * (hx * v1.y + lx * v2.y) + ly * (hx * v3.y + lx * v4.y);
          }
        }
      }
      stf(stage + (2 * cp) * (P * P) + ph * P + pw, acc.x * hdr.inv_count);
      stf(stage + (2 * cp + 1) * (P * P) + ph * P + pw, acc.y * hdr.inv_count);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(NTHREADS, 1) roi_align_fwd_slab2(const Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* slab = reinterpret_cast<float*>(smem_raw);
  const size_t slab_bytes = (size_t)NPAIR * p.pair_stride * sizeof(float2);
  WarpArea<T>* areas = reinterpret_cast<WarpArea<T>*>(smem_raw + ((slab_bytes + 127) / 128) * 128);
  __shared__ int s_next;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  WarpArea<T>* wa = areas + warp;
  const int nslab = p.C / CS;
  const int HW = p.H * p.W;
  const T* feat = reinterpret_cast<const T*>(p.feat);
  T* out = reinterpret_cast<T*>(p.out);
  const int q = lane >> 2, cp = lane & 3;
  const float2* pl = reinterpret_cast<const float2*>(slab) + (size_t)cp * p.pair_stride;

  const long long u_begin = p.units_total * blockIdx.x / gridDim.x;
  const long long u_end = p.units_total * (blockIdx.x + 1) / gridDim.x;
  long long u = u_begin;
  int n = 0;
  while (u < u_end) {
    while (n < p.N && (long long)p.img_off[n + 1] * nslab <= u) ++n;
    if (n >= p.N) break;
    const int r_base = p.img_off[n];
    const int Rn = p.img_off[n + 1] - r_base;
    const long long local = u - (long long)r_base * nslab;
    const int k = (int)(local / Rn);
    const int r0 = (int)(local - (long long)k * Rn);
    const long long seg_end_u = min(u_end, (long long)r_base * nslab + (long long)(k + 1) * Rn);
    const int r1 = r0 + (int)(seg_end_u - u);

    __syncthreads();  // every warp has finished reading the previous slab
    if (tid == 0) s_next = r0;
    load_slab<T>(feat + ((long long)n * p.C + (long long)k * CS) * HW, slab, HW, p.pair_stride, tid);
    __syncthreads();

    while (true) {
      int r = 0;
      if (lane == 0) r = atomicAdd(&s_next, 1);
      r = __shfl_sync(0xffffffffu, r, 0);
      if (r >= r1) break;
      build_tables<T>(p.rois + (long long)(r_base + r) * 5, p, wa, lane);
      if (lane == 0) bulk_wait_read_all();  // the previous bulk store has drained this warp's staging block
      __syncwarp();
      if (!(p.debug & 1) && lane < 28) {
        const int mode = wa->hdr.mode;
        if (mode == 1) {
          const int gh = wa->hdr.gh;
          if (gh == 1) task<T, 1>(pl, p.W, wa, q, cp, wa->stage);
          else if (gh == 2) task<T, 2>(pl, p.W, wa, q, cp, wa->stage);
          else task<T, 0>(pl, p.W, wa, q, cp, wa->stage);
        } else if (mode == 2) {
          task_direct<T>(pl, p.H, p.W, wa->hdr, q, cp, wa->stage);
        } else {
          for (int i = 0; i < P; ++i) {
            stf(wa->stage + (2 * cp) * (P * P) + q * P + i, 0.f);
            stf(wa->stage + (2 * cp + 1) * (P * P) + q * P + i, 0.f);
            stf(wa->stage + (2 * cp) * (P * P) + (q + 7) * P + i, 0.f);
            stf(wa->stage + (2 * cp + 1) * (P * P) + (q + 7) * P + i, 0.f);
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0 && !(p.debug & 2)) {
        bulk_store(out + ((long long)(r_base + r) * p.C + (long long)k * CS) * (P * P), wa->stage,
                   (uint32_t)(CS * P * P * sizeof(T)));
        bulk_commit();
      }
    }
    u = seg_end_u;
  }
  if (lane == 0) bulk_wait_all();
}

template <typename T>
static size_t smem_total(int HW) {
  const size_t slab_bytes = (size_t)NPAIR * pair_stride_host(HW) * sizeof(float2);
  return ((slab_bytes + 127) / 128) * 128 + (size_t)NWARPS * sizeof(WarpArea<T>);
}

}  // namespace v2

bool fwd_slab2_fits(int C, int H, int W, int dtype) {
  if (C % v2::CS) return false;
  const size_t need = dtype == UNIT_F32 ? v2::smem_total<float>(H * W) : v2::smem_total<__nv_bfloat16>(H * W);
  return need <= 227 * 1024;
}

template <typename T>
static int launch_t(const void* feat, const float* rois, void* out, int N, int C, int H, int W, int R, float scale,
                    int sr, int aligned, const int* img_off, cudaStream_t st) {
  v2::Params p;
  p.feat = feat;
  p.rois = rois;
  p.out = out;
  p.img_off = img_off;
  p.N = N;
  p.C = C;
  p.H = H;
  p.W = W;
  p.R = R;
  p.scale = scale;
  p.sampling_ratio = sr;
  p.aligned = aligned;
  p.pair_stride = v2::pair_stride_host(H * W);
  p.units_total = (long long)R * (C / v2::CS);
  p.debug = switches().roi_debug;
  const size_t smem = v2::smem_total<T>(H * W);
  UNIT_CUDA(cudaFuncSetAttribute(v2::roi_align_fwd_slab2<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = p.units_total / 48;  // at least 4 RoIs per warp
  if (grid < 1) grid = 1;
  if (grid > sm_count()) grid = sm_count();
  v2::roi_align_fwd_slab2<T><<<(int)grid, v2::NTHREADS, smem, st>>>(p);
  UNIT_CHECK_LAUNCH("roi_align_fwd_slab2");
  return UNIT_OK;
}

int launch_fwd_slab2(const void* feat, const float* rois, void* out, int N, int C, int H, int W, int R, float scale,
                     int sr, int aligned, int dtype, const int* img_off, cudaStream_t st) {
  if (dtype == UNIT_F32) return launch_t<float>(feat, rois, out, N, C, H, W, R, scale, sr, aligned, img_off, st);
  return launch_t<__nv_bfloat16>(feat, rois, out, N, C, H, W, R, scale, sr, aligned, img_off, st);
}

}  // namespace roi
}  // namespace unit
