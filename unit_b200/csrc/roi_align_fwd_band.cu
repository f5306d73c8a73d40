// ROIAlign forward, slab-resident band-gather kernel.
//
// Work unit = (image, 8-channel slab, RoI).  A persistent CTA (one per SM) owns a contiguous range of units and keeps
// the slab -- 8 full H x W planes, stored as two channel QUADS interleaved per pixel (float4) -- resident in shared
// memory; its 12 warps pull RoIs of that slab from a shared counter, each warp a self-contained pipeline (no CTA
// barrier in steady state).
//
// The average over a bin's sampling grid is separable, and gathering from shared memory makes dynamic indexing free:
//   out[ph][pw] = sum_y Wy[ph][y] * ( sum_j wb[pw][j] * F[y][bx[pw] + j] )
// where wb[pw][.] are bin pw's x-samples pre-summed into a dense band of nb <= 8 columns starting at bx[pw].
//   * a tiny pre-kernel builds every RoI's tables ONCE (the main kernel would otherwise rebuild them for each of
//     the C/8 slabs): the x-bands, and the y-samples as (hy, ly) with an "this sample's lower tap row is one past the
//     previous sample's" flag (same fp32 operation order as torchvision for every coordinate, so floor / validity
//     decisions are the reference's);
//   * lane (pw, half) owns output COLUMN pw of channel quad `half`: its band lives in registers, so one footprint row
//     costs nb x (LDS.128 + 2 FFMA2) for four channels;
//   * the walk is SAMPLE-major: for every bin ph, for every y-sample of the bin: one broadcast LDS.64 of the sample's
//     (hy, ly); if its row advanced, the two live row sums shift (hc <- hn) and the next row sum is formed from pixels
//     that were requested one row earlier; then acc += hy * hc + ly * hn.  One accumulator set is live (a bin is
//     finished before the next starts), there are no shuffles, and every branch is warp-uniform;
//   * tables arrive by TMA bulk load (mbarrier): the x-part is consumed into registers at once, the y-part is double
//     buffered, so the NEXT RoI's tables load while this one computes;
//   * the [8 ch][14][14] block -- 6272 contiguous bytes of the NCHW output -- leaves as one TMA bulk store.
// Every feature byte is read from HBM/L2 once per slab, every output byte is written once, fully coalesced.
#include "roi_slab.cuh"

namespace unit {
namespace roi {
namespace band {

using v2::CS;
using v2::P;
using v2::Params;
constexpr int NQ = CS / 4;   // channel quads per slab
#ifndef UNIT_FWD_NW
#define UNIT_FWD_NW 12
#endif
constexpr int NW = UNIT_FWD_NW;
constexpr int NT = NW * 32;
constexpr int MAXB = 8;      // band columns per bin: sampling grid <= 7 with unit sample steps
constexpr int MAXGY = 4;     // y-samples per bin with tables: RoI height <= 56 feature rows (896 px at 1/16)
constexpr int MAXSY = P * MAXGY;
constexpr uint32_t F_ADV = 1u;  // flag in the mantissa LSB of hy (<= 1 ulp on that weight)

struct __align__(16) RoiTabX {
  int mode, nb, y0, gh;  // mode 0: zero output, 1: tables, 2: direct evaluation
  int gw, nrows, pad0, pad1;
  float inv_count, start_w, start_h, bin_w, bin_h, padf[3];
  int bx[16];            // first column of bin pw's band
  float wb[MAXB][16];    // wb[j][pw]: weight of column bx[pw] + j, already divided by the sample count
  // lane -> (pw | half << 4 | shadow << 7): which output column / channel quad a lane owns for THIS RoI.  An LDS.128
  // is served in four phases of 8 lanes; the pre-kernel places the 28 (pw, half) items so that the 8 pixels of a
  // phase fall into different 16-byte bank groups wherever the RoI's column stride allows it.
  unsigned char lmap[32];
};
struct __align__(16) RoiTabY {
  float2 yt[MAXSY + 8];  // y-sample s: (hy | F_ADV, ly); F_ADV: its lower tap row is the previous sample's + 1
};
struct __align__(16) RoiTab {
  RoiTabX x;
  RoiTabY y;
};
static_assert(sizeof(RoiTabX) == 672 && sizeof(RoiTabY) == 512 && sizeof(RoiTab) == 1184, "RoiTab layout");

template <typename T>
struct __align__(128) WarpArea {
#ifndef UNIT_FWD_DIRECT  // experiment: lanes store straight to global memory, no staging block (more warps fit)
  T stage[CS * P * P];
#endif
  RoiTabX tx;
  RoiTabY ty[2];
  uint64_t bar;
};

typedef unsigned long long f2;
__device__ __forceinline__ f2 pack2(float a, float b) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void fma2(f2& acc, f2 a, f2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ float2 unpack2(f2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------ table pre-kernel
// one warp per RoI
__global__ void __launch_bounds__(128)
roi_tables_kernel(const float* __restrict__ rois, RoiTab* __restrict__ tabs, int R, int N, int H, int W, float scale,
                  int sampling_ratio, int aligned, int* __restrict__ img_off) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* roi = rois + (long long)r * 5;
  RoiTab* t = tabs + r;
  const int n = (int)roi[0];
  if (lane == 0) {
    // per-image RoI offsets from the (sorted) batch-index column, img_off[k] = first RoI of image >= k: the RoI where
    // the index steps writes the entries it passes, the last RoI also the tail (was a launch of its own)
    const int cur = min(max(n, 0), N);
    const int prev = r > 0 ? min(max((int)roi[-5], 0), N) : -1;
    for (int k = prev + 1; k <= cur; ++k) img_off[k] = r;
    if (r == R - 1)
      for (int k = cur + 1; k <= N; ++k) img_off[k] = R;
  }
  const Geom g = roi_geom(roi, scale, P, P, sampling_ratio, aligned);
  int mode = 1;
  if (g.gw <= 0 || g.gh <= 0 || n < 0 || n >= N) mode = 0;
  else if (g.gh > MAXGY || g.gw > MAXB - 1 || H < 2 || W < 2) mode = 2;
  bool bad = false;
  int width = 0;
  int bx_lane = lane;  // this lane's band start (lane = pw); the lane map below reads it by shuffle
  if (mode == 1) {
    const float inv = 1.f / g.count;
    // ---- x: lane pw pre-sums the gw samples of its bin into a dense band
    if (lane < P) {
      float w[MAXB];
#pragma unroll
      for (int j = 0; j < MAXB; ++j) w[j] = 0.f;
      int bx = 0;
      for (int ix = 0; ix < g.gw; ++ix) {
        int lo, hi;
        float l, h;
        axis_tap(sample_coord(g.start_w, g.bin_w, lane, ix, g.gw), W, lo, hi, l, h);
        if (lo >= W - 1) {  // clamped (value F[W-1]) or beyond (weights 0): lower tap W-2 with weights (0, h)
          lo = W - 2;
          l = h;
          h = 0.f;
        }
        if (ix == 0) bx = lo;
        const int d = lo - bx;
        if (d < 0 || d + 1 >= MAXB) {
          bad = true;
        } else {
#pragma unroll
          for (int j = 0; j < MAXB; ++j) {
            if (j == d) w[j] += h * inv;
            if (j == d + 1) w[j] += l * inv;
          }
          width = max(width, d + 2);
        }
      }
      t->x.bx[lane] = bx;
      bx_lane = bx;
#pragma unroll
      for (int j = 0; j < MAXB; ++j) t->x.wb[j][lane] = w[j];
    }
    // ---- y: samples in order; F_ADV marks a sample whose lower tap row is one past the previous sample's
    const int nsy = P * g.gh;
    const float inv_gh = 1.f / (float)g.gh;
    int y0 = 0, last_lo = 0;
    for (int s0 = 0; s0 < nsy; s0 += 32) {
      const int s = s0 + lane;
      int lo2[2] = {0x3fffffff, 0x3fffffff};
      float hh = 0.f, ll = 0.f;
#pragma unroll
      for (int k = 0; k < 2; ++k) {  // k = 0: this sample, k = 1: the previous one (its lower tap row only)
        const int sk = s - k;
        if (sk >= 0 && sk < nsy) {
          int hi;
          float l, h;
          const int ph = (int)(((float)sk + 0.5f) * inv_gh);
          axis_tap(sample_coord(g.start_h, g.bin_h, ph, sk - ph * g.gh, g.gh), H, lo2[k], hi, l, h);
          if (lo2[k] >= H - 1) {
            lo2[k] = H - 2;
            l = h;
            h = 0.f;
          }
          if (k == 0) {
            hh = h;
            ll = l;
          }
        }
      }
      if (s < nsy) {
        const bool adv = s > 0 && lo2[0] != lo2[1];
        if (s > 0 && (lo2[0] - lo2[1] > 1 || lo2[0] < lo2[1])) bad = true;
        const uint32_t bits = (__float_as_uint(hh) & ~1u) | (adv ? F_ADV : 0u);
        t->y.yt[s] = make_float2(__uint_as_float(bits), ll);
      }
      if (s0 == 0) y0 = __shfl_sync(0xffffffffu, lo2[0], 0);
      const int src = min(nsy - 1 - s0, 31);
      if (s0 + 32 >= nsy) last_lo = __shfl_sync(0xffffffffu, lo2[0], src);
    }
    width = __reduce_max_sync(0xffffffffu, width);
    if (lane == 0) {
      t->x.y0 = y0;
      t->x.nrows = last_lo + 2 - y0;
      t->x.nb = width;
    }
  }
  if (__any_sync(0xffffffffu, bad)) mode = 2;
  // ---- lane map (see RoiTabX::lmap).  Bank group of item (pw, half) = (bx[pw] + half * qs) mod 8 with qs = 1 (mod 8)
  // sixteen-byte units between the quad planes; rows and band taps shift every lane alike.  The greedy assignment
  // (28 items into 4 phases of 8 lanes; same pixel already in the phase: free (broadcast); empty bank group: free;
  // otherwise one more wavefront; ties go to the emptier, then the earlier phase) runs on the whole warp: lane
  // 8 * ph + r holds unit_at[ph][r], lane 8 * ph + k holds slot[ph][k], the choice is one warp minimum per item.
  // (The serial form on lane 0 with its arrays in local memory took 11.6 us for 1024 RoIs.)
  const int myph = lane >> 3, myr = lane & 7;
  int my_unit = -1, my_slot = 0, mycnt = 0;  // mycnt = cnt[myph]
#pragma unroll
  for (int i = 0; i < 2 * P; ++i) {
    const int pw = i % P, half = i / P;
    const int bxp = __shfl_sync(0xffffffffu, bx_lane, pw);
    const int u = (mode == 1 ? bxp : pw) + half * 4097;  // distinct per (pixel, half); 4097 = 1 (mod 8)
    unsigned key = 0xffffffffu;
    if (myr == (u & 7) && mycnt < 8) {
      const int cost = my_unit == u ? 0 : (my_unit < 0 ? 1 : 100);
      key = (unsigned)((cost * 16 + mycnt) * 4 + myph);
    }
    const unsigned kmin = __reduce_min_sync(0xffffffffu, key);
    const int best = (int)(kmin & 3u), at = (int)((kmin >> 2) & 15u);  // phase and its fill count
    if (lane == 8 * best + at) my_slot = pw | (half << 4);
    if (myph == best) {
      if (myr == (u & 7) && my_unit < 0) my_unit = u;
      ++mycnt;
    }
  }
  {
    const int slot0 = __shfl_sync(0xffffffffu, my_slot, 8 * myph);
    t->x.lmap[lane] = (unsigned char)(myr < mycnt ? my_slot : (slot0 | 0x80));
  }
  if (lane == 0) {
    t->x.mode = mode;
    t->x.gw = g.gw;
    t->x.gh = g.gh;
    t->x.inv_count = 1.f / g.count;
    t->x.start_w = g.start_w;
    t->x.start_h = g.start_h;
    t->x.bin_w = g.bin_w;
    t->x.bin_h = g.bin_h;
  }
}

// ------------------------------------------------------------------------------------------------ slab loader
// global planar [c][HW] -> shared [c/4][qs] float4 (4 channels per pixel)
template <typename T>
__device__ __forceinline__ void load_slab_quad(const T* __restrict__ src, float4* __restrict__ slab, int HW, int qs,
                                               int tid) {
  constexpr int U = 8;  // pixels per thread per round: 32 loads in flight
  const int total = NQ * HW;
  for (int i0 = tid; i0 < total; i0 += U * NT) {
    float4 v[U];
    int dst[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int i = i0 + k * NT;
      dst[k] = -1;
      if (i < total) {
        const int q = i / HW, o = i - q * HW;
        const T* s = src + (size_t)(4 * q) * HW + o;
        v[k] = make_float4(ldf(s), ldf(s + HW), ldf(s + 2 * (size_t)HW), ldf(s + 3 * (size_t)HW));
        dst[k] = q * qs + o;
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k)
      if (dst[k] >= 0) slab[dst[k]] = v[k];
  }
}

// ------------------------------------------------------------------------------------------------ lane tasks
// NB > 0: band width known at compile time; NB == 0: any width <= MAXB (warp-uniform nb).
// GH > 0: y-samples per bin known at compile time; GH == 0: any count <= MAXGY (warp-uniform gh).
// `yt` is the warp's y-table in shared memory: every lane reads the same entry (one broadcast LDS.64 per sample).
template <typename T>
__device__ __forceinline__ void store4(T* __restrict__ dst, f2 a, f2 b, bool writer) {
  if (writer) {
    const float2 x = unpack2(a), y = unpack2(b);
    stf(dst, x.x);
    stf(dst + P * P, x.y);
    stf(dst + 2 * P * P, y.x);
    stf(dst + 3 * P * P, y.y);
  }
}

template <typename T, int NB, int GH>
__device__ __forceinline__ void task_band(const float4* __restrict__ rowp, int W, int nrows, int gh, int nb,
                                          const float (&w)[MAXB], const float2* __restrict__ yt,
                                          T* __restrict__ dst, bool writer) {
  constexpr int NV = NB > 0 ? NB : MAXB;
  float4 v[NV];
  // The walk requests one row ahead; past the footprint's last row the pointer stops advancing, so the look-ahead
  // re-reads that row (never used) instead of running into the other warps' work areas.
  int rows_left = nrows;  // warp-uniform
  auto load_row = [&]() {
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (NB > 0 || j < nb) v[j] = rowp[j];
    --rows_left;
    rowp += rows_left > 0 ? W : 0;
  };
  auto row_sum = [&](f2& hA, f2& hB) {
    hA = hB = 0ull;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (NB > 0 || j < nb) {
        const f2 wj = pack2(w[j], w[j]);
        fma2(hA, wj, pack2(v[j].x, v[j].y));
        fma2(hB, wj, pack2(v[j].z, v[j].w));
      }
    }
  };
  // prologue: the two rows under the first sample, and the pixels of the row after them are already on their way
  f2 hcA, hcB, hnA, hnB;
  load_row();
  row_sum(hcA, hcB);
  load_row();
  row_sum(hnA, hnB);
  load_row();  // one row ahead
  const int ng = GH > 0 ? GH : gh;
  float2 en = *yt;  // the table entry is requested one sample ahead (the table has spare entries past the last one)
#pragma unroll 1
  for (int ph = 0; ph < P; ++ph) {
    f2 aA = 0ull, aB = 0ull;
#pragma unroll
    for (int i = 0; i < (GH > 0 ? GH : MAXGY); ++i) {
      if (GH > 0 || i < ng) {
        const float2 e = en;
        en = *++yt;
        if (__float_as_uint(e.x) & F_ADV) {  // warp-uniform
          hcA = hnA;
          hcB = hnB;
          row_sum(hnA, hnB);
          load_row();
        }
        const f2 hy = pack2(e.x, e.x), ly = pack2(e.y, e.y);
        fma2(aA, hy, hcA);
        fma2(aB, hy, hcB);
        fma2(aA, ly, hnA);
        fma2(aB, ly, hnB);
      }
    }
    store4<T>(dst, aA, aB, writer);
    dst += P;
  }
}

// direct evaluation (grid larger than the tables or irregular sample steps): lane (pw, half) fills its column
template <typename T>
__device__ __noinline__ void task_direct(const float4* __restrict__ q, int H, int W, int gw, int gh, float inv_count,
                                         float start_w, float start_h, float bin_w, float bin_h, int pw,
                                         T* __restrict__ dst) {
  for (int ph = 0; ph < P; ++ph) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int iy = 0; iy < gh; ++iy) {
      int ylo, yhi;
      float ly, hy;
      const bool vy = axis_tap(sample_coord(start_h, bin_h, ph, iy, gh), H, ylo, yhi, ly, hy);
      for (int ix = 0; ix < gw; ++ix) {
        int xlo, xhi;
        float lx, hx;
        const bool vx = axis_tap(sample_coord(start_w, bin_w, pw, ix, gw), W, xlo, xhi, lx, hx);
        if (vy && vx) {
          const float4 v1 = q[ylo * W + xlo], v2 = q[ylo * W + xhi], v3 = q[yhi * W + xlo], v4 = q[yhi * W + xhi];
          acc.x += hy * (hx * v1.x + lx * v2.x) + ly * (hx * v3.x + lx * v4.x);
          acc.y += hy * (hx * v1.y + lx * v2.y) + ly * (hx * v3.y + lx * v4.y);
          acc.z += hy * (hx * v1.z + lx * v2.z) + ly * (hx * v3.z + lx * v4.z);
          acc.w += hy * (hx * v1.w + lx * v2.w) + ly * (hx * v3.w + lx * v4.w);
        }
      }
    }
    stf(dst + ph * P, acc.x * inv_count);
    stf(dst + ph * P + P * P, acc.y * inv_count);
    stf(dst + ph * P + 2 * P * P, acc.z * inv_count);
    stf(dst + ph * P + 3 * P * P, acc.w * inv_count);
  }
}

__host__ __device__ inline int quad_stride(int HW) {
  int s = HW + MAXB;  // zero padding: zero-weight band columns of the last row read past the plane
  while ((s & 7) != 1) ++s;  // quad planes one bank group apart: best fit for the lane map (roi_tables_kernel)
  return s;
}

template <typename T>
__global__ void __launch_bounds__(NT, 1) roi_align_fwd_band(const Params p, const RoiTab* __restrict__ tabs) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = p.H * p.W;
  const int qs = quad_stride(HW);
  float4* slab = reinterpret_cast<float4*>(smem_raw);
  const size_t slab_bytes = ((size_t)NQ * qs * sizeof(float4) + 127) / 128 * 128;
  WarpArea<T>* areas = reinterpret_cast<WarpArea<T>*>(smem_raw + slab_bytes);
  __shared__ int s_next;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  WarpArea<T>* wa = areas + warp;
  const int nslab = p.C / CS;
  const T* feat = reinterpret_cast<const T*>(p.feat);
  T* out = reinterpret_cast<T*>(p.out);
  uint32_t parity = 0;
  int ycur = 0;  // y-table buffer of the RoI being computed

  for (int i = tid; i < NQ * (qs - HW); i += NT) {
    const int q = i / (qs - HW), o = i - q * (qs - HW);
    slab[(size_t)q * qs + HW + o] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (lane == 0) mbar_init(&wa->bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");

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
    load_slab_quad<T>(feat + ((long long)n * p.C + (long long)k * CS) * HW, slab, HW, qs, tid);
    __syncthreads();

    auto fetch = [&]() -> int {
      int r = 0;
      if (lane == 0) r = atomicAdd(&s_next, 1);
      return __shfl_sync(0xffffffffu, r, 0);
    };
    auto issue_tab = [&](int r, int yb) {  // x-part into the single buffer, y-part into y buffer `yb`
      if (lane == 0) {
        mbar_expect_tx(&wa->bar, (uint32_t)sizeof(RoiTab));
        bulk_load(&wa->tx, &tabs[r_base + r].x, (uint32_t)sizeof(RoiTabX), &wa->bar);
        bulk_load(&wa->ty[yb], &tabs[r_base + r].y, (uint32_t)sizeof(RoiTabY), &wa->bar);
      }
    };
    int cur = fetch();
    if (cur < r1) issue_tab(cur, ycur);
    while (cur < r1) {
      mbar_wait(&wa->bar, parity);
      parity ^= 1u;
      // ---- consume the x-part into registers; the y-part stays in its buffer for the walk
      const RoiTabX* t = &wa->tx;
      const int mode = t->mode, gw = t->gw, gh = t->gh, nb = t->nb, y0 = t->y0, nrows = t->nrows;
      // which (output column, channel quad) this lane owns for this RoI; 4 lanes shadow an owner and never store
      const int lm = t->lmap[lane];
      const bool writer = (lm & 0x80) == 0;
      const int half = (lm >> 4) & 1, pw = lm & 15;
      const float4* quad = slab + (size_t)half * qs;
#ifndef UNIT_FWD_DIRECT
      T* stage_lane = wa->stage + (4 * half) * (P * P) + pw;
#endif
      const int bx = t->bx[pw];
      float w[MAXB];
#pragma unroll
      for (int j = 0; j < MAXB; ++j) w[j] = t->wb[j][pw];
      float inv_count = 0.f, start_w = 0.f, start_h = 0.f, bin_w = 0.f, bin_h = 0.f;
      if (mode == 2) {
        inv_count = t->inv_count;
        start_w = t->start_w;
        start_h = t->start_h;
        bin_w = t->bin_w;
        bin_h = t->bin_h;
      }
      const float2* yt = wa->ty[ycur].yt;
      __syncwarp();  // the x buffer is free: fetch the next RoI and start loading its tables
      const int nxt = fetch();
      if (nxt < r1) issue_tab(nxt, ycur ^ 1);
      ycur ^= 1;
#ifndef UNIT_FWD_DIRECT
      if (lane == 0) v2::bulk_wait_read_all();  // the previous bulk store has drained this warp's staging block
      __syncwarp();
#else
      T* stage_lane = out + ((long long)(r_base + cur) * p.C + (long long)k * CS + 4 * half) * (P * P) + pw;
#endif
      if (!(p.debug & 1)) {
        if (mode == 1) {
          const float4* rowp = quad + (size_t)y0 * p.W + bx;
#define UNIT_BAND_CASE(NBV)                                                                            \
  {                                                                                                    \
    if (gh == 1) task_band<T, NBV, 1>(rowp, p.W, nrows, gh, nb, w, yt, stage_lane, writer);            \
    else if (gh == 2) task_band<T, NBV, 2>(rowp, p.W, nrows, gh, nb, w, yt, stage_lane, writer);       \
    else task_band<T, NBV, 0>(rowp, p.W, nrows, gh, nb, w, yt, stage_lane, writer);                    \
  }
          if (nb <= 2) UNIT_BAND_CASE(2)
          else if (nb == 3) UNIT_BAND_CASE(3)
          else if (nb == 4) UNIT_BAND_CASE(4)
          else if (nb == 5) UNIT_BAND_CASE(5)
          else UNIT_BAND_CASE(0)
#undef UNIT_BAND_CASE
        } else if (mode == 2) {
          if (writer) task_direct<T>(quad, p.H, p.W, gw, gh, inv_count, start_w, start_h, bin_w, bin_h, pw, stage_lane);
        } else if (writer) {
          for (int i = 0; i < P; ++i) {
#pragma unroll
            for (int c = 0; c < 4; ++c) stf(stage_lane + c * (P * P) + i * P, 0.f);
          }
        }
      }
#ifndef UNIT_FWD_DIRECT
      v2::fence_async_smem();
      __syncwarp();
      if (lane == 0 && !(p.debug & 2)) {
        v2::bulk_store(out + ((long long)(r_base + cur) * p.C + (long long)k * CS) * (P * P), wa->stage,
                       (uint32_t)(CS * P * P * sizeof(T)));
        v2::bulk_commit();
      }
#endif
      cur = nxt;
    }
    u = seg_end_u;
  }
#ifndef UNIT_FWD_DIRECT
  if (lane == 0) v2::bulk_wait_all();
#endif
}

template <typename T>
static size_t smem_total(int HW) {
  const size_t slab_bytes = ((size_t)NQ * quad_stride(HW) * sizeof(float4) + 127) / 128 * 128;
  return slab_bytes + (size_t)NW * sizeof(WarpArea<T>);
}

}  // namespace band

bool fwd_band_fits(int C, int H, int W, int dtype) {
  if (C % band::CS) return false;
  const size_t need = dtype == UNIT_F32 ? band::smem_total<float>(H * W) : band::smem_total<__nv_bfloat16>(H * W);
  return need <= 227 * 1024 - 64;
}

size_t fwd_band_workspace_bytes(int R) { return (size_t)(R > 0 ? R : 1) * sizeof(band::RoiTab); }

template <typename T>
static int launch_band_t(const void* feat, const float* rois, void* out, void* tabs_ws, int N, int C, int H, int W,
                         int R, float scale, int sr, int aligned, int* img_off, cudaStream_t st) {
  band::RoiTab* tabs = reinterpret_cast<band::RoiTab*>(tabs_ws);
  band::roi_tables_kernel<<<cdiv(R, 4), 128, 0, st>>>(rois, tabs, R, N, H, W, scale, sr, aligned, img_off);
  UNIT_CHECK_LAUNCH("roi_tables_kernel");
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
  p.pair_stride = 0;
  p.units_total = (long long)R * (C / band::CS);
  p.debug = switches().roi_debug;
  const size_t smem = band::smem_total<T>(H * W);
  UNIT_CUDA(cudaFuncSetAttribute(band::roi_align_fwd_band<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = p.units_total / 48;  // at least 4 RoIs per warp
  if (grid < 1) grid = 1;
  if (grid > sm_count()) grid = sm_count();
  band::roi_align_fwd_band<T><<<(int)grid, band::NT, smem, st>>>(p, tabs);
  UNIT_CHECK_LAUNCH("roi_align_fwd_band");
  return UNIT_OK;
}

int launch_fwd_band(const void* feat, const float* rois, void* out, void* tabs_ws, int N, int C, int H, int W, int R,
                    float scale, int sr, int aligned, int dtype, int* img_off, cudaStream_t st) {
  if (dtype == UNIT_F32)
    return launch_band_t<float>(feat, rois, out, tabs_ws, N, C, H, W, R, scale, sr, aligned, img_off, st);
  return launch_band_t<__nv_bfloat16>(feat, rois, out, tabs_ws, N, C, H, W, R, scale, sr, aligned, img_off, st);
}

}  // namespace roi
}  // namespace unit
