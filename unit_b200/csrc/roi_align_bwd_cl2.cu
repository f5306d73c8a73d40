// ROIAlign backward, channel-lane kernel, table-driven form (fp32 I/O, C % 64 == 0,
// 14 x 14 bins).  Same data flow as roi_align_bwd_cl.cu -- a lane owns two channels, the warp executes one uniform walk
// over the RoI's y- and x-samples, grad_out arrives as [64 ch x 2 bin rows] TMA tiles, every (RoI, footprint cell,
// channel pair) is reduced once into a channel-last fp32 image that lives in L2 -- but the per-item work around the
// arithmetic is gone:
//   * a pre-kernel builds each RoI's sample tables ONCE (the previous kernel rebuilt them for each of the C/64 channel
//     blocks, loading the RoI and running roi_geom per item): x- and y-samples as (h, l) pairs whose mantissa LSB says
//     "the window advances after this sample"; they arrive per item by TMA bulk load into a double-buffered table area;
//   * the walks are compact nested loops (bin -> sample) reading one broadcast LDS.64 per sample: no ballots, no
//     shuffles, no per-sample bookkeeping of (bin, sample-in-bin), ~1/3 of the code size (the old kernel's 62 KB of
//     SASS thrashed the 32 KB instruction cache: 16 % of its stall samples were stall_no_inst).
// Exactly one vector reduction per (RoI, footprint cell, channel pair), as before.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "roi_common.cuh"

namespace unit {
namespace roi {
namespace cl2 {

constexpr int P = 14;
constexpr int CB = 64;                 // channels per work item (two per lane: l and l + 32)
constexpr int CHUNK_ROWS = 2;          // bin rows per TMA tile
constexpr int CHUNK_F = CHUNK_ROWS * P;
constexpr int NCH = P / CHUNK_ROWS;    // tiles per item
constexpr int NST = 2;                 // tile buffers per warp
#ifndef UNIT_BWD2_NW
#define UNIT_BWD2_NW 12
#endif
constexpr int NW = UNIT_BWD2_NW;
constexpr int NT = NW * 32;
constexpr int MAXG = 6;                // sampling grid with tables: RoI side <= 84 feature cells
constexpr int MAXS = P * MAXG;
constexpr uint32_t TILE_BYTES = CB * CHUNK_F * sizeof(float);
constexpr uint32_t F_ADV = 1u;

// per-RoI tables (global, built once per launch)
struct __align__(16) RoiTab {
  int mode, gw, gh, n;      // mode 0: nothing to do, 1: tables, 2: direct evaluation
  int x0, y0, pad0, pad1;   // first footprint cell
  float inv_count, start_w, start_h, bin_w, bin_h, padf[3];
  float2 xt[MAXS];          // x-sample s: (hx / count | F_ADV, lx / count); F_ADV: last sample whose lower tap is its column
  float2 yt[MAXS + 4];      // y-sample s: (hy | F_ADV, ly);                F_ADV: last sample whose lower tap is its row
};
static_assert(sizeof(RoiTab) == 64 + 8 * MAXS + 8 * (MAXS + 4), "RoiTab layout");
static_assert(sizeof(RoiTab) % 16 == 0, "RoiTab must be a multiple of 16 bytes (TMA bulk copy)");

struct __align__(128) WarpArea {
  float stage[NST][CB * CHUNK_F];
  RoiTab tab[2];
  uint64_t bar[NST];
  uint64_t tbar[2];
};

struct Params {
  const float* rois;
  float* scratch;  // [N][H*W][C] fp32, channel (64b + l) at 64b + 2l, channel (64b + 32 + l) at 64b + 2l + 1
  int* counter;
  const RoiTab* tabs;
  int N, C, H, W, R;
  float scale;
  int sampling_ratio, aligned;
  int evict_first;
};

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
__device__ __forceinline__ void tma_tile_load(void* sdst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(sdst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

typedef unsigned long long f2;  // two packed fp32: (channel l, channel l + 32) of this lane
__device__ __forceinline__ f2 pack2(float a, float b) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void fma2(f2& acc, f2 a, f2 b) {  // acc += a * b  (both halves)
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ f2 fma2r(f2 a, f2 b, f2 c) {  // a * b + c
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ void red2(char* addr, f2 v) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(x), "f"(y) : "memory");
}

// ------------------------------------------------------------------------------------------------ table pre-kernel
// One axis of one RoI, whole warp.  Entry s = (h * ws | F_ADV, l * ws); samples clamped to the last cell (value
// F[size-1]) are re-expressed as lo = size-2 with weights (0, h), so lo + 1 is always inside.  Returns true when
// consecutive lower taps are not 0 or 1 apart (sampling_ratio > 0 on large RoIs, fp32 rounding): direct path.
__device__ __forceinline__ bool build_axis(float start, float bin, int g, int size, float ws, float2* tab,
                                           int* first_lo, int lane) {
  constexpr int ROUNDS = (MAXS + 31) / 32;
  const int ns = P * g;
  const float inv_g = 1.f / (float)g;
  int lo[ROUNDS];
  float hw[ROUNDS], lw[ROUNDS];
#pragma unroll
  for (int k = 0; k < ROUNDS; ++k) {
    lo[k] = 0x3fffffff;
    hw[k] = lw[k] = 0.f;
    const int s = 32 * k + lane;
    if (s < ns) {
      int hi;
      float l, h;
      const int pb = (int)(((float)s + 0.5f) * inv_g);  // s / g
      axis_tap(sample_coord(start, bin, pb, s - pb * g, g), size, lo[k], hi, l, h);
      if (lo[k] >= size - 1) {
        lo[k] = size - 2;
        l = h;
        h = 0.f;
      }
      hw[k] = h * ws;
      lw[k] = l * ws;
    }
  }
  bool jump = false;
#pragma unroll
  for (int k = 0; k < ROUNDS; ++k) {
    if (32 * k < ns) {  // warp-uniform
      const int s = 32 * k + lane;
      int nlo = __shfl_down_sync(0xffffffffu, lo[k], 1);
      const int nx = k + 1 < ROUNDS ? __shfl_sync(0xffffffffu, lo[k + 1 < ROUNDS ? k + 1 : k], 0) : 0x3fffffff;
      if (lane == 31) nlo = nx;
      const bool adv = s < ns && (s == ns - 1 || nlo != lo[k]);
      jump |= (s < ns - 1) && (nlo - lo[k] > 1 || nlo < lo[k]);
      if (s < ns) tab[s] = make_float2(__uint_as_float((__float_as_uint(hw[k]) & ~1u) | (adv ? F_ADV : 0u)), lw[k]);
      if (s == 0) *first_lo = lo[k];
    }
  }
  return jump;
}

// One launch prepares the main kernel's inputs: blocks [0, tab_blocks) build the RoI tables (one warp per RoI), the
// others clear the workspace head (work counter + channel-last image) -- two kernels before, back to back on the
// critical path of the step.
__global__ void __launch_bounds__(128)
bwd_prepare_kernel(const float* __restrict__ rois, RoiTab* __restrict__ tabs, int R, int N, int H, int W, float scale,
                   int sampling_ratio, int aligned, int tab_blocks, float4* __restrict__ zero, long long n4) {
  if ((int)blockIdx.x >= tab_blocks) {
    const long long stride = (long long)(gridDim.x - tab_blocks) * blockDim.x;
    for (long long i = (long long)(blockIdx.x - tab_blocks) * blockDim.x + threadIdx.x; i < n4; i += stride)
      zero[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* roi = rois + (long long)r * 5;
  RoiTab* t = tabs + r;
  const int n = (int)roi[0];
  const Geom g = roi_geom(roi, scale, P, P, sampling_ratio, aligned);
  int mode = 1;
  if (g.gw <= 0 || g.gh <= 0 || n < 0 || n >= N) mode = 0;
  else if (g.gw > MAXG || g.gh > MAXG) mode = 2;
  if (mode == 1) {
    bool jump = build_axis(g.start_w, g.bin_w, g.gw, W, 1.f / g.count, t->xt, &t->x0, lane);
    jump |= build_axis(g.start_h, g.bin_h, g.gh, H, 1.f, t->yt, &t->y0, lane);
    if (__any_sync(0xffffffffu, jump)) mode = 2;
    // closing y entry: zero weights + F_ADV, so the row that only received upper taps is swept inside the sample loop
    if (lane == 0) t->yt[P * g.gh] = make_float2(__uint_as_float(F_ADV), 0.f);
  }
  if (lane == 0) {
    t->mode = mode;
    t->gw = g.gw;
    t->gh = g.gh;
    t->n = n;
    t->inv_count = 1.f / g.count;
    t->start_w = g.start_w;
    t->start_h = g.start_h;
    t->bin_w = g.bin_w;
    t->bin_h = g.bin_h;
  }
}

// Rare path (sampling grid > MAXG or a sample step > 1 cell): scalar taps of one bin row straight from the staged tile.
__device__ __noinline__ void direct_row(const Params& p, const RoiTab* t, const float* ta, const float* tb, float* img,
                                        int ph) {
  for (int pw = 0; pw < P; ++pw) {
    const float gA = ta[pw] * t->inv_count;
    const float gB = tb[pw] * t->inv_count;
    for (int iy = 0; iy < t->gh; ++iy) {
      int ylo, yhi;
      float ly, hy;
      const bool vy = axis_tap(sample_coord(t->start_h, t->bin_h, ph, iy, t->gh), p.H, ylo, yhi, ly, hy);
      for (int ix = 0; ix < t->gw; ++ix) {
        int xlo, xhi;
        float lx, hx;
        const bool vx = axis_tap(sample_coord(t->start_w, t->bin_w, pw, ix, t->gw), p.W, xlo, xhi, lx, hx);
        if (vy && vx) {
          red2(reinterpret_cast<char*>(img + ((size_t)ylo * p.W + xlo) * p.C), pack2(gA * hy * hx, gB * hy * hx));
          red2(reinterpret_cast<char*>(img + ((size_t)ylo * p.W + xhi) * p.C), pack2(gA * hy * lx, gB * hy * lx));
          red2(reinterpret_cast<char*>(img + ((size_t)yhi * p.W + xlo) * p.C), pack2(gA * ly * hx, gB * ly * hx));
          red2(reinterpret_cast<char*>(img + ((size_t)yhi * p.W + xhi) * p.C), pack2(gA * ly * lx, gB * ly * lx));
        }
      }
    }
  }
}

// One x-sample of the horizontal sweep as ONE block of PTX, so that the window keeps its registers on both paths and
// the (warp-uniform) branch needs no reconvergence bookkeeping:   c0 += hx * v;  c1 += lx * v;
// last sample of its column (F_ADV): reduce c0 into *cell, cell += cstep, c0 = c1, c1 = 0.
__device__ __forceinline__ void sweep_step(f2& c0, f2& c1, char*& cell, float2 e, f2 v, long long cstep,
                                           uint64_t red_policy) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      ".reg .b64 eh, el;\n"
      ".reg .b32 fl;\n"
      ".reg .f32 ra, rb;\n"
      "mov.b64 eh, {%3, %3};\n"
      "mov.b64 el, {%4, %4};\n"
      "fma.rn.f32x2 %0, eh, %5, %0;\n"
      "fma.rn.f32x2 %1, el, %5, %1;\n"
      "mov.b32 fl, %3;\n"
      "and.b32 fl, fl, 1;\n"
      "setp.eq.u32 q, fl, 0;\n"
      "@q bra.uni SWEEP_NEXT;\n"
#ifndef UNIT_BWD_NORED
      "mov.b64 {ra, rb}, %0;\n"
#ifndef UNIT_BWD2_NO_RED_HINT
      "red.global.add.L2::cache_hint.v2.f32 [%2], {ra, rb}, %7;\n"  // evict_last: the image outlives the tile stream
#else
      "red.global.add.v2.f32 [%2], {ra, rb};\n"
#endif
#endif
      "add.s64 %2, %2, %6;\n"
      "mov.b64 %0, %1;\n"
      "mov.b64 %1, 0;\n"
      "SWEEP_NEXT:\n"
      "}\n"
      : "+l"(c0), "+l"(c1), "+l"(cell)
      : "f"(e.x), "f"(e.y), "l"(v), "l"(cstep), "l"(red_policy)
      : "memory");
}

// The same step for the fully unrolled sweeps.  No "memory" clobber: the reduction targets an image this kernel never
// reads, and without the clobber the table loads of later samples are hoisted over the block (the loop form prefetches
// one entry by hand and pays two register moves per sample for it).
__device__ __forceinline__ void sweep_step_u(f2& c0, f2& c1, char*& cell, float2 e, f2 v, long long cstep,
                                             uint64_t red_policy) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      ".reg .b64 eh, el;\n"
      ".reg .b32 fl;\n"
      ".reg .f32 ra, rb;\n"
      "mov.b64 eh, {%3, %3};\n"
      "mov.b64 el, {%4, %4};\n"
      "fma.rn.f32x2 %0, eh, %5, %0;\n"
      "fma.rn.f32x2 %1, el, %5, %1;\n"
      "mov.b32 fl, %3;\n"
      "and.b32 fl, fl, 1;\n"
      "setp.eq.u32 q, fl, 0;\n"
      "@q bra.uni SWEEP_NEXT_U;\n"
#ifndef UNIT_BWD_NORED
      "mov.b64 {ra, rb}, %0;\n"
#ifndef UNIT_BWD2_NO_RED_HINT
      "red.global.add.L2::cache_hint.v2.f32 [%2], {ra, rb}, %7;\n"
#else
      "red.global.add.v2.f32 [%2], {ra, rb};\n"
#endif
#endif
      "add.s64 %2, %2, %6;\n"
      "mov.b64 %0, %1;\n"
      "mov.b64 %1, 0;\n"
      "SWEEP_NEXT_U:\n"
      "}\n"
      : "+l"(c0), "+l"(c1), "+l"(cell)
      : "f"(e.x), "f"(e.y), "l"(v), "l"(cstep), "l"(red_policy));
}

// Sweep with a compile-time sample count per bin (gw = 1: RoIs up to 224 px wide, gw = 2: up to 448 px): straight-line
// code, table entries at constant offsets.
template <int GW>
__device__ __forceinline__ void sweep_t(const float2* __restrict__ xt, const f2 (&v)[P], char* cell, long long cstep,
                                        uint64_t red_policy) {
  f2 c0 = 0ull, c1 = 0ull;
#pragma unroll
  for (int pw = 0; pw < P; ++pw) {
#pragma unroll
    for (int ix = 0; ix < GW; ++ix) sweep_step_u(c0, c1, cell, xt[pw * GW + ix], v[pw], cstep, red_policy);
  }
#ifndef UNIT_BWD_NORED
  red2(cell, c0);  // the column that only received upper taps
#endif
}

// Horizontal sweep of one finished feature row: v[pw] (two channels per lane) is spread over the row's cells with a
// two-column window (c0, c1); c0 is complete when the lower tap column advances -> one red.v2 per (cell, lane).
__device__ __forceinline__ void sweep(const float2* __restrict__ xt, const f2 (&v)[P], char* cell, long long cstep,
                                      int gw, uint64_t red_policy) {
  f2 c0 = 0ull, c1 = 0ull;
  float2 en = *xt;
#pragma unroll
  for (int pw = 0; pw < P; ++pw) {
#pragma unroll 1
    for (int ix = 0; ix < gw; ++ix) {
      const float2 e = en;
      en = *++xt;  // one entry ahead (the table has spare entries past the last sample)
      sweep_step(c0, c1, cell, e, v[pw], cstep, red_policy);
    }
  }
#ifndef UNIT_BWD_NORED
  red2(cell, c0);  // the column that only received upper taps
#endif
}

__global__ void __launch_bounds__(NT, 1) roi_align_bwd_cl2(const __grid_constant__ CUtensorMap gmap, const Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WarpArea* wa = reinterpret_cast<WarpArea*>(smem_raw) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int nblk = p.C / CB;
  const long long n_items = (long long)p.R * nblk;
  uint64_t policy;
  if (p.evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(policy));
  uint64_t red_policy;  // the 34 MB gradient image stays resident (evict_last) against the 0.8 GB tile stream
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(red_policy));
  if (lane == 0) {
    for (int b = 0; b < NST; ++b) mbar_init(&wa->bar[b], 1);
    mbar_init(&wa->tbar[0], 1);
    mbar_init(&wa->tbar[1], 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

  auto fetch = [&]() -> long long {
    long long it = 0;
    if (lane == 0) it = atomicAdd(p.counter, 1);
    return __shfl_sync(0xffffffffu, it, 0);
  };
  auto issue = [&](long long it, int k, int b) {  // tile k of item `it` -> buffer b (lane 0 only)
    const int r = (int)(it / nblk);
    const int cb = (int)(it - (long long)r * nblk);
    mbar_expect_tx(&wa->bar[b], TILE_BYTES);
    tma_tile_load(wa->stage[b], &gmap, &wa->bar[b], CHUNK_F * k, r * p.C + cb * CB, policy);
  };
  auto refill = [&](long long cur, long long nxt, int k, int b) {  // buffer b gets the tile NST ahead (lane 0 only)
    if (k + NST < NCH) issue(cur, k + NST, b);
    else if (nxt < n_items) issue(nxt, k + NST - NCH, b);
  };
  auto issue_tab = [&](long long it, int tb) {  // tables of item `it` -> table buffer tb (lane 0 only)
    const int r = (int)(it / nblk);
    mbar_expect_tx(&wa->tbar[tb], (uint32_t)sizeof(RoiTab));
    bulk_load(&wa->tab[tb], p.tabs + r, (uint32_t)sizeof(RoiTab), &wa->tbar[tb]);
  };

  long long cur = fetch();
  long long nxt = cur < n_items ? fetch() : n_items;
  if (lane == 0 && cur < n_items) {
    for (int k = 0; k < NST; ++k) issue(cur, k, k);
    issue_tab(cur, 0);
  }
  uint32_t cc = 0;     // tiles consumed by this warp: buffer = cc % NST, mbarrier phase parity = (cc / NST) & 1
  uint32_t items = 0;  // items consumed: table buffer = items & 1, parity = (items >> 1) & 1
  const long long cstep = (long long)p.C * 4;
  const long long rstep = cstep * p.W;
  while (cur < n_items) {
    const int tb = items & 1;
    if (lane == 0 && nxt < n_items) issue_tab(nxt, tb ^ 1);  // its previous user (item - 1) is finished
    mbar_wait(&wa->tbar[tb], (items >> 1) & 1u);
    ++items;
    const RoiTab* t = &wa->tab[tb];
    const int r = (int)(cur / nblk);
    const int cb = (int)(cur - (long long)r * nblk);
    const int mode = t->mode;
    float* img = p.scratch + (size_t)(mode ? t->n : 0) * p.H * p.W * p.C + cb * CB + 2 * lane;

    if (mode == 1) {
      const int gw = t->gw, gh = t->gh;
      char* rowp = reinterpret_cast<char*>(img) + ((long long)t->y0 * p.W + t->x0) * cstep;  // first footprint cell
      const float2* xt = t->xt;
      const float2* yt = t->yt;
      f2 vlo[P], vhi[P];
#pragma unroll
      for (int i = 0; i < P; ++i) vlo[i] = vhi[i] = 0ull;
      float2 en = *yt;
      int b = 0;
#pragma unroll 1
      for (int ph = 0; ph < P; ++ph) {
        // ---- this bin row's 14 gradients of the lane's two channels
        const int half = ph & (CHUNK_ROWS - 1);
        if (half == 0) {
          b = cc % NST;
          mbar_wait(&wa->bar[b], (cc / NST) & 1u);
          ++cc;
        }
        f2 g2[P];
        {
          const float* ta = wa->stage[b] + lane * CHUNK_F + half * P;
#pragma unroll
          for (int j = 0; j < P; ++j) g2[j] = pack2(ta[j], ta[j + 32 * CHUNK_F]);
        }
        if (half == CHUNK_ROWS - 1) {
          __syncwarp();  // every lane is done with the tile: refill the buffer NST tiles ahead in the stream
          if (lane == 0) refill(cur, nxt, ph / CHUNK_ROWS, b);
        }
        const int nsb = gh + (ph == P - 1 ? 1 : 0);  // the last bin also takes the closing entry of the y-table
#pragma unroll 1
        for (int i = 0; i < nsb; ++i) {
          const float2 e = en;
          en = *++yt;
          const f2 hy = pack2(e.x, e.x), ly = pack2(e.y, e.y);
#pragma unroll
          for (int pw = 0; pw < P; ++pw) fma2(vlo[pw], hy, g2[pw]);
          if (__float_as_uint(e.x) & F_ADV) {  // last sample whose lower tap is this feature row: sweep it
            if (gw == 1) sweep_t<1>(xt, vlo, rowp, cstep, red_policy);
            else if (gw == 2) sweep_t<2>(xt, vlo, rowp, cstep, red_policy);
            else sweep(xt, vlo, rowp, cstep, gw, red_policy);
            rowp += rstep;
#pragma unroll
            for (int pw = 0; pw < P; ++pw) {
              vlo[pw] = fma2r(ly, g2[pw], vhi[pw]);
              vhi[pw] = 0ull;
            }
          } else {
#pragma unroll
            for (int pw = 0; pw < P; ++pw) fma2(vhi[pw], ly, g2[pw]);
          }
        }
      }
    } else {
      // degenerate / foreign RoIs (mode 0) only drain their tiles; mode 2 evaluates every tap directly
#pragma unroll 1
      for (int k = 0; k < NCH; ++k) {
        const int b = cc % NST;
        mbar_wait(&wa->bar[b], (cc / NST) & 1u);
        ++cc;
        if (mode == 2) {
          for (int half = 0; half < CHUNK_ROWS; ++half) {
            const float* ta = wa->stage[b] + lane * CHUNK_F + half * P;
            direct_row(p, t, ta, ta + 32 * CHUNK_F, img, CHUNK_ROWS * k + half);
          }
        }
        __syncwarp();
        if (lane == 0) refill(cur, nxt, k, b);
      }
    }
    __syncwarp();  // every lane has finished reading this item's tables before the buffer is handed on
    cur = nxt;
    nxt = cur < n_items ? fetch() : n_items;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

}  // namespace cl2

size_t bwd_cl2_table_bytes(int R) { return (size_t)(R > 0 ? R : 1) * sizeof(cl2::RoiTab) + 256; }

// kernels of roi_align_bwd_cl.cu reused for the epilogue
void launch_bwd_cl_unpermute(const float* scratch, void* gfeat, int N, int C, int HW, bool bf16, cudaStream_t st);

// workspace: [256 B counter][N*H*W*C fp32 channel-last image (+64 B)][R tables]
int launch_bwd_cl2(const void* gout, const float* rois, void* gfeat, void* ws, int N, int C, int H, int W, int R,
                   float scale, int sr, int aligned, cudaStream_t st) {
  using namespace cl2;
  CUtensorMap map;
  if (int rc = ensure_driver_context(gout)) return rc;
  {
    cuuint64_t dims[2] = {(cuuint64_t)(P * P), (cuuint64_t)R * C};
    cuuint64_t strides[1] = {(cuuint64_t)(P * P) * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)CHUNK_F, (cuuint32_t)CB};
    cuuint32_t estr[2] = {1, 1};
    const int promo = switches().bwd_promo;
    const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                     : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                     : promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                  : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    CUresult rc = encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(gout), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UNIT_REQUIRE(rc == CUDA_SUCCESS, "roi_align_bwd: cuTensorMapEncodeTiled failed (%d)", (int)rc);
  }
  const long long total = (long long)N * C * H * W;  // multiple of 64
  const size_t img_bytes = ((size_t)total * 4 + 256 + 255) / 256 * 256;
  Params p;
  p.rois = rois;
  p.counter = (int*)ws;
  p.scratch = (float*)((char*)ws + 256);
  RoiTab* tabs = reinterpret_cast<RoiTab*>((char*)ws + 256 + img_bytes);
  p.tabs = tabs;
  p.N = N;
  p.C = C;
  p.H = H;
  p.W = W;
  p.R = R;
  p.scale = scale;
  p.sampling_ratio = sr;
  p.aligned = aligned;
  // tiles: evict_normal, so the 64-byte DRAM granules two neighbouring tiles share are fetched once (with the image
  // pinned by the reductions' evict_last hint): DRAM read 1.17 GB -> 0.94 GB, same time (profiles/r02_roi_align_bwd.md)
  p.evict_first = switches().bwd2_evict_first;
  {
    const int tab_blocks = cdiv(R, 4);
    const long long n4 = total / 4 + 16;
    const int zero_blocks = (int)std::min<long long>((n4 + 127) / 128, (long long)sm_count() * 16);
    bwd_prepare_kernel<<<tab_blocks + zero_blocks, 128, 0, st>>>(rois, tabs, R, N, H, W, scale, sr, aligned, tab_blocks,
                                                                  (float4*)ws, n4);
    UNIT_CHECK_LAUNCH("bwd_prepare_kernel");
  }
  const size_t smem = (size_t)NW * sizeof(WarpArea);
  const long long items = (long long)R * (C / CB);
  long long grid = (items + NW - 1) / NW;
  if (grid > sm_count()) grid = sm_count();
  if (grid < 1) grid = 1;
  UNIT_CUDA(cudaFuncSetAttribute(roi_align_bwd_cl2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  roi_align_bwd_cl2<<<(int)grid, NT, smem, st>>>(map, p);
  UNIT_CHECK_LAUNCH("roi_align_bwd_cl2");
  launch_bwd_cl_unpermute(p.scratch, gfeat, N, C, H * W, false, st);
  UNIT_CHECK_LAUNCH("unpermute_kernel");
  return UNIT_OK;
}

}  // namespace roi
}  // namespace unit
