// ROIAlign backward, channel-lane kernel (fp32 or bf16 I/O, C % 64 == 0, 14 x 14 bins).
//
// The gradient of one RoI is separable:  dF[c][y][x] = sum_ph sum_pw Wy[ph][y] * g[c][ph][pw] * Wx[pw][x], with the SAME
// Wy / Wx for every channel.  So a lane owns CHANNELS (two of them), and the whole warp executes one warp-uniform
// walk over the RoI's y- and x-samples: no lane ever idles because a RoI is narrow, no branch diverges, and the table
// entries are broadcast shared-memory loads.
//
//   work item   (RoI, block of 64 channels), pulled from a global counter by the 12 warps of 148 persistent CTAs;
//               consecutive items are the 16 channel blocks of the same RoI, so concurrent warps never collide.
//   input       grad_out[r][c0 .. c0+63][2k .. 2k+1][0 .. 13] -- a [64 x 28] fp32 tile (7168 B) per 2 bin rows --
//               arrives by TMA (cp.async.bulk.tensor.2d, L2 evict-first), double buffered per warp, across items.
//               Lane l reads rows l and l+32 with LDS.128: the 112-byte row pitch makes that conflict free.
//   vertical    per y-sample (hy, ly): Vlo[pw] += hy * g[pw], Vhi[pw] += ly * g[pw]  (14 packed FFMA2 each);
//               when the lower tap row advances, Vlo is complete for that feature row.
//   horizontal  the finished row is swept once over the x-samples with a two-column window (cur, next); when the
//               lower tap column advances, `cur` is complete: ONE red.global.add.v2.f32 per (cell, lane), 256
//               contiguous bytes per warp, into a channel-last fp32 image of grad_feat that lives in L2.
//   epilogue    a tiled transpose turns the channel-last image into NCHW.
//   bf16 I/O    grad_out rows are 392 bytes, not a legal TMA stride, so the tensor map views channel PAIRS (784-byte
//               rows of 2 x 196 bf16) cut into seven 112-byte boxes that arrive through a 4-slot ring (see BF_*
//               below); lane l owns channels 2l and 2l+1, accumulation and the L2 image stay fp32, the transpose
//               rounds to bf16 once.
// Exactly one vector reduction per (RoI, footprint cell, channel pair) replaces torchvision's 4 * gh * gw scalar atomics
// per output element (roi_align_kernel.cu, bilinear_interpolate_gradient + atomicAdd).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include <algorithm>

#include "roi_common.cuh"

namespace unit {
namespace roi {
namespace cl {

constexpr int P = 14;
constexpr int CB = 64;                   // channels per work item (two per lane: l and l + 32)
#ifndef UNIT_BWD_ROWS
#define UNIT_BWD_ROWS 2
#endif
constexpr int CHUNK_ROWS = UNIT_BWD_ROWS;  // bin rows per TMA tile (the last tile of a 4-row split is zero-filled past row 13)
constexpr int CHUNK_F = CHUNK_ROWS * P;    // floats per channel and tile
#ifndef UNIT_BWD_NST
#define UNIT_BWD_NST 2
#endif
#ifndef UNIT_BWD_NW
#define UNIT_BWD_NW 12
#endif
constexpr int NST = UNIT_BWD_NST;        // tile buffers per warp
constexpr int NW = UNIT_BWD_NW;
constexpr int NT = NW * 32;
constexpr int MAXG = 6;                  // sampling grid with tables: RoI side <= 84 feature cells (1344 px at 1/16)
constexpr int MAXS = P * MAXG;
constexpr uint32_t TILE_BYTES = CB * CHUNK_F * sizeof(float);
// bf16 I/O: the tensor map views channel PAIRS (rows of 2 x 196 bf16 = 784 B); a pair row is exactly 7 boxes of 56
// elements (112 B, every box start 16-byte aligned): boxes 0-2 = bin rows 0-11 of the even channel, box 3 = its rows
// 12-13 followed by rows 0-1 of the odd channel, boxes 4-6 = rows 2-13 of the odd channel.  A warp walks the bin rows in
// 7 steps of two; step s needs the boxes (E, O) = (0,3) (0,4) (1,4) (1,5) (2,5) (2,6) (3,6).  They arrive in the order
// 0 3 4 1 5 2 6 3 (box 3 twice) into a ring of 4 slots: arrival a -> slot a & 3, mbarrier parity (a >> 2) & 1.
constexpr int BF_BOX_ELEMS = 4 * P;                          // 56
constexpr uint32_t BF_BOX_BYTES = (CB / 2) * BF_BOX_ELEMS * 2;  // 3584: 32 channel pairs x 112 B
constexpr int BF_PITCH = BF_BOX_ELEMS * 2;                   // bytes per channel pair and box
constexpr int BF_SLOTS = 4;
static_assert(BF_SLOTS * BF_BOX_BYTES <= NST * TILE_BYTES, "the bf16 ring must fit the staging buffers");
constexpr int NBAR = NST > BF_SLOTS ? NST : BF_SLOTS;
template <bool BF>
struct Tile {
  static constexpr int ROWS = BF ? 2 : CHUNK_ROWS;
  static constexpr int NCH = (P + ROWS - 1) / ROWS;
};
__device__ __forceinline__ int bf_box(int a) { return (0x36251430u >> (4 * a)) & 7; }     // arrival -> box
__device__ __forceinline__ int bf_slot_even(int s) { return (0x3113300u >> (4 * s)) & 7; }  // step -> slot of E
__device__ __forceinline__ int bf_slot_odd(int s) { return (0x2200221u >> (4 * s)) & 7; }   // step -> slot of O

struct __align__(128) WarpArea {
  float stage[NST][CB * CHUNK_F];
  float4 xt[MAXS];  // (hx / count, hx / count, lx / count, lx / count)
  float4 yt[MAXS + 4];  // (hy, hy, ly, ly); entry 14*gh is a zero sentinel whose advance bit is set
  uint32_t xadv[4], yadv[4];  // bit s: sample s is the last one whose lower tap is its cell (window advances after it)
  uint64_t bar[NBAR];
  int x0, y0, gw, gh, mode;
  float inv_count, start_w, start_h, bin_w, bin_h;
};

struct Params {
  const float* rois;
  float* scratch;  // [N][H*W][C] fp32, channel (64b + l) at 64b + 2l, channel (64b + 32 + l) at 64b + 2l + 1
  int* counter;
  int N, C, H, W, R;
  float scale;
  int sampling_ratio, aligned;
  int evict_first;  // L2 policy of the grad_out tiles
  int sweep3;       // use the unrolled sweep for gw == 3 as well
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

typedef unsigned long long f2;  // two packed fp32: (channel l, channel l + 32) of this lane
__device__ __forceinline__ f2 pack2(float a, float b) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void fma2(f2& acc, f2 a, f2 b) {  // acc += a * b  (both halves)
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void red2(char* addr, f2 v) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(x), "f"(y) : "memory");
}

// One x-sample of the horizontal sweep, as ONE block of PTX so that the window registers keep their names on both
// paths (no copies at the join):   c0 += eh * v;  c1 += el * v;
// last sample of its column: flush c0 to *cell, cell += cstep, c0 = c1, c1 = 0.
__device__ __forceinline__ void sweep_step(float2& c0, float2& c1, char*& cell, float4 e, float2 v, long long cstep) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      ".reg .b64 a, b, eh, el, vv;\n"
      ".reg .b32 fl;\n"
      "mov.b64 a, {%0, %1};\n"
      "mov.b64 b, {%2, %3};\n"
      "mov.b64 eh, {%5, %6};\n"
      "mov.b64 el, {%7, %8};\n"
      "mov.b64 vv, {%9, %10};\n"
      "fma.rn.f32x2 a, eh, vv, a;\n"
      "fma.rn.f32x2 b, el, vv, b;\n"
      "mov.b64 {%0, %1}, a;\n"
      "mov.b64 {%2, %3}, b;\n"
      "mov.b32 fl, %5;\n"
      "and.b32 fl, fl, 1;\n"
      "setp.eq.u32 q, fl, 0;\n"
      "@q bra.uni SWEEP_NEXT;\n"
#ifndef UNIT_BWD_NORED
      "red.global.add.v2.f32 [%4], {%0, %1};\n"
#endif
      "add.s64 %4, %4, %11;\n"
      "mov.f32 %0, %2;\n"
      "mov.f32 %1, %3;\n"
      "mov.f32 %2, 0f00000000;\n"
      "mov.f32 %3, 0f00000000;\n"
      "SWEEP_NEXT:\n"
      "}\n"
      : "+f"(c0.x), "+f"(c0.y), "+f"(c1.x), "+f"(c1.y), "+l"(cell)
      : "f"(e.x), "f"(e.y), "f"(e.z), "f"(e.w), "f"(v.x), "f"(v.y), "l"(cstep));
}

// Upper-tap half of one y-sample for 7 bins, same idea:  not last sample of its row: hi[i] += ly * g[i];
// last sample (the row was just swept): lo[i] = hi[i] + ly * g[i], hi[i] = 0.
__device__ __forceinline__ void ystep_upper7(f2* lo, f2* hi, const f2* g, f2 ly, uint32_t adv) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.u32 q, %22, 0;\n"
      "@q bra.uni Y_ADV;\n"
      "fma.rn.f32x2 %7, %21, %14, %7;\n"
      "fma.rn.f32x2 %8, %21, %15, %8;\n"
      "fma.rn.f32x2 %9, %21, %16, %9;\n"
      "fma.rn.f32x2 %10, %21, %17, %10;\n"
      "fma.rn.f32x2 %11, %21, %18, %11;\n"
      "fma.rn.f32x2 %12, %21, %19, %12;\n"
      "fma.rn.f32x2 %13, %21, %20, %13;\n"
      "bra.uni Y_END;\n"
      "Y_ADV:\n"
      "fma.rn.f32x2 %0, %21, %14, %7;\n"
      "fma.rn.f32x2 %1, %21, %15, %8;\n"
      "fma.rn.f32x2 %2, %21, %16, %9;\n"
      "fma.rn.f32x2 %3, %21, %17, %10;\n"
      "fma.rn.f32x2 %4, %21, %18, %11;\n"
      "fma.rn.f32x2 %5, %21, %19, %12;\n"
      "fma.rn.f32x2 %6, %21, %20, %13;\n"
      "mov.b64 %7, 0;\n"
      "mov.b64 %8, 0;\n"
      "mov.b64 %9, 0;\n"
      "mov.b64 %10, 0;\n"
      "mov.b64 %11, 0;\n"
      "mov.b64 %12, 0;\n"
      "mov.b64 %13, 0;\n"
      "Y_END:\n"
      "}\n"
      : "+l"(lo[0]), "+l"(lo[1]), "+l"(lo[2]), "+l"(lo[3]), "+l"(lo[4]), "+l"(lo[5]), "+l"(lo[6]), "+l"(hi[0]),
        "+l"(hi[1]), "+l"(hi[2]), "+l"(hi[3]), "+l"(hi[4]), "+l"(hi[5]), "+l"(hi[6])
      : "l"(g[0]), "l"(g[1]), "l"(g[2]), "l"(g[3]), "l"(g[4]), "l"(g[5]), "l"(g[6]), "l"(ly), "r"(adv));
}

// One axis of the RoI (whole warp): table entry s = (h * ws, h * ws, l * ws, l * ws), and bit s of adv[] marks the
// last sample whose lower tap is this cell.  Samples clamped to the last cell (value F[size-1]) are re-expressed as
// lo = size-2 with weights (0, h), so lo + 1 is always inside.  With `sentinel`, entry ns is (0,0,0,0) with its advance
// bit set (flushes the cell that only received upper taps).  Returns true when consecutive lower taps are not 0 or 1
// apart (only possible for sampling_ratio > 0 or fp32 rounding) -- such RoIs take the direct path.
__device__ __forceinline__ bool build_axis(float start, float bin, int g, int size, float ws, float4* tab,
                                           uint32_t* advw, int* first_lo, bool sentinel, int lane) {
  constexpr int ROUNDS = (MAXS + 1 + 31) / 32;
  const int ns = P * g;
  const int ne = ns + (sentinel ? 1 : 0);
  const float inv_g = 1.f / (float)g;
  int lo[ROUNDS];
  float hw[ROUNDS], lw[ROUNDS];
#pragma unroll
  for (int k = 0; k < ROUNDS; ++k) {
    lo[k] = 0x3fffffff;
    hw[k] = lw[k] = 0.f;
    const int s = 32 * k + lane;
    if (32 * k < ns && s < ns) {
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
    if (32 * k < ne) {  // warp-uniform
      const int s = 32 * k + lane;
      int nlo = __shfl_down_sync(0xffffffffu, lo[k], 1);
      const int nx = k + 1 < ROUNDS ? __shfl_sync(0xffffffffu, lo[k + 1 < ROUNDS ? k + 1 : k], 0) : 0x3fffffff;
      if (lane == 31) nlo = nx;
      const bool adv = (s < ns && (s == ns - 1 || nlo != lo[k])) || (sentinel && s == ns);
      jump |= (s < ns - 1) && (nlo - lo[k] > 1 || nlo < lo[k]);
      const uint32_t m = __ballot_sync(0xffffffffu, adv);
      // the advance flag also rides in the mantissa LSB of the first copy of h (<= 1 ulp on one channel's weight)
      if (s < ne)
        tab[s] = make_float4(__uint_as_float((__float_as_uint(hw[k]) & ~1u) | (adv ? 1u : 0u)), hw[k], lw[k], lw[k]);
      if (lane == 0) advw[k] = m;
      if (s == 0) *first_lo = lo[k];
    }
  }
  return jump;
}

// Rare path (sampling grid > MAXG or a sample step > 1 cell): scalar taps straight from the staged tile.
// Rare path (sampling grid > MAXG or a sample step > 1 cell): scalar taps of one bin row, straight from the staged
// tile.  ta / tb: the row's 14 values of the lane's two channels (fp32 or bf16).
template <typename E>
__device__ __noinline__ void direct_row(const Params& p, const WarpArea* wa, const E* ta, const E* tb, float* img,
                                        int ph) {
  for (int pw = 0; pw < P; ++pw) {
    const float gA = (float)ta[pw] * wa->inv_count;
    const float gB = (float)tb[pw] * wa->inv_count;
    for (int iy = 0; iy < wa->gh; ++iy) {
      int ylo, yhi;
      float ly, hy;
      const bool vy = axis_tap(sample_coord(wa->start_h, wa->bin_h, ph, iy, wa->gh), p.H, ylo, yhi, ly, hy);
      for (int ix = 0; ix < wa->gw; ++ix) {
        int xlo, xhi;
        float lx, hx;
        const bool vx = axis_tap(sample_coord(wa->start_w, wa->bin_w, pw, ix, wa->gw), p.W, xlo, xhi, lx, hx);
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

// Horizontal sweep of one finished feature row: v[pw] (two channels per lane) is spread over the row's cells with a
// two-column window; c0 is complete when the lower tap column advances (flag = mantissa LSB of the table entry).
// GW = 1, 2: unrolled; GW = 0: one compact loop for any grid (instruction footprint matters: 32 KB L1.5 I-cache).
template <int GW>
__device__ __forceinline__ void sweep(const float4* __restrict__ xt, const f2 (&v)[P], char* cell, long long cstep,
                                      int gw) {
  float2 c0 = make_float2(0.f, 0.f), c1 = c0;
#pragma unroll
  for (int pw = 0; pw < P; ++pw) {
    float2 vv;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(vv.x), "=f"(vv.y) : "l"(v[pw]));
    if (GW > 0) {
#pragma unroll
      for (int ix = 0; ix < GW; ++ix) sweep_step(c0, c1, cell, xt[pw * GW + ix], vv, cstep);
    } else {
#pragma unroll 1
      for (int ix = 0; ix < gw; ++ix) sweep_step(c0, c1, cell, *xt++, vv, cstep);
    }
  }
#ifndef UNIT_BWD_NORED
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(cell), "f"(c0.x), "f"(c0.y));  // upper-tap-only column
#else
  asm volatile("" ::"l"(cell), "f"(c0.x), "f"(c0.y));
#endif
}

template <bool BF>
__global__ void __launch_bounds__(NT, 1)
roi_align_bwd_cl(const __grid_constant__ CUtensorMap gmap, const Params p) {
  constexpr int ROWS = Tile<BF>::ROWS, NCH = Tile<BF>::NCH;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WarpArea* wa = reinterpret_cast<WarpArea*>(smem_raw) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int nblk = p.C / CB;
  const long long n_items = (long long)p.R * nblk;
  uint64_t policy;
  if (p.evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(policy));
  if (lane == 0) {
    for (int b = 0; b < NBAR; ++b) mbar_init(&wa->bar[b], 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

  auto fetch = [&]() -> long long {
    long long it = 0;
    if (lane == 0) it = atomicAdd(p.counter, 1);
    return __shfl_sync(0xffffffffu, it, 0);
  };
  // tile k of item `it` -> buffer b (lane 0 only)
  auto issue = [&](long long it, int k, int b) {
    const int r = (int)(it / nblk);
    const int cb = (int)(it - (long long)r * nblk);
    mbar_expect_tx(&wa->bar[b], TILE_BYTES);
#ifdef UNIT_BWD_SAMETILE  // timing experiment: every tile comes from the same (L2-resident) place
    tma_tile_load(wa->stage[b], &gmap, &wa->bar[b], 0, (r & 7) * CB, policy);
#else
    tma_tile_load(wa->stage[b], &gmap, &wa->bar[b], CHUNK_F * k, r * p.C + cb * CB, policy);
#endif
  };
  // refill buffer b with the tile NST ahead of tile k of the current item (lane 0 only)
  auto refill = [&](long long cur, long long nxt, int k, int b) {
    if (k + NST < NCH) issue(cur, k + NST, b);
    else if (nxt < n_items) issue(nxt, k + NST - NCH, b);
  };

  // bf16 ring (lane 0 only): arrival a of item `it`; after its last use arrival `dead` hands its slot to arrival dead + 4
  char* const ring = reinterpret_cast<char*>(wa->stage);
  auto bf_issue = [&](long long it, int a) {
    const int r = (int)(it / nblk);
    const int cb = (int)(it - (long long)r * nblk);
    uint64_t* bar = &wa->bar[a & 3];
    mbar_expect_tx(bar, BF_BOX_BYTES);
    tma_tile_load(ring + (a & 3) * BF_BOX_BYTES, &gmap, bar, BF_BOX_ELEMS * bf_box(a), (r * p.C + cb * CB) >> 1, policy);
  };
  auto bf_refill = [&](long long cur, long long nxt, int dead) {
    if (dead + 4 < 8) bf_issue(cur, dead + 4);
    else if (nxt < n_items) bf_issue(nxt, dead - 4);
  };
  auto bf_wait = [&](int s) {  // the arrivals first used by step s
    if (s == 0) {
      mbar_wait(&wa->bar[0], 0u);
      mbar_wait(&wa->bar[1], 0u);
    } else {
      mbar_wait(&wa->bar[(s + 1) & 3], (uint32_t)((s + 1) >> 2) & 1u);
    }
  };
  auto bf_release = [&](long long cur, long long nxt, int s) {  // arrivals whose last use was step s
    bf_refill(cur, nxt, s == 0 ? 1 : (s == 1 ? 0 : s));
    if (s == 6) bf_refill(cur, nxt, 7);
  };

  long long cur = fetch();
  long long nxt = cur < n_items ? fetch() : n_items;
  if (lane == 0 && cur < n_items) {
    if (BF) {
      for (int a = 0; a < BF_SLOTS; ++a) bf_issue(cur, a);
    } else {
      for (int k = 0; k < NST; ++k) issue(cur, k, k);
    }
  }
  uint32_t cc = 0;  // tiles consumed by this warp: buffer = cc % NST, mbarrier phase parity = (cc / NST) & 1
  while (cur < n_items) {
    const int r = (int)(cur / nblk);
    const int cb = (int)(cur - (long long)r * nblk);
    const float* roi = p.rois + (long long)r * 5;
    const int n = (int)roi[0];
    // ---- tables of this RoI
    const Geom g = roi_geom(roi, p.scale, P, P, p.sampling_ratio, p.aligned);
    int mode = 1;
    if (g.gw <= 0 || g.gh <= 0 || n < 0 || n >= p.N) mode = 0;
    else if (g.gw > MAXG || g.gh > MAXG) mode = 2;
    if (mode == 1) {
      bool jump = build_axis(g.start_w, g.bin_w, g.gw, p.W, 1.f / g.count, wa->xt, wa->xadv, &wa->x0, false, lane);
      jump |= build_axis(g.start_h, g.bin_h, g.gh, p.H, 1.f, wa->yt, wa->yadv, &wa->y0, true, lane);
      if (__any_sync(0xffffffffu, jump)) mode = 2;
    }
    if (mode == 2 && lane == 0) {
      wa->gw = g.gw;
      wa->gh = g.gh;
      wa->inv_count = 1.f / g.count;
      wa->start_w = g.start_w;
      wa->start_h = g.start_h;
      wa->bin_w = g.bin_w;
      wa->bin_h = g.bin_h;
    }
    __syncwarp();
    float* img = p.scratch + (size_t)(mode ? n : 0) * p.H * p.W * p.C + cb * CB + 2 * lane;

    if (mode == 1) {
      const int gw = g.gw, gh = g.gh;
      const long long cstep = (long long)p.C * 4;
      const long long rstep = cstep * p.W;
      char* rowp = reinterpret_cast<char*>(img) + ((long long)wa->y0 * p.W + wa->x0) * cstep;  // first footprint cell
      // advance bits as ballots: warp-uniform by construction, so the window branches below are uniform branches
      uint32_t ym[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) ym[i] = __ballot_sync(0xffffffffu, (wa->yadv[i] >> lane) & 1u);
      const float4* xt = wa->xt;
      const ulonglong2* yt = reinterpret_cast<const ulonglong2*>(wa->yt);
      f2 vlo[P], vhi[P], g2[P];
#pragma unroll
      for (int i = 0; i < P; ++i) vlo[i] = vhi[i] = g2[i] = 0ull;
      const int nsy = P * gh;
      int iy = 0, ph = 0, b = 0;
      for (int sy = 0; sy <= nsy; ++sy) {  // sample nsy: the zero sentinel that flushes the last row
        if (iy == 0 && ph < P) {
          if (BF) {  // 14 bf16 of channel 2l and of channel 2l+1: 7 + 7 words, widened by shifts
            const int s = ph >> 1, q = ph & 1;
            if (q == 0) bf_wait(s);
            const uint32_t* te = reinterpret_cast<const uint32_t*>(ring + bf_slot_even(s) * BF_BOX_BYTES +
                                                                   lane * BF_PITCH + (2 * (s & 1) + q) * (2 * P));
            const uint32_t* to = reinterpret_cast<const uint32_t*>(ring + bf_slot_odd(s) * BF_BOX_BYTES +
                                                                   lane * BF_PITCH + (2 * ((s + 1) & 1) + q) * (2 * P));
#pragma unroll
            for (int i = 0; i < P / 2; ++i) {
              const uint32_t e = te[i], o = to[i];
              g2[2 * i] = ((unsigned long long)(o << 16) << 32) | (unsigned long long)(e << 16);
              g2[2 * i + 1] = ((unsigned long long)(o & 0xffff0000u) << 32) | (unsigned long long)(e & 0xffff0000u);
            }
            if (q == 1) {
              __syncwarp();  // every lane is done with this step's rows
              if (lane == 0) bf_release(cur, nxt, s);
            }
          } else {
            const int half = ph % ROWS;
            if (half == 0) {
              b = cc % NST;
              mbar_wait(&wa->bar[b], (cc / NST) & 1u);
              ++cc;
            }
            const float* ta = wa->stage[b] + lane * CHUNK_F + half * P;
#pragma unroll
            for (int j = 0; j < P; ++j) g2[j] = pack2(ta[j], ta[j + 32 * CHUNK_F]);
            if (half == ROWS - 1 || ph == P - 1) {
              __syncwarp();  // every lane is done with the tile: refill the buffer NST tiles ahead in the stream
              if (lane == 0) refill(cur, nxt, ph / ROWS, b);
            }
          }
        }
        const ulonglong2 t = yt[sy];
        const uint32_t w = sy < 32 ? ym[0] : (sy < 64 ? ym[1] : ym[2]);
        const bool adv = (w >> (sy & 31)) & 1u;
        if (++iy == gh) {
          iy = 0;
          ++ph;
        }
#pragma unroll
        for (int pw = 0; pw < P; ++pw) fma2(vlo[pw], t.x, g2[pw]);
        if (adv) {  // last sample whose lower tap is this feature row: the row is complete, sweep it
          if (gw == 1) sweep<1>(xt, vlo, rowp, cstep, gw);
          else if (gw == 2) sweep<2>(xt, vlo, rowp, cstep, gw);
          else if (gw == 3 && p.sweep3) sweep<3>(xt, vlo, rowp, cstep, gw);
          else sweep<0>(xt, vlo, rowp, cstep, gw);
          rowp += rstep;
        }
        // upper taps: vhi += ly * g, or (after a sweep) vlo = vhi + ly * g, vhi = 0
        ystep_upper7(vlo, vhi, g2, t.y, adv);
        ystep_upper7(vlo + 7, vhi + 7, g2 + 7, t.y, adv);
      }
    } else {
      // degenerate / foreign RoIs (mode 0) only drain their tiles; mode 2 evaluates every tap directly
#pragma unroll 1
      for (int k = 0; k < NCH; ++k) {
        if (BF) {
          bf_wait(k);
          if (mode == 2) {
            for (int q = 0; q < 2; ++q) {
              const __nv_bfloat16* te = reinterpret_cast<const __nv_bfloat16*>(
                  ring + bf_slot_even(k) * BF_BOX_BYTES + lane * BF_PITCH + (2 * (k & 1) + q) * (2 * P));
              const __nv_bfloat16* to = reinterpret_cast<const __nv_bfloat16*>(
                  ring + bf_slot_odd(k) * BF_BOX_BYTES + lane * BF_PITCH + (2 * ((k + 1) & 1) + q) * (2 * P));
              direct_row<__nv_bfloat16>(p, wa, te, to, img, 2 * k + q);
            }
          }
          __syncwarp();
          if (lane == 0) bf_release(cur, nxt, k);
          continue;
        }
        const int b = cc % NST;
        mbar_wait(&wa->bar[b], (cc / NST) & 1u);
        ++cc;
        if (mode == 2) {
          for (int half = 0; half < ROWS && ROWS * k + half < P; ++half) {
            const float* ta = wa->stage[b] + lane * CHUNK_F + half * P;
            direct_row<float>(p, wa, ta, ta + 32 * CHUNK_F, img, ROWS * k + half);
          }
        }
        __syncwarp();
        if (lane == 0) refill(cur, nxt, k, b);
      }
    }
    __syncwarp();  // every lane has finished reading this item's tables before the next item rebuilds them
    cur = nxt;
    nxt = cur < n_items ? fetch() : n_items;
  }
}

// scratch [N][HW][C] -> grad_feat [N][C][HW].  fp32: channels permuted within 64-blocks (lane l wrote l, l + 32);
// bf16: natural order (lane l wrote 2l, 2l + 1), rounded to bf16 here (once).
template <bool BF>
__global__ void __launch_bounds__(256)
unpermute_kernel(const float* __restrict__ scratch, void* __restrict__ out, int C, int HW) {
  __shared__ float tile[CB][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int hw0 = blockIdx.x * 32, cb = blockIdx.y, n = blockIdx.z;
  const float* src = scratch + (size_t)n * HW * C + cb * CB;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int hw = hw0 + ty + 8 * j;
    if (hw < HW) {
      const float2 v = *reinterpret_cast<const float2*>(src + (size_t)hw * C + 2 * tx);
      tile[BF ? 2 * tx : tx][ty + 8 * j] = v.x;
      tile[BF ? 2 * tx + 1 : tx + 32][ty + 8 * j] = v.y;
    }
  }
  __syncthreads();
  const size_t base = ((size_t)n * C + cb * CB) * HW;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = ty + 8 * j;
    if (hw0 + tx < HW) {
      if (BF) reinterpret_cast<__nv_bfloat16*>(out)[base + (size_t)c * HW + hw0 + tx] = __float2bfloat16_rn(tile[c][tx]);
      else reinterpret_cast<float*>(out)[base + (size_t)c * HW + hw0 + tx] = tile[c][tx];
    }
  }
}

__global__ void zero4_kernel(float4* __restrict__ p, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
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

}  // namespace cl

// epilogue kernel shared with roi_align_bwd_cl2.cu
void launch_bwd_cl_unpermute(const float* scratch, void* gfeat, int N, int C, int HW, bool bf16, cudaStream_t st) {
  dim3 tgrid((HW + 31) / 32, C / cl::CB, N);
  if (bf16) cl::unpermute_kernel<true><<<tgrid, 256, 0, st>>>(scratch, gfeat, C, HW);
  else cl::unpermute_kernel<false><<<tgrid, 256, 0, st>>>(scratch, gfeat, C, HW);
}

bool bwd_cl_fits(int C, int H, int W, int R, int dtype, const void* gout) {
  return (dtype == UNIT_F32 || dtype == UNIT_BF16) && (C % cl::CB) == 0 && H >= 2 && W >= 2 &&
         ((uintptr_t)gout & 15) == 0 &&
         (long long)R * C < (1ll << 31) && cl::encode_fn() != nullptr;
}

// workspace: [256 B counter][N*H*W*C fp32 channel-last image]
int launch_bwd_cl(const void* gout, const float* rois, void* gfeat, void* ws, int N, int C, int H, int W, int R,
                  float scale, int sr, int aligned, int dtype, cudaStream_t st) {
  using namespace cl;
  const bool bf = dtype == UNIT_BF16;
  CUtensorMap map;
  if (int rc = ensure_driver_context(gout)) return rc;
  {
    // fp32: rows = channels (196 floats, 784 B).  bf16: rows = channel PAIRS (2 x 196 bf16, 784 B).
    cuuint64_t dims[2] = {(cuuint64_t)(bf ? 2 * P * P : P * P), (cuuint64_t)R * C / (bf ? 2 : 1)};
    cuuint64_t strides[1] = {(cuuint64_t)(P * P) * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)(bf ? BF_BOX_ELEMS : CHUNK_F), (cuuint32_t)(bf ? CB / 2 : CB)};
    cuuint32_t estr[2] = {1, 1};
    const int promo = switches().bwd_promo;
    const CUtensorMapL2promotion pr = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                     : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                     : promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                  : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    CUresult rc = encode_fn()(&map, bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                              const_cast<void*>(gout), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UNIT_REQUIRE(rc == CUDA_SUCCESS, "roi_align_bwd: cuTensorMapEncodeTiled failed (%d)", (int)rc);
  }
  Params p;
  p.rois = rois;
  p.counter = (int*)ws;
  p.scratch = (float*)((char*)ws + 256);
  p.N = N;
  p.C = C;
  p.H = H;
  p.W = W;
  p.R = R;
  p.scale = scale;
  p.sampling_ratio = sr;
  p.aligned = aligned;
  p.evict_first = switches().bwd_evict_first;
  p.sweep3 = switches().bwd_sweep3;
  const long long total = (long long)N * C * H * W;  // multiple of 64
  const long long n4 = total / 4 + 16;
  const int zgrid = (int)std::min<long long>((n4 + 255) / 256, (long long)sm_count() * 8);
  zero4_kernel<<<zgrid, 256, 0, st>>>((float4*)ws, n4);
  UNIT_CHECK_LAUNCH("zero4_kernel");
  const size_t smem = (size_t)NW * sizeof(WarpArea);
  const long long items = (long long)R * (C / CB);
  long long grid = (items + NW - 1) / NW;
  if (grid > sm_count()) grid = sm_count();
  if (grid < 1) grid = 1;
  dim3 tgrid((H * W + 31) / 32, C / CB, N);
  if (bf) {
    UNIT_CUDA(cudaFuncSetAttribute(roi_align_bwd_cl<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_align_bwd_cl<true><<<(int)grid, NT, smem, st>>>(map, p);
    UNIT_CHECK_LAUNCH("roi_align_bwd_cl");
    unpermute_kernel<true><<<tgrid, 256, 0, st>>>(p.scratch, gfeat, C, H * W);
  } else {
    UNIT_CUDA(cudaFuncSetAttribute(roi_align_bwd_cl<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_align_bwd_cl<false><<<(int)grid, NT, smem, st>>>(map, p);
    UNIT_CHECK_LAUNCH("roi_align_bwd_cl");
    unpermute_kernel<false><<<tgrid, 256, 0, st>>>(p.scratch, gfeat, C, H * W);
  }
  UNIT_CHECK_LAUNCH("unpermute_kernel");
  return UNIT_OK;
}

}  // namespace roi
}  // namespace unit
