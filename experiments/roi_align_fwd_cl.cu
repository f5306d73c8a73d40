// ROIAlign forward, channel-lane kernel (fp32, C % 64 == 0, 14 x 14 bins) -- the mirror
// image of the channel-lane backward (roi_align_bwd_cl.cu).
//
// The average over a bin's sampling grid is separable with the SAME weights for every channel, so a lane owns CHANNELS
// (two of them) and the whole warp executes one warp-uniform walk over the RoI's x- and y-samples: no lane idles
// because an RoI is narrow, no branch diverges, table entries are broadcast shared-memory loads.
//
//   pre-pass    the feature map is copied once per call to channel-last [N][H*W][C] (34 MB at the bench workload: it
//               stays in L2), channels of a 64-block interleaved so that a lane's two channels are adjacent.
//   work item   (RoI, block of 64 channels), pulled from a global counter by the 12 warps of 148 persistent CTAs.
//   input       the footprint rows arrive as [8 pixels x 64 channels] tiles (2 KB) by TMA tensor loads into a 4-tile
//               ring per warp (mbarrier per tile, two tiles of look-ahead); a pixel is one LDS.64 per lane.
//   horizontal  a footprint row is swept once over the x-samples with a two-pixel register window (cur, next):
//               H[pw] += hx * cur + lx * next; when the lower tap column advances, cur = next and the next pixel
//               is read.  (One PTX block per sample, so the window registers keep their names on both paths.)
//   vertical    consecutive row results Hp (row y), Hn (row y+1) are combined with the y-samples whose lower tap is
//               row y: O[pw] += hy * Hp[pw] + ly * Hn[pw]; a finished bin row goes to the staging tile.
//   output      out[r][c0 .. c0+63][2k .. 2k+1][0 .. 13] -- a [64 x 28] tile per two bin rows -- leaves by TMA tensor
//               store.
// Feature reads come from L2 (the footprint of every RoI, ~0.8 GB of L2 traffic for the bench workload) instead of a
// shared-memory slab, which frees the SM of the 134 KB slab and of its load / barrier phases.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "roi_common.cuh"

namespace unit {
namespace roi {
namespace fcl {

constexpr int P = 14;
constexpr int CB = 64;                   // channels per work item (two per lane: l and l + 32)
constexpr int CHUNK_ROWS = 2;            // bin rows per output tile
constexpr int CHUNK_F = CHUNK_ROWS * P;  // 28 floats per channel and tile
constexpr int NW = 12;
constexpr int NT = NW * 32;
constexpr int MAXG = 6;                  // sampling grid with tables: RoI side <= 84 feature cells
constexpr int MAXS = P * MAXG;
constexpr int BOXP = 16;                 // pixels per input tile
constexpr int BOXP_LOG2 = 4;
constexpr int RING = 2;                  // input tiles per warp
constexpr int TILE_F = BOXP * CB;        // 1024 floats = 4 KB
constexpr uint32_t TILE_BYTES = TILE_F * sizeof(float);
constexpr uint32_t RING_MASK = RING * TILE_BYTES - 1;

struct __align__(128) WarpArea {
  float ring[RING][TILE_F];
  float stage[CB * CHUNK_F];
  float4 xt[MAXS];  // (hx / count | ADV in the mantissa LSB, hx / count, lx / count, lx / count)
  float4 yt[MAXS];  // (hy | ADV, hy, ly, ly)
  uint32_t xadv[4], yadv[4];
  uint64_t bar[RING];
  int x0, y0;
  int total, nt, chan0, width;           // tile stream of the current item (slow path bookkeeping)
  int next_t, next_pix, row_pix;         // tile-in-row / first pixel of the next tile to issue, first pixel of its row
  uint32_t tbase;
};

struct Params {
  const float* rois;
  const float* ft;  // channel-last copy of the features (direct path only; the tiles come through the tensor map)
  int* counter;
  int N, C, H, W, R;
  float scale;
  int sampling_ratio, aligned;
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
__device__ __forceinline__ void tma_tile_load(void* sdst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(sdst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_tile_store(const CUtensorMap* map, const void* ssrc, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(ssrc)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {  // a * b + c on both halves
  unsigned long long ua, ub, uc, ud;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ua) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ub) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(uc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ua), "l"(ub), "l"(uc));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(ud));
  return r;
}

// One x-sample of the horizontal sweep as ONE block of PTX:  h += eh * cur + el * next; last sample of its column:
// cur = next, next = the following pixel of the ring (off = its byte offset, k = its index in the row).
__device__ __forceinline__ void sweep_step(float2& h, float2& cur, float2& nxt, uint32_t& off, int& k, float4 e,
                                           uint32_t lane_base) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      ".reg .b64 a, c, n, eh, el;\n"
      ".reg .b32 fl, ad;\n"
      "mov.b64 a, {%0, %1};\n"
      "mov.b64 c, {%2, %3};\n"
      "mov.b64 n, {%4, %5};\n"
      "mov.b64 eh, {%8, %9};\n"
      "mov.b64 el, {%10, %11};\n"
      "fma.rn.f32x2 a, eh, c, a;\n"
      "fma.rn.f32x2 a, el, n, a;\n"
      "mov.b64 {%0, %1}, a;\n"
      "mov.b32 fl, %8;\n"
      "and.b32 fl, fl, 1;\n"
      "setp.eq.u32 q, fl, 0;\n"
      "@q bra.uni SWEEP_NEXT;\n"
      "mov.f32 %2, %4;\n"
      "mov.f32 %3, %5;\n"
      "add.u32 %6, %6, 256;\n"
      "and.b32 %6, %6, %13;\n"
      "add.u32 ad, %12, %6;\n"
      "ld.shared.v2.f32 {%4, %5}, [ad];\n"
      "add.s32 %7, %7, 1;\n"
      "SWEEP_NEXT:\n"
      "}\n"
      : "+f"(h.x), "+f"(h.y), "+f"(cur.x), "+f"(cur.y), "+f"(nxt.x), "+f"(nxt.y), "+r"(off), "+r"(k)
      : "f"(e.x), "f"(e.y), "f"(e.z), "f"(e.w), "r"(lane_base), "r"(RING_MASK));
}

// One axis of the RoI (whole warp): table entry s = (h * ws | ADV, h * ws, l * ws, l * ws); ADV (mantissa LSB, also bit
// s of advw[]) marks the last sample whose lower tap is this cell.  Samples clamped to the last cell (value F[size-1])
// are re-expressed as lo = size-2 with weights (0, h), so lo + 1 is always inside.  Returns true when consecutive
// lower taps are not 0 or 1 apart (sampling_ratio > 0 on large RoIs, fp32 rounding): such RoIs take the direct path.
__device__ __forceinline__ bool build_axis(float start, float bin, int g, int size, float ws, float4* tab,
                                           uint32_t* advw, int* first_lo, int lane) {
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
    uint32_t m = 0u;
    if (32 * k < ns) {  // warp-uniform
      const int s = 32 * k + lane;
      int nlo = __shfl_down_sync(0xffffffffu, lo[k], 1);
      const int nx = k + 1 < ROUNDS ? __shfl_sync(0xffffffffu, lo[k + 1 < ROUNDS ? k + 1 : k], 0) : 0x3fffffff;
      if (lane == 31) nlo = nx;
      const bool adv = s < ns && (s == ns - 1 || nlo != lo[k]);
      jump |= (s < ns - 1) && (nlo - lo[k] > 1 || nlo < lo[k]);
      m = __ballot_sync(0xffffffffu, adv);
      if (s < ns)
        tab[s] = make_float4(__uint_as_float((__float_as_uint(hw[k]) & ~1u) | (adv ? 1u : 0u)), hw[k], lw[k], lw[k]);
      if (s == 0) *first_lo = lo[k];
    }
    if (lane == 0) advw[k] = m;
  }
  return jump;
}

// rare path: every tap read straight from the channel-last copy in global memory (L2)
__device__ __noinline__ void direct_rows(const Params& p, const float* __restrict__ ftl, const Geom& g, int ph0,
                                         float* __restrict__ stage, int lane) {
  const float inv = 1.f / g.count;
  for (int half = 0; half < CHUNK_ROWS; ++half) {
    const int ph = ph0 + half;
    for (int pw = 0; pw < P; ++pw) {
      float2 acc = make_float2(0.f, 0.f);
      for (int iy = 0; iy < g.gh; ++iy) {
        int ylo, yhi;
        float ly, hy;
        const bool vy = axis_tap(sample_coord(g.start_h, g.bin_h, ph, iy, g.gh), p.H, ylo, yhi, ly, hy);
        for (int ix = 0; ix < g.gw; ++ix) {
          int xlo, xhi;
          float lx, hx;
          const bool vx = axis_tap(sample_coord(g.start_w, g.bin_w, pw, ix, g.gw), p.W, xlo, xhi, lx, hx);
          if (vy && vx) {
            const float2 v1 = *reinterpret_cast<const float2*>(ftl + ((size_t)ylo * p.W + xlo) * p.C);
            const float2 v2 = *reinterpret_cast<const float2*>(ftl + ((size_t)ylo * p.W + xhi) * p.C);
            const float2 v3 = *reinterpret_cast<const float2*>(ftl + ((size_t)yhi * p.W + xlo) * p.C);
            const float2 v4 = *reinterpret_cast<const float2*>(ftl + ((size_t)yhi * p.W + xhi) * p.C);
            acc.x += hy * (hx * v1.x + lx * v2.x) + ly * (hx * v3.x + lx * v4.x);
            acc.y += hy * (hx * v1.y + lx * v2.y) + ly * (hx * v3.y + lx * v4.y);
          }
        }
      }
      stage[lane * CHUNK_F + half * P + pw] = acc.x * inv;
      stage[(lane + 32) * CHUNK_F + half * P + pw] = acc.y * inv;
    }
  }
}

// Slow path of the input ring (about once per 16 pixels): issue the tiles < want of the current item (lane 0), then
// wait until tile `need` is resident.  st = issued << 16 | ready (tiles issued / waited for so far); returns the new st.
// Callers guarantee that the slots the new tiles go to hold tiles every lane has finished with.
__device__ __noinline__ int tiles_advance(WarpArea* wa, const CUtensorMap* fmap, int want, int need, int st, int lane) {
  int issued = st >> 16, ready = st & 0xffff;
  __syncwarp();
  if (want > issued) {
    if (lane == 0) {
      int t = wa->next_t, pix = wa->next_pix, rpix = wa->row_pix;
      const int nt = wa->nt;
      for (int jj = issued; jj < want; ++jj) {
        const uint32_t slot = (wa->tbase + jj) % RING;
        mbar_expect_tx(&wa->bar[slot], TILE_BYTES);
        tma_tile_load(wa->ring[slot], fmap, &wa->bar[slot], wa->chan0, pix);
        pix += BOXP;
        if (++t == nt) {
          t = 0;
          rpix += wa->width;
          pix = rpix;
        }
      }
      wa->next_t = t;
      wa->next_pix = pix;
      wa->row_pix = rpix;
    }
    issued = want;
  }
  while (ready <= need) {
    const uint32_t q = wa->tbase + ready;
    mbar_wait(&wa->bar[q % RING], (q / RING) & 1u);
    ++ready;
  }
  return (issued << 16) | ready;
}

// Ring bookkeeping in front of a bin (out of line: the sweep stays small).  Two tiles are resident: the one holding
// `cur` and the next.  (a) if the bin can reach a pixel that has not been waited for, wait for its tile; (b) once `cur`
// has entered the newest issued tile the older one is free and the following tile is requested.  Returns the new st and
// the next value of k + G that needs another look.
__device__ __noinline__ int2 tiles_check(WarpArea* wa, const CUtensorMap* fmap, int k, int G, int jr, int nt, int total,
                                         int st, int lane) {
  const int tc = jr + ((k - 1) >> BOXP_LOG2);  // tile holding cur
  const int want = max(min(tc + RING, total), st >> 16);
  const int need = jr + min((k + G) >> BOXP_LOG2, nt - 1);
  st = tiles_advance(wa, fmap, want, need, st, lane);
  const int issued = st >> 16, ready = st & 0xffff;
  int kcheck = 0x3fffffff;
  if (ready < jr + nt) kcheck = (ready - jr) << BOXP_LOG2;                                  // (a)
  if (issued < total) kcheck = min(kcheck, ((issued - 1 - jr) << BOXP_LOG2) + 1 + G);       // (b), in k + G units
  return make_int2(st, kcheck);
}

// horizontal sweep of one footprint row (two channels per lane).  GW > 0: unrolled; GW == 0: any sampling grid.
template <int GW>
__device__ __forceinline__ void sweep_row(WarpArea* wa, const CUtensorMap* fmap, int gw, int jr, int nt, int total,
                                          int& st, float2 (&h)[P], float2& cur, float2& nxt, uint32_t& off, int& k,
                                          uint32_t lane_base, int lane) {
  const float4* xt = wa->xt;
  const int G = GW > 0 ? GW : gw;
  int kcheck = 0;  // look at the ring before the first bin
#pragma unroll
  for (int pw = 0; pw < P; ++pw) {
    h[pw] = make_float2(0.f, 0.f);
    if (k + G >= kcheck) {
      const int2 r = tiles_check(wa, fmap, k, G, jr, nt, total, st, lane);
      st = r.x;
      kcheck = r.y;
    }
    if (GW > 0) {
#pragma unroll
      for (int ix = 0; ix < GW; ++ix) sweep_step(h[pw], cur, nxt, off, k, xt[pw * GW + ix], lane_base);
    } else {
#pragma unroll 1
      for (int ix = 0; ix < gw; ++ix) sweep_step(h[pw], cur, nxt, off, k, xt[pw * gw + ix], lane_base);
    }
  }
}

__global__ void __launch_bounds__(NT, 1)
roi_align_fwd_cl(const __grid_constant__ CUtensorMap fmap, const __grid_constant__ CUtensorMap omap, const Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WarpArea* wa = reinterpret_cast<WarpArea*>(smem_raw) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int nblk = p.C / CB;
  const long long n_items = (long long)p.R * nblk;
  if (lane == 0) {
    for (int b = 0; b < RING; ++b) mbar_init(&wa->bar[b], 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const uint32_t lane_base = smem_u32(&wa->ring[0][0]) + 8u * lane;
  uint32_t tbase = 0;  // ring tiles consumed by this warp in earlier items (slot = (tbase + j) % RING)

  while (true) {
    long long it = 0;
    if (lane == 0) it = atomicAdd(p.counter, 1);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= n_items) break;
    const int r = (int)(it / nblk);
    const int cb = (int)(it - (long long)r * nblk);
    const float* roi = p.rois + (long long)r * 5;
    const int n = (int)roi[0];
    const Geom g = roi_geom(roi, p.scale, P, P, p.sampling_ratio, p.aligned);
    int mode = 1;
    if (g.gw <= 0 || g.gh <= 0 || n < 0 || n >= p.N) mode = 0;
    else if (g.gw > MAXG || g.gh > MAXG) mode = 2;
    if (mode == 1) {
      bool jump = build_axis(g.start_w, g.bin_w, g.gw, p.W, 1.f / g.count, wa->xt, wa->xadv, &wa->x0, lane);
      jump |= build_axis(g.start_h, g.bin_h, g.gh, p.H, 1.f, wa->yt, wa->yadv, &wa->y0, lane);
      if (__any_sync(0xffffffffu, jump)) mode = 2;
    }
    __syncwarp();
    const int orow = r * p.C + cb * CB;  // first row of this item in the [R*C][196] view of the output

    if (mode == 1) {
      const int gw = g.gw, gh = g.gh;
      const int x0 = wa->x0, y0 = wa->y0;
      const int npx = __popc(wa->xadv[0]) + __popc(wa->xadv[1]) + __popc(wa->xadv[2]) + 1;
      const int nrows = __popc(wa->yadv[0]) + __popc(wa->yadv[1]) + __popc(wa->yadv[2]) + 1;
      const int nt = (npx + BOXP - 1) / BOXP;
      const int total = nrows * nt;
      if (lane == 0) {
        wa->total = total;
        wa->nt = nt;
        wa->next_t = 0;
        wa->next_pix = wa->row_pix = (n * p.H + y0) * p.W + x0;
        wa->chan0 = cb * CB;
        wa->width = p.W;
        wa->tbase = tbase;
      }
      int st = 0;  // tiles of this item issued << 16 | waited for
      uint32_t ym[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) ym[i] = __ballot_sync(0xffffffffu, (wa->yadv[i] >> lane) & 1u);
      const float4* yt = wa->yt;
      float2 hp[P], hn[P], o[P];
#pragma unroll
      for (int i = 0; i < P; ++i) hp[i] = o[i] = make_float2(0.f, 0.f);
      int sy = 0, iy = 0, ph = 0;
      for (int ri = 0; ri < nrows; ++ri) {
        const int jr = ri * nt;  // first tile of this row; every earlier tile has been waited for and is free
        st = tiles_advance(wa, &fmap, max(min(jr + RING, total), st >> 16), jr, st, lane);
        uint32_t off = ((tbase + jr) % RING) * TILE_BYTES;
        float2 cur, nxt;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(cur.x), "=f"(cur.y) : "r"(lane_base + off));
        off = (off + 256u) & RING_MASK;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(nxt.x), "=f"(nxt.y) : "r"(lane_base + off));
        int k = 1;  // index (in the row) of the pixel held by nxt
        if (gw == 1) sweep_row<1>(wa, &fmap, gw, jr, nt, total, st, hn, cur, nxt, off, k, lane_base, lane);
        else if (gw == 2) sweep_row<2>(wa, &fmap, gw, jr, nt, total, st, hn, cur, nxt, off, k, lane_base, lane);
        else sweep_row<0>(wa, &fmap, gw, jr, nt, total, st, hn, cur, nxt, off, k, lane_base, lane);
        if (ri > 0) {
          bool adv;
          do {  // the y-samples whose lower tap is row ri - 1: O += hy * H(ri - 1) + ly * H(ri)
            const float4 t = yt[sy];
            const uint32_t w = sy < 32 ? ym[0] : (sy < 64 ? ym[1] : ym[2]);
            adv = (w >> (sy & 31)) & 1u;
            ++sy;
            const float2 hy = make_float2(t.x, t.y), ly = make_float2(t.z, t.w);
#pragma unroll
            for (int pw = 0; pw < P; ++pw) {
              o[pw] = ffma2(hy, hp[pw], o[pw]);
              o[pw] = ffma2(ly, hn[pw], o[pw]);
            }
            if (++iy == gh) {  // bin row ph is complete
              iy = 0;
              const int half = ph & 1;
              if (half == 0) {
                if (lane == 0) bulk_wait_read_all();  // the previous tile store has drained the staging tile
                __syncwarp();
              }
              float* sa = wa->stage + lane * CHUNK_F + half * P;
#pragma unroll
              for (int pw = 0; pw < P; ++pw) {
                sa[pw] = o[pw].x;
                sa[pw + 32 * CHUNK_F] = o[pw].y;
                o[pw] = make_float2(0.f, 0.f);
              }
              if (half == 1) {
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_tile_store(&omap, wa->stage, CHUNK_F * (ph >> 1), orow);
                  bulk_commit();
                }
              }
              ++ph;
            }
          } while (!adv);
        }
#pragma unroll
        for (int pw = 0; pw < P; ++pw) hp[pw] = hn[pw];
      }
      tbase += (uint32_t)total;
    } else {
      // degenerate RoIs (mode 0) produce zeros; mode 2 evaluates every tap directly
      const float* ftl = p.ft + ((size_t)(mode == 2 ? n : 0) * p.H * p.W) * p.C + cb * CB + 2 * lane;
      for (int k = 0; k < P / CHUNK_ROWS; ++k) {
        if (lane == 0) bulk_wait_read_all();
        __syncwarp();
        if (mode == 2) {
          direct_rows(p, ftl, g, CHUNK_ROWS * k, wa->stage, lane);
        } else {
          for (int i = lane; i < CB * CHUNK_F; i += 32) wa->stage[i] = 0.f;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_tile_store(&omap, wa->stage, CHUNK_F * k, orow);
          bulk_commit();
        }
      }
    }
  }
  if (lane == 0) bulk_wait_all();
}

// features [N][C][HW] -> channel-last [N][HW][C], channel (64b + l) at 64b + 2l and (64b + 32 + l) at 64b + 2l + 1
__global__ void __launch_bounds__(256)
permute_kernel(const float* __restrict__ feat, float* __restrict__ ft, int C, int HW) {
  __shared__ float tile[CB][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int hw0 = blockIdx.x * 32, cb = blockIdx.y, n = blockIdx.z;
  const float* src = feat + ((size_t)n * C + cb * CB) * HW;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = ty + 8 * j;
    if (hw0 + tx < HW) tile[c][tx] = src[(size_t)c * HW + hw0 + tx];
  }
  __syncthreads();
  float* dst = ft + (size_t)n * HW * C + cb * CB;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int hw = hw0 + ty + 8 * j;
    if (hw < HW)
      *reinterpret_cast<float2*>(dst + (size_t)hw * C + 2 * tx) = make_float2(tile[tx][ty + 8 * j], tile[tx + 32][ty + 8 * j]);
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

}  // namespace fcl

bool fwd_cl_fits(int N, int C, int H, int W, int R, int dtype, const void* out) {
  return dtype == UNIT_F32 && (C % fcl::CB) == 0 && H >= 2 && W >= 2 && ((uintptr_t)out & 15) == 0 &&
         (long long)R * C < (1ll << 31) && (long long)N * H * W < (1ll << 30) && fcl::encode_fn() != nullptr;
}

// channel-last copy of the features + the work-item counter
size_t fwd_cl_workspace_bytes(int N, int C, int H, int W) { return (size_t)N * C * H * W * 4 + 256; }

int launch_fwd_cl(const void* feat, const float* rois, void* out, void* ws, int N, int C, int H, int W, int R,
                  float scale, int sr, int aligned, cudaStream_t st) {
  using namespace fcl;
  float* ft = (float*)((char*)ws + 256);
  UNIT_CUDA(cudaMemsetAsync(ws, 0, 256, st));
  dim3 tgrid((H * W + 31) / 32, C / CB, N);
  permute_kernel<<<tgrid, 256, 0, st>>>((const float*)feat, ft, C, H * W);
  UNIT_CHECK_LAUNCH("permute_kernel");
  CUtensorMap fmap, omap;
  {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)N * H * W};
    cuuint64_t strides[1] = {(cuuint64_t)C * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)CB, (cuuint32_t)BOXP};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode_fn()(&fmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ft, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UNIT_REQUIRE(rc == CUDA_SUCCESS, "roi_align_fwd: cuTensorMapEncodeTiled (features) failed (%d)", (int)rc);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)(P * P), (cuuint64_t)R * C};
    cuuint64_t strides[1] = {(cuuint64_t)(P * P) * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)CHUNK_F, (cuuint32_t)CB};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode_fn()(&omap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UNIT_REQUIRE(rc == CUDA_SUCCESS, "roi_align_fwd: cuTensorMapEncodeTiled (output) failed (%d)", (int)rc);
  }
  Params p;
  p.rois = rois;
  p.ft = ft;
  p.counter = (int*)ws;
  p.N = N;
  p.C = C;
  p.H = H;
  p.W = W;
  p.R = R;
  p.scale = scale;
  p.sampling_ratio = sr;
  p.aligned = aligned;
  const size_t smem = (size_t)NW * sizeof(WarpArea);
  UNIT_CUDA(cudaFuncSetAttribute(roi_align_fwd_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long items = (long long)R * (C / CB);
  long long grid = (items + NW - 1) / NW;
  if (grid > sm_count()) grid = sm_count();
  if (grid < 1) grid = 1;
  roi_align_fwd_cl<<<(int)grid, NT, smem, st>>>(fmap, omap, p);
  UNIT_CHECK_LAUNCH("roi_align_fwd_cl");
  return UNIT_OK;
}

}  // namespace roi
}  // namespace unit
