// ROIAlign backward, slab-resident kernel.
//
// Same decomposition as roi_align_fwd.cu, transposed.  A persistent CTA keeps the GRADIENT tile of one
// (image, 8-channel slab) -- channel-pair interleaved float2, 134 KB for the 50x84 map -- in shared memory.  Each of
// its warps pulls RoIs of that slab, streams the RoI's [8][14][14] grad_out block in with one TMA bulk load
// (cp.async.bulk.shared::cta.global + mbarrier, overlapped with the table build; the following RoI's block is
// prefetched into L2), and scatters it column-wise: lane (c, cp) OWNS feature column x0 + c of channel pair cp for
// this RoI.  For each of the 14 bin rows it gathers the horizontally spread value u = sum_s w_s * g[row][bin(s)] over
// the samples touching its column (two contiguous sample ranges of the x-table), then spreads u vertically through a
// two-row register window, so that every tile cell (y, x) of the RoI's footprint receives exactly ONE 64-bit
// shared-memory CAS-add (both channels at once).  Lanes of a warp never collide (distinct columns / planes); only
// different warps (different RoIs) can, and rarely.  Per RoI that is footprint-many atomics on shared memory instead
// of torchvision's 4*gh*gw global atomics per output element (~1e9 for 2x512 RoIs).
// The tile reaches HBM once per segment (plain stores when the CTA owns the whole (image, slab), red.global else).
#include "roi_slab.cuh"

namespace unit {
namespace roi {
namespace v2 {

constexpr int BMAXG = 4;             // sampling grid handled with tables in the backward (RoI side <= 56 px)
constexpr int BMAXS = P * BMAXG;
constexpr int BNWARPS = 12;
constexpr int BNTHREADS = BNWARPS * 32;

struct __align__(16) XSamp {
  float h, l;  // weights of column lo and lo+1, already scaled by 1/count
  int pw;      // bin column of the sample
  int pad;
};

template <typename T>
struct __align__(16) BwdArea {
  int gw, gh, mode, x0, y0, ncols, nsx, nsy;
  float inv_count, start_w, start_h, bin_w, bin_h;
  int pad[3];
  XSamp xs[BMAXS];
  int colstart[BMAXS + 4];   // colstart[c] = first sample whose lower tap is >= x0 + c   (ncols + 1 entries)
  float2 ys[BMAXS + 2];      // (hy | LASTROW sign, ly)
  __align__(16) T stage[CS * P * P];
  uint64_t bar;
  uint64_t pad2;
};

__device__ __forceinline__ void smem_add2(float2* addr, float2 v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    const float x = __uint_as_float((unsigned)(assumed & 0xffffffffull)) + v.x;
    const float y = __uint_as_float((unsigned)(assumed >> 32)) + v.y;
    const unsigned long long nv = ((unsigned long long)__float_as_uint(y) << 32) | (unsigned long long)__float_as_uint(x);
    old = atomicCAS(a, assumed, nv);
  } while (old != assumed);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
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
      "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(sdst)),
               "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

// Tables of one RoI for the column-owner scatter (whole warp).
template <typename T>
__device__ __forceinline__ void build_bwd_tables(const float* __restrict__ roi, const Params& p, BwdArea<T>* ba,
                                                 int lane) {
  const Geom g = roi_geom(roi, p.scale, P, P, p.sampling_ratio, p.aligned);
  int mode = 1;
  if (g.gw <= 0 || g.gh <= 0) mode = 0;
  else if (g.gw > BMAXG || g.gh > BMAXG || p.W < 2 || p.H < 2) mode = 2;
  bool jump = false;
  if (mode == 1) {
    const float inv = 1.f / g.count;
    // ---- x samples: weights, bin index, and the column -> first-sample index (colstart)
    const int ns = P * g.gw;
    const float inv_gw = 1.f / (float)g.gw;
    int x0 = 0;
    for (int s0 = 0; s0 < ns; s0 += 31) {
      const int s = s0 + lane;
      int lo = 0x3fffffff, hi;
      float l = 0.f, h = 0.f;
      int pw = 0;
      if (s < ns) {
        pw = (int)(((float)s + 0.5f) * inv_gw);
        axis_tap(sample_coord(g.start_w, g.bin_w, pw, s - pw * g.gw, g.gw), p.W, lo, hi, l, h);
        if (lo >= p.W - 1) {  // clamped / beyond: 0 * F[W-2] + (1 or 0) * F[W-1]
          lo = p.W - 2;
          l = h;
          h = 0.f;
        }
      }
      if (s0 == 0) x0 = __shfl_sync(0xffffffffu, lo, 0);
      const int nlo = __shfl_down_sync(0xffffffffu, lo, 1);
      if (s < ns && lane < 31) {
        XSamp e;
        e.h = h * inv;
        e.l = l * inv;
        e.pw = pw;
        e.pad = 0;
        ba->xs[s] = e;
        const int c = lo - x0;
        if (c < 0 || c + 2 >= BMAXS + 4) {
          jump = true;  // cannot happen for unit sample steps; keeps the table writes in bounds
        } else {
        if (s == 0) ba->colstart[0] = 0;
        if (s < ns - 1) {
          const int d = nlo - lo;
          jump |= (d > 1 || d < 0);
          if (d == 1) ba->colstart[c + 1] = s + 1;
        } else {
          ba->colstart[c + 1] = ns;      // end of the last lower-tap column
          ba->colstart[c + 2] = ns;      // the column that only receives upper taps
          ba->ncols = c + 2;
        }
        }
      }
    }
    // ---- y samples: (hy | LASTROW, ly)
    const int nsy = P * g.gh;
    const float inv_gh = 1.f / (float)g.gh;
    for (int s0 = 0; s0 < nsy; s0 += 31) {
      const int s = s0 + lane;
      int lo = 0x3fffffff, hi;
      float l = 0.f, h = 0.f;
      if (s < nsy) {
        const int ph = (int)(((float)s + 0.5f) * inv_gh);
        axis_tap(sample_coord(g.start_h, g.bin_h, ph, s - ph * g.gh, g.gh), p.H, lo, hi, l, h);
        if (lo >= p.H - 1) {
          lo = p.H - 2;
          l = h;
          h = 0.f;
        }
      }
      const int nlo = __shfl_down_sync(0xffffffffu, lo, 1);
      if (s < nsy && lane < 31) {
        const bool last = (s == nsy - 1) || (nlo != lo);
        if (s < nsy - 1) jump |= (nlo - lo > 1 || nlo < lo);
        float2 e;
        e.x = __uint_as_float(__float_as_uint(h) | (last ? 0x80000000u : 0u));
        e.y = l;
        ba->ys[s] = e;
        if (s == 0) ba->y0 = lo;
      }
    }
    if (lane == 0) {
      ba->x0 = x0;
      ba->nsx = ns;
      ba->nsy = nsy;
    }
  }
  if (__any_sync(0xffffffffu, jump)) mode = 2;
  if (lane == 0) {
    ba->gw = g.gw;
    ba->gh = g.gh;
    ba->inv_count = 1.f / g.count;
    ba->mode = mode;
    ba->start_w = g.start_w;
    ba->start_h = g.start_h;
    ba->bin_w = g.bin_w;
    ba->bin_h = g.bin_h;
  }
}

// lane task: feature column x0 + c of channel pair cp
template <typename T>
__device__ __forceinline__ void bwd_column(float2* __restrict__ tile_pair, int W, const BwdArea<T>* ba, int c, int cp) {
  const int gh = ba->gh;
  const int cs0 = c > 0 ? ba->colstart[c - 1] : 0;
  const int cs1 = ba->colstart[c];
  const int cs2 = ba->colstart[c + 1];
  const int lo_begin = c > 0 ? cs0 : cs1;  // samples [lo_begin, cs1) contribute through their upper tap
  const T* g0 = ba->stage + (2 * cp) * (P * P);
  const T* g1 = g0 + P * P;
  float2* cell = tile_pair + ba->y0 * W + ba->x0 + c;
  float2 dlo = make_float2(0.f, 0.f), dhi = dlo;
  const float2* ys = ba->ys;
  for (int ph = 0; ph < P; ++ph) {
    float2 u = make_float2(0.f, 0.f);
    const T* r0 = g0 + ph * P;
    const T* r1 = g1 + ph * P;
    for (int s = lo_begin; s < cs1; ++s) {
      const XSamp e = ba->xs[s];
      u = ffma2(e.l, make_float2(ldf(r0 + e.pw), ldf(r1 + e.pw)), u);
    }
    for (int s = cs1; s < cs2; ++s) {
      const XSamp e = ba->xs[s];
      u = ffma2(e.h, make_float2(ldf(r0 + e.pw), ldf(r1 + e.pw)), u);
    }
    for (int iy = 0; iy < gh; ++iy) {
      const float2 t = *ys++;
      dlo = ffma2(fabsf(t.x), u, dlo);
      dhi = ffma2(t.y, u, dhi);
      if (__float_as_uint(t.x) >> 31) {
        if (dlo.x != 0.f || dlo.y != 0.f) smem_add2(cell, dlo);
        cell += W;
        dlo = dhi;
        dhi = make_float2(0.f, 0.f);
      }
    }
  }
  if (dlo.x != 0.f || dlo.y != 0.f) smem_add2(cell, dlo);  // the row that only receives upper taps
}

template <typename T>
__device__ __noinline__ void bwd_direct(float2* __restrict__ pl, int H, int W, const BwdArea<T>* ba, int q, int cp) {
  for (int half = 0; half < 2; ++half) {
    const int ph = q + 7 * half;
    for (int pw = 0; pw < P; ++pw) {
      const float g0 = ldf(ba->stage + (2 * cp) * (P * P) + ph * P + pw) * ba->inv_count;
      const float g1 = ldf(ba->stage + (2 * cp + 1) * (P * P) + ph * P + pw) * ba->inv_count;
      for (int iy = 0; iy < ba->gh; ++iy) {
        int ylo, yhi;
        float ly, hy;
        const bool vy = axis_tap(sample_coord(ba->start_h, ba->bin_h, ph, iy, ba->gh), H, ylo, yhi, ly, hy);
        for (int ix = 0; ix < ba->gw; ++ix) {
          int xlo, xhi;
          float lx, hx;
          const bool vx = axis_tap(sample_coord(ba->start_w, ba->bin_w, pw, ix, ba->gw), W, xlo, xhi, lx, hx);
          if (vy && vx) {
            smem_add2(pl + ylo * W + xlo, make_float2(g0 * hy * hx, g1 * hy * hx));
            smem_add2(pl + ylo * W + xhi, make_float2(g0 * hy * lx, g1 * hy * lx));
            smem_add2(pl + yhi * W + xlo, make_float2(g0 * ly * hx, g1 * ly * hx));
            smem_add2(pl + yhi * W + xhi, make_float2(g0 * ly * lx, g1 * ly * lx));
          }
        }
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(BNTHREADS, 1) roi_align_bwd_slab2(const Params p, float* __restrict__ gfeat32) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* slab = reinterpret_cast<float*>(smem_raw);
  const size_t slab_bytes = (size_t)NPAIR * p.pair_stride * sizeof(float2);
  BwdArea<T>* areas = reinterpret_cast<BwdArea<T>*>(smem_raw + ((slab_bytes + 127) / 128) * 128);
  __shared__ int s_next;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  BwdArea<T>* ba = areas + warp;
  const int nslab = p.C / CS;
  const int HW = p.H * p.W;
  const T* gout = reinterpret_cast<const T*>(p.feat);  // grad_out [R,C,14,14]
  float2* tile = reinterpret_cast<float2*>(slab);
  uint32_t parity = 0;
  constexpr uint32_t BLOCK_BYTES = CS * P * P * sizeof(T);
  if (lane == 0) mbar_init(&ba->bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

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
    const T* gbase = gout + ((long long)r_base * p.C + (long long)k * CS) * (P * P);
    const long long roi_stride = (long long)p.C * (P * P);

    for (int i = tid; i < NPAIR * p.pair_stride * 2; i += BNTHREADS) slab[i] = 0.f;
    if (tid == 0) s_next = r0;
    __syncthreads();

    int r = 0;
    if (lane == 0) r = atomicAdd(&s_next, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    while (r < r1) {
      int rn = 0;
      if (lane == 0) {
        mbar_expect_tx(&ba->bar, BLOCK_BYTES);
        bulk_load(ba->stage, gbase + (long long)r * roi_stride, BLOCK_BYTES, &ba->bar);
        rn = atomicAdd(&s_next, 1);
        if (rn < r1) bulk_prefetch_l2(gbase + (long long)rn * roi_stride, BLOCK_BYTES);
      }
      rn = __shfl_sync(0xffffffffu, rn, 0);
      build_bwd_tables<T>(p.rois + (long long)(r_base + r) * 5, p, ba, lane);
      __syncwarp();
      mbar_wait(&ba->bar, parity);
      parity ^= 1;
      const int mode = ba->mode;
      if (mode == 1) {
        const int ntask = ba->ncols * NPAIR;
        for (int t = lane; t < ntask; t += 32) {
          const int cp = t & (NPAIR - 1), c = t >> 2;
          bwd_column<T>(tile + (size_t)cp * p.pair_stride, p.W, ba, c, cp);
        }
      } else if (mode == 2 && lane < 28) {
        bwd_direct<T>(tile + (size_t)(lane & 3) * p.pair_stride, p.H, p.W, ba, lane >> 2, lane & 3);
      }
      __syncwarp();  // every lane is done with the staging block before the next bulk load overwrites it
      r = rn;
    }
    __syncthreads();  // the tile is complete
    float* dst = gfeat32 + ((long long)n * p.C + (long long)k * CS) * HW;
    const bool whole = (r0 == 0 && r1 == Rn);
    for (int e = tid; e < CS * HW; e += BNTHREADS) {
      const int c = e / HW, o = e - c * HW;
      const float v = slab[((size_t)(c >> 1) * p.pair_stride + o) * 2 + (c & 1)];
      if (whole) dst[e] = v;
      else if (v != 0.f) atomicAdd(dst + e, v);
    }
    __syncthreads();
    u = seg_end_u;
  }
}

__global__ void zero_f32_kernel(float* __restrict__ p, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = 0.f;
}
__global__ void cvt_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

template <typename T>
static size_t bwd_smem_total(int HW) {
  const size_t slab_bytes = (size_t)NPAIR * pair_stride_host(HW) * sizeof(float2);
  return ((slab_bytes + 127) / 128) * 128 + (size_t)BNWARPS * sizeof(BwdArea<T>);
}

}  // namespace v2

bool bwd_slab2_fits(int C, int H, int W, int dtype) {
  if (C % v2::CS) return false;
  const size_t need = dtype == UNIT_F32 ? v2::bwd_smem_total<float>(H * W) : v2::bwd_smem_total<__nv_bfloat16>(H * W);
  return need <= 227 * 1024;
}

// bf16 needs an fp32 accumulation image of grad_feat: N*C*H*W*4 bytes after the offsets
size_t bwd_slab2_workspace_bytes(int N, int C, int H, int W, int dtype) {
  return dtype == UNIT_BF16 ? (size_t)N * C * H * W * 4 : 0;
}

template <typename T>
static int launch_bwd_t(const void* gout, const float* rois, float* gfeat32, int N, int C, int H, int W, int R,
                        float scale, int sr, int aligned, const int* img_off, cudaStream_t st) {
  v2::Params p;
  p.feat = gout;
  p.rois = rois;
  p.out = nullptr;
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
  p.debug = 0;
  const size_t smem = v2::bwd_smem_total<T>(H * W);
  UNIT_CUDA(cudaFuncSetAttribute(v2::roi_align_bwd_slab2<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long grid = p.units_total / 44;
  if (grid < 1) grid = 1;
  if (grid > sm_count()) grid = sm_count();
  v2::roi_align_bwd_slab2<T><<<(int)grid, v2::BNTHREADS, smem, st>>>(p, gfeat32);
  UNIT_CHECK_LAUNCH("roi_align_bwd_slab2");
  return UNIT_OK;
}

int launch_bwd_slab2(const void* gout, const float* rois, void* gfeat, void* f32_scratch, int N, int C, int H, int W,
                     int R, float scale, int sr, int aligned, int dtype, const int* img_off, cudaStream_t st) {
  const long long total = (long long)N * C * H * W;
  const int zgrid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 8);
  if (dtype == UNIT_F32) {
    v2::zero_f32_kernel<<<zgrid, 256, 0, st>>>((float*)gfeat, total);
    UNIT_CHECK_LAUNCH("zero_f32_kernel");
    return launch_bwd_t<float>(gout, rois, (float*)gfeat, N, C, H, W, R, scale, sr, aligned, img_off, st);
  }
  v2::zero_f32_kernel<<<zgrid, 256, 0, st>>>((float*)f32_scratch, total);
  UNIT_CHECK_LAUNCH("zero_f32_kernel");
  int rc = launch_bwd_t<__nv_bfloat16>(gout, rois, (float*)f32_scratch, N, C, H, W, R, scale, sr, aligned, img_off, st);
  if (rc) return rc;
  v2::cvt_f32_bf16_kernel<<<zgrid, 256, 0, st>>>((const float*)f32_scratch, (__nv_bfloat16*)gfeat, total);
  UNIT_CHECK_LAUNCH("cvt_f32_bf16_kernel");
  return UNIT_OK;
}

}  // namespace roi
}  // namespace unit
