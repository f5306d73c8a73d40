// ROIAlign backward, slab-resident kernel.
//
// Work item = (RoI, group of 16 eight-channel slabs), pulled from a global counter by the 15 warps of each of the 148
// persistent CTAs.  A warp builds the RoI's sampling tables ONCE (same fp32 operation order as the forward and as
// torchvision) and reuses them for the 16 slabs; each slab's [8][14][14] grad_out block -- 6272 contiguous bytes --
// streams in through a double-buffered TMA bulk load (cp.async.bulk.shared::cta.global + mbarrier).
// The scatter is column-owned: lane (pair of feature columns, channel pair) gathers, per bin row, the horizontally
// spread values of its two columns from three consecutive lower-tap groups of the x-table, spreads them vertically
// through a two-row register window, and emits exactly ONE 16-byte `red.global.add.v4.f32` per footprint cell pair
// (2 columns x 2 channels) into an L2-resident, channel-pair interleaved fp32 image of grad_feat, which a final
// streaming kernel converts to NCHW.  That is footprint/2 vector atomics per RoI and channel pair instead of
// torchvision's 4*gh*gw scalar atomics per output element (~1e9 for 2x512 RoIs), with no shared-memory tile at all.
#include "roi_slab.cuh"

namespace unit {
namespace roi {
namespace v2 {

constexpr int BMAXG = 4;             // sampling grid handled with tables in the backward (RoI side <= 56 px)
constexpr int BMAXS = P * BMAXG;
constexpr int BNWARPS = 15;
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
  int colstart[BMAXS + 8];   // colstart[c + 2] = first sample whose lower tap is >= x0 + c, for c in [-2, ncols + 1]
  float2 ys[BMAXS + 2];      // (hy | LASTROW sign, ly)
  __align__(16) T stage[2][CS * P * P];
  uint64_t bar[2];
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
        if (c < 0 || c + 5 >= BMAXS + 8) {
          jump = true;  // cannot happen for unit sample steps; keeps the table writes in bounds
        } else {
          if (s == 0) {
            ba->colstart[0] = 0;  // c = -2
            ba->colstart[1] = 0;  // c = -1
            ba->colstart[2] = 0;  // c = 0
          }
          if (s < ns - 1) {
            const int d = nlo - lo;
            jump |= (d > 1 || d < 0);
            if (d == 1) ba->colstart[c + 3] = s + 1;
          } else {
            ba->colstart[c + 3] = ns;  // end of the last lower-tap column
            ba->colstart[c + 4] = ns;  // the column that only receives upper taps
            ba->colstart[c + 5] = ns;
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

// lane task: feature columns (xA, xA+1) (xA even) of channel pair cp.  cA = xA - x0 may be -1.
// Gathers, per bin row, the horizontally spread values of both columns from three consecutive lower-tap groups of
// the x-table, spreads them vertically through a two-row register window, and issues ONE 16-byte
// red.global.add.v4.f32 per (row, column pair) into the pair-interleaved fp32 scratch image.
template <typename T>
__device__ __forceinline__ void bwd_column_pair(float* __restrict__ scratch_pair, int W, const BwdArea<T>* ba,
                                                const T* __restrict__ stage, int cA, int cp) {
  const int gh = ba->gh;
  const int* cs = ba->colstart + 2;  // cs[c] valid for c in [-2, ncols + 1]
  const int s0 = cs[cA - 1], s1 = cs[cA], s2 = cs[cA + 1], s3 = cs[cA + 2];
  const T* g0 = stage + (2 * cp) * (P * P);
  const T* g1 = g0 + P * P;
  float* cell = scratch_pair + ((size_t)ba->y0 * W + (ba->x0 + cA)) * 2;
  float2 loA = make_float2(0.f, 0.f), loB = loA, hiA = loA, hiB = loA;
  const float2* ys = ba->ys;
  for (int ph = 0; ph < P; ++ph) {
    float2 uA = make_float2(0.f, 0.f), uB = uA;
    const T* r0 = g0 + ph * P;
    const T* r1 = g1 + ph * P;
    for (int s = s0; s < s1; ++s) {
      const XSamp e = ba->xs[s];
      uA = ffma2(e.l, make_float2(ldf(r0 + e.pw), ldf(r1 + e.pw)), uA);
    }
    for (int s = s1; s < s2; ++s) {
      const XSamp e = ba->xs[s];
      const float2 g = make_float2(ldf(r0 + e.pw), ldf(r1 + e.pw));
      uA = ffma2(e.h, g, uA);
      uB = ffma2(e.l, g, uB);
    }
    for (int s = s2; s < s3; ++s) {
      const XSamp e = ba->xs[s];
      uB = ffma2(e.h, make_float2(ldf(r0 + e.pw), ldf(r1 + e.pw)), uB);
    }
    for (int iy = 0; iy < gh; ++iy) {
      const float2 t = *ys++;
      const float hy = fabsf(t.x);
      loA = ffma2(hy, uA, loA);
      loB = ffma2(hy, uB, loB);
      hiA = ffma2(t.y, uA, hiA);
      hiB = ffma2(t.y, uB, hiB);
      if (__float_as_uint(t.x) >> 31) {
        if (loA.x != 0.f || loA.y != 0.f || loB.x != 0.f || loB.y != 0.f)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "f"(loA.x), "f"(loA.y),
                       "f"(loB.x), "f"(loB.y)
                       : "memory");
        cell += 2 * W;
        loA = hiA;
        loB = hiB;
        hiA = make_float2(0.f, 0.f);
        hiB = make_float2(0.f, 0.f);
      }
    }
  }
  if (loA.x != 0.f || loA.y != 0.f || loB.x != 0.f || loB.y != 0.f)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "f"(loA.x), "f"(loA.y), "f"(loB.x),
                 "f"(loB.y)
                 : "memory");
}

template <typename T>
__device__ __noinline__ void bwd_direct(float* __restrict__ scratch_pair, int H, int W, const BwdArea<T>* ba,
                                        const T* __restrict__ stage, int q, int cp) {
  for (int half = 0; half < 2; ++half) {
    const int ph = q + 7 * half;
    for (int pw = 0; pw < P; ++pw) {
      const float g0 = ldf(stage + (2 * cp) * (P * P) + ph * P + pw) * ba->inv_count;
      const float g1 = ldf(stage + (2 * cp + 1) * (P * P) + ph * P + pw) * ba->inv_count;
      for (int iy = 0; iy < ba->gh; ++iy) {
        int ylo, yhi;
        float ly, hy;
        const bool vy = axis_tap(sample_coord(ba->start_h, ba->bin_h, ph, iy, ba->gh), H, ylo, yhi, ly, hy);
        for (int ix = 0; ix < ba->gw; ++ix) {
          int xlo, xhi;
          float lx, hx;
          const bool vx = axis_tap(sample_coord(ba->start_w, ba->bin_w, pw, ix, ba->gw), W, xlo, xhi, lx, hx);
          if (vy && vx) {
            const int idx[4] = {ylo * W + xlo, ylo * W + xhi, yhi * W + xlo, yhi * W + xhi};
            const float w[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              atomicAdd(scratch_pair + 2 * (size_t)idx[k], g0 * w[k]);
              atomicAdd(scratch_pair + 2 * (size_t)idx[k] + 1, g1 * w[k]);
            }
          }
        }
      }
    }
  }
}

constexpr int SLABS_PER_ITEM = 16;  // a warp reuses one RoI's tables for 16 consecutive 8-channel slabs

template <typename T>
__global__ void __launch_bounds__(BNTHREADS, 1)
roi_align_bwd_slab2(const Params p, float* __restrict__ scratch, int* __restrict__ item_counter) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdArea<T>* areas = reinterpret_cast<BwdArea<T>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  BwdArea<T>* ba = areas + warp;
  const int nslab = p.C / CS;
  const int HW = p.H * p.W;
  const T* gout = reinterpret_cast<const T*>(p.feat);  // grad_out [R,C,14,14]
  constexpr uint32_t BLOCK_BYTES = CS * P * P * sizeof(T);
  uint32_t parity = 0;  // bit b = phase parity of barrier b
  if (lane == 0) {
    mbar_init(&ba->bar[0], 1);
    mbar_init(&ba->bar[1], 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

  const int groups = (nslab + SLABS_PER_ITEM - 1) / SLABS_PER_ITEM;
  const long long n_items = (long long)p.R * groups;
  while (true) {
    long long item = 0;
    if (lane == 0) item = atomicAdd(item_counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    const int r = (int)(item / groups);
    const int k0 = (int)(item - (long long)r * groups) * SLABS_PER_ITEM;
    const int k1 = min(nslab, k0 + SLABS_PER_ITEM);
    const float* roi = p.rois + (long long)r * 5;
    const int n = (int)roi[0];
    const T* gbase = gout + (long long)r * p.C * (P * P);
    if (lane == 0) {
      for (int j = 0; j < 2 && k0 + j < k1; ++j) {
        mbar_expect_tx(&ba->bar[j], BLOCK_BYTES);
        bulk_load(ba->stage[j], gbase + (long long)(k0 + j) * CS * (P * P), BLOCK_BYTES, &ba->bar[j]);
      }
    }
    build_bwd_tables<T>(roi, p, ba, lane);
    __syncwarp();
    const int mode = ba->mode;
    const bool in_img = n >= 0 && n < p.N;
    const int xe = ba->x0 & ~1;                                   // even column the first pair starts at
    const int npair = mode == 1 ? ((ba->x0 + ba->ncols - 1 - xe) >> 1) + 1 : 0;
    for (int k = k0; k < k1; ++k) {
      const int buf = (k - k0) & 1;
      mbar_wait(&ba->bar[buf], (parity >> buf) & 1u);
      parity ^= 1u << buf;
      float* spair_base = scratch + ((size_t)n * (p.C / 2) + (size_t)k * NPAIR) * HW * 2;
      if (in_img) {
        if (mode == 1) {
          const int ntask = npair * NPAIR;
          for (int t = lane; t < ntask; t += 32) {
            const int cp = t & (NPAIR - 1), pr = t >> 2;
            bwd_column_pair<T>(spair_base + (size_t)cp * HW * 2, p.W, ba, ba->stage[buf], xe + 2 * pr - ba->x0, cp);
          }
        } else if (mode == 2 && lane < 28) {
          bwd_direct<T>(spair_base + (size_t)(lane & 3) * HW * 2, p.H, p.W, ba, ba->stage[buf], lane >> 2, lane & 3);
        }
      }
      __syncwarp();  // every lane is done with this staging buffer
      if (lane == 0 && k + 2 < k1) {
        mbar_expect_tx(&ba->bar[buf], BLOCK_BYTES);
        bulk_load(ba->stage[buf], gbase + (long long)(k + 2) * CS * (P * P), BLOCK_BYTES, &ba->bar[buf]);
      }
    }
  }
}

// scratch [N][C/2][HW][2] fp32 -> grad_feat [N][C][HW] (T)
template <typename T>
__global__ void unpair_kernel(const float2* __restrict__ scratch, T* __restrict__ out, long long n_pairs_hw, int HW) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pairs_hw;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pair = i / HW;
    const int o = (int)(i - pair * HW);
    const float2 v = scratch[i];
    stf(out + (2 * pair) * HW + o, v.x);
    stf(out + (2 * pair + 1) * HW + o, v.y);
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
static size_t bwd_smem_total() { return (size_t)BNWARPS * sizeof(BwdArea<T>); }

}  // namespace v2

bool bwd_slab2_fits(int C, int H, int W, int dtype) {
  (void)dtype;
  return (C % v2::CS) == 0 && (W % 2) == 0 && H >= 2 && W >= 2;
}

// fp32 pair-interleaved accumulation image of grad_feat + the work-item counter
size_t bwd_slab2_workspace_bytes(int N, int C, int H, int W, int dtype) {
  (void)dtype;
  return (size_t)N * C * H * W * 4 + 256;
}

template <typename T>
static int launch_bwd_t(const void* gout, const float* rois, void* gfeat, void* ws, int N, int C, int H, int W, int R,
                        float scale, int sr, int aligned, cudaStream_t st) {
  v2::Params p;
  p.feat = gout;
  p.rois = rois;
  p.out = nullptr;
  p.img_off = nullptr;
  p.N = N;
  p.C = C;
  p.H = H;
  p.W = W;
  p.R = R;
  p.scale = scale;
  p.sampling_ratio = sr;
  p.aligned = aligned;
  p.pair_stride = 0;
  p.units_total = 0;
  p.debug = 0;
  const long long total = (long long)N * C * H * W;
  int* counter = (int*)ws;
  float* scratch = (float*)((char*)ws + 256);
  const int zgrid = (int)std::min<long long>((total + 64 + 255) / 256, (long long)sm_count() * 8);
  v2::zero_f32_kernel<<<zgrid, 256, 0, st>>>((float*)ws, total + 64);
  UNIT_CHECK_LAUNCH("zero_f32_kernel");
  const size_t smem = v2::bwd_smem_total<T>();
  UNIT_CUDA(cudaFuncSetAttribute(v2::roi_align_bwd_slab2<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int groups = (C / v2::CS + v2::SLABS_PER_ITEM - 1) / v2::SLABS_PER_ITEM;
  long long grid = ((long long)R * groups + v2::BNWARPS - 1) / v2::BNWARPS;
  if (grid > sm_count()) grid = sm_count();
  if (grid < 1) grid = 1;
  v2::roi_align_bwd_slab2<T><<<(int)grid, v2::BNTHREADS, smem, st>>>(p, scratch, counter);
  UNIT_CHECK_LAUNCH("roi_align_bwd_slab2");
  v2::unpair_kernel<T><<<zgrid, 256, 0, st>>>((const float2*)scratch, (T*)gfeat, total / 2, H * W);
  UNIT_CHECK_LAUNCH("unpair_kernel");
  return UNIT_OK;
}

int launch_bwd_slab2(const void* gout, const float* rois, void* gfeat, void* ws, int N, int C, int H, int W, int R,
                     float scale, int sr, int aligned, int dtype, const int* img_off, cudaStream_t st) {
  (void)img_off;
  if (dtype == UNIT_F32) return launch_bwd_t<float>(gout, rois, gfeat, ws, N, C, H, W, R, scale, sr, aligned, st);
  return launch_bwd_t<__nv_bfloat16>(gout, rois, gfeat, ws, N, C, H, W, R, scale, sr, aligned, st);
}

}  // namespace roi
}  // namespace unit
