// ROIAlign backward, slab-resident kernel v2.
//
// Transpose of roi_align_fwd.cu with the same decomposition: a persistent CTA keeps the GRADIENT tile of one
// (image, 8-channel slab) -- channel-pair interleaved float2, 134 KB for the 50x84 map -- in shared memory; each of
// its 12 warps pulls RoIs of that slab, streams the RoI's [8][14][14] grad_out block in with one TMA bulk load
// (cp.async.bulk.shared::cta.global + mbarrier, overlapped with the table build), and lane (q, cp) scatters rows
// {q, q+7} of channels {2cp, 2cp+1}: horizontal spread in registers over a two-column window, vertical spread with
// one 64-bit shared-memory CAS-add per tap (both channels at once).  The tile goes to HBM once per segment:
// torchvision issues gh*gw*4 global atomics per output element (~1e9 for 2x512 RoIs), this kernel issues none in the
// steady state (only the per-segment tile merge uses red.global).
#include "roi_slab.cuh"

namespace unit {
namespace roi {
namespace v2 {

__device__ __forceinline__ void smem_add2(float2* addr, float2 v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    float2 cur;
    cur.x = __uint_as_float((unsigned)(assumed & 0xffffffffull));
    cur.y = __uint_as_float((unsigned)(assumed >> 32));
    cur.x += v.x;
    cur.y += v.y;
    const unsigned long long nv =
        ((unsigned long long)__float_as_uint(cur.y) << 32) | (unsigned long long)__float_as_uint(cur.x);
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

template <int GH>
struct BTaps {
  float2* alo[GH > 0 ? GH : 1];
  float2* ahi[GH > 0 ? GH : 1];
  float2* blo[GH > 0 ? GH : 1];
  float2* bhi[GH > 0 ? GH : 1];
  float ah[GH > 0 ? GH : 1], al[GH > 0 ? GH : 1], bh[GH > 0 ? GH : 1], bl[GH > 0 ? GH : 1];
};

template <int GH>
__device__ __forceinline__ void flush(const BTaps<GH>& t, float2* __restrict__ pl, const YTap* ya, const YTap* yb,
                                      int gh, int col, float2 da, float2 db) {
  const bool za = (da.x == 0.f && da.y == 0.f), zb = (db.x == 0.f && db.y == 0.f);
  if (GH > 0) {
#pragma unroll
    for (int i = 0; i < GH; ++i) {
      if (!za) {
        smem_add2(t.alo[i] + col, make_float2(t.ah[i] * da.x, t.ah[i] * da.y));
        if (t.al[i] != 0.f) smem_add2(t.ahi[i] + col, make_float2(t.al[i] * da.x, t.al[i] * da.y));
      }
      if (!zb) {
        smem_add2(t.blo[i] + col, make_float2(t.bh[i] * db.x, t.bh[i] * db.y));
        if (t.bl[i] != 0.f) smem_add2(t.bhi[i] + col, make_float2(t.bl[i] * db.x, t.bl[i] * db.y));
      }
    }
  } else {
    for (int i = 0; i < gh; ++i) {
      const YTap a = ya[i], b = yb[i];
      if (!za) {
        smem_add2(pl + a.lo + col, make_float2(a.h * da.x, a.h * da.y));
        if (a.l != 0.f) smem_add2(pl + a.hi + col, make_float2(a.l * da.x, a.l * da.y));
      }
      if (!zb) {
        smem_add2(pl + b.lo + col, make_float2(b.h * db.x, b.h * db.y));
        if (b.l != 0.f) smem_add2(pl + b.hi + col, make_float2(b.l * db.x, b.l * db.y));
      }
    }
  }
}

template <typename T, int GH>
__device__ __forceinline__ void bwd_task(float2* __restrict__ pl, const WarpArea<T>* wa, int q, int cp) {
  const int gh = wa->hdr.gh;
  const YTap* ya = wa->ytab + q * gh;
  const YTap* yb = wa->ytab + (q + 7) * gh;
  BTaps<GH> t;
  if (GH > 0) {
#pragma unroll
    for (int i = 0; i < GH; ++i) {
      const YTap a = ya[i], b = yb[i];
      t.alo[i] = pl + a.lo;
      t.ahi[i] = pl + a.hi;
      t.blo[i] = pl + b.lo;
      t.bhi[i] = pl + b.hi;
      t.ah[i] = a.h;
      t.al[i] = a.l;
      t.bh[i] = b.h;
      t.bl[i] = b.l;
    }
  }
  int col = wa->hdr.x0;
  int remaining = wa->hdr.nsamp;
  const T* ga = wa->stage + (2 * cp) * (P * P) + q * P;  // grad_out row q, channel 2cp (2cp+1 is P*P further)
  const T* gb = ga + 7 * P;
  float2 g_a = make_float2(ldf(ga), ldf(ga + P * P)), g_b = make_float2(ldf(gb), ldf(gb + P * P));
  float2 dlo_a = make_float2(0.f, 0.f), dlo_b = dlo_a, dhi_a = dlo_a, dhi_b = dlo_a;
  const float2* xt = wa->xtab;
  float2 e = xt[0];
  while (remaining > 0) {
    const float2 en = xt[1];
    ++xt;
    const bool lastcol = (__float_as_uint(e.x) >> 31) != 0;
    const bool end = (__float_as_uint(e.y) >> 31) != 0;
    const float hx = fabsf(e.x), lx = fabsf(e.y);  // already scaled by 1/count
    dlo_a = ffma2(hx, g_a, dlo_a);
    dlo_b = ffma2(hx, g_b, dlo_b);
    dhi_a = ffma2(lx, g_a, dhi_a);
    dhi_b = ffma2(lx, g_b, dhi_b);
    --remaining;
    if (end && remaining > 0) {
      ++ga;
      ++gb;
      g_a = make_float2(ldf(ga), ldf(ga + P * P));
      g_b = make_float2(ldf(gb), ldf(gb + P * P));
    }
    if (lastcol) {
      flush<GH>(t, pl, ya, yb, gh, col, dlo_a, dlo_b);
      ++col;
      dlo_a = dhi_a;
      dlo_b = dhi_b;
      dhi_a = make_float2(0.f, 0.f);
      dhi_b = make_float2(0.f, 0.f);
    }
    e = en;
  }
  // the upper tap of the last column (col == last lo + 1 <= W-1)
  flush<GH>(t, pl, ya, yb, gh, col, dlo_a, dlo_b);
}

template <typename T>
__device__ __noinline__ void bwd_task_direct(float2* __restrict__ pl, int H, int W, const Header& hdr, int q, int cp,
                                             const T* __restrict__ stage) {
  for (int half = 0; half < 2; ++half) {
    const int ph = q + 7 * half;
    for (int pw = 0; pw < P; ++pw) {
      const float g0 = ldf(stage + (2 * cp) * (P * P) + ph * P + pw) * hdr.inv_count;
      const float g1 = ldf(stage + (2 * cp + 1) * (P * P) + ph * P + pw) * hdr.inv_count;
      for (int iy = 0; iy < hdr.gh; ++iy) {
        int ylo, yhi;
        float ly, hy;
        const bool vy = axis_tap(sample_coord(hdr.start_h, hdr.bin_h, ph, iy, hdr.gh), H, ylo, yhi, ly, hy);
        for (int ix = 0; ix < hdr.gw; ++ix) {
          int xlo, xhi;
          float lx, hx;
          const bool vx = axis_tap(sample_coord(hdr.start_w, hdr.bin_w, pw, ix, hdr.gw), W, xlo, xhi, lx, hx);
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
struct BwdArea {
  WarpArea<T> wa;
  uint64_t bar;
  uint64_t pad;
};

template <typename T>
__global__ void __launch_bounds__(NTHREADS, 1) roi_align_bwd_slab2(const Params p, float* __restrict__ gfeat32) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* slab = reinterpret_cast<float*>(smem_raw);
  const size_t slab_bytes = (size_t)NPAIR * p.pair_stride * sizeof(float2);
  BwdArea<T>* areas = reinterpret_cast<BwdArea<T>*>(smem_raw + ((slab_bytes + 127) / 128) * 128);
  __shared__ int s_next;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  BwdArea<T>* ba = areas + warp;
  WarpArea<T>* wa = &ba->wa;
  const int nslab = p.C / CS;
  const int HW = p.H * p.W;
  const T* gout = reinterpret_cast<const T*>(p.feat);  // grad_out [R,C,14,14]
  const int q = lane >> 2, cp = lane & 3;
  float2* pl = reinterpret_cast<float2*>(slab) + (size_t)cp * p.pair_stride;
  uint32_t parity = 0;
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

    for (int i = tid; i < NPAIR * p.pair_stride * 2; i += NTHREADS) slab[i] = 0.f;
    if (tid == 0) s_next = r0;
    __syncthreads();

    while (true) {
      int r = 0;
      if (lane == 0) r = atomicAdd(&s_next, 1);
      r = __shfl_sync(0xffffffffu, r, 0);
      if (r >= r1) break;
      if (lane == 0) {
        mbar_expect_tx(&ba->bar, (uint32_t)(CS * P * P * sizeof(T)));
        bulk_load(wa->stage, gout + ((long long)(r_base + r) * p.C + (long long)k * CS) * (P * P),
                  (uint32_t)(CS * P * P * sizeof(T)), &ba->bar);
      }
      build_tables<T>(p.rois + (long long)(r_base + r) * 5, p, wa, lane);
      __syncwarp();
      mbar_wait(&ba->bar, parity);
      parity ^= 1;
      if (lane < 28) {
        const int mode = wa->hdr.mode;
        if (mode == 1) {
          const int gh = wa->hdr.gh;
          if (gh == 1) bwd_task<T, 1>(pl, wa, q, cp);
          else if (gh == 2) bwd_task<T, 2>(pl, wa, q, cp);
          else bwd_task<T, 0>(pl, wa, q, cp);
        } else if (mode == 2) {
          bwd_task_direct<T>(pl, p.H, p.W, wa->hdr, q, cp, wa->stage);
        }
      }
      __syncwarp();  // every lane is done with the staging block before the next bulk load overwrites it
    }
    __syncthreads();  // the tile is complete
    // merge the tile into grad_feat (fp32): plain stores when this CTA owns the whole (image, slab), atomics otherwise
    float* dst = gfeat32 + ((long long)n * p.C + (long long)k * CS) * HW;
    const bool whole = (r0 == 0 && r1 == Rn);
    for (int e = tid; e < CS * HW; e += NTHREADS) {
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
  return ((slab_bytes + 127) / 128) * 128 + (size_t)NWARPS * sizeof(BwdArea<T>);
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
  // every (image, slab) must be visited even when it has few RoIs: the grid walks contiguous unit ranges
  long long grid = p.units_total / 48;
  if (grid < 1) grid = 1;
  if (grid > sm_count()) grid = sm_count();
  v2::roi_align_bwd_slab2<T><<<(int)grid, v2::NTHREADS, smem, st>>>(p, gfeat32);
  UNIT_CHECK_LAUNCH("roi_align_bwd_slab2");
  return UNIT_OK;
}

// grad_feat must be zero-filled by the caller of this function (images without RoIs and split slabs rely on it).
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
