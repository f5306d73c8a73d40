// Micro-benchmark: does a DRAM read stream overlap with fp32 reductions into an L2-resident image?  (B200, sm_100a)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_red_mix read_red_mix.cu && ./read_red_mix
// The ROIAlign backward reads 0.82-1.17 GB of grad_out from DRAM and reduces 0.83 GB into a 34 MB image; its time
// equals (time of the reads alone) + (time of the reductions alone).  This kernel reproduces the mix without any of the
// kernel's arithmetic: every warp streams 512-byte chunks of a 1.2 GB buffer (ld.global.nc.v4, L1 no-allocate) and issues
// RED_PER reductions of 256 bytes (red.global.add.v2.f32) at pseudo-random places of the image per chunk.
#include <cstdio>
#include <cuda_runtime.h>

template <bool READ, int RED_PER, bool HINT>
__global__ void k(const float4* __restrict__ src, size_t n_chunks, float* img, size_t n_cells, float* sink) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  unsigned state = (unsigned)warp * 2654435761u + 12345u;
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  float acc = 0.f;
  for (size_t c = warp; c < n_chunks; c += nwarps) {
    float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
    if (READ) {
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "l"(src + c * 32 + lane));
      acc += v.x;
    }
#pragma unroll
    for (int r = 0; r < RED_PER; ++r) {
      state = state * 1664525u + 1013904223u;
      float* cell = img + (size_t)(state % n_cells) * 64 + 2 * lane;
      if (HINT)
        asm volatile("red.global.add.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(cell), "f"(v.x), "f"(v.y), "l"(pol)
                     : "memory");
      else
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(cell), "f"(v.x), "f"(v.y) : "memory");
    }
  }
  if (acc == 123.456f) *sink = acc;
}

template <bool READ, int RED_PER, bool HINT>
float run(const char* name, const float4* src, size_t n_chunks, float* img, size_t n_cells, float* sink) {
  const int blocks = 148 * 4, threads = 384;
  k<READ, RED_PER, HINT><<<blocks, threads>>>(src, n_chunks, img, n_cells, sink);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<READ, RED_PER, HINT><<<blocks, threads>>>(src, n_chunks, img, n_cells, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double rd = READ ? (double)n_chunks * 512 : 0.0, red = (double)n_chunks * RED_PER * 256;
  printf("%-34s %7.3f ms   read %6.3f GB (%6.0f GB/s)   red %6.3f GB (%6.0f GB/s)\n", name, ms, rd / 1e9, rd / ms / 1e6,
         red / 1e9, red / ms / 1e6);
  return ms;
}

int main() {
  const size_t n_chunks = (size_t)1200 * 1000 * 1000 / 512;  // 1.2 GB
  const size_t n_img = (size_t)2 * 1024 * 50 * 84;           // floats: 34.4 MB
  float4* src;
  float *img, *sink;
  cudaMalloc(&src, n_chunks * 512);
  cudaMalloc(&img, n_img * 4);
  cudaMalloc(&sink, 4);
  cudaMemset(src, 0, n_chunks * 512);
  cudaMemset(img, 0, n_img * 4);
  const size_t n_cells = n_img / 64;
  const float r = run<true, 0, false>("read only", src, n_chunks, img, n_cells, sink);
  const float d1 = run<false, 1, false>("red only (1 per chunk: 0.6 GB)", src, n_chunks, img, n_cells, sink);
  const float m1 = run<true, 1, false>("read + red (1 per chunk)", src, n_chunks, img, n_cells, sink);
  const float h1 = run<true, 1, true>("read + red evict_last hint", src, n_chunks, img, n_cells, sink);
  const float d2 = run<false, 2, false>("red only (2 per chunk: 1.2 GB)", src, n_chunks, img, n_cells, sink);
  const float m2 = run<true, 2, false>("read + red (2 per chunk)", src, n_chunks, img, n_cells, sink);
  printf("overlap: mixed/(read+red) = %.2f (1 per chunk), %.2f (2 per chunk); max(read,red)/mixed = %.2f, %.2f; hint %.3f ms\n",
         m1 / (r + d1), m2 / (r + d2), (r > d1 ? r : d1) / m1, (r > d2 ? r : d2) / m2, h1);
  return 0;
}
