// Micro-benchmark: how fast can the SMs reduce into an L2-resident fp32 image?  (B200, sm_100a)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_rate red_rate.cu && ./red_rate
// Every warp adds `bytes_per_warp_instr` contiguous bytes per instruction at pseudo-random 256-byte aligned places of a
// 34 MB image (same footprint as the ROIAlign backward scratch), with .f32 / .v2.f32 / .v4.f32 reductions and with plain
// stores for comparison.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>  // 0: red.f32 (128 B / warp), 1: red.v2 (256 B), 2: red.v4 (512 B), 3: st.v2 (256 B)
__global__ void k(float* img, size_t n_floats, int iters) {
  const int lane = threadIdx.x & 31;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  unsigned state = warp * 2654435761u + 12345u;
  const size_t n_cells = n_floats / 128;  // 512-byte cells
  for (int i = 0; i < iters; ++i) {
    state = state * 1664525u + 1013904223u;
    float* cell = img + (size_t)(state % n_cells) * 128;
    if (MODE == 0) {
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(cell + lane), "f"(1.0f) : "memory");
    } else if (MODE == 1) {
      asm volatile("red.global.add.v2.f32 [%0], {%1, %1};" ::"l"(cell + 2 * lane), "f"(1.0f) : "memory");
    } else if (MODE == 2) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(cell + 4 * lane), "f"(1.0f) : "memory");
    } else {
      asm volatile("st.global.v2.f32 [%0], {%1, %1};" ::"l"(cell + 2 * lane), "f"(1.0f) : "memory");
    }
  }
}

template <int MODE>
void run(const char* name, float* img, size_t n, int bytes_per_instr) {
  const int blocks = 148 * 4, threads = 384, iters = 2000;
  k<MODE><<<blocks, threads>>>(img, n, 10);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<blocks, threads>>>(img, n, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double instr = (double)blocks * threads / 32 * iters;
  printf("%-12s %8.3f ms  %7.2f G warp-instr/s  %8.1f GB/s  %7.1f G sectors/s\n", name, ms, instr / ms / 1e6,
         instr * bytes_per_instr / ms / 1e6, instr * bytes_per_instr / 32 / ms / 1e6);
}

int main() {
  const size_t n = (size_t)2 * 1024 * 50 * 84;  // floats: 34.4 MB
  float* img;
  cudaMalloc(&img, n * 4);
  cudaMemset(img, 0, n * 4);
  run<0>("red.f32", img, n, 128);
  run<1>("red.v2.f32", img, n, 256);
  run<2>("red.v4.f32", img, n, 512);
  run<3>("st.v2.f32", img, n, 256);
  return 0;
}
