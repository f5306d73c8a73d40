// Micro-benchmark: TMA bulk reduction (cp.reduce.async.bulk ... .add.f32, shared -> global) into an L2-resident fp32
// image, against red.global.add.v2.f32 from registers (tools/micro/red_rate.cu).  B200, sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bulk red_bulk.cu && ./red_bulk
// Every warp owns a CHUNK-byte buffer in shared memory and reduces it into pseudo-random CHUNK-aligned places of a
// 34 MB image; lane 0 issues the bulk op, `DEPTH` groups are kept in flight per warp.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int CHUNK, int DEPTH>
__global__ void k(float* img, size_t n_floats, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* buf = reinterpret_cast<float*>(smem + (size_t)warp * CHUNK);
  for (int i = lane; i < CHUNK / 4; i += 32) buf[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  unsigned state = gw * 2654435761u + 12345u;
  const size_t n_cells = n_floats * 4 / CHUNK;
  if (lane == 0) {
    for (int i = 0; i < iters; ++i) {
      state = state * 1664525u + 1013904223u;
      char* dst = reinterpret_cast<char*>(img) + (size_t)(state % n_cells) * CHUNK;
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
                   "r"((uint32_t)__cvta_generic_to_shared(buf)), "r"(CHUNK)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <int CHUNK, int DEPTH>
void run(float* img, size_t n) {
  const int blocks = 148, threads = 384;
  const int iters = 2000 * 256 / CHUNK * 4;
  const size_t smem = (size_t)(threads / 32) * CHUNK;
  cudaFuncSetAttribute(k<CHUNK, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<CHUNK, DEPTH><<<blocks, threads, smem>>>(img, n, 10);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<CHUNK, DEPTH><<<blocks, threads, smem>>>(img, n, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double ops = (double)blocks * threads / 32 * iters;
  printf("bulk.add.f32 chunk %5d B depth %d: %8.3f ms  %7.2f G ops/s  %8.1f GB/s  %7.1f G sectors/s  (%s)\n", CHUNK, DEPTH,
         ms, ops / ms / 1e6, ops * CHUNK / ms / 1e6, ops * CHUNK / 32 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const size_t n = (size_t)2 * 1024 * 50 * 84;  // floats: 34.4 MB
  float* img;
  cudaMalloc(&img, n * 4);
  cudaMemset(img, 0, n * 4);
  run<256, 1>(img, n);
  run<256, 4>(img, n);
  run<1024, 2>(img, n);
  run<1024, 4>(img, n);
  run<4096, 2>(img, n);
  run<4096, 4>(img, n);
  run<16384, 1>(img, n);
  return 0;
}
