// Microbenchmark: is straight-line code slow the FIRST time it runs in a kernel (instruction-cache cold start), and is it
// cold again at every launch?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache_cold icache_cold.cu
// A 2048-instruction dependent FFMA chain (32 KB of SASS) is executed twice per launch by one warp per SM; both passes are
// timed. A dependent FFMA issues every ~4 cycles when the code is cached, i.e. ~8.2 k cycles per pass.
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__device__ __forceinline__ float chain(float x, float a) {
#pragma unroll
  for (int i = 0; i < N; ++i) x = fmaf(x, a, (float)(i & 15) * 0.125f);
  return x;
}

__global__ void probe(float a, long long* out, float* sink) {
  float x = (float)threadIdx.x;
  long long t[3];
  t[0] = clock64();
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    x = chain<2048>(x, a);
    sink[blockIdx.x * 32 + threadIdx.x] = x;      // the store needs x: the clock read below follows the chain
    t[pass + 1] = clock64();
  }
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t[1] - t[0]; out[blockIdx.x * 2 + 1] = t[2] - t[1]; }
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 148 * 2 * sizeof(long long));
  cudaMalloc(&sink, 148 * 32 * sizeof(float));
  long long h[148 * 2];
  for (int rep = 0; rep < 5; ++rep) {
    for (int k = 0; k < (rep < 3 ? 1 : 10); ++k) probe<<<148, 32>>>(0.999f, out, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double a = 0, b = 0;
    for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
    printf("launch group %d (%s): first pass %7.0f cycles, second pass %7.0f cycles (2048 dependent FFMAs each)\n", rep,
           rep < 3 ? "single launch" : "last of 10 back-to-back", a / 148, b / 148);
  }
  return 0;
}
