// Microbenchmark: MUFU throughput per SM for tanh.approx.f32, ex2.approx.f32, rcp.approx.f32 and
// tanh.approx.bf16x2, 8 warps per SM, 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i);
  uint32_t u[8];
  for (int i = 0; i < 8; ++i) u[i] = 0x3c003c00u + threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u[i]));
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  const char* names[5] = {"tanh.approx.f32", "ex2.approx.f32", "rcp.approx.f32", "tanh.approx.bf16x2", "fma.f32"};
  for (int op = 0; op < 5; ++op) {
    if (op == 0) k<0><<<148, 256>>>(out, iters, cyc);
    if (op == 1) k<1><<<148, 256>>>(out, iters, cyc);
    if (op == 2) k<2><<<148, 256>>>(out, iters, cyc);
    if (op == 3) k<3><<<148, 256>>>(out, iters, cyc);
    if (op == 4) k<4><<<148, 256>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double ops = 256.0 * 8 * iters;
    printf("%-20s %.2f lane-ops per cycle per SM (%.1f cycles per warp instruction per scheduler)\n", names[op], ops / h,
           (double)h / (2.0 * 8 * iters));
  }
  return 0;
}
