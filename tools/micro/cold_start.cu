// Microbenchmark: latency of the FIRST global load of a kernel, back-to-back launches (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cold_start cold_start.cu && ./cold_start
// Question: a conv kernel's first TMA tile / first statistics load lands ~4 k cycles after kernel entry even when the
// data sits in L2. Is that a per-launch cold start (TLB / L1 invalidation), and does it depend on how many SMs start at once?
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ long long clock_after(int dep) {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(dep) : "memory");      // the operand orders the read after the load
  return t;
}

__global__ void probe(const int* __restrict__ buf, size_t stride_ints, long long* out, int second_offset_ints) {
  const int* p = buf + (size_t)blockIdx.x * stride_ints;
  long long t0 = clock64();
  int v = __ldcg(p);
  v = __shfl_sync(0xffffffffu, v, 0);      // forces the wait for the load before the next clock read
  out[1024 * 4 - 1 - blockIdx.x] = v;       // in-order issue: the store needs v, the clock read follows it
  long long t1 = clock64();
  int w = __ldcg(p + second_offset_ints + (v & 1));       // dependent second load, other line / other page
  w = __shfl_sync(0xffffffffu, w, 0);
  out[1024 * 4 - 200 - blockIdx.x] = w;
  long long t2 = clock64();
  int x = __ldcg(p + 64 + (w & 1));                       // third: same page as the first, different line
  x = __shfl_sync(0xffffffffu, x, 0);
  out[1024 * 4 - 400 - blockIdx.x] = x;
  long long t3 = clock64();
  if (threadIdx.x != 0) return;
  out[blockIdx.x * 4 + 0] = t1 - t0;
  out[blockIdx.x * 4 + 3] = v + w + x;
  out[blockIdx.x * 4 + 1] = t2 - t1;
  out[blockIdx.x * 4 + 2] = t3 - t2;
}

int main() {
  const size_t bytes = 1ull << 30;
  int* buf;
  long long* out;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 0, bytes);
  cudaMalloc(&out, 1024 * 4 * sizeof(long long));
  long long h[1024 * 4];
  struct Case { const char* name; int grid; size_t stride; int second; };
  Case cases[] = {
      {"1 CTA, same line every launch", 1, 0, 1 << 20},
      {"148 CTAs, each its own 4 MB region (own pages)", 148, (4u << 20) / 4, 1 << 19},
      {"148 CTAs, all the same line", 148, 0, 1 << 20},
      {"148 CTAs, 256 B apart (one page)", 148, 64, 1 << 20},
  };
  for (const Case& c : cases) {
    for (int rep = 0; rep < 6; ++rep) {
      probe<<<c.grid, 32>>>(buf, c.stride, out, c.second);
      cudaDeviceSynchronize();
      cudaMemcpy(h, out, c.grid * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
      double a = 0, b = 0, d = 0, mx = 0;
      for (int i = 0; i < c.grid; ++i) { a += h[i * 4]; b += h[i * 4 + 1]; d += h[i * 4 + 2]; if (h[i * 4] > mx) mx = (double)h[i * 4]; }
      printf("%-50s launch %d: first load %6.0f cycles (max %6.0f) | 2nd (other page) %6.0f | 3rd (same page) %6.0f\n", c.name, rep,
             a / c.grid, mx, b / c.grid, d / c.grid);
    }
  }
  // back-to-back launches without host sync in between (as in a graph): only the last one is read
  for (int rep = 0; rep < 3; ++rep) {
    for (int k = 0; k < 20; ++k) probe<<<148, 32>>>(buf, (4u << 20) / 4, out, 1 << 19);
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, 148 * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
    double a = 0, b = 0;
    for (int i = 0; i < 148; ++i) { a += h[i * 4]; b += h[i * 4 + 1]; }
    printf("20 launches back to back, last one: first load %6.0f | 2nd %6.0f\n", a / 148, b / 148);
  }
  return 0;
}
