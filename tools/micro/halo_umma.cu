// Microbenchmark / correctness probe for the "halo-resident" implicit-GEMM A operand:
// ONE TMA box load brings a (8+2) x (16+2) pixel x 64-channel halo tile (128B-swizzled, 128 B per
// pixel) into shared memory; the nine 3x3 taps are then nine tcgen05.mma A descriptors that start at
// pixel (ky*10 + kx) of that tile - i.e. at a 128-byte row that is NOT 1024-byte aligned - with
// SBO = 10 pixels = 1280 B between the 8-row groups (one group = 8 consecutive x of one output row).
// Question 1: does the hardware read the right elements (swizzle on absolute smem address bits)?
// Question 2: does the MMA run at the same rate as with 1024-byte aligned 8-row atoms?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o halo_umma halo_umma.cu -lcuda && ./halo_umma
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7u) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (clock64() - t0 > 2000000000LL) { printf("timeout bar %u\n", bar); __trap(); }
  }
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(id), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr int HALO_W = 10, HALO_H = 18;
constexpr int HALO_BYTES = HALO_W * HALO_H * 128;        // 23040
constexpr int HALO_STRIDE = 23552;                       // rounded up to 1024

// ---------------------------------------------------------------------------------------------- correctness
__global__ void __launch_bounds__(128, 1)
halo_conv_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap wmap, int x0, int y0,
                 int base_off_mode, float* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sW = base + HALO_STRIDE;       // W: 9 tiles of 64 rows x 128 B
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full = smem_u32(&bars[0]), done = smem_u32(&bars[1]);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(done));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(HALO_BYTES + 9 * 8192) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(sA), "l"(reinterpret_cast<uint64_t>(&amap)), "r"(full), "r"(0), "r"(x0 - 1), "r"(y0 - 1), "r"(0) : "memory");
    for (int t = 0; t < 9; ++t)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(sW + t * 8192), "l"(reinterpret_cast<uint64_t>(&wmap)), "r"(full), "r"(t * 64), "r"(0) : "memory");
    mbar_wait(full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t id = idesc(128, 64);
    for (int t = 0; t < 9; ++t) {
      const int ky = t / 3, kx = t % 3;
      for (int k = 0; k < 4; ++k) {
        const uint32_t a_addr = sA + (ky * HALO_W + kx) * 128 + k * 32;
        const uint32_t bo = base_off_mode ? ((a_addr >> 7) & 7u) : 0u;
        umma(tmem, desc_sw128(a_addr, HALO_W * 128, bo), desc_sw128(sW + t * 8192 + k * 32, 1024, 0), id, (t | k) ? 1u : 0u);
      }
    }
    commit(done);
  }
  __syncwarp();
  mbar_wait(done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------------------------- rate
// mode 0: aligned atoms (SBO 1024, each "tap" = its own 16 KB tile); mode 1: halo addressing (SBO 1280, tap offsets)
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int mode, int nwarps, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (warp < nwarps) {
    const uint32_t id = idesc(128, N);
    uint32_t elected;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
    const uint32_t sA = base + warp * 2 * HALO_STRIDE;        // each warp its own A region (>= 9*... see host)
    const uint32_t sW = base + 4 * HALO_STRIDE + warp * (N * 128);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int t = i % 9;
      const int ky = t / 3, kx = t % 3;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t da;
        if (mode == 0) da = desc_sw128(sA + (t & 1) * 16384 + k * 32, 1024, 0);
        else da = desc_sw128(sA + (ky * HALO_W + kx) * 128 + k * 32, HALO_W * 128, 0);
        const uint64_t db = desc_sw128(sW + k * 32, 1024, 0);
        if (elected) umma(tmem + (uint32_t)(warp * N), da, db, id, (i | k) ? 1u : 0u);
      }
    }
    const long long t1 = clock64();
    if (elected) commit(smem_u32(&bar[warp]));
    __syncwarp();
    mbar_wait(smem_u32(&bar[warp]), 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 3) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  const int H = 32, W = 32, C = 64, N = 64, K = 9 * 64;
  std::vector<__nv_bfloat16> hx((size_t)H * W * C), hw((size_t)N * K);
  std::vector<float> fx(hx.size()), fw(hw.size());
  srand(1);
  for (size_t i = 0; i < hx.size(); ++i) { fx[i] = (float)(rand() % 7 - 3); hx[i] = __float2bfloat16(fx[i]); }
  for (size_t i = 0; i < hw.size(); ++i) { fw[i] = (float)(rand() % 5 - 2); hw[i] = __float2bfloat16(fw[i]); }
  __nv_bfloat16 *dx, *dw;
  float* dout;
  CK(cudaMalloc(&dx, hx.size() * 2));
  CK(cudaMalloc(&dw, hw.size() * 2));
  CK(cudaMalloc(&dout, 128 * 64 * 4));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap amap, wmap;
  {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 1};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, HALO_W, HALO_H, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&amap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dw, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode W failed %d\n", (int)r); return 1; }
  }
  CK(cudaFuncSetAttribute(halo_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  const int origins[4][2] = {{0, 0}, {8, 16}, {24, 16}, {16, 0}};
  for (int mode = 0; mode < 2; ++mode)
    for (int o = 0; o < 4; ++o) {
      const int x0 = origins[o][0], y0 = origins[o][1];
      CK(cudaMemset(dout, 0, 128 * 64 * 4));
      halo_conv_kernel<<<1, 128, 110 * 1024>>>(amap, wmap, x0, y0, mode, dout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d origin (%d,%d): kernel error %s\n", mode, x0, y0, cudaGetErrorString(e)); return 1; }
      std::vector<float> got(128 * 64);
      CK(cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost));
      int bad = 0;
      double maxerr = 0;
      for (int r = 0; r < 128; ++r) {
        const int y = y0 + r / 8, x = x0 + r % 8;
        for (int n = 0; n < N; ++n) {
          float ref = 0;
          for (int t = 0; t < 9; ++t) {
            const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            for (int c = 0; c < C; ++c) ref += fx[((size_t)yy * W + xx) * C + c] * fw[(size_t)n * K + t * 64 + c];
          }
          const double err = fabs((double)ref - got[r * 64 + n]);
          if (err > maxerr) maxerr = err;
          bad += err > 0.5;
        }
      }
      printf("halo conv: base_offset mode %d, tile origin (x %2d, y %2d): %d / %d wrong, max err %.1f\n", mode, x0, y0, bad,
             128 * 64, maxerr);
    }

  long long* d;
  CK(cudaMalloc(&d, 148 * 2 * sizeof(long long)));
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int iters = 1800;
  for (int nw = 1; nw <= 2; ++nw)
    for (int mode = 0; mode < 2; ++mode)
      for (int Nn : {64, 128, 256}) {
        if (nw * 2 * Nn > 512 && nw == 2 && Nn == 256) {}
        rate_kernel<<<148, 128, 200 * 1024>>>(Nn, iters, mode, nw, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("rate error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[2];
        CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        printf("rate: %d issuing warp(s), %s A addressing, N %3d: %.1f cycles per MMA per warp (issue %.1f)\n", nw,
               mode ? "halo (SBO 1280, tap offsets)" : "aligned (SBO 1024)", Nn, (double)h[1] / iters / 4,
               (double)h[0] / iters / 4);
      }
  return 0;
}
