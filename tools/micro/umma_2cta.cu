// Correctness + rate probe for tcgen05.mma.cta_group::2 (a CTA pair on one TPC sharing the B operand):
// D[256 x N] = A[256 x K] * B[N x K]^T, K = 64 per stage. CTA r of the pair holds A rows [128r, 128r+128)
// and B rows [N/2 * r, N/2 * (r+1)) in ITS shared memory at the same offsets; the leader (rank 0) issues
// the MMAs; the result rows [128r, 128r+128) land in CTA r's TMEM. TMA loads of both CTAs signal the
// LEADER's mbarrier (cta_group::2 form, peer bit cleared); tcgen05.commit multicasts to both CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_2cta umma_2cta.cu -lcuda && ./umma_2cta
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
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
    if (clock64() - t0 > 2000000000LL) { printf("timeout bar %x block %d\n", bar, blockIdx.x); __trap(); }
  }
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_gemm_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, int kblocks,
                 int iters, float* out, long long* cyc, int feed) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  constexpr int STAGES = 4;
  constexpr int A_BYTES = 128 * 128, B_BYTES = (N / 2) * 128, ST = A_BYTES + B_BYTES;
  __shared__ uint64_t full[STAGES], empty[STAGES], done;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t r = cluster_rank();
  const int pair = blockIdx.x >> 1;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "n"(N < 32 ? 32 : N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;

  long long t0 = 0, t1 = 0;
  if (warp == 0 && lane == 0 && feed) {
    // producer (both CTAs): loads its A rows and its half of B, signals the LEADER's full barrier
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it)
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
        const uint32_t fb = smem_u32(&full[s]) & PEER_MASK;
        if (r == 0)      // the leader arms its barrier for the bytes of BOTH CTAs
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(2 * ST) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(base + s * ST), "l"(reinterpret_cast<uint64_t>(&amap)), "r"(fb), "r"(kb * 64), "r"(pair * 256 + (int)r * 128) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(base + s * ST + A_BYTES), "l"(reinterpret_cast<uint64_t>(&bmap)), "r"(fb), "r"(kb * 64), "r"((int)r * (N / 2)) : "memory");
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
  } else if (warp == 2 && lane == 0 && r == 0) {
    // MMA issuer (leader only)
    const uint32_t id = idesc(256, N);
    int s = 0; uint32_t ph = 0;
    t0 = clock64();
    for (int it = 0; it < iters; ++it)
      for (int kb = 0; kb < kblocks; ++kb) {
        if (feed) mbar_wait(smem_u32(&full[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = base + s * ST, sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (kb | k) ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                       ::"r"(tmem), "l"(desc_sw128(sa + k * 32)), "l"(desc_sw128(sb + k * 32)), "r"(id), "r"(acc), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&empty[s])), "h"((uint16_t)3) : "memory");
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&done)), "h"((uint16_t)3) : "memory");
  }
  __syncwarp();
  mbar_wait(smem_u32(&done), 0);
  t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 2 && lane == 0 && r == 0 && cyc) cyc[pair] = t1 - t0;
  if (out) {
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const size_t row = (size_t)pair * 256 + r * 128 + warp * 32 + lane;
      for (int j = 0; j < 8; ++j) out[row * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(N < 32 ? 32 : N) : "memory");
}

static void encode2d(CUtensorMap* m, void* p, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, strides, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

template <int N>
static void run(int pairs, int kblocks, bool check) {
  const int M = pairs * 256, K = kblocks * 64;
  std::vector<__nv_bfloat16> ha((size_t)M * K), hb((size_t)N * K);
  std::vector<float> fa(ha.size()), fb(hb.size());
  srand(7);
  for (size_t i = 0; i < ha.size(); ++i) { fa[i] = (float)(rand() % 5 - 2); ha[i] = __float2bfloat16(fa[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (float)(rand() % 5 - 2); hb[i] = __float2bfloat16(fb[i]); }
  __nv_bfloat16 *da, *db;
  float* dout;
  long long* dcyc;
  CK(cudaMalloc(&da, ha.size() * 2));
  CK(cudaMalloc(&db, hb.size() * 2));
  CK(cudaMalloc(&dout, (size_t)M * N * 4));
  CK(cudaMalloc(&dcyc, pairs * 8));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap amap, bmap;
  encode2d(&amap, da, K, M, 128);
  encode2d(&bmap, db, K, N, N / 2);
  const size_t smem = 4 * (128 * 128 + (N / 2) * 128) + 2048;
  CK(cudaFuncSetAttribute(pair_gemm_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pair_gemm_kernel<N><<<pairs * 2, 128, smem>>>(amap, bmap, kblocks, 1, check ? dout : nullptr, dcyc, 1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N %d: kernel error %s\n", N, cudaGetErrorString(e)); exit(1); }
  if (check) {
    std::vector<float> got((size_t)M * N);
    CK(cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int k = 0; k < K; ++k) ref += fa[(size_t)m * K + k] * fb[(size_t)n * K + k];
        bad += fabsf(ref - got[(size_t)m * N + n]) > 0.5f;
      }
    printf("cta_group::2 GEMM M=%d N=%d K=%d: %ld / %ld wrong\n", M, N, K, bad, (long)M * N);
  } else {
    const int iters = 50;
    for (int feed = 1; feed >= 0; --feed) {
      pair_gemm_kernel<N><<<pairs * 2, 128, smem>>>(amap, bmap, kblocks, iters, nullptr, dcyc, feed);
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N %d: kernel error %s\n", N, cudaGetErrorString(e)); exit(1); }
      long long h;
      CK(cudaMemcpy(&h, dcyc, 8, cudaMemcpyDeviceToHost));
      printf("cta_group::2 rate, N=%3d, %s, %d pairs: %.1f cycles per M=256 MMA\n", N,
             feed ? "TMA-fed from L2 (A 16 KB + B half per CTA per 4 MMAs)" : "operands resident (no loads)", pairs,
             (double)h / ((double)iters * kblocks * 4));
    }
  }
  cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(dcyc);
}

int main() {
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  run<64>(2, 3, true);
  run<128>(2, 3, true);
  run<256>(3, 5, true);
  run<64>(74, 64, false);
  run<128>(74, 64, false);
  run<256>(74, 64, false);
  return 0;
}
