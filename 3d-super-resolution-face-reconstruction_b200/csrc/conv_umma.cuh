// Implicit-GEMM convolution on Blackwell tensor cores (tcgen05.mma, TMEM accumulators, TMA).
//
// Replaces every nn.Conv2d of the reference UNet (model/sr/sr3_modules/unet.py:62,71,87,102,
// 120-121) except the 6->64 head and the 64->3 tail, which are CUDA-core kernels.
//
// GEMM view:  D[M, N] = A[M, K] * W[N, K]^T
//   M = output pixels. One tile owns 128 of them, chosen as a (bb x bh x bw) box of the NHWC
//       activation so that ONE 4-D TMA box load, shifted by the filter tap, fetches the
//       A tile of that tap: rows = pixels, 64 channels (128 B) per row, 128B-swizzled - i.e.
//       exactly the K-major SWIZZLE_128B operand layout tcgen05 wants. Out-of-bounds rows/cols
//       (the conv's zero padding, or batch rows past B) are zero-filled by the TMA unit.
//   N = output channels, BLOCK_N per tile (64/128/256) = TMEM columns of the fp32 accumulator.
//   K = sum over "taps" of 64-channel blocks: the 9 (or 1) filter taps of the main source,
//       then optional 1x1 segments of other sources (the ResnetBlock's res_conv folded into
//       block2's GEMM, unet.py:103-110). Weights are pre-packed [Cout][K] in the same order.
// Stride-2 convs use four "parity" tensor maps over the same buffer (base offset + doubled
// strides) so a tap is again one plain box load. Upsample(nearest 2x)+conv3x3 (unet.py:58-65)
// is folded: each output parity (y&1, x&1) is a 2x2 conv over the LOW-resolution source with
// pre-summed weights, so the 4x tensor is never materialised (num_par = 4 tile groups).
//
// The kernel is PERSISTENT: grid = min(tiles, SMs); a CTA walks tiles blockIdx.x, +gridDim.x, ...
// Warp roles (256 threads): warp 0 = TMA producer (runs ahead across tiles through a STAGES-deep
// smem ring), warp 1 = MMA issuer (one lane; alternates between two TMEM accumulators),
// warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM -> regs -> +bias[t] (+residual) -> bf16
// stores, and per-channel sum / sum-of-squares of the fp32 outputs for the NEXT GroupNorm), which
// overlaps the MMAs of the following tile.
#pragma once
#include "common.cuh"

namespace b200sr3 {

constexpr int CONV_BLOCK_M = 128;
constexpr int CONV_BLOCK_K = 64;   // bf16 elements = one 128-byte swizzle row
constexpr int CONV_MAX_TAPS = 16;
constexpr int CONV_THREADS = 256;

struct ConvTap {
  int16_t map;      // which A tensor map
  int16_t dw, dh;   // box origin shift in the source's (W, H) index space
  int16_t cblocks;  // number of 64-channel K blocks this tap contributes
};

struct alignas(64) ConvParams {
  CUtensorMap a_map[4];
  CUtensorMap w_map;
  ConvTap taps[CONV_MAX_TAPS];     // num_par groups of num_taps entries
  int num_taps;                    // taps per group
  int num_kblocks;                 // K blocks per tile
  int num_par;                     // 1, or 4 when a nearest-2x upsample is folded in
  int tiles_w, tiles_h, tiles_b, tiles_n, total_tiles;
  int bw, bh, bb;                  // box (tile) extent in pixels, bw*bh*bb == 128
  int B, Hout, Wout, Cout;         // pixel space the tiles walk ([B,Hout,Wout]; source res if num_par==4)
  int out_H, out_W;                // dims of the tensor `out` points at ((2*Hout, 2*Wout) if num_par==4)
  const float* bias;               // [Cout] (may be null)
  int bias_t_stride;               // if non-zero, row ctl->t of a [rows][stride] table is used
  const StepCtl* ctl;
  const bf16* residual;            // same shape as out, added in the epilogue (may be null)
  bf16* out;
  // GroupNorm statistics of the output (null: not wanted): stat_partial[b][slot][c] = (sum, sumsq)
  // of the fp32 outputs over the pixels of image b that ONE CTA produced. A CTA owns a contiguous
  // run of tiles, so an image is covered by a handful of CTAs; CTA number j (in launch order) of
  // those writes slot j and the last one zero-fills the unused slots. No atomics, no fences: the
  // consumer (gn_apply_kernel) adds the stat_slots slots. Sums are carried as 2^-24 fixed point in
  // int64 (exactly associative), so the statistics do not depend on how tiles were split over
  // CTAs - i.e. not on the batch size either: a face's result is bit-identical in any batch.
  long long* stat_partial;
  int stat_slots;
  int stat_atomic;      // as ConvHaloParams::stat_atomic: one accumulator, added to with RED.64, zeroed by the engine
  int seg_len;                     // tiles per (n tile, image[-pair]) segment = tiles_w*tiles_h*num_par
  // optional role timing (B200SR3_CONV_TIMING=1 in b200sr3_conv2d): [grid][8] cycle counters
  unsigned long long* dbg;
  // ablation switches for tools/conv_bench.py (B200SR3_CONV_ABLATE bit mask; results are then wrong):
  // 1 = no global stores, 2 = no A loads, 4 = no W loads, 8 = no bias/residual loads, 16 = no TMEM loads
  int ablate;
};

#ifdef __CUDACC__
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint (ns): the hardware parks the thread until the phase completes or the time is up,
// so a waiting role does not spend issue slots. Measured on B200: with the hint-less form in a loop, the four epilogue
// warps and the single-thread roles waiting for their first work took enough issue slots from the (lower-numbered)
// transform warps to stretch a bar.sync to ~1 k cycles and the GroupNorm table build to ~4 k.
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// Non-blocking probe of a phase (used to look one pipeline stage ahead).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (-> cudaErrorLaunchFailure), never as a
// hung GPU. ~2 s at 2 GHz.
// The spinning part lives out of line: inlined, its clock arithmetic, printf and trap sat in the middle of every hot loop
// (~30 instructions per wait site, most painfully in the single-thread MMA issue loop).
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait_hint(bar, parity, 100000u)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("b200sr3: mbarrier timeout (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x,
             blockIdx.y, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
// One lane of a fully converged warp. Unlike `lane == 0`, the compiler knows exactly one thread is
// active under this predicate, so tcgen05 / TMA operands stay in uniform registers (no per-lane
// serialisation loop around every UTCHMMA / UTMALDG).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, single CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 64 bf16 (128 B),
// 8-row groups 1024 B apart (SBO); LBO is unused for swizzled K-major layouts (encoded as 1);
// bits [46,48) = descriptor version 1 (Blackwell); bits [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16: D fp32 (bits 4-5 = 1), A/B bf16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15,16 = 0), N>>3 at bit 17, M>>4 at bit 24.
__device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

}  // namespace ptx

template <int BLOCK_N, int STAGES>
struct ConvSmem {
  static constexpr int A_BYTES = CONV_BLOCK_M * CONV_BLOCK_K * 2;   // 16 KB
  static constexpr int B_BYTES = BLOCK_N * CONV_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAT_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int STAT_BYTES = 4 * 2 * BLOCK_N * 8;             // [warp][sum|sq][col] int64 fixed point
  static constexpr int BAR_OFFSET = STAT_OFFSET + STAT_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;            // + barriers + align slack
};

// lane l ends up with the sum over the warp's 32 lanes of v[l] (31 shuffles).
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float send = upper ? v[j] : v[j + s];
      const float keep = upper ? v[j + s] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// role timing helpers (two accumulators per thread; only when p.dbg is set)
#ifndef B200SR3_ROLE_TIMING
#define B200SR3_ROLE_TIMING 0      // compile-time option (build.py --timing): counters and ablation switches cost hot-loop code
#endif
#define UMMA_DBG (B200SR3_ROLE_TIMING && p.dbg != nullptr)
#define UMMA_ABLATE(bits) (B200SR3_ROLE_TIMING && (p.ablate & (bits)))
#define DBG_DECL() unsigned long long dbg_acc[2] = {0ull, 0ull}; long long dbg_t0 = 0
#define DBG_T0() do { if (UMMA_DBG) dbg_t0 = clock64(); } while (0)
#define DBG_ACC(i) do { if (UMMA_DBG) dbg_acc[i] += (unsigned long long)(clock64() - dbg_t0); } while (0)
#define DBG_FLUSH(slot, n) do { if (UMMA_DBG) for (int _i = 0; _i < (n); ++_i) p.dbg[blockIdx.x * 8 + (slot) + _i] = dbg_acc[_i]; } while (0)

// MMA_WARPS: a single warp can issue one tcgen05.mma per ~86 cycles (measured, tools/micro/
// umma_rate*.cu), which only saturates the tensor pipe at N = 256 (128 cycles each). For N <= 128 two
// warps issue concurrently, each into its own accumulators and from its own half of the smem ring
// (stage s belongs to warp s % MMA_WARPS): tiles alternate between the warps, the producer
// interleaves their K blocks. Measured: N = 128 reaches the full 4096 MAC/cycle/SM this way.
template <int BLOCK_N, int STAGES, int MMA_WARPS>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_umma_kernel(const __grid_constant__ ConvParams p) {
  static_assert(MMA_WARPS == 1 || MMA_WARPS == 2, "one or two MMA warps");
  static_assert(STAGES % MMA_WARPS == 0 && 2 * MMA_WARPS * BLOCK_N <= 512, "ring / TMEM budget");
  constexpr int NBUF = 2 * MMA_WARPS;       // TMEM accumulators (two per MMA warp)
  constexpr int RING = STAGES / MMA_WARPS;  // smem stages per MMA warp
  using S = ConvSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  long long* sstat = reinterpret_cast<long long*>(smem_gen + S::STAT_OFFSET);
  const uint32_t bar_base = smem_base + S::BAR_OFFSET;
  // barrier slots (8 B each): full[STAGES], empty[STAGES], tmem_full[NBUF], tmem_empty[NBUF]; then
  // the TMEM address.
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto tmem_empty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + NBUF + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * NBUF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.w_map);
    ptx::prefetch_tmap(&p.a_map[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      ptx::mbar_init(tmem_full_bar(b), 1);
      ptx::mbar_init(tmem_empty_bar(b), 128);
    }
    ptx::fence_barrier_init();
  }
  pdl_launch_dependents();
  if (warp == 3) ptx::tmem_alloc(tmem_slot, NBUF * BLOCK_N);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // A CTA owns the contiguous tile range [tile_begin, tile_end). Tile order: w fastest, then h,
  // parity group, image(-pair), n tile - so consecutive tiles belong to the same image and share
  // input halos in L2, and the GroupNorm partial sums of an image stay in one CTA for long runs.
  const int tile_begin = (int)(((long long)blockIdx.x * p.total_tiles) / gridDim.x);
  const int tile_end = (int)(((long long)(blockIdx.x + 1) * p.total_tiles) / gridDim.x);
  struct Tile { int n_tile, w0, h0, b0, par; };
  auto decode = [&](int tile) {
    Tile t;
    t.w0 = (tile % p.tiles_w) * p.bw; tile /= p.tiles_w;
    t.h0 = (tile % p.tiles_h) * p.bh; tile /= p.tiles_h;
    t.par = tile % p.num_par; tile /= p.num_par;
    t.b0 = (tile % p.tiles_b) * p.bb;
    t.n_tile = tile / p.tiles_b;
    return t;
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      // ---------------------------------------------------------------- TMA producer
      // Tiles are taken MMA_WARPS at a time; their K blocks are interleaved so that every MMA warp's
      // ring fills at the same pace. Ring j uses stages j, j + MMA_WARPS, ...
      int rs[MMA_WARPS];          // next ring-local stage
      uint32_t rphase[MMA_WARPS];
#pragma unroll
      for (int j = 0; j < MMA_WARPS; ++j) { rs[j] = 0; rphase[j] = 0; }
      DBG_DECL();
      for (int tile0 = tile_begin; tile0 < tile_end; tile0 += MMA_WARPS) {
        Tile t[MMA_WARPS];
#pragma unroll
        for (int j = 0; j < MMA_WARPS; ++j) t[j] = decode(min(tile0 + j, tile_end - 1));
        int kb = 0;
        for (int ti = 0; ti < p.num_taps; ++ti) {
          const int cblocks = p.taps[ti].cblocks;      // identical in every parity group
          for (int cb = 0; cb < cblocks; ++cb, ++kb) {
#pragma unroll
            for (int j = 0; j < MMA_WARPS; ++j) {
              if (tile0 + j >= tile_end) continue;
              const ConvTap tap = p.taps[t[j].par * p.num_taps + ti];
              const int stage = rs[j] * MMA_WARPS + j;
              DBG_T0();
              ptx::mbar_wait(empty_bar(stage), rphase[j] ^ 1u);
              DBG_ACC(0);
              const uint32_t sa = smem_base + stage * S::STAGE_BYTES;
              const uint32_t sb = sa + S::A_BYTES;
              ptx::mbar_expect_tx(full_bar(stage), (UMMA_ABLATE(2) ? 0 : S::A_BYTES) + (UMMA_ABLATE(4) ? 0 : S::B_BYTES));
              if (!UMMA_ABLATE(2))
                ptx::tma_load_4d(sa, &p.a_map[tap.map], full_bar(stage), cb * CONV_BLOCK_K, t[j].w0 + tap.dw,
                                 t[j].h0 + tap.dh, t[j].b0);
              if (!UMMA_ABLATE(4))
                ptx::tma_load_2d(sb, &p.w_map, full_bar(stage), kb * CONV_BLOCK_K,
                                 t[j].par * p.Cout + t[j].n_tile * BLOCK_N);
              if (++rs[j] == RING) { rs[j] = 0; rphase[j] ^= 1u; }
            }
          }
        }
      }
      DBG_FLUSH(0, 1);
    }
  } else if (warp <= MMA_WARPS) {
    {
      // ---------------------------------------------------------------- MMA issuer(s)
      // The whole warp walks the loop (all values are warp-uniform); one elected lane issues.
      const int mw = warp - 1;                  // this warp takes tiles tile_begin + mw, + MMA_WARPS, ...
      const uint32_t idesc = ptx::make_idesc_bf16(CONV_BLOCK_M, BLOCK_N);
      int lstage = 0;                           // ring-local stage
      uint32_t phase = 0;
      int it = mw;
      bool ready = false;
      DBG_DECL();
      for (int tile = tile_begin + mw; tile < tile_end; tile += MMA_WARPS, it += MMA_WARPS) {
        const int buf = it % NBUF;
        const uint32_t use = (uint32_t)(it / NBUF);
        DBG_T0();
        ptx::mbar_wait(tmem_empty_bar(buf), (use & 1u) ^ 1u);   // epilogue has drained this accumulator
        DBG_ACC(1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BLOCK_N);
        for (int kb = 0; kb < p.num_kblocks; ++kb) {
          const int stage = lstage * MMA_WARPS + mw;
          DBG_T0();
          if (!ready) ptx::mbar_wait(full_bar(stage), phase);
          DBG_ACC(0);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + stage * S::STAGE_BYTES;
          const uint32_t sb = sa + S::A_BYTES;
          const bool elected = ptx::elect_one();
          if (elected) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              // advancing 16 bf16 (32 B) along K inside the 128 B swizzle row: +2 in the address field
              const uint64_t da = ptx::make_sw128_desc(sa + k * 32);
              const uint64_t db = ptx::make_sw128_desc(sb + k * 32);
              ptx::umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          // Whatever this warp does between the last MMA of one stage and the first of the next is
          // exposed (issue is in order), so probe the NEXT stage's barrier now, while two MMAs are
          // in flight, and skip the blocking wait at the top of the loop when it has landed.
          const int nl = (lstage + 1 == RING) ? 0 : lstage + 1;
          const uint32_t nphase = (lstage + 1 == RING) ? (phase ^ 1u) : phase;
          const bool more = (kb + 1 < p.num_kblocks) || (tile + MMA_WARPS < tile_end);
          ready = more && __all_sync(0xffffffffu, ptx::mbar_test_wait(full_bar(nl * MMA_WARPS + mw), nphase));
          if (elected) {
#pragma unroll
            for (int k = 2; k < 4; ++k) {
              const uint64_t da = ptx::make_sw128_desc(sa + k * 32);
              const uint64_t db = ptx::make_sw128_desc(sb + k * 32);
              ptx::umma_bf16(d_tmem, da, db, idesc, 1u);
            }
            ptx::umma_commit(empty_bar(stage));   // frees the smem slot when these MMAs retire
            if (kb == p.num_kblocks - 1) ptx::umma_commit(tmem_full_bar(buf));   // accumulator complete
          }
          __syncwarp();
          lstage = nl;
          phase = nphase;
        }
      }
      if (lane == 0 && mw == 0) DBG_FLUSH(2, 2);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int wq = warp & 3;                  // TMEM lane quarter this warp may read
    const int row = wq * 32 + lane;           // GEMM row = pixel inside the box
    const int tid_e = threadIdx.x - 128;
    const int lw = row % p.bw;
    const int lh = (row / p.bw) % p.bh;
    const int lb = row / (p.bw * p.bh);
    const bool do_stats = p.stat_partial != nullptr;
    const float* bias = p.bias;
    if (bias && p.bias_t_stride) bias += (size_t)p.ctl->t * p.bias_t_stride;
    const int osc = p.num_par == 4 ? 2 : 1;
    // running per-channel (sum, sumsq) of this warp's rows: sacc[warp][sum|sq][col], lane-private
    long long* sacc = sstat + wq * (2 * BLOCK_N);
    if (do_stats)
      for (int i = lane; i < 2 * BLOCK_N; i += 32) sacc[i] = 0;

    int it = 0;
    DBG_DECL();
    const long long dbg_start = UMMA_DBG ? clock64() : 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const Tile t = decode(tile);
      const int buf = it % NBUF;
      const uint32_t use = (uint32_t)(it / NBUF);
      const int n0 = t.n_tile * BLOCK_N;
      const int b = t.b0 + lb, h = t.h0 + lh, w = t.w0 + lw;
      const bool valid = (b < p.B) && (h < p.Hout) && (w < p.Wout);
      const size_t pix = ((size_t)b * p.out_H + (size_t)(h * osc + (t.par >> 1))) * p.out_W +
                         (size_t)(w * osc + (t.par & 1));
      bf16* out_row = p.out + pix * p.Cout + n0;
      const bf16* res_row = p.residual ? p.residual + pix * p.Cout + n0 : nullptr;

      DBG_T0();
      ptx::mbar_wait(tmem_full_bar(buf), use & 1u);
      DBG_ACC(0);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * BLOCK_N);
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        if (!UMMA_ABLATE(16)) {
          ptx::tmem_ld32(taddr + (uint32_t)c0, v);
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (uint32_t)(c0 + j);
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (bias && !UMMA_ABLATE(8)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n0 + c0 + j));
            f[j] += bv.x; f[j + 1] += bv.y; f[j + 2] += bv.z; f[j + 3] += bv.w;
          }
        }
        if (valid) {
          if (res_row && !UMMA_ABLATE(8)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float r[8];
              unpack8(*reinterpret_cast<const uint4*>(res_row + c0 + j), r);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[j + e] += r[e];
            }
          }
          if (!UMMA_ABLATE(1)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) *reinterpret_cast<uint4*>(out_row + c0 + j) = pack8(f + j);
          }
        }
        if (do_stats) {
          float q[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            f[j] = valid ? f[j] : 0.f;
            q[j] = f[j] * f[j];
          }
          const float s_sum = warp_transpose_sum(f, lane);
          const float s_sq = warp_transpose_sum(q, lane);
          sacc[c0 + lane] += __float2ll_rn(s_sum * STAT_FIXED_SCALE);
          sacc[BLOCK_N + c0 + lane] += __float2ll_rn(s_sq * STAT_FIXED_SCALE);
        }
      }
      // all of this thread's TMEM reads have completed: hand the accumulator back to the MMA warp
      ptx::tc_fence_before();
      ptx::mbar_arrive(tmem_empty_bar(buf));

      if (do_stats) {
        const int seg = tile / p.seg_len;
        if (tile + 1 == tile_end || (tile + 1) / p.seg_len != seg) {
          // ---- the CTA's run over this (n tile, image[-pair]) segment ends: publish its partial sums
          const long long G = gridDim.x, T = p.total_tiles;
          const int first_cta = (int)((((long long)seg * p.seg_len + 1) * G - 1) / T);
          const int last_cta = (int)((((long long)(seg + 1) * p.seg_len) * G - 1) / T);
          const int slot = (int)blockIdx.x - first_cta;
          const int nimg = p.bb >= 2 ? 2 : 1;       // host guarantees bb <= 2 when stats are fused
          const int wpi = 4 / nimg;
          epi_bar_sync();
          for (int item = tid_e; item < nimg * 2 * BLOCK_N; item += 128) {
            const int col = item % BLOCK_N;
            const int st = (item / BLOCK_N) & 1;
            const int ib = item / (2 * BLOCK_N);
            long long a = 0;
            for (int ww = 0; ww < wpi; ++ww) a += sstat[((ib * wpi + ww) * 2 + st) * BLOCK_N + col];
            const int bi = t.b0 + ib;
            if (bi < p.B && p.stat_atomic) {
              atomicAdd(reinterpret_cast<unsigned long long*>(p.stat_partial + ((size_t)bi * p.Cout + (n0 + col)) * 2 + st),
                        (unsigned long long)a);
            } else if (bi < p.B) {
              long long* dst = p.stat_partial + (((size_t)bi * p.stat_slots + slot) * p.Cout + (n0 + col)) * 2 + st;
              *dst = a;
              if ((int)blockIdx.x == last_cta)
                for (int sl = slot + 1; sl < p.stat_slots; ++sl) dst[(size_t)(sl - slot) * p.Cout * 2] = 0;
            }
          }
          epi_bar_sync();
          for (int i = lane; i < 2 * BLOCK_N; i += 32) sacc[i] = 0;
        }
      }
    }
    if (UMMA_DBG && tid_e == 0) {
      p.dbg[blockIdx.x * 8 + 4] = dbg_acc[0];                      // epilogue: waiting for an accumulator
      p.dbg[blockIdx.x * 8 + 5] = (unsigned long long)(clock64() - dbg_start);   // epilogue: total
      p.dbg[blockIdx.x * 8 + 6] = (unsigned long long)(tile_end - tile_begin);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 3) ptx::tmem_dealloc(tmem_base, NBUF * BLOCK_N);
}
#endif  // __CUDACC__

}  // namespace b200sr3
