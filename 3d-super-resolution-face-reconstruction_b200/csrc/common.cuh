// Shared declarations for the b200sr3 CUDA engine (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#ifndef __CUDA_ARCH__
#define B200_HOST 1
#endif

namespace b200sr3 {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------- errors
struct Error : std::runtime_error {
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};

#define CUDA_CHECK(expr)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      throw ::b200sr3::Error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) +      \
                             " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")");         \
  } while (0)

#define REQUIRE(cond, msg)                                                                    \
  do {                                                                                        \
    if (!(cond)) throw ::b200sr3::Error(std::string(msg) + " [" #cond "]");                   \
  } while (0)

// ------------------------------------------------------------------------------- step control
// Lives in device memory; every kernel of a captured step reads it, so one CUDA graph serves
// all T steps. Written by the host before a chain, advanced on the device after each step.
struct StepCtl {
  int t;               // current timestep (row of the schedule / noise-bias tables)
  int T;               // schedule length; row T of the bias table is the scratch row
  int noise_mode;      // 0: zeros, 1: injected list, 2: philox, 3: direct pointer (one step)
  int no_clip;         // 1: p_mean_variance(clip_denoised=False), x0 is not clamped (diffusion.py:175-176); 0 = the default
  const float* noise;  // list base (mode 1) or this step's z (mode 3)
  unsigned long long seed;
  long long numel;     // B*3*R*R, stride between entries of the injected list
  long long row0;      // global index of this call's batch row 0: the Philox stream is keyed by GLOBAL (row, y, x), so a
                       // face's noise does not depend on how a batch is sharded over ranks or chunked over calls
};

// ------------------------------------------------------------------------------- launches
// Every kernel of the per-step chain can be launched with programmatic dependent launch (PDL,
// B200SR3_PDL=1): kernel N+1 may be scheduled while kernel N drains, runs its prologue (barrier init,
// TMEM allocation, tensor-map prefetch) and blocks in pdl_wait() until N has completed and flushed.
// Each of those kernels calls pdl_launch_dependents() and, before it touches global memory,
// pdl_wait(); without the launch attribute both are no-ops. Off by default: see pdl_enabled().
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

// ------------------------------------------------------------------------------- small helpers
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// GroupNorm partial sums travel as int64 fixed point: integer addition is exactly associative, so
// the statistics of an image do not depend on which CTA summed which tile.
#define STAT_FIXED_SCALE 16777216.0f
#define STAT_FIXED_INV (1.0 / 16777216.0)

struct Act {  // NHWC bf16 activation
  bf16* ptr = nullptr;
  long long* stats = nullptr;   // [B][stat_slots][C][2]: per-(image, channel) partial sums and sums of squares,
  int stat_slots = 1;           // 2^-24 fixed point (GroupNorm input statistics; the consumer adds the slots)
  int B = 0, H = 0, W = 0, C = 0;
  size_t elems() const { return (size_t)B * H * W * C; }
};

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// x * sigmoid(x) with one MUFU.EX2 and one MUFU.RCP (both ~2^-22 relative: far below the bf16
// rounding of the stored result); the IEEE division this replaces cost ~15 instructions.
__device__ __forceinline__ float swish_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  // a bf16 is the top half of an fp32: shift / mask, two instructions per pair (__bfloat1622float2 compiles to three)
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return q;
}
// ---- Philox4x32-10 + Box-Muller: the sampler's own noise stream (keyed by pixel, timestep, seed)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
  const float u2 = (float)b * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * __logf(u1));
  float sn, cs;
  __sincosf(6.283185307179586f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}
// The reference's update, op for op (diffusion.py:150-151, 175-176, 159-160, 186-187).
// `lim` = 1 (clip_denoised=True, the only value the reference's sampler uses) or +inf (clip_denoised=False).
__device__ __forceinline__ float posterior_update(float x, float eps, float z, float a, float bc,
                                                  float c1, float c2, float sigma, float lim = 1.0f) {
  float x0 = __fsub_rn(__fmul_rn(a, x), __fmul_rn(bc, eps));
  x0 = fminf(fmaxf(x0, -lim), lim);
  const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, x));
  return __fadd_rn(mean, __fmul_rn(z, sigma));
}
#endif

}  // namespace b200sr3
