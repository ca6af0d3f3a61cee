// Non-GEMM kernels of the SR3 sampling step: GroupNorm(+Swish), head conv, tail conv fused with
// the posterior update, mid-block attention core, nearest upsample, weight packing and the
// noise-embedding bias table. HBM-bound kernels use 16-byte vector accesses on NHWC bf16.
#include <mma.h>

#include <algorithm>

#include "kernels.cuh"

namespace b200sr3 {

// Measured on B200 (bench.py, T=600, B=32): 3.60 ms per sampling step plain, 3.63 ms with PDL edges in the
// graph - graph replay already starts a node within ~1 us of its predecessor - so PDL is opt-in.
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("B200SR3_PDL"); return e && e[0] == '1'; }();
  return on;
}

// =============================================================================== GroupNorm
// Statistics arrive as per-(image, channel) (sum, sumsq) pairs "chansum[b][c][2]", produced in the
// epilogue of the conv that wrote the tensor (conv_umma.cuh) or, for tensors no tensor-core conv
// can cover (the head conv, spatial sizes below 8x8), by chan_stats_kernel below. Groups may
// straddle the seam of a channel concat (unet.py:261), which is why the sums are kept per channel
// and grouped only here.
//
// Both kernels run 384-thread CTAs: 384 is divisible by every channel-vector count C/8 the UNet
// produces (8,16,24,32,48,64,96,128), so a thread keeps ONE 8-channel column for its whole life —
// no index division in the loop, scale/shift live in registers, and each thread keeps four
// independent 16-byte loads in flight.
constexpr int GN_THREADS = 384;

// Per-CTA partial (sum, sumsq) per channel; the last CTA of an image (ticket) reduces the
// partials in a fixed order (deterministic) into chansum.
__global__ void __launch_bounds__(GN_THREADS) chan_stats_kernel(ChanStatsPlan g) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[GN_THREADS * 16];
  __shared__ int is_last;
  const int C = g.C;
  const int Cv = C >> 3;
  const int rows = GN_THREADS / Cv;
  const int tid = threadIdx.x;
  const int cv = tid % Cv, r = tid / Cv;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int ppc = (g.HW + g.chunks - 1) / g.chunks;
  const int p_begin = chunk * ppc;
  const int p_end = min(g.HW, p_begin + ppc);

  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  {
    const bf16* base = g.src + (size_t)b * g.HW * C + cv * 8;
    int p = p_begin + r;
    for (; p + 3 * rows < p_end; p += 4 * rows) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(p + u * rows) * C));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
      }
    }
    for (; p < p_end; p += rows) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)p * C)), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[tid * 16 + i] = s[i]; red[tid * 16 + 8 + i] = q[i]; }
  __syncthreads();
  // reduce over rows: thread (j < Cv*16) owns one (channel-vector, slot) pair
  for (int j = tid; j < Cv * 16; j += GN_THREADS) {
    const int v = j >> 4, slot = j & 15;
    float a = 0.f;
    for (int rr = 0; rr < rows; ++rr) a += red[(rr * Cv + v) * 16 + slot];
    // layout [c][2] with slot<8 -> sum of channel v*8+slot, else sumsq
    const size_t o = (size_t)(v * 8 + (slot & 7)) * 2 + (slot >> 3);
    if (g.chunks == 1) g.chansum[(size_t)b * C * 2 + o] = __float2ll_rn(a * STAT_FIXED_SCALE);
    else g.partial[(size_t)(b * g.chunks + chunk) * C * 2 + o] = a;
  }
  if (g.chunks == 1) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(&g.ticket[b], 1) == g.chunks - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int j = tid; j < C * 2; j += GN_THREADS) {
    float a = 0.f;
    for (int k = 0; k < g.chunks; ++k) a += __ldcg(g.partial + (size_t)(b * g.chunks + k) * C * 2 + j);
    g.chansum[(size_t)b * C * 2 + j] = __float2ll_rn(a * STAT_FIXED_SCALE);
  }
  if (tid == 0) g.ticket[b] = 0;
}

// y = [swish]((x - mean_g) * rstd_g * gamma + beta) over [src0 | src1]; grid (apply_chunks, B):
// a CTA forms the group statistics of its image from chansum, then streams a pixel range.
__global__ void __launch_bounds__(GN_THREADS, 3) gn_apply_kernel(GnPlan g, int apply_chunks) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float chan[1024 * 2];
  __shared__ float gstat[64 * 2];
  const int C = g.C0 + g.C1;
  const int Cv = C >> 3;
  const int rows = GN_THREADS / Cv;
  const int tid = threadIdx.x;
  const int cv = tid % Cv, r = tid / Cv;
  const int b = blockIdx.y;
  for (int c = tid; c < C; c += GN_THREADS) {
    // add the producers' partial-sum slots in a fixed order
    const long long* sp;
    int ns, cs;
    if (c < g.C0) { sp = g.stats0 + ((size_t)b * g.slots0 * g.C0 + c) * 2; ns = g.slots0; cs = g.C0; }
    else { sp = g.stats1 + ((size_t)b * g.slots1 * g.C1 + (c - g.C0)) * 2; ns = g.slots1; cs = g.C1; }
    long long a = 0, d = 0;
    for (int k = 0; k < ns; ++k) {
      const longlong2 v = *reinterpret_cast<const longlong2*>(sp + (size_t)k * cs * 2);
      a += v.x;
      d += v.y;
    }
    chan[2 * c] = (float)((double)a * STAT_FIXED_INV);
    chan[2 * c + 1] = (float)((double)d * STAT_FIXED_INV);
  }
  __syncthreads();
  const int cg = C / g.groups;
  if (tid < g.groups) {
    float a = 0.f, d = 0.f;
    for (int k = 0; k < cg; ++k) { a += chan[2 * (tid * cg + k)]; d += chan[2 * (tid * cg + k) + 1]; }
    const float inv_n = 1.0f / ((float)g.HW * (float)cg);
    const float mean = a * inv_n;
    const float var = fmaxf(d * inv_n - mean * mean, 0.f);
    gstat[2 * tid] = mean;
    gstat[2 * tid + 1] = rsqrtf(var + 1e-5f);
  }
  __syncthreads();
  const int ppc = (g.HW + apply_chunks - 1) / apply_chunks;
  const int p_begin = blockIdx.x * ppc;
  const int p_end = min(g.HW, p_begin + ppc);
  const int c = cv * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int grp = (c + i) / cg;
    sc[i] = gstat[2 * grp + 1] * __ldg(g.gamma + c + i);
    sh[i] = __ldg(g.beta + c + i) - gstat[2 * grp] * sc[i];
  }
  const bf16* base;
  int cs;
  if (c < g.C0) { base = g.src0 + c; cs = g.C0; } else { base = g.src1 + (c - g.C0); cs = g.C1; }
  base += (size_t)b * g.HW * cs;
  bf16* dst = g.dst + (size_t)b * g.HW * C + c;
  const bool do_swish = g.swish != 0;
  auto emit = [&](const uint4& v, int p) {
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = fmaf(f[i], sc[i], sh[i]);
      f[i] = do_swish ? swish_f(y) : y;
    }
    *reinterpret_cast<uint4*>(dst + (size_t)p * C) = pack8(f);
  };
  int p = p_begin + r;
  for (; p + 3 * rows < p_end; p += 4 * rows) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(p + u * rows) * cs));
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(v[u], p + u * rows);
  }
  for (; p < p_end; p += rows) emit(__ldg(reinterpret_cast<const uint4*>(base + (size_t)p * cs)), p);
}

// Same statistics as gn_apply_kernel's preamble, for the fused path: one CTA per image. The kernel is pure
// latency (it sits between two convs 57 times per step), so every (slot, channel) partial is fetched by its
// own thread - one L2 round trip instead of `slots` dependent ones - and summed with shared-memory int64
// atomics (integer addition: the order does not matter, the result is exact and deterministic).
__global__ void __launch_bounds__(512) gn_scale_shift_kernel(GnPlan g, float2* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ unsigned long long isum[1024 * 2];
  __shared__ float gstat[64 * 2];
  const int C = g.C0 + g.C1;
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  for (int i = tid; i < 2 * C; i += 512) isum[i] = 0ull;
  __syncthreads();
  {
    const int n0 = g.slots0 * g.C0, n1 = g.slots1 * g.C1;
    const long long* s0 = g.stats0 + (size_t)b * n0 * 2;
    const long long* s1 = g.C1 ? g.stats1 + (size_t)b * n1 * 2 : nullptr;
    for (int i = tid; i < n0 + n1; i += 512) {
      const bool second = i >= n0;
      const int k = second ? i - n0 : i;
      const int c = second ? g.C0 + k % g.C1 : k % g.C0;
      const longlong2 v = *reinterpret_cast<const longlong2*>((second ? s1 : s0) + (size_t)k * 2);
      atomicAdd(&isum[2 * c], (unsigned long long)v.x);
      atomicAdd(&isum[2 * c + 1], (unsigned long long)v.y);
    }
  }
  __syncthreads();
  const int cg = C / g.groups;
  if (tid < g.groups) {
    float a = 0.f, d = 0.f;
    for (int k = 0; k < cg; ++k) {
      a += (float)((double)(long long)isum[2 * (tid * cg + k)] * STAT_FIXED_INV);
      d += (float)((double)(long long)isum[2 * (tid * cg + k) + 1] * STAT_FIXED_INV);
    }
    const float inv_n = 1.0f / ((float)g.HW * (float)cg);
    const float mean = a * inv_n;
    const float var = fmaxf(d * inv_n - mean * mean, 0.f);
    gstat[2 * tid] = mean;
    gstat[2 * tid + 1] = rsqrtf(var + 1e-5f);
  }
  __syncthreads();
  for (int c = tid; c < C; c += 512) {
    const int grp = c / cg;
    const float sc = gstat[2 * grp + 1] * __ldg(g.gamma + c);
    out[(size_t)b * C + c] = make_float2(sc, __ldg(g.beta + c) - gstat[2 * grp] * sc);
  }
}

void launch_gn_scale_shift(const GnPlan& g, float2* out, cudaStream_t s) {
  const int C = g.C0 + g.C1;
  REQUIRE(C <= 1024 && C % g.groups == 0 && g.groups <= 64, "GroupNorm: unsupported channel count");
  REQUIRE(g.stats0 && (g.C1 == 0 || g.stats1), "GroupNorm: missing channel statistics");
  launch_pdl(gn_scale_shift_kernel, dim3(g.B), dim3(512), 0, s, g, out);
  CUDA_CHECK(cudaGetLastError());
}

int chan_stats_chunks(int HW, int C) {
  long long per_img = (long long)HW * C;
  long long ch = per_img / ((long long)GN_THREADS * 8 * 16);   // ~16 vectors per thread
  if (ch < 1) ch = 1;
  if (ch > 64) ch = 64;
  if (ch > HW) ch = HW;
  return (int)ch;
}

void launch_chan_stats(const ChanStatsPlan& g, cudaStream_t s) {
  REQUIRE(g.C % 8 == 0 && g.C <= 1024 && GN_THREADS % (g.C / 8) == 0, "chan_stats: unsupported channel count");
  launch_pdl(chan_stats_kernel, dim3(g.chunks, g.B), dim3(GN_THREADS), 0, s, g);
  CUDA_CHECK(cudaGetLastError());
}

void launch_gn_apply(const GnPlan& g, cudaStream_t s) {
  const int C = g.C0 + g.C1;
  REQUIRE(C % 8 == 0 && g.C0 % 8 == 0 && C <= 1024 && C % g.groups == 0 && g.groups <= 64,
          "GroupNorm: unsupported channel count");
  REQUIRE(GN_THREADS % (C / 8) == 0, "GroupNorm: C/8 must divide 384");
  REQUIRE(g.stats0 && (g.C1 == 0 || g.stats1), "GroupNorm: missing channel statistics");
  long long ch = (long long)g.HW * C / ((long long)GN_THREADS * 8 * 16);     // ~16 vectors per thread
  if (ch < 1) ch = 1;
  if (ch > g.HW) ch = g.HW;
  if (ch > 1024) ch = 1024;
  launch_pdl(gn_apply_kernel, dim3((int)ch, g.B), dim3(GN_THREADS), 0, s, g, (int)ch);
  CUDA_CHECK(cudaGetLastError());
}

// =============================================================================== head conv
// 3x3, pad 1, Cin = c_cond + c_x (6) read straight from the fp32 NCHW sampler state, so x_t is
// never rounded to bf16 on the way in. K = 54: CUDA cores.
__global__ void __launch_bounds__(256) head_conv_kernel(const float* __restrict__ cond,
                                                        const float* __restrict__ x, int c_cond, int c_x,
                                                        const float* __restrict__ w_kc,
                                                        const float* __restrict__ bias, int B, int R,
                                                        int Cout, bf16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float w_s[];   // [K][Cout]
  const int Cin = c_cond + c_x;
  const int K = Cin * 9;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) w_s[i] = w_kc[i];
  __syncthreads();
  const int tpp = Cout >> 3;                     // threads per pixel
  const int ppb = blockDim.x / tpp;              // pixels per block
  const int cgp = threadIdx.x % tpp;
  const long long pix = (long long)blockIdx.x * ppb + threadIdx.x / tpp;
  const long long npix = (long long)B * R * R;
  if (pix >= npix || threadIdx.x / tpp >= ppb) return;
  const int xw = (int)(pix % R);
  const int yh = (int)((pix / R) % R);
  const int b = (int)(pix / ((long long)R * R));
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = bias[cgp * 8 + i];
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = yh + tap / 3 - 1, xx = xw + tap % 3 - 1;
    if (yy < 0 || yy >= R || xx < 0 || xx >= R) continue;
    for (int ci = 0; ci < Cin; ++ci) {
      const float v = (ci < c_cond) ? __ldg(cond + (((size_t)b * c_cond + ci) * R + yy) * R + xx)
                                    : __ldg(x + (((size_t)b * c_x + (ci - c_cond)) * R + yy) * R + xx);
      const float4* wr = reinterpret_cast<const float4*>(w_s + (tap * Cin + ci) * Cout + cgp * 8);
      const float4 w0 = wr[0], w1 = wr[1];
      acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
      acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
      acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
      acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
    }
  }
  *reinterpret_cast<uint4*>(out + pix * Cout + cgp * 8) = pack8(acc);
}

// Fast path (Cout = 64, R % 32 == 0): a CTA owns an 8 x 32 pixel tile, one pixel per thread with all
// 64 output channels in registers; the input tile (+halo) and the [K][64] weights sit in shared
// memory (weights are read as warp-wide broadcasts). A warp is one tile row, so the per-channel
// (sum, sumsq) of the fp32 outputs reduce with the same shuffle transpose the tcgen05 epilogue
// uses and leave the CTA as 128 int64 fixed-point atomics (exactly associative -> deterministic).
constexpr int HEAD_TW = 32, HEAD_TH = 8;

__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
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

__global__ void __launch_bounds__(HEAD_TW * HEAD_TH, 2)
head_conv64_kernel(const float* __restrict__ cond, const float* __restrict__ x, int c_cond, int c_x,
                   const float* __restrict__ w_kc, const float* __restrict__ bias, int B, int R,
                   bf16* __restrict__ out, unsigned long long* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  // Persistent: a CTA loads the [K][64] weights once and walks a contiguous run of tiles, so the
  // statistics of an image leave the CTA as ONE set of 128 atomics per (CTA, image) instead of one per tile.
  constexpr int COUT = 64;
  extern __shared__ float hs[];
  const int Cin = c_cond + c_x;
  const int K = Cin * 9;
  float* w_s = hs;                                  // [K][64]
  float* in_s = hs + K * COUT;                      // [Cin][TH+2][TW+2]
  __shared__ unsigned long long sred[2 * COUT];
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  const int tiles_x = R / HEAD_TW, tiles_y = R / HEAD_TH;
  const int total = B * tiles_x * tiles_y;
  const int t_begin = (int)(((long long)blockIdx.x * total) / gridDim.x);
  const int t_end = (int)(((long long)(blockIdx.x + 1) * total) / gridDim.x);

  for (int i = tid; i < K * COUT / 4; i += blockDim.x)
    reinterpret_cast<float4*>(w_s)[i] = __ldg(reinterpret_cast<const float4*>(w_kc) + i);
  if (tid < 2 * COUT) sred[tid] = 0ull;
  constexpr int IW = HEAD_TW + 2, IH = HEAD_TH + 2;
  int cur_b = -1;
  for (int tile = t_begin; tile < t_end; ++tile) {
    const int x0 = (tile % tiles_x) * HEAD_TW;
    const int y0 = ((tile / tiles_x) % tiles_y) * HEAD_TH;
    const int b = tile / (tiles_x * tiles_y);
    __syncthreads();                                // previous tile's reads of in_s (and sred flush) are done
    if (stats && b != cur_b) {
      if (cur_b >= 0 && tid < 2 * COUT) {
        atomicAdd(stats + (size_t)cur_b * COUT * 2 + tid, sred[tid]);
        sred[tid] = 0ull;
      }
      cur_b = b;
    }
    for (int i = tid; i < Cin * IH * IW; i += blockDim.x) {
      const int ix = i % IW, iy = (i / IW) % IH, ci = i / (IW * IH);
      const int gx = x0 + ix - 1, gy = y0 + iy - 1;
      float v = 0.f;
      if (gx >= 0 && gx < R && gy >= 0 && gy < R)
        v = (ci < c_cond) ? __ldg(cond + (((size_t)b * c_cond + ci) * R + gy) * R + gx)
                          : __ldg(x + (((size_t)b * c_x + (ci - c_cond)) * R + gy) * R + gx);
      in_s[i] = v;
    }
    __syncthreads();

    float acc[COUT];
#pragma unroll
    for (int j = 0; j < COUT; j += 4) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + j));
      acc[j] = bv.x; acc[j + 1] = bv.y; acc[j + 2] = bv.z; acc[j + 3] = bv.w;
    }
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
      for (int ci = 0; ci < Cin; ++ci) {
        const float v = in_s[(ci * IH + ty + ky) * IW + tx + kx];
        const float4* wr = reinterpret_cast<const float4*>(w_s + (tap * Cin + ci) * COUT);
#pragma unroll
        for (int j = 0; j < COUT / 4; ++j) {
          const float4 w = wr[j];
          acc[4 * j] = fmaf(v, w.x, acc[4 * j]);
          acc[4 * j + 1] = fmaf(v, w.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(v, w.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(v, w.w, acc[4 * j + 3]);
        }
      }
    }
    bf16* dst = out + (((size_t)b * R + y0 + ty) * R + x0 + tx) * COUT;
#pragma unroll
    for (int j = 0; j < COUT; j += 8) *reinterpret_cast<uint4*>(dst + j) = pack8(acc + j);

    if (stats) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float f[32], q[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { f[j] = acc[half * 32 + j]; q[j] = f[j] * f[j]; }
        const float s_sum = warp_transpose_sum32(f, tx);
        const float s_sq = warp_transpose_sum32(q, tx);
        atomicAdd(&sred[(half * 32 + tx) * 2], (unsigned long long)__float2ll_rn(s_sum * STAT_FIXED_SCALE));
        atomicAdd(&sred[(half * 32 + tx) * 2 + 1], (unsigned long long)__float2ll_rn(s_sq * STAT_FIXED_SCALE));
      }
    }
  }
  __syncthreads();
  if (stats && cur_b >= 0 && tid < 2 * COUT) atomicAdd(stats + (size_t)cur_b * COUT * 2 + tid, sred[tid]);
}

// ---- head conv on the tensor cores: operand preparation
// The 6 -> 64 head conv must not round x_t to bf16 (the sampler state is fp32 all the way), so the
// tensor-core version feeds a SPLIT operand: channels [0,n) hold hi = bf16(v), [n,2n) hold lo = bf16(v - hi)
// and [2n,3n) hold hi again; the packed weights carry (w_hi, w_hi, w_lo) in those positions, so the GEMM
// forms w_hi*v_hi + w_hi*v_lo + w_lo*v_hi = w*v up to 2^-16 relative, accumulated in fp32.
template <int N>
__global__ void __launch_bounds__(256) head_pack_kernel(const float* __restrict__ cond, const float* __restrict__ x,
                                                        int c_cond, int B, int R, bf16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long npix = (long long)B * R * R;
  const size_t plane = (size_t)R * R;
  const int c_x = N - c_cond;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(pix / (long long)plane);
    const size_t off = (size_t)(pix - (long long)b * (long long)plane);
    float row[32];                                  // 3N <= 24 used values, fully unrolled -> registers
#pragma unroll
    for (int i = 0; i < 32; ++i) row[i] = 0.f;
#pragma unroll
    for (int c = 0; c < N; ++c) {
      const float v = (c < c_cond) ? __ldg(cond + ((size_t)b * c_cond + c) * plane + off)
                                   : __ldg(x + ((size_t)b * c_x + (c - c_cond)) * plane + off);
      const float hi = __bfloat162float(__float2bfloat16_rn(v));
      row[c] = hi;
      row[N + c] = v - hi;
      row[2 * N + c] = hi;
    }
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)pix * 64);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = pack8(row + 8 * i);
#pragma unroll
    for (int i = 4; i < 8; ++i) dst[i] = make_uint4(0u, 0u, 0u, 0u);
  }
}
void launch_head_pack(const float* cond, const float* x, int c_cond, int c_x, int B, int R, bf16* out, cudaStream_t s) {
  const int n = c_cond + c_x;
  REQUIRE(n == 2 || n == 6 || n == 8, "head pack: in_channel must be 2, 6 or 8");
  const long long npix = (long long)B * R * R;
  long long blocks = (npix + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (n == 6) launch_pdl(head_pack_kernel<6>, dim3((unsigned)blocks), dim3(256), 0, s, cond, x, c_cond, B, R, out);
  else if (n == 2) launch_pdl(head_pack_kernel<2>, dim3((unsigned)blocks), dim3(256), 0, s, cond, x, c_cond, B, R, out);
  else launch_pdl(head_pack_kernel<8>, dim3((unsigned)blocks), dim3(256), 0, s, cond, x, c_cond, B, R, out);
}
// OIHW fp32 [Cout][n][3][3] -> [Cout][9*64] bf16 in the split layout above
__global__ void pack_head_split_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int Cout, int n) {
  const int total = Cout * 9 * 64;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = idx % 64, tap = (idx / 64) % 9, o = idx / (64 * 9);
    float v = 0.f;
    if (c < 3 * n) {
      const float w = src[((size_t)o * n + (c % n)) * 9 + tap];
      const float hi = __bfloat162float(__float2bfloat16_rn(w));
      v = (c < 2 * n) ? hi : (w - hi);
    }
    dst[idx] = __float2bfloat16_rn(v);
  }
}
void launch_pack_head_split_weight(const float* src, bf16* dst, int Cout, int n, cudaStream_t s) {
  pack_head_split_weight_kernel<<<(Cout * 9 * 64 + 255) / 256, 256, 0, s>>>(src, dst, Cout, n);
  CUDA_CHECK(cudaGetLastError());
}

bool head_conv_fast_path(int c_in, int R, int Cout) {
  return Cout == 64 && R % HEAD_TW == 0 && R % HEAD_TH == 0 &&
         (size_t)(c_in * 9 * 64 + c_in * (HEAD_TH + 2) * (HEAD_TW + 2)) * sizeof(float) <= 48 * 1024;
}

void launch_head_conv(const float* cond, const float* x, int c_cond, int c_x, const float* w_kc,
                      const float* bias, int B, int R, int Cout, bf16* out, long long* stats, cudaStream_t s) {
  if (head_conv_fast_path(c_cond + c_x, R, Cout)) {
    const int cin = c_cond + c_x;
    const size_t smem = (size_t)(cin * 9 * 64 + cin * (HEAD_TH + 2) * (HEAD_TW + 2)) * sizeof(float);
    if (stats) CUDA_CHECK(cudaMemsetAsync(stats, 0, (size_t)B * 64 * 2 * sizeof(long long), s));
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned tiles = (unsigned)(B * (R / HEAD_TH) * (R / HEAD_TW));
    const unsigned blocks = std::min<unsigned>(tiles, (unsigned)(2 * sms));
    launch_pdl(head_conv64_kernel, dim3(blocks), dim3(HEAD_TW * HEAD_TH), smem, s, cond, x, c_cond, c_x, w_kc, bias, B, R,
               out, reinterpret_cast<unsigned long long*>(stats));
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  REQUIRE(stats == nullptr, "head conv: the generic path does not produce statistics");
  REQUIRE(Cout % 8 == 0 && Cout <= 2048, "head conv: Cout must be a multiple of 8");
  const int tpp = Cout / 8;
  const int threads = 256 / tpp * tpp > 0 ? (256 / tpp) * tpp : tpp;
  const int ppb = threads / tpp;
  const long long npix = (long long)B * R * R;
  const size_t smem = (size_t)(c_cond + c_x) * 9 * Cout * sizeof(float);
  REQUIRE(smem <= 48 * 1024, "head conv: weights exceed 48 KB of shared memory");
  launch_pdl(head_conv_kernel, dim3((unsigned)((npix + ppb - 1) / ppb)), dim3(threads), smem, s, cond, x, c_cond, c_x,
             w_kc, bias, B, R, Cout, out);
  CUDA_CHECK(cudaGetLastError());
}

// =============================================================================== tail + update
// (philox4x32_10, box_muller and posterior_update live in common.cuh: the halo conv's tail epilogue
// uses them too.)
// One thread per pixel: eps[oc] = bias[oc] + sum_{tap,c} src[pix+tap][c] * w[oc][tap][c]
// then the reference's update, op for op (diffusion.py:150-151, 175-176, 159-160, 186-187).
template <int OC>
__global__ void __launch_bounds__(128) tail_kernel(TailPlan t) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float w_s[];   // [OC][9][C]
  const int C = t.C, R = t.R;
  for (int i = threadIdx.x; i < OC * 9 * C; i += blockDim.x) w_s[i] = t.w[i];
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long npix = (long long)t.B * R * R;
  if (pix >= npix) return;
  const int xw = (int)(pix % R);
  const int yh = (int)((pix / R) % R);
  const int b = (int)(pix / ((long long)R * R));
  float acc[OC];
#pragma unroll
  for (int o = 0; o < OC; ++o) acc[o] = t.bias[o];
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = yh + tap / 3 - 1, xx = xw + tap % 3 - 1;
    if (yy < 0 || yy >= R || xx < 0 || xx >= R) continue;
    const bf16* src = t.src + (((size_t)b * R + yy) * R + xx) * C;
    for (int c = 0; c < C; c += 8) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + c)), f);
#pragma unroll
      for (int o = 0; o < OC; ++o) {
        const float4* wr = reinterpret_cast<const float4*>(w_s + (o * 9 + tap) * C + c);
        const float4 w0 = wr[0], w1 = wr[1];
        acc[o] = fmaf(f[0], w0.x, acc[o]); acc[o] = fmaf(f[1], w0.y, acc[o]);
        acc[o] = fmaf(f[2], w0.z, acc[o]); acc[o] = fmaf(f[3], w0.w, acc[o]);
        acc[o] = fmaf(f[4], w1.x, acc[o]); acc[o] = fmaf(f[5], w1.y, acc[o]);
        acc[o] = fmaf(f[6], w1.z, acc[o]); acc[o] = fmaf(f[7], w1.w, acc[o]);
      }
    }
  }
  const size_t plane = (size_t)R * R;
  const size_t base = (size_t)b * OC * plane + (size_t)yh * R + xw;
  if (t.eps_out) {
#pragma unroll
    for (int o = 0; o < OC; ++o) t.eps_out[base + o * plane] = acc[o];
  }
  if (t.x == nullptr) return;
  const int ts = t.ctl->t, T = t.ctl->T;
  const float a = t.coefs[ts], bc = t.coefs[T + ts], c1 = t.coefs[2 * T + ts], c2 = t.coefs[3 * T + ts];
  const float sigma = expf(0.5f * t.coefs[4 * T + ts]);
  float z[OC];
#pragma unroll
  for (int o = 0; o < OC; ++o) z[o] = 0.f;
  const int mode = t.ctl->noise_mode;
  if (ts > 0) {
    if (mode == 1 || mode == 3) {
      const float* zp = t.ctl->noise;
      if (zp) {
        if (mode == 1) zp += (size_t)(T - ts) * (size_t)t.ctl->numel;
#pragma unroll
        for (int o = 0; o < OC; ++o) z[o] = __ldg(zp + base + o * plane);
      }
    } else if (mode == 2) {
      const unsigned long long seed = t.ctl->seed;
      const long long gpix = pix + t.ctl->row0 * (long long)R * R;      // keyed by the GLOBAL batch row
      const uint4 r = philox4x32_10(make_uint4((uint32_t)gpix, (uint32_t)(gpix >> 32), (uint32_t)ts, 0x5352u),
                                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      const float2 g0 = box_muller(r.x, r.y), g1 = box_muller(r.z, r.w);
      const float zz[4] = {g0.x, g0.y, g1.x, g1.y};
#pragma unroll
      for (int o = 0; o < OC; ++o) z[o] = zz[o & 3];
    }
  }
#pragma unroll
  for (int o = 0; o < OC; ++o) {
    const float xv = t.x[base + o * plane];
    t.x[base + o * plane] = posterior_update(xv, acc[o], z[o], a, bc, c1, c2, sigma, t.ctl->no_clip ? __int_as_float(0x7f800000) : 1.0f);
  }
}

// Fast path (C/8 a power of two <= 32, i.e. C = 64): C/8 lanes share a pixel, each owning 8
// channels, so a tap is one fully coalesced 16-byte load per lane; the OC partial dot products are
// combined with xor-shuffles and lane o < OC of the group applies the posterior update of channel o.
template <int OC>
__global__ void __launch_bounds__(256) tail_split_kernel(TailPlan t) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float w_s[];   // [OC][9][C]
  const int C = t.C, R = t.R;
  for (int i = threadIdx.x; i < OC * 9 * C; i += blockDim.x) w_s[i] = t.w[i];
  __syncthreads();
  const int lpp = C >> 3;                       // lanes per pixel
  const int sub = threadIdx.x & (lpp - 1);
  const long long pix = (long long)blockIdx.x * (blockDim.x / lpp) + threadIdx.x / lpp;
  const long long npix = (long long)t.B * R * R;
  const bool live = pix < npix;                 // whole pixel groups are live or dead together
  const int xw = live ? (int)(pix % R) : 0;
  const int yh = live ? (int)((pix / R) % R) : 0;
  const int b = live ? (int)(pix / ((long long)R * R)) : 0;
  float acc[OC];
#pragma unroll
  for (int o = 0; o < OC; ++o) acc[o] = 0.f;
  if (live) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = yh + tap / 3 - 1, xx = xw + tap % 3 - 1;
      if (yy < 0 || yy >= R || xx < 0 || xx >= R) continue;
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(t.src + (((size_t)b * R + yy) * R + xx) * C + sub * 8)), f);
#pragma unroll
      for (int o = 0; o < OC; ++o) {
        const float4* wr = reinterpret_cast<const float4*>(w_s + (o * 9 + tap) * C + sub * 8);
        const float4 w0 = wr[0], w1 = wr[1];
        acc[o] = fmaf(f[0], w0.x, acc[o]); acc[o] = fmaf(f[1], w0.y, acc[o]);
        acc[o] = fmaf(f[2], w0.z, acc[o]); acc[o] = fmaf(f[3], w0.w, acc[o]);
        acc[o] = fmaf(f[4], w1.x, acc[o]); acc[o] = fmaf(f[5], w1.y, acc[o]);
        acc[o] = fmaf(f[6], w1.z, acc[o]); acc[o] = fmaf(f[7], w1.w, acc[o]);
      }
    }
  }
  for (int m = 1; m < lpp; m <<= 1) {
#pragma unroll
    for (int o = 0; o < OC; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], m);
  }
  if (!live || sub >= OC) return;
  float eps = 0.f;
#pragma unroll
  for (int o = 0; o < OC; ++o) if (o == sub) eps = acc[o] + t.bias[o];
  const size_t plane = (size_t)R * R;
  const size_t idx = ((size_t)b * OC + sub) * plane + (size_t)yh * R + xw;
  if (t.eps_out) t.eps_out[idx] = eps;
  if (t.x == nullptr) return;
  const int ts = t.ctl->t, T = t.ctl->T;
  const float a = t.coefs[ts], bc = t.coefs[T + ts], c1 = t.coefs[2 * T + ts], c2 = t.coefs[3 * T + ts];
  const float sigma = expf(0.5f * t.coefs[4 * T + ts]);
  float z = 0.f;
  const int mode = t.ctl->noise_mode;
  if (ts > 0) {
    if (mode == 1 || mode == 3) {
      const float* zp = t.ctl->noise;
      if (zp) {
        if (mode == 1) zp += (size_t)(T - ts) * (size_t)t.ctl->numel;
        z = __ldg(zp + idx);
      }
    } else if (mode == 2) {
      const unsigned long long seed = t.ctl->seed;
      const long long gpix = pix + t.ctl->row0 * (long long)R * R;      // keyed by the GLOBAL batch row
      const uint4 r = philox4x32_10(make_uint4((uint32_t)gpix, (uint32_t)(gpix >> 32), (uint32_t)ts, 0x5352u),
                                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      const float2 g0 = box_muller(r.x, r.y), g1 = box_muller(r.z, r.w);
      const float zz[4] = {g0.x, g0.y, g1.x, g1.y};
      z = zz[sub & 3];
    }
  }
  t.x[idx] = posterior_update(t.x[idx], eps, z, a, bc, c1, c2, sigma, t.ctl->no_clip ? __int_as_float(0x7f800000) : 1.0f);
}

void launch_tail(const TailPlan& t, cudaStream_t s) {
  REQUIRE(t.C % 8 == 0, "tail conv: C must be a multiple of 8");
  const long long npix = (long long)t.B * t.R * t.R;
  const size_t smem = (size_t)t.OC * 9 * t.C * sizeof(float);
  REQUIRE(smem <= 48 * 1024, "tail conv: weights exceed 48 KB of shared memory");
  const int lpp = t.C / 8;
  if (lpp >= 4 && lpp <= 32 && (lpp & (lpp - 1)) == 0) {
    const int ppb = 256 / lpp;
    const unsigned blocks = (unsigned)((npix + ppb - 1) / ppb);
    switch (t.OC) {
      case 1: launch_pdl(tail_split_kernel<1>, dim3(blocks), dim3(256), smem, s, t); break;
      case 3: launch_pdl(tail_split_kernel<3>, dim3(blocks), dim3(256), smem, s, t); break;
      case 4: launch_pdl(tail_split_kernel<4>, dim3(blocks), dim3(256), smem, s, t); break;
      default: throw Error("tail conv: out_channel must be 1, 3 or 4");
    }
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  const unsigned blocks = (unsigned)((npix + 127) / 128);
  switch (t.OC) {
    case 1: launch_pdl(tail_kernel<1>, dim3(blocks), dim3(128), smem, s, t); break;
    case 3: launch_pdl(tail_kernel<3>, dim3(blocks), dim3(128), smem, s, t); break;
    case 4: launch_pdl(tail_kernel<4>, dim3(blocks), dim3(128), smem, s, t); break;
    default: throw Error("tail conv: out_channel must be 1, 3 or 4");
  }
  CUDA_CHECK(cudaGetLastError());
}

__global__ void philox_fill_kernel(float* x, int B, int C, int R, unsigned long long seed, int t, long long row0) {
  const long long npix = (long long)B * R * R;
  const size_t plane = (size_t)R * R;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += (long long)gridDim.x * blockDim.x) {
    const long long gpix = pix + row0 * (long long)plane;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)gpix, (uint32_t)(gpix >> 32), (uint32_t)t, 0x5352u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float2 g0 = box_muller(r.x, r.y), g1 = box_muller(r.z, r.w);
    const float zz[4] = {g0.x, g0.y, g1.x, g1.y};
    const int b = (int)(pix / plane);
    const size_t base = (size_t)b * C * plane + (size_t)(pix % plane);
    for (int c = 0; c < C; ++c) x[base + c * plane] = zz[c & 3];
  }
}
void launch_philox_fill(float* x, int B, int C, int R, unsigned long long seed, int t, long long row0, cudaStream_t s) {
  REQUIRE(C >= 1 && C <= 4, "philox fill: one Philox block yields the (up to 4) channels of a pixel");
  const long long npix = (long long)B * R * R;
  long long blocks = (npix + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  philox_fill_kernel<<<(int)blocks, 256, 0, s>>>(x, B, C, R, seed, t, row0);
  CUDA_CHECK(cudaGetLastError());
}

__global__ void ctl_advance_kernel(StepCtl* ctl) {
  pdl_launch_dependents();
  pdl_wait();
  ctl->t -= 1;
}
void launch_ctl_advance(StepCtl* ctl, cudaStream_t s) {
  launch_pdl(ctl_advance_kernel, dim3(1), dim3(1), 0, s, ctl);
  CUDA_CHECK(cudaGetLastError());
}

__global__ void posterior_update_kernel(const float* x, const float* eps, const float* z, float a, float b,
                                        float c1, float c2, float logvar, long long n, float* out) {
  const float sigma = expf(0.5f * logvar);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = posterior_update(x[i], eps[i], z ? z[i] : 0.f, a, b, c1, c2, sigma);
}
void launch_posterior_update(const float* x, const float* eps, const float* z, float a, float b, float c1,
                             float c2, float logvar, long long n, float* out, cudaStream_t s) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  posterior_update_kernel<<<(int)blocks, 256, 0, s>>>(x, eps, z, a, b, c1, c2, logvar, n, out);
  CUDA_CHECK(cudaGetLastError());
}

// =============================================================================== attention
// One warp per query token. qkv is the 1x1 conv output [B, HW, 3C] (q | k | v along channels,
// unet.py:128-129). Scores and softmax in fp32; HW <= 64 tokens for every shipped config.
constexpr int ATT_WARPS = 8;
constexpr int ATT_MAXV = 4;   // C <= 1024
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_kernel(const bf16* __restrict__ qkv,
                                                                   bf16* __restrict__ out, int HW, int C) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sc_s[];   // [ATT_WARPS][HW]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int qi = blockIdx.x * ATT_WARPS + warp;
  if (qi >= HW) return;
  float* sc = sc_s + warp * HW;
  const int Cv = C >> 3;
  const size_t row = (size_t)3 * C;
  const bf16* base = qkv + (size_t)b * HW * row;
  float q[ATT_MAXV][8];
#pragma unroll
  for (int i = 0; i < ATT_MAXV; ++i) {
    const int v = lane + 32 * i;
    if (v < Cv) unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)qi * row + v * 8)), q[i]);
  }
  const float scale = rsqrtf((float)C);
  float mx = -INFINITY;
  for (int j = 0; j < HW; ++j) {
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < ATT_MAXV; ++i) {
      const int v = lane + 32 * i;
      if (v < Cv) {
        float k[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)j * row + C + v * 8)), k);
#pragma unroll
        for (int e = 0; e < 8; ++e) d = fmaf(q[i][e], k[e], d);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    d *= scale;
    if (lane == 0) sc[j] = d;
    mx = fmaxf(mx, d);
  }
  __syncwarp();
  float sum = 0.f;
  for (int j = lane; j < HW; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncwarp();
  const float inv = 1.0f / sum;
  float acc[ATT_MAXV][8];
#pragma unroll
  for (int i = 0; i < ATT_MAXV; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;
  for (int j = 0; j < HW; ++j) {
    const float pj = sc[j];
#pragma unroll
    for (int i = 0; i < ATT_MAXV; ++i) {
      const int v = lane + 32 * i;
      if (v < Cv) {
        float vv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)j * row + 2 * C + v * 8)), vv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[i][e] = fmaf(pj, vv[e], acc[i][e]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < ATT_MAXV; ++i) {
    const int v = lane + 32 * i;
    if (v < Cv) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[i][e] *= inv;
      *reinterpret_cast<uint4*>(out + ((size_t)b * HW + qi) * C + v * 8) = pack8(acc[i]);
    }
  }
}

// Tensor-core version for HW in {16, 32, 48, 64} tokens: a CTA owns 16 queries of one image. Q (16 rows)
// and K (all tokens) are staged in shared memory with coalesced 16-byte loads (rows padded by 8
// elements against bank conflicts), S = Q K^T runs on mma.sync (bf16 in, fp32 out: a 0.01%-of-FLOPs
// op, the legacy tensor path is plenty), the softmax is fp32, P is rounded to bf16 and O = P V reuses
// K's buffer for V.
constexpr int ATT_MMA_THREADS = 256;      // 8 warps: the kernel is bound by the latency of staging K and V (2 x 64 KB per CTA)
__global__ void __launch_bounds__(ATT_MMA_THREADS) attention_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                        int HW, int C) {
  constexpr int NT = ATT_MMA_THREADS, NW = ATT_MMA_THREADS / 32;
  pdl_launch_dependents();
  pdl_wait();
  using namespace nvcuda;
  extern __shared__ __align__(128) uint8_t att_smem[];
  const int ld = C + 8;                                        // padded row length (elements)
  bf16* q_s = reinterpret_cast<bf16*>(att_smem);               // [16][ld]
  bf16* kv_s = q_s + 16 * ld;                                  // [HW][ld]
  float* s_s = reinterpret_cast<float*>(kv_s + (size_t)HW * ld);   // [16][HW]
  bf16* p_s = reinterpret_cast<bf16*>(s_s + 16 * HW);          // [16][HW]
  float* scr = reinterpret_cast<float*>(p_s + 16 * HW);        // [NW][256]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, q0 = blockIdx.x * 16;
  const size_t row = (size_t)3 * C;
  const bf16* base = qkv + (size_t)b * HW * row;
  const int cv = C >> 3;
  for (int i = tid; i < 16 * cv; i += NT) {
    const int r = i / cv, c = i - r * cv;
    *reinterpret_cast<uint4*>(q_s + r * ld + c * 8) = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(q0 + r) * row + c * 8));
  }
  for (int i = tid; i < HW * cv; i += NT) {
    const int r = i / cv, c = i - r * cv;
    *reinterpret_cast<uint4*>(kv_s + r * ld + c * 8) = __ldg(reinterpret_cast<const uint4*>(base + (size_t)r * row + C + c * 8));
  }
  __syncthreads();
  // ---- S = Q K^T: one 16x16 tile of keys per warp pass
  for (int t = warp; t < HW / 16; t += NW) {
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
    wmma::fill_fragment(acc, 0.f);
    for (int k = 0; k < C; k += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::row_major> a;
      wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::col_major> kb;
      wmma::load_matrix_sync(a, q_s + k, ld);
      wmma::load_matrix_sync(kb, kv_s + (size_t)t * 16 * ld + k, ld);
      wmma::mma_sync(acc, a, kb, acc);
    }
    wmma::store_matrix_sync(s_s + t * 16, acc, HW, wmma::mem_row_major);
  }
  __syncthreads();
  // ---- V replaces K while the softmax runs
  for (int i = tid; i < HW * cv; i += NT) {
    const int r = i / cv, c = i - r * cv;
    *reinterpret_cast<uint4*>(kv_s + r * ld + c * 8) = __ldg(reinterpret_cast<const uint4*>(base + (size_t)r * row + 2 * C + c * 8));
  }
  if (tid < 128) {
    // 8 lanes per query row (unet.py:133-135: scores / sqrt(C), softmax over all keys)
    const int r = tid >> 3, sub = tid & 7;
    const float scale = rsqrtf((float)C);
    float v[8];
    float mx = -INFINITY;
    int n = 0;
    for (int j = sub; j < HW; j += 8) { v[n] = s_s[r * HW + j] * scale; mx = fmaxf(mx, v[n]); ++n; }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = 0; i < n; ++i) { v[i] = expf(v[i] - mx); sum += v[i]; }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    n = 0;
    for (int j = sub; j < HW; j += 8) p_s[r * HW + j] = __float2bfloat16_rn(v[n++] * inv);
  }
  __syncthreads();
  // ---- O = P V: 16-channel tiles round-robin over the warps
  float* my = scr + warp * 256;
  for (int t = warp; t < C / 16; t += NW) {
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
    wmma::fill_fragment(acc, 0.f);
    for (int k = 0; k < HW; k += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::row_major> a;
      wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::row_major> vb;
      wmma::load_matrix_sync(a, p_s + k, HW);
      wmma::load_matrix_sync(vb, kv_s + (size_t)k * ld + t * 16, ld);
      wmma::mma_sync(acc, a, vb, acc);
    }
    wmma::store_matrix_sync(my, acc, 16, wmma::mem_row_major);
    __syncwarp();
    const int r = lane >> 1, h = lane & 1;
    *reinterpret_cast<uint4*>(out + ((size_t)b * HW + q0 + r) * C + t * 16 + h * 8) = pack8(my + r * 16 + h * 8);
    __syncwarp();
  }
}

void launch_attention(const bf16* qkv, bf16* out, int B, int HW, int C, cudaStream_t s) {
  REQUIRE(C % 8 == 0 && C <= 1024, "attention: C must be a multiple of 8 and <= 1024");
  if (HW % 16 == 0 && HW <= 64 && C % 16 == 0) {
    const size_t smem = (size_t)(16 + HW) * (C + 8) * 2 + (size_t)16 * HW * 4 + (size_t)16 * HW * 2 + (ATT_MMA_THREADS / 32) * 256 * 4;
    static bool attr_set = false;
    if (!attr_set) {
      CUDA_CHECK(cudaFuncSetAttribute(attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    launch_pdl(attention_mma_kernel, dim3(HW / 16, B), dim3(ATT_MMA_THREADS), smem, s, qkv, out, HW, C);
    CUDA_CHECK(cudaGetLastError());
    return;
  }
  const size_t smem = (size_t)ATT_WARPS * HW * sizeof(float);
  REQUIRE(smem <= 160 * 1024, "attention: too many tokens");
  if (smem > 48 * 1024)
    CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  launch_pdl(attention_kernel, dim3(ceil_div(HW, ATT_WARPS), B), dim3(ATT_WARPS * 32), smem, s, qkv, out, HW, C);
  CUDA_CHECK(cudaGetLastError());
}

// =============================================================================== weight packing
__global__ void pack_conv_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int Cout, int Cin,
                                        int taps, int cin_pad, int k_off, int k_total) {
  const long long total = (long long)Cout * taps * cin_pad;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cin_pad);
    const int tap = (int)((idx / cin_pad) % taps);
    const int o = (int)(idx / ((long long)cin_pad * taps));
    const float v = (c < Cin) ? src[((size_t)o * Cin + c) * taps + tap] : 0.f;
    dst[(size_t)o * k_total + k_off + tap * cin_pad + c] = __float2bfloat16_rn(v);
  }
}
void launch_pack_conv_weight(const float* src, bf16* dst, int Cout, int Cin, int taps, int cin_pad, int k_off,
                             int k_total, cudaStream_t s) {
  const long long total = (long long)Cout * taps * cin_pad;
  pack_conv_weight_kernel<<<(int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256), 256, 0, s>>>(
      src, dst, Cout, Cin, taps, cin_pad, k_off, k_total);
  CUDA_CHECK(cudaGetLastError());
}

// Upsample(nearest 2x) + conv3x3 folded into four 2x2 convs over the low-resolution source
// (unet.py:58-65): dst[(par*Cout + o)][t*cin_pad + c] = sum of the 3x3 taps (kh, kw) that land on
// source offset t = (a, b) for output parity par = (py, px). Summed in fp32, rounded once.
__global__ void pack_upfold_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int Cout, int Cin,
                                          int cin_pad) {
  const long long total = 4LL * Cout * 4 * cin_pad;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cin_pad);
    const int t = (int)((idx / cin_pad) % 4);
    const int o = (int)((idx / (4LL * cin_pad)) % Cout);
    const int par = (int)(idx / (4LL * cin_pad * Cout));
    float v = 0.f;
    if (c < Cin) {
      const int py = par >> 1, px = par & 1, a = t >> 1, b = t & 1;
      // py = 0: a = 0 <- kh {0}, a = 1 <- kh {1, 2};  py = 1: a = 0 <- kh {0, 1}, a = 1 <- kh {2}
      const int kh0 = (py == 0) ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2);
      const int kh1 = (py == 0) ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
      const int kw0 = (px == 0) ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2);
      const int kw1 = (px == 0) ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
      const float* w = src + ((size_t)o * Cin + c) * 9;
      for (int kh = kh0; kh <= kh1; ++kh)
        for (int kw = kw0; kw <= kw1; ++kw) v += w[kh * 3 + kw];
    }
    dst[idx] = __float2bfloat16_rn(v);
  }
}
void launch_pack_upfold_weight(const float* src, bf16* dst, int Cout, int Cin, int cin_pad, cudaStream_t s) {
  const long long total = 4LL * Cout * 4 * cin_pad;
  const long long blocks = (total + 255) / 256;
  pack_upfold_weight_kernel<<<(int)(blocks > 8192 ? 8192 : blocks), 256, 0, s>>>(src, dst, Cout, Cin, cin_pad);
  CUDA_CHECK(cudaGetLastError());
}

__global__ void pack_head_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin) {
  const int total = Cout * Cin * 9;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int o = idx % Cout;
    const int ci = (idx / Cout) % Cin;
    const int tap = idx / (Cout * Cin);
    dst[idx] = src[((size_t)o * Cin + ci) * 9 + tap];   // dst[(tap*Cin+ci)*Cout + o]
  }
}
void launch_pack_head_weight(const float* src, float* dst, int Cout, int Cin, cudaStream_t s) {
  pack_head_weight_kernel<<<ceil_div(Cout * Cin * 9, 256), 256, 0, s>>>(src, dst, Cout, Cin);
  CUDA_CHECK(cudaGetLastError());
}

__global__ void pack_tail_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int OC, int Cin) {
  const int total = OC * 9 * Cin;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = idx % Cin;
    const int tap = (idx / Cin) % 9;
    const int o = idx / (Cin * 9);
    dst[idx] = src[((size_t)o * Cin + c) * 9 + tap];    // dst[(o*9+tap)*Cin + c]
  }
}
void launch_pack_tail_weight(const float* src, float* dst, int OC, int Cin, cudaStream_t s) {
  pack_tail_weight_kernel<<<ceil_div(OC * Cin * 9, 256), 256, 0, s>>>(src, dst, OC, Cin);
  CUDA_CHECK(cudaGetLastError());
}

// =============================================================================== noise table
__global__ void __launch_bounds__(256) noise_table_kernel(NoiseTablePlan p, int row0) {
  extern __shared__ float sm[];   // pe[inner] | h[4*inner] | e[inner]
  const int inner = p.inner, hid = 4 * inner, count = inner / 2;
  float* pe = sm;
  float* h = sm + inner;
  float* e = h + hid;
  const int row = row0 + blockIdx.x;
  const float nl = p.nl[row];
  for (int j = threadIdx.x; j < count; j += blockDim.x) {
    const float step = (float)j / (float)count;
    const float enc = nl * expf(-9.210340371976184f * step);
    pe[j] = sinf(enc);
    pe[count + j] = cosf(enc);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hid; i += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < inner; ++j) a = fmaf(p.w1[i * inner + j], pe[j], a);
    a += p.b1[i];
    h[i] = a / (1.0f + expf(-a));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < inner; i += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < hid; ++j) a = fmaf(p.w3[i * hid + j], h[j], a);
    e[i] = a + p.b3[i];
  }
  __syncthreads();
  float* dst = p.table + (size_t)row * p.total;
  for (int n = threadIdx.x; n < p.total; n += blockDim.x) {
    float a = 0.f;
    const float* wr = p.wall + (size_t)n * inner;
    for (int j = 0; j < inner; ++j) a = fmaf(wr[j], e[j], a);
    dst[n] = a + p.ball[n];
  }
}
void launch_noise_table(const NoiseTablePlan& p, int row0, int rows, cudaStream_t s) {
  if (rows <= 0) return;
  noise_table_kernel<<<rows, 256, 6 * p.inner * sizeof(float), s>>>(p, row0);
  CUDA_CHECK(cudaGetLastError());
}

// =============================================================================== layout helpers
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int B, int C, int H, int W) {
  const long long total = (long long)B * C * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long long p = idx / C;
    const int x = (int)(p % W); p /= W;
    const int y = (int)(p % H);
    const int b = (int)(p / H);
    dst[idx] = __float2bfloat16_rn(src[(((size_t)b * C + c) * H + y) * W + x]);
  }
}
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int B, int C, int H, int W) {
  const long long total = (long long)B * C * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    long long p = idx / W;
    const int y = (int)(p % H); p /= H;
    const int c = (int)(p % C);
    const int b = (int)(p / C);
    dst[idx] = __bfloat162float(src[(((size_t)b * H + y) * W + x) * C + c]);
  }
}
static int grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  return (int)(blocks > 148LL * 32 ? 148LL * 32 : (blocks < 1 ? 1 : blocks));
}
void launch_nchw_to_nhwc(const float* src, bf16* dst, int B, int C, int H, int W, cudaStream_t s) {
  nchw_to_nhwc_kernel<<<grid_for((long long)B * C * H * W), 256, 0, s>>>(src, dst, B, C, H, W);
  CUDA_CHECK(cudaGetLastError());
}
void launch_nhwc_to_nchw(const bf16* src, float* dst, int B, int C, int H, int W, cudaStream_t s) {
  nhwc_to_nchw_kernel<<<grid_for((long long)B * C * H * W), 256, 0, s>>>(src, dst, B, C, H, W);
  CUDA_CHECK(cudaGetLastError());
}
__global__ void fill_f32_kernel(float* p, float v, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
void launch_fill_f32(float* p, float v, long long n, cudaStream_t s) {
  fill_f32_kernel<<<grid_for(n), 256, 0, s>>>(p, v, n);
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace b200sr3
