// Launchers of the non-GEMM kernels (definitions in kernels.cu). All activations are NHWC bf16;
// the sampler state (x_t, cond, noise, eps) is fp32 NCHW as at the reference boundary.
#pragma once
#include "common.cuh"

namespace b200sr3 {

// ---- GroupNorm (unet.py:84, 117): y = [swish]((x - mean) * rstd * gamma + beta)
// The input may be the channel concat of two tensors (unet.py:261 feeds cat((x, skip), 1) to the
// block); groups may straddle the seam, so statistics are kept per channel ("chansum": per
// (image, channel) sum and sum of squares over the pixels) and grouped in the apply kernel.
// chansum is produced by the conv that wrote the tensor, or by launch_chan_stats.
struct GnPlan {
  const bf16* src0 = nullptr;   // [B,H,W,C0]
  const bf16* src1 = nullptr;   // [B,H,W,C1] or null
  int B = 0, HW = 0, C0 = 0, C1 = 0, groups = 32;
  const long long* stats0 = nullptr;  // partial sums of src0 [B][slots0][C0][2], 2^-24 fixed point
  const long long* stats1 = nullptr;  // partial sums of src1 [B][slots1][C1][2]
  int slots0 = 1, slots1 = 1;
  const float* gamma = nullptr;   // [C0+C1]
  const float* beta = nullptr;
  bf16* dst = nullptr;            // [B,H,W,C0+C1]
  int swish = 1;
};
void launch_gn_apply(const GnPlan& g, cudaStream_t s);

// (scale, shift) table of a GroupNorm whose apply step is fused into the consuming conv
// (conv_halo.cuh): out[b][c] = (rstd_g * gamma_c, beta_c - mean_g * rstd_g * gamma_c). Uses src0/src1
// statistics, C0, C1, B, HW, groups, gamma, beta of the plan; src/dst pointers are ignored.
void launch_gn_scale_shift(const GnPlan& g, float2* out, cudaStream_t s);

struct ChanStatsPlan {
  const bf16* src = nullptr;     // [B,H,W,C]
  int B = 0, HW = 0, C = 0;
  int chunks = 1;                // CTAs per image
  float* partial = nullptr;      // [B][chunks][C][2] (unused when chunks == 1)
  int* ticket = nullptr;         // [B], zero on entry, self-resetting
  long long* chansum = nullptr;  // [B][C][2], 2^-24 fixed point (one slot)
};
int chan_stats_chunks(int HW, int C);
void launch_chan_stats(const ChanStatsPlan& g, cudaStream_t s);

// ---- head conv (unet.py:196-197, downs.0): fp32 NCHW cond/x -> 3x3 conv -> NHWC bf16
// `stats` (optional, fast path only): [B][Cout][2] int64 fixed-point sums of the output, zeroed here.
bool head_conv_fast_path(int c_in, int R, int Cout);
void launch_head_conv(const float* cond, const float* x, int c_cond, int c_x, const float* w_kc,
                      const float* bias, int B, int R, int Cout, bf16* out, long long* stats, cudaStream_t s);

// Tensor-core head conv: split (hi, lo, hi) bf16 operand [B,R,R,64] from the fp32 NCHW inputs and the matching
// (w_hi, w_hi, w_lo) weight packing [Cout][9*64]; the conv itself is a halo conv (conv_halo.cuh).
void launch_head_pack(const float* cond, const float* x, int c_cond, int c_x, int B, int R, bf16* out, cudaStream_t s);
void launch_pack_head_split_weight(const float* src, bf16* dst, int Cout, int n, cudaStream_t s);

// ---- tail (unet.py:233 final conv on the normalised tensor) fused with the posterior update
// (diffusion.py:144-187): eps = conv3x3(src) ; x0 = clamp(A x - B eps) ; x' = c1 x0 + c2 x + sigma z
struct TailPlan {
  const bf16* src = nullptr;   // GN+Swish output [B,R,R,C]
  const float* w = nullptr;    // [OC][9][C] fp32
  const float* bias = nullptr; // [OC]
  int B = 0, R = 0, C = 0, OC = 3;
  float* eps_out = nullptr;    // optional fp32 NCHW [B,OC,R,R]
  float* x = nullptr;          // fp32 NCHW state, updated in place (null: eps only)
  const float* coefs = nullptr;   // [5][T]: A, Bc, C1, C2, LV
  const StepCtl* ctl = nullptr;
};
void launch_tail(const TailPlan& t, cudaStream_t s);
void launch_ctl_advance(StepCtl* ctl, cudaStream_t s);
// x[b][c][y][x] = N(0,1) from the same Philox stream the update kernel uses, keyed at step t
// N(0,1) draws of the sampler's own stream at key t (t = T: x_T; t < T: z_t), rows [row0, row0 + B) of the global batch
void launch_philox_fill(float* x, int B, int C, int R, unsigned long long seed, int t, long long row0, cudaStream_t s);

// ---- standalone posterior update (kernel-level parity of diffusion.py:144-187 given eps)
void launch_posterior_update(const float* x, const float* eps, const float* z, float a, float b,
                             float c1, float c2, float logvar, long long n, float* out, cudaStream_t s);

// ---- mid-block attention core (unet.py:123-139): softmax(Q K^T / sqrt(C)) V over HW tokens
void launch_attention(const bf16* qkv, bf16* out, int B, int HW, int C, cudaStream_t s);

// ---- weights / tables
// OIHW fp32 -> dst[o][k_off + tap*cin_pad + c] bf16 (zero padded to cin_pad)
void launch_pack_conv_weight(const float* src, bf16* dst, int Cout, int Cin, int taps, int cin_pad,
                             int k_off, int k_total, cudaStream_t s);
// OIHW fp32 3x3 -> [4*Cout][4*cin_pad] bf16: the four parity 2x2 convs of a folded nearest-2x upsample
void launch_pack_upfold_weight(const float* src, bf16* dst, int Cout, int Cin, int cin_pad, cudaStream_t s);
// OIHW fp32 -> [tap*Cin + c][o] fp32 (head) / [o][tap][c] fp32 (tail)
void launch_pack_head_weight(const float* src, float* dst, int Cout, int Cin, cudaStream_t s);
void launch_pack_tail_weight(const float* src, float* dst, int OC, int Cin, cudaStream_t s);

// Noise-embedding bias table (unet.py:18-31 PositionalEncoding, 179-184 noise_level_mlp,
// 34-50 FeatureWiseAffine): row r = Wall * mlp(pe(nl[r])) + ball for r in [row0, row0+rows).
struct NoiseTablePlan {
  const float* nl = nullptr;    // [rows_total] noise level per row
  const float* w1 = nullptr;    // [4*inner][inner]
  const float* b1 = nullptr;
  const float* w3 = nullptr;    // [inner][4*inner]
  const float* b3 = nullptr;
  const float* wall = nullptr;  // [total][inner]
  const float* ball = nullptr;  // [total]
  int inner = 64, total = 0;
  float* table = nullptr;       // [rows_total][total]
};
void launch_noise_table(const NoiseTablePlan& p, int row0, int rows, cudaStream_t s);

// ---- layout conversion (tests / introspection only)
void launch_nchw_to_nhwc(const float* src, bf16* dst, int B, int C, int H, int W, cudaStream_t s);
void launch_nhwc_to_nchw(const bf16* src, float* dst, int B, int C, int H, int W, cudaStream_t s);
void launch_fill_f32(float* p, float v, long long n, cudaStream_t s);

// ---- SR -> MICA hand-off (handoff.cu): core/metrics.py:16-42 tensor2img, model/sr3d/model.py:374 cv2.resize(224),
// :127-131 cv2.dnn.blobFromImages(112, swapRB), :105-124 create_tensor_blob - bit-exact byte/integer kernels
void launch_tensor2img(const float* x, int B, int C, int H, int W, uint8_t* img, cudaStream_t s);
void launch_mica_handoff(const uint8_t* img, int B, int R, uint8_t* up, float* image, float* blob, cudaStream_t s);
void launch_tensor_blob(const float* x, int B, int R, float* blob, cudaStream_t s);

}  // namespace b200sr3
