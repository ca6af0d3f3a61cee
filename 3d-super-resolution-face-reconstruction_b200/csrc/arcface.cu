// MICA identity encoder on the device: ArcFace iResNet-100 (model/mica/arcface.py:40-200, eval
// mode), F.normalize (model/sr3d/model.py:167) and the MappingNetwork regressor (model/mica/generator.py:31-60), i.e.
// everything between the SR -> MICA hand-off blob (handoff.cu) and the FLAME decoder (licensed assets, out of scope).
//
// The 100 convolutions run on the halo-resident tcgen05 kernel of the SR UNet (conv_halo.cuh):
//   * every BatchNorm that FOLLOWS a conv (bn2 after conv1, bn3 after conv2, downsample.1, the stem's bn1) is folded into
//     that conv's bf16 weights (per-output-channel scale) and fp32 bias (shift) when the weights are packed;
//   * the BatchNorm / PReLU in FRONT of a conv (bn1 before conv1, prelu before conv2) cannot be folded - the conv pads
//     the transformed tensor with zeros - and is applied to the halo tile in shared memory by the kernel's transform
//     warps (PRELU mode: y = scale * prelu(x; slope) + shift per channel, padding left zero), exactly where the UNet path
//     applies GroupNorm + Swish. No normalised tensor is ever written to HBM;
//   * the residual add rides the GEMM: `out += identity` is an identity-matrix K segment, the stride-2 blocks'
//     downsample (conv1x1 stride 2 + BN) a 1x1 K segment over the even/even input view;
//   * stride-2 convs read the four input-parity views (conv_halo.cu); 56 / 28 / 14 / 7 px images use partial tiles.
// The stem (3 -> 64 at 112 px) is the UNet's head kernel: split-precision operand built in shared memory from the fp32
// blob. fc (with bn2 and the `features` BatchNorm1d folded in), the L2 normalisation and the 5 small Linear layers of
// the regressor are plain CUDA-core kernels (0.4 % of the FLOPs).
#include "arcface.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace b200sr3 {

static const int ARC_LAYERS[4] = {3, 13, 30, 3};      // arcface.py:167
static const int ARC_PLANES[4] = {64, 128, 256, 512};
static const int ARC_RES = 112;

// ------------------------------------------------------------------------------- small kernels
// BatchNorm (eval) -> per-channel (scale, shift): y = x * scale + shift (arcface.py: eps = 1e-5 everywhere)
__global__ void bn_fold_kernel(const float* w, const float* b, const float* mean, const float* var, int C, float* scale,
                               float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = w[c] / sqrtf(var[c] + 1e-5f);
  scale[c] = sc;
  shift[c] = b[c] - mean[c] * sc;
}
// (scale, shift) rows the conv's transform warps read; slope rows
__global__ void xf_row_kernel(const float* scale, const float* shift, int C, float2* row) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) row[c] = make_float2(scale ? scale[c] : 1.f, shift ? shift[c] : 0.f);
}
// dst[o][...] = src[o][...] * oscale[o]  (fold a following BatchNorm's scale into conv weights, fp32)
__global__ void scale_rows_kernel(const float* __restrict__ src, const float* __restrict__ oscale, float* __restrict__ dst,
                                  long long per_row, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i] * oscale[i / per_row];
}
__global__ void add_vec2_kernel(const float* a, const float* b, float* o, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + (b ? b[i] : 0.f);
}
// in-place PReLU over an NHWC bf16 tensor (the stem's activation, arcface.py:189)
__global__ void prelu_nhwc_kernel(bf16* __restrict__ x, const float* __restrict__ slope, int C, long long n_vec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
    uint4* p = reinterpret_cast<uint4*>(x) + i;
    float f[8];
    unpack8(*p, f);
    const int c0 = (int)((i * 8) % C);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = f[e] > 0.f ? f[e] : f[e] * __ldg(slope + c0 + e);
    *p = pack8(f);
  }
}
// fc with bn2 (BatchNorm2d in front) and `features` (BatchNorm1d behind) folded in. The activation is NHWC, the
// reference flattens NCHW (arcface.py:195): W'[o][(y*7+x)*C + c] = fs[o] * W[o][c*49 + y*7+x] * s2[c],
// bias'[o] = fs[o] * (b[o] + sum_k W[o][k] * t2[c(k)]) + ft[o]. One CTA per output row.
__global__ void __launch_bounds__(256) fc_fold_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                                      const float* __restrict__ s2, const float* __restrict__ t2,
                                                      const float* __restrict__ fs, const float* __restrict__ ft, int C,
                                                      int HW, float* __restrict__ Wf, float* __restrict__ bf) {
  const int o = blockIdx.x;
  const int K = C * HW;
  float corr = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const int c = k / HW, pix = k - c * HW;
    const float w = W[(size_t)o * K + k];
    corr += w * t2[c];
    Wf[(size_t)o * K + (size_t)pix * C + c] = fs[o] * w * s2[c];
  }
  __shared__ float red[256];
  red[threadIdx.x] = corr;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) bf[o] = fs[o] * (b[o] + red[0]) + ft[o];
}
// y[b][o] = bias[o] + sum_k W[o][k] * x[b][k], x bf16 [B][K] (the NHWC activation), W fp32. A CTA owns 8 outputs x 8
// images and splits K over its 256 threads (64 fp32 accumulators each), then reduces through shared memory.
__global__ void __launch_bounds__(256) fc_kernel(const bf16* __restrict__ x, const float* __restrict__ W,
                                                 const float* __restrict__ bias, int B, int K, int O, float* __restrict__ y) {
  const int o0 = blockIdx.x * 8, b0 = blockIdx.y * 8;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k = threadIdx.x; k < K; k += 256) {
    float w[8], v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = o0 + i < O ? __ldg(W + (size_t)(o0 + i) * K + k) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = b0 + j < B ? __bfloat162float(x[(size_t)(b0 + j) * K + k]) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(w[i], v[j], acc[i][j]);
  }
  __shared__ float red[8][64];      // per warp, the 64 (output, image) sums
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = acc[i][j];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
      if (lane == 0) red[warp][i * 8 + j] = a;
    }
  __syncthreads();
  if (threadIdx.x < 64) {
    float a = 0.f;
#pragma unroll
    for (int wi = 0; wi < 8; ++wi) a += red[wi][threadIdx.x];      // fixed order: deterministic
    const int i = threadIdx.x >> 3, j = threadIdx.x & 7;
    if (o0 + i < O && b0 + j < B) y[(size_t)(b0 + j) * O + o0 + i] = a + bias[o0 + i];
  }
}
// F.normalize (model/sr3d/model.py:167): x / max(||x||_2, 1e-12) per row. One warp per row.
__global__ void l2_normalize_kernel(const float* __restrict__ x, int B, int D, float* __restrict__ y) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) { const float v = x[(size_t)row * D + k]; s = fmaf(v, v, s); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  for (int k = lane; k < D; k += 32) y[(size_t)row * D + k] = x[(size_t)row * D + k] * inv;
}
// y[b][o] = act(bias[o] + W[o][:] . x[b][:]), fp32, one warp per (b, o); act = leaky_relu(0.2) or identity
// (generator.py:50-60)
__global__ void linear_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                              int B, int K, int O, int leaky, float* __restrict__ y) {
  const long long idx = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (idx >= (long long)B * O) return;
  const int b = (int)(idx / O), o = (int)(idx - (long long)b * O);
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s = fmaf(__ldg(W + (size_t)o * K + k), x[(size_t)b * K + k], s);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
    s += bias[o];
    y[(size_t)b * O + o] = (leaky && s < 0.f) ? 0.2f * s : s;
  }
}
__global__ void identity_rows_kernel(bf16* __restrict__ dst, int Cout, int k_off, int k_total) {
  const long long total = (long long)Cout * Cout;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cout), o = (int)(idx / Cout);
    dst[(size_t)o * k_total + k_off + c] = __float2bfloat16_rn(c == o ? 1.f : 0.f);
  }
}

// ------------------------------------------------------------------------------- construction
MicaEncoder::MicaEncoder(int device, int z_dim, int map_hidden_dim, int map_layers, int n_shape)
    : device_(device), z_dim_(z_dim), map_hidden_(map_hidden_dim), map_layers_(map_layers), n_shape_(n_shape) {
  REQUIRE(z_dim == 512, "mica: the ArcFace embedding has 512 features (arcface.py:76)");
  REQUIRE(map_hidden_dim >= 1 && n_shape >= 1 && map_layers >= 0 && map_layers <= 5,
          "mica: mapping network with 0..5 hidden layers (more adds skip connections, generator.py:35-38)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw Error(std::string("no CUDA device available (") + cudaGetErrorString(e) + "): b200sr3 has no CPU fallback");
  REQUIRE(device >= 0 && device < ndev, "device index out of range");
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) throw Error(std::string("device '") + prop.name + "' is not sm_100: b200sr3 is built for B200 only");
  CUDA_CHECK(cudaSetDevice(device));
  conv_init_device();
  CUDA_CHECK(cudaStreamCreateWithFlags(&capture_stream_, cudaStreamNonBlocking));

  auto bn = [&](const std::string& k, int c) {
    add_tensor(k + ".weight", {c}); add_tensor(k + ".bias", {c});
    add_tensor(k + ".running_mean", {c}); add_tensor(k + ".running_var", {c});
  };
  // state_dict keys of the reference Arcface module (arcface.py:90-111; num_batches_tracked is accepted and ignored)
  add_tensor("arcface.conv1.weight", {64, 3, 3, 3});
  bn("arcface.bn1", 64);
  add_tensor("arcface.prelu.weight", {64});
  int inplanes = 64;
  for (int li = 0; li < 4; ++li)
    for (int bi = 0; bi < ARC_LAYERS[li]; ++bi) {
      Block b;
      b.name = "layer" + std::to_string(li + 1) + "." + std::to_string(bi);
      b.inplanes = inplanes; b.planes = ARC_PLANES[li]; b.stride = bi == 0 ? 2 : 1; b.down = bi == 0;
      const std::string k = "arcface." + b.name;
      bn(k + ".bn1", b.inplanes);
      add_tensor(k + ".conv1.weight", {b.planes, b.inplanes, 3, 3});
      bn(k + ".bn2", b.planes);
      add_tensor(k + ".prelu.weight", {b.planes});
      add_tensor(k + ".conv2.weight", {b.planes, b.planes, 3, 3});
      bn(k + ".bn3", b.planes);
      if (b.down) {
        add_tensor(k + ".downsample.0.weight", {b.planes, b.inplanes, 1, 1});
        bn(k + ".downsample.1", b.planes);
      }
      blocks_.push_back(b);
      inplanes = b.planes;
    }
  bn("arcface.bn2", 512);
  add_tensor("arcface.fc.weight", {512, 512 * 49});
  add_tensor("arcface.fc.bias", {512});
  bn("arcface.features", 512);
  // MappingNetwork (generator.py:40-47)
  for (int i = 0; i <= map_layers_; ++i) {
    add_tensor("regressor.network." + std::to_string(i) + ".weight", {map_hidden_, i == 0 ? z_dim_ : map_hidden_});
    add_tensor("regressor.network." + std::to_string(i) + ".bias", {map_hidden_});
  }
  add_tensor("regressor.output.weight", {n_shape_, map_hidden_});
  add_tensor("regressor.output.bias", {n_shape_});
}

MicaEncoder::~MicaEncoder() {
  cudaSetDevice(device_);
  plans_.clear();
  for (auto& t : tensors_) if (t.dev) cudaFree(t.dev);
  for (void* p : owned_) cudaFree(p);
  if (capture_stream_) cudaStreamDestroy(capture_stream_);
}

void MicaEncoder::add_tensor(const std::string& key, std::vector<int64_t> shape) {
  TensorSpec t;
  t.key = key;
  t.shape = std::move(shape);
  tensor_index_[key] = (int)tensors_.size();
  tensors_.push_back(std::move(t));
}

void MicaEncoder::load_tensor(const std::string& key, const float* data, const int64_t* shape, int ndim) {
  if (key.size() > 20 && key.compare(key.size() - 20, 20, ".num_batches_tracked") == 0) return;      // BatchNorm bookkeeping
  auto it = tensor_index_.find(key);
  if (it == tensor_index_.end()) throw Error("mica load_tensor: unexpected key '" + key + "'");
  TensorSpec& t = tensors_[it->second];
  bool ok = (int)t.shape.size() == ndim;
  for (int i = 0; ok && i < ndim; ++i) ok = t.shape[i] == shape[i];
  if (!ok) throw Error("mica load_tensor: shape mismatch for '" + key + "'");
  REQUIRE(data != nullptr, "mica load_tensor: null data");
  CUDA_CHECK(cudaSetDevice(device_));
  if (!t.dev) CUDA_CHECK(cudaMalloc(&t.dev, t.numel() * sizeof(float)));
  CUDA_CHECK(cudaMemcpy(t.dev, data, t.numel() * sizeof(float), cudaMemcpyDefault));
  t.loaded = true;
  finalized_ = false;
}

float* MicaEncoder::T_(const std::string& key) const {
  auto it = tensor_index_.find(key);
  if (it == tensor_index_.end()) throw Error("internal: unknown tensor '" + key + "'");
  const TensorSpec& t = tensors_[it->second];
  if (!t.loaded) throw Error("mica weights: tensor '" + key + "' was never loaded");
  return t.dev;
}

void MicaEncoder::finalize_weights(cudaStream_t s) {
  CUDA_CHECK(cudaSetDevice(device_));
  for (auto& t : tensors_)
    if (!t.loaded) throw Error("mica finalize_weights: tensor '" + t.key + "' was never loaded");
  plans_.clear();
  for (void* p : owned_) cudaFree(p);
  owned_.clear();
  auto dalloc = [&](size_t bytes) {
    void* p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, bytes));
    owned_.push_back(p);
    return p;
  };
  auto fold = [&](const std::string& k, int c, float*& scale, float*& shift) {
    scale = (float*)dalloc(c * sizeof(float));
    shift = (float*)dalloc(c * sizeof(float));
    bn_fold_kernel<<<ceil_div(c, 128), 128, 0, s>>>(T_(k + ".weight"), T_(k + ".bias"), T_(k + ".running_mean"),
                                                   T_(k + ".running_var"), c, scale, shift);
    CUDA_CHECK(cudaGetLastError());
  };
  auto row = [&](const float* scale, const float* shift, int c) {
    float2* r = (float2*)dalloc(c * sizeof(float2));
    xf_row_kernel<<<ceil_div(c, 128), 128, 0, s>>>(scale, shift, c, r);
    CUDA_CHECK(cudaGetLastError());
    return r;
  };
  size_t scratch_elems = (size_t)512 * 512 * 9;
  float* scratch = (float*)dalloc(scratch_elems * sizeof(float));      // conv weights with the following BN's scale folded in
  auto scaled = [&](const float* w, const float* oscale, long long per_row, int rows) {
    const long long total = per_row * rows;
    REQUIRE((size_t)total <= scratch_elems, "internal: fold scratch too small");
    scale_rows_kernel<<<(int)std::min<long long>((total + 255) / 256, 4096), 256, 0, s>>>(w, oscale, scratch, per_row, total);
    CUDA_CHECK(cudaGetLastError());
    return scratch;
  };
  ones_ = (float*)dalloc(512 * sizeof(float));
  launch_fill_f32(ones_, 1.f, 512, s);

  // ---- stem: conv1 (3 -> 64) with bn1 folded; split-precision head operand (kernels.cu)
  float *s1, *t1;
  fold("arcface.bn1", 64, s1, t1);
  stem_.cout = 64; stem_.taps = 9; stem_.cin_main = CONV_BLOCK_K; stem_.k_total = 9 * CONV_BLOCK_K;
  stem_.w = (bf16*)dalloc((size_t)64 * stem_.k_total * sizeof(bf16));
  launch_pack_head_split_weight(scaled(T_("arcface.conv1.weight"), s1, 27, 64), stem_.w, 64, 3, s);
  stem_.bias = t1;

  for (Block& b : blocks_) {
    const std::string k = "arcface." + b.name;
    float *sa, *ta, *sb, *tb, *sc, *tc;
    fold(k + ".bn1", b.inplanes, sa, ta);
    fold(k + ".bn2", b.planes, sb, tb);
    fold(k + ".bn3", b.planes, sc, tc);
    b.xf1 = row(sa, ta, b.inplanes);            // bn1 in front of conv1: affine, no PReLU
    b.xf2 = row(nullptr, nullptr, b.planes);    // prelu in front of conv2: identity affine + slopes
    b.slope2 = T_(k + ".prelu.weight");
    // conv1 * bn2.scale, bias = bn2.shift
    b.c1 = PackedConv();
    b.c1.cout = b.planes; b.c1.taps = 9; b.c1.cin_main = b.inplanes; b.c1.k_total = 9 * b.inplanes;
    b.c1.w = (bf16*)dalloc((size_t)b.planes * b.c1.k_total * sizeof(bf16));
    launch_pack_conv_weight(scaled(T_(k + ".conv1.weight"), sb, 9LL * b.inplanes, b.planes), b.c1.w, b.planes, b.inplanes,
                            9, b.inplanes, 0, b.c1.k_total, s);
    b.c1.bias = tb;
    // conv2 * bn3.scale (+ shortcut as extra K columns), bias = bn3.shift (+ downsample.1 shift)
    b.c2 = PackedConv();
    const int csc = b.down ? b.inplanes : b.planes;
    b.c2.cout = b.planes; b.c2.taps = 9; b.c2.cin_main = b.planes; b.c2.c_res0 = csc; b.c2.k_total = 9 * b.planes + csc;
    b.c2.res_identity = !b.down;
    b.c2.down_perm = b.stride == 2;
    b.c2.w = (bf16*)dalloc((size_t)b.planes * b.c2.k_total * sizeof(bf16));
    const float* w2 = scaled(T_(k + ".conv2.weight"), sc, 9LL * b.planes, b.planes);
    if (b.stride == 2) pack_conv_weight_by_input_parity(w2, b.c2.w, b.planes, b.planes, s, b.c2.k_total);
    else launch_pack_conv_weight(w2, b.c2.w, b.planes, b.planes, 9, b.planes, 0, b.c2.k_total, s);
    if (b.down) {
      float *sd, *td;
      fold(k + ".downsample.1", b.planes, sd, td);
      launch_pack_conv_weight(scaled(T_(k + ".downsample.0.weight"), sd, b.inplanes, b.planes), b.c2.w, b.planes, b.inplanes,
                              1, b.inplanes, 9 * b.planes, b.c2.k_total, s);
      float* bias = (float*)dalloc(b.planes * sizeof(float));
      add_vec2_kernel<<<ceil_div(b.planes, 128), 128, 0, s>>>(tc, td, bias, b.planes);
      CUDA_CHECK(cudaGetLastError());
      b.c2.bias = bias;
    } else {
      identity_rows_kernel<<<(int)std::min<long long>(((long long)b.planes * b.planes + 255) / 256, 4096), 256, 0, s>>>(
          b.c2.w, b.planes, 9 * b.planes, b.c2.k_total);
      CUDA_CHECK(cudaGetLastError());
      b.c2.bias = tc;
    }
  }
  // ---- fc with bn2 in front and `features` behind
  float *s2, *t2, *fs, *ft;
  fold("arcface.bn2", 512, s2, t2);
  fold("arcface.features", 512, fs, ft);
  fc_w_ = (float*)dalloc((size_t)512 * 512 * 49 * sizeof(float));
  fc_b_ = (float*)dalloc(512 * sizeof(float));
  fc_fold_kernel<<<512, 256, 0, s>>>(T_("arcface.fc.weight"), T_("arcface.fc.bias"), s2, t2, fs, ft, 512, 49, fc_w_, fc_b_);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(s));
  finalized_ = true;
}

// ------------------------------------------------------------------------------- per-batch plan
MicaEncoder::Plan::~Plan() {
  if (graph) cudaGraphExecDestroy(graph);
  for (void* p : allocations) cudaFree(p);
}

MicaEncoder::Plan& MicaEncoder::plan(int B) {
  REQUIRE(finalized_, "mica: weights are not finalized (call b200sr3_mica_finalize_weights)");
  REQUIRE(B >= 1, "mica: B must be positive");
  for (size_t i = 0; i < plans_.size(); ++i)
    if (plans_[i]->B == B) {
      std::rotate(plans_.begin() + i, plans_.begin() + i + 1, plans_.end());
      return *plans_.back();
    }
  while (plans_.size() >= 2) plans_.erase(plans_.begin());
  std::unique_ptr<Plan> pl(new Plan());
  pl->B = B;
  build_plan(*pl);
  plans_.push_back(std::move(pl));
  return *plans_.back();
}

void MicaEncoder::build_plan(Plan& pl) {
  CUDA_CHECK(cudaSetDevice(device_));
  const int B = pl.B;
  auto dalloc = [&](size_t bytes) {
    void* p = nullptr;
    bytes = (bytes + 255) / 256 * 256;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) throw Error(std::string("mica workspace allocation failed: ") + cudaGetErrorString(e));
    pl.allocations.push_back(p);
    return p;
  };
  auto act = [&](int H, int W, int C) {
    Act a;
    a.B = B; a.H = H; a.W = W; a.C = C;
    a.ptr = (bf16*)dalloc(a.elems() * sizeof(bf16));
    return a;
  };
  pl.blob = (float*)dalloc((size_t)B * 3 * ARC_RES * ARC_RES * sizeof(float));
  pl.emb = (float*)dalloc((size_t)B * 512 * sizeof(float));
  pl.ident = (float*)dalloc((size_t)B * 512 * sizeof(float));
  pl.shape = (float*)dalloc((size_t)B * n_shape_ * sizeof(float));
  float* h0 = (float*)dalloc((size_t)B * map_hidden_ * sizeof(float));
  float* h1 = (float*)dalloc((size_t)B * map_hidden_ * sizeof(float));

  // stem (arcface.py:187-189): conv1 + folded bn1 on the head kernel, then PReLU in place
  Act cur = act(ARC_RES, ARC_RES, 64);
  {
    Act virt;
    virt.B = B; virt.H = ARC_RES; virt.W = ARC_RES; virt.C = CONV_BLOCK_K;
    HaloHead hh;
    hh.cond = nullptr; hh.x = pl.blob; hh.cc = 0; hh.cx = 3;
    HaloConvExtra ex;
    ex.head = &hh;
    pl.ops.push_back(make_conv_halo_op("stem.conv1", {HaloSource{virt, 9, -1}}, false, stem_, stem_.bias, 0, nullptr, cur,
                                       nullptr, 0, true, nullptr, ex));
    pl.ops.back().flops = 2.0 * B * ARC_RES * ARC_RES * 64.0 * 27.0;
    const float* slope = T_("arcface.prelu.weight");
    bf16* ptr = cur.ptr;
    const long long n_vec = (long long)cur.elems() / 8;
    pl.ops.push_back(Op{"stem.prelu", false, [ptr, slope, n_vec](cudaStream_t s) {
      prelu_nhwc_kernel<<<(int)std::min<long long>((n_vec + 255) / 256, 148 * 8), 256, 0, s>>>(ptr, slope, 64, n_vec);
      CUDA_CHECK(cudaGetLastError());
    }});
    pl.layer_out["stem"] = cur;
  }
  for (const Block& b : blocks_) {
    REQUIRE(cur.C == b.inplanes, "internal: arcface channel plan mismatch");
    // out = conv1(bn1(x)) with bn2 folded (arcface.py:60-62)
    Act t = act(cur.H, cur.W, b.planes);
    HaloConvExtra e1;
    e1.prelu_slope = ones_;           // bn1: affine only
    e1.partial_tiles = true;
    pl.ops.push_back(make_conv_halo_op(b.name + ".conv1", {HaloSource{cur, 9, 0}}, false, b.c1, b.c1.bias, 0, nullptr, t, b.xf1,
                                       b.inplanes, false, nullptr, e1));
    // out = conv2(prelu(out)) with bn3 folded, + identity / downsample(x) as K segments (arcface.py:63-69)
    Act y = act(cur.H / b.stride, cur.W / b.stride, b.planes);
    HaloConvExtra e2;
    e2.prelu_slope = b.slope2;        // prelu in front of conv2
    e2.partial_tiles = true;
    e2.stride = b.stride;
    pl.ops.push_back(make_conv_halo_op(b.name + ".conv2", {HaloSource{t, 9, 0}, HaloSource{cur, 1, -1, !b.down}}, false, b.c2,
                                       b.c2.bias, 0, nullptr, y, b.xf2, b.planes, false, nullptr, e2));
    cur = y;
    pl.layer_out[b.name] = cur;
  }
  REQUIRE(cur.H == 7 && cur.W == 7 && cur.C == 512, "internal: arcface output shape");
  {
    const bf16* x = cur.ptr;
    const float *W = fc_w_, *bias = fc_b_;
    float *emb = pl.emb, *ident = pl.ident, *shape = pl.shape;
    pl.ops.push_back(Op{"fc", false, [=](cudaStream_t s) {
      fc_kernel<<<dim3(512 / 8, ceil_div(B, 8)), 256, 0, s>>>(x, W, bias, B, 512 * 49, 512, emb);
      CUDA_CHECK(cudaGetLastError());
    }});
    pl.ops.back().flops = 2.0 * B * 512.0 * 512.0 * 49.0;
    pl.ops.push_back(Op{"normalize", false, [=](cudaStream_t s) {
      l2_normalize_kernel<<<ceil_div(B, 8), 256, 0, s>>>(emb, B, 512, ident);
      CUDA_CHECK(cudaGetLastError());
    }});
    const float* in = ident;
    int K = z_dim_;
    for (int i = 0; i <= map_layers_ + 1; ++i) {
      const bool last = i == map_layers_ + 1;
      const std::string key = last ? std::string("regressor.output") : "regressor.network." + std::to_string(i);
      const float *Wl = T_(key + ".weight"), *bl = T_(key + ".bias");
      const int O = last ? n_shape_ : map_hidden_;
      float* out = last ? shape : ((i & 1) ? h1 : h0);
      const int leaky = last ? 0 : 1;
      const float* xin = in;
      const int Kc = K;
      pl.ops.push_back(Op{key, false, [=](cudaStream_t s) {
        const long long warps = (long long)B * O;
        linear_kernel<<<(int)((warps + 7) / 8), 256, 0, s>>>(xin, Wl, bl, B, Kc, O, leaky, out);
        CUDA_CHECK(cudaGetLastError());
      }});
      pl.ops.back().flops = 2.0 * B * (double)O * Kc;
      in = out;
      K = O;
    }
  }
  pl.n_conv = 1 + 2 * (int64_t)blocks_.size();
}

void MicaEncoder::encode(const float* blob, int B, float* embedding, float* identity, float* shape_code, cudaStream_t s) {
  CUDA_CHECK(cudaSetDevice(device_));
  REQUIRE(blob != nullptr, "mica encode: null input");
  Plan& pl = plan(B);
  CUDA_CHECK(cudaMemcpyAsync(pl.blob, blob, (size_t)B * 3 * ARC_RES * ARC_RES * sizeof(float), cudaMemcpyDefault, s));
  if (!pl.graph && use_graph_) {
    cudaGraph_t g = nullptr;
    CUDA_CHECK(cudaStreamSynchronize(s));
    CUDA_CHECK(cudaStreamBeginCapture(capture_stream_, cudaStreamCaptureModeThreadLocal));
    try {
      for (auto& op : pl.ops) op.run(capture_stream_);
    } catch (...) {
      cudaStreamEndCapture(capture_stream_, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    CUDA_CHECK(cudaStreamEndCapture(capture_stream_, &g));
    cudaError_t e = cudaGraphInstantiate(&pl.graph, g, 0);
    cudaGraphDestroy(g);
    CUDA_CHECK(e);
  }
  if (pl.graph) CUDA_CHECK(cudaGraphLaunch(pl.graph, s));
  else for (auto& op : pl.ops) op.run(s);
  last_total = (int64_t)pl.ops.size();
  last_conv = pl.n_conv;
  if (embedding) CUDA_CHECK(cudaMemcpyAsync(embedding, pl.emb, (size_t)B * 512 * sizeof(float), cudaMemcpyDefault, s));
  if (identity) CUDA_CHECK(cudaMemcpyAsync(identity, pl.ident, (size_t)B * 512 * sizeof(float), cudaMemcpyDefault, s));
  if (shape_code) CUDA_CHECK(cudaMemcpyAsync(shape_code, pl.shape, (size_t)B * n_shape_ * sizeof(float), cudaMemcpyDefault, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
}

int MicaEncoder::profile(int B, int max_ops, float* ms, double* flops, char* names, int names_len, cudaStream_t s) {
  CUDA_CHECK(cudaSetDevice(device_));
  Plan& pl = plan(B);
  const int n = (int)pl.ops.size();
  REQUIRE(n <= max_ops, "mica profile: output arrays too small");
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) CUDA_CHECK(cudaEventCreate(&e));
  for (int rep = 0; rep < 2; ++rep) {
    CUDA_CHECK(cudaEventRecord(ev[0], s));
    for (int i = 0; i < n; ++i) {
      pl.ops[i].run(s);
      CUDA_CHECK(cudaEventRecord(ev[i + 1], s));
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
  std::string all;
  for (int i = 0; i < n; ++i) {
    CUDA_CHECK(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
    if (flops) flops[i] = pl.ops[i].flops;
    all += pl.ops[i].name;
    all += '\n';
  }
  for (auto& e : ev) cudaEventDestroy(e);
  if (names && names_len > 0) {
    strncpy(names, all.c_str(), (size_t)names_len - 1);
    names[names_len - 1] = 0;
  }
  return n;
}

void MicaEncoder::layer_output(const std::string& layer, float* dst, int* C, int* H, int* W, cudaStream_t s) {
  CUDA_CHECK(cudaSetDevice(device_));
  for (auto pi = plans_.rbegin(); pi != plans_.rend(); ++pi) {
    auto it = (*pi)->layer_out.find(layer);
    if (it == (*pi)->layer_out.end()) continue;
    const Act& a = it->second;
    if (C) *C = a.C;
    if (H) *H = a.H;
    if (W) *W = a.W;
    if (dst) {
      launch_nhwc_to_nchw(a.ptr, dst, a.B, a.C, a.H, a.W, s);
      CUDA_CHECK(cudaStreamSynchronize(s));
    }
    return;
  }
  throw Error("mica layer_output: no activation named '" + layer + "' (run encode first)");
}

}  // namespace b200sr3
