// Engine: mirrors UNet.__init__/forward (model/sr/sr3_modules/unet.py:161-265) and
// GaussianDiffusion.p_sample_loop (model/sr/sr3_modules/diffusion.py:189-215) as a static
// launch plan over NHWC bf16 buffers.
#include "engine.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace b200sr3 {

// ------------------------------------------------------------------------------- tiny kernels
__global__ void add_vec_kernel(const float* a, const float* b, float* o, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + (b ? b[i] : 0.f);
}
static void add_vec(const float* a, const float* b, float* o, int n, cudaStream_t s) {
  add_vec_kernel<<<ceil_div(n, 256), 256, 0, s>>>(a, b, o, n);
  CUDA_CHECK(cudaGetLastError());
}
// OIHW fp32 (channel slice [c0, c0+cseg) of cin_total) -> packed bf16 K-major row segment. K position `tap` holds filter
// tap order.t[tap] (identity for every conv but the parity-ordered stride-2 ones).
struct TapOrder { int t[9]; };
static const TapOrder TAPS_NATURAL = {{0, 1, 2, 3, 4, 5, 6, 7, 8}};
static const TapOrder TAPS_BY_INPUT_PARITY = {{0, 2, 6, 8, 3, 5, 1, 7, 4}};      // PackedConv::down_perm
__global__ void pack_slice_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int Cout, int cseg, int c0,
                                  int cin_total, int taps, int cpad, int k_off, int k_total, TapOrder order) {
  const long long total = (long long)Cout * taps * cpad;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cpad);
    const int tap = (int)((idx / cpad) % taps);
    const int o = (int)(idx / ((long long)cpad * taps));
    const float v = (c < cseg) ? src[((size_t)o * cin_total + c0 + c) * taps + order.t[tap]] : 0.f;
    dst[(size_t)o * k_total + k_off + tap * cpad + c] = __float2bfloat16_rn(v);
  }
}
static void pack_slice(const float* src, bf16* dst, int Cout, int cseg, int c0, int cin_total, int taps, int cpad,
                       int k_off, int k_total, cudaStream_t s, const TapOrder& order = TAPS_NATURAL) {
  const long long total = (long long)Cout * taps * cpad;
  const long long blocks = std::min<long long>((total + 255) / 256, 8192);
  pack_slice_kernel<<<(int)blocks, 256, 0, s>>>(src, dst, Cout, cseg, c0, cin_total, taps, cpad, k_off, k_total, order);
  CUDA_CHECK(cudaGetLastError());
}

void pack_conv_weight_by_input_parity(const float* w, bf16* dst, int Cout, int Cin, cudaStream_t s, int k_total) {
  pack_slice(w, dst, Cout, Cin, 0, Cin, 9, Cin, 0, k_total ? k_total : 9 * Cin, s, TAPS_BY_INPUT_PARITY);
}

// Identity block appended along K: out += 1.0 * x, i.e. the ResnetBlock's identity shortcut
// (unet.py:110 with res_conv = nn.Identity) rides the GEMM as one more 1x1 segment instead of being
// fetched by the epilogue threads. bf16 x * 1.0 is exact in the fp32 accumulator.
__global__ void pack_identity_kernel(bf16* __restrict__ dst, int Cout, int cpad, int k_off, int k_total) {
  const long long total = (long long)Cout * cpad;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cpad);
    const int o = (int)(idx / cpad);
    dst[(size_t)o * k_total + k_off + c] = __float2bfloat16_rn(c == o ? 1.f : 0.f);
  }
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

Workspace::~Workspace() {
  if (graph) cudaGraphExecDestroy(graph);
  for (void* p : allocations) cudaFree(p);
}

// ------------------------------------------------------------------------------- construction
Engine::Engine(const b200sr3_config& cfg, int device) : cfg_(cfg), device_(device) {
  REQUIRE(cfg.n_mults >= 1 && cfg.n_mults <= B200SR3_MAX_LEVELS, "config: bad channel_multiplier length");
  REQUIRE(cfg.n_attn_res >= 0 && cfg.n_attn_res <= B200SR3_MAX_LEVELS, "config: bad attn_res length");
  REQUIRE(cfg.inner_channel % 8 == 0 && cfg.inner_channel >= 8, "config: inner_channel must be a multiple of 8");
  REQUIRE(cfg.inner_channel % 2 == 0 && cfg.inner_channel <= 256, "config: inner_channel out of range");
  REQUIRE(cfg.norm_groups >= 1 && cfg.norm_groups <= 64, "config: norm_groups out of range");
  REQUIRE(cfg.out_channel == 1 || cfg.out_channel == 3 || cfg.out_channel == 4, "config: out_channel must be 1, 3 or 4");
  REQUIRE(cfg.in_channel == (cfg.conditional ? 2 : 1) * cfg.out_channel,
          "config: in_channel must be out_channel (x) plus out_channel (cond) when conditional");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw Error(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                "): b200sr3 has no CPU fallback");
  REQUIRE(device >= 0 && device < ndev, "device index out of range");
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    throw Error(std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                std::to_string(prop.minor) + "; b200sr3 is built for sm_100a (B200) only");
  CUDA_CHECK(cudaSetDevice(device));
  try {
    conv_init_device();
    CUDA_CHECK(cudaStreamCreateWithFlags(&capture_stream_, cudaStreamNonBlocking));
    CUDA_CHECK(cudaMalloc(&ctl_, sizeof(StepCtl)));
    CUDA_CHECK(cudaMemset(ctl_, 0, sizeof(StepCtl)));
    if (const char* g = getenv("B200SR3_NO_GRAPH")) use_graph_ = !(g[0] == '1');
    if (const char* g = getenv("B200SR3_BLOCK_N")) force_block_n_ = atoi(g);
    if (const char* g = getenv("B200SR3_NO_FUSED_STATS")) fuse_stats_ = !(g[0] == '1');
    if (const char* g = getenv("B200SR3_NO_HALO")) use_halo_ = !(g[0] == '1');
    if (const char* g = getenv("B200SR3_MAX_WORKSPACES")) max_workspaces_ = (size_t)std::max(1, atoi(g));
    build_layers();
  } catch (...) {
    release();          // a throwing constructor runs no destructor
    throw;
  }
}

void Engine::release() {
  cudaSetDevice(device_);
  workspaces_.clear();
  for (auto& t : tensors_) if (t.dev) cudaFree(t.dev);
  tensors_.clear();
  for (void* p : owned_) cudaFree(p);
  owned_.clear();
  if (coefs_) cudaFree(coefs_);
  if (nl_) cudaFree(nl_);
  if (table_) cudaFree(table_);
  if (ctl_) cudaFree(ctl_);
  if (capture_stream_) cudaStreamDestroy(capture_stream_);
  coefs_ = nl_ = table_ = nullptr;
  ctl_ = nullptr;
  capture_stream_ = nullptr;
}

Engine::~Engine() { release(); }

void Engine::add_tensor(const std::string& key, std::vector<int64_t> shape) {
  TensorSpec t;
  t.key = key;
  t.shape = std::move(shape);
  tensor_index_[key] = (int)tensors_.size();
  tensors_.push_back(std::move(t));
}

// Walks the reference constructor (unet.py:175-233) to produce the layer list and the
// state_dict key set.
void Engine::build_layers() {
  const int inner = cfg_.inner_channel;
  auto has_attn = [&](int res) {
    for (int i = 0; i < cfg_.n_attn_res; ++i) if (cfg_.attn_res[i] == res) return true;
    return false;
  };
  add_tensor("noise_level_mlp.1.weight", {4 * inner, inner});
  add_tensor("noise_level_mlp.1.bias", {4 * inner});
  add_tensor("noise_level_mlp.3.weight", {inner, 4 * inner});
  add_tensor("noise_level_mlp.3.bias", {inner});

  auto add_conv = [&](const std::string& k, int co, int ci, int ks, bool bias = true) {
    add_tensor(k + ".weight", {co, ci, ks, ks});
    if (bias) add_tensor(k + ".bias", {co});
  };
  auto add_gn = [&](const std::string& k, int c) {
    add_tensor(k + ".weight", {c});
    add_tensor(k + ".bias", {c});
  };
  auto add_res = [&](const std::string& name, int cx, int cskip, int cout, bool attn) {
    const int cin = cx + cskip;
    LayerDesc l{name, LayerKind::Res, cx, cskip, cout, attn};
    layers_.push_back(l);
    const std::string rb = name + ".res_block";
    add_gn(rb + ".block1.block.0", cin);
    add_conv(rb + ".block1.block.3", cout, cin, 3);
    add_tensor(rb + ".noise_func.noise_func.0.weight", {cout, inner});
    add_tensor(rb + ".noise_func.noise_func.0.bias", {cout});
    add_gn(rb + ".block2.block.0", cout);
    add_conv(rb + ".block2.block.3", cout, cout, 3);
    if (cin != cout) add_conv(rb + ".res_conv", cout, cin, 1);
    if (attn) {
      add_gn(name + ".attn.norm", cout);
      add_conv(name + ".attn.qkv", 3 * cout, cout, 1, false);
      add_conv(name + ".attn.out", cout, cout, 1);
    }
    noise_off_[name] = noise_total_;
    noise_total_ += round_up(cout, 8);
  };

  layers_.push_back(LayerDesc{"downs.0", LayerKind::HeadConv, cfg_.in_channel, 0, inner, false});
  add_conv("downs.0", inner, cfg_.in_channel, 3);
  int pre = inner, now_res = cfg_.image_size, idx = 1;
  std::vector<int> feat{pre};
  for (int lvl = 0; lvl < cfg_.n_mults; ++lvl) {
    const bool last = lvl == cfg_.n_mults - 1;
    const bool attn = has_attn(now_res);
    const int ch = inner * cfg_.channel_mults[lvl];
    for (int r = 0; r < cfg_.res_blocks; ++r) {
      add_res("downs." + std::to_string(idx++), pre, 0, ch, attn);
      feat.push_back(ch);
      pre = ch;
    }
    if (!last) {
      const std::string name = "downs." + std::to_string(idx++);
      layers_.push_back(LayerDesc{name, LayerKind::Down, pre, 0, pre, false});
      add_conv(name + ".conv", pre, pre, 3);
      feat.push_back(pre);
      now_res /= 2;
    }
  }
  add_res("mid.0", pre, 0, pre, true);
  add_res("mid.1", pre, 0, pre, false);
  idx = 0;
  for (int lvl = cfg_.n_mults - 1; lvl >= 0; --lvl) {
    const bool last = lvl < 1;
    const bool attn = has_attn(now_res);
    const int ch = inner * cfg_.channel_mults[lvl];
    for (int r = 0; r < cfg_.res_blocks + 1; ++r) {
      const int sk = feat.back();
      feat.pop_back();
      add_res("ups." + std::to_string(idx++), pre, sk, ch, attn);
      pre = ch;
    }
    if (!last) {
      const std::string name = "ups." + std::to_string(idx++);
      layers_.push_back(LayerDesc{name, LayerKind::Up, pre, 0, pre, false});
      add_conv(name + ".conv", pre, pre, 3);
      now_res *= 2;
    }
  }
  layers_.push_back(LayerDesc{"final_conv", LayerKind::Final, pre, 0, cfg_.out_channel, false});
  add_gn("final_conv.block.0", pre);
  add_conv("final_conv.block.3", cfg_.out_channel, pre, 3);
}

void Engine::load_tensor(const std::string& key, const float* data, const int64_t* shape, int ndim) {
  auto it = tensor_index_.find(key);
  if (it == tensor_index_.end()) throw Error("load_tensor: unexpected key '" + key + "'");
  TensorSpec& t = tensors_[it->second];
  bool ok = (int)t.shape.size() == ndim;
  for (int i = 0; ok && i < ndim; ++i) ok = t.shape[i] == shape[i];
  if (!ok) {
    std::string want, got;
    for (auto d : t.shape) want += std::to_string(d) + ",";
    for (int i = 0; i < ndim; ++i) got += std::to_string(shape[i]) + ",";
    throw Error("load_tensor: shape mismatch for '" + key + "': expected [" + want + "] got [" + got + "]");
  }
  REQUIRE(data != nullptr, "load_tensor: null data");
  CUDA_CHECK(cudaSetDevice(device_));
  if (!t.dev) CUDA_CHECK(cudaMalloc(&t.dev, t.numel() * sizeof(float)));
  CUDA_CHECK(cudaMemcpy(t.dev, data, t.numel() * sizeof(float), cudaMemcpyDefault));
  t.loaded = true;
  finalized_ = false;
}

float* Engine::T_(const std::string& key) const {
  auto it = tensor_index_.find(key);
  if (it == tensor_index_.end()) throw Error("internal: unknown tensor '" + key + "'");
  const TensorSpec& t = tensors_[it->second];
  if (!t.loaded) throw Error("weights: tensor '" + key + "' was never loaded");
  return t.dev;
}

void Engine::finalize_weights(cudaStream_t s) {
  CUDA_CHECK(cudaSetDevice(device_));
  for (auto& t : tensors_)
    if (!t.loaded) throw Error("finalize_weights: tensor '" + t.key + "' was never loaded");
  workspaces_.clear();                 // plans hold pointers into the packed weights
  for (void* p : owned_) cudaFree(p);
  owned_.clear();
  convs_.clear();
  auto dalloc = [&](size_t bytes) {
    void* p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, bytes));
    owned_.push_back(p);
    return p;
  };
  const int inner = cfg_.inner_channel;
  ones_ = (float*)dalloc(1024 * sizeof(float));
  launch_fill_f32(ones_, 1.f, 1024, s);
  wall_ = (float*)dalloc((size_t)noise_total_ * inner * sizeof(float));
  ball_ = (float*)dalloc((size_t)noise_total_ * sizeof(float));
  CUDA_CHECK(cudaMemsetAsync(wall_, 0, (size_t)noise_total_ * inner * sizeof(float), s));
  CUDA_CHECK(cudaMemsetAsync(ball_, 0, (size_t)noise_total_ * sizeof(float), s));

  // plain conv (optionally with folded 1x1 res segments appended along K)
  auto pack = [&](const std::string& name, const std::string& wkey, int cout, int cin, int taps,
                  const std::string& reskey, int c_res0, int c_res1) {
    PackedConv pc;
    pc.cout = cout; pc.taps = taps; pc.cin_main = cin; pc.c_res0 = c_res0; pc.c_res1 = c_res1;
    pc.res_identity = c_res0 && reskey == "identity";
    const int cpad = round_up(cin, CONV_BLOCK_K);
    const int r0pad = c_res0 ? round_up(c_res0, CONV_BLOCK_K) : 0;
    const int r1pad = c_res1 ? round_up(c_res1, CONV_BLOCK_K) : 0;
    pc.k_total = taps * cpad + r0pad + r1pad;
    pc.w = (bf16*)dalloc((size_t)cout * pc.k_total * sizeof(bf16));
    pack_slice(T_(wkey + ".weight"), pc.w, cout, cin, 0, cin, taps, cpad, 0, pc.k_total, s);
    if (c_res0 && reskey == "identity") {
      REQUIRE(c_res0 == cout && c_res1 == 0, "internal: identity shortcut needs Cin == Cout");
      const long long total = (long long)cout * r0pad;
      pack_identity_kernel<<<(int)std::min<long long>((total + 255) / 256, 4096), 256, 0, s>>>(pc.w, cout, r0pad,
                                                                                             taps * cpad, pc.k_total);
      CUDA_CHECK(cudaGetLastError());
    } else if (c_res0)
      pack_slice(T_(reskey + ".weight"), pc.w, cout, c_res0, 0, c_res0 + c_res1, 1, r0pad, taps * cpad, pc.k_total, s);
    if (c_res1) pack_slice(T_(reskey + ".weight"), pc.w, cout, c_res1, c_res0, c_res0 + c_res1, 1, r1pad, taps * cpad + r0pad, pc.k_total, s);
    convs_[name] = pc;
    return &convs_[name];
  };

  for (const LayerDesc& l : layers_) {
    switch (l.kind) {
      case LayerKind::HeadConv: {
        head_w_ = (float*)dalloc((size_t)l.cout * l.c_x * 9 * sizeof(float));
        launch_pack_head_weight(T_(l.name + ".weight"), head_w_, l.cout, l.c_x, s);
        head_pc_ = PackedConv();
        if ((l.c_x == 1 || l.c_x == 2 || l.c_x == 3 || l.c_x == 4 || l.c_x == 6 || l.c_x == 8) && l.cout % 64 == 0) {
          head_pc_.cout = l.cout; head_pc_.taps = 9; head_pc_.cin_main = CONV_BLOCK_K; head_pc_.k_total = 9 * CONV_BLOCK_K;
          head_pc_.w = (bf16*)dalloc((size_t)l.cout * head_pc_.k_total * sizeof(bf16));
          launch_pack_head_split_weight(T_(l.name + ".weight"), head_pc_.w, l.cout, l.c_x, s);
        }
        break;
      }
      case LayerKind::Down: {
        PackedConv* pc = pack(l.name + ".conv", l.name + ".conv", l.cout, l.c_x, 9, "", 0, 0);
        pc->bias = T_(l.name + ".conv.bias");
        if (l.c_x % CONV_BLOCK_K == 0) {
          // the same weights with the taps ordered by input parity: the halo kernel's stride-2 path (conv_halo.cu)
          PackedConv s2 = *pc;
          s2.down_perm = true;
          s2.w = (bf16*)dalloc((size_t)l.cout * s2.k_total * sizeof(bf16));
          pack_slice(T_(l.name + ".conv.weight"), s2.w, l.cout, l.c_x, 0, l.c_x, 9, l.c_x, 0, s2.k_total, s,
                     TAPS_BY_INPUT_PARITY);
          convs_[l.name + ".conv.s2"] = s2;
        }
        break;
      }
      case LayerKind::Up: {
        // Upsample(nearest 2x) + conv3x3 -> four parity 2x2 convs over the low-res tensor
        PackedConv pc;
        pc.cout = l.cout; pc.taps = 9; pc.cin_main = l.c_x; pc.up_folded = true;
        const int cpad = round_up(l.c_x, CONV_BLOCK_K);
        pc.k_total = 4 * cpad;
        pc.w = (bf16*)dalloc((size_t)4 * l.cout * pc.k_total * sizeof(bf16));
        launch_pack_upfold_weight(T_(l.name + ".conv.weight"), pc.w, l.cout, l.c_x, cpad, s);
        pc.bias = T_(l.name + ".conv.bias");
        convs_[l.name + ".conv"] = pc;
        break;
      }
      case LayerKind::Final: {
        tail_w_ = (float*)dalloc((size_t)l.cout * 9 * l.c_x * sizeof(float));
        launch_pack_tail_weight(T_(l.name + ".block.3.weight"), tail_w_, l.cout, l.c_x, s);
        if (l.c_x % CONV_BLOCK_K == 0 && l.cout <= 4) {
          // the same conv as a tensor-core operand: rows >= out_channel are zero
          tail_pc_ = PackedConv();
          tail_pc_.cout = 16; tail_pc_.taps = 9; tail_pc_.cin_main = l.c_x; tail_pc_.k_total = 9 * l.c_x;
          tail_pc_.w = (bf16*)dalloc((size_t)16 * tail_pc_.k_total * sizeof(bf16));
          CUDA_CHECK(cudaMemsetAsync(tail_pc_.w, 0, (size_t)16 * tail_pc_.k_total * sizeof(bf16), s));
          launch_pack_conv_weight(T_(l.name + ".block.3.weight"), tail_pc_.w, l.cout, l.c_x, 9, l.c_x, 0,
                                  tail_pc_.k_total, s);
        }
        break;
      }
      case LayerKind::Res: {
        const std::string rb = l.name + ".res_block";
        const int cin = l.c_x + l.c_skip;
        pack(l.name + ".c1", rb + ".block1.block.3", l.cout, cin, 9, "", 0, 0);
        // conv1's bias and the FeatureWiseAffine bias both land in the per-timestep table
        const int off = noise_off_[l.name];
        CUDA_CHECK(cudaMemcpyAsync(wall_ + (size_t)off * inner, T_(rb + ".noise_func.noise_func.0.weight"),
                                   (size_t)l.cout * inner * sizeof(float), cudaMemcpyDeviceToDevice, s));
        add_vec(T_(rb + ".noise_func.noise_func.0.bias"), T_(rb + ".block1.block.3.bias"), ball_ + off, l.cout, s);
        const bool has_res = cin != l.cout;
        if (!has_res) REQUIRE(l.c_skip == 0, "identity residual over a concatenated input is not supported");
        PackedConv* c2 = pack(l.name + ".c2", rb + ".block2.block.3", l.cout, l.cout, 9,
                              has_res ? rb + ".res_conv" : std::string("identity"), l.c_x, has_res ? l.c_skip : 0);
        c2->bias = (float*)dalloc((size_t)l.cout * sizeof(float));
        add_vec(T_(rb + ".block2.block.3.bias"), has_res ? T_(rb + ".res_conv.bias") : nullptr, c2->bias, l.cout, s);
        if (l.attn) {
          pack(l.name + ".qkv", l.name + ".attn.qkv", 3 * l.cout, l.cout, 1, "", 0, 0);
          PackedConv* o = pack(l.name + ".out", l.name + ".attn.out", l.cout, l.cout, 1, "identity", l.cout, 0);
          o->bias = T_(l.name + ".attn.out.bias");
        }
        break;
      }
    }
  }
  CUDA_CHECK(cudaStreamSynchronize(s));
  finalized_ = true;
  // the bias table depends on the weights: rebuild it if a schedule is already installed
  if (table_) {
    NoiseTablePlan np{nl_, T_("noise_level_mlp.1.weight"), T_("noise_level_mlp.1.bias"),
                      T_("noise_level_mlp.3.weight"), T_("noise_level_mlp.3.bias"), wall_, ball_, inner,
                      noise_total_, table_};
    launch_noise_table(np, 0, T_sched_, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
}

void Engine::set_schedule(int T, const float* a, const float* bc, const float* c1, const float* c2,
                          const float* lv, const double* sqrt_ac_prev, cudaStream_t s) {
  REQUIRE(T >= 1, "set_schedule: T must be positive");
  REQUIRE(a && bc && c1 && c2 && lv && sqrt_ac_prev, "set_schedule: null table");
  CUDA_CHECK(cudaSetDevice(device_));
  if (coefs_) cudaFree(coefs_);
  if (nl_) cudaFree(nl_);
  if (table_) cudaFree(table_);
  coefs_ = nl_ = table_ = nullptr;
  T_sched_ = T;
  std::vector<float> h((size_t)5 * T);
  const float* src[5] = {a, bc, c1, c2, lv};
  for (int k = 0; k < 5; ++k) memcpy(h.data() + (size_t)k * T, src[k], (size_t)T * sizeof(float));
  CUDA_CHECK(cudaMalloc(&coefs_, h.size() * sizeof(float)));
  CUDA_CHECK(cudaMemcpy(coefs_, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  // diffusion.py:166-167: noise_level for step t is float32(sqrt_alphas_cumprod_prev[t+1])
  std::vector<float> nl((size_t)T + 1, 0.f);
  for (int t = 0; t < T; ++t) nl[t] = (float)sqrt_ac_prev[t + 1];
  CUDA_CHECK(cudaMalloc(&nl_, nl.size() * sizeof(float)));
  CUDA_CHECK(cudaMemcpy(nl_, nl.data(), nl.size() * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMalloc(&table_, (size_t)(T + 1) * noise_total_ * sizeof(float)));
  workspaces_.clear();                 // plans captured the old table pointers
  if (finalized_) {
    NoiseTablePlan np{nl_, T_("noise_level_mlp.1.weight"), T_("noise_level_mlp.1.bias"),
                      T_("noise_level_mlp.3.weight"), T_("noise_level_mlp.3.bias"), wall_, ball_,
                      cfg_.inner_channel, noise_total_, table_};
    launch_noise_table(np, 0, T, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
}

int Engine::num_snapshots() const {
  if (T_sched_ <= 0) return 0;
  const int inter = 1 | (T_sched_ / 10);
  int n = 0;
  for (int t = 0; t < T_sched_; ++t) n += (t % inter == 0);
  return n;
}

// ------------------------------------------------------------------------------- workspace
Workspace& Engine::workspace(int B, int R) {
  REQUIRE(finalized_, "weights are not finalized (call b200sr3_finalize_weights)");
  REQUIRE(B >= 1 && R >= 1, "B and R must be positive");
  REQUIRE((R & (R - 1)) == 0, "R must be a power of two");
  REQUIRE((R >> (cfg_.n_mults - 1)) >= 1, "R is too small for the number of UNet levels");
  if (!table_) {          // no schedule yet: a one-row (scratch) table so unet_forward works
    T_sched_ = 0;
    CUDA_CHECK(cudaMalloc(&nl_, sizeof(float)));
    CUDA_CHECK(cudaMalloc(&table_, (size_t)noise_total_ * sizeof(float)));
  }
  for (size_t i = 0; i < workspaces_.size(); ++i)
    if (workspaces_[i]->B == B && workspaces_[i]->R == R) {      // hit: move to the most-recently-used end
      std::rotate(workspaces_.begin() + i, workspaces_.begin() + i + 1, workspaces_.end());
      return *workspaces_.back();
    }
  while (workspaces_.size() >= max_workspaces_) workspaces_.erase(workspaces_.begin());      // evict the LRU plan
  for (;;) {
    std::unique_ptr<Workspace> ws(new Workspace());
    ws->B = B;
    ws->R = R;
    try {
      build_workspace(*ws);
    } catch (const Error&) {
      // out of device memory: drop the other cached plans (oldest first) and retry before giving up
      ws.reset();
      cudaGetLastError();
      if (workspaces_.empty()) throw;
      workspaces_.erase(workspaces_.begin());
      continue;
    }
    workspaces_.push_back(std::move(ws));
    return *workspaces_.back();
  }
}

void Engine::build_workspace(Workspace& ws) {
  CUDA_CHECK(cudaSetDevice(device_));
  const int B = ws.B, R = ws.R;
  auto dalloc = [&](size_t bytes) {
    void* p = nullptr;
    bytes = (bytes + 255) / 256 * 256;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess)
      throw Error("workspace allocation of " + std::to_string(bytes) + " bytes failed after " +
                  std::to_string(ws.bytes) + ": " + cudaGetErrorString(e));
    ws.allocations.push_back(p);
    ws.bytes += bytes;
    return p;
  };
  // GroupNorm statistics. Default: ONE [B][C][2] int64 accumulator per tensor, carved out of one arena that the first op
  // of the step zeroes; the producing conv's CTAs add into it (ConvStats::atomic). B200SR3_STAT_SLOTS=1: the earlier
  // scheme - one slot per producing CTA, written not added, summed by the consumer (for A/B).
  static const bool stat_slots_mode = [] { const char* e = getenv("B200SR3_STAT_SLOTS"); return e && e[0] == '1'; }();
  const bool stat_atomic = !stat_slots_mode;
  size_t arena_cap = 0;
  {
    size_t ch = 4096;
    for (const LayerDesc& l : layers_) ch += (size_t)(l.kind == LayerKind::Res ? (l.attn ? 12 : 4) : 2) * std::max(l.cout, l.c_x + l.c_skip) + 512;
    arena_cap = ch * (size_t)B * 2 * sizeof(long long);
  }
  uint8_t* arena = stat_atomic ? (uint8_t*)dalloc(arena_cap) : nullptr;
  auto arena_used = std::make_shared<size_t>(0);
  if (stat_atomic) {
    ws.ops.push_back(Op{"stats.zero", false, [arena, arena_used](cudaStream_t s) {
      CUDA_CHECK(cudaMemsetAsync(arena, 0, *arena_used, s));
    }});
  }
  auto act = [&](int H, int W, int C, int stat_slots = 1) {
    Act a;
    a.B = B; a.H = H; a.W = W; a.C = C;
    a.ptr = (bf16*)dalloc(a.elems() * sizeof(bf16));
    if (stat_atomic) {
      const size_t bytes = ((size_t)B * C * 2 * sizeof(long long) + 255) / 256 * 256;
      REQUIRE(*arena_used + bytes <= arena_cap, "internal: statistics arena too small");
      a.stat_slots = 1;
      a.stats = reinterpret_cast<long long*>(arena + *arena_used);
      *arena_used += bytes;
    } else {
      a.stat_slots = stat_slots;
      a.stats = (long long*)dalloc((size_t)B * stat_slots * C * 2 * sizeof(long long));
    }
    return a;
  };
  const int oc = cfg_.out_channel;
  const size_t img = (size_t)B * oc * R * R * sizeof(float);
  ws.cond = (float*)dalloc(img);
  ws.x = (float*)dalloc(img);
  ws.eps = (float*)dalloc(img);
  CUDA_CHECK(cudaMemset(ws.cond, 0, img));
  const int G = cfg_.norm_groups;

  // statistics of a tensor no conv epilogue produced (head conv output, spatial sizes < 8x8)
  auto chan_stats = [&](const std::string& name, const Act& x) {
    auto g = std::make_shared<ChanStatsPlan>();
    g->src = x.ptr; g->B = B; g->HW = x.H * x.W; g->C = x.C;
    g->chunks = chan_stats_chunks(g->HW, g->C);
    g->chansum = x.stats;
    if (g->chunks > 1) {
      g->partial = (float*)dalloc((size_t)B * g->chunks * x.C * 2 * sizeof(float));
      g->ticket = (int*)dalloc((size_t)B * sizeof(int));
      CUDA_CHECK(cudaMemset(g->ticket, 0, (size_t)B * sizeof(int)));
    }
    ws.ops.push_back(Op{name + ".chan_stats", false, [g](cudaStream_t s) { launch_chan_stats(*g, s); }, 0.0,
                        2.0 * (double)B * g->HW * g->C});
  };
  // GroupNorm (+Swish) of [x0 | x1] into a fresh tensor (one HBM pass; statistics come with the inputs)
  auto group_norm = [&](const std::string& name, const Act& x0, const Act* x1, const std::string& gkey, bool swish) {
    auto g = std::make_shared<GnPlan>();
    g->src0 = x0.ptr; g->C0 = x0.C; g->stats0 = x0.stats; g->slots0 = x0.stat_slots;
    g->src1 = x1 ? x1->ptr : nullptr; g->C1 = x1 ? x1->C : 0; g->stats1 = x1 ? x1->stats : nullptr;
    g->slots1 = x1 ? x1->stat_slots : 1;
    g->B = B; g->HW = x0.H * x0.W; g->groups = G;
    const int C = g->C0 + g->C1;
    REQUIRE(C % G == 0, "GroupNorm: channels not divisible by norm_groups");
    g->gamma = T_(gkey + ".weight");
    g->beta = T_(gkey + ".bias");
    Act y = act(x0.H, x0.W, C);
    g->dst = y.ptr;
    g->swish = swish ? 1 : 0;
    const double elems = (double)B * g->HW * C;
    ws.ops.push_back(Op{name + ".gn_apply", false, [g](cudaStream_t s) { launch_gn_apply(*g, s); }, 0.0, 4.0 * elems});
    return y;
  };
  auto conv = [&](const std::string& name, const Act& src, int taps, int stride, bool up, const PackedConv& w,
                  const Act* r0, const Act* r1, const float* bias, int bias_stride, const bf16* residual,
                  int Ho, int Wo, bool want_stats) {
    Act probe;
    probe.B = B; probe.H = Ho; probe.W = Wo; probe.C = w.cout;
    const bool fuse = want_stats && fuse_stats_ && conv_can_fuse_stats(probe, up);
    Act y = act(Ho, Wo, w.cout, fuse ? conv_stat_slots(probe, up) : 1);
    ConvSource cs;
    cs.act = src; cs.taps = taps; cs.stride = stride; cs.upsample2x = up;
    ConvStats st;
    st.partial = y.stats;
    st.slots = y.stat_slots;
    st.atomic = stat_atomic;
    ws.ops.push_back(make_conv_op(name, cs, r0, r1, w, bias, bias_stride, ctl_, residual, y, force_block_n_,
                                  fuse ? &st : nullptr));
    ws.n_conv++;
    if (want_stats && !fuse) chan_stats(name, y);
    return y;
  };

  // GroupNorm whose apply (+Swish) runs inside the consuming halo conv. By default the conv also builds the
  // (scale, shift) table of its current image from the producers' partial sums (conv_halo.cuh); with
  // B200SR3_GN_TABLE_KERNEL=1 the table is a separate one-CTA-per-image launch (the earlier design, kept for A/B).
  struct GnRef { const float2* tab = nullptr; std::shared_ptr<GnPlan> plan; };
  static const bool table_kernel = [] { const char* e = getenv("B200SR3_GN_TABLE_KERNEL"); return e && e[0] == '1'; }();
  auto gn_table = [&](const std::string& name, const Act& x0, const Act* x1, const std::string& gkey) {
    auto g = std::make_shared<GnPlan>();
    g->C0 = x0.C; g->stats0 = x0.stats; g->slots0 = x0.stat_slots;
    g->C1 = x1 ? x1->C : 0; g->stats1 = x1 ? x1->stats : nullptr; g->slots1 = x1 ? x1->stat_slots : 1;
    g->B = B; g->HW = x0.H * x0.W; g->groups = G;
    const int C = g->C0 + g->C1;
    REQUIRE(C % G == 0, "GroupNorm: channels not divisible by norm_groups");
    g->gamma = T_(gkey + ".weight");
    g->beta = T_(gkey + ".bias");
    GnRef r;
    if (table_kernel || G > 32) {
      float2* tab = (float2*)dalloc((size_t)B * C * sizeof(float2));
      ws.ops.push_back(Op{name + ".gn_scale", false, [g, tab](cudaStream_t s) { launch_gn_scale_shift(*g, tab, s); }, 0.0, 0.0});
      r.tab = tab;
    } else {
      r.plan = g;
    }
    return r;
  };
  // halo-resident conv (conv_halo.cuh): 3x3 main source(s) with the GroupNorm+Swish applied in shared
  // memory, raw 1x1 shortcut sources, GroupNorm statistics of the output from the epilogue
  auto conv_halo = [&](const std::string& name, const std::vector<HaloSource>& srcs, bool up, const PackedConv& w,
                       const float* bias, int bias_stride, const GnRef& gn, int gn_C, bool want_stats, int stride = 1,
                       const float* affine_slope = nullptr) {      // non-null: GroupNorm WITHOUT Swish (slopes of 1)
    const Act& a0 = srcs[0].act;
    Act probe;
    probe.B = B; probe.H = up ? 2 * a0.H : a0.H / stride; probe.W = up ? 2 * a0.W : a0.W / stride; probe.C = w.cout;
    Act y = act(probe.H, probe.W, w.cout, want_stats ? conv_halo_stat_slots(probe, up) : 1);
    ConvStats st;
    st.partial = y.stats;
    st.slots = y.stat_slots;
    st.atomic = stat_atomic;
    HaloConvExtra extra;
    extra.gn_from_stats = gn.plan.get();
    extra.stride = stride;
    extra.prelu_slope = affine_slope;
    ws.ops.push_back(make_conv_halo_op(name, srcs, up, w, bias, bias_stride, ctl_, y, gn.tab, gn_C, affine_slope == nullptr,
                                       want_stats ? &st : nullptr, extra));
    ws.n_conv++;
    return y;
  };
  const GnRef no_gn;

  std::vector<Act> feats;
  Act cur;
  for (const LayerDesc& l : layers_) {
    switch (l.kind) {
      case LayerKind::HeadConv: {
        if (use_halo_ && head_pc_.w && conv_halo_eligible(R, R, 1, l.cout)) {
          // downs.0 on the tensor cores, straight from the fp32 NCHW sampler state: the kernel's transform warps build
          // the split-precision operand (x_t is never rounded to bf16) in shared memory - nothing is packed in HBM.
          // (B200SR3_HEAD_PACK=1: the earlier two-launch form, pack kernel + plain halo conv, for A/B.)
          const float* cond = cfg_.conditional ? ws.cond : nullptr;
          const int cc = cfg_.conditional ? oc : 0;
          static const bool head_pack = [] { const char* e = getenv("B200SR3_HEAD_PACK"); return e && e[0] == '1'; }();
          if (head_pack && (l.c_x == 2 || l.c_x == 6 || l.c_x == 8)) {
            Act hp = act(R, R, CONV_BLOCK_K);
            const float* xw = ws.x;
            bf16* hdst = hp.ptr;
            ws.ops.push_back(Op{l.name + ".pack", false, [=](cudaStream_t s) {
              launch_head_pack(cond, xw, cc, oc, B, R, hdst, s);
            }});
            cur = conv_halo(l.name, {HaloSource{hp, 9, -1}}, false, head_pc_, T_(l.name + ".bias"), 0, no_gn, 0, true);
          } else {
            Act virt;                       // the operand the kernel builds on the fly: 64 channels, never in memory
            virt.B = B; virt.H = R; virt.W = R; virt.C = CONV_BLOCK_K;
            HaloHead hh;
            hh.cond = cond; hh.x = ws.x; hh.cc = cc; hh.cx = oc;
            cur = act(R, R, l.cout, conv_halo_stat_slots(Act{nullptr, nullptr, 1, B, R, R, l.cout}, false));
            ConvStats st;
            st.partial = cur.stats;
            st.slots = cur.stat_slots;
            st.atomic = stat_atomic;
            HaloConvExtra extra;
            extra.head = &hh;
            ws.ops.push_back(make_conv_halo_op(l.name, {HaloSource{virt, 9, -1}}, false, head_pc_, T_(l.name + ".bias"), 0,
                                               ctl_, cur, nullptr, 0, true, &st, extra));
            ws.n_conv++;
          }
          ws.ops.back().flops = 2.0 * B * R * R * (double)l.cout * 9.0 * l.c_x;      // reference graph: K = 9 * in_channel
          // HBM-bound: algorithmic bytes of the whole head = the fp32 NCHW inputs in, the bf16 NHWC activation out
          ws.ops.back().bytes = (double)B * R * R * ((double)l.c_x * 4.0 + (double)l.cout * 2.0);
          feats.push_back(cur);
          break;
        }
        cur = act(R, R, l.cout);
        const float* cond = cfg_.conditional ? ws.cond : nullptr;
        const int cc = cfg_.conditional ? oc : 0;
        const float *xw = ws.x, *hw = head_w_, *hb = T_(l.name + ".bias");
        bf16* dst = cur.ptr;
        const int co = l.cout;
        long long* hstats = (fuse_stats_ && head_conv_fast_path(cc + oc, R, co)) ? cur.stats : nullptr;
        ws.ops.push_back(Op{l.name, false, [=](cudaStream_t s) {
          launch_head_conv(cond, xw, cc, oc, hw, hb, B, R, co, dst, hstats, s);
        }});
        if (!hstats) chan_stats(l.name, cur);
        feats.push_back(cur);
        break;
      }
      case LayerKind::Down: {
        const PackedConv& pc = convs_.at(l.name + ".conv");
        static const bool down_umma = [] { const char* e = getenv("B200SR3_DOWN_UMMA"); return e && e[0] == '1'; }();
        if (use_halo_ && !down_umma && convs_.count(l.name + ".conv.s2") &&
            conv_halo_eligible(cur.H / 2, cur.W / 2, cur.C % 64 == 0, pc.cout)) {
          // Downsample (unet.py:68-74) on the halo kernel: four input-parity views, 4 + 2 + 2 + 1 taps
          const PackedConv& s2 = convs_.at(l.name + ".conv.s2");
          cur = conv_halo(l.name, {HaloSource{cur, 9, -1}}, false, s2, s2.bias, 0, no_gn, 0, true, 2);
        } else {
          cur = conv(l.name, cur, 9, 2, false, pc, nullptr, nullptr, pc.bias, 0, nullptr, cur.H / 2, cur.W / 2, true);
        }
        feats.push_back(cur);
        break;
      }
      case LayerKind::Up: {
        const PackedConv& pc = convs_.at(l.name + ".conv");
        if (use_halo_ && conv_halo_eligible(cur.H, cur.W, cur.C % 64 == 0, pc.cout))
          cur = conv_halo(l.name, {HaloSource{cur, 9, -1}}, true, pc, pc.bias, 0, no_gn, 0, true);
        else
          cur = conv(l.name, cur, 9, 1, true, pc, nullptr, nullptr, pc.bias, 0, nullptr, cur.H * 2, cur.W * 2, true);
        break;
      }
      case LayerKind::Res: {
        const bool is_up = l.c_skip > 0;
        Act skip;
        if (is_up) {
          REQUIRE(!feats.empty(), "internal: skip stack underflow");
          skip = feats.back();
          feats.pop_back();
          REQUIRE(skip.C == l.c_skip && skip.H == cur.H, "internal: skip tensor mismatch");
        }
        REQUIRE(cur.C == l.c_x, "internal: channel plan mismatch");
        const std::string rb = l.name + ".res_block";
        const Act xin = cur;
        const PackedConv& c1 = convs_.at(l.name + ".c1");
        const PackedConv& c2 = convs_.at(l.name + ".c2");
        if (use_halo_ && conv_halo_eligible(xin.H, xin.W, (l.c_x % 64 == 0) && (l.c_skip % 64 == 0), l.cout)) {
          // Block = GN -> Swish -> Conv (unet.py:80-91) as ONE kernel each: the GroupNorm apply runs on the
          // halo tile in shared memory; only the tiny (scale, shift) table is a separate launch
          const int cin = l.c_x + l.c_skip;
          const GnRef g1 = gn_table(l.name + ".block1", xin, is_up ? &skip : nullptr, rb + ".block1.block.0");
          std::vector<HaloSource> s1{HaloSource{xin, 9, 0}};
          if (is_up) s1.push_back(HaloSource{skip, 9, l.c_x});
          Act h = conv_halo(l.name + ".conv1", s1, false, c1, table_ + noise_off_.at(l.name), noise_total_, g1, cin,
                            true);
          const GnRef g2 = gn_table(l.name + ".block2", h, nullptr, rb + ".block2.block.0");
          // the shortcut (res_conv 1x1, or identity) is one or two extra raw 1x1 K segments of this GEMM
          std::vector<HaloSource> s2{HaloSource{h, 9, 0}, HaloSource{xin, 1, -1, cin == l.cout}};      // (identity shortcut when dim == dim_out)
          if (is_up) s2.push_back(HaloSource{skip, 1, -1});
          cur = conv_halo(l.name + ".conv2", s2, false, c2, c2.bias, 0, g2, l.cout, true);
        } else {
          Act xn = group_norm(l.name + ".block1", xin, is_up ? &skip : nullptr, rb + ".block1.block.0", true);
          Act h = conv(l.name + ".conv1", xn, 9, 1, false, c1, nullptr, nullptr, table_ + noise_off_.at(l.name),
                       noise_total_, nullptr, xin.H, xin.W, true);
          Act hn = group_norm(l.name + ".block2", h, nullptr, rb + ".block2.block.0", true);
          // the shortcut (res_conv 1x1, or identity) is one or two extra 1x1 K segments of this GEMM
          cur = conv(l.name + ".conv2", hn, 9, 1, false, c2, &xin, is_up ? &skip : nullptr, c2.bias, 0, nullptr, xin.H,
                     xin.W, true);
        }
        if (l.attn) {
          const Act ain = cur;
          const PackedConv& pq = convs_.at(l.name + ".qkv");
          const PackedConv& po = convs_.at(l.name + ".out");
          static const bool attn_umma = [] { const char* e = getenv("B200SR3_ATTN_UMMA"); return e && e[0] == '1'; }();
          GnRef ga;
          const bool on_halo = use_halo_ && !attn_umma && !(ain.H == 4 && ain.W == 4) && ain.C <= 1024 &&
                               conv_halo_eligible(ain.H, ain.W, ain.C % 64 == 0, pq.cout);
          if (on_halo) ga = gn_table(l.name + ".attn", ain, nullptr, l.name + ".attn.norm");
          if (on_halo && ga.plan) {
            // SelfAttention (unet.py:113-142) with its two 1x1 convs on the halo kernel: attn.norm (GroupNorm, no Swish)
            // is applied to the qkv conv's input tile in shared memory (the affine transform with slopes of 1), the
            // residual `+ x` rides the out conv's GEMM as an identity K segment. No normalised tensor in HBM.
            Act qkv = conv_halo(l.name + ".attn.qkv", {HaloSource{ain, 1, 0}}, false, pq, nullptr, 0, ga, ain.C, false, 1, ones_);
            Act ao = act(ain.H, ain.W, ain.C);
            ws.ops.push_back(Op{l.name + ".attn.core", false, [qkv, ao](cudaStream_t s) {
              launch_attention(qkv.ptr, ao.ptr, qkv.B, qkv.H * qkv.W, ao.C, s);
            }});
            cur = conv_halo(l.name + ".attn.out", {HaloSource{ao, 1, -1}, HaloSource{ain, 1, -1, true}}, false, po, po.bias, 0,
                            no_gn, 0, true);
          } else {
            Act an = group_norm(l.name + ".attn", ain, nullptr, l.name + ".attn.norm", false);
            Act qkv = conv(l.name + ".attn.qkv", an, 1, 1, false, pq, nullptr, nullptr, nullptr, 0, nullptr, ain.H, ain.W, false);
            Act ao = act(ain.H, ain.W, ain.C);
            ws.ops.push_back(Op{l.name + ".attn.core", false, [qkv, ao](cudaStream_t s) {
              launch_attention(qkv.ptr, ao.ptr, qkv.B, qkv.H * qkv.W, ao.C, s);
            }});
            cur = conv(l.name + ".attn.out", ao, 1, 1, false, po, &ain, nullptr, po.bias, 0, nullptr, ain.H, ain.W, true);
          }
        }
        if (l.name.compare(0, 6, "downs.") == 0) feats.push_back(cur);
        break;
      }
      case LayerKind::Final: {
        if (use_halo_ && tail_pc_.w && conv_halo_eligible(cur.H, cur.W, cur.C % 64 == 0, 64)) {
          // final_conv = Block(GN -> Swish -> Conv 64 -> 3) + the sampler update as ONE halo conv launch
          const GnRef g = gn_table(l.name, cur, nullptr, l.name + ".block.0");
          Act o16;
          o16.B = B; o16.H = cur.H; o16.W = cur.W; o16.C = 16;
          HaloTail tl;
          tl.x = ws.x; tl.eps_out = nullptr; tl.coefs = coefs_; tl.oc = oc;
          HaloConvExtra extra;
          extra.tail = &tl;
          extra.params_out = &ws.tail_halo;
          extra.gn_from_stats = g.plan.get();
          ws.ops.push_back(make_conv_halo_op(l.name + ".tail", {HaloSource{cur, 9, 0}}, false, tail_pc_,
                                             T_(l.name + ".block.3.bias"), 0, ctl_, o16, g.tab, cur.C, true, nullptr, extra));
          // HBM-bound: the bf16 activation in, the fp32 state read and written (Philox noise costs no bytes)
          ws.ops.back().bytes = (double)B * R * R * ((double)cur.C * 2.0 + (double)oc * 8.0);
          ws.n_conv++;
          break;
        }
        Act fn = group_norm(l.name, cur, nullptr, l.name + ".block.0", true);
        auto tp = std::make_shared<TailPlan>();
        tp->src = fn.ptr; tp->w = tail_w_; tp->bias = T_(l.name + ".block.3.bias");
        tp->B = B; tp->R = R; tp->C = fn.C; tp->OC = oc;
        tp->eps_out = nullptr; tp->x = ws.x; tp->coefs = coefs_; tp->ctl = ctl_;
        ws.tail_plan = tp;
        ws.ops.push_back(Op{l.name + ".tail", false, [tp](cudaStream_t s) { launch_tail(*tp, s); }});
        break;
      }
    }
    if (l.kind != LayerKind::Final) ws.layer_out[l.name] = cur;
  }
  // every halo conv prefetches the next conv's weights into L2 (the last one the first one's: the chain repeats per step)
  static const bool no_prefetch = [] { const char* e = getenv("B200SR3_NO_WEIGHT_PREFETCH"); return e && e[0] == '1'; }();
  if (!no_prefetch) {
    const size_t n = ws.ops.size();
    for (size_t i = 0; i < n; ++i) {
      if (!ws.ops[i].set_prefetch) continue;
      for (size_t d = 1; d <= n; ++d) {
        const Op& nx = ws.ops[(i + d) % n];
        if (nx.weights) { ws.ops[i].set_prefetch(nx.weights, nx.weight_bytes); break; }
      }
    }
  }
}

void Engine::write_ctl(int t, int mode, const float* noise, uint64_t seed, long long numel, long long row0,
                       cudaStream_t s, bool clip) {
  StepCtl c;
  c.t = t; c.T = T_sched_; c.noise_mode = mode; c.no_clip = clip ? 0 : 1;
  c.noise = noise; c.seed = seed; c.numel = numel; c.row0 = row0;
  CUDA_CHECK(cudaMemcpyAsync(ctl_, &c, sizeof(c), cudaMemcpyHostToDevice, s));
}

void Engine::run_ops(Workspace& ws, cudaStream_t s) {
  for (auto& op : ws.ops) op.run(s);
  last_total += ws.n_kernels();
  last_conv += ws.n_conv;
}

void Engine::ensure_graph(Workspace& ws) {
  if (ws.graph || !use_graph_) return;
  // The tail must be in "update" mode while capturing (plans are read at launch = capture time).
  ws.set_tail(ws.x, nullptr);
  cudaGraph_t g = nullptr;
  CUDA_CHECK(cudaStreamBeginCapture(capture_stream_, cudaStreamCaptureModeThreadLocal));
  try {
    for (auto& op : ws.ops) op.run(capture_stream_);
    launch_ctl_advance(ctl_, capture_stream_);
  } catch (...) {
    cudaStreamEndCapture(capture_stream_, &g);
    if (g) cudaGraphDestroy(g);
    throw;
  }
  CUDA_CHECK(cudaStreamEndCapture(capture_stream_, &g));
  cudaError_t e = cudaGraphInstantiate(&ws.graph, g, 0);
  cudaGraphDestroy(g);
  CUDA_CHECK(e);
}

int Engine::profile_step(int B, int R, int max_ops, float* ms, double* flops, double* flops_executed, double* bytes,
                         char* names, int names_len, cudaStream_t s) {
  CUDA_CHECK(cudaSetDevice(device_));
  REQUIRE(T_sched_ > 0, "profile_step: no noise schedule installed");
  Workspace& ws = workspace(B, R);
  const int n = (int)ws.ops.size();
  REQUIRE(n <= max_ops, "profile_step: output arrays too small");
  const size_t numel = (size_t)B * cfg_.out_channel * R * R;
  launch_philox_fill(ws.x, B, cfg_.out_channel, R, 1234, T_sched_, 0, s);
  write_ctl(T_sched_ - 1, B200SR3_NOISE_PHILOX, nullptr, 1234, (long long)numel, 0, s);
  ws.set_tail(ws.x, nullptr);
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) CUDA_CHECK(cudaEventCreate(&e));
  for (int rep = 0; rep < 2; ++rep) {          // first repetition warms up
    CUDA_CHECK(cudaEventRecord(ev[0], s));
    for (int i = 0; i < n; ++i) {
      ws.ops[i].run(s);
      CUDA_CHECK(cudaEventRecord(ev[i + 1], s));
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (rep == 1)
      for (auto& op : ws.ops)
        if (op.report) op.report();      // timing build: role counters (accumulated "waits" are per super tile either way)
  }
  std::string all;
  for (int i = 0; i < n; ++i) {
    CUDA_CHECK(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
    if (flops) flops[i] = ws.ops[i].flops;
    if (flops_executed) flops_executed[i] = ws.ops[i].flops_executed;
    if (bytes) bytes[i] = ws.ops[i].bytes;
    all += ws.ops[i].name;
    all += '\n';
  }
  for (auto& e : ev) cudaEventDestroy(e);
  if (names && names_len > 0) {
    strncpy(names, all.c_str(), (size_t)names_len - 1);
    names[names_len - 1] = 0;
  }
  return n;
}

// ------------------------------------------------------------------------------- entry points
void Engine::unet_forward(const float* cond, const float* x, float noise_level, int B, int R, float* eps,
                          cudaStream_t s) {
  CUDA_CHECK(cudaSetDevice(device_));
  Workspace& ws = workspace(B, R);
  REQUIRE(x && eps, "unet_forward: null pointer");
  REQUIRE(cond || !cfg_.conditional, "unet_forward: cond is required for a conditional model");
  const size_t img = (size_t)B * cfg_.out_channel * R * R * sizeof(float);
  last_total = last_conv = 0;
  if (cond) CUDA_CHECK(cudaMemcpyAsync(ws.cond, cond, img, cudaMemcpyDeviceToDevice, s));
  CUDA_CHECK(cudaMemcpyAsync(ws.x, x, img, cudaMemcpyDeviceToDevice, s));
  // scratch row T of the bias table <- this noise level
  CUDA_CHECK(cudaMemcpyAsync(nl_ + T_sched_, &noise_level, sizeof(float), cudaMemcpyHostToDevice, s));
  NoiseTablePlan np{nl_, T_("noise_level_mlp.1.weight"), T_("noise_level_mlp.1.bias"),
                    T_("noise_level_mlp.3.weight"), T_("noise_level_mlp.3.bias"), wall_, ball_,
                    cfg_.inner_channel, noise_total_, table_};
  launch_noise_table(np, T_sched_, 1, s);
  write_ctl(T_sched_, 0, nullptr, 0, 0, 0, s);
  ws.set_tail(nullptr, ws.eps);
  run_ops(ws, s);
  ws.set_tail(ws.x, nullptr);
  CUDA_CHECK(cudaMemcpyAsync(eps, ws.eps, img, cudaMemcpyDeviceToDevice, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
}

void Engine::step(const float* cond, const float* x_t, const float* noise, int t, int clip_denoised, int B, int R,
                  float* x_tm1, cudaStream_t s) {
  CUDA_CHECK(cudaSetDevice(device_));
  REQUIRE(T_sched_ > 0, "step: no noise schedule installed (call b200sr3_set_schedule)");
  REQUIRE(t >= 0 && t < T_sched_, "step: t out of range");
  Workspace& ws = workspace(B, R);
  REQUIRE(x_t && x_tm1, "step: null pointer");
  REQUIRE(cond || !cfg_.conditional, "step: cond is required for a conditional model");
  const size_t img = (size_t)B * cfg_.out_channel * R * R * sizeof(float);
  last_total = last_conv = 0;
  if (cond) CUDA_CHECK(cudaMemcpyAsync(ws.cond, cond, img, cudaMemcpyDeviceToDevice, s));
  CUDA_CHECK(cudaMemcpyAsync(ws.x, x_t, img, cudaMemcpyDeviceToDevice, s));
  write_ctl(t, 3, noise, 0, (long long)(img / sizeof(float)), 0, s, clip_denoised != 0);
  ws.set_tail(ws.x, nullptr);
  run_ops(ws, s);
  CUDA_CHECK(cudaMemcpyAsync(x_tm1, ws.x, img, cudaMemcpyDeviceToDevice, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
}

void Engine::philox_normal(uint64_t seed, int t, int64_t row_offset, int B, int R, float* out, cudaStream_t s) {
  CUDA_CHECK(cudaSetDevice(device_));
  REQUIRE(out != nullptr && B >= 1 && R >= 1 && t >= 0 && row_offset >= 0, "philox_normal: bad argument");
  launch_philox_fill(out, B, cfg_.out_channel, R, seed, t, row_offset, s);
  CUDA_CHECK(cudaStreamSynchronize(s));
}

void Engine::sample(const float* cond, int noise_mode, const float* noise, uint64_t seed, int64_t row_offset, int B,
                    int R, float* out, float* snapshots, cudaStream_t s) {
  CUDA_CHECK(cudaSetDevice(device_));
  REQUIRE(T_sched_ > 0, "sample: no noise schedule installed (call b200sr3_set_schedule)");
  REQUIRE(noise_mode == B200SR3_NOISE_INJECTED || noise_mode == B200SR3_NOISE_PHILOX, "sample: bad noise_mode");
  REQUIRE(noise_mode != B200SR3_NOISE_INJECTED || noise != nullptr, "sample: injected mode needs a noise list");
  REQUIRE(out != nullptr, "sample: null output");
  REQUIRE(cond || !cfg_.conditional, "sample: cond is required for a conditional model");
  REQUIRE(row_offset >= 0, "sample: row_offset must be >= 0");
  Workspace& ws = workspace(B, R);
  const int T = T_sched_;
  const size_t numel = (size_t)B * cfg_.out_channel * R * R;
  const size_t img = numel * sizeof(float);
  last_total = last_conv = 0;
  ensure_graph(ws);
  if (cond) CUDA_CHECK(cudaMemcpyAsync(ws.cond, cond, img, cudaMemcpyDeviceToDevice, s));
  // x_T (diffusion.py:205): first entry of the injected list, or a Philox draw keyed at t = T
  if (noise_mode == B200SR3_NOISE_INJECTED) {
    CUDA_CHECK(cudaMemcpyAsync(ws.x, noise, img, cudaMemcpyDeviceToDevice, s));
  } else {
    // same generator as the update kernel, keyed at t = T (never a real step)
    launch_philox_fill(ws.x, B, cfg_.out_channel, R, seed, T, row_offset, s);
  }
  write_ctl(T - 1, noise_mode, noise, seed, (long long)numel, row_offset, s);
  ws.set_tail(ws.x, nullptr);
  const int inter = 1 | (T / 10);
  int snap = 0;
  for (int t = T - 1; t >= 0; --t) {
    if (ws.graph) {
      CUDA_CHECK(cudaGraphLaunch(ws.graph, s));
      last_total += ws.n_kernels() + 1;
      last_conv += ws.n_conv;
    } else {
      run_ops(ws, s);
      launch_ctl_advance(ctl_, s);
      last_total += 1;
    }
    if (snapshots && t % inter == 0)
      CUDA_CHECK(cudaMemcpyAsync(snapshots + (size_t)(snap++) * numel, ws.x, img, cudaMemcpyDeviceToDevice, s));
  }
  CUDA_CHECK(cudaMemcpyAsync(out, ws.x, img, cudaMemcpyDeviceToDevice, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
}

void Engine::sample_host(const float* cond_host, uint64_t seed, int64_t row_offset, int B, int R, float* out_host,
                         cudaStream_t s) {
  CUDA_CHECK(cudaSetDevice(device_));
  Workspace& ws = workspace(B, R);
  const size_t img = (size_t)B * cfg_.out_channel * R * R * sizeof(float);
  REQUIRE(out_host != nullptr, "sample_host: null output");
  REQUIRE(cond_host || !cfg_.conditional, "sample_host: cond is required for a conditional model");
  if (cond_host) CUDA_CHECK(cudaMemcpyAsync(ws.eps, cond_host, img, cudaMemcpyHostToDevice, s));
  // ws.eps doubles as the staging buffer: sample() copies cond -> ws.cond and out <- ws.x
  sample(cond_host ? ws.eps : nullptr, B200SR3_NOISE_PHILOX, nullptr, seed, row_offset, B, R, ws.eps, nullptr, s);
  CUDA_CHECK(cudaMemcpyAsync(out_host, ws.eps, img, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
}

void Engine::layer_output(const std::string& layer, float* dst, int* C, int* H, int* W, cudaStream_t s) {
  CUDA_CHECK(cudaSetDevice(device_));
  for (auto wi = workspaces_.rbegin(); wi != workspaces_.rend(); ++wi) {      // most recently used plan first
    Workspace& ws = **wi;
    auto it = ws.layer_out.find(layer);
    if (it == ws.layer_out.end()) continue;
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
  throw Error("layer_output: no activation named '" + layer + "' (run a forward first)");
}

}  // namespace b200sr3
