// extern "C" surface declared in include/b200sr3.h. Exceptions never cross the boundary: they
// are turned into a non-zero return code plus a thread-local message.
#include <cstdio>
#include <cstring>
#include <string>

#include "arcface.cuh"
#include "engine.cuh"

using namespace b200sr3;

struct b200sr3_handle {
  Engine* engine;
};
struct b200sr3_mica {
  MicaEncoder* enc;
};

static thread_local std::string g_last_error;

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    cudaGetLastError();   // clear a sticky-less launch error so the next call starts clean
    return 1;
  } catch (...) {
    g_last_error = "unknown error";
    return 2;
  }
}

static Engine& E(b200sr3_handle* h) {
  if (!h || !h->engine) throw Error("null handle");
  return *h->engine;
}

extern "C" {

const char* b200sr3_last_error(void) { return g_last_error.c_str(); }
int b200sr3_abi_version(void) { return B200SR3_ABI_VERSION; }

int b200sr3_create(const b200sr3_config* cfg, int device, b200sr3_handle** out) {
  return guarded([&] {
    REQUIRE(cfg && out, "create: null argument");
    *out = nullptr;
    Engine* e = new Engine(*cfg, device);
    *out = new b200sr3_handle{e};
  });
}

int b200sr3_destroy(b200sr3_handle* h) {
  return guarded([&] {
    if (!h) return;
    delete h->engine;
    delete h;
  });
}

int b200sr3_num_tensors(b200sr3_handle* h) {
  int n = -1;
  guarded([&] { n = E(h).num_tensors(); });
  return n;
}

int b200sr3_tensor_info(b200sr3_handle* h, int index, const char** key, int64_t shape[4], int* ndim) {
  return guarded([&] {
    const TensorSpec& t = E(h).tensor(index);
    if (key) *key = t.key.c_str();
    if (ndim) *ndim = (int)t.shape.size();
    if (shape)
      for (size_t i = 0; i < 4; ++i) shape[i] = i < t.shape.size() ? t.shape[i] : 1;
  });
}

int b200sr3_load_tensor(b200sr3_handle* h, const char* key, const float* data, const int64_t* shape, int ndim) {
  return guarded([&] {
    REQUIRE(key && shape, "load_tensor: null argument");
    E(h).load_tensor(key, data, shape, ndim);
  });
}

int b200sr3_finalize_weights(b200sr3_handle* h, void* stream) {
  return guarded([&] { E(h).finalize_weights((cudaStream_t)stream); });
}

int b200sr3_set_schedule(b200sr3_handle* h, int T, const float* sqrt_recip_ac, const float* sqrt_recipm1_ac,
                         const float* coef1, const float* coef2, const float* post_logvar,
                         const double* sqrt_ac_prev, void* stream) {
  return guarded([&] {
    E(h).set_schedule(T, sqrt_recip_ac, sqrt_recipm1_ac, coef1, coef2, post_logvar, sqrt_ac_prev, (cudaStream_t)stream);
  });
}

int b200sr3_unet_forward(b200sr3_handle* h, const float* cond, const float* x, float noise_level, int B, int R,
                         float* eps, void* stream) {
  return guarded([&] { E(h).unet_forward(cond, x, noise_level, B, R, eps, (cudaStream_t)stream); });
}

int b200sr3_step(b200sr3_handle* h, const float* cond, const float* x_t, const float* noise, int t, int clip_denoised,
                 int B, int R, float* x_tm1, void* stream) {
  return guarded([&] { E(h).step(cond, x_t, noise, t, clip_denoised, B, R, x_tm1, (cudaStream_t)stream); });
}

int b200sr3_sample(b200sr3_handle* h, const float* cond, int noise_mode, const float* noise, uint64_t seed,
                   int64_t row_offset, int B, int R, float* out, float* snapshots, void* stream) {
  return guarded([&] {
    E(h).sample(cond, noise_mode, noise, seed, row_offset, B, R, out, snapshots, (cudaStream_t)stream);
  });
}

int b200sr3_philox_normal(b200sr3_handle* h, uint64_t seed, int t, int64_t row_offset, int B, int R, float* out,
                          void* stream) {
  return guarded([&] { E(h).philox_normal(seed, t, row_offset, B, R, out, (cudaStream_t)stream); });
}

int b200sr3_num_snapshots(b200sr3_handle* h) {
  int n = -1;
  guarded([&] { n = E(h).num_snapshots(); });
  return n;
}

int b200sr3_sample_host(b200sr3_handle* h, const float* cond_host, uint64_t seed, int64_t row_offset, int B, int R,
                        float* out_host, void* stream) {
  return guarded([&] { E(h).sample_host(cond_host, seed, row_offset, B, R, out_host, (cudaStream_t)stream); });
}

int b200sr3_layer_output(b200sr3_handle* h, const char* layer, float* dst, int* C, int* H, int* W, void* stream) {
  return guarded([&] {
    REQUIRE(layer, "layer_output: null name");
    E(h).layer_output(layer, dst, C, H, W, (cudaStream_t)stream);
  });
}

int b200sr3_last_launch_count(b200sr3_handle* h, int64_t* total, int64_t* conv) {
  return guarded([&] {
    if (total) *total = E(h).last_total;
    if (conv) *conv = E(h).last_conv;
  });
}

int b200sr3_profile_step(b200sr3_handle* h, int B, int R, int max_ops, float* ms, double* flops, double* flops_executed,
                         double* bytes, char* names, int names_len, int* n_ops, void* stream) {
  return guarded([&] {
    REQUIRE(ms && n_ops, "profile_step: null argument");
    *n_ops = E(h).profile_step(B, R, max_ops, ms, flops, flops_executed, bytes, names, names_len, (cudaStream_t)stream);
  });
}

int b200sr3_conv2d(int device, const float* x, const float* w, const float* bias, const float* residual, int B,
                   int Cin, int H, int W, int Cout, int k, int stride, int upsample2x, float* y, int iters,
                   float* avg_ms, void* stream) {
  return guarded([&] {
    REQUIRE(x && w && y, "conv2d: null pointer");
    REQUIRE(k == 1 || k == 3, "conv2d: k must be 1 or 3");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw Error(std::string("no CUDA device available (") + cudaGetErrorString(e) + "): no CPU fallback");
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    REQUIRE(prop.major == 10, "conv2d: device is not sm_100");
    CUDA_CHECK(cudaSetDevice(device));
    conv_init_device();
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<void*> tmp;
    auto dalloc = [&](size_t bytes) {
      void* p = nullptr;
      CUDA_CHECK(cudaMalloc(&p, bytes));
      tmp.push_back(p);
      return p;
    };
    struct Cleanup {
      std::vector<void*>& v;
      ~Cleanup() { for (void* p : v) cudaFree(p); }
    } cleanup{tmp};

    Act in;
    in.B = B; in.H = H; in.W = W; in.C = Cin;
    in.ptr = (bf16*)dalloc(in.elems() * sizeof(bf16));
    launch_nchw_to_nhwc(x, in.ptr, B, Cin, H, W, s);
    Act src = in;
    Act out;
    out.B = B; out.C = Cout;
    out.H = (upsample2x ? 2 * H : H) / stride;
    out.W = (upsample2x ? 2 * W : W) / stride;
    out.ptr = (bf16*)dalloc(out.elems() * sizeof(bf16));
    bf16* res = nullptr;
    if (residual) {
      res = (bf16*)dalloc(out.elems() * sizeof(bf16));
      launch_nchw_to_nhwc(residual, res, B, Cout, out.H, out.W, s);
    }
    PackedConv pc;
    const int taps = k * k;
    const int cpad = (Cin + CONV_BLOCK_K - 1) / CONV_BLOCK_K * CONV_BLOCK_K;
    pc.cout = Cout; pc.taps = taps; pc.cin_main = Cin;
    if (upsample2x) {
      REQUIRE(k == 3 && stride == 1, "conv2d: upsample2x needs a 3x3 stride-1 conv");
      pc.up_folded = true;
      pc.k_total = 4 * cpad;
      pc.w = (bf16*)dalloc((size_t)4 * Cout * pc.k_total * sizeof(bf16));
      launch_pack_upfold_weight(w, pc.w, Cout, Cin, cpad, s);
    } else {
      pc.k_total = taps * cpad;
      pc.w = (bf16*)dalloc((size_t)Cout * pc.k_total * sizeof(bf16));
      launch_pack_conv_weight(w, pc.w, Cout, Cin, taps, cpad, 0, pc.k_total, s);
    }
    ConvSource cs;
    cs.act = src; cs.taps = taps; cs.stride = stride; cs.upsample2x = upsample2x != 0;
    int force_bn = 0;
    if (const char* g = getenv("B200SR3_BLOCK_N")) force_bn = atoi(g);
    // measurement switches: B200SR3_CONV2D_STATS=1 fuses the GroupNorm statistics into the epilogue
    // (as the engine does); B200SR3_CONV_TIMING=1 prints per-role cycle counters of the last launch.
    ConvStats st;
    const char* ev = getenv("B200SR3_CONV2D_STATS");
    const bool want_stats = ev && ev[0] == '1' && conv_can_fuse_stats(out, upsample2x != 0);
    if (want_stats) {
      st.slots = conv_stat_slots(out, upsample2x != 0);
      st.partial = (long long*)dalloc((size_t)B * st.slots * Cout * 2 * sizeof(long long));
    }
    ev = getenv("B200SR3_CONV_TIMING");
    const bool timing = ev && ev[0] == '1';
    if (timing) {
      st.dbg = (unsigned long long*)dalloc(256 * 8 * sizeof(unsigned long long));
      CUDA_CHECK(cudaMemset(st.dbg, 0, 256 * 8 * sizeof(unsigned long long)));
    }
    Op op = make_conv_op("conv2d", cs, nullptr, nullptr, pc, bias, 0, nullptr, res, out, force_bn,
                         (want_stats || timing) ? &st : nullptr);
    op.run(s);
    launch_nhwc_to_nchw(out.ptr, y, B, Cout, out.H, out.W, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (iters > 0 && avg_ms) {
      cudaEvent_t e0, e1;
      CUDA_CHECK(cudaEventCreate(&e0));
      CUDA_CHECK(cudaEventCreate(&e1));
      CUDA_CHECK(cudaEventRecord(e0, s));
      for (int i = 0; i < iters; ++i) op.run(s);
      CUDA_CHECK(cudaEventRecord(e1, s));
      CUDA_CHECK(cudaEventSynchronize(e1));
      float ms = 0.f;
      CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
      *avg_ms = ms / iters;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    if (timing) {
      std::vector<unsigned long long> h(256 * 8);
      CUDA_CHECK(cudaMemcpy(h.data(), st.dbg, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      double acc[8] = {0};
      int n = 0;
      for (int c = 0; c < 256; ++c) {
        if (h[c * 8 + 6] == 0) continue;
        ++n;
        for (int k = 0; k < 8; ++k) acc[k] += (double)h[c * 8 + k];
      }
      if (n) {
        const double tiles = acc[6] / n;
        fprintf(stderr,
                "conv timing (avg over %d CTAs, %.1f tiles each, cycles per tile): producer waits empty %.0f | "
                "mma waits full %.0f, waits tmem %.0f | epilogue waits accum %.0f, total %.0f\n",
                n, tiles, acc[0] / n / tiles, acc[2] / n / tiles, acc[3] / n / tiles, acc[4] / n / tiles,
                acc[5] / n / tiles);
      }
    }
  });
}


int b200sr3_conv_block(int device, const float* x0, int C0, const float* x1, int C1, const float* gamma,
                       const float* beta, int groups, int swish, const float* w, const float* bias, const float* r0,
                       int Cr0, const float* r1, int Cr1, const float* wres, int B, int H, int W, int Cout,
                       int resample, float* y, float* stats_out, int iters, float* avg_ms, void* stream) {
  return guarded([&] {
    REQUIRE(x0 && w && y, "conv_block: null pointer");
    REQUIRE((x1 != nullptr) == (C1 > 0) && (r0 != nullptr) == (Cr0 > 0) && (r1 != nullptr) == (Cr1 > 0),
            "conv_block: pointer / channel count mismatch");
    REQUIRE((Cr0 + Cr1 > 0) == (wres != nullptr), "conv_block: shortcut weights / sources mismatch");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw Error(std::string("no CUDA device available (") + cudaGetErrorString(e) + "): no CPU fallback");
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    REQUIRE(prop.major == 10, "conv_block: device is not sm_100");
    CUDA_CHECK(cudaSetDevice(device));
    conv_init_device();
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<void*> tmp;
    auto dalloc = [&](size_t bytes) {
      void* p = nullptr;
      CUDA_CHECK(cudaMalloc(&p, bytes));
      tmp.push_back(p);
      return p;
    };
    struct Cleanup {
      std::vector<void*>& v;
      ~Cleanup() { for (void* p : v) cudaFree(p); }
    } cleanup{tmp};
    REQUIRE(resample >= 0 && resample <= 2, "conv_block: resample is 0 (none), 1 (nearest 2x up) or 2 (stride 2)");
    const bool up = resample == 1, down = resample == 2;
    REQUIRE(!down || (H % 2 == 0 && W % 2 == 0), "conv_block: a stride-2 conv needs even H and W");
    const int TH = down ? H / 2 : H, TW = down ? W / 2 : W;      // the grid the tiles walk
    REQUIRE(conv_halo_eligible(TH, TW, (C0 % 64 == 0) && (C1 % 64 == 0) && (Cr0 % 64 == 0) && (Cr1 % 64 == 0), Cout),
            "conv_block: shape not supported by the halo conv (H % 16, W % 8, W >= 16, channels % 64)");
    REQUIRE(!up || (Cr0 + Cr1 == 0), "conv_block: a folded upsample has no shortcut");
    REQUIRE(!down || (Cr0 + Cr1 + C1 == 0 && gamma == nullptr), "conv_block: a stride-2 conv is one raw source");

    auto make_act = [&](const float* src, int C, bool want_stats) {
      Act a;
      a.B = B; a.H = H; a.W = W; a.C = C;
      a.ptr = (bf16*)dalloc(a.elems() * sizeof(bf16));
      launch_nchw_to_nhwc(src, a.ptr, B, C, H, W, s);
      if (want_stats) {
        a.stat_slots = 1;
        a.stats = (long long*)dalloc((size_t)B * C * 2 * sizeof(long long));
        ChanStatsPlan g;
        g.src = a.ptr; g.B = B; g.HW = H * W; g.C = C; g.chunks = 1; g.chansum = a.stats;
        launch_chan_stats(g, s);
      }
      return a;
    };
    const bool has_gn = gamma != nullptr;
    REQUIRE(!has_gn || beta != nullptr, "conv_block: GroupNorm needs gamma and beta");
    std::vector<HaloSource> srcs;
    Act a0 = make_act(x0, C0, has_gn);
    srcs.push_back(HaloSource{a0, 9, has_gn ? 0 : -1});
    Act a1;
    if (C1) { a1 = make_act(x1, C1, has_gn); srcs.push_back(HaloSource{a1, 9, has_gn ? C0 : -1}); }
    if (Cr0) srcs.push_back(HaloSource{make_act(r0, Cr0, false), 1, -1});
    if (Cr1) srcs.push_back(HaloSource{make_act(r1, Cr1, false), 1, -1});
    float2* gn = nullptr;
    GnPlan gplan;
    bool gn_in_kernel = false;
    if (has_gn) {
      gn = (float2*)dalloc((size_t)B * (C0 + C1) * sizeof(float2));
      GnPlan g;
      g.C0 = C0; g.C1 = C1; g.B = B; g.HW = H * W; g.groups = groups;
      g.stats0 = a0.stats; g.slots0 = 1; g.stats1 = C1 ? a1.stats : nullptr; g.slots1 = 1;
      float* dg = (float*)dalloc((size_t)(C0 + C1) * sizeof(float));
      float* db = (float*)dalloc((size_t)(C0 + C1) * sizeof(float));
      CUDA_CHECK(cudaMemcpyAsync(dg, gamma, (size_t)(C0 + C1) * sizeof(float), cudaMemcpyDefault, s));
      CUDA_CHECK(cudaMemcpyAsync(db, beta, (size_t)(C0 + C1) * sizeof(float), cudaMemcpyDefault, s));
      g.gamma = dg; g.beta = db;
      gplan = g;
      const char* tk = getenv("B200SR3_GN_TABLE_KERNEL");
      if (tk && tk[0] == '1') launch_gn_scale_shift(g, gn, s);      // the separate-launch table (A/B)
      else gn_in_kernel = true;                                      // default: the conv builds the table itself
    }
    Act out;
    out.B = B; out.C = Cout; out.H = up ? 2 * H : TH; out.W = up ? 2 * W : TW;
    out.ptr = (bf16*)dalloc(out.elems() * sizeof(bf16));
    PackedConv pc;
    const int cin = C0 + C1;
    pc.cout = Cout; pc.taps = 9; pc.cin_main = cin; pc.c_res0 = Cr0; pc.c_res1 = Cr1;
    if (up) {
      pc.up_folded = true;
      pc.k_total = 4 * cin;
      pc.w = (bf16*)dalloc((size_t)4 * Cout * pc.k_total * sizeof(bf16));
      launch_pack_upfold_weight(w, pc.w, Cout, cin, cin, s);
    } else if (down) {
      pc.down_perm = true;
      pc.k_total = 9 * cin;
      pc.w = (bf16*)dalloc((size_t)Cout * pc.k_total * sizeof(bf16));
      pack_conv_weight_by_input_parity(w, pc.w, Cout, cin, s);
    } else {
      pc.k_total = 9 * cin + Cr0 + Cr1;
      pc.w = (bf16*)dalloc((size_t)Cout * pc.k_total * sizeof(bf16));
      launch_pack_conv_weight(w, pc.w, Cout, cin, 9, cin, 0, pc.k_total, s);
      if (Cr0 + Cr1) launch_pack_conv_weight(wres, pc.w, Cout, Cr0 + Cr1, 1, Cr0 + Cr1, 9 * cin, pc.k_total, s);
    }
    ConvStats st;
    if (stats_out) {
      st.slots = conv_halo_stat_slots(out, up);
      st.partial = (long long*)dalloc((size_t)B * st.slots * Cout * 2 * sizeof(long long));
    }
    const char* tev = getenv("B200SR3_CONV_TIMING");
    bool timing = tev && tev[0] == '1';
#if !B200SR3_ROLE_TIMING
    if (timing) {
      fprintf(stderr, "b200sr3: this build has no role counters; build `build.py --timing` and set B200SR3_LIB to libb200sr3_timing.so\n");
      timing = false;
    }
#endif
    if (timing) {
      st.dbg = (unsigned long long*)dalloc(256 * 16 * sizeof(unsigned long long));
      CUDA_CHECK(cudaMemset(st.dbg, 0, 256 * 16 * sizeof(unsigned long long)));
    }
    HaloConvExtra extra;
    extra.gn_from_stats = gn_in_kernel ? &gplan : nullptr;
    extra.stride = down ? 2 : 1;
    Op op = make_conv_halo_op("conv_block", srcs, up, pc, bias, 0, nullptr, out, gn, cin, swish != 0,
                              (stats_out || timing) ? &st : nullptr, extra);
    op.run(s);
    launch_nhwc_to_nchw(out.ptr, y, B, Cout, out.H, out.W, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (stats_out) {
      std::vector<long long> h((size_t)B * st.slots * Cout * 2);
      CUDA_CHECK(cudaMemcpy(h.data(), st.partial, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      std::vector<float> f((size_t)B * Cout * 2);
      for (int b = 0; b < B; ++b)
        for (int c = 0; c < Cout * 2; ++c) {
          long long a = 0;
          for (int k = 0; k < st.slots; ++k) a += h[((size_t)b * st.slots + k) * Cout * 2 + c];
          f[(size_t)b * Cout * 2 + c] = (float)((double)a * STAT_FIXED_INV);
        }
      CUDA_CHECK(cudaMemcpy(stats_out, f.data(), f.size() * sizeof(float), cudaMemcpyDefault));
    }
    if (iters > 0 && avg_ms) {
      cudaEvent_t e0, e1;
      CUDA_CHECK(cudaEventCreate(&e0));
      CUDA_CHECK(cudaEventCreate(&e1));
      CUDA_CHECK(cudaEventRecord(e0, s));
      for (int i = 0; i < iters; ++i) op.run(s);
      CUDA_CHECK(cudaEventRecord(e1, s));
      CUDA_CHECK(cudaEventSynchronize(e1));
      float ms = 0.f;
      CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
      *avg_ms = ms / iters;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    if (timing) {
      halo_report_timing(st.dbg, nullptr);
    }
  });
}

int b200sr3_tensor2img(const float* x, int B, int C, int H, int W, uint8_t* img, void* stream) {
  return guarded([&] { launch_tensor2img(x, B, C, H, W, img, (cudaStream_t)stream); });
}

int b200sr3_mica_handoff(const uint8_t* img, int B, int R, uint8_t* up224, float* image224, float* arcface_blob,
                         void* stream) {
  return guarded([&] { launch_mica_handoff(img, B, R, up224, image224, arcface_blob, (cudaStream_t)stream); });
}

int b200sr3_tensor_blob(const float* x, int B, int R, float* arcface_blob, void* stream) {
  return guarded([&] { launch_tensor_blob(x, B, R, arcface_blob, (cudaStream_t)stream); });
}


// ---- MICA identity encoder (arcface.cu)
static MicaEncoder& M(b200sr3_mica* h) {
  if (!h || !h->enc) throw Error("null mica handle");
  return *h->enc;
}

int b200sr3_mica_create(int device, int z_dim, int map_hidden_dim, int map_layers, int n_shape, b200sr3_mica** out) {
  return guarded([&] {
    REQUIRE(out, "mica_create: null argument");
    *out = nullptr;
    MicaEncoder* e = new MicaEncoder(device, z_dim, map_hidden_dim, map_layers, n_shape);
    *out = new b200sr3_mica{e};
  });
}

int b200sr3_mica_destroy(b200sr3_mica* h) {
  return guarded([&] {
    if (!h) return;
    delete h->enc;
    delete h;
  });
}

int b200sr3_mica_num_tensors(b200sr3_mica* h) {
  int n = -1;
  guarded([&] { n = M(h).num_tensors(); });
  return n;
}

int b200sr3_mica_tensor_info(b200sr3_mica* h, int index, const char** key, int64_t shape[4], int* ndim) {
  return guarded([&] {
    const TensorSpec& t = M(h).tensor(index);
    if (key) *key = t.key.c_str();
    if (ndim) *ndim = (int)t.shape.size();
    if (shape)
      for (size_t i = 0; i < 4; ++i) shape[i] = i < t.shape.size() ? t.shape[i] : 1;
  });
}

int b200sr3_mica_load_tensor(b200sr3_mica* h, const char* key, const float* data, const int64_t* shape, int ndim) {
  return guarded([&] {
    REQUIRE(key && shape, "mica_load_tensor: null argument");
    M(h).load_tensor(key, data, shape, ndim);
  });
}

int b200sr3_mica_finalize_weights(b200sr3_mica* h, void* stream) {
  return guarded([&] { M(h).finalize_weights((cudaStream_t)stream); });
}

int b200sr3_mica_encode(b200sr3_mica* h, const float* arcface_blob, int B, float* embedding, float* identity,
                        float* shape_code, void* stream) {
  return guarded([&] { M(h).encode(arcface_blob, B, embedding, identity, shape_code, (cudaStream_t)stream); });
}

int b200sr3_mica_layer_output(b200sr3_mica* h, const char* layer, float* dst, int* C, int* H, int* W, void* stream) {
  return guarded([&] {
    REQUIRE(layer, "mica_layer_output: null name");
    M(h).layer_output(layer, dst, C, H, W, (cudaStream_t)stream);
  });
}

int b200sr3_mica_profile(b200sr3_mica* h, int B, int max_ops, float* ms, double* flops, char* names, int names_len,
                         int* n_ops, int64_t* launches, int64_t* conv_launches, void* stream) {
  return guarded([&] {
    REQUIRE(ms && n_ops, "mica_profile: null argument");
    *n_ops = M(h).profile(B, max_ops, ms, flops, names, names_len, (cudaStream_t)stream);
    if (launches) *launches = M(h).last_total;
    if (conv_launches) *conv_launches = M(h).last_conv;
  });
}

}  // extern "C"