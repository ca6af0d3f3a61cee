// Host side of the tcgen05 implicit-GEMM convolution: TMA descriptor construction, tile-shape
// choice and the launch closure. The kernel itself is in conv_umma.cuh.
#include <memory>
#include <mutex>

#include "engine.cuh"

namespace b200sr3 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is resolved at run time so that the library still loads (symbols only) on a host
// without a driver.
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  return fn;
}

// 4-D map over an NHWC bf16 activation: dims (C, W', H', B) with arbitrary element strides so
// that the stride-2 "parity" views can be expressed; box = (64, bw, bh, bb), 128B swizzle.
static void encode_act_map(CUtensorMap* m, const bf16* base, int C, int Wd, int Hd, int Bd, size_t sw,
                           size_t sh, size_t sb, int bw, int bh, int bb) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)Bd};
  cuuint64_t strides[3] = {(cuuint64_t)(sw * 2), (cuuint64_t)(sh * 2), (cuuint64_t)(sb * 2)};
  cuuint32_t box[4] = {(cuuint32_t)CONV_BLOCK_K, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw Error("cuTensorMapEncodeTiled(activation) failed with CUresult " + std::to_string((int)r));
}

static void encode_weight_map(CUtensorMap* m, const bf16* w, int k_total, int cout, int block_n) {
  cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)cout};
  cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
  cuuint32_t box[2] = {(cuuint32_t)CONV_BLOCK_K, (cuuint32_t)block_n};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(w), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw Error("cuTensorMapEncodeTiled(weights) failed with CUresult " + std::to_string((int)r));
}

// shared with conv_halo.cu
CUresult encode_tiled(CUtensorMap* m, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr,
                      CUtensorMapSwizzle sw, CUtensorMapL2promotion l2) {
  return encode_fn()(m, dt, rank, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

template <int BN, int ST, int MW>
static void launch_conv(const ConvParams& p, int grid, cudaStream_t s) {
  launch_pdl(conv_umma_kernel<BN, ST, MW>, dim3(grid), dim3(CONV_THREADS), ConvSmem<BN, ST>::TOTAL, s, p);
}

// smem ring depth per BLOCK_N: 192 KB of stages in every case
constexpr int ST64 = 8, ST128 = 6, ST256 = 4;
static int g_num_sms = 148;

void conv_init_device() {
  CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel<64, ST64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ConvSmem<64, ST64>::TOTAL));
  CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel<128, ST128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ConvSmem<128, ST128>::TOTAL));
  CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel<256, ST256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ConvSmem<256, ST256>::TOTAL));
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  conv_halo_init_device();
}

bool conv_can_fuse_stats(const Act& out, bool upsample2x) {
  // a warp's 32 accumulator rows must belong to one image, and a tile to at most two
  const int hw = upsample2x ? (out.H / 2) * (out.W / 2) : out.H * out.W;
  return hw >= 64;
}

Op make_conv_op(const std::string& name, const ConvSource& main, const Act* res0, const Act* res1,
                const PackedConv& w, const float* bias, int bias_t_stride, const StepCtl* ctl,
                const bf16* residual, const Act& out, int force_block_n, const ConvStats* stats) {
  const Act& a = main.act;
  const bool up = main.upsample2x;
  REQUIRE(main.stride == 1 || main.stride == 2, "conv: stride must be 1 or 2");
  REQUIRE(main.taps == 9 || main.taps == 1, "conv: kernel must be 1x1 or 3x3");
  REQUIRE(a.C % 8 == 0 && out.C % 8 == 0, "conv: channel counts must be multiples of 8");
  REQUIRE(a.C == w.cin_main && out.C == w.cout && main.taps == w.taps, "conv: weight/activation mismatch");
  if (up) {
    REQUIRE(main.stride == 1 && main.taps == 9 && !res0 && !res1 && w.up_folded,
            "conv: folded upsample supports a plain 3x3 stride-1 conv only");
    REQUIRE(out.H == 2 * a.H && out.W == 2 * a.W && out.B == a.B, "conv: output shape mismatch");
  } else {
    REQUIRE(out.H * main.stride == a.H && out.W * main.stride == a.W && out.B == a.B, "conv: output shape mismatch");
    REQUIRE(!w.up_folded, "conv: weights were packed for a folded upsample");
  }
  REQUIRE(is_pow2(out.W) && is_pow2(out.H), "conv: spatial dims must be powers of two");

  auto pp = std::make_shared<ConvParams>();
  ConvParams& p = *pp;
  memset(&p, 0, sizeof(p));
  // ---- tile box over the pixel space the tiles walk (the source grid when the upsample is folded)
  const int PH = up ? a.H : out.H, PW = up ? a.W : out.W;
  p.bw = std::min(PW, CONV_BLOCK_M);
  p.bh = std::min(PH, CONV_BLOCK_M / p.bw);
  p.bb = CONV_BLOCK_M / (p.bw * p.bh);
  p.tiles_w = PW / p.bw;
  p.tiles_h = PH / p.bh;
  p.tiles_b = ceil_div(out.B, p.bb);
  p.B = out.B; p.Hout = PH; p.Wout = PW; p.Cout = out.C;
  p.out_H = out.H; p.out_W = out.W;
  p.num_par = up ? 4 : 1;
  p.bias = bias; p.bias_t_stride = bias_t_stride; p.ctl = ctl;
  p.residual = residual; p.out = out.ptr;

  // ---- A maps + tap list (order must match the packed weight's K order)
  int nt = 0, kblocks = 0;
  const int cb_main = ceil_div(a.C, CONV_BLOCK_K);
  if (up) {
    encode_act_map(&p.a_map[0], a.ptr, a.C, a.W, a.H, a.B, (size_t)a.C, (size_t)a.W * a.C,
                   (size_t)a.H * a.W * a.C, p.bw, p.bh, p.bb);
    // output (2i+py, 2j+px) reads source rows {i-1, i} (py = 0) or {i, i+1} (py = 1); same for columns
    for (int par = 0; par < 4; ++par) {
      const int py = par >> 1, px = par & 1;
      for (int t = 0; t < 4; ++t) {
        ConvTap& tp = p.taps[nt++];
        tp.map = 0;
        tp.dh = (int16_t)((t >> 1) - 1 + py);
        tp.dw = (int16_t)((t & 1) - 1 + px);
        tp.cblocks = (int16_t)cb_main;
      }
    }
    p.num_taps = 4;
    kblocks = 4 * cb_main;
  } else if (main.stride == 1) {
    encode_act_map(&p.a_map[0], a.ptr, a.C, a.W, a.H, a.B, (size_t)a.C, (size_t)a.W * a.C,
                   (size_t)a.H * a.W * a.C, p.bw, p.bh, p.bb);
    for (int t = 0; t < main.taps; ++t) {
      ConvTap& tp = p.taps[nt++];
      tp.map = 0;
      tp.dh = main.taps == 9 ? (int16_t)(t / 3 - 1) : 0;
      tp.dw = main.taps == 9 ? (int16_t)(t % 3 - 1) : 0;
      tp.cblocks = (int16_t)cb_main;
      kblocks += cb_main;
    }
  } else {
    REQUIRE(main.taps == 9 && !res0 && !res1, "conv: stride 2 supports plain 3x3 only");
    // parity view (ph, pw): element (c, j, i, b) = src[b][2i+ph][2j+pw][c]
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw)
        encode_act_map(&p.a_map[ph * 2 + pw], a.ptr + ((size_t)ph * a.W + pw) * a.C, a.C, a.W / 2, a.H / 2, a.B,
                       (size_t)2 * a.C, (size_t)2 * a.W * a.C, (size_t)a.H * a.W * a.C, p.bw, p.bh, p.bb);
    for (int t = 0; t < 9; ++t) {
      const int kh = t / 3, kw = t % 3;       // input row = 2*ho + kh - 1
      ConvTap& tp = p.taps[nt++];
      tp.map = (int16_t)(((kh == 1) ? 0 : 1) * 2 + ((kw == 1) ? 0 : 1));
      tp.dh = (kh == 0) ? -1 : 0;
      tp.dw = (kw == 0) ? -1 : 0;
      tp.cblocks = (int16_t)cb_main;
      kblocks += cb_main;
    }
  }
  if (!up) {
    const Act* rs[2] = {res0, res1};
    const int rc[2] = {w.c_res0, w.c_res1};
    for (int i = 0; i < 2; ++i) {
      if (!rs[i]) { REQUIRE(rc[i] == 0, "conv: missing res_conv source"); continue; }
      REQUIRE(rs[i]->C == rc[i] && rs[i]->H == out.H && rs[i]->W == out.W && rs[i]->B == out.B,
              "conv: res_conv source mismatch");
      const Act& r = *rs[i];
      encode_act_map(&p.a_map[1 + i], r.ptr, r.C, r.W, r.H, r.B, (size_t)r.C, (size_t)r.W * r.C,
                     (size_t)r.H * r.W * r.C, p.bw, p.bh, p.bb);
      ConvTap& tp = p.taps[nt++];
      tp.map = (int16_t)(1 + i); tp.dh = 0; tp.dw = 0;
      tp.cblocks = (int16_t)ceil_div(r.C, CONV_BLOCK_K);
      kblocks += tp.cblocks;
    }
    p.num_taps = nt;
  }
  REQUIRE(nt <= CONV_MAX_TAPS, "conv: too many taps");
  REQUIRE(kblocks * CONV_BLOCK_K == w.k_total, "conv: packed weight K does not match the tap list");
  p.num_kblocks = kblocks;

  // ---- N tile: the candidate with the lowest modelled time. Per K block a tile costs
  // max(MMA cycles = 2*bn, operand bytes / L2-to-SM rate) and a CTA runs ceil(tiles / SMs) tiles.
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b * p.num_par;
  int bn = 64;
  if (force_block_n) {
    bn = force_block_n;
  } else {
    double best = 1e30;
    const int cand[3] = {256, 128, 64};
    for (int c : cand) {
      if (out.C % c != 0) continue;
      // tensor time per K block (measured: 128 cycles per N=256 MMA; with two issuing warps 64 per
      // N=128 and ~48 per N=64 MMA) vs operand bytes at ~96 B/cycle/SM from L2
      const double mma = c == 256 ? 512.0 : (c == 128 ? 256.0 : 192.0);
      const double per_kb = std::max(mma, (16384.0 + 128.0 * c) / 96.0);
      const double rounds = (double)ceil_div(m_tiles * (out.C / c), g_num_sms);
      const double cost = rounds * (kblocks * per_kb + 400.0 + 6.0 * c);
      if (cost < best) { best = cost; bn = c; }
    }
  }
  REQUIRE(bn == 64 || bn == 128 || bn == 256, "conv: BLOCK_N must be 64, 128 or 256");
  REQUIRE(out.C % bn == 0, "conv: BLOCK_N must divide Cout");
  p.tiles_n = out.C / bn;
  p.total_tiles = m_tiles * p.tiles_n;
  encode_weight_map(&p.w_map, w.w, w.k_total, w.cout * p.num_par, bn);

  const int grid = std::min(p.total_tiles, g_num_sms);
  p.seg_len = p.tiles_w * p.tiles_h * p.num_par;
  if (stats) p.dbg = stats->dbg;
  if (const char* ab = getenv("B200SR3_CONV_ABLATE")) p.ablate = atoi(ab);
  if (stats && stats->partial) {
    REQUIRE(conv_can_fuse_stats(out, up), "conv: statistics cannot be fused for this shape");
    REQUIRE(p.bb <= 2, "conv: internal tile shape error");
    // the most CTAs any segment is spread over (same owner formula as the kernel)
    int need = 1;
    const long long G = grid, T = p.total_tiles;
    for (long long seg = 0; seg * p.seg_len < T; ++seg) {
      const int first_cta = (int)(((seg * p.seg_len + 1) * G - 1) / T);
      const int last_cta = (int)((((seg + 1) * p.seg_len) * G - 1) / T);
      need = std::max(need, last_cta - first_cta + 1);
    }
    REQUIRE(stats->atomic ? stats->slots == 1 : need <= stats->slots, "conv: statistics scratch has too few slots");
    p.stat_partial = stats->partial;
    p.stat_slots = stats->slots;
    p.stat_atomic = stats->atomic ? 1 : 0;
  }

  Op op;
  op.name = name;
  op.is_conv = true;
  {
    const double m = (double)out.B * out.H * out.W;
    const double kres = (double)((res0 ? res0->C : 0) + (res1 ? res1->C : 0));
    // reference graph: full-resolution 3x3 for Upsample; a res_conv counts, an identity shortcut that
    // merely rides the GEMM (unet.py:101, nn.Identity) does not
    op.flops = 2.0 * m * (double)out.C * ((double)main.taps * a.C + (w.res_identity ? 0.0 : kres));
    op.flops_executed = 2.0 * m * (double)out.C * ((double)(up ? 4 : main.taps) * a.C + kres);
  }
  op.run = [pp, grid, bn](cudaStream_t s) {
    if (bn == 256) launch_conv<256, ST256, 1>(*pp, grid, s);
    else if (bn == 128) launch_conv<128, ST128, 2>(*pp, grid, s);
    else launch_conv<64, ST64, 2>(*pp, grid, s);
  };
  return op;
}

int conv_stat_slots(const Act& out, bool upsample2x) {
  // upper bound over every BLOCK_N choice: the fewest tiles (tiles_n = 1) give the shortest runs
  const int PH = upsample2x ? out.H / 2 : out.H, PW = upsample2x ? out.W / 2 : out.W;
  const int bw = std::min(PW, CONV_BLOCK_M);
  const int bh = std::min(PH, CONV_BLOCK_M / bw);
  const int bb = CONV_BLOCK_M / (bw * bh);
  const int seg_len = (PH / bh) * (PW / bw) * (upsample2x ? 4 : 1);
  const long long T = (long long)seg_len * ceil_div(out.B, bb);
  const long long G = std::min<long long>(T, g_num_sms);
  const long long min_run = std::max<long long>(1, T / G);
  return (int)((seg_len - 1) / min_run + 2);
}

}  // namespace b200sr3
