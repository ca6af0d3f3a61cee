// Host side of the halo-resident convolution (conv_halo.cuh): tensor maps, segment list, tile-shape
// choice, launch closure.
#include <cstdio>
#include <memory>
#include <mutex>
#include <vector>

#include "engine.cuh"

namespace b200sr3 {

CUresult encode_tiled(CUtensorMap* m, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr,
                      CUtensorMapSwizzle sw, CUtensorMapL2promotion l2);   // conv_umma.cu

static int g_halo_sms = 148;

template <int BN, int MT, bool GN, int GEO = 0, int CG = 1>
static void set_attr() {
  CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<BN, MT, GN, GEO, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  HaloSmem<BN, MT, GEO, CG>::TOTAL));
}

void conv_halo_init_device() {
  set_attr<64, 1, false>();  set_attr<64, 1, true>();
  set_attr<64, 2, false>();  set_attr<64, 2, true>();
  set_attr<128, 1, false>(); set_attr<128, 1, true>();
  set_attr<128, 2, false>(); set_attr<128, 2, true>();
  set_attr<256, 1, false>(); set_attr<256, 1, true>();
  set_attr<16, 1, false>();  set_attr<16, 1, true>();
  set_attr<16, 2, false>();  set_attr<16, 2, true>();
  set_attr<64, 1, false, 1>();  set_attr<64, 1, true, 1>();
  set_attr<128, 1, false, 1>(); set_attr<128, 1, true, 1>();
  set_attr<64, 1, false, 2>();  set_attr<64, 1, true, 2>();
  set_attr<64, 1, false, 0, 2>();  set_attr<64, 1, true, 0, 2>();
  set_attr<128, 1, false, 0, 2>(); set_attr<128, 1, true, 0, 2>();
  set_attr<256, 1, false, 0, 2>(); set_attr<256, 1, true, 0, 2>();
  CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<64, 1, true, 0, 1, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloSmem<64, 1, 0, 1>::TOTAL));
  CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<64, 2, true, 0, 1, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloSmem<64, 2, 0, 1>::TOTAL));
  CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<128, 1, true, 0, 1, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloSmem<128, 1, 0, 1>::TOTAL));
  CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<256, 1, true, 0, 1, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloSmem<256, 1, 0, 1>::TOTAL));
  CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<64, 1, true, 1, 1, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloSmem<64, 1, 1, 1>::TOTAL));
  CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<128, 1, true, 1, 1, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloSmem<128, 1, 1, 1>::TOTAL));
  CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<64, 1, true, 0, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  HaloSmem<64, 1, 0, 1, true>::TOTAL));
  CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<128, 1, true, 0, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  HaloSmem<128, 1, 0, 1, true>::TOTAL));
  CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<64, 1, false, 0, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  HaloSmem<64, 1, 0, 1>::TOTAL));
  CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<64, 2, false, 0, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  HaloSmem<64, 2, 0, 1>::TOTAL));
  int dev = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&g_halo_sms, cudaDevAttrMultiProcessorCount, dev));
}

template <int BN, int MT, int GEO = 0>
static void launch_halo(const ConvHaloParams& p, bool gn, int grid, cudaStream_t s) {
  if (gn) launch_pdl(conv_halo_kernel<BN, MT, true, GEO>, dim3(grid), dim3(halo_threads(BN)), HaloSmem<BN, MT, GEO>::TOTAL, s, p);
  else launch_pdl(conv_halo_kernel<BN, MT, false, GEO>, dim3(grid), dim3(halo_threads(BN)), HaloSmem<BN, MT, GEO>::TOTAL, s, p);
}
// CTA pairs (cta_group::2): a cluster of two CTAs per tile pair
template <int BN>
static void launch_halo_pair(const ConvHaloParams& p, bool gn, int grid, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(halo_threads(BN));
  cfg.dynamicSmemBytes = HaloSmem<BN, 1, 0, 2>::TOTAL;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (gn) CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_halo_kernel<BN, 1, true, 0, 2>, p));
  else CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_halo_kernel<BN, 1, false, 0, 2>, p));
}

// Role counters of one launch (timing build, conv_halo.cuh HALO_DBG): averages over the CTAs that ran tiles.
void halo_report_timing(const unsigned long long* dbg_dev, const char* label) {
  std::vector<unsigned long long> h(256 * 16);
  CUDA_CHECK(cudaMemcpy(h.data(), dbg_dev, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  double acc[16] = {0};
  int n = 0;
  for (int c = 0; c < 256; ++c) {
    if (h[c * 16 + 2] == 0) continue;
    ++n;
    for (int k = 0; k < 16; ++k) acc[k] += (double)h[c * 16 + k];
  }
  if (!n) return;
  const double sup = acc[2] / n;
  auto per = [&](int k) { return acc[k] / n / sup; };
  if (label) fprintf(stderr, "[%s] ", label);
  fprintf(stderr,
          "halo timing (avg over %d CTAs, %.1f super tiles each; cycles per super tile): total %.0f | A producer "
          "waits empty %.0f | MMA waits A %.0f, W %.0f, TMEM %.0f | epilogue waits accum %.0f | transform waits "
          "A full %.0f, works %.0f\n",
          n, sup, per(1), per(0), per(4), per(5), per(6), per(8), per(10), per(11));
  double mx14 = 0;
  for (int c = 0; c < 256; ++c) mx14 = std::max(mx14, (double)h[c * 16 + 14]);
  fprintf(stderr, "  per CTA (cycles): entry -> prologue done %.0f | -> GroupNorm table: first barrier passed %.0f, ready %.0f | -> first halo landed %.0f\n",
          acc[15] / n, acc[9] / n, acc[3] / n, acc[7] / n);
  fprintf(stderr, "  per CTA (cycles): entry -> first MMA %.0f | entry -> last MMA issued %.0f | entry -> epilogue done %.0f "
                  "(slowest CTA %.0f)\n", acc[12] / n, acc[13] / n, acc[14] / n, mx14);
}

static bool geo1(int H, int W) { return H == 8 && W == 8; }   // two whole 8x8 images per tile
static bool geo2(int H, int W) { return H == 4 && W == 4; }   // five whole 4x4 images per tile
static int geo_imgs(int H, int W) { return geo1(H, W) ? 2 : (geo2(H, W) ? 5 : 1); }

bool conv_halo_eligible(int H, int W, int c_multiple_of_64_all, int cout) {
  if (!c_multiple_of_64_all || cout % 64 != 0) return false;
  return geo1(H, W) || geo2(H, W) || ((H % HALO_TH == 0) && (W % HALO_TW == 0) && W >= 16 && H >= 16);
}

// CTAs a (n tile, image) segment of seg_len_super super tiles can be spread over when total_super
// super tiles are split into contiguous runs over `grid` CTAs (same owner formula as the kernel).
static int slots_needed(long long seg_len_super, long long total_super, long long grid) {
  int need = 1;
  for (long long seg = 0; seg * seg_len_super < total_super; ++seg) {
    const int first_cta = (int)(((seg * seg_len_super + 1) * grid - 1) / total_super);
    const int last_cta = (int)((((seg + 1) * seg_len_super) * grid - 1) / total_super);
    need = std::max(need, last_cta - first_cta + 1);
  }
  return need;
}

int conv_halo_stat_slots(const Act& out, bool upsample2x) {
  const int PH = upsample2x ? out.H / 2 : out.H, PW = upsample2x ? out.W / 2 : out.W;
  const int gi = geo_imgs(PH, PW);
  const bool g1 = gi > 1;
  const long long seg_len = (g1 ? 1 : (long long)(PH / HALO_TH) * (PW / HALO_TW)) * (upsample2x ? 4 : 1);
  const long long units = (out.B + gi - 1) / gi;
  int need = 1;
  for (int mt = 1; mt <= 2; ++mt) {
    if (seg_len % mt) continue;
    for (int tn = 1; tn <= 16; tn *= 2) {
      const long long total = seg_len / mt * units * tn;
      need = std::max(need, slots_needed(seg_len / mt, total, std::min<long long>(total, g_halo_sms)));
      // CTA pairs: super tile = the pair's two tiles, one run per cluster, two slots per cluster
      if (mt == 2) need = std::max(need, 2 * slots_needed(seg_len / 2, total, std::min<long long>(total, g_halo_sms / 2)));
    }
  }
  return need;
}

Op make_conv_halo_op(const std::string& name, const std::vector<HaloSource>& srcs, bool upsample2x,
                     const PackedConv& w, const float* bias, int bias_t_stride, const StepCtl* ctl, const Act& out,
                     const float2* gn, int gn_C, bool gn_swish, const ConvStats* stats, const HaloConvExtra& extra) {
  const HaloTail* tail = extra.tail;
  std::shared_ptr<ConvHaloParams>* params_out = extra.params_out;
  const GnPlan* gn_from_stats = extra.gn_from_stats;
  const int stride = extra.stride;
  const HaloHead* head = extra.head;
  const float* prelu_slope = extra.prelu_slope;
  const bool partial_tiles = extra.partial_tiles;
  REQUIRE(!srcs.empty() && (int)srcs.size() <= HALO_MAX_SEGS, "halo conv: 1..4 sources");
  const Act& a0 = srcs[0].act;
  REQUIRE(stride == 1 || stride == 2, "halo conv: stride 1 or 2");
  // stride 2 (Downsample, unet.py:68-74): the tiles walk the OUTPUT grid and the input is read through its four
  // (row parity, column parity) sub-grids, each a strided view with its own tensor map - see the segment list below
  REQUIRE(stride == 1 || (srcs[0].ntaps == 9 && !upsample2x && !tail && a0.H % 2 == 0 && a0.W % 2 == 0 && w.down_perm &&
                          srcs.size() <= 3),
          "halo conv: a stride-2 conv is one 3x3 source with parity-ordered weights (plus at most two 1x1 shortcut sources)");
  const bool prelu = prelu_slope != nullptr;      // transform = per-channel affine + PReLU (ArcFace), not GroupNorm + Swish
  // (scale, shift) of the affine + PReLU transform: a ready row (`gn`, ArcFace's folded BatchNorm), or - `gn_from_stats` -
  // the GroupNorm table the kernel builds per image, here WITHOUT the Swish (attn.norm, unet.py:117,127: slopes all 1)
  REQUIRE(!prelu || (((gn != nullptr) != (gn_from_stats != nullptr)) && !tail && !upsample2x && !head),
          "halo conv: the affine + PReLU transform takes a ready (scale, shift) row or a GroupNorm plan");
  REQUIRE(!partial_tiles || (!(stats && stats->partial) && !tail && !upsample2x),
          "halo conv: partial tiles are for convs that publish no statistics");
  REQUIRE(stride == 2 || !w.down_perm, "halo conv: parity-ordered weights belong to a stride-2 conv");
  // head: the single source is virtual (64 split-precision channels built in shared memory from the fp32 inputs)
  REQUIRE(!head || (srcs.size() == 1 && srcs[0].ntaps == 9 && srcs[0].gn_off < 0 && srcs[0].act.C == CONV_BLOCK_K &&
                    stride == 1 && !upsample2x && !tail && head->x && head->cc >= 0 && head->cx >= 1 &&
                    (head->cc + head->cx <= 4 || head->cc + head->cx == 6 || head->cc + head->cx == 8) &&
                    (head->cc == 0 || head->cond)),
          "halo conv: the head is one raw 3x3 over <= 8 fp32 input channels");
  const int PH = a0.H / stride, PW = a0.W / stride;
  auto pp = std::make_shared<ConvHaloParams>();
  ConvHaloParams& p = *pp;
  memset(&p, 0, sizeof(p));
  p.num_par = upsample2x ? 4 : 1;
  REQUIRE(out.B == a0.B && out.H == (upsample2x ? 2 : 1) * PH && out.W == (upsample2x ? 2 : 1) * PW,
          "halo conv: output shape mismatch");
  // partial tiles (ArcFace: 56, 28, 14, 7 px): always the one-image geometry; tiles hanging over the right / bottom edge
  // load zeros there (TMA fill = the conv's padding) and their stores are clipped by the output tensor map
  const bool g2 = !partial_tiles && geo2(PH, PW);
  const bool g1 = !partial_tiles && (geo1(PH, PW) || g2);      // a multi-image geometry (one tile per unit of 2 or 5 whole images)
  const int gi = partial_tiles ? 1 : geo_imgs(PH, PW);
  REQUIRE(g1 || partial_tiles || (PH % HALO_TH == 0 && PW % HALO_TW == 0 && PW >= 16), "halo conv: unsupported spatial size");
  REQUIRE(!(g1 && tail), "halo conv: the tail runs on the one-image geometry");
  if (tail) {
    REQUIRE(out.C == 16 && w.cout == 16 && tail->oc >= 1 && tail->oc <= 4 && !upsample2x && !(stats && stats->partial),
            "halo conv: the tail is a plain 3x3 conv with <= 4 output channels padded to 16");
  } else {
    REQUIRE(out.C == w.cout && out.C % 64 == 0, "halo conv: Cout must be a multiple of 64");
  }

  // ---- segments; K order of the packed weights: [tap][all main channels] then the 1x1 shortcut blocks
  // a 1x1 conv (the attention's qkv / out convs, unet.py:120-121): its first source is the conv's input ("main", one tap)
  const bool conv1x1 = w.taps == 1;
  REQUIRE(!conv1x1 || (srcs[0].ntaps == 1 && stride == 1 && !upsample2x && !tail && !head), "halo conv: a 1x1 conv reads 1x1 sources");
  auto is_main = [&](size_t i) { return conv1x1 ? i == 0 : srcs[i].ntaps != 1; };
  int c_main = 0;
  for (size_t i = 0; i < srcs.size(); ++i)
    if (is_main(i)) c_main += srcs[i].act.C;
  REQUIRE(c_main == w.cin_main && c_main % CONV_BLOCK_K == 0, "halo conv: main channel count mismatch");
  const int main_taps = upsample2x ? 4 : (conv1x1 ? 1 : 9);
  int c_seen = 0, k_short = main_taps * c_main, kblocks = 0;
  bool any_gn = false;
  // tensor map of one source view: (C, Wd, Hd, B) with the given pixel strides, box = the geometry's halo
  auto encode_src = [&](int mi, const bf16* base, int C, int Wd, int Hd, size_t sW, size_t sH, size_t sB) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)a0.B};
    cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sB * 2};
    cuuint32_t box[4] = {(cuuint32_t)CONV_BLOCK_K, (cuuint32_t)HALO_W, (cuuint32_t)HALO_H, 1};
    if (g2) {             // natural order, box = the 5 x 5 grids (top zero row, left zero column) of five images
      box[1] = 5; box[2] = 5; box[3] = 5;
    } else if (g1) {      // (C, W, B, H)-ordered view, box = 10 x 2 images x 10
      std::swap(dims[2], dims[3]);
      std::swap(strides[1], strides[2]);
      box[2] = 2; box[3] = 10;
    }
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode_tiled(&p.a_map[mi], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled(halo) failed with CUresult " + std::to_string((int)r));
  };
  const int pitch = g2 ? 5 : (g1 ? 20 : HALO_W);      // pixels per halo-tile row (HaloGeo<GEO>::PITCH)
  if (stride == 2) {
    // Output (oy, ox) reads input (2 oy + ky - 1, 2 ox + kx - 1): ky = 1 is row oy of the even-row view, ky = 0 / 2 are
    // rows oy - 1 / oy of the odd-row view, likewise in x. So the conv is four small convs over the parity views, whose
    // taps are 2x2 / 1x2 / 2x1 / 1x1 windows of the usual halo tile (top-left corner = view pixel (oy-1, ox-1)); the
    // view's coordinate -1 is the conv's zero padding (TMA zero fill). Weights arrive parity-ordered (PackedConv::down_perm):
    // taps (0,0) (0,2) (2,0) (2,2) | (1,0) (1,2) | (0,1) (2,1) | (1,1).
    const Act& a = a0;
    REQUIRE(a.B == out.B && a.C % CONV_BLOCK_K == 0 && a.C == w.cin_main, "halo conv: source shape mismatch");
    static const int S2[4][4] = {{1, 1, 4, 2}, {0, 1, 2, 2}, {1, 0, 2, 1}, {0, 0, 1, 1}};      // row parity, col parity, taps, taps per row
    const int pix0[4] = {0, pitch, 1, pitch + 1};
    int tap0 = 0;
    for (int i = 0; i < 4; ++i) {
      const int py = S2[i][0], px = S2[i][1];
      encode_src(i, a.ptr + ((size_t)py * a.W + px) * a.C, a.C, PW, PH, (size_t)2 * a.C, (size_t)2 * a.W * a.C,
                 (size_t)a.H * a.W * a.C);
      HaloSeg& sg = p.seg[i];
      sg.map = i; sg.cblocks = a.C / CONV_BLOCK_K; sg.gn_off = srcs[0].gn_off;      // (every view shares the channel table)
      sg.ntaps = S2[i][2]; sg.tap_w = S2[i][3]; sg.pix0 = pix0[i];
      sg.k_base = tap0 * a.C; sg.k_tap_stride = a.C;
      tap0 += sg.ntaps;
    }
    any_gn |= srcs[0].gn_off >= 0;
    k_short = 9 * a.C;
    p.num_segs = 4;
    // 1x1 stride-2 shortcut sources (arcface.py:133-137 downsample): input (2 oy, 2 ox) = the even / even view's centre tap
    for (size_t i = 1; i < srcs.size(); ++i) {
      const Act& r = srcs[i].act;
      REQUIRE(srcs[i].ntaps == 1 && srcs[i].gn_off < 0 && r.B == out.B && r.H == a.H && r.W == a.W && r.C % CONV_BLOCK_K == 0,
              "halo conv: a stride-2 shortcut source is a raw 1x1 over the conv's input grid");
      encode_src(p.num_segs, r.ptr, r.C, PW, PH, (size_t)2 * r.C, (size_t)2 * r.W * r.C, (size_t)r.H * r.W * r.C);
      HaloSeg& sg = p.seg[p.num_segs];
      sg.map = p.num_segs; sg.cblocks = r.C / CONV_BLOCK_K; sg.gn_off = -1;
      sg.ntaps = 1; sg.tap_w = 1; sg.pix0 = pitch + 1; sg.k_base = k_short; sg.k_tap_stride = 0;
      k_short += r.C;
      ++p.num_segs;
    }
  } else
  for (size_t i = 0; i < srcs.size(); ++i) {
    const HaloSource& s = srcs[i];
    const Act& a = s.act;
    REQUIRE(a.B == out.B && a.H == PH && a.W == PW && a.C % CONV_BLOCK_K == 0, "halo conv: source shape mismatch");
    REQUIRE(s.ntaps == 9 || s.ntaps == 1, "halo conv: a source is 3x3 or 1x1");
    REQUIRE(!(upsample2x && s.ntaps == 1), "halo conv: a folded upsample has no shortcut");
    if (!head) encode_src((int)i, a.ptr, a.C, a.W, a.H, (size_t)a.C, (size_t)a.W * a.C, (size_t)a.H * a.W * a.C);
    HaloSeg& sg = p.seg[i];
    sg.map = (int)i;
    sg.cblocks = a.C / CONV_BLOCK_K;
    sg.gn_off = s.gn_off;
    if (is_main(i)) {
      REQUIRE(k_short == main_taps * c_main && kblocks == (c_seen / CONV_BLOCK_K) * main_taps,
              "halo conv: main sources must come first");
      sg.ntaps = main_taps;
      sg.tap_w = upsample2x ? 2 : (conv1x1 ? 1 : 3);
      sg.pix0 = upsample2x ? -1 : (conv1x1 ? pitch + 1 : 0);      // folded upsample: the 2x2 window depends on the output parity
      sg.k_base = c_seen;
      sg.k_tap_stride = c_main;
      c_seen += a.C;
    } else {
      sg.ntaps = 1;
      sg.tap_w = 1;
      sg.pix0 = pitch + 1;                // centre pixel
      sg.k_base = k_short;
      sg.k_tap_stride = 0;
      k_short += a.C;
    }
    kblocks += sg.ntaps * sg.cblocks;
    any_gn |= s.gn_off >= 0;
    if (s.gn_off >= 0)
      REQUIRE((gn != nullptr || gn_from_stats != nullptr) && s.gn_off + a.C <= gn_C, "halo conv: GroupNorm table too small");
  }
  if (stride == 1) p.num_segs = (int)srcs.size();
  REQUIRE(k_short == w.k_total, "halo conv: packed weight K does not match the segment list");
  REQUIRE(w.up_folded == upsample2x, "halo conv: weight packing / upsample mismatch");
  p.tiles_w = g1 ? 1 : (PW + HALO_TW - 1) / HALO_TW;
  p.tiles_h = g1 ? 1 : (PH + HALO_TH - 1) / HALO_TH;
  p.units = (out.B + gi - 1) / gi;
  p.inv_tiles_w = 1.0f / (float)p.tiles_w; p.inv_tiles_h = 1.0f / (float)p.tiles_h;
  p.inv_num_par = 1.0f / (float)p.num_par; p.inv_units = 1.0f / (float)p.units;
  p.B = out.B; p.H = PH; p.W = PW; p.Cout = out.C;
  p.out_H = out.H; p.out_W = out.W;
  p.bias = bias; p.bias_t_stride = bias_t_stride; p.ctl = ctl;
  p.out = out.ptr;
  REQUIRE(!any_gn || gn_C <= 1024, "halo conv: the fused GroupNorm handles at most 1024 channels");
  REQUIRE(!any_gn || gn_swish || prelu, "halo conv: the fused GroupNorm is always followed by Swish (unet.py:84-86)");
  REQUIRE(!prelu || any_gn, "halo conv: PReLU slopes without a transformed source");
  p.gn = any_gn ? gn : nullptr; p.gn_C = gn_C; p.gn_swish = gn_swish ? 1 : 0;
  p.gn_b_stride = prelu ? 0 : gn_C;
  p.xf_slope = prelu_slope;
  if (any_gn && gn_from_stats) {
    const GnPlan& g = *gn_from_stats;
    REQUIRE(g.C0 + g.C1 == gn_C && g.HW == PH * PW && g.B == out.B, "halo conv: GroupNorm plan does not match the sources");
    REQUIRE(g.groups >= 1 && g.groups <= 32 && gn_C % g.groups == 0, "halo conv: the in-kernel GroupNorm table handles <= 32 groups");
    REQUIRE(g.stats0 && (g.C1 == 0 || g.stats1) && g.gamma && g.beta, "halo conv: GroupNorm plan is missing statistics or affine parameters");
    p.gn = nullptr;
    p.gn_stats0 = g.stats0; p.gn_stats1 = g.stats1; p.gn_slots0 = g.slots0; p.gn_slots1 = g.slots1;
    p.gn_C0 = g.C0; p.gn_groups = g.groups; p.gn_gamma = g.gamma; p.gn_beta = g.beta;
  }
  // output maps for the epilogue's tensor stores: box = one warp's slab (64 channels x 8 x 4 pixels);
  // a folded upsample writes output parity (py, px) through a view with doubled pixel strides
  if (tail) {
    p.tail_x = tail->x; p.tail_eps = tail->eps_out; p.coefs = tail->coefs; p.tail_oc = tail->oc;
  }
  if (head) {
    p.head_cond = head->cond; p.head_x = head->x; p.head_cc = head->cc; p.head_cx = head->cx;
  }
  const bool is_head = head != nullptr;
  for (int par = 0; par < p.num_par && !tail; ++par) {
    const int sc = upsample2x ? 2 : 1;
    const int py = par >> 1, px = par & 1;
    bf16* base = out.ptr + ((size_t)py * out.W + px) * out.C;
    cuuint64_t dims[4] = {(cuuint64_t)out.C, (cuuint64_t)(out.W / sc), (cuuint64_t)(out.H / sc), (cuuint64_t)out.B};
    cuuint64_t strides[3] = {(cuuint64_t)sc * out.C * 2, (cuuint64_t)sc * out.W * out.C * 2,
                             (cuuint64_t)out.H * out.W * out.C * 2};
    cuuint32_t box[4] = {(cuuint32_t)CONV_BLOCK_K, (cuuint32_t)HALO_TW, 4, 1};
    if (g2) {             // unused: the 4x4 geometry stores its interior rows from registers
      box[1] = 4; box[2] = 4; box[3] = 1;
    } else if (g1) {      // (C, W, B, H)-ordered view; a warp's slab is 8 x 2 images x 2 rows
      std::swap(dims[2], dims[3]);
      std::swap(strides[1], strides[2]);
      box[2] = 2; box[3] = 2;
    }
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode_tiled(&p.o_map[par], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled(output) failed with CUresult " + std::to_string((int)r));
  }

  // ---- (BLOCK_N, MT): lowest modelled time. Per 64-channel block a super tile costs
  // max(MMA cycles, L2->SM bytes / rate); a CTA runs ceil(super tiles / SMs) of them.
  const long long tiles_img = (long long)p.tiles_w * p.tiles_h;
  const long long m_tiles = tiles_img * p.num_par * p.units;
  int bn = 0, mt = 0, cg = 1;
  {
    int fbn = 0, fmt = 0;
    if (const char* e = getenv("B200SR3_HALO_BN")) fbn = atoi(e);
    if (const char* e = getenv("B200SR3_HALO_MT")) fmt = atoi(e);
    double best = 1e30;
    int fcg = 0;
    if (const char* e = getenv("B200SR3_HALO_CG")) fcg = atoi(e);
    const char* e128 = getenv("B200SR3_HALO_128X2");      // opt-in: +3 % in burst timing, -0.6 % in the full step
    const bool allow_128x2 = e128 && e128[0] == '1';
    int cg_min_bn = 64;
    if (const char* e = getenv("B200SR3_HALO_CG_MIN_BN")) cg_min_bn = atoi(e);
    const int cand[10][3] = {{256, 1, 2}, {128, 1, 2}, {64, 1, 2}, {256, 1, 1}, {128, 2, 1}, {128, 1, 1}, {64, 2, 1}, {64, 1, 1}, {16, 2, 1}, {16, 1, 1}};
    // CTA pairs are opt-in (B200SR3_HALO_CG=2, optionally only for BLOCK_N >= B200SR3_HALO_CG_MIN_BN): in burst timing
    // they are 0-60 % SLOWER than the one-CTA shapes on every layer of the R=128 UNet
    // (profiles/r01e_cta_pair_halo_bench.txt) - the pair runs in lock step through cross-CTA barriers and gives up the
    // MT=2 weight reuse that the Cout=64 layers rely on. Pass 0 looks for a pair shape, pass 1 for a one-CTA shape.
    const bool pair_ok = fcg == 2 && !(g1 || tail || tiles_img % 2 != 0);
    for (int pass = pair_ok ? 0 : 1; pass < 2 && bn == 0; ++pass)
    for (auto& c : cand) {
      if ((c[2] == 2) != (pass == 0)) continue;
      if (c[2] == 2 && c[0] < cg_min_bn) continue;
      if ((c[0] == 16) != (tail != nullptr)) continue;
      if (is_head && c[0] != 64) continue;
      if (prelu && (c[2] != 1 || (c[1] == 2 && c[0] != 64))) continue;      // instantiated: (64,1) (64,2) (128,1) (256,1); two-image: (64,1) (128,1)
      if (prelu && g2) continue;
      if (g1 && (c[1] != 1 || c[0] > 128)) continue;
      if (g2 && c[0] != 64) continue;
      if (c[0] == 128 && c[1] == 2 && !allow_128x2) continue;
      if (out.C % c[0] != 0 || tiles_img % c[1] != 0) continue;
      const int tiles_per_super = c[1] * c[2];
      if ((fbn && c[0] != fbn) || (fmt && c[1] != fmt)) continue;
      // cycles per (M=128 per CTA) MMA, measured; a CTA pair feeds (4 + N/64) KB per MMA instead of (4 + N/32) KB
      const double mma_cyc = c[2] == 2 ? (c[0] == 256 ? 128.0 : (c[0] == 128 ? 64.0 : 40.0))
                                       : (c[0] == 256 ? 128.0 : (c[0] == 128 ? 64.0 : (c[0] == 64 ? 48.0 : 36.0)));
      double per_super = 0.0;
      for (int i = 0; i < p.num_segs; ++i) {
        const double mma = p.seg[i].ntaps * c[1] * 4 * mma_cyc;
        const double bytes = c[1] * (double)(g2 ? 16000 : (g1 ? 25600 : HALO_BYTES)) + p.seg[i].ntaps * c[0] * 128.0 / c[2];
        per_super += p.seg[i].cblocks * std::max(mma, bytes / 56.0);
      }
      per_super += 300.0 + c[1] * c[0] * 5.0;                                      // epilogue drain, not overlapped at the end
      const long long supers = m_tiles / tiles_per_super * (out.C / c[0]);
      const long long units_avail = g_halo_sms / c[2];
      const double rounds = (double)((supers + units_avail - 1) / units_avail);
      const double cost = rounds * per_super;
      if (cost < best) { best = cost; bn = c[0]; mt = c[1]; cg = c[2]; }
    }
    REQUIRE(bn != 0, "halo conv: no tile shape fits (check B200SR3_HALO_BN / B200SR3_HALO_MT)");
  }
  // Deep halo ring (conv_halo.cuh, DEEP): conv2 layers whose K loop is mostly single-tap shortcut stages. Rules from the
  // same-box A/B in profiles/r02d_ab.txt; B200SR3_HALO_DEEP=0 switches it off, =1 forces it wherever it is instantiated.
  bool deep = false;
  {
    int sc_blocks = 0, main_blocks = 0;
    for (int i = 0; i < p.num_segs; ++i) (p.seg[i].ntaps == 1 ? sc_blocks : main_blocks) += p.seg[i].cblocks;
    const char* e = getenv("B200SR3_HALO_DEEP");
    const int force = e ? atoi(e) : -1;
    const bool can = !g1 && cg == 1 && !tail && !is_head && any_gn && !prelu && stride == 1 && !upsample2x && (bn == 64 || bn == 128) &&
                     !getenv("B200SR3_HALO_BN") && !getenv("B200SR3_HALO_MT");
    if (can && force != 0) {
      if (bn == 64 && (sc_blocks >= 2 || force == 1)) { deep = true; mt = 1; }
      else if (bn == 128 && mt == 1 && ((sc_blocks >= 1 && main_blocks <= 2) || force == 1)) deep = true;
    }
  }
  // Resident weights (conv_halo.cuh, ConvHaloParams::w_resident): one-tile, three-stage shape whose 12-stage weight ring
  // holds every weight tile of the layer. OPT-IN (B200SR3_W_RESIDENT=1; =2 restricts it to layers without 1x1 shortcut
  // segments over a concatenated input): measured and rejected - the only shape with room for all weight tiles gives up the
  // two-tile super tiles / the deep halo ring, and that costs more than the weight stream (same box, profiles/r03u_ab.txt:
  // 64 -> 64 53.3 -> 59.2 us, + identity 61.0 -> 67.3, + res128 80.6 -> 99.2, + res192 89.2 -> 113.5; identical results).
  p.w_resident = 0;
  {
    int n_w = 0, sc_blocks = 0;
    for (int i = 0; i < p.num_segs; ++i) {
      n_w += p.seg[i].cblocks * p.seg[i].ntaps;
      if (p.seg[i].ntaps == 1) sc_blocks += p.seg[i].cblocks;
    }
    const char* e = getenv("B200SR3_W_RESIDENT");
    const int mode = e ? atoi(e) : 0;
    static_assert(HaloSmem<64, 1, 0, 1, false>::W_STAGES >= 12, "resident weights count on a 12-stage ring");
    const bool can = !g1 && !g2 && cg == 1 && !tail && !is_head && !prelu && bn == 64 && out.C == 64 && p.num_par == 1 &&
                     n_w <= 12 && !getenv("B200SR3_HALO_BN") && !getenv("B200SR3_HALO_MT");
    if (can && mode != 0 && !(mode == 2 && sc_blocks >= 2)) { deep = false; mt = 1; p.w_resident = n_w; }
  }
  p.tiles_n = out.C / bn;
  p.total_super = (int)(m_tiles / (mt * cg) * p.tiles_n);
  p.seg_len_super = (int)(tiles_img * p.num_par / (mt * cg));
  {
    cuuint64_t dims[2] = {(cuuint64_t)w.k_total, (cuuint64_t)w.cout * p.num_par};
    cuuint64_t strides[1] = {(cuuint64_t)w.k_total * 2};
    cuuint32_t box[2] = {(cuuint32_t)CONV_BLOCK_K, (cuuint32_t)(bn / cg)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_tiled(&p.w_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w.w, dims, strides, box, estr,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled(weights) failed with CUresult " + std::to_string((int)r));
  }
  const int grid = cg * std::min(p.total_super, g_halo_sms / cg);
  REQUIRE((long long)p.total_super * (grid / cg + 1) < (1ll << 31) && m_tiles * p.tiles_n < (1ll << 22),
          "halo conv: too many tiles for the kernel's 32-bit / float-reciprocal tile arithmetic");
  if (const char* ab = getenv("B200SR3_CONV_ABLATE")) p.ablate = atoi(ab);
  p.pdl = pdl_enabled() ? 1 : 0;
  if (stats) p.dbg = stats->dbg;
  if (stats && stats->partial) {
    REQUIRE(stats->atomic ? stats->slots == 1 : cg * slots_needed(p.seg_len_super, p.total_super, grid / cg) <= stats->slots,
            "halo conv: statistics scratch has too few slots");
    p.stat_partial = stats->partial;
    p.stat_slots = stats->slots;
    p.stat_atomic = stats->atomic ? 1 : 0;
  }

  Op op;
  op.name = name;
  op.is_conv = true;
  {
    const double m = (double)out.B * out.H * out.W;
    double k = 0, k_exec = 0;
    for (const HaloSource& s : srcs) {      // (a stride-2 conv has one 9-tap source: 9 C either way)
      // reference graph: full 3x3 at output resolution; a res_conv segment counts, an identity shortcut
      // that merely rides the GEMM (unet.py:101, nn.Identity) does not
      if (!(s.identity || (s.ntaps == 1 && !conv1x1 && w.res_identity))) k += (double)s.ntaps * s.act.C;
      k_exec += (double)(s.ntaps == 1 ? 1 : main_taps) * s.act.C;
    }
    op.flops = 2.0 * m * (double)(tail ? tail->oc : out.C) * k;
    op.flops_executed = 2.0 * m * (double)out.C * k_exec;
  }
  REQUIRE(!(prelu && g2), "halo conv: the affine + PReLU transform is not instantiated for the five-image geometry");
  op.run = [pp, grid, bn, mt, cg, any_gn, g1, g2, is_head, deep, prelu](cudaStream_t s) {
    if (prelu && g1) {
      if (bn == 128) launch_pdl(conv_halo_kernel<128, 1, true, 1, 1, false, false, true>, dim3(grid), dim3(halo_threads(128)), HaloSmem<128, 1, 1, 1>::TOTAL, s, *pp);
      else launch_pdl(conv_halo_kernel<64, 1, true, 1, 1, false, false, true>, dim3(grid), dim3(halo_threads(64)), HaloSmem<64, 1, 1, 1>::TOTAL, s, *pp);
    } else if (prelu) {
      if (bn == 256) launch_pdl(conv_halo_kernel<256, 1, true, 0, 1, false, false, true>, dim3(grid), dim3(halo_threads(256)), HaloSmem<256, 1, 0, 1>::TOTAL, s, *pp);
      else if (bn == 128) launch_pdl(conv_halo_kernel<128, 1, true, 0, 1, false, false, true>, dim3(grid), dim3(halo_threads(128)), HaloSmem<128, 1, 0, 1>::TOTAL, s, *pp);
      else if (mt == 2) launch_pdl(conv_halo_kernel<64, 2, true, 0, 1, false, false, true>, dim3(grid), dim3(halo_threads(64)), HaloSmem<64, 2, 0, 1>::TOTAL, s, *pp);
      else launch_pdl(conv_halo_kernel<64, 1, true, 0, 1, false, false, true>, dim3(grid), dim3(halo_threads(64)), HaloSmem<64, 1, 0, 1>::TOTAL, s, *pp);
    } else if (deep) {
      if (bn == 64) launch_pdl(conv_halo_kernel<64, 1, true, 0, 1, false, true>, dim3(grid), dim3(halo_threads(64)), HaloSmem<64, 1, 0, 1, true>::TOTAL, s, *pp);
      else launch_pdl(conv_halo_kernel<128, 1, true, 0, 1, false, true>, dim3(grid), dim3(halo_threads(128)), HaloSmem<128, 1, 0, 1, true>::TOTAL, s, *pp);
    } else if (is_head) {
      if (mt == 2) launch_pdl(conv_halo_kernel<64, 2, false, 0, 1, true>, dim3(grid), dim3(halo_threads(64)), HaloSmem<64, 2, 0, 1>::TOTAL, s, *pp);
      else launch_pdl(conv_halo_kernel<64, 1, false, 0, 1, true>, dim3(grid), dim3(halo_threads(64)), HaloSmem<64, 1, 0, 1>::TOTAL, s, *pp);
    } else if (cg == 2) {
      if (bn == 256) launch_halo_pair<256>(*pp, any_gn, grid, s);
      else if (bn == 128) launch_halo_pair<128>(*pp, any_gn, grid, s);
      else launch_halo_pair<64>(*pp, any_gn, grid, s);
    } else if (g2) launch_halo<64, 1, 2>(*pp, any_gn, grid, s);
    else if (g1 && bn == 128) launch_halo<128, 1, 1>(*pp, any_gn, grid, s);
    else if (g1) launch_halo<64, 1, 1>(*pp, any_gn, grid, s);
    else if (bn == 256) launch_halo<256, 1>(*pp, any_gn, grid, s);
    else if (bn == 128 && mt == 2) launch_halo<128, 2>(*pp, any_gn, grid, s);
    else if (bn == 128) launch_halo<128, 1>(*pp, any_gn, grid, s);
    else if (bn == 64 && mt == 2) launch_halo<64, 2>(*pp, any_gn, grid, s);
    else if (bn == 64) launch_halo<64, 1>(*pp, any_gn, grid, s);
    else if (mt == 2) launch_halo<16, 2>(*pp, any_gn, grid, s);
    else launch_halo<16, 1>(*pp, any_gn, grid, s);
  };
  op.weights = w.w;
  op.weight_bytes = (size_t)w.cout * p.num_par * w.k_total * sizeof(bf16);
  op.set_prefetch = [pp, grid](const void* ptr, size_t bytes) {
    pp->pf_ptr = static_cast<const uint8_t*>(ptr);
    pp->pf_total = (unsigned)std::min<size_t>(bytes, 64u << 20);
    pp->pf_slice = (unsigned)(((pp->pf_total + grid - 1) / grid + 127) / 128 * 128);
  };
#if B200SR3_ROLE_TIMING
  // timing build: role counters of the launches named in B200SR3_TIMING_OPS ("all" or a comma-separated list of op
  // names), printed by Engine::profile_step (tools/profile_step.py)
  if (const char* sel = getenv("B200SR3_TIMING_OPS")) {
    const std::string list = std::string(",") + sel + ",";
    if (std::string(sel) == "all" || list.find("," + name + ",") != std::string::npos) {
      unsigned long long* buf = nullptr;
      CUDA_CHECK(cudaMalloc(&buf, 256 * 16 * sizeof(unsigned long long)));
      CUDA_CHECK(cudaMemset(buf, 0, 256 * 16 * sizeof(unsigned long long)));
      std::shared_ptr<unsigned long long> hold(buf, [](unsigned long long* q) { cudaFree(q); });
      pp->dbg = buf;
      const std::string label = name;
      op.report = [hold, label]() {
        halo_report_timing(hold.get(), label.c_str());
        cudaMemset(hold.get(), 0, 256 * 16 * sizeof(unsigned long long));
      };
    }
  }
#endif
  if (params_out) *params_out = pp;
  return op;
}

}  // namespace b200sr3
