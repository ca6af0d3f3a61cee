// Halo-resident implicit-GEMM 3x3 convolution (tcgen05 / TMEM / TMA), second generation.
//
// conv_umma.cuh fetches one 16 KB A tile per filter tap, i.e. it pulls the input through L2 nine
// times; measured on B200 that L2->SM traffic, not the tensor pipe, bounds every conv with
// Cout <= 256 (profiles/r01b). Here ONE 4-D TMA box load brings the (8+2) x (16+2) pixel x
// 64-channel HALO of an 8x16-pixel output tile into shared memory (128 B per pixel, 128B-swizzled)
// and the nine taps are nine tcgen05 A descriptors into that same tile: tap (ky,kx) starts at
// pixel ky*10+kx - a 128-byte row that is not 1024-byte aligned - with SBO = 10 pixels (1280 B)
// between the 8-row groups (a group = 8 consecutive x of one output row). The hardware swizzles on
// absolute shared-memory address bits, so this reads exactly the shifted window
// (tools/micro/halo_umma.cu proves it bit-exactly and shows the MMA rate is unchanged).
// A traffic drops from 9 x 16 KB to 22.5 KB per 64-channel block.
//
// Because the input now sits in shared memory ONCE per use, GroupNorm-apply + Swish
// (unet.py:84-86, Block = GN -> Swish -> Conv) is fused in: four "transform" warps rewrite the
// halo tile in place, y = swish(x * scale[b][c] + shift[b][c]) for in-image pixels (the conv's zero
// padding must stay zero, so out-of-image pixels are left as TMA zero-filled them), between the
// TMA arrival and the MMAs. The normalised tensor is never written to HBM; the channel concat
// cat((x, skip), 1) (unet.py:261) is two K segments with their own tensor maps.
//
// K loop: segments (main source(s) with 9 taps [or the 4 parity taps of a folded nearest-2x
// upsample], then raw 1x1 shortcut sources, unet.py:103-110) x 64-channel blocks x taps. Weights
// stream through their own ring, one [BLOCK_N x 64] tile per (tap, block). A CTA processes MT
// horizontally adjacent tiles against each weight tile (MT x fewer weight bytes from L2).
//
// Warp roles (512 threads): 0-3 and 8-11 = GroupNorm/Swish transform, 4-7 = epilogue, 12 = halo TMA
// producer, 13 = weight TMA producer, 14 = TMEM allocator, 15 = MMA issuer. The single-thread roles
// carry the HIGHEST warp ids on purpose: a scheduler (warp id % 4) prefers its highest-numbered ready
// warp, and with the issuer below the ALU-heavy transform / epilogue warps the tensor pipe starved.
//
// Epilogue: each of the four warps owns 32 accumulator rows = a 4-row x 8-pixel slab of the tile. Per
// 64 output channels it moves TMEM -> registers -> (+bias[t], bf16) -> a 4 KB 128B-swizzled staging slab
// in shared memory, hands the slab to a TMA tensor store (asynchronous, fully coalesced; the four
// output parities of a folded upsample are four strided tensor maps) and forms the GroupNorm statistics
// of the output from the slab: lane l adds up columns 2l, 2l+1 over the slab's 32 rows (conflict-free
// LDS.32), i.e. the cross-row reduction costs no shuffles. Statistics are those of the bf16-rounded
// tensor - exactly what the consumer normalises - and accumulate as int64 fixed point.
#pragma once
#include "conv_umma.cuh"

namespace b200sr3 {

constexpr int HALO_W = 10, HALO_H = 18;                  // (8+2) x (16+2) pixels
constexpr int HALO_TW = 8, HALO_TH = 16;                 // output tile
constexpr int HALO_BYTES = HALO_W * HALO_H * 128;        // 23040: what one TMA box load delivers
constexpr int HALO_STRIDE = 23552;                       // rounded up to the 1024 B swizzle period
// Epilogue warp sets (alternating 64-channel units). Measured on B200: a second set does not pay - the
// BLOCK_N = 64 layers are bound by shared-memory bandwidth (UMMA operand reads + TMA writes + transform
// + staging = ~370 KB per tile at 128 B/clk), not by epilogue issue slots - so every shape runs one set.
constexpr int halo_esets(int block_n) { return (void)block_n, 1; }
constexpr int halo_threads(int block_n) { return 512 + 128 * (halo_esets(block_n) - 1); }
// halo ring depth: 3 (load / transform / MMA). A 6-deep ring for one-tile BLOCK_N = 64 shapes was tried for
// the single-tap shortcut segments and does not pay: those layers are shared-memory-bandwidth bound.
// DEEP ring (template flag, one-tile GEO 0 shapes with BLOCK_N <= 128): conv2 of a ResnetBlock whose shortcut is a res_conv
// over a concatenated input (unet.py:103-110,261) walks 1 + (2..6) halo stages per tile, all but the first single-tap. With
// three stages the NEXT tile's main stage cannot be fetched and transformed while this tile's MMAs run (role counters,
// profiles/r02c_roles.txt: the MMA issuer waited 4.2 k of 11.1 k cycles per super tile on `64->64 + res192 @128`; ncu:
// shared-memory pipe 50 %, TMA 17 %, DRAM 44 % - a latency-bound pipeline, not a bandwidth-bound one). Measured on B200
// (profiles/r02d_ab.txt, burst): 99.1 -> 89.5 us (64 + res192), 89.3 -> 81.2 (64 + res128), 65.5 -> 61.7 / 54.0 -> 49.2 /
// 49.4 -> 46.3 (128-channel conv2 layers); layers of 9-tap blocks only lose 4-13 % (fewer weight stages fit), so the host
// picks it per layer.
constexpr int halo_a_stages(int block_n, int mt, int geo, bool deep = false) {
  return (deep && geo == 0 && mt == 1) ? (block_n == 64 ? 6 : (block_n == 128 ? 5 : 3)) : 3;
}
constexpr int HALO_MAX_SEGS = 6;      // e.g. a stride-2 conv (four parity views) plus two shortcut sources

// Tile geometry. GEO 0: an 8 x 16 pixel tile of one image, halo 10 x 18. GEO 1 (8 x 8 images): a tile is
// TWO whole images laid side by side in the halo tile - pixel (hy, image, hx) at hy*20 + image*10 + hx,
// fetched by one TMA box over a (C, W, B, H)-ordered view - so that the 8-row groups (one image row of
// one image each, ordered y-major, image-minor) are again exactly 10 pixels apart.
// GEO 2 (4 x 4 images): a tile is FIVE whole images; accumulator row r is LINEAR pixel r of their 5x5 grids laid end to
// end - pixel = image*25 + hy*5 + hx with hy = y+1, hx = x+1, i.e. each image carries its TOP zero row and LEFT zero
// column only, exactly what one (C, W, H, B) box of 5 x 5 x 5 starting at (-1,-1) delivers. The right neighbour of x = 3
// is then the next row's zero column and the row below y = 3 is the next image's zero row (for the last image: ten pixels
// after the box that are zeroed once and never written), so every border read of an interior pixel lands on a zero. The
// 8-row groups are dense (8 pixels apart) and tap (ky,kx) is the window shifted by (ky-1)*5 + (kx-1) pixels. 80 of the
// 128 rows are outputs; zero-row / zero-column / tail rows compute garbage that is neither stored nor counted. A tile is
// preceded by a 1 KB guard and its stage is long enough for the +6-pixel window of row 127.
template <int GEO> struct HaloGeo;
template <> struct HaloGeo<0> {
  static constexpr int PITCH = 10, ROWS = 18, NPIX = 180, BYTES = 180 * 128, STRIDE = 23552, IMGS = 1;
  static constexpr int GROUP = 10, TMA_OFF = 0, MMA_OFF = 0;      // pixels between 8-row groups; byte offsets in a stage
};
template <> struct HaloGeo<1> {
  static constexpr int PITCH = 20, ROWS = 10, NPIX = 200, BYTES = 200 * 128, STRIDE = 26624, IMGS = 2;
  static constexpr int GROUP = 10, TMA_OFF = 0, MMA_OFF = 0;
};
template <> struct HaloGeo<2> {
  static constexpr int PITCH = 5, ROWS = 25, NPIX = 125, BYTES = 125 * 128, STRIDE = 18432, IMGS = 5;
  static constexpr int GROUP = 8, TMA_OFF = 1024, MMA_OFF = 1024 - 6 * 128;      // tap (0,0) of row 0 is pixel -6
};

struct HaloSeg {
  int map;           // a_map index
  int cblocks;       // 64-channel blocks of this source
  int ntaps;         // 9: 3x3; 4: parity 2x2 of a folded upsample; 1: 1x1 shortcut (centre pixel); 4/2/2/1: the four
                     // input-parity views of a stride-2 conv (unet.py:68-74)
  int k_base;        // K column of (tap 0, block 0) in the packed weights
  int k_tap_stride;  // K columns between consecutive taps
  int gn_off;        // channel offset of this source in the GN table; < 0: raw input, no transform
  // the taps' windows into the halo tile: tap i reads the window that starts at pixel pix0 + (i / tap_w) * PITCH + i % tap_w
  // (pixel 0 = the halo's top-left corner, i.e. filter tap (0,0) of a 3x3). pix0 < 0: the folded upsample, whose 2x2
  // window depends on the tile's output parity.
  int tap_w;
  int pix0;
};

struct alignas(64) ConvHaloParams {
  CUtensorMap a_map[HALO_MAX_SEGS];
  CUtensorMap w_map;
  CUtensorMap o_map[4];            // output, box (64 ch, 8, 4, 1); one per output parity when num_par == 4
  HaloSeg seg[HALO_MAX_SEGS];
  int num_segs;
  int num_par;                     // 1, or 4 (folded nearest-2x upsample: output parity (y&1, x&1))
  int tiles_w, tiles_h;            // 8x16 tiles per image over the pixel space the tiles walk
  int B, H, W, Cout;               // pixel space the tiles walk (the SOURCE grid when num_par == 4)
  int units;                       // images (GEO 0) or image pairs (GEO 1) the tile index walks
  float inv_tiles_w, inv_tiles_h, inv_num_par, inv_units;      // 1 / x of the four divisors of the tile decode
  int out_H, out_W;
  int tiles_n, total_super;        // N tiles; super tiles (MT tiles each) in the launch
  int seg_len_super;               // super tiles per (n tile, image) = tiles_w*tiles_h*num_par/MT
  const float* bias;               // [Cout] or a [rows][stride] table indexed by ctl->t
  int bias_t_stride;
  const StepCtl* ctl;
  bf16* out;                       // NHWC [B][out_H][out_W][Cout]
  const float2* gn;                // [B][gn_C] (scale, shift) of the fused GroupNorm (null: none, or built in-kernel)
  int gn_C;
  int gn_swish;
  int gn_b_stride;                 // float2 elements between consecutive images' rows of `gn` (0: one row for every image)
  // PRELU (template flag): the transform is y = scale * prelu(x; slope) + shift per channel instead of GroupNorm + Swish
  // - the BatchNorm / PReLU in front of a conv of the ArcFace iResNet (model/mica/arcface.py:58-66), whose zero
  // padding must stay zero just like the UNet's. (scale, shift) come from `gn` (one row, gn_b_stride = 0).
  const float* xf_slope;           // [gn_C] PReLU slopes (1 = no PReLU)
  // In-kernel (scale, shift) table: the transform warps turn the producers' per-channel partial sums
  // ([B][slots][C][2] int64 fixed point, as written by a conv epilogue) into the table of the current image themselves,
  // which removes the gn_scale_shift launch that otherwise sits between every two convs (same arithmetic, same bits).
  const long long* gn_stats0;      // channels [0, gn_C0) of the normalised (concatenated) input; null: use `gn`
  const long long* gn_stats1;      // channels [gn_C0, gn_C)
  int gn_slots0, gn_slots1, gn_C0, gn_groups;
  const float* gn_gamma;           // [gn_C]
  const float* gn_beta;
  long long* stat_partial;         // as ConvParams::stat_partial
  int stat_slots;
  // stat_atomic: one [B][Cout][2] accumulator (stat_slots == 1), zeroed once per sampling step by the engine, that every
  // CTA adds its int64 fixed-point sums into (RED.64). Integer addition is associative, so the result is as
  // batch-invariant as the slot scheme, and the CONSUMER's GroupNorm table needs one load per channel instead of a chain
  // of `slots` dependent L2 round trips (role counters, profiles/r02i_roles_in_situ.txt: 10-13 k cycles to the table
  // with 5-6 slots - the largest single piece of every GroupNorm conv's fill time).
  int stat_atomic;
  // BLOCK_N == 16 ("tail"): the conv is final_conv (unet.py:233, Cout = out_channel <= 4 padded to 16) and
  // the epilogue applies the sampler update (diffusion.py:144-187) to the fp32 NCHW state instead of
  // storing an activation: eps -> x0 = clamp(A x - B eps) -> mean -> + sigma z.
  // HEAD (template flag): the conv is downs.0 (unet.py:187, in_channel -> inner_channel) and its A operand never exists
  // in HBM: the transform warps read the fp32 NCHW sampler inputs cat([cond, x], 1) (diffusion.py:170) for the halo
  // pixels themselves and write the split-precision bf16 rows (hi | lo | hi, see launch_pack_head_split_weight)
  // straight into the swizzled tile. No TMA load, no 64-channel padded operand in memory.
  const float* head_cond;          // fp32 NCHW [B][head_cc][H][W] or null
  const float* head_x;             // fp32 NCHW [B][head_cx][H][W]
  int head_cc, head_cx;            // head_cc + head_cx <= 8
  float* tail_x;                   // fp32 NCHW [B][tail_oc][H][W], updated in place (null: eps only)
  float* tail_eps;                 // optional fp32 NCHW eps output
  const float* coefs;              // [5][T]: A, Bc, C1, C2, LV
  int tail_oc;
  // measurement only (B200SR3_CONV_ABLATE bit mask; results are then wrong): 1 = no global stores,
  // 2 = transform arrives without touching the tile, 4 = no weight loads, 8 = no halo loads,
  // 16 = no TMEM loads, 32 = no statistics math
  int pdl;                         // launched with the programmatic-dependent-launch attribute (B200SR3_PDL=1)
  int ablate;
  // L2 prefetch of the NEXT conv's weights (constants): every CTA requests its slice [blockIdx.x * pf_slice, + pf_slice)
  // of them while this kernel runs. In situ the first tile of a BLOCK_N = 256 layer waited ~16 % of its time for weight
  // tiles (profiles/r02k_roles_in_situ.txt, `MMA waits W`): all 148 CTAs walk the weight matrix in the same order, so
  // every tile's first request is an HBM miss that a three-stage ring cannot cover.
  const uint8_t* pf_ptr;
  unsigned pf_slice, pf_total;
  // Resident weights (0: off, else the number of weight tiles a super tile walks, <= the ring's stages): every super tile
  // of the launch uses the SAME weight tiles (one n tile, one parity) and they all fit the ring, so the producer loads them
  // once per CTA and the MMA issuer neither waits for nor recycles a stage after the first super tile. With no weight loads
  // at all the 64 -> 64 layers at 128^2 ran 11-13 % faster and left the power cap (tools/ablate_loads.sh): their weight
  // stream (74 KB per super tile through L2 -> shared memory) competes with the MMAs' operand reads for the port. Opt-in
  // only: the one shape with room for 12 resident tiles has one-tile super tiles and a three-stage halo ring, and is slower
  // (conv_halo.cu, profiles/r03u_ab.txt).
  int w_resident;
  // optional role timing (B200SR3_CONV_TIMING=1 in b200sr3_conv_block): [grid][16] cycle counters
  unsigned long long* dbg;
};

#ifdef __CUDACC__
template <int BLOCK_N, int MT, int GEO = 0, int CG = 1, bool DEEP = false>
struct HaloSmem {
  static constexpr int A_STAGE = MT * HaloGeo<GEO>::STRIDE;
  static constexpr int AST = halo_a_stages(BLOCK_N, MT, GEO, DEEP);
  static constexpr int A_BYTES = AST * A_STAGE;
  static constexpr int W_STAGE = (BLOCK_N / CG) * 128;           // a CTA of a cta_group::2 pair holds half the weight tile
  static constexpr int NSTG = (BLOCK_N == 256 || MT == 2) ? 1 : 2;   // staging slabs per epilogue warp
  static constexpr int ESETS = halo_esets(BLOCK_N);
  static constexpr int STG_BYTES = 4 * ESETS * NSTG * 4096;
  // (scale, shift) of the current image (pair), <= 1024 channels; every 8-channel group is followed by
  // 16 bytes of padding so that the eight groups a warp reads at once fall into different banks
  static constexpr int GN_BYTES = HaloGeo<GEO>::IMGS * 1024 * 10 + 2048;     // + (mean, rstd) of <= 5 x 32 groups
  static constexpr int GSTAT_OFFSET_IN_GN = HaloGeo<GEO>::IMGS * 1024 * 10;
  // the current n tile's BLOCK_N bias values, staged by the epilogue warps (<= 1 KB, behind the <= 1280 B of group statistics
  // that only the five-image geometry keeps)
  static constexpr int BIAS_OFFSET_IN_GN = GSTAT_OFFSET_IN_GN + (HaloGeo<GEO>::IMGS == 5 ? 1280 : 0);
  static_assert(BIAS_OFFSET_IN_GN + BLOCK_N * 4 <= GN_BYTES, "bias row must fit behind the GroupNorm table");
  static constexpr int BUDGET = 227 * 1024 - 1024 - 512;          // minus alignment slack and barriers
  static constexpr int W_FIT = (BUDGET - A_BYTES - STG_BYTES - GN_BYTES) / W_STAGE;
  static constexpr int W_STAGES = W_FIT > 12 ? 12 : W_FIT;
  static constexpr int W_OFFSET = A_BYTES;
  static constexpr int STG_OFFSET = W_OFFSET + W_STAGES * W_STAGE;
  static constexpr int GN_OFFSET = STG_OFFSET + STG_BYTES;
  static constexpr int BAR_OFFSET = GN_OFFSET + GN_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;
  static_assert(W_STAGES >= 3, "weight ring too shallow");
  static_assert(BLOCK_N * 16 * HaloGeo<GEO>::IMGS <= NSTG * 4096, "statistics hand-over must fit a warp's slabs");
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// ---- cta_group::2 (CTA pair) helpers; semantics verified in tools/micro/umma_2cta.cu
constexpr uint32_t HALO_PEER_MASK = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address -> rank 0's copy
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {      // bar: shared::cluster address (any CTA of the cluster)
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {        // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// MUFU.TANH: max relative error 2^-11, far below the bf16 rounding of the value it produces
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Packed fp32 pairs (FFMA2 / FADD2, sm_100): one issue slot for two IEEE round-to-nearest operations, bit-identical to
// two scalar fmaf / adds. The transform and epilogue warps are bound by their own instruction latency (one or two warps per
// scheduler), not by the FMA pipe, so halving the issue count of their arithmetic is what shortens them.
__device__ __forceinline__ void fma2(float& dx, float& dy, float ax, float ay, float bx, float by, float cx, float cy) {
  asm("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%6, %7};\n\t"
      "fma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}"
      : "=f"(dx), "=f"(dy) : "f"(ax), "f"(ay), "f"(bx), "f"(by), "f"(cx), "f"(cy));
}
__device__ __forceinline__ void add2(float& dx, float& dy, float ax, float ay, float bx, float by) {
  asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
      : "=f"(dx), "=f"(dy) : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// A descriptor into a halo tile: 8-row groups 10 pixels (1280 B) apart
__device__ __forceinline__ uint64_t make_halo_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((HALO_W * 128) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Split-precision operand row of the head conv: N fp32 inputs -> hi | lo | hi (bf16-exact floats), zero padded to 32
// (launch_pack_head_split_weight carries w_hi | w_hi | w_lo in the same positions).
template <int N>
__device__ __forceinline__ void halo_head_row(const float* v, float* row) {
#pragma unroll
  for (int i = 0; i < 32; ++i) row[i] = 0.f;
#pragma unroll
  for (int c = 0; c < N; ++c) {
    const float hi = __bfloat162float(__float2bfloat16_rn(v[c]));
    row[c] = hi;
    row[N + c] = v[c] - hi;
    row[2 * N + c] = hi;
  }
}

// Role timing is a COMPILE-TIME option (-DB200SR3_ROLE_TIMING=1, `build.py --timing` -> libb200sr3_timing.so): with the
// counters compiled in, every hot loop carries parameter loads, clock reads and branches, and the kernel's hot code set
// grows - measured, unrolling the nine taps of the issue loop alone (more code, fewer instructions executed) cost 6 %.
#ifndef B200SR3_ROLE_TIMING
#define B200SR3_ROLE_TIMING 0
#endif
#define HALO_DBG (B200SR3_ROLE_TIMING && p.dbg != nullptr)
#define HALO_ABLATE(bits) (B200SR3_ROLE_TIMING && (p.ablate & (bits)))      // measurement switches: same build option
#define HDBG_DECL() unsigned long long hd[4] = {0ull, 0ull, 0ull, 0ull}; long long hd_t0 = 0
#define HDBG_T0() do { if (HALO_DBG) hd_t0 = clock64(); } while (0)
#define HDBG_ACC(i) do { if (HALO_DBG) hd[i] += (unsigned long long)(clock64() - hd_t0); } while (0)
#ifdef HALO_EPI_PROFILE      // one-off experiment: the epilogue's sections take the transform's counter slots
#define HEPI_T0() HDBG_T0()
#define HEPI_ACC(i) do { HDBG_ACC(i); HDBG_T0(); } while (0)
#else
#define HEPI_T0() do { } while (0)
#define HEPI_ACC(i) do { } while (0)
#endif
#define HDBG_FLUSH(slot, n) do { if (HALO_DBG) for (int _i = 0; _i < (n); ++_i) p.dbg[blockIdx.x * 16 + (slot) + _i] = hd[_i]; } while (0)

// Table slot of channel c (float index; one image's row starts at float2 index im * gn_pitch): 20 floats per 8 channels -
// four (scale c, scale c+1, shift c, shift c+1) quads, so that a 16-byte load hands the transform two register PAIRS for
// its packed FFMA2s, and 4 floats of padding, so that the eight 64-byte groups a warp reads at once fall into different banks.
__device__ __forceinline__ int gn_tab_idx(int c) { return (c >> 3) * 20 + ((c & 6) << 1) + (c & 1); }
__device__ __forceinline__ void gn_tab_put(float2* row, int c, float2 v) {
  float* f = reinterpret_cast<float*>(row) + gn_tab_idx(c);
  f[0] = v.x; f[2] = v.y;
}
__device__ __forceinline__ float2 gn_tab_get(const float2* row, int c) {
  const float* f = reinterpret_cast<const float*>(row) + gn_tab_idx(c);
  return make_float2(f[0], f[2]);
}

// (scale, shift) table of the fused GroupNorm for the IMGS images starting at image b0, built by the 256 transform threads
// (tt = 0..255) into shared memory, image im's row at gtab + im * gn_pitch, channel slots as gn_tab_idx says.
// From the producers' statistics it is ONE pass with no intermediate barrier: TPG adjacent lanes own a (image, group).
// Lane s takes the group's channels s, s + TPG, ...: for each it adds up the producer's int64 partial sums (exact, any
// number of slots), converts to float and accumulates the group sums; a butterfly over the TPG lanes gives every lane
// the group's mean and 1/std; the lane then writes (scale, shift) = (rstd * gamma, beta - mean * rstd * gamma) of its
// channels (pre-halved for the tanh form of Swish). gamma / beta are requested together with the statistics, so the
// table costs one global round trip. The summation order is fixed by the lane mapping: a face's table does not depend
// on its batch or tile. Statistics are read through L2 only (nothing stale in L1 under programmatic dependent launch).
template <int IMGS>
__device__ __forceinline__ void halo_build_gn_table(const ConvHaloParams& p, float2* gtab, int gn_pitch, int tt, int b0,
                                                    bool do_swish, long long t_entry = 0) {
  asm volatile("bar.sync 2, 256;" ::: "memory");      // everyone is done with the previous table
  if (HALO_DBG && tt == 0 && t_entry) p.dbg[blockIdx.x * 16 + 9] = (unsigned long long)(clock64() - t_entry);   // [9] table: past the first barrier
  if (p.gn_stats0 && IMGS == 5) {
    // Five images per tile (4x4 level): 160 (image, group) pairs over 256 lanes leave one lane per group, which then
    // walked its 16-32 channels' statistics four at a time - 4 to 8 dependent L2 round trips, 14 k (C = 512) to 33 k
    // (C = 1024) cycles in front of the first MMA (profiles/r02x_roles_R64.txt). Channel-parallel instead, three phases
    // over shared memory: (1) lane -> channels tt + 256 j, all five images: load the sums (ten independent loads in
    // flight per round, gamma / beta requested with the first round), park them as floats in the channel's table slot;
    // (2) lane -> (image, group): mean and 1/std from the parked sums, in channel order (fixed summation order: batch-
    // invariant); (3) lane -> its channels again: (scale, shift) over the slot. No integer division anywhere: with
    // `idx / C` per item the index arithmetic alone cost ~500 cycles per item (measured: 11 k cycles for C = 512).
    const int C = p.gn_C, C0 = p.gn_C0;
    const int cg = C / p.gn_groups;
    const float inv_cg = 1.0f / (float)cg;
    float2* gstat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(gtab) + IMGS * 1024 * 10);
    float gm[4], bt[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = tt + 256 * j;
      gm[j] = c < C ? __ldg(p.gn_gamma + c) : 0.f;
      bt[j] = c < C ? __ldg(p.gn_beta + c) : 0.f;
    }
    // statistics of (image im, channel c): slot 0 at base + im * img_stride (images past the batch repeat the last one)
    auto chan_base = [&](int c, int& slots, int& stride, size_t& img_stride) {
      const bool second = c >= C0;
      const int cs = second ? C - C0 : C0, cl = second ? c - C0 : c;
      slots = second ? p.gn_slots1 : p.gn_slots0;
      stride = cs * 2;
      img_stride = (size_t)slots * cs * 2;
      return (second ? p.gn_stats1 : p.gn_stats0) + (size_t)b0 * img_stride + (size_t)cl * 2;
    };
    const int last_im = min(IMGS - 1, p.B - 1 - b0);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (tt + 512 * half >= C) break;
      longlong2 v[2][IMGS];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int c = tt + 256 * (2 * half + jj);
        int slots, stride;
        size_t img_stride;
        const long long* base = chan_base(min(c, C - 1), slots, stride, img_stride);
#pragma unroll
        for (int im = 0; im < IMGS; ++im)
          v[jj][im] = c < C ? __ldcg(reinterpret_cast<const longlong2*>(base + (size_t)min(im, last_im) * img_stride))
                            : make_longlong2(0, 0);
      }
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int c = tt + 256 * (2 * half + jj);
        if (c >= C) continue;
        int slots, stride;
        size_t img_stride;
        const long long* base = chan_base(c, slots, stride, img_stride);
#pragma unroll
        for (int im = 0; im < IMGS; ++im) {
          longlong2 a = v[jj][im];
#pragma unroll 1
          for (int sl = 1; sl < slots; ++sl) {      // slot mode only (B200SR3_STAT_SLOTS=1)
            const longlong2 w = __ldcg(reinterpret_cast<const longlong2*>(base + (size_t)min(im, last_im) * img_stride + (size_t)sl * stride));
            a.x += w.x; a.y += w.y;
          }
          gn_tab_put(gtab + im * gn_pitch, c,
                     make_float2(__ll2float_rn(a.x) * (1.0f / 16777216.0f), __ll2float_rn(a.y) * (1.0f / 16777216.0f)));
        }
      }
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");
    if (tt < IMGS * p.gn_groups) {
      int im = 0, g = tt;
      while (g >= p.gn_groups) { g -= p.gn_groups; ++im; }
      float a = 0.f, d = 0.f;
      for (int k = 0; k < cg; ++k) {
        const int c = g * cg + k;
        const float2 sv = gn_tab_get(gtab + im * gn_pitch, c);
        a += sv.x; d += sv.y;
      }
      const float inv_n = 1.0f / ((float)(p.H * p.W) * (float)cg);
      const float mean = a * inv_n;
      gstat[tt] = make_float2(mean, rsqrtf(fmaxf(d * inv_n - mean * mean, 0.f) + 1e-5f));
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");
    const float hs = do_swish ? 0.5f : 1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = tt + 256 * j;
      if (c < C) {
        int g = __float2int_rz(((float)c + 0.5f) * inv_cg);      // c / cg (exact: c < 1024, cg <= 64)
#pragma unroll
        for (int im = 0; im < IMGS; ++im) {
          const float2 mr = gstat[im * p.gn_groups + g];
          float2 v;
          v.x = mr.y * gm[j];
          v.y = bt[j] - mr.x * v.x;
          v.x *= hs; v.y *= hs;
          gn_tab_put(gtab + im * gn_pitch, c, v);
        }
      }
    }
  } else if (p.gn_stats0) {
    const int C = p.gn_C, C0 = p.gn_C0;
    const int cg = C / p.gn_groups;
    constexpr int TPG = IMGS == 1 ? 8 : (IMGS == 2 ? 4 : 1);
    constexpr int KEEP = 4;                       // channels per lane whose gamma / beta stay in registers
    const int gi = tt / TPG, sub = tt - gi * TPG;
    const bool live = gi < IMGS * p.gn_groups;
    int im = 0, g = live ? gi : 0;
    while (g >= p.gn_groups) { g -= p.gn_groups; ++im; }
    const int b = min(b0 + im, p.B - 1);
    float a = 0.f, d = 0.f, gam[KEEP], bet[KEEP];
#pragma unroll
    for (int q = 0; q < KEEP; ++q) gam[q] = bet[q] = 0.f;
    // statistics of channel c: pointer to slot 0 of image b, slot count and slot stride (in int64 units)
    auto chan_ptr = [&](int c, int& slots, int& stride) {
      const bool second = c >= C0;
      const int cs = second ? C - C0 : C0, cl = second ? c - C0 : c;
      slots = second ? p.gn_slots1 : p.gn_slots0;
      stride = cs * 2;
      return (second ? p.gn_stats1 : p.gn_stats0) + ((size_t)b * slots * cs + cl) * 2;
    };
    auto finish = [&](longlong2 v, const long long* sp, int slots, int stride, float& fa, float& fd) {
      long long sa = v.x, sd = v.y;
#pragma unroll 1      // slots beyond the first (rare, 1-3): a rolled loop - unrolled 16-way it put ~500 instructions here
      for (int sl = 1; sl < slots; ++sl) {
        sp += stride;
        const longlong2 w = __ldcg(reinterpret_cast<const longlong2*>(sp));
        sa += w.x; sd += w.y;
      }
      // int64 -> float directly: one rounding, then an exact power-of-two scale (the same bits as going through double)
      fa = __ll2float_rn(sa) * (1.0f / 16777216.0f);
      fd = __ll2float_rn(sd) * (1.0f / 16777216.0f);
    };
    if (live) {
      longlong2 v0[KEEP];
      const long long* sp[KEEP];
      int nsl[KEEP], str[KEEP];
#pragma unroll
      for (int q = 0; q < KEEP; ++q) {          // slot 0, gamma and beta of the first KEEP channels: all in flight together
        const int k = sub + q * TPG;
        v0[q] = make_longlong2(0, 0); sp[q] = nullptr; nsl[q] = 0; str[q] = 0;
        if (k < cg) {
          const int c = g * cg + k;
          sp[q] = chan_ptr(c, nsl[q], str[q]);
          v0[q] = __ldcg(reinterpret_cast<const longlong2*>(sp[q]));
          gam[q] = __ldg(p.gn_gamma + c); bet[q] = __ldg(p.gn_beta + c);
        }
      }
#pragma unroll
      for (int q = 0; q < KEEP; ++q) {
        if (sub + q * TPG < cg) {
          float fa, fd;
          finish(v0[q], sp[q], nsl[q], str[q], fa, fd);
          a += fa; d += fd;
        }
      }
      for (int k = sub + KEEP * TPG; k < cg; k += TPG) {
        int ns, st;
        const long long* ptr = chan_ptr(g * cg + k, ns, st);
        float fa, fd;
        finish(__ldcg(reinterpret_cast<const longlong2*>(ptr)), ptr, ns, st, fa, fd);
        a += fa; d += fd;
      }
    }
#pragma unroll
    for (int o = TPG >> 1; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    if (live) {
      const float inv_n = 1.0f / ((float)(p.H * p.W) * (float)cg);
      const float mean = a * inv_n;
      const float var = fmaxf(d * inv_n - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      const float hs = do_swish ? 0.5f : 1.0f;
      auto put = [&](int c, float gm, float bt) {
        float2 v;
        v.x = rstd * gm;
        v.y = bt - mean * v.x;
        v.x *= hs; v.y *= hs;
        gn_tab_put(gtab + im * gn_pitch, c, v);
      };
#pragma unroll
      for (int q = 0; q < KEEP; ++q) {
        const int k = sub + q * TPG;
        if (k < cg) put(g * cg + k, gam[q], bet[q]);
      }
      for (int k = sub + KEEP * TPG; k < cg; k += TPG) put(g * cg + k, __ldg(p.gn_gamma + g * cg + k), __ldg(p.gn_beta + g * cg + k));
    }
  } else if (p.gn) {
    for (int idx = tt; idx < IMGS * p.gn_C; idx += 256) {
      const int im = idx / p.gn_C;
      const int c = idx - im * p.gn_C;
      float2 v = __ldg(p.gn + (size_t)min(b0 + im, p.B - 1) * p.gn_b_stride + c);
      if (do_swish) { v.x *= 0.5f; v.y *= 0.5f; }
      gn_tab_put(gtab + im * gn_pitch, c, v);
    }
  }
  asm volatile("bar.sync 2, 256;" ::: "memory");
}

// CG = 2: two CTAs of a cluster (one TPC) run ONE M = 256 tile pair with tcgen05 cta_group::2. Each CTA loads its own
// halo, transforms it, and drains its own 128 accumulator rows, but only HALF of every weight tile: the pair shares the
// B operand, which halves the weight bytes through each CTA's shared-memory port and cuts the operand bytes a CTA feeds
// per MMA from (4 + N/32) KB to (4 + N/64) KB. The leader (rank 0) issues the MMAs; its barriers collect both CTAs'
// weight loads (TMA in cta_group::2 form signals the leader's barrier), transform arrivals and epilogue releases;
// tcgen05.commit multicasts the "slot free" / "accumulator ready" arrivals to both CTAs.
template <int BLOCK_N, int MT, bool FUSE_GN, int GEO, int CG = 1, bool HEAD = false, bool DEEP = false, bool PRELU = false>
__global__ void __launch_bounds__(halo_threads(BLOCK_N), 1)
conv_halo_kernel(const __grid_constant__ ConvHaloParams p) {
  using S = HaloSmem<BLOCK_N, MT, GEO, CG, DEEP>;
  static_assert(!PRELU || (FUSE_GN && GEO != 2 && CG == 1 && !HEAD && !DEEP && BLOCK_N != 16), "affine + PReLU transform: plain one- / two-image shapes");
  static_assert(!DEEP || (GEO == 0 && MT == 1 && CG == 1 && !HEAD && (BLOCK_N == 64 || BLOCK_N == 128)), "deep ring: one-tile shapes");
  static_assert(CG == 1 || (CG == 2 && MT == 1 && GEO == 0 && BLOCK_N >= 64), "CTA pairs run one 8x16 tile per CTA");
  static_assert(!HEAD || (!FUSE_GN && GEO == 0 && CG == 1 && BLOCK_N == 64), "the head conv is a raw 3x3, Cout = 64");
  constexpr bool XF = FUSE_GN || CG == 2 || HEAD;      // transform warps active (in a pair they also forward "halo landed" to the leader)
  constexpr int KSTEPS = HEAD ? 2 : 4;         // K = 16 steps per 64-channel block: the head's split operand fills 3 * 8 <= 32 channels
  using G = HaloGeo<GEO>;
  static_assert(GEO == 0 || (BLOCK_N != 16 && MT == 1), "the multi-image geometries run plain convs, one tile at a time");
  static_assert(GEO != 2 || BLOCK_N == 64, "the 4x4 geometry runs BLOCK_N = 64");
  constexpr int ESETS = S::ESETS;
  // warp groups: 0 = transform A, 1 = epilogue A, 2 = transform B, [3 = epilogue B], last = single-thread roles
  constexpr int LW = 4 * (2 + ESETS);        // first warp of the last group: LW+0 halo producer, +1 weight producer,
                                              // +2 TMEM allocator, +3 MMA issuer
  constexpr int AST = S::AST, WST = S::W_STAGES;
  constexpr int NBUF = 2;
  static_assert(NBUF * MT * BLOCK_N <= 512, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + S::BAR_OFFSET;
  // barriers: a_full[AST], a_ready[AST], a_empty[AST], w_full[WST], w_empty[WST], tmem_full[2], tmem_empty[2]
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_ready = [&](int s) { return bar_base + 8u * (AST + s); };
  auto a_empty = [&](int s) { return bar_base + 8u * (2 * AST + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (3 * AST + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (3 * AST + WST + s); };
  auto tmem_full = [&](int b) { return bar_base + 8u * (3 * AST + 2 * WST + b); };
  auto tmem_empty = [&](int b) { return bar_base + 8u * (3 * AST + 2 * WST + NBUF + b); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * AST + 2 * WST + 2 * NBUF);
  static_assert(8 * (3 * AST + 2 * 12 + 2 * NBUF) + 4 <= 512, "barrier area");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  {
    // Touch every 64-byte line of the kernel parameters at once: the constant cache starts cold at every launch, and the
    // roles would otherwise take its misses one after the other (tile decode, segment list, GroupNorm fields, ...).
    const int* pw = reinterpret_cast<const int*>(&p);
    int touch = 0;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(ConvHaloParams) / 64); ++i) touch += pw[i * 16];
    asm volatile("" ::"r"(touch));
  }
  const long long t_entry = HALO_DBG ? clock64() : 0;      // role timing: kernel entry on this SM (after the parameter touch)

  // contiguous run of super tiles; order: x tile fastest, y tile, parity, image, n tile. (32-bit arithmetic: the host
  // checks total_super * grid < 2^31; a 64-bit division is a ~500-cycle subroutine in front of the first TMA load.)
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;      // rank in the CTA pair; tile = 2 * super + rank
  const int unit_id = (int)blockIdx.x / CG, num_units = (int)gridDim.x / CG;
  const int sup_begin = (int)(((uint32_t)unit_id * (uint32_t)p.total_super) / (uint32_t)num_units);
  const int sup_end = (int)(((uint32_t)(unit_id + 1) * (uint32_t)p.total_super) / (uint32_t)num_units);
  struct Tile { int n_tile, x0, y0, b, par, unit; };
  // n / d and n % d for n < 2^22 through the host's float reciprocal (quotient off by at most one, then corrected): an
  // integer division by a run-time value is ~150 cycles of dependent instructions, and a decode chains four of them in
  // front of the first TMA load of every role
  auto divmod = [](uint32_t n, uint32_t d, float inv, uint32_t& r) {
    uint32_t q = __float2uint_rz(__uint2float_rn(n) * inv);
    int rr = (int)(n - q * d);
    if (rr < 0) { --q; rr += (int)d; } else if ((uint32_t)rr >= d) { ++q; rr -= (int)d; }
    r = (uint32_t)rr;
    return q;
  };
  auto decode = [&](int sup, int mt) {
    uint32_t tile = (uint32_t)((sup * MT + mt) * CG) + crank, r;
    Tile t;
    tile = divmod(tile, (uint32_t)p.tiles_w, p.inv_tiles_w, r); t.x0 = (int)r * HALO_TW;
    tile = divmod(tile, (uint32_t)p.tiles_h, p.inv_tiles_h, r); t.y0 = (int)r * HALO_TH;
    tile = divmod(tile, (uint32_t)p.num_par, p.inv_num_par, r); t.par = (int)r;
    tile = divmod(tile, (uint32_t)p.units, p.inv_units, r); t.b = (int)r * G::IMGS; t.unit = (int)r;
    t.n_tile = (int)tile;
    return t;
  };
  // Every role walks the same tile sequence; after the first decode the walk is carried forward with compares instead of
  // being decoded again (a decode is ~80 dependent instructions, and the epilogue ran three of them per super tile).
  auto tile_next = [&](Tile& t) {
#pragma unroll
    for (int s = 0; s < CG; ++s) {
      t.x0 += HALO_TW;
      if (t.x0 == p.tiles_w * HALO_TW) {
        t.x0 = 0; t.y0 += HALO_TH;
        if (t.y0 == p.tiles_h * HALO_TH) {
          t.y0 = 0;
          if (++t.par == p.num_par) {
            t.par = 0; t.b += G::IMGS;
            if (++t.unit == p.units) { t.unit = 0; t.b = 0; ++t.n_tile; }
          }
        }
      }
    }
  };

  const Tile tile0 = decode(sup_begin, 0);      // decoded once; every role starts its walk from it

  // ---- prologue. Each single-thread role initialises the barriers it produces into, so the two TMA producers can start
  // loading before the CTA-wide rendezvous (they only ARRIVE on it): measured on B200, the first global access of a
  // kernel takes ~3 k cycles (caches and TLBs start cold at every launch) and used to begin after ~2 k cycles of setup.
  // One halo stage = the MT tiles' boxes of one 64-channel block, all completing on a_full(stage).
  auto issue_halo = [&](int stage, int map, int cb, const Tile* t) {
    if (HEAD) {      // nothing to fetch: "full" only says that the slot is free for the transform warps to fill
      ptx::mbar_arrive(a_full(stage));
      return;
    }
    ptx::mbar_expect_tx(a_full(stage), HALO_ABLATE(8) ? 0 : MT * G::BYTES);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if HALO_ABLATE(8) break;
      if (GEO == 0)
        ptx::tma_load_4d(smem_base + stage * S::A_STAGE + m * G::STRIDE, &p.a_map[map], a_full(stage),
                         cb * CONV_BLOCK_K, t[m].x0 - 1, t[m].y0 - 1, t[m].b);
      else if (GEO == 2)      // natural (C, W, H, B) order: the 5 x 5 grids of five images, end to end
        ptx::tma_load_4d(smem_base + stage * S::A_STAGE + m * G::STRIDE + G::TMA_OFF, &p.a_map[map], a_full(stage),
                         cb * CONV_BLOCK_K, -1, -1, t[m].b);
      else      // view ordered (C, W, B, H): both images of the pair in one box
        ptx::tma_load_4d(smem_base + stage * S::A_STAGE + m * G::STRIDE, &p.a_map[map], a_full(stage),
                         cb * CONV_BLOCK_K, -1, t[m].b, -1);
    }
  };
  // The first halo tile is on the path to the first MMA (it lands ~3.5 k cycles after it is requested, then has to be
  // transformed): the producer thread requests it before anything else - its own barriers, one decode, one TMA.
  bool halo0_issued = false;
  if (warp == LW) {
    halo0_issued = CG == 1 && !p.pdl && sup_begin < sup_end;      // uniform over the warp
    if (lane == 0) {
      if (!HEAD) ptx::prefetch_tmap(&p.a_map[p.seg[0].map]);
      for (int s = 0; s < AST; ++s) {
        ptx::mbar_init(a_full(s), 1);
        ptx::mbar_init(a_ready(s), CG == 2 ? 16 : 256);      // pair: one arrival per transform warp of both CTAs
        ptx::mbar_init(a_empty(s), 1);
      }
      ptx::fence_barrier_init();
      if (halo0_issued) {
        Tile t[MT];
#pragma unroll
        Tile w0 = tile0;
#pragma unroll
        for (int m = 0; m < MT; ++m) { t[m] = w0; tile_next(w0); }
        issue_halo(0, p.seg[0].map, 0, t);
      }
      if (!HEAD) for (int i = 0; i < p.num_segs; ++i) ptx::prefetch_tmap(&p.a_map[i]);
    }
    __syncwarp();
  }
  if (warp == LW + 1 && lane == 0) ptx::prefetch_tmap(&p.w_map);
  if (warp == 4 && lane == 0)
    for (int i = 0; i < p.num_par; ++i) ptx::prefetch_tmap(&p.o_map[i]);
  if (warp == LW + 1 && lane == 0) {
    for (int s = 0; s < WST; ++s) {
      ptx::mbar_init(w_full(s), 1);
      ptx::mbar_init(w_empty(s), 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == LW + 2) {
    if (lane == 0) {
      for (int b = 0; b < NBUF; ++b) {
        ptx::mbar_init(tmem_full(b), 1);
        ptx::mbar_init(tmem_empty(b), CG == 2 ? 8 : 128 * ESETS);   // pair: one arrival per epilogue warp of both CTAs
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    if (CG == 2) tmem_alloc_pair(tmem_slot, NBUF * MT * BLOCK_N);
    else ptx::tmem_alloc(tmem_slot, NBUF * MT * BLOCK_N);
  }
  // Transform warps build the first image's GroupNorm table NOW, before the CTA-wide rendezvous and before the TMA
  // producers queue the first halo and weight tiles of all 148 CTAs: measured on B200, a global load issued once that
  // burst is in flight takes ~3 k cycles instead of ~1 k, and the table sat on the path to the first MMA. (Not under
  // programmatic dependent launch: the statistics may not be final before griddepcontrol.wait.)
  int tab_b_early = -1;
  if (FUSE_GN && CG == 1 && !p.pdl && (warp < 4 || (warp >= 8 && warp < 12)) && sup_begin < sup_end) {
    const int tt = warp < 4 ? (int)threadIdx.x : (int)threadIdx.x - 128;
    tab_b_early = tile0.b;
    halo_build_gn_table<G::IMGS>(p, reinterpret_cast<float2*>(smem_gen + S::GN_OFFSET), p.gn_C + 2 * (p.gn_C >> 3), tt,
                                 tab_b_early, !PRELU && p.gn_swish != 0, t_entry);
    if (HALO_DBG && tt == 0) p.dbg[blockIdx.x * 16 + 3] = (unsigned long long)(clock64() - t_entry);   // [3] first table ready
  }
  if (GEO == 2) {
    // pixels 125 .. 134 of every stage: the zero row below the last image (TMA never writes them, nobody else does)
    for (int idx = (int)threadIdx.x; idx < AST * 10 * 8; idx += (int)blockDim.x) {
      const int st = idx / 80, rem = idx - st * 80;
      *reinterpret_cast<uint4*>(smem_gen + st * S::A_STAGE + G::TMA_OFF + (G::NPIX + rem / 8) * 128 + (rem % 8) * 16) =
          make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  if (CG == 2) {
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer's barriers are initialised before anything arrives on them
    ptx::tc_fence_after();
  } else if (warp == LW || warp == LW + 1) {
    asm volatile("bar.arrive 3, %0;" ::"n"(halo_threads(BLOCK_N)) : "memory");      // producers do not wait
  } else {
    ptx::tc_fence_before();
    asm volatile("bar.sync 3, %0;" ::"n"(halo_threads(BLOCK_N)) : "memory");
    ptx::tc_fence_after();
  }
  // Everything above overlapped the previous kernel's tail (programmatic dependent launch); its results are needed from
  // here on - except by the weight producer, which only reads constants and starts filling its ring at once.
  if (p.pdl && warp != LW + 1) pdl_wait();
  uint32_t tmem_base = 0;
  if (CG == 2 || (warp != LW && warp != LW + 1)) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (HALO_DBG && threadIdx.x == 0) p.dbg[blockIdx.x * 16 + 15] = (unsigned long long)(clock64() - t_entry);   // [15] prologue

  if (warp == LW) {
    if (ptx::elect_one()) {
      // ---------------------------------------------------------------- halo producer
      int as = 0; uint32_t aphase = 0;
      HDBG_DECL();
      const long long hd_start = HALO_DBG ? clock64() : 0;
      Tile walk = tile0;
      for (int sup = sup_begin; sup < sup_end; ++sup) {
        Tile t[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) { t[m] = walk; tile_next(walk); }
        for (int sg = 0; sg < p.num_segs; ++sg) {
          const HaloSeg seg = p.seg[sg];
          for (int cb = 0; cb < seg.cblocks; ++cb) {
            if (!(halo0_issued && sup == sup_begin && sg == 0 && cb == 0)) {      // (that one went out in the prologue)
              HDBG_T0();
              ptx::mbar_wait(a_empty(as), aphase ^ 1u);
              HDBG_ACC(0);
              issue_halo(as, seg.map, cb, t);
            }
            if (++as == AST) { as = 0; aphase ^= 1u; }
          }
        }
      }
      if (HALO_DBG) { hd[1] = (unsigned long long)(clock64() - hd_start); hd[2] = (unsigned long long)(sup_end - sup_begin); }
      HDBG_FLUSH(0, 3);      // [0] A producer waits a_empty, [1] total, [2] super tiles
    }
  } else if (warp == LW + 1) {
    if (ptx::elect_one()) {
      // ---------------------------------------------------------------- weight producer
      // The ring is NOT filled at once: a full ring is 100-400 KB per SM queued in front of the first halo tile and of
      // the GroupNorm statistics (measured: every other first access of the kernel then takes ~3 k cycles instead of
      // ~1 k). Two stages go out, the rest follows once the first halo tile has landed (a_full(0), phase 0 - a
      // non-consuming wait; its second completion needs MMAs that need more than two weight stages, so it cannot pass us).
      int ws = 0; uint32_t wphase = 0;
      int early = CG == 1 ? 2 : -1;
      Tile walk = tile0;
      for (int sup = sup_begin; sup < sup_end; ++sup) {
        const Tile t = walk;
#pragma unroll
        for (int m = 0; m < MT; ++m) tile_next(walk);
        const int wrow = t.par * p.Cout + t.n_tile * BLOCK_N + (int)crank * (BLOCK_N / CG);
        for (int sg = 0; sg < p.num_segs; ++sg) {
          const HaloSeg seg = p.seg[sg];
          for (int cb = 0; cb < seg.cblocks; ++cb) {
            for (int tap = 0; tap < seg.ntaps; ++tap) {
              if (early == 0) ptx::mbar_wait(a_full(0), 0u);
              if (early >= 0) --early;
              ptx::mbar_wait(w_empty(ws), wphase ^ 1u);
              if (CG == 2) {
                // both halves of the tile complete on the LEADER's barrier, which the leader arms for both
                if (crank == 0) ptx::mbar_expect_tx(w_full(ws), CG * S::W_STAGE);
                tma_load_2d_pair(smem_base + S::W_OFFSET + ws * S::W_STAGE, &p.w_map, w_full(ws) & HALO_PEER_MASK,
                                 seg.k_base + tap * seg.k_tap_stride + cb * CONV_BLOCK_K, wrow);
              } else {
                ptx::mbar_expect_tx(w_full(ws), HALO_ABLATE(4) ? 0 : S::W_STAGE);
                if (!HALO_ABLATE(4))
                  ptx::tma_load_2d(smem_base + S::W_OFFSET + ws * S::W_STAGE, &p.w_map, w_full(ws),
                                   seg.k_base + tap * seg.k_tap_stride + cb * CONV_BLOCK_K, wrow);
              }
              if (++ws == WST) { ws = 0; wphase ^= 1u; }
            }
          }
        }
        if (p.w_resident) break;      // every later super tile finds its weight tiles where the first one left them
      }
    }
  } else if (warp == LW + 3) {
    // ------------------------------------------------------------------ MMA issuer
    // One elected thread runs the whole loop. A single thread retires dependent scalar instructions at
    // ~5 cycles each while an N=64 MMA occupies the tensor pipe for only 48 cycles, so the loop is kept
    // minimal: descriptors are (constant high word, low word advanced by adds), the tap window walks
    // the halo tile incrementally, and the NEXT weight stage's barrier is probed before this tap's
    // MMAs are issued so that its latency overlaps them.
    if (crank == 0 && ptx::elect_one()) {
      constexpr uint32_t A_HI = (uint32_t)((G::GROUP * 128) >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t B_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      const uint32_t idesc = ptx::make_idesc_bf16(CONV_BLOCK_M * CG, BLOCK_N);
      const uint32_t a_lo0 = (((smem_base + G::MMA_OFF) & 0x3FFFFu) >> 4) | 0x10000u;
      const uint32_t w_lo0 = (((smem_base + S::W_OFFSET) & 0x3FFFFu) >> 4) | 0x10000u;
      auto desc = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
      int as = 0, ws = 0;
      uint32_t aphase = 0, wphase = 0;
      bool ready = false;
      int it = 0;
      HDBG_DECL();
      long long t_first_mma = 0;
      const bool dbg_on = HALO_DBG;
      uint32_t b_cur = w_lo0, w_full_cur = w_full(0), w_empty_cur = w_empty(0);      // running per-stage values of ws
      const int wres = p.w_resident;
      const int wst_eff = wres ? wres : WST;      // resident: stage j holds tile j of EVERY super tile, phase 0 stays complete
      Tile walk = tile0;
      for (int sup = sup_begin; sup < sup_end; ++sup, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        const int par = walk.par;
#pragma unroll
        for (int m = 0; m < MT; ++m) tile_next(walk);
        HDBG_T0();
        ptx::mbar_wait(tmem_empty(buf), (use & 1u) ^ 1u);
        HDBG_ACC(2);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * MT * BLOCK_N);
        uint32_t accum = 0;
        for (int sg = 0; sg < p.num_segs; ++sg) {
          const int ntaps = p.seg[sg].ntaps, cblocks = p.seg[sg].cblocks;
          const int ntx = p.seg[sg].tap_w;
          const uint32_t pix0 = p.seg[sg].pix0 >= 0 ? (uint32_t)p.seg[sg].pix0 : (uint32_t)((par >> 1) * G::PITCH + (par & 1));
          for (int cb = 0; cb < cblocks; ++cb) {
            HDBG_T0();
            ptx::mbar_wait(XF ? a_ready(as) : a_full(as), aphase);
            HDBG_ACC(0);
            uint32_t a_lo = a_lo0 + (uint32_t)as * (uint32_t)(S::A_STAGE >> 4) + pix0 * 8u;
            int tx = 0;
            if (dbg_on && t_first_mma == 0) t_first_mma = clock64();
            // The tap loop is THE critical instruction stream of the Cout = 64 layers: eight 48-cycle MMAs per tap leave
            // ~380 cycles, and one thread retires ~5 cycles per dependent instruction. Everything that is not an MMA is
            // kept to running pointers (no multiplications, no parameter reloads, no timing code).
            for (int tap = 0; tap < ntaps; ++tap) {
              if (!ready) {
                if (dbg_on) { HDBG_T0(); ptx::mbar_wait(w_full_cur, wphase); HDBG_ACC(1); }
                else ptx::mbar_wait(w_full_cur, wphase);
              }
              ptx::tc_fence_after();
              const uint32_t b_lo = b_cur;
              const uint32_t wcur = w_empty_cur;
              b_cur += (uint32_t)(S::W_STAGE >> 4); w_full_cur += 8u; w_empty_cur += 8u;
              if (++ws == wst_eff) { ws = 0; if (!wres) wphase ^= 1u; b_cur = w_lo0; w_full_cur = w_full(0); w_empty_cur = w_empty(0); }
              ready = ptx::mbar_test_wait(w_full_cur, wphase);      // look one stage ahead
#pragma unroll
              for (int m = 0; m < MT; ++m) {
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  if (CG == 2)
                    umma_bf16_pair(d_tmem + (uint32_t)(m * BLOCK_N), desc(a_lo + (uint32_t)(m * (G::STRIDE >> 4) + 2 * k), A_HI),
                                   desc(b_lo + 2u * k, B_HI), idesc, accum | (uint32_t)k);
                  else
                    ptx::umma_bf16(d_tmem + (uint32_t)(m * BLOCK_N), desc(a_lo + (uint32_t)(m * (G::STRIDE >> 4) + 2 * k), A_HI),
                                   desc(b_lo + 2u * k, B_HI), idesc, accum | (uint32_t)k);
                }
              }
              if (CG == 2) umma_commit_pair(wcur); else if (!wres) ptx::umma_commit(wcur);
              accum = 1;
              a_lo += 8u;
              if (++tx == ntx) { tx = 0; a_lo += (uint32_t)(G::PITCH - ntx) * 8u; }
            }
            if (CG == 2) umma_commit_pair(a_empty(as)); else ptx::umma_commit(a_empty(as));
            if (++as == AST) { as = 0; aphase ^= 1u; }
          }
        }
        if (CG == 2) umma_commit_pair(tmem_full(buf)); else ptx::umma_commit(tmem_full(buf));
      }
      // all of this CTA's MMAs are issued: only the last epilogue remains, let the next kernel's CTAs
      // be scheduled (they run their prologue and block in pdl_wait until this grid has completed)
      pdl_launch_dependents();
      HDBG_FLUSH(4, 3);      // [4] MMA waits A ready, [5] waits W full, [6] waits TMEM empty
      if (HALO_DBG) {           // [12] kernel entry -> first MMA issued ("fill"), [13] entry -> last MMA issued
        p.dbg[blockIdx.x * 16 + 12] = (unsigned long long)(t_first_mma - t_entry);
        p.dbg[blockIdx.x * 16 + 13] = (unsigned long long)(clock64() - t_entry);
      }
    }
  } else if (warp == LW + 2) {
    // ------------------------------------------------------------------ (TMEM allocator, otherwise idle) L2 prefetch
    if (p.pf_ptr != nullptr && lane == 0) {
      const unsigned off = blockIdx.x * p.pf_slice;
      if (off < p.pf_total) {
        __nanosleep(2000);      // behind this kernel's own first loads
        const unsigned n = min(p.pf_slice, p.pf_total - off) & ~15u;
        for (unsigned o = 0; o < n; o += 32768u) {
          const unsigned sz = min(32768u, n - o);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.pf_ptr + off + o), "r"(sz) : "memory");
        }
      }
    }
  } else if (BLOCK_N == 16 && warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ tail epilogue: sampler update
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const int lx = row & 7, ly = row >> 3;
    const int OC = p.tail_oc;
    const size_t plane = (size_t)p.H * p.W;
    float a = 0.f, bc = 0.f, c1 = 0.f, c2 = 0.f, sigma = 0.f, lim = 1.0f;
    int ts = 0, T = 0, mode = 0;
    if (p.tail_x) {
      ts = p.ctl->t; T = p.ctl->T; mode = p.ctl->noise_mode;
      if (p.ctl->no_clip) lim = __int_as_float(0x7f800000);      // clip_denoised=False: clamp to +-inf
      a = p.coefs[ts]; bc = p.coefs[T + ts]; c1 = p.coefs[2 * T + ts]; c2 = p.coefs[3 * T + ts];
      sigma = expf(0.5f * p.coefs[4 * T + ts]);
    }
    int it = 0;
    HDBG_DECL();
    Tile walk = tile0;
    const long long row0 = p.tail_x ? p.ctl->row0 : 0;
    const unsigned long long seed = p.tail_x ? p.ctl->seed : 0ull;
    const float* zp = nullptr;                      // injected noise of this step (modes 1 and 3), null: none
    if (p.tail_x && ts > 0 && (mode == 1 || mode == 3) && p.ctl->noise)
      zp = p.ctl->noise + (mode == 1 ? (size_t)(T - ts) * (size_t)p.ctl->numel : (size_t)0);
    float bias_o[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) bias_o[o] = (p.bias && o < OC) ? __ldg(p.bias + o) : 0.f;
    for (int sup = sup_begin; sup < sup_end; ++sup, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      // Everything that does not depend on the accumulator happens BEFORE the wait for it: the state x and the injected
      // noise of this thread's MT pixels are requested (L2 round trips of ~1 k cycles in situ; read inside the update
      // loop they were serialised behind each other by the in-place stores) and the Philox / Box-Muller draws are made.
      size_t base[MT];
      float xv[MT][4], z[MT][4];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const Tile t = walk;
        tile_next(walk);
        const int y = t.y0 + ly, x = t.x0 + lx;
        base[m] = (size_t)t.b * OC * plane + (size_t)y * p.W + x;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          xv[m][o] = (p.tail_x && o < OC) ? p.tail_x[base[m] + o * plane] : 0.f;
          z[m][o] = (zp && o < OC) ? __ldg(zp + base[m] + o * plane) : 0.f;
        }
        if (p.tail_x && ts > 0 && mode == 2) {
          const long long pix = ((row0 + (long long)t.b) * p.H + y) * p.W + x;      // global row: sharding-invariant noise
          const uint4 r = philox4x32_10(make_uint4((uint32_t)pix, (uint32_t)(pix >> 32), (uint32_t)ts, 0x5352u),
                                        make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
          const float2 g0 = box_muller(r.x, r.y), g1 = box_muller(r.z, r.w);
          z[m][0] = g0.x; z[m][1] = g0.y; z[m][2] = g1.x; z[m][3] = g1.y;
        }
      }
      HDBG_T0();
      ptx::mbar_wait(tmem_full(buf), use & 1u);
      HDBG_ACC(0);
      ptx::tc_fence_after();
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((buf * MT + m) * BLOCK_N), v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          if (o < OC) {
            const float eps = __uint_as_float(v[o]) + bias_o[o];
            if (p.tail_eps) p.tail_eps[base[m] + o * plane] = eps;
            if (p.tail_x) p.tail_x[base[m] + o * plane] = posterior_update(xv[m][o], eps, z[m][o], a, bc, c1, c2, sigma, lim);
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(tmem_empty(buf));
    }
    if (threadIdx.x == 128) HDBG_FLUSH(8, 1);
  } else if (BLOCK_N != 16 && ((warp >= 4 && warp < 8) || (ESETS == 2 && warp >= 12 && warp < 16))) {
    // ------------------------------------------------------------------ epilogue
    constexpr int NCH = BLOCK_N >= 64 ? BLOCK_N / 64 : 1;
    constexpr int NSTG = S::NSTG;
    const int wq = warp & 3;
#define HALO_EPI_SYNC() asm volatile("bar.sync 1, %0;" ::"n"(128 * ESETS) : "memory")
    const int eset = warp >= 12 ? 1 : 0;                   // which epilogue set this warp belongs to
    const int tid_e = (int)threadIdx.x - (eset ? 256 : 128);     // 0 .. 128*ESETS-1
    uint8_t* slab_gen = smem_gen + S::STG_OFFSET + (eset * 4 + wq) * (NSTG * 4096);
    int unit = 0;                                          // 64-channel units seen so far (same count in every set)
    const bool do_stats = p.stat_partial != nullptr;
    const float* bias = p.bias;
    if (bias && p.bias_t_stride) bias += (size_t)p.ctl->t * p.bias_t_stride;
    // the n tile's bias row sits in shared memory: per 64-channel chunk a warp reads it with 16 broadcast LDS.128 instead of
    // 16 global loads, each of which also cost a pair of descriptor moves (R2UR) inside this divergent region
    float* bias_s = reinterpret_cast<float*>(smem_gen + S::GN_OFFSET + S::BIAS_OFFSET_IN_GN);
    int bias_n0 = -1;
    constexpr int IMGS = G::IMGS;
    // this lane's columns (2l, 2l+1) of each 64-channel chunk, per image of the tile: sum0, sum1, sq0, sq1
    constexpr int NACC = GEO == 2 ? 2 : IMGS;      // GEO 2: a warp's rows touch at most two images
    long long acc[NACC][NCH][4];
#pragma unroll
    for (int im = 0; im < NACC; ++im)
#pragma unroll
      for (int c = 0; c < NCH; ++c) acc[im][c][0] = acc[im][c][1] = acc[im][c][2] = acc[im][c][3] = 0;
    // where lane l finds columns (2l, 2l+1) of slab row r: chunk (l >> 2) ^ (r & 7), word l & 3
    const uint32_t col_chunk = (uint32_t)(lane >> 2), col_word = (uint32_t)(lane & 3) * 4u;

    // GEO 2: this warp's 32 accumulator rows are linear pixels 32 wq .. 32 wq + 31 of the five 5x5 grids: they belong to
    // image g2_first (rows < g2_rb) and image g2_first + 1, and only interior pixels (bit set in g2_valid) are outputs.
    // acc[0] / acc[1] then hold the sums of the warp's first / second image. Interior rows are stored straight from
    // registers (128 contiguous bytes per row and 64-channel chunk): a tensor store whose box starts at (-1,-1) to clip
    // the ring faults on this part ("illegal instruction"), and the level is far too small for the store path to matter.
    int g2_first = 0, g2_rb = 32;
    uint32_t g2_valid = 0u;
    if (GEO == 2) {
      g2_first = (32 * wq) / 25;
      g2_rb = 25 * (g2_first + 1) - 32 * wq;
      for (int r = 0; r < 32; ++r) {
        const int px = 32 * wq + r, q = px % 25;
        if (px < G::NPIX && q >= 5 && q % 5 != 0) g2_valid |= 1u << r;
      }
    }
    auto data_off = [&](int st) { return (eset * 4 + wq) * (NSTG * 4096) + st * 4096; };

    int it = 0, stg = 0;
    HDBG_DECL();
    Tile walk = tile0;
    for (int sup = sup_begin; sup < sup_end; ++sup, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      HDBG_T0();
      ptx::mbar_wait(tmem_full(buf), use & 1u);
      HDBG_ACC(0);
      ptx::tc_fence_after();
      const Tile t0 = walk;
      const int n0 = t0.n_tile * BLOCK_N;
      if (bias && n0 != bias_n0) {          // (uniform over the epilogue warps: they walk the same super tiles)
        HALO_EPI_SYNC();                    // every warp is done with the previous row
        for (int i = tid_e; i < BLOCK_N; i += 128 * ESETS) bias_s[i] = __ldg(bias + n0 + i);
        HALO_EPI_SYNC();
        bias_n0 = n0;
      }
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const Tile t = walk;
        tile_next(walk);      // after the loop: the first tile of the next super tile
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((buf * MT + m) * BLOCK_N);
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
          if (ESETS == 2 && ((unit++ & 1) != eset)) continue;      // the other set's unit
          const uint32_t sl = smem_base + S::STG_OFFSET + (uint32_t)data_off(stg);
          uint8_t* slg = smem_gen + S::STG_OFFSET + data_off(stg);
          // the tensor store that last read this slab must have finished reading it
          if (lane == 0) bulk_wait_read<NSTG - 1>();
          __syncwarp();
          HEPI_T0();
          bf16* g2_dst = nullptr;             // GEO 2: this thread's output row in global memory (null: not an output)
          if (GEO == 2 && ((g2_valid >> lane) & 1u) && !HALO_ABLATE(1)) {
            const int img = g2_first + (lane >= g2_rb ? 1 : 0);
            const int q = (32 * wq + lane) % 25, y = q / 5 - 1, x = q % 5 - 1;
            if (t.b + img < p.B) {
              const int sc = p.num_par == 4 ? 2 : 1;
              const int oy = sc * y + (t.par >> 1), ox = sc * x + (t.par & 1);
              g2_dst = p.out + (((size_t)(t.b + img) * p.out_H + oy) * p.out_W + ox) * p.Cout + n0 + cc * 64;
            }
          }
          // both 32-column halves of the chunk are requested before the one wait (a second wait cost a TMEM round trip)
          // (only where the statistics accumulators leave 64 registers for it: elsewhere the second half is requested
          // once the first has been packed)
          constexpr bool LD64 = GEO != 2 && NACC * NCH <= 2;
          uint32_t v2[2][32];
          if (!HALO_ABLATE(16)) {
            ptx::tmem_ld32(taddr + (uint32_t)(cc * 64), v2[0]);
            if (LD64) ptx::tmem_ld32(taddr + (uint32_t)(cc * 64 + 32), v2[1]);
            ptx::tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) { v2[0][j] = (uint32_t)j; v2[1][j] = (uint32_t)(32 + j); }
          }
          HEPI_ACC(1);
          if (LD64 && ESETS == 1 && m == MT - 1 && cc == NCH - 1) {
            // this thread's last TMEM read of the super tile has completed: hand the accumulator back to the MMA warp
            // NOW, not after the store and the statistics of this chunk (in situ the MMA issuer of the 64 -> 64 layers and
            // of the head waited 17-33 % of its time for this arrival, profiles/r03c_roles_in_situ.txt)
            ptx::tc_fence_before();
            if (CG == 2) {
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(tmem_empty(buf) & HALO_PEER_MASK);      // the leader's barrier
            } else {
              ptx::mbar_arrive(tmem_empty(buf));
            }
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t* v = v2[LD64 ? half : 0];
            if (!LD64 && half == 1) {
              if (!HALO_ABLATE(16)) {
                ptx::tmem_ld32(taddr + (uint32_t)(cc * 64 + 32), v);
                ptx::tmem_ld_wait();
              }
              if (ESETS == 1 && m == MT - 1 && cc == NCH - 1) {
                ptx::tc_fence_before();
                if (CG == 2) {
                  __syncwarp();
                  if (lane == 0) mbar_arrive_cluster(tmem_empty(buf) & HALO_PEER_MASK);
                } else {
                  ptx::mbar_arrive(tmem_empty(buf));
                }
              }
            }
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (bias) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(bias_s + cc * 64 + half * 32 + j);
                add2(f[j], f[j + 1], f[j], f[j + 1], bv.x, bv.y);
                add2(f[j + 2], f[j + 3], f[j + 2], f[j + 3], bv.z, bv.w);
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 pk = pack8(f + 8 * j);
              *reinterpret_cast<uint4*>(slg + lane * 128 + (((half * 4 + j) ^ (lane & 7)) << 4)) = pk;
              if (GEO == 2 && g2_dst) *reinterpret_cast<uint4*>(g2_dst + (half * 4 + j) * 8) = pk;
            }
          }
          HEPI_ACC(2);
          fence_proxy_async_smem();       // generic-proxy writes -> visible to the TMA store
          if (GEO == 2) {
            __syncwarp();                   // rows were stored from registers; the slab only feeds the statistics
          } else {
            __syncwarp();
            if (lane == 0 && !HALO_ABLATE(1)) {
              if (GEO == 0) tma_store_4d(&p.o_map[t.par], sl, n0 + cc * 64, t.x0, t.y0 + 4 * wq, t.b);
              else tma_store_4d(&p.o_map[t.par], sl, n0 + cc * 64, 0, t.b, 2 * wq);      // (C, W, B, H) view
              bulk_commit();
            }
          }
          if (GEO == 2 && do_stats && !HALO_ABLATE(32)) {
            // every element goes to fixed point on its own: which rows of an image a warp sees depends on the image's
            // place in the tile, and float partial sums would make a face's statistics depend on its batch position
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const uint32_t w = *reinterpret_cast<const uint32_t*>(slg + r * 128 + (((col_chunk ^ (uint32_t)(r & 7)) << 4) | col_word));
              const bool valid = (g2_valid >> r) & 1u;           // select, never multiply: border rows may hold NaN
              const float lo = valid ? __uint_as_float(w << 16) : 0.f, hi = valid ? __uint_as_float(w & 0xffff0000u) : 0.f;
              const long long i0 = __float2ll_rn(lo * STAT_FIXED_SCALE), i1 = __float2ll_rn(hi * STAT_FIXED_SCALE);
              const long long i2 = __float2ll_rn(lo * lo * STAT_FIXED_SCALE), i3 = __float2ll_rn(hi * hi * STAT_FIXED_SCALE);
              if (r >= g2_rb) { acc[1][cc][0] += i0; acc[1][cc][1] += i1; acc[1][cc][2] += i2; acc[1][cc][3] += i3; }
              else { acc[0][cc][0] += i0; acc[0][cc][1] += i1; acc[0][cc][2] += i2; acc[0][cc][3] += i3; }
            }
          }
          HEPI_ACC(3);
          if (GEO != 2 && do_stats && !HALO_ABLATE(32)) {
            float s0[IMGS], s1[IMGS], q0[IMGS], q1[IMGS];
#pragma unroll
            for (int im = 0; im < IMGS; ++im) s0[im] = s1[im] = q0[im] = q1[im] = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const int im = (IMGS == 2) ? ((r >> 3) & 1) : 0;      // GEO 1: 8-row groups alternate between the images
              const uint32_t w = *reinterpret_cast<const uint32_t*>(slg + r * 128 + (((col_chunk ^ (uint32_t)(r & 7)) << 4) | col_word));
              const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
              add2(s0[im], s1[im], s0[im], s1[im], lo, hi);
              fma2(q0[im], q1[im], lo, hi, lo, hi, q0[im], q1[im]);
            }
#pragma unroll
            for (int im = 0; im < IMGS; ++im) {
              acc[im][cc][0] += __float2ll_rn(s0[im] * STAT_FIXED_SCALE);
              acc[im][cc][1] += __float2ll_rn(s1[im] * STAT_FIXED_SCALE);
              acc[im][cc][2] += __float2ll_rn(q0[im] * STAT_FIXED_SCALE);
              acc[im][cc][3] += __float2ll_rn(q1[im] * STAT_FIXED_SCALE);
            }
          }
          stg = (stg + 1 == NSTG) ? 0 : stg + 1;
        }
      }
      if (ESETS != 1) {      // (two epilogue sets skip each other's chunks: release after the loop)
        ptx::tc_fence_before();
        if (CG == 2) {
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tmem_empty(buf) & HALO_PEER_MASK);      // the leader's barrier
        } else {
          ptx::mbar_arrive(tmem_empty(buf));
        }
      }

      if (do_stats) {
        const int seg = t0.n_tile * p.units + t0.unit;      // = sup / seg_len_super: a segment is one (n tile, unit)
        if (sup + 1 == sup_end || walk.unit != t0.unit || walk.n_tile != t0.n_tile) {
          // the CTA's run over this (n tile, image) segment ends: publish its partial sums. The four
          // warps' sums meet in the (drained) staging slabs: [column][sum|sq] int64 per warp.
          const uint32_t GU = (uint32_t)num_units, T = (uint32_t)p.total_super;
          const int first_unit = (int)((((uint32_t)seg * (uint32_t)p.seg_len_super + 1u) * GU - 1u) / T);
          const int last_unit = (int)((((uint32_t)(seg + 1) * (uint32_t)p.seg_len_super) * GU - 1u) / T);
          const int slot = (unit_id - first_unit) * CG + (int)crank;       // a pair publishes two slots
          const bool is_last = unit_id == last_unit && (int)crank == CG - 1;
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
          long long* mine = reinterpret_cast<long long*>(slab_gen);
#pragma unroll
          for (int im = 0; im < IMGS; ++im)
#pragma unroll
            for (int cc = 0; cc < NCH; ++cc) {
              const int col = im * BLOCK_N + cc * 64 + 2 * lane;
              long long v[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (GEO == 2) v[k] = im == g2_first ? acc[0][cc][k] : (im == g2_first + 1 ? acc[1][cc][k] : 0ll);
                else v[k] = acc[im < NACC ? im : 0][cc][k];
              }
              *reinterpret_cast<longlong2*>(mine + col * 2) = make_longlong2(v[0], v[2]);
              *reinterpret_cast<longlong2*>(mine + col * 2 + 2) = make_longlong2(v[1], v[3]);
            }
#pragma unroll
          for (int im = 0; im < NACC; ++im)
#pragma unroll
            for (int cc = 0; cc < NCH; ++cc) acc[im][cc][0] = acc[im][cc][1] = acc[im][cc][2] = acc[im][cc][3] = 0;
          HALO_EPI_SYNC();
          for (int item = tid_e; item < IMGS * 2 * BLOCK_N; item += 128 * ESETS) {
            const int im = item / (2 * BLOCK_N), within = item - im * 2 * BLOCK_N;
            if (t0.b + im >= p.B) continue;            // odd batch: the pair's second image does not exist
            long long a = 0;
#pragma unroll
            for (int ww = 0; ww < 4 * ESETS; ++ww)
              a += reinterpret_cast<const long long*>(smem_gen + S::STG_OFFSET + ww * (NSTG * 4096))[item];
            if (p.stat_atomic) {
              atomicAdd(reinterpret_cast<unsigned long long*>(p.stat_partial + ((size_t)(t0.b + im) * p.Cout + n0) * 2 + within),
                        (unsigned long long)a);
              continue;
            }
            long long* dst = p.stat_partial + (((size_t)(t0.b + im) * p.stat_slots + slot) * p.Cout + n0) * 2 + within;
            *dst = a;
            if (is_last)
              for (int sl2 = slot + 1; sl2 < p.stat_slots; ++sl2) dst[(size_t)(sl2 - slot) * p.Cout * 2] = 0;
          }
          HALO_EPI_SYNC();
        }
      }
    }
    if (lane == 0) bulk_wait_all();       // the staging slabs must outlive the stores that read them
#ifdef HALO_EPI_PROFILE
    if (tid_e == 0) HDBG_FLUSH(8, 4);      // [8] waits accumulator, [9] waits slab, [10] load + pack + store, [11] statistics
#else
    if (tid_e == 0) HDBG_FLUSH(8, 1);      // [8] epilogue waits accumulator
#endif
    if (tid_e == 0 && HALO_DBG) p.dbg[blockIdx.x * 16 + 14] = (unsigned long long)(clock64() - t_entry);   // [14] entry -> epilogue done
#undef HALO_EPI_SYNC
  } else if (XF && (warp < 4 || (warp >= 8 && warp < 12))) {
    // ------------------------------------------------------------------ GroupNorm + Swish transform
    // thread -> (16-byte channel chunk j, pixels p_first + 32 i): its 8 channels' (scale, shift) sit in
    // registers for the whole halo tile; a warp touches 4 full 128-byte pixel rows per access. The
    // (scale, shift) table of the current image (pair) is staged in shared memory once, so a stage does
    // not start with an L2 round trip.
    constexpr int IMGS = G::IMGS;
    const int tt = warp < 4 ? (int)threadIdx.x : (int)threadIdx.x - 128;      // 0..255
    const int j = tt & 7;
    const int p_first = tt >> 3;
    constexpr bool do_swish = !PRELU;    // Block = GroupNorm -> Swish -> Conv (unet.py:84-86): the host rejects a fused GroupNorm without Swish
    float2* gtab = reinterpret_cast<float2*>(smem_gen + S::GN_OFFSET);       // [IMGS][gn_C (padded)], halved if swish
    const int gn_pitch = p.gn_C + 2 * (p.gn_C >> 3);
    int tab_b = tab_b_early;
    int as = 0; uint32_t aphase = 0;
    bool first_halo_seen = false;
    HDBG_DECL();
    Tile walk = tile0;
    for (int sup = sup_begin; sup < sup_end; ++sup) {
      Tile t[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) { t[m] = walk; tile_next(walk); }
      if (FUSE_GN && ((PRELU && !p.gn_stats0) ? tab_b < 0 : t[0].b != tab_b)) {      // (a ready affine + PReLU row is the same for every image)
        halo_build_gn_table<G::IMGS>(p, gtab, gn_pitch, tt, t[0].b, do_swish);
        tab_b = t[0].b;
      }
      for (int sg = 0; sg < p.num_segs; ++sg) {
        const HaloSeg seg = p.seg[sg];
        for (int cb = 0; cb < seg.cblocks; ++cb) {
          constexpr int NTAB = GEO == 2 ? 1 : IMGS;      // GEO 2 fetches (scale, shift) per pixel item instead
          float sc[NTAB][8], sh[NTAB][8];
          float sl[PRELU ? 8 : 1];
          if (PRELU && seg.gn_off >= 0) {
            const float4* s4 = reinterpret_cast<const float4*>(p.xf_slope + seg.gn_off + cb * CONV_BLOCK_K + j * 8);
            const float4 s0 = __ldg(s4), s1 = __ldg(s4 + 1);
            sl[0] = s0.x; sl[1 % (PRELU ? 8 : 1)] = s0.y; sl[2 % (PRELU ? 8 : 1)] = s0.z; sl[3 % (PRELU ? 8 : 1)] = s0.w;
            sl[4 % (PRELU ? 8 : 1)] = s1.x; sl[5 % (PRELU ? 8 : 1)] = s1.y; sl[6 % (PRELU ? 8 : 1)] = s1.z; sl[7 % (PRELU ? 8 : 1)] = s1.w;
          }
          HDBG_T0();
          if (seg.gn_off >= 0 && GEO != 2) {
#pragma unroll
            for (int im = 0; im < NTAB; ++im) {
              const int c0 = seg.gn_off + cb * CONV_BLOCK_K + j * 8;
              const float4* g4 = reinterpret_cast<const float4*>(gtab + im * gn_pitch + c0 + 2 * (c0 >> 3));
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 v = g4[i];
                sc[im][2 * i] = v.x; sc[im][2 * i + 1] = v.y; sh[im][2 * i] = v.z; sh[im][2 * i + 1] = v.w;
              }
            }
          }
          ptx::mbar_wait(a_full(as), aphase);
          HDBG_ACC(0);
          if (HALO_DBG && tt == 0 && !first_halo_seen) { first_halo_seen = true; p.dbg[blockIdx.x * 16 + 7] = (unsigned long long)(clock64() - t_entry); }   // [7] first halo landed
          HDBG_T0();
          if (HEAD) {
            // one halo pixel per thread (180 of the 256): the pixel's fp32 inputs (cond channels, then x channels) -> a
            // 64-byte split-precision row hi | lo | hi; out-of-image pixels are the conv's zero padding. All loads of the
            // MT tiles are issued before any arithmetic.
            const int n = p.head_cc + p.head_cx;
            const size_t plane = (size_t)p.H * p.W;
            float v[MT][8];
            const int hy = tt / G::PITCH, hx = tt - hy * G::PITCH;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              const int gy = t[m].y0 - 1 + hy, gx = t[m].x0 - 1 + hx;
              const bool ok = tt < G::NPIX && (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W;
              const size_t off = (size_t)gy * p.W + gx;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                v[m][c] = 0.f;
                if (ok && c < n)
                  v[m][c] = c < p.head_cc ? __ldg(p.head_cond + ((size_t)t[m].b * p.head_cc + c) * plane + off)
                                          : __ldg(p.head_x + ((size_t)t[m].b * p.head_cx + (c - p.head_cc)) * plane + off);
              }
            }
            if (tt < G::NPIX) {
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                float row[32];
                if (n == 6) halo_head_row<6>(v[m], row);      // (compile-time channel counts: constant register indices)
                else if (n == 2) halo_head_row<2>(v[m], row);
                else if (n == 8) halo_head_row<8>(v[m], row);
                else if (n == 3) halo_head_row<3>(v[m], row);
                else if (n == 4) halo_head_row<4>(v[m], row);
                else halo_head_row<1>(v[m], row);
                uint8_t* dst = smem_gen + as * S::A_STAGE + m * G::STRIDE + tt * 128;
#pragma unroll
                for (int jj = 0; jj < 2 * KSTEPS; ++jj)
                  *reinterpret_cast<uint4*>(dst + ((jj ^ (tt & 7)) << 4)) = pack8(row + 8 * jj);
              }
            }
            fence_proxy_async_smem();
          }
          if (seg.gn_off >= 0 && !HALO_ABLATE(2)) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              // GEO 0: chunks 32 pixels apart cover the 180-pixel halo (the ring belongs to neighbouring
              // tiles or is zero padding). GEO 1: only the 2 x 64 image pixels are visited - the ring is
              // all padding. All loads are issued before any math: a warp must cover LDS + MUFU latency
              // with its own instruction-level parallelism.
              constexpr int NCHK = GEO == 0 ? (G::NPIX + 31) / 32 : (GEO == 1 ? 4 : 3);
              uint8_t* tile = smem_gen + as * S::A_STAGE + m * G::STRIDE + G::TMA_OFF;
              const int gx0 = t[m].x0 - 1, gy0 = t[m].y0 - 1;
              uint4 v[NCHK];
              uint4* ptr[NCHK];
              bool ok[NCHK];
              int img[NCHK];
#pragma unroll
              for (int i = 0; i < NCHK; ++i) {
                int px;
                if (GEO == 0) {
                  px = p_first + 32 * i;
                  const int hy = px / G::PITCH, hx = px - hy * G::PITCH;
                  img[i] = 0;
                  ok[i] = px < G::NPIX && (unsigned)(gy0 + hy) < (unsigned)p.H && (unsigned)(gx0 + hx) < (unsigned)p.W;
                } else if (GEO == 1) {
                  const int q = p_first + 32 * i;                       // (y, image, x) = (q >> 4, (q >> 3) & 1, q & 7)
                  img[i] = (q >> 3) & 1;
                  px = ((q >> 4) + 1) * G::PITCH + img[i] * 10 + (q & 7) + 1;
                  ok[i] = t[m].b + img[i] < p.B;
                } else {
                  const int q = p_first + 32 * i;                       // (image, y, x) = (q >> 4, (q >> 2) & 3, q & 3), q < 80
                  img[i] = min(q >> 4, IMGS - 1);
                  px = img[i] * 25 + (((q >> 2) & 3) + 1) * G::PITCH + (q & 3) + 1;
                  ok[i] = q < 80 && t[m].b + img[i] < p.B;
                }
                ptr[i] = reinterpret_cast<uint4*>(tile + px * 128 + ((j ^ (px & 7)) << 4));
                v[i] = ok[i] ? *ptr[i] : make_uint4(0u, 0u, 0u, 0u);
              }
#pragma unroll
              for (int i = 0; i < NCHK; ++i) {
                float f[8];
                unpack8(v[i], f);
                if (GEO == 2) {      // this item's image: its 8 channels' (scale, shift) from the staged table
                  const int c0 = seg.gn_off + cb * CONV_BLOCK_K + j * 8;
                  const float4* g4 = reinterpret_cast<const float4*>(gtab + img[i] * gn_pitch + c0 + 2 * (c0 >> 3));
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float4 tv = g4[k];
                    sc[0][2 * k] = tv.x; sc[0][2 * k + 1] = tv.y; sh[0][2 * k] = tv.z; sh[0][2 * k + 1] = tv.w;
                  }
                }
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                  float sc0 = sc[0][e], sh0 = sh[0][e], sc1 = sc[0][e + 1], sh1 = sh[0][e + 1];
                  if (GEO == 1 && img[i] == 1) { sc0 = sc[NTAB - 1][e]; sh0 = sh[NTAB - 1][e]; sc1 = sc[NTAB - 1][e + 1]; sh1 = sh[NTAB - 1][e + 1]; }
                  if (PRELU) {      // y = scale * prelu(x) + shift (arcface.py:60-64: bn1 before conv1, prelu before conv2)
                    const float x0 = f[e], x1 = f[e + 1];
                    fma2(f[e], f[e + 1], x0 > 0.f ? x0 : x0 * sl[e % (PRELU ? 8 : 1)], x1 > 0.f ? x1 : x1 * sl[(e + 1) % (PRELU ? 8 : 1)],
                         sc0, sc1, sh0, sh1);
                    continue;
                  }
                  float h0, h1;
                  fma2(h0, h1, f[e], f[e + 1], sc0, sc1, sh0, sh1);
                  // x*sigmoid(x) = h + h*tanh(h), h = x/2 (sc/sh arrive pre-halved): ONE MUFU per element
                  if (do_swish) fma2(f[e], f[e + 1], h0, h1, tanh_approx(h0), tanh_approx(h1), h0, h1);
                  else { f[e] = h0; f[e + 1] = h1; }
                }
                if (ok[i]) *ptr[i] = pack8(f);
              }
            }
            fence_proxy_async_smem();     // generic-proxy writes -> visible to the tensor core's async reads
          }
          if (CG == 2) {
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(a_ready(as) & HALO_PEER_MASK);      // the leader's barrier
          } else {
            ptx::mbar_arrive(a_ready(as));
          }
          HDBG_ACC(1);
          if (++as == AST) { as = 0; aphase ^= 1u; }
        }
      }
    }
#ifndef HALO_EPI_PROFILE
    if (tt == 0) HDBG_FLUSH(10, 2);      // [10] transform waits A full (incl. table loads), [11] transforming
#endif
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();          // no CTA of the pair exits while the other may still signal it
  if (warp == LW + 2) {
    if (CG == 2) tmem_dealloc_pair(tmem_base, NBUF * MT * BLOCK_N);
    else ptx::tmem_dealloc(tmem_base, NBUF * MT * BLOCK_N);
  }
}
#endif  // __CUDACC__

}  // namespace b200sr3
