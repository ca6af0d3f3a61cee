// Host-side engine: layer plan of the reference UNet, packed weights, per-(B,R) workspace with
// a static buffer plan, TMA descriptors and one CUDA graph per sampling step.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b200sr3.h"
#include "common.cuh"
#include "conv_umma.cuh"
#include "conv_halo.cuh"
#include "kernels.cuh"

namespace b200sr3 {

enum class LayerKind { HeadConv, Res, Down, Up, Final };

struct LayerDesc {
  std::string name;   // reference module path: "downs.3", "mid.0", "ups.18", "final_conv"
  LayerKind kind;
  int c_x = 0;        // channels arriving on the main path
  int c_skip = 0;     // channels of the skip tensor concatenated after it (ups only)
  int cout = 0;
  bool attn = false;
};

struct TensorSpec {
  std::string key;
  std::vector<int64_t> shape;
  float* dev = nullptr;
  bool loaded = false;
  size_t numel() const { size_t n = 1; for (auto d : shape) n *= (size_t)d; return n; }
};

struct PackedConv {     // weights of one tensor-core conv
  bf16* w = nullptr;    // [Cout][k_total]
  int cout = 0, k_total = 0;
  int taps = 9;         // main segment: 9 or 1
  int cin_main = 0;     // channels of the main source
  int c_res0 = 0, c_res1 = 0;   // folded 1x1 res_conv segments (0 = none)
  bool res_identity = false;    // the 1x1 segment is an identity matrix (ResnetBlock with dim == dim_out, unet.py:101):
                                // executed by the GEMM, but NOT a conv of the reference graph - no algorithmic FLOPs
  bool up_folded = false;       // [4*Cout][4*Cin]: parity 2x2 convs of Upsample(nearest 2x)+conv3x3
  bool down_perm = false;       // taps ordered by input parity for the halo kernel's stride-2 path (conv_halo.cu):
                                // (0,0) (0,2) (2,0) (2,2) | (1,0) (1,2) | (0,1) (2,1) | (1,1)
  float* bias = nullptr;        // static bias [Cout] (null when the bias comes from the table)
};

// A launch that is part of the per-step kernel chain.
struct Op {
  std::string name;
  bool is_conv = false;
  std::function<void(cudaStream_t)> run;
  double flops = 0.0;   // algorithmic FLOPs (2*MAC on the reference graph) of one launch
  double bytes = 0.0;   // algorithmic HBM bytes of one launch (HBM-bound kernels)
  double flops_executed = 0.0;   // what the tensor pipe runs: + identity-shortcut segments and K padding, - upsample folding
  std::function<void()> report;  // timing build only: prints (and resets) the launch's role counters
  const void* weights = nullptr; // packed weights of this launch and their size: the PREVIOUS conv prefetches them into L2
  size_t weight_bytes = 0;
  std::function<void(const void*, size_t)> set_prefetch;   // (halo convs) which constants to prefetch while this launch runs
};

class Engine;

struct Workspace {
  int B = 0, R = 0;
  std::vector<void*> allocations;
  size_t bytes = 0;
  float* cond = nullptr;   // fp32 NCHW
  float* x = nullptr;      // fp32 NCHW sampler state
  float* eps = nullptr;    // fp32 NCHW (unet_forward output)
  std::vector<Op> ops;     // one UNet forward + update, in order
  // the last op (final conv + sampler update) writes either the updated state or eps; exactly one of
  // the two plans is live and the launch closure reads it at launch time
  std::shared_ptr<TailPlan> tail_plan;
  std::shared_ptr<ConvHaloParams> tail_halo;
  void set_tail(float* x, float* eps) {
    if (tail_plan) { tail_plan->x = x; tail_plan->eps_out = eps; }
    if (tail_halo) { tail_halo->tail_x = x; tail_halo->tail_eps = eps; }
  }
  std::map<std::string, Act> layer_out;
  cudaGraphExec_t graph = nullptr;
  int64_t n_conv = 0;
  int64_t n_kernels() const {      // launches of the step that are kernels ("stats.zero" is a memset node)
    int64_t n = 0;
    for (const Op& op : ops) n += op.name != "stats.zero";
    return n;
  }
  ~Workspace();
};

class Engine {
 public:
  Engine(const b200sr3_config& cfg, int device);
  ~Engine();

  int num_tensors() const { return (int)tensors_.size(); }
  const TensorSpec& tensor(int i) const { return tensors_.at(i); }
  void load_tensor(const std::string& key, const float* data, const int64_t* shape, int ndim);
  void finalize_weights(cudaStream_t s);
  void set_schedule(int T, const float* a, const float* bc, const float* c1, const float* c2,
                    const float* lv, const double* sqrt_ac_prev, cudaStream_t s);

  void unet_forward(const float* cond, const float* x, float noise_level, int B, int R, float* eps,
                    cudaStream_t s);
  void step(const float* cond, const float* x_t, const float* noise, int t, int clip_denoised, int B, int R,
            float* x_tm1, cudaStream_t s);
  // row_offset: global index of batch row 0 (Philox counter), so shards / chunks of one logical batch draw the rows'
  // own noise whatever the split
  void sample(const float* cond, int noise_mode, const float* noise, uint64_t seed, int64_t row_offset, int B, int R,
              float* out, float* snapshots, cudaStream_t s);
  void sample_host(const float* cond_host, uint64_t seed, int64_t row_offset, int B, int R, float* out_host,
                   cudaStream_t s);
  void philox_normal(uint64_t seed, int t, int64_t row_offset, int B, int R, float* out, cudaStream_t s);
  int num_snapshots() const;
  // One eager step with a CUDA event between consecutive launches: per-op device time.
  int profile_step(int B, int R, int max_ops, float* ms, double* flops, double* flops_executed, double* bytes,
                   char* names, int names_len, cudaStream_t s);
  void layer_output(const std::string& layer, float* dst, int* C, int* H, int* W, cudaStream_t s);

  int64_t last_total = 0, last_conv = 0;
  int device() const { return device_; }

 private:
  friend struct Workspace;
  void build_layers();
  void add_tensor(const std::string& key, std::vector<int64_t> shape);
  float* T_(const std::string& key) const;       // device pointer of a loaded tensor
  Workspace& workspace(int B, int R);
  void build_workspace(Workspace& ws);
  void write_ctl(int t, int mode, const float* noise, uint64_t seed, long long numel, long long row0, cudaStream_t s,
                 bool clip = true);
  void release();                                 // frees everything the engine owns (destructor, failed constructor)
  void run_ops(Workspace& ws, cudaStream_t s);
  void ensure_graph(Workspace& ws);

  b200sr3_config cfg_;
  int device_ = 0;
  std::vector<LayerDesc> layers_;
  std::vector<TensorSpec> tensors_;
  std::map<std::string, int> tensor_index_;
  bool finalized_ = false;

  std::map<std::string, PackedConv> convs_;       // "<layer>.c1", ".c2", ".conv", ".qkv", ".out"
  std::map<std::string, int> noise_off_;          // layer -> column offset in the bias table
  int noise_total_ = 0;
  float *head_w_ = nullptr, *tail_w_ = nullptr;
  PackedConv tail_pc_;        // final_conv as a 16-row bf16 GEMM operand (halo conv tail)
  PackedConv head_pc_;        // downs.0 as a split-precision bf16 GEMM operand (halo conv head)
  float *wall_ = nullptr, *ball_ = nullptr;
  float* ones_ = nullptr;     // 1024 ones: PReLU slopes of the affine-only (GroupNorm without Swish) transform
  std::vector<void*> owned_;

  int T_sched_ = 0;
  float* coefs_ = nullptr;   // [5][T]
  float* nl_ = nullptr;      // [T+1]
  float* table_ = nullptr;   // [T+1][noise_total]
  StepCtl* ctl_ = nullptr;
  cudaStream_t capture_stream_ = nullptr;
  bool use_graph_ = true;
  int force_block_n_ = 0;
  bool use_halo_ = true;       // B200SR3_NO_HALO=1: first-generation conv + separate GroupNorm apply everywhere
  bool fuse_stats_ = true;     // B200SR3_NO_FUSED_STATS=1: GroupNorm statistics by chan_stats_kernel instead

  // per-(B,R) workspaces, least recently used first; bounded (B200SR3_MAX_WORKSPACES, default 3): ragged last batches
  // and varying evaluation batch sizes must not accumulate ~2 GB plans until cudaMalloc fails
  std::vector<std::unique_ptr<Workspace>> workspaces_;
  size_t max_workspaces_ = 3;
};

// Builds the launch closure of one tensor-core convolution (conv_umma.cu).
struct ConvSource {
  Act act;            // source activation
  int taps = 9;       // 9: 3x3 pad 1; 1: 1x1
  int stride = 1;     // 1 or 2 (main source only)
  bool upsample2x = false;   // the conv reads nearest-2x(act) (unet.py:58-65), folded into 4 parity convs
};
struct ConvStats {    // where the epilogue leaves the GroupNorm statistics of the output
  long long* partial = nullptr;   // [B][slots][Cout][2], 2^-24 fixed point
  int slots = 0;
  bool atomic = false;            // slots == 1 and the launches ADD into it (the caller zeroes it once per step)
  unsigned long long* dbg = nullptr;   // role timing counters [148][8] (measurement only)
};
bool conv_can_fuse_stats(const Act& out, bool upsample2x);
int conv_stat_slots(const Act& out, bool upsample2x);
Op make_conv_op(const std::string& name, const ConvSource& main, const Act* res0, const Act* res1,
                const PackedConv& w, const float* bias, int bias_t_stride, const StepCtl* ctl,
                const bf16* residual, const Act& out, int force_block_n, const ConvStats* stats);
// Shared-memory opt-in for every conv kernel instance (once per process / device).
void conv_init_device();

// Halo-resident 3x3 conv with fused GroupNorm+Swish (conv_halo.cuh / conv_halo.cu).
struct HaloSource {
  Act act;
  int ntaps = 9;      // 9: 3x3 main source (folded to 4 parity taps when upsample2x); 1: 1x1 source - a shortcut behind the
                      // 3x3 sources, or, when the weights are a 1x1 conv (PackedConv::taps == 1), the conv's own input
  int gn_off = -1;    // channel offset in the GroupNorm (scale, shift) table; < 0: no transform
  bool identity = false;   // 1x1 source whose weights are an identity matrix (a residual add riding the GEMM): executed,
                           // but not a conv of the reference graph - no algorithmic FLOPs
};
bool conv_halo_eligible(int H, int W, int c_multiple_of_64_all, int cout);
int conv_halo_stat_slots(const Act& out, bool upsample2x);
struct HaloHead {      // downs.0 (unet.py:187) reading cat([cond, x], 1) (diffusion.py:170) as fp32 NCHW, no packed operand
  const float* cond = nullptr;   // [B][cc][R][R] or null
  const float* x = nullptr;      // [B][cx][R][R]
  int cc = 0, cx = 0;
};
struct HaloTail {      // final_conv fused with the sampler update (conv_halo.cuh, BLOCK_N == 16)
  float* x = nullptr;          // fp32 NCHW state, updated in place (null: eps only)
  float* eps_out = nullptr;    // optional fp32 NCHW eps
  const float* coefs = nullptr;
  int oc = 3;
};
// Everything about a halo conv that most call sites leave at its default.
struct HaloConvExtra {
  const HaloTail* tail = nullptr;                        // final_conv + sampler update (BLOCK_N == 16)
  std::shared_ptr<ConvHaloParams>* params_out = nullptr; // receives the launch parameters (the engine retargets the tail)
  const GnPlan* gn_from_stats = nullptr;                 // non-null: the kernel builds the (scale, shift) table itself (gn ignored)
  int stride = 1;                                        // 2: Downsample (unet.py:68-74), weights with down_perm
  const HaloHead* head = nullptr;                        // non-null: downs.0 straight from the fp32 NCHW inputs
  const float* prelu_slope = nullptr;                    // non-null: transform = (scale, shift) row or GroupNorm plan + PReLU slopes, no Swish
  bool partial_tiles = false;                            // any H, W (no statistics): ArcFace's 56 / 28 / 14 / 7 px
};
Op make_conv_halo_op(const std::string& name, const std::vector<HaloSource>& srcs, bool upsample2x,
                     const PackedConv& w, const float* bias, int bias_t_stride, const StepCtl* ctl, const Act& out,
                     const float2* gn, int gn_C, bool gn_swish, const ConvStats* stats,
                     const HaloConvExtra& extra = HaloConvExtra());
void conv_halo_init_device();
void halo_report_timing(const unsigned long long* dbg_dev, const char* label);
// OIHW fp32 [Cout][Cin][3][3] -> bf16 [Cout][9*Cin] with the taps in PackedConv::down_perm order (engine.cu)
// (k_total > 9*Cin: the row also holds shortcut columns behind the taps)
void pack_conv_weight_by_input_parity(const float* w, bf16* dst, int Cout, int Cin, cudaStream_t s, int k_total = 0);

}  // namespace b200sr3
