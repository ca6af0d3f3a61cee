// MICA identity encoder: ArcFace iResNet-100 + F.normalize + MappingNetwork (arcface.cu).
#pragma once
#include "engine.cuh"

namespace b200sr3 {

class MicaEncoder {
 public:
  MicaEncoder(int device, int z_dim, int map_hidden_dim, int map_layers, int n_shape);
  ~MicaEncoder();

  int num_tensors() const { return (int)tensors_.size(); }
  const TensorSpec& tensor(int i) const { return tensors_.at(i); }
  // keys: "arcface.<reference Arcface state_dict key>" and "regressor.<reference MappingNetwork state_dict key>"
  void load_tensor(const std::string& key, const float* data, const int64_t* shape, int ndim);
  void finalize_weights(cudaStream_t s);
  // blob: fp32 [B,3,112,112] (host or device). Outputs (device or host, each optional): the raw 512-d embedding
  // (arcface.py:199), F.normalize of it (model/sr3d/model.py:167), the regressor's shape code [B,n_shape].
  void encode(const float* blob, int B, float* embedding, float* identity, float* shape_code, cudaStream_t s);
  int profile(int B, int max_ops, float* ms, double* flops, char* names, int names_len, cudaStream_t s);
  void layer_output(const std::string& layer, float* dst, int* C, int* H, int* W, cudaStream_t s);
  int n_shape() const { return n_shape_; }
  int64_t last_total = 0, last_conv = 0;

 private:
  struct Block {          // IBasicBlock (arcface.py:40-69)
    std::string name;     // "layer3.17"
    int inplanes = 0, planes = 0, stride = 1;
    bool down = false;    // downsample = conv1x1(stride) + BN on the identity path (arcface.py:133-137)
    PackedConv c1, c2;
    const float2* xf1 = nullptr;   // (scale, shift) of bn1, applied in front of conv1
    const float2* xf2 = nullptr;   // (1, 0): only the PReLU acts in front of conv2
    const float* slope2 = nullptr; // prelu.weight
  };
  struct Plan {           // activations + launch list for one batch size
    int B = 0;
    std::vector<void*> allocations;
    float *blob = nullptr, *emb = nullptr, *ident = nullptr, *shape = nullptr;
    std::vector<Op> ops;
    std::map<std::string, Act> layer_out;
    cudaGraphExec_t graph = nullptr;
    int64_t n_conv = 0;
    ~Plan();
  };
  void add_tensor(const std::string& key, std::vector<int64_t> shape);
  float* T_(const std::string& key) const;
  Plan& plan(int B);
  void build_plan(Plan& pl);

  int device_ = 0, z_dim_ = 512, map_hidden_ = 300, map_layers_ = 3, n_shape_ = 300;
  std::vector<TensorSpec> tensors_;
  std::map<std::string, int> tensor_index_;
  bool finalized_ = false;
  bool use_graph_ = true;
  std::vector<Block> blocks_;
  PackedConv stem_;
  float *fc_w_ = nullptr, *fc_b_ = nullptr, *ones_ = nullptr;
  std::vector<void*> owned_;
  std::vector<std::unique_ptr<Plan>> plans_;
  cudaStream_t capture_stream_ = nullptr;
};

}  // namespace b200sr3
