// Model handle of libhft_sm100.so: dimensions, the state_dict schema, registered + derived weights, workspace.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <string>
#include <vector>
#include "hft_internal.h"

namespace hft {

struct WeightSpec {
  std::string name;
  long long numel;
};

// Indices into Model::w (fp32 device copies) for one attention block / FFN / LayerNorm.
struct AttnW { int q_w, q_b, k_w, k_b, v_w, v_b, o_w, o_b; };
struct FfnW { int w1, b1, w2, b2; };
struct LnW { int g, b; };

struct EncLayerW { LnW ln; AttnW sa; FfnW ff; };
struct DecLayerW { LnW ln; AttnW sa; AttnW ca; FfnW ff; bool has_sa; };

// Derived fp32 tensors the kernels consume (built by hft_model_set_weights).
struct FusedAttn {
  float* qkv_w = nullptr;   // [3H,H]  rows: q | k | v
  float* qkv_b = nullptr;   // [3H]
};

struct Model {
  hft_dims d;
  int H, P, heads, dh, nbin, nframe, nnote, nvel, W, nproc;
  std::vector<WeightSpec> spec;
  std::vector<float*> w;              // device fp32 copies, same order as spec
  float* arena = nullptr;             // backing store of w
  bool weights_set = false;

  // schema indices
  int conv_w, conv_b, tok_w, tok_b, pos_freq;
  std::vector<EncLayerW> enc;
  int dec_pos_freq;
  DecLayerW dec0;
  std::vector<DecLayerW> dec;
  int head_freq[8];                   // onset w,b offset w,b mpe w,b velocity w,b
  int pos_time;
  std::vector<EncLayerW> tim;
  int head_time[8];

  // derived (fp32)
  float* front_w = nullptr;           // [H,65] collapsed conv+Linear
  float* front_b = nullptr;           // [H]
  std::vector<FusedAttn> enc_qkv, tim_qkv, dec_sa_qkv;      // self-attention fused QKV
  std::vector<FusedAttn> dec_ca_kv;                         // cross-attention fused [2H,H] (k | v), index 0 = layer zero
  float* q0 = nullptr;                // [n_note,H] = fc_q(pos_embedding_freq) of layer zero (input independent)
  float* headA_w = nullptr;           // [3+V,H]
  float* headA_b = nullptr;
  float* headB_w = nullptr;
  float* headB_b = nullptr;
  float* derived_arena = nullptr;

  // workspace
  int max_batch = 8;
  int ws_batch = 0;
  float* ws = nullptr;
  size_t ws_bytes = 0;

  // tensor-core path state (fwd_tc.cu)
  void* tc = nullptr;
};

int model_build_schema(Model* m);
int forward_f32(Model* m, const float* spec, long long sb, long long sbin, long long st, int B, const hft_outputs* o, cudaStream_t s, int stage = 0, float* enc_io = nullptr);
int forward_tc(Model* m, int precision, const float* spec, long long sb, long long sbin, long long st, int B, const hft_outputs* o, cudaStream_t s);
int tc_prepare_weights(Model* m, cudaStream_t s);
void tc_destroy(Model* m);
void tc_release_workspace(Model* m);

}  // namespace hft
