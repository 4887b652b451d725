// hft_model_* entry points: state_dict schema, weight registration and derived weights, forward dispatch.
// Mirrors the parameter layout of Model_SPEC2MIDI (reference hftt_code/model/model_spec2midi.py:9-378; the key
// names are the drop-in contract: m_training.py:275 load_state_dict, amt.py:24-25 pickled modules).
#include "common.cuh"
#include "model.h"
#include "hft_internal.h"

#include <math.h>

namespace hft {

static int add(Model* m, const std::string& name, long long numel) {
  m->spec.push_back(WeightSpec{name, numel});
  return (int)m->spec.size() - 1;
}

static LnW add_ln(Model* m, const std::string& p) {
  LnW l;
  l.g = add(m, p + ".layer_norm.weight", m->H);
  l.b = add(m, p + ".layer_norm.bias", m->H);
  return l;
}
static AttnW add_attn(Model* m, const std::string& p) {
  AttnW a;
  long long hh = (long long)m->H * m->H;
  a.q_w = add(m, p + ".fc_q.weight", hh); a.q_b = add(m, p + ".fc_q.bias", m->H);
  a.k_w = add(m, p + ".fc_k.weight", hh); a.k_b = add(m, p + ".fc_k.bias", m->H);
  a.v_w = add(m, p + ".fc_v.weight", hh); a.v_b = add(m, p + ".fc_v.bias", m->H);
  a.o_w = add(m, p + ".fc_o.weight", hh); a.o_b = add(m, p + ".fc_o.bias", m->H);
  return a;
}
static FfnW add_ffn(Model* m, const std::string& p) {
  FfnW f;
  long long hp = (long long)m->H * m->P;
  f.w1 = add(m, p + ".positionwise_feedforward.fc_1.weight", hp); f.b1 = add(m, p + ".positionwise_feedforward.fc_1.bias", m->P);
  f.w2 = add(m, p + ".positionwise_feedforward.fc_2.weight", hp); f.b2 = add(m, p + ".positionwise_feedforward.fc_2.bias", m->H);
  return f;
}
static EncLayerW add_enc_layer(Model* m, const std::string& p) {
  EncLayerW e;
  e.ln = add_ln(m, p);
  e.sa = add_attn(m, p + ".self_attention");
  e.ff = add_ffn(m, p);
  return e;
}
static void add_heads(Model* m, const std::string& p, const char* suffix, int* idx) {
  const char* names[3] = {"onset", "offset", "mpe"};
  for (int i = 0; i < 3; ++i) {
    idx[2 * i] = add(m, p + ".fc_" + names[i] + "_" + suffix + ".weight", m->H);
    idx[2 * i + 1] = add(m, p + ".fc_" + names[i] + "_" + suffix + ".bias", 1);
  }
  idx[6] = add(m, p + ".fc_velocity_" + suffix + ".weight", (long long)m->nvel * m->H);
  idx[7] = add(m, p + ".fc_velocity_" + suffix + ".bias", m->nvel);
}

int model_build_schema(Model* m) {
  const hft_dims& d = m->d;
  m->H = d.hid_dim; m->P = d.pf_dim; m->heads = d.n_heads; m->dh = d.hid_dim / d.n_heads;
  m->nbin = d.n_bin; m->nframe = d.n_frame; m->nnote = d.n_note; m->nvel = d.n_velocity;
  m->W = d.n_frame + 2 * d.n_margin; m->nproc = 2 * d.n_margin + 1;
  const std::string e = "encoder_spec2midi", dd = "decoder_spec2midi";
  int cnn_dim = d.cnn_channel * (m->nproc - (d.cnn_kernel - 1));
  m->conv_w = add(m, e + ".conv.weight", (long long)d.cnn_channel * d.cnn_kernel);
  m->conv_b = add(m, e + ".conv.bias", d.cnn_channel);
  m->tok_w = add(m, e + ".tok_embedding_freq.weight", (long long)m->H * cnn_dim);
  m->tok_b = add(m, e + ".tok_embedding_freq.bias", m->H);
  m->pos_freq = add(m, e + ".pos_embedding_freq.weight", (long long)m->nbin * m->H);
  for (int i = 0; i < d.n_enc_layers; ++i) m->enc.push_back(add_enc_layer(m, e + ".layers_freq." + std::to_string(i)));
  m->dec_pos_freq = add(m, dd + ".pos_embedding_freq.weight", (long long)m->nnote * m->H);
  {
    const std::string p = dd + ".layer_zero_freq";
    m->dec0.has_sa = false;
    m->dec0.ln = add_ln(m, p);
    m->dec0.ca = add_attn(m, p + ".encoder_attention");
    m->dec0.ff = add_ffn(m, p);
  }
  for (int i = 0; i < d.n_dec_layers - 1; ++i) {
    const std::string p = dd + ".layers_freq." + std::to_string(i);
    DecLayerW l;
    l.has_sa = true;
    l.ln = add_ln(m, p);
    l.sa = add_attn(m, p + ".self_attention");
    l.ca = add_attn(m, p + ".encoder_attention");
    l.ff = add_ffn(m, p);
    m->dec.push_back(l);
  }
  add_heads(m, dd, "freq", m->head_freq);
  m->pos_time = add(m, dd + ".pos_embedding_time.weight", (long long)m->nframe * m->H);
  for (int i = 0; i < d.n_dec_layers; ++i) m->tim.push_back(add_enc_layer(m, dd + ".layers_time." + std::to_string(i)));
  add_heads(m, dd, "time", m->head_time);
  return HFT_OK;
}

// ---- derived-weight kernels --------------------------------------------------------------------------------
// Collapse Conv2d(1,C,(1,kw)) + Linear(C*n_out -> H) (no non-linearity in between, model_spec2midi.py:73-85) into one
// n_proc-tap filter per hidden unit:  Wc[h][j] = sum_c sum_i W[h][c*n_out + j-i] cw[c][i],  bc[h] = b[h] + sum W cb.
__global__ void collapse_front_kernel(const float* __restrict__ tok_w, const float* __restrict__ tok_b, const float* __restrict__ conv_w,
                                      const float* __restrict__ conv_b, int H, int C, int kw, int n_out, int n_proc,
                                      float* __restrict__ Wc, float* __restrict__ bc) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * (n_proc + 1)) return;
  int h = idx / (n_proc + 1), j = idx % (n_proc + 1);
  const float* wrow = tok_w + (long long)h * C * n_out;
  if (j == n_proc) {
    double acc = tok_b[h];
    for (int c = 0; c < C; ++c)
      for (int k = 0; k < n_out; ++k) acc += (double)wrow[c * n_out + k] * (double)conv_b[c];
    bc[h] = (float)acc;
  } else {
    double acc = 0.0;
    for (int c = 0; c < C; ++c)
      for (int i = 0; i < kw; ++i) {
        int k = j - i;
        if (k >= 0 && k < n_out) acc += (double)wrow[c * n_out + k] * (double)conv_w[c * kw + i];
      }
    Wc[h * n_proc + j] = (float)acc;
  }
}

// Up to kCopyCap device-to-device slice copies in ONE launch (the fused Q|K|V / K|V / head matrices are re-assembled from the parameter arena
// after every optimiser step: 54 copies for the reduced model, which as separate launches cost 4 us each).  blockIdx.y = slice.
constexpr int kCopyCap = 96;
struct CopyBatch {
  const float* src[kCopyCap];
  float* dst[kCopyCap];
  int n[kCopyCap];
  int count;
};
__global__ void copy_batch_kernel(const __grid_constant__ CopyBatch b) {
  const int e = blockIdx.y, n = b.n[e];
  const float* __restrict__ src = b.src[e];
  float* __restrict__ dst = b.dst[e];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// q0[n][o] = sum_h pos[n][h] W[o][h] + b[o]   (fc_q of the constant pitch queries, model_spec2midi.py:154-155,260)
__global__ void q0_kernel(const float* __restrict__ pos, const float* __restrict__ w, const float* __restrict__ b, int n_note, int H,
                          float* __restrict__ q0) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_note * H) return;
  int n = idx / H, o = idx % H;
  float acc = 0.f;
  for (int h = 0; h < H; ++h) acc = fmaf(pos[n * H + h], w[o * H + h], acc);
  q0[idx] = acc + b[o];
}

static void copy_flush(CopyBatch& cb, cudaStream_t s) {
  if (cb.count == 0) return;
  int mx = 1;
  for (int i = 0; i < cb.count; ++i) mx = cb.n[i] > mx ? cb.n[i] : mx;
  int gx = (mx + 1023) / 1024;
  if (gx > 64) gx = 64;
  copy_batch_kernel<<<dim3(gx, cb.count), 256, 0, s>>>(cb);
  count_launch();
  cb.count = 0;
}

static int derive_weights(Model* m, cudaStream_t s) {
  const int H = m->H, V = m->nvel;
  const long long hh = (long long)H * H;
  size_t n_self = m->enc.size() + m->tim.size() + m->dec.size();
  size_t n_cross = 1 + m->dec.size();
  size_t floats = (size_t)H * m->nproc + H + n_self * (3 * hh + 3 * H) + n_cross * (2 * hh + 2 * H) + (size_t)m->nnote * H +
                  2 * ((size_t)(3 + V) * H + (3 + V)) + 64 * (8 + 2 * (n_self + n_cross));   // + per-slice alignment slack
  if (!m->derived_arena) HFT_CHECK_CUDA(cudaMalloc(&m->derived_arena, floats * sizeof(float)));
  float* p = m->derived_arena;
  CopyBatch cb{};
  auto dcopy = [&](const float* src, float* dst, long long n, cudaStream_t st) {
    if (cb.count == kCopyCap) copy_flush(cb, st);
    cb.src[cb.count] = src; cb.dst[cb.count] = dst; cb.n[cb.count] = (int)n; ++cb.count;
  };
  auto take = [&](size_t n) { float* r = p; p += (n + 63) & ~(size_t)63; return r; };   // 256-byte aligned slices
  m->front_w = take((size_t)H * m->nproc);
  m->front_b = take(H);
  int n_out = m->nproc - (m->d.cnn_kernel - 1);
  collapse_front_kernel<<<(H * (m->nproc + 1) + 127) / 128, 128, 0, s>>>(m->w[m->tok_w], m->w[m->tok_b], m->w[m->conv_w], m->w[m->conv_b], H,
                                                                           m->d.cnn_channel, m->d.cnn_kernel, n_out, m->nproc, m->front_w, m->front_b);
  auto fuse_self = [&](const AttnW& a, FusedAttn& f) {
    f.qkv_w = take(3 * hh);
    f.qkv_b = take(3 * H);
    dcopy(m->w[a.q_w], f.qkv_w, hh, s); dcopy(m->w[a.k_w], f.qkv_w + hh, hh, s); dcopy(m->w[a.v_w], f.qkv_w + 2 * hh, hh, s);
    dcopy(m->w[a.q_b], f.qkv_b, H, s); dcopy(m->w[a.k_b], f.qkv_b + H, H, s); dcopy(m->w[a.v_b], f.qkv_b + 2 * H, H, s);
  };
  auto fuse_cross = [&](const AttnW& a, FusedAttn& f) {
    f.qkv_w = take(2 * hh);
    f.qkv_b = take(2 * H);
    dcopy(m->w[a.k_w], f.qkv_w, hh, s); dcopy(m->w[a.v_w], f.qkv_w + hh, hh, s);
    dcopy(m->w[a.k_b], f.qkv_b, H, s); dcopy(m->w[a.v_b], f.qkv_b + H, H, s);
  };
  m->enc_qkv.resize(m->enc.size()); m->tim_qkv.resize(m->tim.size()); m->dec_sa_qkv.resize(m->dec.size());
  m->dec_ca_kv.resize(n_cross);
  for (size_t i = 0; i < m->enc.size(); ++i) fuse_self(m->enc[i].sa, m->enc_qkv[i]);
  for (size_t i = 0; i < m->tim.size(); ++i) fuse_self(m->tim[i].sa, m->tim_qkv[i]);
  for (size_t i = 0; i < m->dec.size(); ++i) fuse_self(m->dec[i].sa, m->dec_sa_qkv[i]);
  fuse_cross(m->dec0.ca, m->dec_ca_kv[0]);
  for (size_t i = 0; i < m->dec.size(); ++i) fuse_cross(m->dec[i].ca, m->dec_ca_kv[i + 1]);
  m->q0 = take((size_t)m->nnote * H);
  q0_kernel<<<(m->nnote * H + 127) / 128, 128, 0, s>>>(m->w[m->dec_pos_freq], m->w[m->dec0.ca.q_w], m->w[m->dec0.ca.q_b], m->nnote, H, m->q0);
  auto fuse_heads = [&](const int* idx, float*& hw, float*& hb) {
    hw = take((size_t)(3 + V) * H);
    hb = take(3 + V);
    for (int i = 0; i < 3; ++i) { dcopy(m->w[idx[2 * i]], hw + (size_t)i * H, H, s); dcopy(m->w[idx[2 * i + 1]], hb + i, 1, s); }
    dcopy(m->w[idx[6]], hw + (size_t)3 * H, (long long)V * H, s);
    dcopy(m->w[idx[7]], hb + 3, V, s);
  };
  fuse_heads(m->head_freq, m->headA_w, m->headA_b);
  fuse_heads(m->head_time, m->headB_w, m->headB_b);
  copy_flush(cb, s);
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}

int derive_weights_public(Model* m, cudaStream_t s) { return derive_weights(m, s); }   // after in-place parameter updates (train_f32.cu)

}  // namespace hft

using namespace hft;

extern "C" int hft_model_create(hft_model** out, const hft_dims* dims) {
  HFT_REQUIRE(out && dims, HFT_ERR_ARG, "hft_model_create: NULL argument");
  const hft_dims& d = *dims;
  HFT_REQUIRE(d.n_heads > 0 && d.hid_dim % d.n_heads == 0, HFT_ERR_ARG, "hft_model_create: hid_dim %d not divisible by n_heads %d", d.hid_dim, d.n_heads);
  int dh = d.hid_dim / d.n_heads;
  HFT_REQUIRE(dh == 32 || dh == 64, HFT_ERR_UNSUPPORTED, "hft_model_create: head_dim %d unsupported (32 or 64)", dh);
  HFT_REQUIRE(d.hid_dim % 64 == 0 && d.hid_dim <= 256 && d.pf_dim % 64 == 0 && d.pf_dim <= 512, HFT_ERR_UNSUPPORTED,
              "hft_model_create: hid_dim %d / pf_dim %d unsupported (multiples of 64, <= 256 / 512)", d.hid_dim, d.pf_dim);
  HFT_REQUIRE(d.n_bin == 256 && d.n_frame == 128 && d.n_margin == 32 && d.n_note == 88 && d.n_velocity == 128, HFT_ERR_UNSUPPORTED,
              "hft_model_create: only the reference geometry is supported (n_bin 256, n_frame 128, margin 32, 88 notes, 128 velocities)");
  HFT_REQUIRE(d.n_enc_layers >= 1 && d.n_dec_layers >= 1 && d.cnn_kernel >= 1 && d.cnn_kernel <= 2 * d.n_margin + 1 && d.cnn_channel >= 1,
              HFT_ERR_ARG, "hft_model_create: bad layer / cnn configuration");
  Model* m = new Model();
  m->d = d;
  model_build_schema(m);
  *out = reinterpret_cast<hft_model*>(m);
  return HFT_OK;
}

extern "C" int hft_model_destroy(hft_model* model) {
  if (!model) return HFT_OK;
  Model* m = reinterpret_cast<Model*>(model);
  tc_destroy(m);
  cudaFree(m->arena); cudaFree(m->derived_arena); cudaFree(m->ws);
  delete m;
  return HFT_OK;
}

extern "C" int hft_model_num_weights(const hft_model* model) { return model ? (int)reinterpret_cast<const Model*>(model)->spec.size() : 0; }
extern "C" const char* hft_model_weight_name(const hft_model* model, int i) {
  const Model* m = reinterpret_cast<const Model*>(model);
  return (m && i >= 0 && i < (int)m->spec.size()) ? m->spec[i].name.c_str() : nullptr;
}
extern "C" int64_t hft_model_weight_numel(const hft_model* model, int i) {
  const Model* m = reinterpret_cast<const Model*>(model);
  return (m && i >= 0 && i < (int)m->spec.size()) ? m->spec[i].numel : -1;
}

extern "C" int hft_model_set_weights(hft_model* model, const float* const* weights_dev, int n_weights, void* stream) {
  HFT_REQUIRE(model && weights_dev, HFT_ERR_ARG, "hft_model_set_weights: NULL argument");
  Model* m = reinterpret_cast<Model*>(model);
  HFT_REQUIRE(n_weights == (int)m->spec.size(), HFT_ERR_ARG, "hft_model_set_weights: got %d tensors, the schema has %d", n_weights, (int)m->spec.size());
  cudaStream_t s = (cudaStream_t)stream;
  if (!m->arena) {
    size_t total = 0;
    for (auto& w : m->spec) total += (size_t)((w.numel + 3) & ~3ll);
    HFT_CHECK_CUDA(cudaMalloc(&m->arena, total * sizeof(float)));
    m->w.resize(m->spec.size());
    float* p = m->arena;
    for (size_t i = 0; i < m->spec.size(); ++i) { m->w[i] = p; p += (m->spec[i].numel + 3) & ~3ll; }
  }
  for (size_t i = 0; i < m->spec.size(); ++i) {
    HFT_REQUIRE(weights_dev[i] != nullptr, HFT_ERR_ARG, "hft_model_set_weights: tensor %s is NULL", m->spec[i].name.c_str());
    HFT_CHECK_CUDA(cudaMemcpyAsync(m->w[i], weights_dev[i], m->spec[i].numel * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  int rc = derive_weights(m, s);
  if (rc != HFT_OK) return rc;
  rc = tc_prepare_weights(m, s);
  if (rc != HFT_OK) return rc;
  m->weights_set = true;
  return HFT_OK;
}

// The two halves of the forward as separate calls (the reference exposes them as modules: Encoder_SPEC2MIDI.forward model_spec2midi.py:60-106,
// Decoder_SPEC2MIDI.forward :145-216).  fp32 CUDA-core kernels; the fused hft_forward is the hot path.
extern "C" int hft_forward_encoder(hft_model* model, const float* spec_dev, int64_t stride_b, int64_t stride_bin, int64_t stride_t, int32_t batch,
                                   float* enc_out_dev, void* stream) {
  HFT_REQUIRE(model && enc_out_dev && (spec_dev || batch == 0) && batch >= 0, HFT_ERR_ARG, "hft_forward_encoder: bad argument");
  Model* m = reinterpret_cast<Model*>(model);
  HFT_REQUIRE(m->weights_set, HFT_ERR_STATE, "hft_forward_encoder: call hft_model_set_weights first");
  reset_launch_count();
  const long long per = (long long)m->nframe * m->nbin * m->H;
  for (int b0 = 0; b0 < batch; b0 += m->max_batch) {
    const int bc = batch - b0 < m->max_batch ? batch - b0 : m->max_batch;
    { int rc = forward_f32(m, spec_dev + (long long)b0 * stride_b, stride_b, stride_bin, stride_t, bc, nullptr, (cudaStream_t)stream, 1, enc_out_dev + b0 * per); if (rc != HFT_OK) return rc; }
  }
  return HFT_OK;
}

extern "C" int hft_forward_decoder(hft_model* model, const float* enc_dev, int32_t batch, const hft_outputs* outputs, void* stream) {
  HFT_REQUIRE(model && outputs && (enc_dev || batch == 0) && batch >= 0, HFT_ERR_ARG, "hft_forward_decoder: bad argument");
  Model* m = reinterpret_cast<Model*>(model);
  HFT_REQUIRE(m->weights_set, HFT_ERR_STATE, "hft_forward_decoder: call hft_model_set_weights first");
  reset_launch_count();
  const long long fn = (long long)m->nframe * m->nnote, per = (long long)m->nframe * m->nbin * m->H;
  for (int b0 = 0; b0 < batch; b0 += m->max_batch) {
    const int bc = batch - b0 < m->max_batch ? batch - b0 : m->max_batch;
    hft_outputs o = *outputs;
    auto adv = [&](float*& p, long long n) { if (p) p += (long long)b0 * n; };
    adv(o.onset_A, fn); adv(o.offset_A, fn); adv(o.mpe_A, fn); adv(o.velocity_A, fn * m->nvel);
    adv(o.attention, (long long)m->nframe * m->heads * m->nnote * m->nbin);
    adv(o.onset_B, fn); adv(o.offset_B, fn); adv(o.mpe_B, fn); adv(o.velocity_B, fn * m->nvel);
    if (o.velocity_A_argmax) o.velocity_A_argmax += (long long)b0 * fn;
    if (o.velocity_B_argmax) o.velocity_B_argmax += (long long)b0 * fn;
    { int rc = forward_f32(m, nullptr, 0, 0, 0, bc, &o, (cudaStream_t)stream, 2, const_cast<float*>(enc_dev) + b0 * per); if (rc != HFT_OK) return rc; }
  }
  return HFT_OK;
}

extern "C" int hft_model_set_max_batch(hft_model* model, int32_t max_batch) {
  HFT_REQUIRE(model && max_batch >= 1, HFT_ERR_ARG, "hft_model_set_max_batch: bad argument");
  Model* m = reinterpret_cast<Model*>(model);
  if (max_batch < m->max_batch) hft_model_release_workspace(model);      // a smaller bound gives the memory back (work spaces only grow otherwise)
  m->max_batch = max_batch;
  return HFT_OK;
}

// Frees the activation work spaces of all precision modes (48 segments per call in fp16x3: 14 GB); weights and derived tensors stay.
// The device is synchronised first: a forward still in flight may be using them.
extern "C" int hft_model_release_workspace(hft_model* model) {
  HFT_REQUIRE(model, HFT_ERR_ARG, "hft_model_release_workspace: NULL model");
  Model* m = reinterpret_cast<Model*>(model);
  HFT_CHECK_CUDA(cudaDeviceSynchronize());
  tc_release_workspace(m);
  cudaFree(m->ws);
  m->ws = nullptr; m->ws_bytes = 0;
  return HFT_OK;
}

extern "C" int hft_forward(hft_model* model, int precision, const float* spec_dev, int64_t stride_b, int64_t stride_bin, int64_t stride_t,
                           int32_t batch, const hft_outputs* outputs, void* stream) {
  HFT_REQUIRE(model && outputs, HFT_ERR_ARG, "hft_forward: NULL argument");
  Model* m = reinterpret_cast<Model*>(model);
  HFT_REQUIRE(m->weights_set, HFT_ERR_STATE, "hft_forward: call hft_model_set_weights first");
  HFT_REQUIRE(batch >= 0 && (spec_dev || batch == 0), HFT_ERR_ARG, "hft_forward: bad batch / spec");
  HFT_REQUIRE(precision == HFT_PREC_F32 || precision == HFT_PREC_BF16 || precision == HFT_PREC_F16 || precision == HFT_PREC_F16X3 || precision == HFT_PREC_MIXED, HFT_ERR_ARG, "hft_forward: unknown precision %d", precision);
  reset_launch_count();
  cudaStream_t s = (cudaStream_t)stream;
  const long long fn = (long long)m->nframe * m->nnote;
  for (int b0 = 0; b0 < batch; b0 += m->max_batch) {
    int bc = batch - b0 < m->max_batch ? batch - b0 : m->max_batch;
    hft_outputs o = *outputs;
    auto adv = [&](float*& p, long long per) { if (p) p += (long long)b0 * per; };
    adv(o.onset_A, fn); adv(o.offset_A, fn); adv(o.mpe_A, fn); adv(o.velocity_A, fn * m->nvel);
    adv(o.attention, (long long)m->nframe * m->heads * m->nnote * m->nbin);
    adv(o.onset_B, fn); adv(o.offset_B, fn); adv(o.mpe_B, fn); adv(o.velocity_B, fn * m->nvel);
    if (o.velocity_A_argmax) o.velocity_A_argmax += (long long)b0 * fn;
    if (o.velocity_B_argmax) o.velocity_B_argmax += (long long)b0 * fn;
    const float* sp = spec_dev + (long long)b0 * stride_b;
    int rc = (precision == HFT_PREC_F32) ? forward_f32(m, sp, stride_b, stride_bin, stride_t, bc, &o, s)
                                         : forward_tc(m, precision, sp, stride_b, stride_bin, stride_t, bc, &o, s);
    if (rc != HFT_OK) return rc;
  }
  return HFT_OK;
}
