// fp32 CUDA-core forward of Model_SPEC2MIDI (reference hftt_code/model/model_spec2midi.py:15-35) -- the
// HFT_PREC_F32 path: IEEE fp32 FMA arithmetic end to end, used for the 2e-3 parity budget and as the on-device
// cross-check of the tensor-core path.  Kernels: collapsed front filter, tiled SGEMM (+bias, +ReLU), two-pass
// softmax attention, residual+LayerNorm, time re-layout, head finish.
#include "common.cuh"
#include "f32_kernels.cuh"
#include "model.h"

#include <math.h>

namespace hft {

// ------------------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------------------
static int gemm(cudaStream_t s, const float* A, int lda, const float* Wt, int ldw, const float* bias, float* C, int ldc, long long M, int N, int K,
                bool relu) {
  HFT_REQUIRE(K % GBK == 0 && lda % 4 == 0 && ldw % 4 == 0, HFT_ERR_UNSUPPORTED, "sgemm: K=%d lda=%d ldw=%d not supported", K, lda, ldw);
  const int bn = sgemm_tile_n(N);
  dim3 grid((N + bn - 1) / bn, (unsigned)((M + GBM - 1) / GBM));
  LaunchScope ls(HFT_KCLASS_GEMM, s);
  if (bn == 64) {
    if (relu) sgemm_tn_kernel<true, 64><<<grid, 256, 0, s>>>(A, lda, Wt, ldw, bias, C, ldc, (int)M, N, K);
    else sgemm_tn_kernel<false, 64><<<grid, 256, 0, s>>>(A, lda, Wt, ldw, bias, C, ldc, (int)M, N, K);
  } else {
    if (relu) sgemm_tn_kernel<true, 128><<<grid, 256, 0, s>>>(A, lda, Wt, ldw, bias, C, ldc, (int)M, N, K);
    else sgemm_tn_kernel<false, 128><<<grid, 256, 0, s>>>(A, lda, Wt, ldw, bias, C, ldc, (int)M, N, K);
  }
  return HFT_OK;
}

static int attention(Model* m, cudaStream_t s, const float* Q, int ldq, long long q_seq_stride, const float* K, const float* V, int ldkv, long long S,
                     int Lq, int Lk, float* ctx, float* probs) {
  const int dh = m->dh;
  size_t smem = (size_t)2 * Lk * dh * sizeof(float);
  int threads = (Lq + 31) / 32 * 32;
  HFT_REQUIRE(threads <= 256 && smem <= 200 * 1024, HFT_ERR_UNSUPPORTED, "attention: Lq=%d Lk=%d unsupported", Lq, Lk);
  dim3 grid((unsigned)S, m->heads);
  float inv_scale = 1.f / sqrtf((float)dh);
  LaunchScope ls(HFT_KCLASS_ATTENTION, s);
  if (dh == 64) {
    HFT_CHECK_CUDA(cudaFuncSetAttribute(attn_f32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attn_f32_kernel<64><<<grid, threads, smem, s>>>(Q, ldq, q_seq_stride, K, V, ldkv, Lq, Lk, m->heads, inv_scale, ctx, m->H, probs);
  } else {
    HFT_CHECK_CUDA(cudaFuncSetAttribute(attn_f32_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attn_f32_kernel<32><<<grid, threads, smem, s>>>(Q, ldq, q_seq_stride, K, V, ldkv, Lq, Lk, m->heads, inv_scale, ctx, m->H, probs);
  }
  return HFT_OK;
}

static void add_ln(Model* m, cudaStream_t s, const float* x, const float* r, long long r_rows, const LnW& ln, long long rows, float* y) {
  LaunchScope ls(HFT_KCLASS_NORM, s);
  add_ln_f32_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(x, r, r_rows, m->w[ln.g], m->w[ln.b], m->H, rows, y);
}

struct Ws32 {
  float *X, *BIG, *CTX, *TMP, *HID, *T, *DQ, *U;
};

static int ensure_ws(Model* m, int B, Ws32& w) {
  const long long Re = (long long)B * m->nframe * m->nbin, Rd = (long long)B * m->nframe * m->nnote;
  const int H = m->H, P = m->P;
  size_t need = ((size_t)Re * (6 * H + P) + (size_t)Rd * 5 * H) * sizeof(float);
  if (m->ws_bytes < need) {
    cudaFree(m->ws);
    m->ws = nullptr; m->ws_bytes = 0;
    HFT_CHECK_CUDA(cudaMalloc(&m->ws, need));
    m->ws_bytes = need;
  }
  float* p = m->ws;
  w.X = p; p += Re * H;
  w.BIG = p; p += Re * 3 * H;
  w.CTX = p; p += Re * H;
  w.TMP = p; p += Re * H;
  w.HID = p; p += Re * P;
  w.T = p; p += Rd * H;
  w.DQ = p; p += Rd * 3 * H;
  w.U = p;
  return HFT_OK;
}

#define HFT_TRY(x) do { int _rc = (x); if (_rc != HFT_OK) return _rc; } while (0)

// EncoderLayer (model_spec2midi.py:230-245) over S sequences of L tokens stored in x [S*L, H]
static int encoder_layer_f32(Model* m, cudaStream_t s, Ws32& w, float* x, long long S, int L, const EncLayerW& lw, const FusedAttn& qkv) {
  const int H = m->H, P = m->P;
  const long long R = S * L;
  HFT_TRY(gemm(s, x, H, qkv.qkv_w, H, qkv.qkv_b, w.BIG, 3 * H, R, 3 * H, H, false));
  HFT_TRY(attention(m, s, w.BIG, 3 * H, (long long)L * 3 * H, w.BIG + H, w.BIG + 2 * H, 3 * H, S, L, L, w.CTX, nullptr));
  HFT_TRY(gemm(s, w.CTX, H, m->w[lw.sa.o_w], H, m->w[lw.sa.o_b], w.TMP, H, R, H, H, false));
  add_ln(m, s, x, w.TMP, R, lw.ln, R, x);
  HFT_TRY(gemm(s, x, H, m->w[lw.ff.w1], H, m->w[lw.ff.b1], w.HID, P, R, P, H, true));
  HFT_TRY(gemm(s, w.HID, P, m->w[lw.ff.w2], P, m->w[lw.ff.b2], w.TMP, H, R, H, P, false));
  add_ln(m, s, x, w.TMP, R, lw.ln, R, x);
  return HFT_OK;
}

// stage 0: whole forward.  stage 1: encoder only (model_spec2midi.py:60-106), the memory [B, n_frame, n_bin, hid] is copied to enc_io.
// stage 2: decoder only (model_spec2midi.py:145-216) on the memory read from enc_io.
int forward_f32(Model* m, const float* spec, long long sb, long long sbin, long long st, int B, const hft_outputs* o, cudaStream_t s, int stage, float* enc_io) {
  if (B == 0) return HFT_OK;
  Ws32 w;
  HFT_TRY(ensure_ws(m, B, w));
  const int H = m->H, P = m->P, F = m->nframe, NB = m->nbin, NN = m->nnote, V = m->nvel;
  const long long Se = (long long)B * F, Re = Se * NB, Rd = Se * NN;
  const float sqrtH = sqrtf((float)H);
  HFT_REQUIRE(m->nproc == 65, HFT_ERR_UNSUPPORTED, "front kernel is built for n_margin 32");
  if (stage == 2) {
    HFT_CHECK_CUDA(cudaMemcpyAsync(w.X, enc_io, (size_t)Re * H * sizeof(float), cudaMemcpyDeviceToDevice, s));
  } else {
    // encoder front
    {
      LaunchScope ls(HFT_KCLASS_FRONT, s);
      front_f32_kernel<65><<<dim3(NB, B), 256, 0, s>>>(spec, sb, sbin, st, m->front_w, m->front_b, m->w[m->pos_freq], sqrtH, H, F, NB, w.X);
    }
    for (size_t l = 0; l < m->enc.size(); ++l) HFT_TRY(encoder_layer_f32(m, s, w, w.X, Se, NB, m->enc[l], m->enc_qkv[l]));
    if (stage == 1) {
      HFT_CHECK_CUDA(cudaMemcpyAsync(enc_io, w.X, (size_t)Re * H * sizeof(float), cudaMemcpyDeviceToDevice, s));
      HFT_CHECK_CUDA(cudaGetLastError());
      return HFT_OK;
    }
  }
  // decoder: layer zero (cross-attention of the constant pitch queries + FFN), model_spec2midi.py:255-272
  const int n_cross = 1 + (int)m->dec.size();
  {
    const DecLayerW& lw = m->dec0;
    HFT_TRY(gemm(s, w.X, H, m->dec_ca_kv[0].qkv_w, H, m->dec_ca_kv[0].qkv_b, w.BIG, 2 * H, Re, 2 * H, H, false));
    float* probs = (n_cross == 1) ? o->attention : nullptr;
    HFT_TRY(attention(m, s, m->q0, H, 0, w.BIG, w.BIG + H, 2 * H, Se, NN, NB, w.CTX, probs));
    HFT_TRY(gemm(s, w.CTX, H, m->w[lw.ca.o_w], H, m->w[lw.ca.o_b], w.TMP, H, Rd, H, H, false));
    add_ln(m, s, w.TMP, m->w[m->dec_pos_freq], NN, lw.ln, Rd, w.T);
    HFT_TRY(gemm(s, w.T, H, m->w[lw.ff.w1], H, m->w[lw.ff.b1], w.HID, P, Rd, P, H, true));
    HFT_TRY(gemm(s, w.HID, P, m->w[lw.ff.w2], P, m->w[lw.ff.b2], w.TMP, H, Rd, H, P, false));
    add_ln(m, s, w.T, w.TMP, Rd, lw.ln, Rd, w.T);
  }
  // decoder layers 1..L-1: self-attention -> cross-attention -> FFN, model_spec2midi.py:283-306
  for (size_t l = 0; l < m->dec.size(); ++l) {
    const DecLayerW& lw = m->dec[l];
    HFT_TRY(gemm(s, w.T, H, m->dec_sa_qkv[l].qkv_w, H, m->dec_sa_qkv[l].qkv_b, w.DQ, 3 * H, Rd, 3 * H, H, false));
    HFT_TRY(attention(m, s, w.DQ, 3 * H, (long long)NN * 3 * H, w.DQ + H, w.DQ + 2 * H, 3 * H, Se, NN, NN, w.CTX, nullptr));
    HFT_TRY(gemm(s, w.CTX, H, m->w[lw.sa.o_w], H, m->w[lw.sa.o_b], w.TMP, H, Rd, H, H, false));
    add_ln(m, s, w.T, w.TMP, Rd, lw.ln, Rd, w.T);
    HFT_TRY(gemm(s, w.T, H, m->w[lw.ca.q_w], H, m->w[lw.ca.q_b], w.DQ, H, Rd, H, H, false));
    HFT_TRY(gemm(s, w.X, H, m->dec_ca_kv[l + 1].qkv_w, H, m->dec_ca_kv[l + 1].qkv_b, w.BIG, 2 * H, Re, 2 * H, H, false));
    float* probs = ((int)l + 2 == n_cross) ? o->attention : nullptr;
    HFT_TRY(attention(m, s, w.DQ, H, (long long)NN * H, w.BIG, w.BIG + H, 2 * H, Se, NN, NB, w.CTX, probs));
    HFT_TRY(gemm(s, w.CTX, H, m->w[lw.ca.o_w], H, m->w[lw.ca.o_b], w.TMP, H, Rd, H, H, false));
    add_ln(m, s, w.T, w.TMP, Rd, lw.ln, Rd, w.T);
    HFT_TRY(gemm(s, w.T, H, m->w[lw.ff.w1], H, m->w[lw.ff.b1], w.HID, P, Rd, P, H, true));
    HFT_TRY(gemm(s, w.HID, P, m->w[lw.ff.w2], P, m->w[lw.ff.b2], w.TMP, H, Rd, H, P, false));
    add_ln(m, s, w.T, w.TMP, Rd, lw.ln, Rd, w.T);
  }
  // heads A (:172-175)
  const int NH = 3 + V;
  HFT_TRY(gemm(s, w.T, H, m->headA_w, H, m->headA_b, w.HID, NH, Rd, NH, H, false));
  {
    LaunchScope ls(HFT_KCLASS_HEADS, s);
    heads_finish_f32_kernel<<<(unsigned)((Rd * NH + 255) / 256), 256, 0, s>>>(w.HID, NH, V, F, NN, Rd, false, o->onset_A, o->offset_A, o->mpe_A, o->velocity_A);
    if (o->velocity_A_argmax) heads_argmax_f32_kernel<<<(unsigned)((Rd + 255) / 256), 256, 0, s>>>(w.HID, NH, V, F, NN, Rd, false, o->velocity_A_argmax);
  }
  // time re-layout + SAtime (:189-198)
  {
    LaunchScope ls(HFT_KCLASS_NORM, s);
    time_relayout_f32_kernel<<<(unsigned)((Rd * H + 255) / 256), 256, 0, s>>>(w.T, m->w[m->pos_time], sqrtH, F, NN, H, Rd * H, w.U);
  }
  for (size_t l = 0; l < m->tim.size(); ++l) HFT_TRY(encoder_layer_f32(m, s, w, w.U, (long long)B * NN, F, m->tim[l], m->tim_qkv[l]));
  // heads B (:203-206)
  HFT_TRY(gemm(s, w.U, H, m->headB_w, H, m->headB_b, w.HID, NH, Rd, NH, H, false));
  {
    LaunchScope ls(HFT_KCLASS_HEADS, s);
    heads_finish_f32_kernel<<<(unsigned)((Rd * NH + 255) / 256), 256, 0, s>>>(w.HID, NH, V, F, NN, Rd, true, o->onset_B, o->offset_B, o->mpe_B, o->velocity_B);
    if (o->velocity_B_argmax) heads_argmax_f32_kernel<<<(unsigned)((Rd + 255) / 256), 256, 0, s>>>(w.HID, NH, V, F, NN, Rd, true, o->velocity_B_argmax);
  }
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}

}  // namespace hft
