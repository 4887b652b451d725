// Training step of Model_SPEC2MIDI on sm_100a, fp32 CUDA-core kernels (reference hftt_code/training/train.py:89-160:
// forward in train mode, the 8-term loss, loss.backward(), optimizer.step(); optimiser and initialisation
// hftt_code/training/m_training.py:31-33,141,146).  The forward keeps a tape (layer inputs, fused Q|K|V, attention row
// log-sum-exps, contexts, pre-LayerNorm sums, FFN hidden activations); the backward recomputes the attention
// probabilities from Q, K and the log-sum-exp instead of storing [S, heads, Lq, Lk].
// Gradients land in one flat fp32 vector laid out like the model's parameter arena (state_dict order, every tensor
// padded to 4 floats) so that the data-parallel configuration all-reduces ONE bucket over NCCL.
// Dropout (the reference trains with p = 0.1): counter-based masks, hash(seed, site, element) -- see Drop in f32_kernels.cuh;
// the backward regenerates them.  Sites are numbered in forward order (hft_sm100.h documents the numbering) so that a
// restatement can reproduce the masks; p = 0 (the parity configuration of SURVEY.md 8d config 5) skips every mask.
#include "common.cuh"
#include "model.h"
#include "train_kernels.cuh"
#include "tc_train_attn.cuh"
#include "tc_train_gemm.cuh"
#include "tc_train_dw.cuh"

#include <cudaTypedefs.h>
#include <math.h>
#include <vector>

namespace hft {

#define HFT_TRY(x) do { int _rc = (x); if (_rc != HFT_OK) return _rc; } while (0)

struct LayerTape {           // EncoderLayer / the blocks of a DecoderLayer
  float *xin, *qkv, *lse, *ctx, *s1, *x1, *hid, *s2;
};
struct DecTape {
  // self-attention block (layers >= 1)
  float *tin, *qkv, *lse_s, *ctx_s, *s0, *t0;
  // cross-attention block
  float *qc, *kv, *lse_c, *ctx_c, *s1, *t1;
  // FFN block
  float *hid, *s2;
};

struct Trainer {
  Model* m = nullptr;
  int B = 0;                     // capacity: the tape is sized for B segments
  int Bc = 0;                    // segments of the current step (<= B: the last batch of an epoch may be partial, train.py:72 with drop_last=False)
  float* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<LayerTape> enc, tim;
  std::vector<DecTape> dec;      // index 0 = layer zero
  float *x_enc = nullptr, *t_out = nullptr, *u0 = nullptr, *u_out = nullptr;
  float *logits_a = nullptr, *logits_b = nullptr;
  float *head_w[2] = {nullptr, nullptr}, *head_b[2] = {nullptr, nullptr}, *g_head_w = nullptr, *g_head_b = nullptr;
  float *g_front_w = nullptr, *g_front_b = nullptr;
  // gradient work buffers
  float *gX = nullptr, *gBIG = nullptr, *gCTX = nullptr, *gHID = nullptr, *gT = nullptr, *gDQ = nullptr, *gU = nullptr, *gLOG = nullptr, *dD = nullptr, *gQ0 = nullptr;
  float* gM = nullptr;           // masked copy of a gradient (sub-layer branch under dropout)
  int NP = 144;                  // padded head width (3 + V = 131 -> multiple of 16)
  float p_drop = 0.f;
  uint32_t seed = 0;
  Drop drop(int site) const {
    Drop d{0u, seed, (uint32_t)site, 1.f};
    if (p_drop > 0.f) { d.thresh = (uint32_t)((double)p_drop * 4294967296.0); d.scale = 1.f / (1.f - p_drop); }
    return d;
  }
  // site numbering (forward order)
  int site_enc(int l) const { return 1 + 4 * l; }
  int site_dec0() const { return 1 + 4 * (int)m->enc.size(); }
  int site_dec(int l) const { return site_dec0() + 4 + 6 * l; }           // l = 0 is the first DecoderLayer after layer zero
  int site_time_emb() const { return site_dec0() + 4 + 6 * (int)m->dec.size(); }
  int site_time(int l) const { return site_time_emb() + 1 + 4 * l; }
};

// ---- launch helpers -------------------------------------------------------------------------------------------------
// tcgen05 kernels of the training step (tc_train_attn.cuh: attention, head_dim 32; tc_train_gemm.cuh: Linear forward / input gradient;
// split-fp16 products, fp32-class results).  HFT_TRAIN_TC=0 selects the fp32 CUDA-core kernels instead (experiment switch;
// hft_train_attention() / hft_train_linear() can force either).
static int g_train_tc_force = -1;
static bool train_tc_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HFT_TRAIN_TC"); v = (e && e[0] == '0') ? 0 : 1; }
  return g_train_tc_force >= 0 ? g_train_tc_force == 1 : v == 1;
}
static bool tgemm_ok(const float* A, int lda, const float* W, int ldw, bool w_kn, const float* C, int ldc, int N, int K, const float* mask, int ldm) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return train_tc_enabled() && (K == 64 || K == 128) && N % 16 == 0 && N >= 16 && N * (K / 8) <= 2048 && lda % 4 == 0 && ldc % 4 == 0 && (w_kn || ldw % 4 == 0) && al(A) && al(W) &&
         al(C) && (!mask || (ldm % 4 == 0 && al(mask))) && tc::tgemm_smem(K / 64, N) <= 113 * 1024;
}
// ring depths of the tcgen05 weight-gradient kernel (one CTA per SM, 220 KB of shared memory): operand slots first (HFT_DW_NS, default 2),
// the raw fp32 ring takes what is left, at most 8 stages (HFT_DW_NR caps it)
static int tdw_op_slots() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HFT_DW_NS"); v = e ? atoi(e) : 2; if (v < 2 || v > 4) v = 2; }
  return v;
}
static int tdw_raw_stages(int N, int K, int n_slots) {
  const long long room = 220 * 1024 - (long long)tc::tdw_smem(K / 64, N, n_slots, 0);
  long long n = room / (long long)tc::tdw_raw_bytes(K / 64, N);
  if (n > 8) n = 8;
  if (const char* e = getenv("HFT_DW_NR")) { int v = atoi(e); if (v >= 1 && v < n) n = v; }
  return (int)n;
}
static void tdw_rings(int N, int K, int* n_slots, int* n_raw) {
  *n_slots = tdw_op_slots();
  while (*n_slots > 2 && tdw_raw_stages(N, K, *n_slots) < 3) --*n_slots;
  *n_raw = tdw_raw_stages(N, K, *n_slots);
}
// HFT_TRAIN_TC_DW=0: keep the fp32 CUDA-core weight-gradient kernel (experiment switch); hft_train_linear_wgrad(use_tc) forces either.
// 2-D row-major fp32 tensor [rows, cols] with row pitch ld (floats), box = box_cols x box_rows, no swizzle (the raw operand stages of tdw_kernel)
static int make_map_f32(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_cols, int box_rows) {
  static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
  }
  HFT_REQUIRE(enc != nullptr, HFT_ERR_STATE, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HFT_REQUIRE(r == CUDA_SUCCESS, HFT_ERR_STATE, "cuTensorMapEncodeTiled (fp32) failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r, rows, cols, ld, box_cols, box_rows);
  return HFT_OK;
}
static bool tdw_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HFT_TRAIN_TC_DW"); v = (e && e[0] == '0') ? 0 : 1; }
  return g_train_tc_force >= 0 ? g_train_tc_force == 1 : (v == 1 && train_tc_enabled());
}
static bool tdw_ok(const float* dY, int ldy, const float* X, int ldx, int N, int K) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int G = ((N + 63) / 64 + 1) / 2;
  return tdw_enabled() && (K == 64 || K == 128) && N % 8 == 0 && N >= 8 && N + K <= 256 && G * (64 + 3 * K) <= 512 && ldy % 4 == 0 && ldx % 4 == 0 && al(dY) && al(X) &&
         tdw_raw_stages(N, K, 2) >= 2;
}
static int tgemm_launch(cudaStream_t s, const tc::TGemmArgs& a0) {
  static const int sms = num_sms();
  tc::TGemmArgs a = a0;
  a.tmem_cols = a.N <= 32 ? 32 : a.N <= 64 ? 64 : a.N <= 128 ? 128 : 256;
  const size_t smem = tc::tgemm_smem(a.K / 64, a.N);
  // two CTAs per SM; HFT_TRAIN_GEMM_CTAS=3 cuts the register budget for three where shared memory (<= 75 KB each) and TMEM (<= 128 columns each)
  // allow it (experiment switch: measured slower, 46.5 against 41.2 us for [262144, 64] x [64, 64] -- the 85-register budget spills)
  static int max_b = -1;
  if (max_b < 0) { const char* e = getenv("HFT_TRAIN_GEMM_CTAS"); max_b = (e && atoi(e) == 3) ? 3 : 2; }
  const int minb = (max_b == 3 && smem <= 75 * 1024 && a.tmem_cols <= 128) ? 3 : 2;
  const long long tiles = (a.M + 127) / 128;
  const unsigned grid = (unsigned)(tiles < (long long)minb * sms ? tiles : (long long)minb * sms);
  LaunchScope ls(HFT_KCLASS_GEMM, s);
#define HFT_TGEMM_LAUNCH(KBv, Bv)                                                  \
  do {                                                                             \
    HFT_SET_MAX_SMEM((tc::tgemm_kernel<KBv, Bv>), 113 * 1024);                     \
    tc::tgemm_kernel<KBv, Bv><<<grid, tc::kGThreads, smem, s>>>(a);                \
  } while (0)
  if (a.K == 64 && minb == 3) HFT_TGEMM_LAUNCH(1, 3);
  else if (a.K == 64) HFT_TGEMM_LAUNCH(1, 2);
  else if (minb == 3) HFT_TGEMM_LAUNCH(2, 3);
  else HFT_TGEMM_LAUNCH(2, 2);
#undef HFT_TGEMM_LAUNCH
  return HFT_OK;
}

static int gemm_tn(cudaStream_t s, const float* A, int lda, const float* Wt, int ldw, const float* bias, float* C, int ldc, long long M, int N, int K,
                   bool relu, bool accum = false) {
  if (M > 0 && tgemm_ok(A, lda, Wt, ldw, false, C, ldc, N, K, nullptr, 0)) {
    tc::TGemmArgs a{};
    a.A = A; a.lda = lda; a.W = Wt; a.ldw = ldw; a.w_kn = 0; a.bias = bias; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K; a.relu = relu; a.accum = accum;
    return tgemm_launch(s, a);
  }
  HFT_REQUIRE(K % GBK == 0 && lda % 4 == 0 && ldw % 4 == 0, HFT_ERR_UNSUPPORTED, "train sgemm_tn: K=%d lda=%d ldw=%d", K, lda, ldw);
  const int bn = sgemm_tile_n(N);
  dim3 grid((N + bn - 1) / bn, (unsigned)((M + GBM - 1) / GBM));
  LaunchScope ls(HFT_KCLASS_GEMM, s);
  if (bn == 64) {
    if (relu) sgemm_tn_kernel<true, 64><<<grid, 256, 0, s>>>(A, lda, Wt, ldw, bias, C, ldc, (int)M, N, K, accum);
    else sgemm_tn_kernel<false, 64><<<grid, 256, 0, s>>>(A, lda, Wt, ldw, bias, C, ldc, (int)M, N, K, accum);
  } else {
    if (relu) sgemm_tn_kernel<true, 128><<<grid, 256, 0, s>>>(A, lda, Wt, ldw, bias, C, ldc, (int)M, N, K, accum);
    else sgemm_tn_kernel<false, 128><<<grid, 256, 0, s>>>(A, lda, Wt, ldw, bias, C, ldc, (int)M, N, K, accum);
  }
  return HFT_OK;
}
// C[M,N] (+)= A[M,K] * W[K,N]
static int gemm_nn(cudaStream_t s, const float* A, int lda, const float* W, int ldb, float* C, int ldc, long long M, int N, int K, bool accum,
                   const float* mask = nullptr, int ldm = 0, float mask_scale = 1.f) {
  if (M > 0 && tgemm_ok(A, lda, W, ldb, true, C, ldc, N, K, mask, ldm)) {
    tc::TGemmArgs a{};
    a.A = A; a.lda = lda; a.W = W; a.ldw = ldb; a.w_kn = 1; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K; a.accum = accum; a.mask = mask; a.ldm = ldm;
    a.mask_scale = mask_scale;
    return tgemm_launch(s, a);
  }
  HFT_REQUIRE(K % GBK == 0 && N % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0, HFT_ERR_UNSUPPORTED, "train sgemm_nn: N=%d K=%d lda=%d ldb=%d", N, K, lda, ldb);
  const int bn = sgemm_tile_n(N);
  dim3 grid((N + bn - 1) / bn, (unsigned)((M + GBM - 1) / GBM));
  LaunchScope ls(HFT_KCLASS_GEMM, s);
  if (bn == 64) sgemm_nn_kernel<64><<<grid, 256, 0, s>>>(A, lda, W, ldb, C, ldc, (int)M, N, K, accum, mask, ldm, mask_scale);
  else sgemm_nn_kernel<128><<<grid, 256, 0, s>>>(A, lda, W, ldb, C, ldc, (int)M, N, K, accum, mask, ldm, mask_scale);
  return HFT_OK;
}
// dW[N,K] += dY[M,N]^T X[M,K]; db[N] += colsum(dY)
static int gemm_dw(cudaStream_t s, const float* dY, int ldy, const float* X, int ldx, float* dW, int ldw, float* db, long long M, int N, int K) {
  HFT_REQUIRE(ldy % 4 == 0 && ldx % 4 == 0, HFT_ERR_UNSUPPORTED, "train dw gemm: ldy=%d ldx=%d", ldy, ldx);
  if (M > 0 && tdw_ok(dY, ldy, X, ldx, N, K)) {
    static const int sms_tc = num_sms();
    tc::TDwArgs a{};
    a.dY = dY; a.ldy = ldy; a.X = X; a.ldx = ldx; a.dW = dW; a.ldw = ldw; a.db = db; a.M = M; a.N = N; a.K = K;
    CUtensorMap map_y, map_x;
    HFT_TRY(make_map_f32(&map_y, dY, M, N, ldy, N, tc::kDwRows));
    HFT_TRY(make_map_f32(&map_x, X, M, K, ldx, K, tc::kDwRows));
    tdw_rings(N, K, &a.n_slots, &a.n_raw);
    long long ctas = sms_tc;                                        // one persistent CTA per SM, at least four 32-row stages each
    if (ctas > (M + 4 * tc::kDwRows - 1) / (4 * tc::kDwRows)) ctas = (M + 4 * tc::kDwRows - 1) / (4 * tc::kDwRows);
    a.rows_per_cta = ((M + ctas - 1) / ctas + tc::kDwRows - 1) / tc::kDwRows * tc::kDwRows;
    ctas = (M + a.rows_per_cta - 1) / a.rows_per_cta;
    const size_t smem = tc::tdw_smem(K / 64, N, a.n_slots, a.n_raw);
    const int tps = (N + K + 127) / 128;                              // 8-float chunks per producer thread and stage (1 or 2)
    LaunchScope ls(HFT_KCLASS_GEMM, s);
#define HFT_TDW_LAUNCH(KBv, Tv)                                                                       \
  do {                                                                                                \
    HFT_SET_MAX_SMEM((tc::tdw_kernel<KBv, Tv>), 220 * 1024);                                          \
    tc::tdw_kernel<KBv, Tv><<<(unsigned)ctas, tc::kDwThreads, smem, s>>>(map_y, map_x, a);                          \
  } while (0)
    if (K == 64 && tps == 1) HFT_TDW_LAUNCH(1, 1);
    else if (K == 64) HFT_TDW_LAUNCH(1, 2);
    else HFT_TDW_LAUNCH(2, 2);
#undef HFT_TDW_LAUNCH
    return HFT_OK;
  }
  const int gx = (N + DWT - 1) / DWT, gy = (K + DWT - 1) / DWT;
  // split M so that the grid fills the chip two (256-thread, 128-register) CTAs deep (the tile count gx * gy is 1..6 for this model: without the split a
  // [64 x 64] dW ran on 44..128 of the 148 SMs), with at least four stages per CTA so the closing atomics stay the minor part
  static const int sms = num_sms();
  long long splits = (2LL * sms + gx * gy - 1) / (gx * gy);
  if (splits > (M + 4 * DWR - 1) / (4 * DWR)) splits = (M + 4 * DWR - 1) / (4 * DWR);
  if (splits < 1) splits = 1;
  long long rps = ((M + splits - 1) / splits + DWR - 1) / DWR * DWR;
  splits = (M + rps - 1) / rps;
  HFT_SET_MAX_SMEM(dw_gemm_kernel, DW_SMEM);
  LaunchScope ls(HFT_KCLASS_GEMM, s);
  dw_gemm_kernel<<<dim3(gx, gy, (unsigned)splits), DWTHREADS, DW_SMEM, s>>>(dY, ldy, X, ldx, dW, ldw, db, M, N, K, rps);
  return HFT_OK;
}
static long long ln_grid(long long rows, int rpw) {              // CTAs of 8 warps x rpw rows, at most 16 per SM (grid-stride beyond that)
  static const int sms = num_sms();
  const long long need = (rows + 8 * rpw - 1) / (8 * rpw);
  return need < 16LL * sms ? need : 16LL * sms;
}
static void ln_fwd(Model* m, cudaStream_t s, const float* x, const float* r, long long r_rows, const LnW& ln, long long rows, float* y, float* sum_out,
                   Drop drop = Drop{0, 0, 0, 1.f}, bool drop_x = false) {
  LaunchScope ls(HFT_KCLASS_NORM, s);
  const float *g = m->w[ln.g], *b = m->w[ln.b];
  if (m->H == 64) add_ln_v8_kernel<8><<<(unsigned)ln_grid(rows, 4), 256, 0, s>>>(x, r, r_rows, g, b, rows, y, sum_out, drop, drop_x);
  else if (m->H == 128) add_ln_v8_kernel<16><<<(unsigned)ln_grid(rows, 2), 256, 0, s>>>(x, r, r_rows, g, b, rows, y, sum_out, drop, drop_x);
  else if (m->H == 256) add_ln_v8_kernel<32><<<(unsigned)ln_grid(rows, 1), 256, 0, s>>>(x, r, r_rows, g, b, rows, y, sum_out, drop, drop_x);
  else add_ln_f32_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(x, r, r_rows, g, b, m->H, rows, y, sum_out, drop, drop_x);
}
// element-wise dropout in place (forward of an embedding / hidden dropout, or the backward of one)
static void dropout_inplace(cudaStream_t s, float* x, long long n, const Drop& d) {
  if (!d.thresh) return;
  LaunchScope ls(HFT_KCLASS_NORM, s);
  dropout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, n, d);
}
// gradient of the sub-layer branch under dropout: g itself when p = 0, else a masked copy in `scratch`
static const float* branch_grad(cudaStream_t s, const float* g, float* scratch, long long n, const Drop& d) {
  if (!d.thresh) return g;
  LaunchScope ls(HFT_KCLASS_NORM, s);
  dropout_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(g, scratch, n, d);
  return scratch;
}
static void ln_bwd(Model* m, cudaStream_t s, const float* dy, const float* sum, const LnW& ln, long long rows, float* ds, float* G) {
  LaunchScope ls(HFT_KCLASS_NORM, s);
  float *dg = G + (m->w[ln.g] - m->arena), *db = G + (m->w[ln.b] - m->arena);
  static const int sms = num_sms();
  // few, long-lived CTAs: every CTA closes with 2 H atomics
  const unsigned grid = (unsigned)(rows < 4LL * sms * 64 ? (rows + 63) / 64 : 4LL * sms);
  if (m->H == 64) ln_bwd_v8_kernel<8><<<grid, 256, 0, s>>>(dy, sum, m->w[ln.g], rows, ds, dg, db);
  else if (m->H == 128) ln_bwd_v8_kernel<16><<<grid, 256, 0, s>>>(dy, sum, m->w[ln.g], rows, ds, dg, db);
  else if (m->H == 256) ln_bwd_v8_kernel<32><<<grid, 256, 0, s>>>(dy, sum, m->w[ln.g], rows, ds, dg, db);
  else ln_bwd_kernel<<<(unsigned)((rows + 63) / 64), 256, 0, s>>>(dy, sum, m->w[ln.g], m->H, rows, ds, dg, db);
}
static void colsum(cudaStream_t s, const float* in, long long rows, long long cols, float* out) {
  long long splits = rows >= 64 ? 16 : 1;
  long long rps = (rows + splits - 1) / splits;
  LaunchScope ls(HFT_KCLASS_NORM, s);
  colsum_kernel<<<dim3((unsigned)((cols + 255) / 256), (unsigned)splits), 256, 0, s>>>(in, rows, cols, rps, out);
}

static bool attn_r2_enabled() {                    // HFT_TRAIN_ATTN_R2=0: the one-row-per-thread kernels
  static int v = -1;
  if (v < 0) { const char* e = getenv("HFT_TRAIN_ATTN_R2"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

struct AttnDims { int dh, heads, H; };
static bool tattn_ok(const AttnDims& d, int ldq, int ldkv, int Lq, int Lk) {
  return train_tc_enabled() && d.dh == 32 && Lq >= 1 && Lk >= 1 && Lq <= 256 && Lk <= 256 && ldq % 4 == 0 && ldkv % 4 == 0 && d.H % 4 == 0;
}

static int attn_fwd(const AttnDims& m, cudaStream_t s, const float* Q, int ldq, long long q_seq_stride, const float* K, const float* V, int ldkv, long long S,
                    int Lq, int Lk, float* ctx, float* lse, Drop drop = Drop{0, 0, 0, 1.f}) {
  const int dh = m.dh;
  const float inv_scale = 1.f / sqrtf((float)dh);
  LaunchScope ls(HFT_KCLASS_ATTENTION, s);
  if (tattn_ok(m, ldq, ldkv, Lq, Lk)) {
    tc::TAttnArgs a{};
    a.Q = Q; a.ldq = ldq; a.q_seq_stride = q_seq_stride; a.K = K; a.V = V; a.ldkv = ldkv; a.Lq = Lq; a.Lk = Lk; a.heads = m.heads; a.c = inv_scale;
    a.ctx = ctx; a.ldo = m.H; a.lse = lse; a.drop = drop;
    const int lkp = (Lk + 127) / 128 * 128;
    HFT_SET_MAX_SMEM(tc::tattn_fwd_kernel, tc::tattn_fwd_smem(256));
    tc::tattn_fwd_kernel<<<dim3((unsigned)S, m.heads, (Lq + 127) / 128), tc::kTThreads, tc::tattn_fwd_smem(lkp), s>>>(a, lkp);
    return HFT_OK;
  }
  size_t smem = (size_t)2 * Lk * dh * sizeof(float);
  int threads = (Lq + 31) / 32 * 32;
  HFT_REQUIRE(threads <= 256 && smem <= 200 * 1024, HFT_ERR_UNSUPPORTED, "train attention: Lq=%d Lk=%d unsupported", Lq, Lk);
  dim3 grid((unsigned)S, m.heads);
  if (dh == 64) {
    HFT_CHECK_CUDA(cudaFuncSetAttribute(attn_f32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attn_f32_kernel<64><<<grid, threads, smem, s>>>(Q, ldq, q_seq_stride, K, V, ldkv, Lq, Lk, m.heads, inv_scale, ctx, m.H, nullptr, lse, drop);
  } else if (attn_r2_enabled()) {                    // two query rows per thread, one-pass softmax
    HFT_CHECK_CUDA(cudaFuncSetAttribute(attn_f32_r2_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attn_f32_r2_kernel<32><<<grid, ((Lq + 1) / 2 + 31) / 32 * 32, smem, s>>>(Q, ldq, q_seq_stride, K, V, ldkv, Lq, Lk, m.heads, inv_scale, ctx, m.H, lse, drop);
  } else {
    HFT_CHECK_CUDA(cudaFuncSetAttribute(attn_f32_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attn_f32_kernel<32><<<grid, threads, smem, s>>>(Q, ldq, q_seq_stride, K, V, ldkv, Lq, Lk, m.heads, inv_scale, ctx, m.H, nullptr, lse, drop);
  }
  return HFT_OK;
}
static int attn_fwd(Model* m, cudaStream_t s, const float* Q, int ldq, long long q_seq_stride, const float* K, const float* V, int ldkv, long long S,
                    int Lq, int Lk, float* ctx, float* lse, Drop drop = Drop{0, 0, 0, 1.f}) {
  return attn_fwd(AttnDims{m->dh, m->heads, m->H}, s, Q, ldq, q_seq_stride, K, V, ldkv, S, Lq, Lk, ctx, lse, drop);
}

template <int DH>
static int attn_bwd_t(const AttnDims& m, cudaStream_t s, const float* Q, int ldq, long long qss, const float* K, const float* V, int ldkv, const float* dO,
                      const float* O, const float* lse, long long S, int Lq, int Lk, float* dQ, int lddq, float* dK, float* dV, int lddkv, float* Dbuf,
                      Drop drop) {
  const float c = 1.f / sqrtf((float)DH);
  LaunchScope ls(HFT_KCLASS_ATTENTION, s);
  if (DH == 32 && tattn_ok(m, ldq, ldkv, Lq, Lk) && lddq % 4 == 0 && lddkv % 4 == 0) {
    tc::TAttnArgs a{};
    a.Q = Q; a.ldq = ldq; a.q_seq_stride = qss; a.K = K; a.V = V; a.ldkv = ldkv; a.Lq = Lq; a.Lk = Lk; a.heads = m.heads; a.c = c;
    a.ctx = const_cast<float*>(O); a.ldo = m.H; a.lse = const_cast<float*>(lse); a.dO = dO; a.dQ = dQ; a.lddq = lddq; a.dK = dK; a.dV = dV; a.lddkv = lddkv;
    a.Dbuf = Dbuf; a.drop = drop;
    HFT_SET_MAX_SMEM(tc::tattn_bwd_kernel<false>, tc::tattn_bwd_smem(256));
    HFT_SET_MAX_SMEM(tc::tattn_bwd_kernel<true>, tc::tattn_bwd_smem(256));
    const int lcp_k = (Lk + 63) / 64 * 64, lcp_q = (Lq + 63) / 64 * 64;
    tc::tattn_bwd_kernel<false><<<dim3((unsigned)S, m.heads, (Lq + 127) / 128), tc::kTThreads, tc::tattn_bwd_smem(lcp_k), s>>>(a, lcp_k);   // dQ, D
    tc::tattn_bwd_kernel<true><<<dim3((unsigned)S, m.heads, (Lk + 127) / 128), tc::kTThreads, tc::tattn_bwd_smem(lcp_q), s>>>(a, lcp_q);    // dK, dV
    return HFT_OK;
  }
  dim3 grid((unsigned)S, m.heads);
  const size_t smem1 = (size_t)2 * Lk * DH * sizeof(float);
  const size_t smem2 = ((size_t)2 * Lq * DH + 2 * Lq) * sizeof(float);
  const int t1 = (Lq + 31) / 32 * 32, t2 = (Lk + 31) / 32 * 32;
  HFT_REQUIRE(t1 <= 256 && t2 <= 256 && smem1 <= 200 * 1024 && smem2 <= 200 * 1024, HFT_ERR_UNSUPPORTED, "train attention backward: Lq=%d Lk=%d", Lq, Lk);
  constexpr int DHR = DH <= 32 ? DH : 32;          // the two-row kernel exists for head_dim <= 32 only
  if (DH <= 32 && attn_r2_enabled() && lddq % 4 == 0) {
    HFT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_r2_kernel<DHR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attn_bwd_dq_r2_kernel<DHR><<<grid, ((Lq + 1) / 2 + 31) / 32 * 32, smem1, s>>>(Q, ldq, qss, K, V, ldkv, dO, O, m.H, lse, Lq, Lk, m.heads, c, dQ, lddq, Dbuf, drop);
  } else {
    HFT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attn_bwd_dq_kernel<DH><<<grid, t1, smem1, s>>>(Q, ldq, qss, K, V, ldkv, dO, O, m.H, lse, Lq, Lk, m.heads, c, dQ, lddq, Dbuf, drop);
  }
  if (DH == 32 && attn_r2_enabled() && lddkv % 4 == 0) {        // thread pairs: two keys x half the head dimension each
    HFT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attn_bwd_dkv_pair_kernel<<<grid, t2, smem2, s>>>(Q, ldq, qss, K, V, ldkv, dO, m.H, lse, Dbuf, Lq, Lk, m.heads, c, dK, dV, lddkv, drop);
  } else if (DH <= 32) {
    HFT_CHECK_CUDA(cudaFuncSetAttribute((attn_bwd_dkv_kernel<DH, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attn_bwd_dkv_kernel<DH, 0><<<grid, t2, smem2, s>>>(Q, ldq, qss, K, V, ldkv, dO, m.H, lse, Dbuf, Lq, Lk, m.heads, c, dK, dV, lddkv, drop);
  } else {
    HFT_CHECK_CUDA(cudaFuncSetAttribute((attn_bwd_dkv_kernel<DH, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    HFT_CHECK_CUDA(cudaFuncSetAttribute((attn_bwd_dkv_kernel<DH, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attn_bwd_dkv_kernel<DH, 1><<<grid, t2, smem2, s>>>(Q, ldq, qss, K, V, ldkv, dO, m.H, lse, Dbuf, Lq, Lk, m.heads, c, dK, dV, lddkv, drop);
    attn_bwd_dkv_kernel<DH, 2><<<grid, t2, smem2, s>>>(Q, ldq, qss, K, V, ldkv, dO, m.H, lse, Dbuf, Lq, Lk, m.heads, c, dK, dV, lddkv, drop);
  }
  return HFT_OK;
}
static int attn_bwd(const AttnDims& m, cudaStream_t s, const float* Q, int ldq, long long qss, const float* K, const float* V, int ldkv, const float* dO, const float* O,
                    const float* lse, long long S, int Lq, int Lk, float* dQ, int lddq, float* dK, float* dV, int lddkv, float* Dbuf, Drop drop) {
  if (m.dh == 64) return attn_bwd_t<64>(m, s, Q, ldq, qss, K, V, ldkv, dO, O, lse, S, Lq, Lk, dQ, lddq, dK, dV, lddkv, Dbuf, drop);
  return attn_bwd_t<32>(m, s, Q, ldq, qss, K, V, ldkv, dO, O, lse, S, Lq, Lk, dQ, lddq, dK, dV, lddkv, Dbuf, drop);
}
static int attn_bwd(Model* m, cudaStream_t s, const float* Q, int ldq, long long qss, const float* K, const float* V, int ldkv, const float* dO, const float* O,
                    const float* lse, long long S, int Lq, int Lk, float* dQ, int lddq, float* dK, float* dV, int lddkv, float* Dbuf, Drop drop) {
  return attn_bwd(AttnDims{m->dh, m->heads, m->H}, s, Q, ldq, qss, K, V, ldkv, dO, O, lse, S, Lq, Lk, dQ, lddq, dK, dV, lddkv, Dbuf, drop);
}

// gradient slot of a registered parameter inside the flat gradient vector
static inline float* gof(Model* m, float* G, int idx) { return G + (m->w[idx] - m->arena); }

// ---- tape allocation ------------------------------------------------------------------------------------------------
static int alloc_tape(Trainer& t) {
  Model* m = t.m;
  const long long Re = (long long)t.B * m->nframe * m->nbin, Rd = (long long)t.B * m->nframe * m->nnote;
  const long long H = m->H, P = m->P, heads = m->heads;
  const long long Se = (long long)t.B * m->nframe, St = (long long)t.B * m->nnote;
  const size_t n_enc = m->enc.size(), n_tim = m->tim.size(), n_dec = 1 + m->dec.size();
  auto layer_floats = [&](long long R, long long S, long long L) { return R * (8 * H + P) + S * heads * L + 64 * 8; };
  size_t fl = 0;
  fl += n_enc * layer_floats(Re, Se, m->nbin) + Re * H;                                   // encoder layers + encoder output
  fl += n_tim * layer_floats(Rd, St, m->nframe) + 2 * Rd * H;                             // time layers + u0 / u_out
  fl += n_dec * (Rd * (12 * H + P) + Re * 2 * H + 2 * Se * heads * m->nnote + 64 * 16) + Rd * H;   // decoder layers + t_out
  fl += 2 * Rd * t.NP + 2 * ((size_t)t.NP * H + t.NP) + (size_t)t.NP * H + t.NP + (size_t)H * m->nproc + H + 64 * 16;
  fl += Re * (2 * H + 3 * H + H + P) + Rd * (H + 3 * H + H) + Rd * t.NP + Re * heads + (size_t)m->nnote * H + 64 * 12;              // gradient work buffers
  t.arena_bytes = fl * sizeof(float);
  HFT_CHECK_CUDA(cudaMalloc(&t.arena, t.arena_bytes));
  float* p = t.arena;
  auto take = [&](long long n) { float* r = p; p += (n + 63) & ~63ll; return r; };
  auto mk_layer = [&](LayerTape& L, long long R, long long S, long long Lq) {
    L.xin = take(R * H); L.qkv = take(R * 3 * H); L.lse = take(S * heads * Lq); L.ctx = take(R * H); L.s1 = take(R * H); L.x1 = take(R * H);
    L.hid = take(R * P); L.s2 = take(R * H);
  };
  t.enc.resize(n_enc); t.tim.resize(n_tim); t.dec.resize(n_dec);
  for (auto& L : t.enc) mk_layer(L, Re, Se, m->nbin);
  t.x_enc = take(Re * H);
  for (size_t i = 0; i < n_dec; ++i) {
    DecTape& D = t.dec[i];
    D.tin = take(Rd * H); D.qkv = take(Rd * 3 * H); D.lse_s = take(Se * heads * m->nnote); D.ctx_s = take(Rd * H); D.s0 = take(Rd * H); D.t0 = take(Rd * H);
    D.qc = take(Rd * H); D.kv = take(Re * 2 * H); D.lse_c = take(Se * heads * m->nnote); D.ctx_c = take(Rd * H); D.s1 = take(Rd * H); D.t1 = take(Rd * H);
    D.hid = take(Rd * P); D.s2 = take(Rd * H);
  }
  t.t_out = take(Rd * H);
  t.u0 = take(Rd * H);
  for (auto& L : t.tim) mk_layer(L, Rd, St, m->nframe);
  t.u_out = take(Rd * H);
  t.logits_a = take(Rd * t.NP); t.logits_b = take(Rd * t.NP);
  for (int i = 0; i < 2; ++i) { t.head_w[i] = take((long long)t.NP * H); t.head_b[i] = take(t.NP); }
  t.g_head_w = take((long long)t.NP * H); t.g_head_b = take(t.NP);
  t.g_front_w = take(H * m->nproc); t.g_front_b = take(H);
  t.gX = take(Re * H); t.gBIG = take(Re * 3 * H); t.gCTX = take(Re * H); t.gHID = take(Re * P);
  t.gT = take(Rd * H); t.gDQ = take(Rd * 3 * H); t.gU = take(Rd * H); t.gLOG = take(Rd * t.NP); t.dD = take(Re * heads); t.gQ0 = take((long long)m->nnote * H);
  t.gM = take(Re * H);
  HFT_REQUIRE((size_t)(p - t.arena) * sizeof(float) <= t.arena_bytes, HFT_ERR_STATE, "trainer tape overflow (%zu > %zu)", (size_t)(p - t.arena) * sizeof(float), t.arena_bytes);
  return HFT_OK;
}

// ---- forward with tape ----------------------------------------------------------------------------------------------
// EncoderLayer (model_spec2midi.py:230-245): x = L.xin -> out
// dropout sites of the layer: base + 0 attention probabilities, + 1 attention sub-layer output, + 2 FFN hidden, + 3 FFN sub-layer output
static int enc_layer_fwd(Trainer& t, cudaStream_t s, LayerTape& L, long long S, int Lq, const EncLayerW& lw, const FusedAttn& qkv, float* out, float* tmp, int base) {
  Model* m = t.m;
  const int H = m->H, P = m->P;
  const long long R = S * Lq;
  HFT_TRY(gemm_tn(s, L.xin, H, qkv.qkv_w, H, qkv.qkv_b, L.qkv, 3 * H, R, 3 * H, H, false));
  HFT_TRY(attn_fwd(m, s, L.qkv, 3 * H, (long long)Lq * 3 * H, L.qkv + H, L.qkv + 2 * H, 3 * H, S, Lq, Lq, L.ctx, L.lse, t.drop(base)));
  HFT_TRY(gemm_tn(s, L.ctx, H, m->w[lw.sa.o_w], H, m->w[lw.sa.o_b], tmp, H, R, H, H, false));
  ln_fwd(m, s, L.xin, tmp, R, lw.ln, R, L.x1, L.s1, t.drop(base + 1));
  HFT_TRY(gemm_tn(s, L.x1, H, m->w[lw.ff.w1], H, m->w[lw.ff.b1], L.hid, P, R, P, H, true));
  dropout_inplace(s, L.hid, R * P, t.drop(base + 2));
  HFT_TRY(gemm_tn(s, L.hid, P, m->w[lw.ff.w2], P, m->w[lw.ff.b2], tmp, H, R, H, P, false));
  ln_fwd(m, s, L.x1, tmp, R, lw.ln, R, out, L.s2, t.drop(base + 3));
  return HFT_OK;
}

// EncoderLayer backward: g = dL/d(out) on entry, dL/d(xin) on exit (in place)
static int enc_layer_bwd(Trainer& t, cudaStream_t s, LayerTape& L, long long S, int Lq, const EncLayerW& lw, const FusedAttn& qkv, float* g, float* gqkv, float* gctx,
                         float* ghid, float* G, int base) {
  Model* m = t.m;
  const int H = m->H, P = m->P;
  const long long R = S * Lq;
  ln_bwd(m, s, g, L.s2, lw.ln, R, g, G);
  const float* gb = branch_grad(s, g, t.gM, R * H, t.drop(base + 3));                       // gradient of the FFN branch (sub-layer dropout)
  HFT_TRY(gemm_dw(s, gb, H, L.hid, P, gof(m, G, lw.ff.w2), P, gof(m, G, lw.ff.b2), R, H, P));
  HFT_TRY(gemm_nn(s, gb, H, m->w[lw.ff.w2], P, ghid, P, R, P, H, false, L.hid, P, t.drop(base + 2).scale));   // through fc_2, the hidden dropout and the ReLU
  HFT_TRY(gemm_dw(s, ghid, P, L.x1, H, gof(m, G, lw.ff.w1), H, gof(m, G, lw.ff.b1), R, P, H));
  HFT_TRY(gemm_nn(s, ghid, P, m->w[lw.ff.w1], H, g, H, R, H, P, true));                     // + residual path already in g
  ln_bwd(m, s, g, L.s1, lw.ln, R, g, G);
  gb = branch_grad(s, g, t.gM, R * H, t.drop(base + 1));                                    // gradient of the attention branch
  HFT_TRY(gemm_dw(s, gb, H, L.ctx, H, gof(m, G, lw.sa.o_w), H, gof(m, G, lw.sa.o_b), R, H, H));
  HFT_TRY(gemm_nn(s, gb, H, m->w[lw.sa.o_w], H, gctx, H, R, H, H, false));
  HFT_TRY(attn_bwd(m, s, L.qkv, 3 * H, (long long)Lq * 3 * H, L.qkv + H, L.qkv + 2 * H, 3 * H, gctx, L.ctx, L.lse, S, Lq, Lq, gqkv, 3 * H, gqkv + H, gqkv + 2 * H,
                   3 * H, t.dD, t.drop(base)));
  HFT_TRY(gemm_dw(s, gqkv, 3 * H, L.xin, H, gof(m, G, lw.sa.q_w), H, gof(m, G, lw.sa.q_b), R, H, H));
  HFT_TRY(gemm_dw(s, gqkv + H, 3 * H, L.xin, H, gof(m, G, lw.sa.k_w), H, gof(m, G, lw.sa.k_b), R, H, H));
  HFT_TRY(gemm_dw(s, gqkv + 2 * H, 3 * H, L.xin, H, gof(m, G, lw.sa.v_w), H, gof(m, G, lw.sa.v_b), R, H, H));
  HFT_TRY(gemm_nn(s, gqkv, 3 * H, qkv.qkv_w, H, g, H, R, H, 3 * H, true));
  return HFT_OK;
}

static int pack_heads(Trainer& t, cudaStream_t s) {
  Model* m = t.m;
  for (int i = 0; i < 2; ++i) {
    const int* idx = i == 0 ? m->head_freq : m->head_time;
    pack_heads_f32_kernel<<<(t.NP * m->H + 255) / 256, 256, 0, s>>>(m->w[idx[0]], m->w[idx[2]], m->w[idx[4]], m->w[idx[6]], m->w[idx[1]], m->w[idx[3]],
                                                                     m->w[idx[5]], m->w[idx[7]], m->nvel, m->H, t.NP, t.head_w[i], t.head_b[i]);
  }
  return HFT_OK;
}

static int train_forward(Trainer& t, const float* spec, long long sb, long long sbin, long long st, cudaStream_t s) {
  Model* m = t.m;
  const int B = t.Bc, H = m->H, P = m->P, F = m->nframe, NB = m->nbin, NN = m->nnote;
  const long long Se = (long long)B * F, Re = Se * NB, Rd = Se * NN;
  const float sqrtH = sqrtf((float)H);
  float* tmp = t.gX;                                      // scratch [Re, H] (the gradient buffers are idle during the forward)
  {
    LaunchScope ls(HFT_KCLASS_FRONT, s);
    front_f32_kernel<65><<<dim3(NB, B), 256, 0, s>>>(spec, sb, sbin, st, m->front_w, m->front_b, m->w[m->pos_freq], sqrtH, H, F, NB, t.enc[0].xin);
  }
  dropout_inplace(s, t.enc[0].xin, Re * H, t.drop(0));                                   // embedding dropout, model_spec2midi.py:95
  for (size_t l = 0; l < m->enc.size(); ++l)
    HFT_TRY(enc_layer_fwd(t, s, t.enc[l], Se, NB, m->enc[l], m->enc_qkv[l], l + 1 < m->enc.size() ? t.enc[l + 1].xin : t.x_enc, tmp, t.site_enc((int)l)));
  // decoder layer zero (model_spec2midi.py:255-272)
  {
    DecTape& D = t.dec[0];
    const DecLayerW& lw = m->dec0;
    HFT_TRY(gemm_tn(s, t.x_enc, H, m->dec_ca_kv[0].qkv_w, H, m->dec_ca_kv[0].qkv_b, D.kv, 2 * H, Re, 2 * H, H, false));
    const int b0 = t.site_dec0();                            // + 0 cross probabilities, + 1 cross output, + 2 hidden, + 3 FFN output
    HFT_TRY(attn_fwd(m, s, m->q0, H, 0, D.kv, D.kv + H, 2 * H, Se, NN, NB, D.ctx_c, D.lse_c, t.drop(b0)));
    HFT_TRY(gemm_tn(s, D.ctx_c, H, m->w[lw.ca.o_w], H, m->w[lw.ca.o_b], tmp, H, Rd, H, H, false));
    ln_fwd(m, s, tmp, m->w[m->dec_pos_freq], NN, lw.ln, Rd, D.t1, D.s1, t.drop(b0 + 1), true);   // the sub-layer output is the first operand here
    HFT_TRY(gemm_tn(s, D.t1, H, m->w[lw.ff.w1], H, m->w[lw.ff.b1], D.hid, P, Rd, P, H, true));
    dropout_inplace(s, D.hid, Rd * P, t.drop(b0 + 2));
    HFT_TRY(gemm_tn(s, D.hid, P, m->w[lw.ff.w2], P, m->w[lw.ff.b2], tmp, H, Rd, H, P, false));
    ln_fwd(m, s, D.t1, tmp, Rd, lw.ln, Rd, m->dec.empty() ? t.t_out : t.dec[1].tin, D.s2, t.drop(b0 + 3));
  }
  for (size_t l = 0; l < m->dec.size(); ++l) {               // model_spec2midi.py:283-306
    DecTape& D = t.dec[l + 1];
    const DecLayerW& lw = m->dec[l];
    float* out = l + 1 < m->dec.size() ? t.dec[l + 2].tin : t.t_out;
    HFT_TRY(gemm_tn(s, D.tin, H, m->dec_sa_qkv[l].qkv_w, H, m->dec_sa_qkv[l].qkv_b, D.qkv, 3 * H, Rd, 3 * H, H, false));
    const int bl = t.site_dec((int)l);                       // + 0 self probabilities, + 1 self output, + 2 cross probabilities, + 3 cross output, + 4 hidden, + 5 FFN output
    HFT_TRY(attn_fwd(m, s, D.qkv, 3 * H, (long long)NN * 3 * H, D.qkv + H, D.qkv + 2 * H, 3 * H, Se, NN, NN, D.ctx_s, D.lse_s, t.drop(bl)));
    HFT_TRY(gemm_tn(s, D.ctx_s, H, m->w[lw.sa.o_w], H, m->w[lw.sa.o_b], tmp, H, Rd, H, H, false));
    ln_fwd(m, s, D.tin, tmp, Rd, lw.ln, Rd, D.t0, D.s0, t.drop(bl + 1));
    HFT_TRY(gemm_tn(s, D.t0, H, m->w[lw.ca.q_w], H, m->w[lw.ca.q_b], D.qc, H, Rd, H, H, false));
    HFT_TRY(gemm_tn(s, t.x_enc, H, m->dec_ca_kv[l + 1].qkv_w, H, m->dec_ca_kv[l + 1].qkv_b, D.kv, 2 * H, Re, 2 * H, H, false));
    HFT_TRY(attn_fwd(m, s, D.qc, H, (long long)NN * H, D.kv, D.kv + H, 2 * H, Se, NN, NB, D.ctx_c, D.lse_c, t.drop(bl + 2)));
    HFT_TRY(gemm_tn(s, D.ctx_c, H, m->w[lw.ca.o_w], H, m->w[lw.ca.o_b], tmp, H, Rd, H, H, false));
    ln_fwd(m, s, D.t0, tmp, Rd, lw.ln, Rd, D.t1, D.s1, t.drop(bl + 3));
    HFT_TRY(gemm_tn(s, D.t1, H, m->w[lw.ff.w1], H, m->w[lw.ff.b1], D.hid, P, Rd, P, H, true));
    dropout_inplace(s, D.hid, Rd * P, t.drop(bl + 4));
    HFT_TRY(gemm_tn(s, D.hid, P, m->w[lw.ff.w2], P, m->w[lw.ff.b2], tmp, H, Rd, H, P, false));
    ln_fwd(m, s, D.t1, tmp, Rd, lw.ln, Rd, out, D.s2, t.drop(bl + 5));
  }
  HFT_TRY(pack_heads(t, s));
  HFT_TRY(gemm_tn(s, t.t_out, H, t.head_w[0], H, t.head_b[0], t.logits_a, t.NP, Rd, t.NP, H, false));
  {
    LaunchScope ls(HFT_KCLASS_NORM, s);
    time_relayout_f32_kernel<<<(unsigned)((Rd * H + 255) / 256), 256, 0, s>>>(t.t_out, m->w[m->pos_time], sqrtH, F, NN, H, Rd * H, t.tim.empty() ? t.u_out : t.tim[0].xin);
  }
  dropout_inplace(s, t.tim.empty() ? t.u_out : t.tim[0].xin, Rd * H, t.drop(t.site_time_emb()));   // model_spec2midi.py:191
  for (size_t l = 0; l < m->tim.size(); ++l)
    HFT_TRY(enc_layer_fwd(t, s, t.tim[l], (long long)B * NN, F, m->tim[l], m->tim_qkv[l], l + 1 < m->tim.size() ? t.tim[l + 1].xin : t.u_out, tmp, t.site_time((int)l)));
  HFT_TRY(gemm_tn(s, t.u_out, H, t.head_w[1], H, t.head_b[1], t.logits_b, t.NP, Rd, t.NP, H, false));
  return HFT_OK;
}

// ---- backward -------------------------------------------------------------------------------------------------------
static int heads_bwd(Trainer& t, cudaStream_t s, int which, const float* x, float* gx, bool accum, float* G) {
  Model* m = t.m;
  const int H = m->H;
  const long long Rd = (long long)t.Bc * m->nframe * m->nnote;
  const int* idx = which == 0 ? m->head_freq : m->head_time;
  HFT_CHECK_CUDA(cudaMemsetAsync(t.g_head_w, 0, (size_t)t.NP * H * sizeof(float), s));
  HFT_CHECK_CUDA(cudaMemsetAsync(t.g_head_b, 0, (size_t)t.NP * sizeof(float), s));
  HFT_TRY(gemm_dw(s, t.gLOG, t.NP, x, H, t.g_head_w, H, t.g_head_b, Rd, t.NP, H));
  unpack_heads_grad_kernel<<<((3 + m->nvel) * H + 255) / 256, 256, 0, s>>>(t.g_head_w, t.g_head_b, m->nvel, H, gof(m, G, idx[0]), gof(m, G, idx[2]), gof(m, G, idx[4]),
                                                                           gof(m, G, idx[6]), gof(m, G, idx[1]), gof(m, G, idx[3]), gof(m, G, idx[5]), gof(m, G, idx[7]));
  HFT_TRY(gemm_nn(s, t.gLOG, t.NP, t.head_w[which], H, gx, H, Rd, H, t.NP, accum));
  return HFT_OK;
}

// FFN + LayerNorm block of a decoder layer: g = dL/d(out) -> dL/d(t1)
static int dec_ffn_bwd(Trainer& t, cudaStream_t s, DecTape& D, const DecLayerW& lw, long long Rd, float* g, float* G, int site_hid, int site_out) {
  Model* m = t.m;
  const int H = m->H, P = m->P;
  ln_bwd(m, s, g, D.s2, lw.ln, Rd, g, G);
  const float* gb = branch_grad(s, g, t.gM, Rd * H, t.drop(site_out));
  HFT_TRY(gemm_dw(s, gb, H, D.hid, P, gof(m, G, lw.ff.w2), P, gof(m, G, lw.ff.b2), Rd, H, P));
  HFT_TRY(gemm_nn(s, gb, H, m->w[lw.ff.w2], P, t.gHID, P, Rd, P, H, false, D.hid, P, t.drop(site_hid).scale));
  HFT_TRY(gemm_dw(s, t.gHID, P, D.t1, H, gof(m, G, lw.ff.w1), H, gof(m, G, lw.ff.b1), Rd, P, H));
  HFT_TRY(gemm_nn(s, t.gHID, P, m->w[lw.ff.w1], H, g, H, Rd, H, P, true));
  return HFT_OK;
}

// Where dL/dlogits comes from: the library's own 8-term loss (train.py:139-151; hft_train_forward_backward) or the gradients autograd hands
// back for the eight head outputs (hft_train_backward: the host computed the loss itself from the tensors hft_train_forward returned).
struct GradSource {
  const float *y_on = nullptr, *y_off = nullptr, *y_mpe = nullptr;
  const long long* y_vel = nullptr;
  float wA = 1.f, wB = 1.f;
  float* loss = nullptr;
  const hft_outputs* g = nullptr;          // non-NULL: output gradients (any member may be NULL = zero gradient)
};

static void head_logit_grads(Trainer& t, cudaStream_t s, int which, const GradSource& gs) {
  Model* m = t.m;
  const int F = m->nframe, NN = m->nnote, V = m->nvel;
  const long long Rd = (long long)t.Bc * F * NN;
  const float* logits = which ? t.logits_b : t.logits_a;
  LaunchScope ls(HFT_KCLASS_HEADS, s);
  if (gs.g) {
    const hft_outputs* g = gs.g;
    heads_outgrad_kernel<<<(unsigned)((Rd + 7) / 8), 256, 0, s>>>(logits, t.NP, V, F, NN, Rd, which != 0, which ? g->onset_B : g->onset_A, which ? g->offset_B : g->offset_A,
                                                                  which ? g->mpe_B : g->mpe_A, which ? g->velocity_B : g->velocity_A, t.gLOG);
  } else {
    loss_grad_kernel<<<(unsigned)((Rd + 7) / 8), 256, 0, s>>>(logits, t.NP, V, F, NN, Rd, which != 0, gs.y_on, gs.y_off, gs.y_mpe, gs.y_vel, which ? gs.wB : gs.wA, t.gLOG, gs.loss);
  }
}

static int train_backward(Trainer& t, const float* spec, long long sb, long long sbin, long long st, const GradSource& gs, float* G, cudaStream_t s) {
  Model* m = t.m;
  const int B = t.Bc, H = m->H, F = m->nframe, NB = m->nbin, NN = m->nnote;
  const long long Se = (long long)B * F, Re = Se * NB, Rd = Se * NN;
  const float sqrtH = sqrtf((float)H);
  if (gs.loss) HFT_CHECK_CUDA(cudaMemsetAsync(gs.loss, 0, sizeof(float), s));
  // ---- heads B + time stack ----
  head_logit_grads(t, s, 1, gs);
  HFT_TRY(heads_bwd(t, s, 1, t.u_out, t.gU, false, G));
  for (int l = (int)m->tim.size() - 1; l >= 0; --l)
    HFT_TRY(enc_layer_bwd(t, s, t.tim[l], (long long)B * NN, F, m->tim[l], m->tim_qkv[l], t.gU, t.gDQ, t.gCTX, t.gHID, G, t.site_time(l)));
  dropout_inplace(s, t.gU, Rd * H, t.drop(t.site_time_emb()));                           // back through the time-embedding dropout
  colsum(s, t.gU, (long long)B * NN, (long long)F * H, gof(m, G, m->pos_time));          // pos_embedding_time
  // ---- heads A + re-layout ----
  head_logit_grads(t, s, 0, gs);
  HFT_TRY(heads_bwd(t, s, 0, t.t_out, t.gT, false, G));
  {
    LaunchScope ls(HFT_KCLASS_NORM, s);
    time_relayout_bwd_kernel<<<(unsigned)((Rd * H + 255) / 256), 256, 0, s>>>(t.gU, sqrtH, F, NN, H, Rd * H, t.gT);
  }
  // ---- decoder ----
  HFT_CHECK_CUDA(cudaMemsetAsync(t.gX, 0, (size_t)Re * H * sizeof(float), s));          // dL/d(encoder output), summed over the cross-attentions
  float* gkv = t.gBIG;                                                                  // [Re, 2H]
  auto cross_kv_bwd = [&](DecTape& D, const DecLayerW& lw, const FusedAttn& kv) -> int {
    HFT_TRY(gemm_dw(s, gkv, 2 * H, t.x_enc, H, gof(m, G, lw.ca.k_w), H, gof(m, G, lw.ca.k_b), Re, H, H));
    HFT_TRY(gemm_dw(s, gkv + H, 2 * H, t.x_enc, H, gof(m, G, lw.ca.v_w), H, gof(m, G, lw.ca.v_b), Re, H, H));
    HFT_TRY(gemm_nn(s, gkv, 2 * H, kv.qkv_w, H, t.gX, H, Re, H, 2 * H, true));
    (void)D;
    return HFT_OK;
  };
  for (int l = (int)m->dec.size() - 1; l >= 0; --l) {
    DecTape& D = t.dec[l + 1];
    const DecLayerW& lw = m->dec[l];
    const int bl = t.site_dec(l);
    HFT_TRY(dec_ffn_bwd(t, s, D, lw, Rd, t.gT, G, bl + 4, bl + 5));
    ln_bwd(m, s, t.gT, D.s1, lw.ln, Rd, t.gT, G);
    const float* gb = branch_grad(s, t.gT, t.gM, Rd * H, t.drop(bl + 3));
    HFT_TRY(gemm_dw(s, gb, H, D.ctx_c, H, gof(m, G, lw.ca.o_w), H, gof(m, G, lw.ca.o_b), Rd, H, H));
    HFT_TRY(gemm_nn(s, gb, H, m->w[lw.ca.o_w], H, t.gCTX, H, Rd, H, H, false));
    HFT_TRY(attn_bwd(m, s, D.qc, H, (long long)NN * H, D.kv, D.kv + H, 2 * H, t.gCTX, D.ctx_c, D.lse_c, Se, NN, NB, t.gDQ, H, gkv, gkv + H, 2 * H, t.dD, t.drop(bl + 2)));
    HFT_TRY(gemm_dw(s, t.gDQ, H, D.t0, H, gof(m, G, lw.ca.q_w), H, gof(m, G, lw.ca.q_b), Rd, H, H));
    HFT_TRY(gemm_nn(s, t.gDQ, H, m->w[lw.ca.q_w], H, t.gT, H, Rd, H, H, true));
    HFT_TRY(cross_kv_bwd(D, lw, m->dec_ca_kv[l + 1]));
    ln_bwd(m, s, t.gT, D.s0, lw.ln, Rd, t.gT, G);
    gb = branch_grad(s, t.gT, t.gM, Rd * H, t.drop(bl + 1));
    HFT_TRY(gemm_dw(s, gb, H, D.ctx_s, H, gof(m, G, lw.sa.o_w), H, gof(m, G, lw.sa.o_b), Rd, H, H));
    HFT_TRY(gemm_nn(s, gb, H, m->w[lw.sa.o_w], H, t.gCTX, H, Rd, H, H, false));
    HFT_TRY(attn_bwd(m, s, D.qkv, 3 * H, (long long)NN * 3 * H, D.qkv + H, D.qkv + 2 * H, 3 * H, t.gCTX, D.ctx_s, D.lse_s, Se, NN, NN, t.gDQ, 3 * H, t.gDQ + H,
                     t.gDQ + 2 * H, 3 * H, t.dD, t.drop(bl)));
    HFT_TRY(gemm_dw(s, t.gDQ, 3 * H, D.tin, H, gof(m, G, lw.sa.q_w), H, gof(m, G, lw.sa.q_b), Rd, H, H));
    HFT_TRY(gemm_dw(s, t.gDQ + H, 3 * H, D.tin, H, gof(m, G, lw.sa.k_w), H, gof(m, G, lw.sa.k_b), Rd, H, H));
    HFT_TRY(gemm_dw(s, t.gDQ + 2 * H, 3 * H, D.tin, H, gof(m, G, lw.sa.v_w), H, gof(m, G, lw.sa.v_b), Rd, H, H));
    HFT_TRY(gemm_nn(s, t.gDQ, 3 * H, m->dec_sa_qkv[l].qkv_w, H, t.gT, H, Rd, H, 3 * H, true));
  }
  {                                                                                     // layer zero
    DecTape& D = t.dec[0];
    const DecLayerW& lw = m->dec0;
    float* g_pos = gof(m, G, m->dec_pos_freq);
    const int b0 = t.site_dec0();
    HFT_TRY(dec_ffn_bwd(t, s, D, lw, Rd, t.gT, G, b0 + 2, b0 + 3));
    ln_bwd(m, s, t.gT, D.s1, lw.ln, Rd, t.gT, G);                                       // s1 = pos_embedding_freq + dropout(fc_o(ctx))
    colsum(s, t.gT, Se, (long long)NN * H, g_pos);
    const float* gb = branch_grad(s, t.gT, t.gM, Rd * H, t.drop(b0 + 1));
    HFT_TRY(gemm_dw(s, gb, H, D.ctx_c, H, gof(m, G, lw.ca.o_w), H, gof(m, G, lw.ca.o_b), Rd, H, H));
    HFT_TRY(gemm_nn(s, gb, H, m->w[lw.ca.o_w], H, t.gCTX, H, Rd, H, H, false));
    HFT_TRY(attn_bwd(m, s, m->q0, H, 0, D.kv, D.kv + H, 2 * H, t.gCTX, D.ctx_c, D.lse_c, Se, NN, NB, t.gDQ, H, gkv, gkv + H, 2 * H, t.dD, t.drop(b0)));
    HFT_CHECK_CUDA(cudaMemsetAsync(t.gQ0, 0, (size_t)NN * H * sizeof(float), s));
    colsum(s, t.gDQ, Se, (long long)NN * H, t.gQ0);                                     // the same projected queries serve every sequence
    HFT_TRY(gemm_dw(s, t.gQ0, H, m->w[m->dec_pos_freq], H, gof(m, G, lw.ca.q_w), H, gof(m, G, lw.ca.q_b), NN, H, H));
    HFT_TRY(gemm_nn(s, t.gQ0, H, m->w[lw.ca.q_w], H, g_pos, H, NN, H, H, true));
    HFT_TRY(cross_kv_bwd(D, lw, m->dec_ca_kv[0]));
  }
  // ---- encoder ----
  for (int l = (int)m->enc.size() - 1; l >= 0; --l)
    HFT_TRY(enc_layer_bwd(t, s, t.enc[l], Se, NB, m->enc[l], m->enc_qkv[l], t.gX, t.gBIG, t.gCTX, t.gHID, G, t.site_enc(l)));
  dropout_inplace(s, t.gX, Re * H, t.drop(0));                                           // back through the embedding dropout
  colsum(s, t.gX, Se, (long long)NB * H, gof(m, G, m->pos_freq));                       // pos_embedding_freq (encoder)
  // ---- front: conv + Linear through the collapsed 65-tap filter ----
  HFT_CHECK_CUDA(cudaMemsetAsync(t.g_front_w, 0, (size_t)H * m->nproc * sizeof(float), s));
  HFT_CHECK_CUDA(cudaMemsetAsync(t.g_front_b, 0, (size_t)H * sizeof(float), s));
  {
    LaunchScope ls(HFT_KCLASS_FRONT, s);
    static int spc = -1;                               // segments per CTA (HFT_FRONT_BWD_SPC; see the kernel: parallelism against same-address atomics)
    if (spc < 0) { const char* e = getenv("HFT_FRONT_BWD_SPC"); spc = e ? atoi(e) : 2; if (spc < 1) spc = 1; }
    const dim3 fgrid(NB, (B + spc - 1) / spc);
    if (H == 64) front_bwd_kernel<65, 4><<<fgrid, 256, 0, s>>>(spec, sb, sbin, st, t.gX, sqrtH, H, F, NB, B, spc, t.g_front_w, t.g_front_b);
    else if (H == 128) front_bwd_kernel<65, 2><<<fgrid, 256, 0, s>>>(spec, sb, sbin, st, t.gX, sqrtH, H, F, NB, B, spc, t.g_front_w, t.g_front_b);
    else front_bwd_kernel<65, 1><<<fgrid, 256, 0, s>>>(spec, sb, sbin, st, t.gX, sqrtH, H, F, NB, B, spc, t.g_front_w, t.g_front_b);
    const int C = m->d.cnn_channel, kw = m->d.cnn_kernel, n_out = m->nproc - (kw - 1);
    const int total = H * C * n_out + H + 32 * (C * kw + C);                     // threads, then one warp per conv_w / conv_b element
    front_chain_bwd_kernel<<<(total + 127) / 128, 128, 0, s>>>(t.g_front_w, t.g_front_b, m->w[m->tok_w], m->w[m->conv_w], m->w[m->conv_b], H, C, kw, n_out, m->nproc,
                                                                gof(m, G, m->tok_w), gof(m, G, m->tok_b), gof(m, G, m->conv_w), gof(m, G, m->conv_b));
  }
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}

int derive_weights_public(Model* m, cudaStream_t s);      // model.cu

}  // namespace hft

using namespace hft;

extern "C" int hft_trainer_create(hft_trainer** out, hft_model* model, int32_t batch) {
  HFT_REQUIRE(out && model && batch >= 1, HFT_ERR_ARG, "hft_trainer_create: bad argument");
  Model* m = reinterpret_cast<Model*>(model);
  HFT_REQUIRE(m->weights_set, HFT_ERR_STATE, "hft_trainer_create: call hft_model_set_weights first");
  HFT_REQUIRE(m->nvel == 128 && m->nproc == 65, HFT_ERR_UNSUPPORTED, "hft_trainer_create: built for 128 velocities and margin 32");
  Trainer* t = new Trainer();
  t->m = m;
  t->B = batch;
  int rc = alloc_tape(*t);
  if (rc != HFT_OK) { cudaFree(t->arena); delete t; return rc; }
  *out = reinterpret_cast<hft_trainer*>(t);
  return HFT_OK;
}

extern "C" int hft_trainer_destroy(hft_trainer* trainer) {
  if (!trainer) return HFT_OK;
  Trainer* t = reinterpret_cast<Trainer*>(trainer);
  cudaFree(t->arena);
  delete t;
  return HFT_OK;
}

extern "C" int hft_trainer_set_dropout(hft_trainer* trainer, float p, uint32_t seed) {
  HFT_REQUIRE(trainer && p >= 0.f && p < 1.f, HFT_ERR_ARG, "hft_trainer_set_dropout: p must be in [0, 1)");
  Trainer* t = reinterpret_cast<Trainer*>(trainer);
  t->p_drop = p;
  t->seed = seed;
  return HFT_OK;
}

extern "C" int hft_dropout_mask(float p, uint32_t seed, int32_t site, int64_t n, float* mask_out_dev, void* stream) {
  HFT_REQUIRE(mask_out_dev && n >= 0 && p >= 0.f && p < 1.f, HFT_ERR_ARG, "hft_dropout_mask: bad argument");
  if (n == 0) return HFT_OK;
  Drop d{0u, seed, (uint32_t)site, 1.f};
  if (p > 0.f) { d.thresh = (uint32_t)((double)p * 4294967296.0); d.scale = 1.f / (1.f - p); }
  dropout_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mask_out_dev, n, d);
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}

extern "C" int64_t hft_model_param_floats(const hft_model* model) {
  const Model* m = reinterpret_cast<const Model*>(model);
  if (!m) return -1;
  int64_t total = 0;
  for (auto& w : m->spec) total += (w.numel + 3) & ~3ll;
  return total;
}
extern "C" int64_t hft_model_param_offset(const hft_model* model, int index) {
  const Model* m = reinterpret_cast<const Model*>(model);
  if (!m || index < 0 || index >= (int)m->spec.size()) return -1;
  int64_t off = 0;
  for (int i = 0; i < index; ++i) off += (m->spec[i].numel + 3) & ~3ll;
  return off;
}
extern "C" float* hft_model_params(hft_model* model) { return model ? reinterpret_cast<Model*>(model)->arena : nullptr; }

extern "C" int hft_model_get_params(hft_model* model, float* params_out_dev, void* stream) {
  HFT_REQUIRE(model && params_out_dev, HFT_ERR_ARG, "hft_model_get_params: NULL argument");
  Model* m = reinterpret_cast<Model*>(model);
  HFT_REQUIRE(m->weights_set, HFT_ERR_STATE, "hft_model_get_params: no weights registered");
  HFT_CHECK_CUDA(cudaMemcpyAsync(params_out_dev, m->arena, (size_t)hft_model_param_floats(model) * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return HFT_OK;
}

extern "C" int hft_model_refresh(hft_model* model, void* stream) {
  HFT_REQUIRE(model, HFT_ERR_ARG, "hft_model_refresh: NULL model");
  Model* m = reinterpret_cast<Model*>(model);
  HFT_REQUIRE(m->weights_set, HFT_ERR_STATE, "hft_model_refresh: no weights registered");
  HFT_TRY(derive_weights_public(m, (cudaStream_t)stream));
  return tc_prepare_weights(m, (cudaStream_t)stream);
}

extern "C" int hft_train_forward_backward(hft_trainer* trainer, const float* spec_dev, int64_t stride_b, int64_t stride_bin, int64_t stride_t,
                                          const float* label_onset_dev, const float* label_offset_dev, const float* label_mpe_dev,
                                          const int64_t* label_velocity_dev, float weight_A, float weight_B, float* loss_dev, float* grads_dev, void* stream) {
  HFT_REQUIRE(trainer && spec_dev && label_onset_dev && label_offset_dev && label_mpe_dev && label_velocity_dev && loss_dev && grads_dev, HFT_ERR_ARG,
              "hft_train_forward_backward: NULL argument");
  Trainer* t = reinterpret_cast<Trainer*>(trainer);
  return hft_train_forward_backward_n(trainer, t->B, spec_dev, stride_b, stride_bin, stride_t, label_onset_dev, label_offset_dev, label_mpe_dev, label_velocity_dev,
                                      weight_A, weight_B, loss_dev, grads_dev, stream);
}

extern "C" int hft_train_forward_backward_n(hft_trainer* trainer, int32_t batch, const float* spec_dev, int64_t stride_b, int64_t stride_bin, int64_t stride_t,
                                            const float* label_onset_dev, const float* label_offset_dev, const float* label_mpe_dev,
                                            const int64_t* label_velocity_dev, float weight_A, float weight_B, float* loss_dev, float* grads_dev, void* stream) {
  HFT_REQUIRE(trainer && spec_dev && label_onset_dev && label_offset_dev && label_mpe_dev && label_velocity_dev && loss_dev && grads_dev, HFT_ERR_ARG,
              "hft_train_forward_backward: NULL argument");
  Trainer* t = reinterpret_cast<Trainer*>(trainer);
  HFT_REQUIRE(batch >= 1 && batch <= t->B, HFT_ERR_ARG, "hft_train_forward_backward: batch %d outside [1, %d] (the capacity given to hft_trainer_create)", batch, t->B);
  t->Bc = batch;
  cudaStream_t s = (cudaStream_t)stream;
  reset_launch_count();
  HFT_CHECK_CUDA(cudaMemsetAsync(grads_dev, 0, (size_t)hft_model_param_floats(reinterpret_cast<hft_model*>(t->m)) * sizeof(float), s));
  HFT_TRY(train_forward(*t, spec_dev, stride_b, stride_bin, stride_t, s));
  GradSource gs;
  gs.y_on = label_onset_dev; gs.y_off = label_offset_dev; gs.y_mpe = label_mpe_dev; gs.y_vel = reinterpret_cast<const long long*>(label_velocity_dev);
  gs.wA = weight_A; gs.wB = weight_B; gs.loss = loss_dev;
  return train_backward(*t, spec_dev, stride_b, stride_bin, stride_t, gs, grads_dev, s);
}

// Train-mode forward on its own (model(input_spec) of train.py:90 with model.train()): fills the tape and writes the eight head outputs
// (sigmoid probabilities / raw velocity logits, the layouts of hft_outputs; attention and the argmax members are ignored).
extern "C" int hft_train_forward(hft_trainer* trainer, int32_t batch, const float* spec_dev, int64_t stride_b, int64_t stride_bin, int64_t stride_t,
                                 const hft_outputs* outputs, void* stream) {
  HFT_REQUIRE(trainer && spec_dev && outputs, HFT_ERR_ARG, "hft_train_forward: NULL argument");
  Trainer* t = reinterpret_cast<Trainer*>(trainer);
  HFT_REQUIRE(batch >= 1 && batch <= t->B, HFT_ERR_ARG, "hft_train_forward: batch %d outside [1, %d]", batch, t->B);
  t->Bc = batch;
  cudaStream_t s = (cudaStream_t)stream;
  reset_launch_count();
  HFT_TRY(train_forward(*t, spec_dev, stride_b, stride_bin, stride_t, s));
  Model* m = t->m;
  const long long Rd = (long long)batch * m->nframe * m->nnote;
  {
    LaunchScope ls(HFT_KCLASS_HEADS, s);
    heads_out_kernel<<<(unsigned)((Rd + 7) / 8), 256, 0, s>>>(t->logits_a, t->NP, m->nvel, m->nframe, m->nnote, Rd, false, outputs->onset_A, outputs->offset_A, outputs->mpe_A,
                                                              outputs->velocity_A);
    heads_out_kernel<<<(unsigned)((Rd + 7) / 8), 256, 0, s>>>(t->logits_b, t->NP, m->nvel, m->nframe, m->nnote, Rd, true, outputs->onset_B, outputs->offset_B, outputs->mpe_B,
                                                              outputs->velocity_B);
  }
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}

// loss.backward() of train.py:157 for a loss the host built from the outputs of the LAST hft_train_forward on this trainer: out_grads holds
// dLoss/d(output) in the output layouts (NULL member = zero); grads_dev is overwritten with dLoss/dParam.  spec_dev must be the same input.
extern "C" int hft_train_backward(hft_trainer* trainer, const float* spec_dev, int64_t stride_b, int64_t stride_bin, int64_t stride_t, const hft_outputs* out_grads,
                                  float* grads_dev, void* stream) {
  HFT_REQUIRE(trainer && spec_dev && out_grads && grads_dev, HFT_ERR_ARG, "hft_train_backward: NULL argument");
  Trainer* t = reinterpret_cast<Trainer*>(trainer);
  HFT_REQUIRE(t->Bc >= 1, HFT_ERR_STATE, "hft_train_backward: no forward on this trainer yet");
  cudaStream_t s = (cudaStream_t)stream;
  reset_launch_count();
  HFT_CHECK_CUDA(cudaMemsetAsync(grads_dev, 0, (size_t)hft_model_param_floats(reinterpret_cast<hft_model*>(t->m)) * sizeof(float), s));
  GradSource gs;
  gs.g = out_grads;
  return train_backward(*t, spec_dev, stride_b, stride_bin, stride_t, gs, grads_dev, s);
}

extern "C" int hft_adam_step(float* params_dev, const float* grads_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t n, float lr, float beta1, float beta2,
                             float eps, int64_t step, float grad_scale, void* stream) {
  HFT_REQUIRE(params_dev && grads_dev && exp_avg_dev && exp_avg_sq_dev && n >= 0 && step >= 1, HFT_ERR_ARG, "hft_adam_step: bad argument");
  if (n == 0) return HFT_OK;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = sqrtf(1.f - powf(beta2, (float)step));
  reset_launch_count();
  {
    LaunchScope ls(HFT_KCLASS_NORM, stream);
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params_dev, grads_dev, exp_avg_dev, exp_avg_sq_dev, n, lr, beta1, beta2, eps, bc1, bc2, grad_scale);
  }
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}

// Component entry: the multi-head attention of the training step alone (forward with the row log-sum-exp, optionally the backward),
// on caller-provided fp32 tensors.  use_tc: 1 = tcgen05 kernels (tc_train_attn.cuh; head_dim 32, Lq, Lk <= 256), 0 = fp32 CUDA-core kernels,
// -1 = what the training step itself would pick.  Lets the tests compare either implementation with torch autograd on the same operands.
extern "C" int hft_train_attention(int32_t use_tc, int32_t dh, int32_t heads, const float* q_dev, int32_t ldq, int64_t q_seq_stride, const float* k_dev,
                                   const float* v_dev, int32_t ldkv, int64_t n_seq, int32_t lq, int32_t lk, float p_drop, uint32_t seed, int32_t site,
                                   float* ctx_dev, float* lse_dev, const float* d_ctx_dev, float* dq_dev, int32_t lddq, float* dk_dev, float* dv_dev,
                                   int32_t lddkv, float* d_buf_dev, void* stream) {
  HFT_REQUIRE(q_dev && k_dev && v_dev && ctx_dev && lse_dev && n_seq >= 1 && lq >= 1 && lk >= 1 && heads >= 1 && (dh == 32 || dh == 64), HFT_ERR_ARG,
              "hft_train_attention: bad argument");
  HFT_REQUIRE(p_drop >= 0.f && p_drop < 1.f, HFT_ERR_ARG, "hft_train_attention: p_drop must be in [0, 1)");
  HFT_REQUIRE(!d_ctx_dev || (dq_dev && dk_dev && dv_dev && d_buf_dev), HFT_ERR_ARG, "hft_train_attention: backward needs dq, dk, dv and d_buf");
  cudaStream_t s = (cudaStream_t)stream;
  const AttnDims d{dh, heads, heads * dh};
  Drop drop{0u, seed, (uint32_t)site, 1.f};
  if (p_drop > 0.f) { drop.thresh = (uint32_t)((double)p_drop * 4294967296.0); drop.scale = 1.f / (1.f - p_drop); }
  reset_launch_count();
  g_train_tc_force = use_tc;
  if (use_tc == 1 && !tattn_ok(d, ldq, ldkv, lq, lk)) {
    g_train_tc_force = -1;
    HFT_REQUIRE(false, HFT_ERR_UNSUPPORTED, "hft_train_attention: the tcgen05 kernels need head_dim 32, Lq, Lk <= 256 and row pitches that are multiples of 4");
  }
  int rc = attn_fwd(d, s, q_dev, ldq, q_seq_stride, k_dev, v_dev, ldkv, n_seq, lq, lk, ctx_dev, lse_dev, drop);
  if (rc == HFT_OK && d_ctx_dev)
    rc = attn_bwd(d, s, q_dev, ldq, q_seq_stride, k_dev, v_dev, ldkv, d_ctx_dev, ctx_dev, lse_dev, n_seq, lq, lk, dq_dev, lddq, dk_dev, dv_dev, lddkv, d_buf_dev, drop);
  g_train_tc_force = -1;
  if (rc != HFT_OK) return rc;
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}

// Component entry: one Linear of the training step on fp32 device tensors -- forward y = x W^T + b (w_kn = 0, W [N, K]; optional ReLU) or
// input gradient dx = dy W (w_kn = 1, W [K, N]; optional ReLU mask: c = mask > 0 ? c * mask_scale : 0) -- optionally accumulated into C.
// use_tc as in hft_train_attention (tcgen05 kernel: K = 64 / 128, N % 16 == 0, N <= 256).
extern "C" int hft_train_linear(int32_t use_tc, int32_t w_kn, const float* a_dev, int32_t lda, const float* w_dev, int32_t ldw, const float* bias_dev,
                                float* c_dev, int32_t ldc, int64_t m, int32_t n, int32_t k, int32_t relu, int32_t accum, const float* mask_dev, int32_t ldm,
                                float mask_scale, void* stream) {
  HFT_REQUIRE(a_dev && w_dev && c_dev && m >= 1 && n >= 1 && k >= 1, HFT_ERR_ARG, "hft_train_linear: bad argument");
  HFT_REQUIRE(!(w_kn && (bias_dev || relu)) && !(!w_kn && mask_dev), HFT_ERR_ARG, "hft_train_linear: bias / ReLU belong to the forward form, the mask to the gradient form");
  cudaStream_t s = (cudaStream_t)stream;
  reset_launch_count();
  g_train_tc_force = use_tc;
  int rc;
  if (use_tc == 1 && !tgemm_ok(a_dev, lda, w_dev, ldw, w_kn != 0, c_dev, ldc, n, k, mask_dev, ldm)) {
    set_error("hft_train_linear: the tcgen05 kernel needs K = 64 or 128, N %% 16 == 0, N <= 256, 16-byte aligned rows");
    rc = HFT_ERR_UNSUPPORTED;
  } else if (w_kn) {
    rc = gemm_nn(s, a_dev, lda, w_dev, ldw, c_dev, ldc, m, n, k, accum != 0, mask_dev, ldm, mask_scale);
  } else {
    rc = gemm_tn(s, a_dev, lda, w_dev, ldw, bias_dev, c_dev, ldc, m, n, k, relu != 0, accum != 0);
  }
  g_train_tc_force = -1;
  if (rc != HFT_OK) return rc;
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}

// Component entry: the weight gradient of one Linear of the training step, dW[n, k] += dY[m, n]^T X[m, k] and db[n] += colsum(dY) (db may be NULL),
// on fp32 device tensors.  use_tc as in hft_train_attention (tcgen05 kernel: k = 64 / 128, n % 8 == 0, n + k <= 256).
extern "C" int hft_train_linear_wgrad(int32_t use_tc, const float* dy_dev, int32_t ldy, const float* x_dev, int32_t ldx, float* dw_dev, int32_t ldw, float* db_dev,
                                      int64_t m, int32_t n, int32_t k, void* stream) {
  HFT_REQUIRE(dy_dev && x_dev && dw_dev && m >= 1 && n >= 1 && k >= 1, HFT_ERR_ARG, "hft_train_linear_wgrad: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  reset_launch_count();
  g_train_tc_force = use_tc;
  int rc;
  if (use_tc == 1 && !tdw_ok(dy_dev, ldy, x_dev, ldx, n, k)) {
    set_error("hft_train_linear_wgrad: the tcgen05 kernel needs K = 64 or 128, N %% 8 == 0, N + K <= 256, 16-byte aligned rows");
    rc = HFT_ERR_UNSUPPORTED;
  } else {
    rc = gemm_dw(s, dy_dev, ldy, x_dev, ldx, dw_dev, ldw, db_dev, m, n, k);
  }
  g_train_tc_force = -1;
  if (rc != HFT_OK) return rc;
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}
