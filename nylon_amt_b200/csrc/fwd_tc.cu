// Tensor-core forward of Model_SPEC2MIDI (reference hftt_code/model/model_spec2midi.py:15-35) -- HFT_PREC_F16X3 (split fp16
// operands, three products: the fp32-parity mode and the bench default), HFT_PREC_BF16 / HFT_PREC_F16 (one product): 16-bit
// operands on tcgen05 with fp32 accumulation in TMEM; softmax, LayerNorm, residuals, sigmoid in fp32.  Activations live in
// HBM as 16-bit row-major [rows, H] tensors (split mode: [rows, 2H] = hi | lo).  Launch sequence per chunk of segments:
//   front_tc_kernel (tc_front.cuh)            collapsed 65-tap filter + positional row -> X
//   per layer: gemm_kernel<EPI_STORE> (QKV)   -> attn2_kernel / attn_probs_kernel (tc_attn2.cuh) -> gemm_kernel<EPI_LN> (fc_o +
//              residual + LayerNorm)          -> ffn_kernel (tc_ffn.cuh: fc_1, ReLU, fc_2, residual, LayerNorm in one launch)
//   gemm_kernel<EPI_HEADS>                    the four heads of a group as one [rows, 192] GEMM, sigmoid / argmax in the epilogue
// Host side: tensor maps are encoded once per (workspace, weights) and passed as __grid_constant__.
#include "common.cuh"
#include "model.h"
#include "tc_attn.cuh"
#include "tc_attn2.cuh"
#include "tc_gemm.cuh"
#include "tc_ffn.cuh"
#include "tc_front.cuh"

#include <cudaTypedefs.h>
#include <math.h>
#include <stdlib.h>
#include <map>

namespace hft {

using namespace tc;

// ---- cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda) ---------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// 2-D row-major 16-bit tensor [rows, cols] with row pitch ld (elements); box = box_cols x box_rows; swizzle = box_cols*2 bytes.
static int make_map(CUtensorMap* m, const void* base, long long rows, long long cols, long long ld, int box_cols, int box_rows, bool bf16) {
  auto enc = get_encode();
  HFT_REQUIRE(enc != nullptr, HFT_ERR_STATE, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : box_cols * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HFT_REQUIRE(r == CUDA_SUCCESS, HFT_ERR_STATE, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r, rows, cols, ld, box_cols, box_rows);
  return HFT_OK;
}

// 3-D view [outer][mid][cols] of a row-major 16-bit tensor whose rows are (outer * mid_n + mid); box = 64 columns x 1 x 32 outer rows
// (the front kernel's tile: 32 consecutive frames of one bin), 128-byte swizzle like the 2-D store boxes.
static int make_map_front(CUtensorMap* m, const void* base, long long outer, long long mid_n, long long cols, long long ld, bool bf16) {
  auto enc = get_encode();
  HFT_REQUIRE(enc != nullptr, HFT_ERR_STATE, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)mid_n, (cuuint64_t)outer};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)mid_n};
  cuuint32_t box[3] = {64, 1, 32};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HFT_REQUIRE(r == CUDA_SUCCESS, HFT_ERR_STATE, "cuTensorMapEncodeTiled (3-D front map) failed (%d)", (int)r);
  return HFT_OK;
}

// Rows per TMA-store box of the GEMM epilogues (HFT_TC_STAGE_ROWS = 32 | 16 | 8).  A warp's staging block is rows x 128 bytes, so
// 16 / 8 rows free 16 / 24 KB of shared memory, which in split mode is what lets the CTA's W slice (128 KB, K = 256) stay resident
// next to a 4- / 5-deep A ring instead of streaming it from L2 for every row tile.
static int gemm_stage_rows() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HFT_TC_STAGE_ROWS"); v = e ? atoi(e) : 32; if (v != 8 && v != 16) v = 32; }
  return v;
}

// ---- small CUDA-core kernels of the 16-bit path -----------------------------------------------------------------
// In x3 (split) mode every 16-bit tensor carries two column blocks: hi = round16(v) at [0, C) and lo = round16(v - hi)
// at [C, 2C); hi + lo reproduces v to ~22 bits with fp16 parts.
template <bool BF16>
__device__ __forceinline__ void put16(uint16_t* dst, long long idx, long long lo_off, bool x3, float v) {
  const uint32_t h = Op16<BF16>::pack(v, 0.f);
  dst[idx] = (uint16_t)(h & 0xffffu);
  if (x3) dst[idx + lo_off] = (uint16_t)(Op16<BF16>::pack(v - Op16<BF16>::lo(h), 0.f) & 0xffffu);
}

// dst[r, :] = 16-bit(src[r % period, :]) for r < valid_rows, 0 beyond   (weights: period = rows; pitch queries: 88)
template <bool BF16>
__global__ void cvt_rows_kernel(const float* __restrict__ src, int period, int valid_rows, int K, int rows, bool x3, uint16_t* __restrict__ dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * K) return;
  int r = (int)(i / K), k = (int)(i % K);
  const int ld = x3 ? 2 * K : K;
  put16<BF16>(dst, (long long)r * ld + k, K, x3, r < valid_rows ? src[(long long)(r % period) * K + k] : 0.f);
}

// heads weight [192, H]: rows 0..V-1 velocity, V..V+2 onset/offset/mpe, rest zero (so velocity stores are 16-byte aligned)
template <bool BF16>
__global__ void pack_heads_kernel(const float* __restrict__ w_on, const float* __restrict__ w_off, const float* __restrict__ w_mpe,
                                  const float* __restrict__ w_vel, const float* __restrict__ b_on, const float* __restrict__ b_off,
                                  const float* __restrict__ b_mpe, const float* __restrict__ b_vel, int V, int H, int rows, bool x3,
                                  uint16_t* __restrict__ w16, float* __restrict__ bias) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H) return;
  int r = i / H, h = i % H;
  float v = 0.f, b = 0.f;
  if (r < V) { v = w_vel[r * H + h]; b = b_vel[r]; }
  else if (r == V) { v = w_on[h]; b = b_on[0]; }
  else if (r == V + 1) { v = w_off[h]; b = b_off[0]; }
  else if (r == V + 2) { v = w_mpe[h]; b = b_mpe[0]; }
  put16<BF16>(w16, (long long)r * (x3 ? 2 * H : H) + h, H, x3, v);
  if (h == 0) bias[r] = b;
}

// front: collapsed 65-tap filter per hidden unit, fp32 FMA, 16-bit store.  grid (n_bin, B).
template <bool BF16, int NPROC>
__global__ void __launch_bounds__(256) front16_kernel(const float* __restrict__ spec, long long sb, long long sbin, long long st,
                                                      const float* __restrict__ Wc, const float* __restrict__ bc, const float* __restrict__ pos,
                                                      float scale, int H, int F, int NB, bool x3, uint16_t* __restrict__ X) {
  __shared__ float s_row[256];
  const int bin = blockIdx.x, b = blockIdx.y;
  const int W = F + NPROC - 1;
  for (int i = threadIdx.x; i < W; i += blockDim.x) s_row[i] = spec[b * sb + bin * sbin + i * st];
  __syncthreads();
  const int groups = blockDim.x / H;
  const int h = threadIdx.x % H, g = threadIdx.x / H;
  if (g >= groups) return;
  float w[NPROC];
#pragma unroll
  for (int j = 0; j < NPROC; ++j) w[j] = Wc[h * NPROC + j];
  const float bias = bc[h], pe = pos[bin * H + h];
  const int fpg = F / groups;
  const int ld = x3 ? 2 * H : H;
  for (int f0 = g * fpg; f0 < (g + 1) * fpg; f0 += 8) {
    float acc[8], sv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = 0.f; sv[i] = s_row[f0 + i]; }
#pragma unroll
    for (int j = 0; j < NPROC; ++j) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(w[j], sv[i], acc[i]);
#pragma unroll
      for (int i = 0; i < 7; ++i) sv[i] = sv[i + 1];
      sv[7] = (j + 1 < NPROC) ? s_row[f0 + j + 8] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) put16<BF16>(X, (((long long)b * F + f0 + i) * NB + bin) * ld + h, H, x3, (acc[i] + bias) * scale + pe);
  }
}

// U[((b*NN+n)*F+f), h] = T[((b*F+f)*NN+n), h] * sqrt(H) + pos_time[f, h]   (model_spec2midi.py:189-191); 8 elements / thread
template <bool BF16>
__global__ void time_relayout16_kernel(const uint16_t* __restrict__ T, const float* __restrict__ pos, float scale, int F, int NN, int H, bool x3,
                                       long long total8, uint16_t* __restrict__ U) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int h8 = H / 8;
  const int ld = x3 ? 2 * H : H;
  int hq = (int)(i % h8);
  long long r = i / h8;
  int f = (int)(r % F);
  long long bn = r / F;
  int n = (int)(bn % NN);
  long long b = bn / NN;
  const uint16_t* src = T + (((b * F + f) * NN + n)) * ld + hq * 8;
  uint4 q = *reinterpret_cast<const uint4*>(src);
  uint4 ql = x3 ? *reinterpret_cast<const uint4*>(src + H) : make_uint4(0, 0, 0, 0);
  const float* pp = pos + f * H + hq * 8;
  uint32_t w4[4] = {q.x, q.y, q.z, q.w}, l4[4] = {ql.x, ql.y, ql.z, ql.w}, o[4], ol[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float a = Op16<BF16>::lo(w4[e]), c = Op16<BF16>::hi(w4[e]);
    if (x3) { a += Op16<BF16>::lo(l4[e]); c += Op16<BF16>::hi(l4[e]); }
    a = a * scale + pp[2 * e];
    c = c * scale + pp[2 * e + 1];
    o[e] = Op16<BF16>::pack(a, c);
    ol[e] = Op16<BF16>::pack(a - Op16<BF16>::lo(o[e]), c - Op16<BF16>::hi(o[e]));
  }
  uint16_t* dst = U + r * ld + hq * 8;
  *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
  if (x3) *reinterpret_cast<uint4*>(dst + H) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
}

// ---- per-precision state -----------------------------------------------------------------------------------------
struct W16 {            // one weight matrix [N, K] (x3: [N, 2K] = hi | lo) in 16-bit + its tensor map (box = 64 x n_tile)
  uint16_t* ptr = nullptr;
  int N = 0, K = 0, n_tile = 0;
  CUtensorMap map;                  // box 64 x n_tile
  CUtensorMap map2;                 // box 64 x n_tile/2: one CTA's half of the W rows in PAIR (cta_group::2) mode
  CUtensorMap map64;                // box 64 x 64: fc_1 rows of one CTA for one 128-unit hidden block (fused FFN)
  const float* bias = nullptr;
};

struct TcLayer { W16 qkv, o, w1, w2; };                    // EncoderLayer
struct TcDecLayer { W16 sa_qkv, sa_o, ca_q, ca_kv, ca_o, w1, w2; };

struct TcState {
  bool bf16 = true;
  bool x3 = false;                  // split operands (hi | lo): fp32-class accuracy at 3 MMAs per product
  bool weights_ready = false;
  uint16_t* warena = nullptr;
  float* head_bias = nullptr;       // [2][192]
  uint16_t* q0_16 = nullptr;        // [128, H] projected pitch queries (rows >= n_note zero)
  std::vector<TcLayer> enc, tim;
  TcDecLayer dec0;
  std::vector<TcDecLayer> dec;
  W16 headA, headB;
  // workspace
  int ws_batch = 0;
  uint16_t* ws = nullptr;
  uint16_t *X, *QKV, *CTX, *HID, *T, *DQ, *U;
  CUtensorMap mX, mCTX, mHID, mT, mU;               // GEMM A operands / TMA-store targets, box 64 x 128
  CUtensorMap gX, gHID, gT, gU, gQKV, gDQ;          // the same targets with stage_rows-row boxes (GEMM epilogues, see gemm_stage_rows())
  CUtensorMap sX3;                                  // X as [b * F + f][bin][column] for the front kernel's stores
  CUtensorMap sX, sHID, sT, sU, sQKV, sDQ, sCTX;    // TMA-store targets (projections, attention context), box 64 x 32 (one epilogue warp)
  CUtensorMap mPosRep;                              // repeated pitch-query table (residual of layer zero)
  uint16_t* pos_rep = nullptr;                      // [11*128, H]: pos_embedding_freq[row % 88] (lcm(88,128) = 1408 rows)
  CUtensorMap mQKV_q, mQKV_kv, mDQ_q, mDQ_kv, mQ0;  // attention operands, box dh x {128, Lk}
  int cm() const { return x3 ? 2 : 1; }
  // HFT_PREC_MIXED: the split (hi | lo) state with a per-site plan of one or three products (see SitePlan)
  bool mixed = false;
};

// Products per site in HFT_PREC_MIXED.  Every tensor keeps the hi | lo layout; a "single" site multiplies only the hi halves.
// The plan comes from the per-site sensitivity study on the seeded-weights fixture (tools/precision_plan.py, plan A): with these
// weights the time stack amplifies upstream rounding ~20x, so everything that feeds time layer 0's scores keeps three products; single are
//   * every P V product except encoder layer 0's,   * the decoder's score products (self and cross),
//   * time layers >= 1 entirely (projections, FFN, both attention products),   * both head GEMMs.
// Worst error of the CPU emulation over 5 signal families: 6.2e-3 on velocity B (budget 2e-2), 1.7e-3 on velocity A, <= 1.1e-3 on the probabilities.
struct SitePlan {
  bool lin = false;       // projections + FFN of the layer
  bool scores = false;    // Q K^T
  bool pv = false;        // P V
};
static SitePlan plan_for(const TcState& t, char kind, int layer) {   // kind: 'e' encoder, 'd' decoder (0 = layer zero), 't' time
  SitePlan p;
  if (!t.mixed) return p;
  if (kind == 'e') { p.pv = layer > 0; }
  else if (kind == 'd') { p.scores = true; p.pv = true; }
  else { p.pv = true; if (layer > 0) { p.lin = true; p.scores = true; } }
  return p;
}

struct TcBoth { TcState st[3]; };   // [0] = fp16, [1] = bf16, [2] = fp16 x3

// UMMA N of a projection: LayerNorm epilogues need the whole row in one tile; otherwise the widest tile dividing N
// (measured: in x3 mode a 256-wide streamed W with a deep A ring beats a 128-wide resident W with a 3-deep A ring).
static int n_tile_for(int N, bool x3 = false, bool full_row = false) {
  if (full_row) return (N == 256 || N == 128 || N == 64) ? N : 0;
  const int cands[3] = {256, 128, 64};
  for (int c : cands)
    if (N % c == 0) return c;
  return 0;
}

template <bool BF16>
static void cvt_rows(const float* src, int period, int valid_rows, int K, int rows, bool x3, uint16_t* dst, cudaStream_t s) {
  const long long n = (long long)rows * K;
  cvt_rows_kernel<BF16><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, period, valid_rows, K, rows, x3, dst);
}

static int prepare_weights(Model* m, TcState& t, cudaStream_t s) {
  const int H = m->H, P = m->P, V = m->nvel;
  const long long hh = (long long)H * H, hp = (long long)H * P;
  const bool bf = t.bf16, x3 = t.x3;
  const int cm = t.cm();
  size_t n_self = m->enc.size() + m->tim.size() + m->dec.size();
  size_t n_dec_layers = 1 + m->dec.size();
  size_t elems = (n_self * (3 * hh + hh) + (m->enc.size() + m->tim.size() + n_dec_layers) * 2 * hp + n_dec_layers * (hh + 2 * hh + hh) +
                  2 * (size_t)192 * H + (size_t)128 * H + (size_t)11 * 128 * H) * cm + 16384;
  if (!t.warena) {
    HFT_CHECK_CUDA(cudaMalloc(&t.warena, elems * 2));
    HFT_CHECK_CUDA(cudaMalloc(&t.head_bias, 2 * 192 * sizeof(float)));
  }
  HFT_CHECK_CUDA(cudaMemsetAsync(t.warena, 0, elems * 2, s));
  uint16_t* p = t.warena;
  auto take = [&](size_t n) { uint16_t* r = p; p += (n + 127) & ~(size_t)127; return r; };
  int rc = HFT_OK;
  auto mk = [&](W16& w, const float* src, int N, int K, const float* bias, bool full_row = false) {
    w.ptr = take((size_t)N * K * cm);
    w.N = N; w.K = K; w.n_tile = n_tile_for(N, x3, full_row); w.bias = bias;
    if (bf) cvt_rows<true>(src, N, N, K, N, x3, w.ptr, s); else cvt_rows<false>(src, N, N, K, N, x3, w.ptr, s);
    int r = make_map(&w.map, w.ptr, N, (long long)K * cm, (long long)K * cm, kBlockK, w.n_tile, bf);
    if (r != HFT_OK) rc = r;
    r = make_map(&w.map2, w.ptr, N, (long long)K * cm, (long long)K * cm, kBlockK, w.n_tile / 2, bf);
    if (r != HFT_OK) rc = r;
    r = make_map(&w.map64, w.ptr, N, (long long)K * cm, (long long)K * cm, kBlockK, 64, bf);
    if (r != HFT_OK) rc = r;
  };
  auto mk_enc = [&](TcLayer& L, const EncLayerW& lw, const FusedAttn& f) {
    mk(L.qkv, f.qkv_w, 3 * H, H, f.qkv_b);
    mk(L.o, m->w[lw.sa.o_w], H, H, m->w[lw.sa.o_b], true);
    mk(L.w1, m->w[lw.ff.w1], P, H, m->w[lw.ff.b1]);
    mk(L.w2, m->w[lw.ff.w2], H, P, m->w[lw.ff.b2], true);
  };
  t.enc.resize(m->enc.size()); t.tim.resize(m->tim.size()); t.dec.resize(m->dec.size());
  for (size_t i = 0; i < m->enc.size(); ++i) mk_enc(t.enc[i], m->enc[i], m->enc_qkv[i]);
  for (size_t i = 0; i < m->tim.size(); ++i) mk_enc(t.tim[i], m->tim[i], m->tim_qkv[i]);
  auto mk_dec = [&](TcDecLayer& L, const DecLayerW& lw, const FusedAttn* sa, const FusedAttn& kv) {
    if (sa) {
      mk(L.sa_qkv, sa->qkv_w, 3 * H, H, sa->qkv_b);
      mk(L.sa_o, m->w[lw.sa.o_w], H, H, m->w[lw.sa.o_b], true);
    }
    mk(L.ca_q, m->w[lw.ca.q_w], H, H, m->w[lw.ca.q_b]);
    mk(L.ca_kv, kv.qkv_w, 2 * H, H, kv.qkv_b);
    mk(L.ca_o, m->w[lw.ca.o_w], H, H, m->w[lw.ca.o_b], true);
    mk(L.w1, m->w[lw.ff.w1], P, H, m->w[lw.ff.b1]);
    mk(L.w2, m->w[lw.ff.w2], H, P, m->w[lw.ff.b2], true);
  };
  mk_dec(t.dec0, m->dec0, nullptr, m->dec_ca_kv[0]);
  for (size_t i = 0; i < m->dec.size(); ++i) mk_dec(t.dec[i], m->dec[i], &m->dec_sa_qkv[i], m->dec_ca_kv[i + 1]);
  // heads: packed + reordered
  HFT_REQUIRE(V + 3 <= 192 && V % 32 == 0, HFT_ERR_UNSUPPORTED, "heads packing expects n_velocity %% 32 == 0 and <= 189");
  auto mk_heads = [&](W16& w, const int* idx, int which) -> int {
    w.ptr = take((size_t)192 * H * cm);
    w.N = 192; w.K = H; w.n_tile = 192; w.bias = t.head_bias + which * 192;
    HFT_CHECK_CUDA(cudaMemsetAsync(t.head_bias + which * 192, 0, 192 * sizeof(float), s));
    if (bf) pack_heads_kernel<true><<<(192 * H + 255) / 256, 256, 0, s>>>(m->w[idx[0]], m->w[idx[2]], m->w[idx[4]], m->w[idx[6]], m->w[idx[1]], m->w[idx[3]], m->w[idx[5]], m->w[idx[7]], V, H, 192, x3, w.ptr, t.head_bias + which * 192);
    else pack_heads_kernel<false><<<(192 * H + 255) / 256, 256, 0, s>>>(m->w[idx[0]], m->w[idx[2]], m->w[idx[4]], m->w[idx[6]], m->w[idx[1]], m->w[idx[3]], m->w[idx[5]], m->w[idx[7]], V, H, 192, x3, w.ptr, t.head_bias + which * 192);
    int r = make_map(&w.map, w.ptr, 192, (long long)H * cm, (long long)H * cm, kBlockK, 192, bf);
    if (r != HFT_OK) rc = r;
    r = make_map(&w.map2, w.ptr, 192, (long long)H * cm, (long long)H * cm, kBlockK, 96, bf);
    if (r != HFT_OK) rc = r;
    return (int)HFT_OK;
  };
  mk_heads(t.headA, m->head_freq, 0);
  mk_heads(t.headB, m->head_time, 1);
  // pitch-query table repeated over lcm(88,128) = 1408 rows: residual operand of layer zero's LayerNorm GEMM
  t.pos_rep = take((size_t)11 * 128 * H * cm);
  if (bf) cvt_rows<true>(m->w[m->dec_pos_freq], m->nnote, 11 * 128, H, 11 * 128, x3, t.pos_rep, s);
  else cvt_rows<false>(m->w[m->dec_pos_freq], m->nnote, 11 * 128, H, 11 * 128, x3, t.pos_rep, s);
  // projected pitch queries of layer zero, zero padded to 128 rows
  t.q0_16 = take((size_t)128 * H * cm);
  if (bf) cvt_rows<true>(m->q0, m->nnote, m->nnote, H, 128, x3, t.q0_16, s); else cvt_rows<false>(m->q0, m->nnote, m->nnote, H, 128, x3, t.q0_16, s);
  HFT_CHECK_CUDA(cudaGetLastError());
  if (rc != HFT_OK) return rc;
  t.weights_ready = true;
  t.ws_batch = 0;                    // tensor maps of q0 / pos_rep are rebuilt with the workspace maps
  return HFT_OK;
}

static int ensure_ws(Model* m, TcState& t, int B) {
  if (t.ws && t.ws_batch >= B) return HFT_OK;
  const long long Re = (long long)B * m->nframe * m->nbin, Rd = (long long)B * m->nframe * m->nnote;
  const int H = m->H, P = m->P, dh = m->dh, cm = t.cm();
  size_t elems = ((size_t)Re * (5 * H + P) + (size_t)Rd * 5 * H) * cm + 8192;
  static_assert(sizeof(uint16_t) == 2, "");
  cudaFree(t.ws);
  t.ws = nullptr;
  HFT_CHECK_CUDA(cudaMalloc(&t.ws, elems * 2));
  uint16_t* p = t.ws;
  auto take = [&](size_t n) { uint16_t* r = p; p += (n + 511) & ~(size_t)511; return r; };
  t.X = take((size_t)Re * H * cm); t.QKV = take((size_t)Re * 3 * H * cm); t.CTX = take((size_t)Re * H * cm); t.HID = take((size_t)Re * P * cm);
  t.T = take((size_t)Rd * H * cm); t.DQ = take((size_t)Rd * 3 * H * cm); t.U = take((size_t)Rd * H * cm);
  const bool bf = t.bf16;
  int rc = HFT_OK;
  auto chk = [&](int r) { if (r != HFT_OK) rc = r; };
  const long long h1 = (long long)H * cm, h3 = (long long)3 * H * cm, p1 = (long long)P * cm;
  chk(make_map(&t.mX, t.X, Re, h1, h1, kBlockK, kBlockM, bf));
  chk(make_map(&t.mCTX, t.CTX, Re, h1, h1, kBlockK, kBlockM, bf));
  chk(make_map(&t.mHID, t.HID, Re, p1, p1, kBlockK, kBlockM, bf));
  chk(make_map(&t.mT, t.T, Rd, h1, h1, kBlockK, kBlockM, bf));
  chk(make_map(&t.mU, t.U, Rd, h1, h1, kBlockK, kBlockM, bf));
  chk(make_map(&t.mQKV_q, t.QKV, Re, h3, h3, dh, 128, bf));
  chk(make_map(&t.mQKV_kv, t.QKV, Re, h3, h3, dh, 256, bf));
  chk(make_map(&t.mDQ_q, t.DQ, Rd, h3, h3, dh, 128, bf));
  chk(make_map(&t.mDQ_kv, t.DQ, Rd, h3, h3, dh, 96, bf));
  chk(make_map(&t.mQ0, t.q0_16, 128, h1, h1, dh, 128, bf));
  chk(make_map(&t.sX, t.X, Re, h1, h1, 64, 32, bf));
  chk(make_map_front(&t.sX3, t.X, Re / m->nbin, m->nbin, h1, h1, bf));
  chk(make_map(&t.sHID, t.HID, Re, p1, p1, 64, 32, bf));
  chk(make_map(&t.sT, t.T, Rd, h1, h1, 64, 32, bf));
  chk(make_map(&t.sU, t.U, Rd, h1, h1, 64, 32, bf));
  chk(make_map(&t.sQKV, t.QKV, Re, h3, h3, 64, 32, bf));
  chk(make_map(&t.sDQ, t.DQ, Rd, h3, h3, 64, 32, bf));
  chk(make_map(&t.sCTX, t.CTX, Re, h1, h1, 64, 32, bf));
  {
    const int sr = gemm_stage_rows();
    chk(make_map(&t.gX, t.X, Re, h1, h1, 64, sr, bf));
    chk(make_map(&t.gHID, t.HID, Re, p1, p1, 64, sr, bf));
    chk(make_map(&t.gT, t.T, Rd, h1, h1, 64, sr, bf));
    chk(make_map(&t.gU, t.U, Rd, h1, h1, 64, sr, bf));
    chk(make_map(&t.gQKV, t.QKV, Re, h3, h3, 64, sr, bf));
    chk(make_map(&t.gDQ, t.DQ, Rd, h3, h3, 64, sr, bf));
  }
  chk(make_map(&t.mPosRep, t.pos_rep, 11 * 128, h1, h1, 64, 128, bf));
  if (rc != HFT_OK) return rc;
  t.ws_batch = B;
  return HFT_OK;
}

// ---- launchers ---------------------------------------------------------------------------------------------------
template <bool BF16, int EPI, int NT, bool PAIR>
static int launch_gemm_t(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mr, const CUtensorMap& mo, const GemmParams& gp, int grid,
                         size_t smem, cudaStream_t s) {
  auto kern = gemm_kernel<BF16, EPI, NT, PAIR>;
  static bool attr_set = false;
  if (!attr_set) { HFT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_set = true; }
  if (!PAIR) {
    kern<<<grid, kGemmThreads, smem, s>>>(ma, mw, mr, mo, gp);
    return HFT_OK;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  HFT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mw, mr, mo, gp));
  return HFT_OK;
}

template <bool BF16, int EPI, bool PAIR>
static int launch_gemm_e(int n_tile, const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mr, const CUtensorMap& mo, const GemmParams& gp,
                         int grid, size_t smem, cudaStream_t s) {
  if (EPI == EPI_HEADS) return launch_gemm_t<BF16, EPI_HEADS, 192, PAIR>(ma, mw, mr, mo, gp, grid, smem, s);
  if (n_tile == 256) return launch_gemm_t<BF16, EPI == EPI_HEADS ? EPI_STORE : EPI, 256, PAIR>(ma, mw, mr, mo, gp, grid, smem, s);
  if (n_tile == 128) return launch_gemm_t<BF16, EPI == EPI_HEADS ? EPI_STORE : EPI, 128, PAIR>(ma, mw, mr, mo, gp, grid, smem, s);
  if (n_tile == 64) return launch_gemm_t<BF16, EPI == EPI_HEADS ? EPI_STORE : EPI, 64, PAIR>(ma, mw, mr, mo, gp, grid, smem, s);
  set_error("tc gemm: unsupported tile width %d", n_tile);
  return HFT_ERR_UNSUPPORTED;
}

template <bool BF16, bool PAIR>
static int launch_gemm_p(int epi, int n_tile, const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mr, const CUtensorMap& mo, const GemmParams& gp,
                         int grid, size_t smem, cudaStream_t s) {
  return epi == EPI_STORE ? launch_gemm_e<BF16, EPI_STORE, PAIR>(n_tile, ma, mw, mr, mo, gp, grid, smem, s)
       : epi == EPI_RELU  ? launch_gemm_e<BF16, EPI_RELU, PAIR>(n_tile, ma, mw, mr, mo, gp, grid, smem, s)
       : epi == EPI_LN    ? launch_gemm_e<BF16, EPI_LN, PAIR>(n_tile, ma, mw, mr, mo, gp, grid, smem, s)
                          : launch_gemm_e<BF16, EPI_HEADS, PAIR>(n_tile, ma, mw, mr, mo, gp, grid, smem, s);
}

// CTA pairs (cta_group::2) when M is a multiple of 256.  Default: split-operand mode only -- measured on B200 (r01): x3 GEMMs
// 1406 -> 1168 ms per hour of audio (W streaming halves), single-product bf16 GEMMs 573 -> 591 ms (HBM-bound, W already resident).
// HFT_TC_PAIR=0 / 1 forces single-CTA tiles / pairs everywhere.
static bool pair_enabled(bool x3) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HFT_TC_PAIR"); v = !e ? 2 : (e[0] == '0' ? 0 : 1); }
  return v == 2 ? x3 : v == 1;
}

// out: STORE tensor map of the output tensor (box 64 x 32: one epilogue warp's rows); resid: LOAD tensor map of the
// residual (box 64 x 128) for EPI_LN (added by the identity MMA).
static int launch_gemm(bool bf16, int epi, const CUtensorMap& ma, const W16& w, long long M, GemmParams gp, const CUtensorMap* mo,
                       const CUtensorMap* mr, cudaStream_t s) {
  HFT_REQUIRE(M % kBlockM == 0 && w.K % kBlockK == 0 && w.n_tile > 0, HFT_ERR_UNSUPPORTED, "tc gemm: M=%lld K=%d N=%d unsupported", M, w.K, w.N);
  static int sms = num_sms();
  const bool pair = pair_enabled(gp.x3 != 0) && M % (2 * kBlockM) == 0 && sms >= 2 * (w.N / w.n_tile);
  const int n_rows_w = pair ? w.n_tile / 2 : w.n_tile;
  gp.m_tiles = (int)(M / (pair ? 2 * kBlockM : kBlockM));
  gp.n_tiles = w.N / w.n_tile;
  gp.k_chunks = w.K / kBlockK;
  gp.bias = w.bias;
  gp.has_resid = (epi == EPI_LN && mr != nullptr) ? 1 : 0;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("HFT_TC_DEBUG"); dbg = e ? atoi(e) : 0; } gp.debug_flags = dbg; }
  if (epi == EPI_LN) HFT_REQUIRE(gp.n_tiles == 1, HFT_ERR_UNSUPPORTED, "tc gemm: LayerNorm epilogue needs the full row in one tile (N=%d)", w.N);
  // W stays resident when this CTA's slice fits beside a >= 3-deep A ring, the identity block and the store staging
  // (measured r01, bf16 K = 256: resident W + 3 A slots 573 ms/h vs streamed W 665 ms/h); otherwise W chunks stream
  // through a ring of their own.
  const bool x3p = gp.x3 && !gp.single;
  const size_t w_bytes = (size_t)n_rows_w * w.K * 2 * (x3p ? 2 : 1);
  const size_t w_chunk = (size_t)n_rows_w * kBlockK * 2;
  const int stage_rows = gp.stage_rows ? gp.stage_rows : 32;
  const size_t fixed = 1024 + (gp.has_resid ? 8192 : 0) + (size_t)kEpiWarps * stage_rows * 128 + kConstBytes + 512;
  const size_t budget = 227 * 1024;
  static int wres = -1;                      // HFT_TC_WRES=0: never keep W resident (experiments)
  if (wres < 0) { const char* e = getenv("HFT_TC_WRES"); wres = (e && e[0] == '0') ? 0 : 1; }
  // split mode needs two A chunks (hi, lo) per k-chunk: below 4 slots the A stream starves (measured r01, pair K = 256: resident W +
  // 3 A slots 1 284 ms/h of GEMM vs streamed W + 6 A / 4 W slots 1 193 ms/h); single-product modes are fine with 3 (573 vs 665 ms/h)
  gp.w_resident = (wres && w_bytes + fixed + (x3p ? 4 : 3) * kChunkA <= budget) ? 1 : 0;
  static int wst = -1;                       // HFT_TC_WSTAGES: W ring depth for 16 KB chunks (experiments; default 4)
  if (wst < 0) { const char* e = getenv("HFT_TC_WSTAGES"); wst = e ? atoi(e) : 4; if (wst < 2 || wst > 8) wst = wst < 2 ? 2 : 8; }
  gp.w_stages = gp.w_resident ? 0 : (w_chunk <= 16384 ? wst : 2);
  const size_t w_smem = gp.w_resident ? w_bytes : gp.w_stages * w_chunk;
  long long a_st = (long long)(budget - fixed - w_smem) / kChunkA;
  gp.a_stages = a_st > 8 ? 8 : (int)a_st;
  HFT_REQUIRE(gp.a_stages >= (x3p ? 3 : 2), HFT_ERR_UNSUPPORTED, "tc gemm: no room for the operand rings (N tile %d, K %d)", w.n_tile, w.K);
  const int units_max = pair ? sms / 2 : sms;
  int units = (units_max / gp.n_tiles) * gp.n_tiles;
  if (units > gp.m_tiles * gp.n_tiles) units = gp.m_tiles * gp.n_tiles;
  const int grid = pair ? 2 * units : units;
  const size_t smem = gemm_smem_bytes(n_rows_w, gp.k_chunks, gp.w_resident, gp.a_stages, gp.w_stages, x3p, gp.has_resid, stage_rows);
  HFT_REQUIRE(smem <= 227 * 1024, HFT_ERR_UNSUPPORTED, "tc gemm: %zu bytes of shared memory needed", smem);
  const CUtensorMap& o = mo ? *mo : ma;
  const CUtensorMap& r = mr ? *mr : ma;
  const CUtensorMap& mw = pair ? w.map2 : w.map;
  LaunchScope ls(HFT_KCLASS_GEMM, s);
  if (pair) return bf16 ? launch_gemm_p<true, true>(epi, w.n_tile, ma, mw, r, o, gp, grid, smem, s) : launch_gemm_p<false, true>(epi, w.n_tile, ma, mw, r, o, gp, grid, smem, s);
  return bf16 ? launch_gemm_p<true, false>(epi, w.n_tile, ma, mw, r, o, gp, grid, smem, s) : launch_gemm_p<false, false>(epi, w.n_tile, ma, mw, r, o, gp, grid, smem, s);
}

template <bool BF16, int DH, int LK, bool PROBS, bool X3>
static int launch_attn_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mo, const AttnParams& ap, long long items,
                         cudaStream_t s) {
  auto kern = attn_kernel<BF16, DH, LK, PROBS, X3>;
  static bool attr_set = false;
  if (!attr_set) { HFT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<DH, LK, X3>::total)); attr_set = true; }
  kern<<<(unsigned)items, kAttnThreads, AttnSmem<DH, LK, X3>::total, s>>>(mq, mk, mv, mo, ap);
  return HFT_OK;
}

template <bool BF16, int DH, bool X3>
static int launch_attn_dh(int LK, bool probs, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mo, const AttnParams& ap,
                          long long items, cudaStream_t s) {
  if (LK == 256) return probs ? launch_attn_t<BF16, DH, 256, true, X3>(mq, mk, mv, mo, ap, items, s) : launch_attn_t<BF16, DH, 256, false, X3>(mq, mk, mv, mo, ap, items, s);
  if (LK == 128) return launch_attn_t<BF16, DH, 128, false, X3>(mq, mk, mv, mo, ap, items, s);
  if (LK == 96) return launch_attn_t<BF16, DH, 96, false, X3>(mq, mk, mv, mo, ap, items, s);
  set_error("tc attention: unsupported key tile %d", LK);
  return HFT_ERR_UNSUPPORTED;
}

// mo: tensor map of the ctx tensor (box 64 x 128) for the TMA-store epilogue, or nullptr for direct stores
static int launch_attn(int heads, int dh, bool bf16, bool x3, int LK, const CUtensorMap& mq, const CUtensorMap& mkv, const CUtensorMap* mo, AttnParams ap,
                       long long n_seq, cudaStream_t s) {
  ap.heads = heads;
  ap.tma_store = (mo != nullptr && dh == 64 && ap.lq % 128 == 0) ? 1 : 0;
  const CUtensorMap& mo_ = mo ? *mo : mq;
  ap.q_tiles = (ap.lq + 127) / 128;
  ap.scale_log2e = 1.4426950408889634f / sqrtf((float)dh);
  const long long items = n_seq * heads * ap.q_tiles;
  HFT_REQUIRE(items < (1ll << 31), HFT_ERR_UNSUPPORTED, "tc attention: too many work items");
  LaunchScope ls(HFT_KCLASS_ATTENTION, s);
  const bool probs = ap.probs != nullptr;
  if (x3) {                                             // split mode is fp16 only
    HFT_REQUIRE(!bf16, HFT_ERR_UNSUPPORTED, "x3 attention is built for fp16 parts");
    if (dh == 64) return launch_attn_dh<false, 64, true>(LK, probs, mq, mkv, mkv, mo_, ap, items, s);
    return launch_attn_dh<false, 32, true>(LK, probs, mq, mkv, mkv, mo_, ap, items, s);
  }
  if (dh == 64) return bf16 ? launch_attn_dh<true, 64, false>(LK, probs, mq, mkv, mkv, mo_, ap, items, s) : launch_attn_dh<false, 64, false>(LK, probs, mq, mkv, mkv, mo_, ap, items, s);
  return bf16 ? launch_attn_dh<true, 32, false>(LK, probs, mq, mkv, mkv, mo_, ap, items, s) : launch_attn_dh<false, 32, false>(LK, probs, mq, mkv, mkv, mo_, ap, items, s);
}

// ---- pipelined attention (tc_attn2.cuh): dh = 64, no probabilities -------------------------------------------------------
template <bool BF16, int NKEY, int NH, bool X3, bool PROBS = false>
static int launch_attn2_t(const CUtensorMap& mq, const CUtensorMap& mkv, const CUtensorMap& mo, const Attn2Params& ap, cudaStream_t s) {
  auto kern = attn2_kernel<BF16, 64, NKEY, NH, X3, PROBS>;
  static bool attr_set = false;
  if (!attr_set) { HFT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn2Smem<64, X3>::total)); attr_set = true; }
  static int sms = num_sms();
  const int grid = ap.n_rounds < sms ? ap.n_rounds : sms;
  kern<<<grid, kAttn2Threads, Attn2Smem<64, X3>::total, s>>>(mq, mkv, mo, ap);
  return HFT_OK;
}

template <bool BF16, bool X3>
static int launch_attn2_k(int LK, const CUtensorMap& mq, const CUtensorMap& mkv, const CUtensorMap& mo, const Attn2Params& ap, cudaStream_t s) {
  if (LK == 256 && ap.probs) return launch_attn2_t<BF16, 128, 2, X3, true>(mq, mkv, mo, ap, s);
  if (LK == 256) return launch_attn2_t<BF16, 128, 2, X3>(mq, mkv, mo, ap, s);
  if (LK == 128) return launch_attn2_t<BF16, 128, 1, X3>(mq, mkv, mo, ap, s);
  if (LK == 96) return launch_attn2_t<BF16, 96, 1, X3>(mq, mkv, mo, ap, s);
  set_error("tc attention: unsupported key tile %d", LK);
  return HFT_ERR_UNSUPPORTED;
}

// mkv_unit: tensor map of the K/V tensor with box dh x 128 (LK 256 / 128) or dh x 96; mo: store map of ctx, box 64 x 32
static int launch_attn2(int heads, bool bf16, bool x3, int LK, const CUtensorMap& mq, const CUtensorMap& mkv_unit, const CUtensorMap& mo, const AttnParams& a,
                        long long n_seq, cudaStream_t s) {
  Attn2Params ap{};
  ap.lq = a.lq; ap.lk = a.lk; ap.q_seq_rows = a.q_seq_rows; ap.q_tiles = (a.lq + 127) / 128; ap.heads = heads;
  ap.q_col0 = a.q_col0; ap.k_col0 = a.k_col0; ap.v_col0 = a.v_col0;
  ap.scale_log2e = 1.4426950408889634f / sqrtf(64.f);
  ap.ctx = a.ctx; ap.ld_ctx = a.ld_ctx; ap.q_lo_off = a.q_lo_off; ap.kv_lo_off = a.kv_lo_off; ap.ctx_lo_off = a.ctx_lo_off;
  ap.tma_store = (a.lq % 128 == 0) ? 1 : 0;
  const long long items = n_seq * heads * ap.q_tiles;
  HFT_REQUIRE(items < (1ll << 30) && ap.q_tiles <= 2, HFT_ERR_UNSUPPORTED, "tc attention: %lld work items / %d query tiles unsupported", items, ap.q_tiles);
  ap.n_items = (int)items;
  ap.n_rounds = (int)((items + 1) / 2);
  ap.shared_kv = ap.q_tiles == 2 ? 1 : 0;
  ap.probs = a.probs;
  ap.s_single = a.s_single; ap.pv_single = a.pv_single;
  HFT_REQUIRE(!ap.probs || LK == 256, HFT_ERR_UNSUPPORTED, "tc attention: probabilities are returned for 256-key sequences only");
  LaunchScope ls(HFT_KCLASS_ATTENTION, s);
  if (x3) {
    HFT_REQUIRE(!bf16, HFT_ERR_UNSUPPORTED, "x3 attention is built for fp16 parts");
    return launch_attn2_k<false, true>(LK, mq, mkv_unit, mo, ap, s);
  }
  return bf16 ? launch_attn2_k<true, false>(LK, mq, mkv_unit, mo, ap, s) : launch_attn2_k<false, false>(LK, mq, mkv_unit, mo, ap, s);
}

// 1 (default): pipelined kernel wherever it applies (dh 64, no probabilities); HFT_TC_ATTN2=0 keeps the one-tile-per-CTA kernel
static bool attn2_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HFT_TC_ATTN2"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// Probabilities-returning attention (tc_attn2.cuh, attn_probs_kernel): persistent single-warpgroup kernel for Lq <= 128, Lk = 256, dh = 64
template <bool BF16, bool X3>
static int launch_attn_probs_t(const CUtensorMap& mq, const CUtensorMap& mkv, const Attn2Params& ap, cudaStream_t s) {
  auto kern = attn_probs_kernel<BF16, X3>;
  static bool attr_set = false;
  if (!attr_set) { HFT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnProbsSmem<X3>::total)); attr_set = true; }
  static int sms = num_sms();
  const int grid = ap.n_items < sms ? ap.n_items : sms;
  kern<<<grid, kAttnProbsThreads, AttnProbsSmem<X3>::total, s>>>(mq, mkv, ap);
  return HFT_OK;
}
// mkv256: K/V tensor map with box dh x 256
static int launch_attn_probs(int heads, bool bf16, bool x3, const CUtensorMap& mq, const CUtensorMap& mkv256, const AttnParams& a, long long n_seq, cudaStream_t s) {
  Attn2Params ap{};
  ap.lq = a.lq; ap.lk = a.lk; ap.q_seq_rows = a.q_seq_rows; ap.q_tiles = 1; ap.heads = heads;
  ap.q_col0 = a.q_col0; ap.k_col0 = a.k_col0; ap.v_col0 = a.v_col0;
  ap.scale_log2e = 1.4426950408889634f / sqrtf(64.f);
  ap.ctx = a.ctx; ap.ld_ctx = a.ld_ctx; ap.q_lo_off = a.q_lo_off; ap.kv_lo_off = a.kv_lo_off; ap.ctx_lo_off = a.ctx_lo_off;
  ap.probs = a.probs;
  ap.s_single = a.s_single; ap.pv_single = a.pv_single;
  const long long items = n_seq * heads;
  HFT_REQUIRE(items < (1ll << 30) && a.lq <= 128 && a.lk == 256, HFT_ERR_UNSUPPORTED, "tc attention (probabilities): lq=%d lk=%d unsupported", a.lq, a.lk);
  ap.n_items = (int)items;
  LaunchScope ls(HFT_KCLASS_ATTENTION, s);
  if (x3) return launch_attn_probs_t<false, true>(mq, mkv256, ap, s);
  return bf16 ? launch_attn_probs_t<true, false>(mq, mkv256, ap, s) : launch_attn_probs_t<false, false>(mq, mkv256, ap, s);
}
// HFT_TC_ATTN_PROBS=0 keeps the one-tile-per-CTA kernel for the probabilities-returning attention
static bool attn_probs_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HFT_TC_ATTN_PROBS"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// The probabilities-returning cross-attention: the pipelined kernel has a PROBS variant (un-normalised rows written during the
// softmax pass, rescaled in place by the same thread), but its row-per-lane global accesses make it slower than the
// one-tile-per-CTA kernel (measured r01: attention 433 -> 447 ms per hour), so it is opt-in: HFT_TC_ATTN2_PROBS=1.
static bool attn2_probs_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HFT_TC_ATTN2_PROBS"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

#define HFT_TRY(x) do { int _rc = (x); if (_rc != HFT_OK) return _rc; } while (0)

// one projection launch: out[:, out_col0 ..] = epilogue(A * W^T); out_width = columns of the output tensor's hi block
static int linear(TcState& t, cudaStream_t s, int epi, const CUtensorMap& a, const W16& w, long long M, const CUtensorMap& out, int out_col0,
                  int out_width, const CUtensorMap* resid = nullptr, const LnW* ln = nullptr, Model* m = nullptr, int resid_period = 0, bool single = false) {
  GemmParams g{};
  g.out_col0 = out_col0;
  g.resid_period = resid_period;
  g.x3 = t.x3 ? 1 : 0;
  g.single = (t.x3 && single) ? 1 : 0;
  g.a_lo_off = w.K; g.w_lo_off = w.K; g.out_lo_off = out_width;
  if (ln) { g.gamma = m->w[ln->g]; g.beta = m->w[ln->b]; }
  const CUtensorMap* o = &out;
  if (gemm_stage_rows() != 32) {                   // the same target through the map with the smaller store box
    const CUtensorMap* from[6] = {&t.sX, &t.sHID, &t.sT, &t.sU, &t.sQKV, &t.sDQ};
    const CUtensorMap* to[6] = {&t.gX, &t.gHID, &t.gT, &t.gU, &t.gQKV, &t.gDQ};
    for (int i = 0; i < 6; ++i)
      if (o == from[i]) { o = to[i]; g.stage_rows = gemm_stage_rows(); break; }
  }
  return launch_gemm(t.bf16, epi, a, w, M, g, o, resid, s);
}

// mkv: K/V tensor map with box dh x LK (one-tile kernel); mkv_unit: box dh x 128 (or dh x 96) for the pipelined kernel
static int attention(Model* m, TcState& t, cudaStream_t s, int LK, const CUtensorMap& mq, const CUtensorMap& mkv, const CUtensorMap& mkv_unit, AttnParams a,
                     long long n_seq, int q_width) {
  a.q_lo_off = q_width;            // hi-block width of the Q tensor (3H for fused QKV buffers, H for the pitch-query table)
  a.kv_lo_off = 3 * m->H;
  a.ctx = t.CTX; a.ld_ctx = m->H * t.cm(); a.ctx_lo_off = m->H;
  if (m->dh == 64 && a.probs != nullptr && a.lq <= 128 && a.lk == 256 && LK == 256 && attn_probs_enabled() && !attn2_probs_enabled())
    return launch_attn_probs(m->heads, t.bf16, t.x3, mq, mkv, a, n_seq, s);
  if (m->dh == 64 && attn2_enabled() && (a.probs == nullptr || attn2_probs_enabled())) return launch_attn2(m->heads, t.bf16, t.x3, LK, mq, mkv_unit, t.sCTX, a, n_seq, s);
  return launch_attn(m->heads, m->dh, t.bf16, t.x3, LK, mq, mkv, &t.mCTX, a, n_seq, s);
}

// Fused FFN block (tc_ffn.cuh): x = LN(x + fc_2(relu(fc_1(x)))) in place, where it applies (hid 256, pf 512, rows % 256 == 0).
// Measured on B200 (r01, same box back to back): bf16 3 989 -> 4 181 x real-time, fp16x3 2 025 -> 2 138 x.  In split mode the
// resident x tile (128 KB hi|lo) leaves room for only 4 W ring slots, so the kernel is bound by the latency of the W stream
// (profiles/r01_prof_ffn_x3_*), not by HBM: it moves 3.5 x fewer bytes than the two GEMMs it replaces.  HFT_TC_FFN=0 disables it.
static bool ffn_fused_enabled(bool x3) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HFT_TC_FFN"); v = (e && e[0] == '0') ? 0 : 1; }
  (void)x3;
  return v == 1;
}
static bool ffn_fusable(Model* m, const TcState& t, long long R) { return ffn_fused_enabled(t.x3) && m->H == kFfnH && m->P == kFfnP && R % (2 * kBlockM) == 0; }

static int ffn_fused(Model* m, TcState& t, cudaStream_t s, const CUtensorMap& mx, const CUtensorMap& sx, long long R, const W16& w1, const W16& w2, const LnW& ln,
                     bool single = false) {
  FfnParams fp{};
  fp.m_tiles = (int)(R / (2 * kBlockM));
  fp.x3 = t.x3 ? 1 : 0;
  fp.single = (t.x3 && single) ? 1 : 0;
  fp.lo_off = m->H; fp.w1_lo_off = m->H; fp.w2_lo_off = m->P;
  fp.b1 = w1.bias; fp.b2 = w2.bias; fp.gamma = m->w[ln.g]; fp.beta = m->w[ln.b];
  static int slots_env = -1;
  if (slots_env < 0) { const char* e = getenv("HFT_TC_FFN_SLOTS"); slots_env = e ? atoi(e) : 0; }
  if (slots_env >= 2 && slots_env < ffn_slots(fp.x3)) fp.slots = slots_env;
  static int sms = num_sms();
  int units = sms / 2;
  if (units > fp.m_tiles) units = fp.m_tiles;
  const size_t smem = ffn_smem_bytes(fp.x3);
  HFT_REQUIRE(smem <= 227 * 1024, HFT_ERR_UNSUPPORTED, "fused ffn: %zu bytes of shared memory needed", smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * units));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  LaunchScope ls(HFT_KCLASS_GEMM, s);
  if (t.bf16) {
    HFT_SET_MAX_SMEM(ffn_kernel<true>, 227 * 1024);
    HFT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ffn_kernel<true>, mx, w1.map64, w2.map2, sx, fp));
  } else {
    HFT_SET_MAX_SMEM(ffn_kernel<false>, 227 * 1024);
    HFT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ffn_kernel<false>, mx, w1.map64, w2.map2, sx, fp));
  }
  return HFT_OK;
}

// EncoderLayer (model_spec2midi.py:230-245) over S sequences of L tokens held in x [S*L, H] (16-bit, updated in place)
static int encoder_layer_tc(Model* m, TcState& t, cudaStream_t s, const CUtensorMap& mx, const CUtensorMap& sx, const CUtensorMap& sqkv, const CUtensorMap& mq,
                            const CUtensorMap& mkv, const CUtensorMap& mkv_unit, int LK, long long S, int L, const TcLayer& lw, const LnW& ln, SitePlan sp = SitePlan()) {
  const int H = m->H, P = m->P;
  const long long R = S * L;
  HFT_TRY(linear(t, s, EPI_STORE, mx, lw.qkv, R, sqkv, 0, 3 * H, nullptr, nullptr, nullptr, 0, sp.lin));
  AttnParams a{};
  a.lq = L; a.lk = L; a.q_seq_rows = L; a.q_col0 = 0; a.k_col0 = H; a.v_col0 = 2 * H; a.probs = nullptr;
  a.s_single = sp.scores; a.pv_single = sp.pv;
  HFT_TRY(attention(m, t, s, LK, mq, mkv, mkv_unit, a, S, 3 * H));
  HFT_TRY(linear(t, s, EPI_LN, t.mCTX, lw.o, R, sx, 0, H, &mx, &ln, m, 0, sp.lin));          // x = LN(x + fc_o(ctx))
  if (ffn_fusable(m, t, R)) return ffn_fused(m, t, s, mx, sx, R, lw.w1, lw.w2, ln, sp.lin);   // x = LN(x + fc_2(relu(fc_1(x)))), hidden kept in TMEM
  HFT_TRY(linear(t, s, EPI_RELU, mx, lw.w1, R, t.sHID, 0, P, nullptr, nullptr, nullptr, 0, sp.lin));
  HFT_TRY(linear(t, s, EPI_LN, t.mHID, lw.w2, R, sx, 0, H, &mx, &ln, m, 0, sp.lin));          // x = LN(x + fc_2(relu(fc_1(x))))
  return HFT_OK;
}

int tc_prepare_weights(Model* m, cudaStream_t s) {
  (void)s;
  if (m->tc) {                                    // weights changed: 16-bit copies are rebuilt lazily per dtype
    TcBoth* b = reinterpret_cast<TcBoth*>(m->tc);
    for (auto& t : b->st) t.weights_ready = false;
  }
  return HFT_OK;
}

void tc_destroy(Model* m) {
  if (!m->tc) return;
  TcBoth* b = reinterpret_cast<TcBoth*>(m->tc);
  for (auto& t : b->st) { cudaFree(t.warena); cudaFree(t.head_bias); cudaFree(t.ws); }
  delete b;
  m->tc = nullptr;
}

// drop the activation workspaces of every precision mode (the 16-bit weight copies stay); the next forward re-allocates for its max_batch
void tc_release_workspace(Model* m) {
  if (!m->tc) return;
  TcBoth* b = reinterpret_cast<TcBoth*>(m->tc);
  for (auto& t : b->st) { cudaFree(t.ws); t.ws = nullptr; t.ws_batch = 0; }
}

static TcState& state_for(Model* m, int precision) {
  if (!m->tc) {
    TcBoth* b = new TcBoth();
    b->st[0].bf16 = false;
    b->st[1].bf16 = true;
    b->st[2].bf16 = false;
    b->st[2].x3 = true;
    m->tc = b;
  }
  return reinterpret_cast<TcBoth*>(m->tc)->st[precision == HFT_PREC_BF16 ? 1 : (precision == HFT_PREC_F16X3 || precision == HFT_PREC_MIXED) ? 2 : 0];
}

int forward_tc(Model* m, int precision, const float* spec, long long sb, long long sbin, long long st, int B, const hft_outputs* o, cudaStream_t s) {
  if (B == 0) return HFT_OK;
  HFT_REQUIRE(m->H == 64 || m->H == 128 || m->H == 256, HFT_ERR_UNSUPPORTED, "tensor-core path supports hid_dim 64 / 128 / 256 (got %d)", m->H);
  TcState& t = state_for(m, precision);
  t.mixed = precision == HFT_PREC_MIXED;          // shares weights and workspace with HFT_PREC_F16X3; only the per-site product counts differ
  HFT_REQUIRE(!t.mixed || (m->dh == 64 && attn2_enabled()), HFT_ERR_UNSUPPORTED, "HFT_PREC_MIXED is built for head_dim 64 (the pipelined attention kernels)");
  const bool bf = t.bf16;
  if (!t.weights_ready) HFT_TRY(prepare_weights(m, t, s));
  HFT_TRY(ensure_ws(m, t, B > m->max_batch ? B : m->max_batch));
  const int H = m->H, F = m->nframe, NB = m->nbin, NN = m->nnote, V = m->nvel;
  const long long Se = (long long)B * F, Re = Se * NB, Rd = Se * NN;
  const float sqrtH = sqrtf((float)H);
  HFT_REQUIRE(m->nproc == 65, HFT_ERR_UNSUPPORTED, "front kernel is built for n_margin 32");
  {
    LaunchScope ls(HFT_KCLASS_FRONT, s);
    static int tc_front = -1;                       // HFT_TC_FRONT=0: the CUDA-core filter kernel (the only one for hid != 256)
    if (tc_front < 0) { const char* e = getenv("HFT_TC_FRONT"); tc_front = (e && e[0] == '0') ? 0 : 1; }
    if (tc_front && H == kFrontH && F == kFrontF) {
      FrontParams fp{};
      fp.spec = spec; fp.sb = sb; fp.sbin = sbin; fp.st = st;
      fp.Wc = m->front_w; fp.bc = m->front_b; fp.pos = m->w[m->pos_freq];
      fp.scale = sqrtH; fp.n_bin = NB; fp.n_tiles = B * NB; fp.x3 = t.x3 ? 1 : 0; fp.lo_off = H;
      static int sms = num_sms();
      const unsigned grid = (unsigned)(fp.n_tiles < sms ? fp.n_tiles : sms);
      const size_t smem = sizeof(FrontSmem) + 1024;
      if (bf) {
        HFT_SET_MAX_SMEM(front_tc_kernel<true>, smem);
        front_tc_kernel<true><<<grid, kFrontThreads, smem, s>>>(t.sX3, fp);
      } else {
        HFT_SET_MAX_SMEM(front_tc_kernel<false>, smem);
        front_tc_kernel<false><<<grid, kFrontThreads, smem, s>>>(t.sX3, fp);
      }
    } else if (bf) front16_kernel<true, 65><<<dim3(NB, B), 256, 0, s>>>(spec, sb, sbin, st, m->front_w, m->front_b, m->w[m->pos_freq], sqrtH, H, F, NB, t.x3, t.X);
    else front16_kernel<false, 65><<<dim3(NB, B), 256, 0, s>>>(spec, sb, sbin, st, m->front_w, m->front_b, m->w[m->pos_freq], sqrtH, H, F, NB, t.x3, t.X);
  }
  for (size_t l = 0; l < m->enc.size(); ++l)
    HFT_TRY(encoder_layer_tc(m, t, s, t.mX, t.sX, t.sQKV, t.mQKV_q, t.mQKV_kv, t.mQKV_q, 256, Se, NB, t.enc[l], m->enc[l].ln, plan_for(t, 'e', (int)l)));

  const int n_cross = 1 + (int)m->dec.size();
  auto ffn = [&](const TcDecLayer& lw, const LnW& ln) -> int {
    if (ffn_fusable(m, t, Rd)) return ffn_fused(m, t, s, t.mT, t.sT, Rd, lw.w1, lw.w2, ln);
    HFT_TRY(linear(t, s, EPI_RELU, t.mT, lw.w1, Rd, t.sHID, 0, m->P));
    HFT_TRY(linear(t, s, EPI_LN, t.mHID, lw.w2, Rd, t.sT, 0, H, &t.mT, &ln, m));
    return HFT_OK;
  };
  auto cross = [&](const TcDecLayer& lw, const LnW& ln, bool zero, float* probs) -> int {
    HFT_TRY(linear(t, s, EPI_STORE, t.mX, lw.ca_kv, Re, t.sQKV, H, 3 * H));   // K | V of the 256-bin memory at columns [H, 3H)
    AttnParams a{};
    a.lq = NN; a.lk = NB; a.k_col0 = H; a.v_col0 = 2 * H; a.probs = probs; a.q_col0 = 0;
    { const SitePlan sp = plan_for(t, 'd', 0); a.s_single = sp.scores; a.pv_single = sp.pv; }
    if (zero) {
      a.q_seq_rows = 0;
      HFT_TRY(attention(m, t, s, 256, t.mQ0, t.mQKV_kv, t.mQKV_q, a, Se, H));
      // t = LN(pos_embedding_freq + fc_o(ctx)): the residual is the constant pitch-query table, period lcm(88,128)/128 = 11 tiles
      HFT_TRY(linear(t, s, EPI_LN, t.mCTX, lw.ca_o, Rd, t.sT, 0, H, &t.mPosRep, &ln, m, 11));
    } else {
      HFT_TRY(linear(t, s, EPI_STORE, t.mT, lw.ca_q, Rd, t.sDQ, 0, 3 * H));
      a.q_seq_rows = NN;
      HFT_TRY(attention(m, t, s, 256, t.mDQ_q, t.mQKV_kv, t.mQKV_q, a, Se, 3 * H));
      HFT_TRY(linear(t, s, EPI_LN, t.mCTX, lw.ca_o, Rd, t.sT, 0, H, &t.mT, &ln, m));
    }
    return HFT_OK;
  };
  HFT_TRY(cross(t.dec0, m->dec0.ln, true, n_cross == 1 ? o->attention : nullptr));
  HFT_TRY(ffn(t.dec0, m->dec0.ln));
  for (size_t l = 0; l < m->dec.size(); ++l) {
    const TcDecLayer& lw = t.dec[l];
    const LnW& ln = m->dec[l].ln;
    HFT_TRY(linear(t, s, EPI_STORE, t.mT, lw.sa_qkv, Rd, t.sDQ, 0, 3 * H));
    AttnParams a{};
    a.lq = NN; a.lk = NN; a.q_seq_rows = NN; a.q_col0 = 0; a.k_col0 = H; a.v_col0 = 2 * H;
    { const SitePlan sp = plan_for(t, 'd', (int)l + 1); a.s_single = sp.scores; a.pv_single = sp.pv; }
    HFT_TRY(attention(m, t, s, 96, t.mDQ_q, t.mDQ_kv, t.mDQ_kv, a, Se, 3 * H));
    HFT_TRY(linear(t, s, EPI_LN, t.mCTX, lw.sa_o, Rd, t.sT, 0, H, &t.mT, &ln, m));
    HFT_TRY(cross(lw, ln, false, ((int)l + 2 == n_cross) ? o->attention : nullptr));
    HFT_TRY(ffn(lw, ln));
  }
  // heads A
  {
    GemmParams g{};
    g.onset = o->onset_A; g.offset = o->offset_A; g.mpe = o->mpe_A; g.velocity = o->velocity_A; g.vel_argmax = o->velocity_A_argmax; g.n_vel = V; g.time_major = 0; g.n_frame = F; g.n_note = NN;
    g.x3 = t.x3; g.a_lo_off = H; g.w_lo_off = H; g.single = t.mixed ? 1 : 0;
    HFT_TRY(launch_gemm(bf, EPI_HEADS, t.mT, t.headA, Rd, g, nullptr, nullptr, s));
  }
  {
    LaunchScope ls(HFT_KCLASS_NORM, s);
    const long long total8 = Rd * H / 8;
    if (bf) time_relayout16_kernel<true><<<(unsigned)((total8 + 255) / 256), 256, 0, s>>>(t.T, m->w[m->pos_time], sqrtH, F, NN, H, t.x3, total8, t.U);
    else time_relayout16_kernel<false><<<(unsigned)((total8 + 255) / 256), 256, 0, s>>>(t.T, m->w[m->pos_time], sqrtH, F, NN, H, t.x3, total8, t.U);
  }
  for (size_t l = 0; l < m->tim.size(); ++l)
    HFT_TRY(encoder_layer_tc(m, t, s, t.mU, t.sU, t.sDQ, t.mDQ_q, t.mDQ_q, t.mDQ_q, 128, (long long)B * NN, F, t.tim[l], m->tim[l].ln, plan_for(t, 't', (int)l)));
  {
    GemmParams g{};
    g.onset = o->onset_B; g.offset = o->offset_B; g.mpe = o->mpe_B; g.velocity = o->velocity_B; g.vel_argmax = o->velocity_B_argmax; g.n_vel = V; g.time_major = 1; g.n_frame = F; g.n_note = NN;
    g.x3 = t.x3; g.a_lo_off = H; g.w_lo_off = H; g.single = t.mixed ? 1 : 0;
    HFT_TRY(launch_gemm(bf, EPI_HEADS, t.mU, t.headB, Rd, g, nullptr, nullptr, s));
  }
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}

}  // namespace hft

// ---- component entry points (include/hft_sm100.h) ---------------------------------------------------------------
using namespace hft;

extern "C" int hft_tc_linear(int bf16, int epi, const void* a16, const void* w16, const float* bias, int64_t M, int32_t N, int32_t K, void* out16,
                             const void* resid16, const float* gamma, const float* beta, void* stream) {
  HFT_REQUIRE(a16 && w16 && bias && out16, HFT_ERR_ARG, "hft_tc_linear: NULL buffer");
  HFT_REQUIRE(epi >= 0 && epi <= 2, HFT_ERR_ARG, "hft_tc_linear: epi %d", epi);
  HFT_REQUIRE(M % 128 == 0 && K % 64 == 0 && N % 64 == 0 && (epi != 2 || ((N == 64 || N == 128 || N == 256) && resid16 && gamma && beta)),
              HFT_ERR_UNSUPPORTED, "hft_tc_linear: M=%lld N=%d K=%d epi=%d unsupported", (long long)M, N, K, epi);
  W16 w;
  w.ptr = (uint16_t*)w16; w.N = N; w.K = K; w.n_tile = n_tile_for(N, false, epi == 2); w.bias = bias;
  CUtensorMap ma, mo, mr;
  HFT_TRY(make_map(&ma, a16, M, K, K, kBlockK, kBlockM, bf16 != 0));
  HFT_TRY(make_map(&w.map, w16, N, K, K, kBlockK, w.n_tile, bf16 != 0));
  HFT_TRY(make_map(&w.map2, w16, N, K, K, kBlockK, w.n_tile / 2, bf16 != 0));
  HFT_TRY(make_map(&mo, out16, M, N, N, 64, 32, bf16 != 0));
  if (epi == 2) HFT_TRY(make_map(&mr, resid16, M, N, N, 64, 128, bf16 != 0));
  GemmParams g{};
  g.gamma = gamma; g.beta = beta;
  reset_launch_count();
  return launch_gemm(bf16 != 0, epi, ma, w, M, g, &mo, epi == 2 ? &mr : nullptr, (cudaStream_t)stream);
}

extern "C" int hft_tc_ffn(int bf16, const void* x16, const void* w1_16, const float* b1, const void* w2_16, const float* b2, const float* gamma, const float* beta,
                          int64_t M, void* out16, void* stream) {
  HFT_REQUIRE(x16 && w1_16 && b1 && w2_16 && b2 && gamma && beta && out16, HFT_ERR_ARG, "hft_tc_ffn: NULL buffer");
  HFT_REQUIRE(M > 0 && M % (2 * kBlockM) == 0, HFT_ERR_UNSUPPORTED, "hft_tc_ffn: M=%lld must be a multiple of 256", (long long)M);
  CUtensorMap mx, mw1, mw2, mo;
  HFT_TRY(make_map(&mx, x16, M, kFfnH, kFfnH, kBlockK, kBlockM, bf16 != 0));
  HFT_TRY(make_map(&mw1, w1_16, kFfnP, kFfnH, kFfnH, kBlockK, 64, bf16 != 0));
  HFT_TRY(make_map(&mw2, w2_16, kFfnH, kFfnP, kFfnP, kBlockK, 128, bf16 != 0));
  HFT_TRY(make_map(&mo, out16, M, kFfnH, kFfnH, 64, 32, bf16 != 0));
  FfnParams fp{};
  fp.m_tiles = (int)(M / (2 * kBlockM));
  fp.x3 = 0; fp.lo_off = kFfnH; fp.w1_lo_off = kFfnH; fp.w2_lo_off = kFfnP;
  fp.b1 = b1; fp.b2 = b2; fp.gamma = gamma; fp.beta = beta;
  int units = num_sms() / 2;
  if (units > fp.m_tiles) units = fp.m_tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * units));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = ffn_smem_bytes(0);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  reset_launch_count();
  LaunchScope ls(HFT_KCLASS_GEMM, stream);
  if (bf16) {
    HFT_CHECK_CUDA(cudaFuncSetAttribute(ffn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    HFT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ffn_kernel<true>, mx, mw1, mw2, mo, fp));
  } else {
    HFT_CHECK_CUDA(cudaFuncSetAttribute(ffn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    HFT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ffn_kernel<false>, mx, mw1, mw2, mo, fp));
  }
  return HFT_OK;
}

extern "C" int hft_tc_attention(int bf16, int32_t dh, int32_t heads, const void* qkv16, int64_t n_seq, int32_t L, void* ctx16, float* probs,
                                void* stream) {
  HFT_REQUIRE(qkv16 && ctx16 && n_seq > 0, HFT_ERR_ARG, "hft_tc_attention: bad argument");
  HFT_REQUIRE((dh == 64 || dh == 32) && (L == 256 || L == 128 || L == 88) && (!probs || L == 256), HFT_ERR_UNSUPPORTED,
              "hft_tc_attention: dh=%d L=%d unsupported", dh, L);
  const int H = heads * dh, LK = L == 88 ? 96 : L;
  CUtensorMap mq, mkv;
  HFT_TRY(make_map(&mq, qkv16, n_seq * L, 3 * H, 3 * H, dh, 128, bf16 != 0));
  HFT_TRY(make_map(&mkv, qkv16, n_seq * L, 3 * H, 3 * H, dh, LK, bf16 != 0));
  AttnParams a{};
  a.lq = L; a.lk = L; a.q_seq_rows = L; a.q_col0 = 0; a.k_col0 = H; a.v_col0 = 2 * H; a.ctx = ctx16; a.ld_ctx = H; a.probs = probs;
  CUtensorMap mo;
  const bool use_tma = dh == 64 && L % 128 == 0;
  if (use_tma) HFT_TRY(make_map(&mo, ctx16, n_seq * L, H, H, 64, 128, bf16 != 0));
  reset_launch_count();
  if (dh == 64 && attn2_enabled() && (!probs || attn2_probs_enabled())) {   // pipelined kernel (tc_attn2.cuh)
    CUtensorMap mku, so;
    HFT_TRY(make_map(&mku, qkv16, n_seq * L, 3 * H, 3 * H, dh, LK == 96 ? 96 : 128, bf16 != 0));
    HFT_TRY(make_map(&so, ctx16, n_seq * L, H, H, 64, 32, bf16 != 0));
    return launch_attn2(heads, bf16 != 0, false, LK, mq, mku, so, a, n_seq, (cudaStream_t)stream);
  }
  return launch_attn(heads, dh, bf16 != 0, false, LK, mq, mkv, use_tma ? &mo : nullptr, a, n_seq, (cudaStream_t)stream);
}
