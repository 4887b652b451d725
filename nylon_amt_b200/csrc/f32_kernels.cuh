// fp32 CUDA-core kernels of the hFT forward (shared by the inference forward fwd_f32.cu and the training step
// train_f32.cu): collapsed front filter, tiled SGEMM (+bias, +ReLU, +accumulate), two-pass softmax attention (optionally
// returning the row log-sum-exp for the backward pass), residual+LayerNorm, time re-layout, head finish.
#pragma once
#include "common.cuh"
#include <math.h>

namespace hft {
namespace {

// ---- dropout (training step only) ---------------------------------------------------------------------------------------
// Counter-based: element `idx` of dropout site `site` is kept iff hash(seed, site, idx) >= thresh (thresh = p * 2^32), and kept
// values are scaled by 1 / (1 - p) (nn.Dropout semantics).  The backward pass regenerates the same decisions, nothing is stored.
struct Drop {
  uint32_t thresh;     // 0: no dropout
  uint32_t seed;
  uint32_t site;
  float scale;
};
__host__ __device__ __forceinline__ bool drop_keep(const Drop& d, unsigned long long idx) {
  uint32_t h = (uint32_t)idx * 0x9E3779B1u ^ (uint32_t)(idx >> 32) * 0x85EBCA77u ^ d.seed ^ (d.site * 0xC2B2AE3Du);
  h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
  return h >= d.thresh;
}
// x[i] = keep ? x[i] * scale : 0   (forward in place; also the backward of an element-wise dropout)
__global__ void dropout_kernel(float* __restrict__ x, long long n, Drop d) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  x[i] = drop_keep(d, (unsigned long long)i) ? x[i] * d.scale : 0.f;
}
// m[i] = keep ? scale : 0   (the multiplier itself; lets a restatement check its own mask arithmetic against the kernels')
__global__ void dropout_mask_kernel(float* __restrict__ m, long long n, Drop d) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  m[i] = drop_keep(d, (unsigned long long)i) ? d.scale : 0.f;
}
// y[i] = keep ? x[i] * scale : 0   (out of place: the sub-layer branch of a gradient that also feeds the residual path)
__global__ void dropout_copy_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, Drop d) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = drop_keep(d, (unsigned long long)i) ? x[i] * d.scale : 0.f;
}

// ------------------------------------------------------------------------------------------------------------
// front: unfold(65) -> conv(1,C,(1,kw)) -> Linear -> *sqrt(H) + pos  (model_spec2midi.py:65-95), collapsed to one
// 65-tap filter per hidden unit.  grid (n_bin, B); X[((b*F+f)*NB+bin)*H + h].
// ------------------------------------------------------------------------------------------------------------
template <int NPROC>
__global__ void __launch_bounds__(256) front_f32_kernel(const float* __restrict__ spec, long long sb, long long sbin, long long st,
                                                        const float* __restrict__ Wc, const float* __restrict__ bc, const float* __restrict__ pos,
                                                        float scale, int H, int F, int NB, float* __restrict__ X) {
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
    for (int i = 0; i < 8; ++i) X[(((long long)b * F + f0 + i) * NB + bin) * H + h] = (acc[i] + bias) * scale + pe;
  }
}

// ------------------------------------------------------------------------------------------------------------
// C[M,N] = A[M,K] * W[N,K]^T + bias (+ReLU).  128x128x16 tiles, 256 threads, 8x8 micro-tile.  K % 16 == 0.
// ------------------------------------------------------------------------------------------------------------
constexpr int GBM = 128, GBN = 128, GBK = 16, GPAD = 4;

// BN: tile width (128, or 64 for the narrow projections of the reduced model so that no half-empty tile is computed)
template <bool RELU, int BN = 128>
__global__ void __launch_bounds__(256) sgemm_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Wt, int ldw,
                                                       const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int N, int K,
                                                       bool accum = false) {
  constexpr int NJ = BN / 16;                      // columns per thread: 8 (two groups of 4, 64 apart) or 4
  __shared__ __align__(16) float As[GBK][GBM + GPAD];
  __shared__ __align__(16) float Ws[GBK][BN + GPAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * BN;
  const int lrow = tid >> 1, lk = (tid & 1) * 8;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][NJ];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
  const bool a_ok = (m0 + lrow) < M, w_ld = lrow < BN, w_ok = w_ld && (n0 + lrow) < N;
  const float* ap = A + (long long)(m0 + lrow) * lda + lk;
  const float* wp = Wt + (long long)(n0 + lrow) * ldw + lk;
  float4 a0, a1, w0, w1;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  a0 = a_ok ? *reinterpret_cast<const float4*>(ap) : z;
  a1 = a_ok ? *reinterpret_cast<const float4*>(ap + 4) : z;
  w0 = w_ok ? *reinterpret_cast<const float4*>(wp) : z;
  w1 = w_ok ? *reinterpret_cast<const float4*>(wp + 4) : z;
  for (int k0 = 0; k0 < K; k0 += GBK) {
    As[lk + 0][lrow] = a0.x; As[lk + 1][lrow] = a0.y; As[lk + 2][lrow] = a0.z; As[lk + 3][lrow] = a0.w;
    As[lk + 4][lrow] = a1.x; As[lk + 5][lrow] = a1.y; As[lk + 6][lrow] = a1.z; As[lk + 7][lrow] = a1.w;
    if (w_ld) {
      Ws[lk + 0][lrow] = w0.x; Ws[lk + 1][lrow] = w0.y; Ws[lk + 2][lrow] = w0.z; Ws[lk + 3][lrow] = w0.w;
      Ws[lk + 4][lrow] = w1.x; Ws[lk + 5][lrow] = w1.y; Ws[lk + 6][lrow] = w1.z; Ws[lk + 7][lrow] = w1.w;
    }
    __syncthreads();
    if (k0 + GBK < K) {                            // register prefetch of the next K slab
      a0 = a_ok ? *reinterpret_cast<const float4*>(ap + k0 + GBK) : z;
      a1 = a_ok ? *reinterpret_cast<const float4*>(ap + k0 + GBK + 4) : z;
      w0 = w_ok ? *reinterpret_cast<const float4*>(wp + k0 + GBK) : z;
      w1 = w_ok ? *reinterpret_cast<const float4*>(wp + k0 + GBK + 4) : z;
    }
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float4 ra0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 ra1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      float4 rb0 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      float4 rb1 = BN == 128 ? *reinterpret_cast<const float4*>(&Ws[k][(BN == 128 ? 64 : 0) + tx * 4]) : z;
      float ra[8] = {ra0.x, ra0.y, ra0.z, ra0.w, ra1.x, ra1.y, ra1.z, ra1.w};
      float rb[8] = {rb0.x, rb0.y, rb0.z, rb0.w, rb1.x, rb1.y, rb1.z, rb1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(ra[i], rb[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool vec = (ldc & 3) == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (row >= M) continue;
#pragma unroll
    for (int g = 0; g < NJ / 4; ++g) {
      const int col0 = n0 + g * 64 + tx * 4;
      float v[4] = {acc[i][4 * g], acc[i][4 * g + 1], acc[i][4 * g + 2], acc[i][4 * g + 3]};
      float* cp = C + (long long)row * ldc + col0;
      if (vec && col0 + 3 < N) {                      // four consecutive columns: 16-byte accesses
        if (bias) { v[0] += bias[col0]; v[1] += bias[col0 + 1]; v[2] += bias[col0 + 2]; v[3] += bias[col0 + 3]; }
        if (RELU) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
        if (accum) { const float4 c4 = *reinterpret_cast<const float4*>(cp); v[0] += c4.x; v[1] += c4.y; v[2] += c4.z; v[3] += c4.w; }
        *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (col0 + j >= N) continue;
          float w = v[j] + (bias ? bias[col0 + j] : 0.f);
          if (RELU) w = fmaxf(w, 0.f);
          if (accum) w += cp[j];
          cp[j] = w;
        }
      }
    }
  }
}

// tile width with the fewest padded columns (ties -> 128)
inline int sgemm_tile_n(int N) { return ((N + 63) / 64 * 64 < (N + 127) / 128 * 128) ? 64 : 128; }

// ------------------------------------------------------------------------------------------------------------
// attention for one (sequence, head): softmax(Q K^T / sqrt(dh)) V  (model_spec2midi.py:342-348), two passes over
// the keys (max+sum, then normalised PV); optionally writes the probabilities [S,heads,Lq,Lk] (:360).
// One thread per query row; the head's K and V slices live in shared memory.
// ------------------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(256) attn_f32_kernel(const float* __restrict__ Q, int ldq, long long q_seq_stride,
                                                       const float* __restrict__ Kp, const float* __restrict__ Vp, int ldkv, int Lq, int Lk,
                                                       int heads, float inv_scale, float* __restrict__ ctx, int ldc, float* __restrict__ probs,
                                                       float* __restrict__ lse = nullptr, Drop drop = Drop{0, 0, 0, 1.f}) {
  extern __shared__ __align__(16) float smem_attn[];
  float* sK = smem_attn;
  float* sV = smem_attn + (size_t)Lk * DH;
  const int seq = blockIdx.x, head = blockIdx.y;
  const float* kbase = Kp + (long long)seq * Lk * ldkv + head * DH;
  const float* vbase = Vp + (long long)seq * Lk * ldkv + head * DH;
  for (int i = threadIdx.x; i < Lk * (DH / 4); i += blockDim.x) {
    int j = i / (DH / 4), c = i % (DH / 4);
    reinterpret_cast<float4*>(sK)[i] = *reinterpret_cast<const float4*>(kbase + (long long)j * ldkv + c * 4);
    reinterpret_cast<float4*>(sV)[i] = *reinterpret_cast<const float4*>(vbase + (long long)j * ldkv + c * 4);
  }
  __syncthreads();
  const int r = threadIdx.x;
  if (r >= Lq) return;
  float q[DH];
  const float* qp = Q + (long long)seq * q_seq_stride + (long long)r * ldq + head * DH;
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) {
    float4 t = *reinterpret_cast<const float4*>(qp + c * 4);
    q[c * 4] = t.x; q[c * 4 + 1] = t.y; q[c * 4 + 2] = t.z; q[c * 4 + 3] = t.w;
  }
  float mx = -INFINITY, sum = 0.f;
  for (int j = 0; j < Lk; ++j) {
    const float4* kr = reinterpret_cast<const float4*>(sK + (size_t)j * DH);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      float4 t = kr[c];
      s = fmaf(q[c * 4], t.x, s); s = fmaf(q[c * 4 + 1], t.y, s); s = fmaf(q[c * 4 + 2], t.z, s); s = fmaf(q[c * 4 + 3], t.w, s);
    }
    s *= inv_scale;
    if (s > mx) { sum *= expf(mx - s); mx = s; }
    sum += expf(s - mx);
  }
  const float inv_sum = 1.f / sum;
  if (lse) lse[((long long)seq * heads + head) * Lq + r] = mx + logf(sum);      // row log-sum-exp, kept for the backward pass
  float acc[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) acc[c] = 0.f;
  float* prow = probs ? probs + (((long long)seq * heads + head) * Lq + r) * Lk : nullptr;
  for (int j = 0; j < Lk; ++j) {
    const float4* kr = reinterpret_cast<const float4*>(sK + (size_t)j * DH);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      float4 t = kr[c];
      s = fmaf(q[c * 4], t.x, s); s = fmaf(q[c * 4 + 1], t.y, s); s = fmaf(q[c * 4 + 2], t.z, s); s = fmaf(q[c * 4 + 3], t.w, s);
    }
    float p = expf(s * inv_scale - mx) * inv_sum;
    if (prow) prow[j] = p;
    if (drop.thresh)                                   // dropout on the probabilities (model_spec2midi.py:348), training only
      p = drop_keep(drop, (unsigned long long)((((long long)seq * heads + head) * Lq + r) * Lk + j)) ? p * drop.scale : 0.f;
    const float4* vr = reinterpret_cast<const float4*>(sV + (size_t)j * DH);
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      float4 t = vr[c];
      acc[c * 4] = fmaf(p, t.x, acc[c * 4]); acc[c * 4 + 1] = fmaf(p, t.y, acc[c * 4 + 1]);
      acc[c * 4 + 2] = fmaf(p, t.z, acc[c * 4 + 2]); acc[c * 4 + 3] = fmaf(p, t.w, acc[c * 4 + 3]);
    }
  }
  float* op = ctx + ((long long)seq * Lq + r) * ldc + head * DH;
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) *reinterpret_cast<float4*>(op + c * 4) = make_float4(acc[c * 4], acc[c * 4 + 1], acc[c * 4 + 2], acc[c * 4 + 3]);
}

// Training forward (no probabilities returned): the broadcast reads of K / V rows from shared memory bound the kernel above
// (one float delivered per warp and clock against four FMA issues), so here every thread owns TWO query rows -- each K / V
// value read feeds both -- and the softmax is the one-pass online form (running max, rescale on a new max): K is read once.
// Row log-sum-exp saved for the backward pass; dropout on the probabilities as above.
template <int DH>
__global__ void __launch_bounds__(128) attn_f32_r2_kernel(const float* __restrict__ Q, int ldq, long long q_seq_stride, const float* __restrict__ Kp,
                                                          const float* __restrict__ Vp, int ldkv, int Lq, int Lk, int heads, float inv_scale,
                                                          float* __restrict__ ctx, int ldc, float* __restrict__ lse, Drop drop) {
  extern __shared__ __align__(16) float smem_attn[];
  float* sK = smem_attn;
  float* sV = smem_attn + (size_t)Lk * DH;
  const int seq = blockIdx.x, head = blockIdx.y;
  const float* kbase = Kp + (long long)seq * Lk * ldkv + head * DH;
  const float* vbase = Vp + (long long)seq * Lk * ldkv + head * DH;
  for (int i = threadIdx.x; i < Lk * (DH / 4); i += blockDim.x) {
    int j = i / (DH / 4), c = i % (DH / 4);
    reinterpret_cast<float4*>(sK)[i] = *reinterpret_cast<const float4*>(kbase + (long long)j * ldkv + c * 4);
    reinterpret_cast<float4*>(sV)[i] = *reinterpret_cast<const float4*>(vbase + (long long)j * ldkv + c * 4);
  }
  __syncthreads();
  const int r0 = 2 * threadIdx.x;
  if (r0 >= Lq) return;
  const bool two = r0 + 1 < Lq;
  float q[2][DH], acc[2][DH];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const float* qp = Q + (long long)seq * q_seq_stride + (long long)(r0 + (two ? u : 0)) * ldq + head * DH;
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      float4 t = *reinterpret_cast<const float4*>(qp + c * 4);
      q[u][c * 4] = t.x * inv_scale; q[u][c * 4 + 1] = t.y * inv_scale; q[u][c * 4 + 2] = t.z * inv_scale; q[u][c * 4 + 3] = t.w * inv_scale;
    }
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[u][c] = 0.f;
  }
  float mx[2] = {-INFINITY, -INFINITY}, sum[2] = {0.f, 0.f};
  const long long pbase = (((long long)seq * heads + head) * Lq + r0) * Lk;
  for (int j = 0; j < Lk; ++j) {
    const float4* kr = reinterpret_cast<const float4*>(sK + (size_t)j * DH);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      const float4 t = kr[c];
      s0 = fmaf(q[0][c * 4], t.x, s0); s0 = fmaf(q[0][c * 4 + 1], t.y, s0); s0 = fmaf(q[0][c * 4 + 2], t.z, s0); s0 = fmaf(q[0][c * 4 + 3], t.w, s0);
      s1 = fmaf(q[1][c * 4], t.x, s1); s1 = fmaf(q[1][c * 4 + 1], t.y, s1); s1 = fmaf(q[1][c * 4 + 2], t.z, s1); s1 = fmaf(q[1][c * 4 + 3], t.w, s1);
    }
    float p[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float sv = u ? s1 : s0;
      if (sv > mx[u]) {                                   // new running max: rescale what has been accumulated
        const float f = expf(mx[u] - sv);
        sum[u] *= f;
#pragma unroll
        for (int c = 0; c < DH; ++c) acc[u][c] *= f;
        mx[u] = sv;
      }
      p[u] = expf(sv - mx[u]);
      sum[u] += p[u];
      if (drop.thresh) p[u] = drop_keep(drop, (unsigned long long)(pbase + (long long)u * Lk + j)) ? p[u] * drop.scale : 0.f;
    }
    const float4* vr = reinterpret_cast<const float4*>(sV + (size_t)j * DH);
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      const float4 t = vr[c];
      acc[0][c * 4] = fmaf(p[0], t.x, acc[0][c * 4]); acc[0][c * 4 + 1] = fmaf(p[0], t.y, acc[0][c * 4 + 1]);
      acc[0][c * 4 + 2] = fmaf(p[0], t.z, acc[0][c * 4 + 2]); acc[0][c * 4 + 3] = fmaf(p[0], t.w, acc[0][c * 4 + 3]);
      acc[1][c * 4] = fmaf(p[1], t.x, acc[1][c * 4]); acc[1][c * 4 + 1] = fmaf(p[1], t.y, acc[1][c * 4 + 1]);
      acc[1][c * 4 + 2] = fmaf(p[1], t.z, acc[1][c * 4 + 2]); acc[1][c * 4 + 3] = fmaf(p[1], t.w, acc[1][c * 4 + 3]);
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (u == 1 && !two) break;
    const float inv = 1.f / sum[u];
    lse[((long long)seq * heads + head) * Lq + r0 + u] = mx[u] + logf(sum[u]);
    float* op = ctx + ((long long)seq * Lq + r0 + u) * ldc + head * DH;
#pragma unroll
    for (int c = 0; c < DH / 4; ++c)
      *reinterpret_cast<float4*>(op + c * 4) = make_float4(acc[u][c * 4] * inv, acc[u][c * 4 + 1] * inv, acc[u][c * 4 + 2] * inv, acc[u][c * 4 + 3] * inv);
  }
}

// ------------------------------------------------------------------------------------------------------------
// y = LayerNorm(x + r) * g + b, eps 1e-5 (model_spec2midi.py:236,242: one LayerNorm module shared by the sites of
// a layer).  One warp per row.  r is indexed by (row % r_rows) so the constant pitch queries can be broadcast.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add_ln_f32_kernel(const float* __restrict__ x, const float* __restrict__ r, long long r_rows,
                                                         const float* __restrict__ g, const float* __restrict__ b, int H, long long rows,
                                                         float* __restrict__ y, float* __restrict__ sum_out = nullptr, Drop drop = Drop{0, 0, 0, 1.f},
                                                         bool drop_x = false) {
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xp = x + row * H;
  const float* rp = r + (row % r_rows) * H;
  float v[8];
  float s = 0.f;
  const int per = H >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < per) {
      float xv = xp[lane + 32 * i], rv = rp[lane + 32 * i];
      if (drop.thresh) {                               // sub-layer dropout before the residual add (model_spec2midi.py:236), training only
        const bool keep = drop_keep(drop, (unsigned long long)(row * H + lane + 32 * i));
        if (drop_x) xv = keep ? xv * drop.scale : 0.f; else rv = keep ? rv * drop.scale : 0.f;
      }
      v[i] = xv + rv;
      s += v[i];
      if (sum_out) sum_out[row * H + lane + 32 * i] = v[i];      // pre-LayerNorm sum, kept for the backward pass
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < per) { float d = v[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)H + 1e-5f);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < per) { int c = lane + 32 * i; y[row * H + c] = (v[i] - mean) * rstd * g[c] + b[c]; }
}

// U[((b*NN+n)*F+f)*H+h] = T[((b*F+f)*NN+n)*H+h] * sqrt(H) + pos_time[f*H+h]   (model_spec2midi.py:189-191)
__global__ void time_relayout_f32_kernel(const float* __restrict__ T, const float* __restrict__ pos, float scale, int F, int NN, int H,
                                         long long total, float* __restrict__ U) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int h = (int)(i % H);
  long long r = i / H;
  int f = (int)(r % F);
  long long bn = r / F;
  int n = (int)(bn % NN);
  long long b = bn / NN;
  U[i] = T[(((b * F + f) * NN + n)) * H + h] * scale + pos[f * H + h];
}

// heads (model_spec2midi.py:172-175 / :203-206): HT[row][0..2] -> sigmoid -> onset/offset/mpe, HT[row][3..] -> velocity.
// time_major: rows are (b,n,f) and outputs are permuted back to [B,F,NN(,V)].
__global__ void heads_finish_f32_kernel(const float* __restrict__ HT, int ldh, int V, int F, int NN, long long rows, bool time_major,
                                        float* __restrict__ onset, float* __restrict__ offset, float* __restrict__ mpe,
                                        float* __restrict__ velocity) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cols = 3 + V;
  if (i >= rows * cols) return;
  long long row = i / cols;
  int c = (int)(i % cols);
  long long orow = row;
  if (time_major) {
    int f = (int)(row % F);
    long long bn = row / F;
    int n = (int)(bn % NN);
    long long b = bn / NN;
    orow = (b * F + f) * NN + n;
  }
  float v = HT[row * ldh + c];
  if (c < 3) {
    float* dst = c == 0 ? onset : (c == 1 ? offset : mpe);
    if (dst) dst[orow] = 1.f / (1.f + expf(-v));
  } else if (velocity) {
    velocity[orow * V + (c - 3)] = v;
  }
}

// argmax over the velocity logits of every row of HT (columns 3 .. 3+V-1), first maximum like torch.argmax (amt.py:107,113)
__global__ void heads_argmax_f32_kernel(const float* __restrict__ HT, int ldh, int V, int F, int NN, long long rows, bool time_major,
                                        signed char* __restrict__ out) {
  long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  long long orow = row;
  if (time_major) {
    int f = (int)(row % F);
    long long bn = row / F;
    int n = (int)(bn % NN);
    orow = ((bn / NN) * F + f) * NN + n;
  }
  const float* p = HT + row * ldh + 3;
  float best = -INFINITY;
  int bi = 0;
  for (int j = 0; j < V; ++j) {
    const float v = p[j];
    if (v > best) { best = v; bi = j; }
  }
  out[orow] = (signed char)bi;
}

}  // namespace
}  // namespace hft
