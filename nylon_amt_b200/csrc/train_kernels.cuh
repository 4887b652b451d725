// fp32 CUDA-core kernels of the training step (reference hftt_code/training/train.py:89-160): the backward passes of the
// hFT layers (Linear, LayerNorm, multi-head attention, ReLU, embeddings, the conv+Linear front), the 8-term loss
// (BCELoss on the six sigmoid outputs + CrossEntropyLoss on the two velocity logit tensors) and Adam.
#pragma once
#include "f32_kernels.cuh"

namespace hft {
namespace {

// ------------------------------------------------------------------------------------------------------------
// C[M,N] (+)= A[M,K] * B[K,N]   (dX = dY * W with W = [out,in] row-major as B).  Optional ReLU mask: C *= (mask > 0).
// 128x128x16 tiles like sgemm_tn_kernel.  K % 16 == 0, N % 4 == 0, lda % 4 == 0, ldb % 4 == 0.
// ------------------------------------------------------------------------------------------------------------
template <int BN = 128>
__global__ void __launch_bounds__(256) sgemm_nn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
                                                       float* __restrict__ C, int ldc, int M, int N, int K, bool accum,
                                                       const float* __restrict__ mask, int ldm, float mask_scale) {
  constexpr int NJ = BN / 16;
  constexpr int BV = BN / 16;                             // floats loaded per thread for the B tile: 8 (two float4) or 4 (one)
  __shared__ __align__(16) float As[GBK][GBM + GPAD];
  __shared__ __align__(16) float Bs[GBK][BN + GPAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * BN;
  const int lrow = tid >> 1, lk = (tid & 1) * 8;          // A tile: 128 rows x 16 k
  const int brow = tid >> 4, bcol = (tid & 15) * BV;      // B tile: 16 k x BN cols
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][NJ];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
  const bool a_ok = (m0 + lrow) < M;
  const bool b_ok0 = (n0 + bcol) < N, b_ok1 = BN == 128 && (n0 + bcol + 4) < N;
  const float* ap = A + (long long)(m0 + lrow) * lda + lk;
  const float* bp = Bm + (long long)brow * ldb + n0 + bcol;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 a0 = a_ok ? *reinterpret_cast<const float4*>(ap) : z, a1 = a_ok ? *reinterpret_cast<const float4*>(ap + 4) : z;
  float4 b0 = b_ok0 ? *reinterpret_cast<const float4*>(bp) : z, b1 = b_ok1 ? *reinterpret_cast<const float4*>(bp + 4) : z;
  for (int k0 = 0; k0 < K; k0 += GBK) {
    As[lk + 0][lrow] = a0.x; As[lk + 1][lrow] = a0.y; As[lk + 2][lrow] = a0.z; As[lk + 3][lrow] = a0.w;
    As[lk + 4][lrow] = a1.x; As[lk + 5][lrow] = a1.y; As[lk + 6][lrow] = a1.z; As[lk + 7][lrow] = a1.w;
    *reinterpret_cast<float4*>(&Bs[brow][bcol]) = b0;
    if (BN == 128) *reinterpret_cast<float4*>(&Bs[brow][bcol + 4]) = b1;
    __syncthreads();
    if (k0 + GBK < K) {
      a0 = a_ok ? *reinterpret_cast<const float4*>(ap + k0 + GBK) : z;
      a1 = a_ok ? *reinterpret_cast<const float4*>(ap + k0 + GBK + 4) : z;
      const float* bq = bp + (long long)(k0 + GBK) * ldb;
      b0 = b_ok0 ? *reinterpret_cast<const float4*>(bq) : z;
      b1 = b_ok1 ? *reinterpret_cast<const float4*>(bq + 4) : z;
    }
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float4 ra0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 ra1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      float4 rb0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float4 rb1 = BN == 128 ? *reinterpret_cast<const float4*>(&Bs[k][(BN == 128 ? 64 : 0) + tx * 4]) : z;
      float ra[8] = {ra0.x, ra0.y, ra0.z, ra0.w, ra1.x, ra1.y, ra1.z, ra1.w};
      float rb[8] = {rb0.x, rb0.y, rb0.z, rb0.w, rb1.x, rb1.y, rb1.z, rb1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(ra[i], rb[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool vec = (ldc & 3) == 0 && (!mask || (ldm & 3) == 0);
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
        if (mask) {
          const float4 mk = *reinterpret_cast<const float4*>(mask + (long long)row * ldm + col0);
          v[0] = mk.x > 0.f ? v[0] * mask_scale : 0.f; v[1] = mk.y > 0.f ? v[1] * mask_scale : 0.f;
          v[2] = mk.z > 0.f ? v[2] * mask_scale : 0.f; v[3] = mk.w > 0.f ? v[3] * mask_scale : 0.f;
        }
        if (accum) { const float4 c4 = *reinterpret_cast<const float4*>(cp); v[0] += c4.x; v[1] += c4.y; v[2] += c4.z; v[3] += c4.w; }
        *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (col0 + j >= N) continue;
          float w = v[j];
          if (mask) w = (mask[(long long)row * ldm + col0 + j] > 0.f) ? w * mask_scale : 0.f;
          if (accum) w += cp[j];
          cp[j] = w;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// dW[N,K] += dY[M,N]^T * X[M,K],  db[N] += colsum(dY)      (weight gradient of y = x W^T + b)
// grid (ceil(N/64), ceil(K/64), splits): every CTA reduces its slice of the M rows into a 64 x 64 tile.  256 threads = four
// row groups of 64 threads; a group owns every fourth 16-row slab of the stage and holds the whole tile as 8 x 8 per thread
// (64 FMAs per four 16-byte shared-memory reads).  Stages of 64 rows arrive through cp.async, double-buffered; at the end the
// four groups' tiles are summed through shared memory and ONE set of fp32 atomics per CTA goes to dW (the M split is what
// fills the chip -- dW itself is only 1..6 tiles -- so the atomics are the serial part and are kept four times rarer than
// the accumulating threads).  N, K arbitrary (guards); ldy % 4 == 0 and ldx % 4 == 0.
// ------------------------------------------------------------------------------------------------------------
constexpr int DWT = 64, DWR = 64, DWTHREADS = 256;
constexpr int DW_SMEM = 2 * 2 * DWR * DWT * 4;            // two stages x (dY | X) x 64 rows x 64 floats = 64 KB

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(DWTHREADS, 2) dw_gemm_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ X, int ldx, float* __restrict__ dW,
                                                            int ldw, float* __restrict__ db, long long M, int N, int K, long long rows_per_split) {
  extern __shared__ __align__(16) float dw_smem[];
  auto Ys = [&](int buf, int r) { return dw_smem + ((size_t)(buf * 2) * DWR + r) * DWT; };
  auto Xs = [&](int buf, int r) { return dw_smem + ((size_t)(buf * 2 + 1) * DWR + r) * DWT; };
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * DWT, k0 = blockIdx.y * DWT;
  const long long m_begin = (long long)blockIdx.z * rows_per_split;
  const long long m_end = (m_begin + rows_per_split < M) ? m_begin + rows_per_split : M;
  const int lr = tid >> 4, lc = (tid & 15) * 4;           // loader: rows lr + 16 h, 4 columns
  const int grp = tid >> 6, t64 = tid & 63;
  const int tn = t64 >> 3, tk = t64 & 7;                  // compute: n = tn*4.. and 32 + tn*4.., k = tk*4.. and 32 + tk*4..
  const bool y_full = n0 + lc + 3 < N, x_full = k0 + lc + 3 < K;
  float acc[8][8], bsum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    bsum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  }
  auto load_stage = [&](int buf, long long m) {
#pragma unroll
    for (int h = 0; h < DWR / 16; ++h) {
      const long long row = m + lr + 16 * h;
      float* ys = Ys(buf, lr + 16 * h) + lc;
      float* xs = Xs(buf, lr + 16 * h) + lc;
      const float* yp = dY + row * ldy + n0 + lc;
      const float* xp = X + row * ldx + k0 + lc;
      if (row < m_end && y_full) cp_async16(ys, yp);
      else {
        float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < m_end) { if (n0 + lc < N) y.x = yp[0]; if (n0 + lc + 1 < N) y.y = yp[1]; if (n0 + lc + 2 < N) y.z = yp[2]; }
        *reinterpret_cast<float4*>(ys) = y;
      }
      if (row < m_end && x_full) cp_async16(xs, xp);
      else {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < m_end) { if (k0 + lc < K) x.x = xp[0]; if (k0 + lc + 1 < K) x.y = xp[1]; if (k0 + lc + 2 < K) x.z = xp[2]; }
        *reinterpret_cast<float4*>(xs) = x;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_stage(0, m_begin);
  int buf = 0;
  for (long long m = m_begin; m < m_end; m += DWR, buf ^= 1) {
    if (m + DWR < m_end) { load_stage(buf ^ 1, m + DWR); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < DWR / 4; ++rr) {
      const int r = rr * 4 + grp;                          // rows interleaved over the groups: the tail stage stays balanced
      const float* yr = Ys(buf, r);
      const float* xr = Xs(buf, r);
      const float4 a0 = *reinterpret_cast<const float4*>(yr + tn * 4), a1 = *reinterpret_cast<const float4*>(yr + 32 + tn * 4);
      const float4 b0 = *reinterpret_cast<const float4*>(xr + tk * 4), b1 = *reinterpret_cast<const float4*>(xr + 32 + tk * 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        bsum[i] += av[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  // sum the four groups' tiles: groups 1..3 park theirs in shared memory ([grp - 1][64 values][64 threads]: conflict-free)
  float* red = dw_smem;
  if (grp > 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[((size_t)(grp - 1) * 72 + i * 8 + j) * 64 + t64] = acc[i][j];
      red[((size_t)(grp - 1) * 72 + 64 + i) * 64 + t64] = bsum[i];
    }
  }
  __syncthreads();
  if (grp > 0) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int g = 0; g < 3; ++g) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] += red[((size_t)g * 72 + i * 8 + j) * 64 + t64];
      bsum[i] += red[((size_t)g * 72 + 64 + i) * 64 + t64];
    }
    const int n = n0 + (i < 4 ? tn * 4 + i : 32 + tn * 4 + i - 4);
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + (j < 4 ? tk * 4 + j : 32 + tk * 4 + j - 4);
      if (k < K) atomicAdd(dW + (long long)n * ldw + k, acc[i][j]);
    }
    if (db && blockIdx.y == 0 && tk == 0) atomicAdd(db + n, bsum[i]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm backward (y = (s - mean) * rstd * g + b, s = the saved pre-LayerNorm sum):
//   ds = rstd * (dy g - mean(dy g) - xhat * mean(dy g xhat)),  dg += sum dy xhat,  db += sum dy.   One warp per row,
// 8 rows per warp; the block's column partials go through shared memory and one atomicAdd per column.
// dy and ds may alias.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ s, const float* __restrict__ g, int H,
                                                     long long rows, float* __restrict__ ds, float* __restrict__ dg, float* __restrict__ db) {
  __shared__ float s_dg[8][256], s_db[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = H >> 5;
  float pg[8], pb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { pg[i] = 0.f; pb[i] = 0.f; }
  const long long row0 = ((long long)blockIdx.x * 8 + warp) * 8;
  for (int rr = 0; rr < 8; ++rr) {
    const long long row = row0 + rr;
    if (row >= rows) break;
    float v[8], d[8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < per) { v[i] = s[row * H + lane + 32 * i]; d[i] = dy[row * H + lane + 32 * i]; sum += v[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < per) { float t = v[i] - mean; q = fmaf(t, t, q); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)H + 1e-5f);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < per) {
        const float xh = (v[i] - mean) * rstd;
        const float dxh = d[i] * g[lane + 32 * i];
        pg[i] = fmaf(d[i], xh, pg[i]);
        pb[i] += d[i];
        v[i] = xh; d[i] = dxh;
        m1 += dxh; m2 = fmaf(dxh, xh, m2);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o); }
    m1 /= (float)H; m2 /= (float)H;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < per) ds[row * H + lane + 32 * i] = rstd * (d[i] - m1 - v[i] * m2);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < per) { s_dg[warp][lane + 32 * i] = pg[i]; s_db[warp][lane + 32 * i] = pb[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += s_dg[w][c]; b += s_db[w][c]; }
    atomicAdd(dg + c, a);
    atomicAdd(db + c, b);
  }
}

// ------------------------------------------------------------------------------------------------------------
// The two LayerNorm kernels of the training step for H = 64 / 128 / 256, eight columns per lane: a row takes LPR = H / 8 lanes (two 16-byte
// loads per lane and tensor), a warp works on 32 / LPR rows at a time -- for the reduced model four rows per warp instruction where the
// one-row-per-warp kernels (add_ln_f32_kernel, ln_bwd_kernel) moved 8 bytes per lane and ran at 1.5-2.2 TB/s.  Same arithmetic, same
// dropout indexing (element row * H + col).
// ------------------------------------------------------------------------------------------------------------
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// y = LayerNorm(x + r) * g + b (see add_ln_f32_kernel); grid-stride over groups of 8 * (32 / LPR) rows per CTA
template <int LPR>
__global__ void __launch_bounds__(256) add_ln_v8_kernel(const float* __restrict__ x, const float* __restrict__ r, long long r_rows, const float* __restrict__ g,
                                                        const float* __restrict__ b, long long rows, float* __restrict__ y, float* __restrict__ sum_out, Drop drop,
                                                        bool drop_x) {
  constexpr int H = LPR * 8, RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane / LPR, c0 = (lane % LPR) * 8;
  float gg[8], bb[8];
  ld8(g + c0, gg);
  ld8(b + c0, bb);
  const long long stride = (long long)gridDim.x * 8 * RPW;
  const long long n_iter = (rows + stride - 1) / stride;     // (whole warps stay in the loop for the shuffles)
  for (long long it = 0; it < n_iter; ++it) {
    const long long row = it * stride + ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + sub;
    const bool live = row < rows;
    float v[8], rv[8];
    if (live) {
      ld8(x + row * H + c0, v);
      ld8(r + (row % r_rows) * H + c0, rv);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i] = 0.f; rv[i] = 0.f; }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (drop.thresh) {
        const bool keep = drop_keep(drop, (unsigned long long)(row * H + c0 + i));
        if (drop_x) v[i] = keep ? v[i] * drop.scale : 0.f; else rv[i] = keep ? rv[i] * drop.scale : 0.f;
      }
      v[i] += rv[i];
      s += v[i];
    }
    if (live && sum_out) st8(sum_out + row * H + c0, v);
    const float mean = group_sum<LPR>(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(group_sum<LPR>(q) / (float)H + 1e-5f);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * gg[i] + bb[i];
    if (live) st8(y + row * H + c0, v);
  }
}

// LayerNorm backward (see ln_bwd_kernel): ds, and dg / db through per-lane column partials -> shared memory -> one atomicAdd per column and CTA
template <int LPR>
__global__ void __launch_bounds__(256) ln_bwd_v8_kernel(const float* __restrict__ dy, const float* __restrict__ s, const float* __restrict__ g, long long rows,
                                                        float* __restrict__ ds, float* __restrict__ dg, float* __restrict__ db) {
  constexpr int H = LPR * 8, RPW = 32 / LPR;
  __shared__ float s_dg[8][H], s_db[8][H];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane / LPR, c0 = (lane % LPR) * 8;
  float gg[8], pg[8], pb[8];
  ld8(g + c0, gg);
#pragma unroll
  for (int i = 0; i < 8; ++i) { pg[i] = 0.f; pb[i] = 0.f; }
  const long long stride = (long long)gridDim.x * 8 * RPW;
  const long long n_iter = (rows + stride - 1) / stride;
  for (long long it = 0; it < n_iter; ++it) {
    const long long row = it * stride + ((long long)blockIdx.x * 8 + warp) * RPW + sub;
    const bool live = row < rows;
    float v[8], d[8];
    if (live) {
      ld8(s + row * H + c0, v);
      ld8(dy + row * H + c0, d);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i] = 0.f; d[i] = 0.f; }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += v[i];
    const float mean = group_sum<LPR>(sum) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float t = v[i] - mean; q = fmaf(t, t, q); }
    const float rstd = rsqrtf(group_sum<LPR>(q) / (float)H + 1e-5f);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float xh = (v[i] - mean) * rstd, dxh = d[i] * gg[i];
      pg[i] = fmaf(d[i], xh, pg[i]);
      pb[i] += d[i];
      v[i] = xh; d[i] = dxh;
      m1 += dxh; m2 = fmaf(dxh, xh, m2);
    }
    m1 = group_sum<LPR>(m1) / (float)H;
    m2 = group_sum<LPR>(m2) / (float)H;
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = rstd * (d[i] - m1 - v[i] * m2);
    if (live) st8(ds + row * H + c0, d);
  }
  // the RPW row groups of the warp hold partials for the same columns: fold them, then the eight warps through shared memory
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) { pg[i] += __shfl_xor_sync(0xffffffffu, pg[i], o); pb[i] += __shfl_xor_sync(0xffffffffu, pb[i], o); }
  }
  if (sub == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { s_dg[warp][c0 + i] = pg[i]; s_db[warp][c0 + i] = pb[i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float a = 0.f, bsum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += s_dg[w][c]; bsum += s_db[w][c]; }
    atomicAdd(dg + c, a);
    atomicAdd(db + c, bsum);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Attention backward for one (sequence, head), scores recomputed from Q, K and the saved row log-sum-exp:
//   P = exp(Q K^T c - lse),  dP = dO V^T,  D_i = dO_i . O_i,  dS = P (dP - D) c,  dQ = dS K,  dK = dS^T Q,  dV = P^T dO.
// Pass 1 (one thread per query row): dQ and D.  Pass 2 (one thread per key row): dK and / or dV, no atomics.
// ------------------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(256) attn_bwd_dq_kernel(const float* __restrict__ Q, int ldq, long long q_seq_stride, const float* __restrict__ Kp,
                                                          const float* __restrict__ Vp, int ldkv, const float* __restrict__ dO, const float* __restrict__ O,
                                                          int ldo, const float* __restrict__ lse, int Lq, int Lk, int heads, float c,
                                                          float* __restrict__ dQ, int lddq, float* __restrict__ Dout, Drop drop) {
  extern __shared__ __align__(16) float smem_bwd[];
  float* sK = smem_bwd;
  float* sV = smem_bwd + (size_t)Lk * DH;
  const int seq = blockIdx.x, head = blockIdx.y;
  const float* kbase = Kp + (long long)seq * Lk * ldkv + head * DH;
  const float* vbase = Vp + (long long)seq * Lk * ldkv + head * DH;
  for (int i = threadIdx.x; i < Lk * (DH / 4); i += blockDim.x) {
    int j = i / (DH / 4), cc = i % (DH / 4);
    reinterpret_cast<float4*>(sK)[i] = *reinterpret_cast<const float4*>(kbase + (long long)j * ldkv + cc * 4);
    reinterpret_cast<float4*>(sV)[i] = *reinterpret_cast<const float4*>(vbase + (long long)j * ldkv + cc * 4);
  }
  __syncthreads();
  const int r = threadIdx.x;
  if (r >= Lq) return;
  float q[DH], go[DH], acc[DH];
  const float* qp = Q + (long long)seq * q_seq_stride + (long long)r * ldq + head * DH;
  const float* gp = dO + ((long long)seq * Lq + r) * ldo + head * DH;
  const float* op = O + ((long long)seq * Lq + r) * ldo + head * DH;
  float D = 0.f;
#pragma unroll
  for (int i = 0; i < DH; ++i) { q[i] = qp[i]; go[i] = gp[i]; D = fmaf(go[i], op[i], D); acc[i] = 0.f; }
  const float l = lse[((long long)seq * heads + head) * Lq + r];
  for (int j = 0; j < Lk; ++j) {
    const float* kr = sK + (size_t)j * DH;
    const float* vr = sV + (size_t)j * DH;
    float s = 0.f, dp = 0.f;
#pragma unroll
    for (int i = 0; i < DH; ++i) { s = fmaf(q[i], kr[i], s); dp = fmaf(go[i], vr[i], dp); }
    if (drop.thresh) dp = drop_keep(drop, (unsigned long long)((((long long)seq * heads + head) * Lq + r) * Lk + j)) ? dp * drop.scale : 0.f;
    const float ds = expf(s * c - l) * (dp - D) * c;
#pragma unroll
    for (int i = 0; i < DH; ++i) acc[i] = fmaf(ds, kr[i], acc[i]);
  }
  float* out = dQ + ((long long)seq * Lq + r) * lddq + head * DH;
#pragma unroll
  for (int i = 0; i < DH; ++i) out[i] = acc[i];
  Dout[((long long)seq * heads + head) * Lq + r] = D;
}

// Pass 1 with two query rows per thread (head_dim 32): every K / V value read from shared memory feeds both rows, which halves
// the shared-memory read traffic that bounds the one-row kernel.
template <int DH>
__global__ void __launch_bounds__(128) attn_bwd_dq_r2_kernel(const float* __restrict__ Q, int ldq, long long q_seq_stride, const float* __restrict__ Kp,
                                                             const float* __restrict__ Vp, int ldkv, const float* __restrict__ dO, const float* __restrict__ O,
                                                             int ldo, const float* __restrict__ lse, int Lq, int Lk, int heads, float c,
                                                             float* __restrict__ dQ, int lddq, float* __restrict__ Dout, Drop drop) {
  extern __shared__ __align__(16) float smem_bwd[];
  float* sK = smem_bwd;
  float* sV = smem_bwd + (size_t)Lk * DH;
  const int seq = blockIdx.x, head = blockIdx.y;
  const float* kbase = Kp + (long long)seq * Lk * ldkv + head * DH;
  const float* vbase = Vp + (long long)seq * Lk * ldkv + head * DH;
  for (int i = threadIdx.x; i < Lk * (DH / 4); i += blockDim.x) {
    int j = i / (DH / 4), cc = i % (DH / 4);
    reinterpret_cast<float4*>(sK)[i] = *reinterpret_cast<const float4*>(kbase + (long long)j * ldkv + cc * 4);
    reinterpret_cast<float4*>(sV)[i] = *reinterpret_cast<const float4*>(vbase + (long long)j * ldkv + cc * 4);
  }
  __syncthreads();
  const int r0 = 2 * threadIdx.x;
  if (r0 >= Lq) return;
  const bool two = r0 + 1 < Lq;
  float q[2][DH], go[2][DH], acc[2][DH];
  float D[2] = {0.f, 0.f}, l[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int r = r0 + (two ? u : 0);
    const float* qp = Q + (long long)seq * q_seq_stride + (long long)r * ldq + head * DH;
    const float* gp = dO + ((long long)seq * Lq + r) * ldo + head * DH;
    const float* op = O + ((long long)seq * Lq + r) * ldo + head * DH;
#pragma unroll
    for (int i = 0; i < DH; ++i) { q[u][i] = qp[i]; go[u][i] = gp[i]; D[u] = fmaf(go[u][i], op[i], D[u]); acc[u][i] = 0.f; }
    l[u] = lse[((long long)seq * heads + head) * Lq + r];
  }
  const long long pbase = (((long long)seq * heads + head) * Lq + r0) * Lk;
  for (int j = 0; j < Lk; ++j) {
    const float4* kr = reinterpret_cast<const float4*>(sK + (size_t)j * DH);
    const float4* vr = reinterpret_cast<const float4*>(sV + (size_t)j * DH);
    float s[2] = {0.f, 0.f}, dp[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < DH / 4; ++i) {
      const float4 k4 = kr[i], v4 = vr[i];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        s[u] = fmaf(q[u][4 * i], k4.x, s[u]); s[u] = fmaf(q[u][4 * i + 1], k4.y, s[u]); s[u] = fmaf(q[u][4 * i + 2], k4.z, s[u]); s[u] = fmaf(q[u][4 * i + 3], k4.w, s[u]);
        dp[u] = fmaf(go[u][4 * i], v4.x, dp[u]); dp[u] = fmaf(go[u][4 * i + 1], v4.y, dp[u]);
        dp[u] = fmaf(go[u][4 * i + 2], v4.z, dp[u]); dp[u] = fmaf(go[u][4 * i + 3], v4.w, dp[u]);
      }
    }
    float ds[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (drop.thresh) dp[u] = drop_keep(drop, (unsigned long long)(pbase + (long long)u * Lk + j)) ? dp[u] * drop.scale : 0.f;
      ds[u] = expf(s[u] * c - l[u]) * (dp[u] - D[u]) * c;
    }
#pragma unroll
    for (int i = 0; i < DH / 4; ++i) {
      const float4 k4 = kr[i];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        acc[u][4 * i] = fmaf(ds[u], k4.x, acc[u][4 * i]); acc[u][4 * i + 1] = fmaf(ds[u], k4.y, acc[u][4 * i + 1]);
        acc[u][4 * i + 2] = fmaf(ds[u], k4.z, acc[u][4 * i + 2]); acc[u][4 * i + 3] = fmaf(ds[u], k4.w, acc[u][4 * i + 3]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (u == 1 && !two) break;
    float* out = dQ + ((long long)seq * Lq + r0 + u) * lddq + head * DH;
#pragma unroll
    for (int i = 0; i < DH / 4; ++i) *reinterpret_cast<float4*>(out + 4 * i) = make_float4(acc[u][4 * i], acc[u][4 * i + 1], acc[u][4 * i + 2], acc[u][4 * i + 3]);
    Dout[((long long)seq * heads + head) * Lq + r0 + u] = D[u];
  }
}

// MODE 0: dK and dV, 1: dK only, 2: dV only (head_dim 64 needs two passes to stay inside the register file)
template <int DH, int MODE>
__global__ void __launch_bounds__(256) attn_bwd_dkv_kernel(const float* __restrict__ Q, int ldq, long long q_seq_stride, const float* __restrict__ Kp,
                                                           const float* __restrict__ Vp, int ldkv, const float* __restrict__ dO, int ldo,
                                                           const float* __restrict__ lse, const float* __restrict__ Din, int Lq, int Lk, int heads, float c,
                                                           float* __restrict__ dK, float* __restrict__ dV, int lddkv, Drop drop) {
  extern __shared__ __align__(16) float smem_bwd[];
  float* sQ = smem_bwd;
  float* sG = smem_bwd + (size_t)Lq * DH;
  float* sL = sG + (size_t)Lq * DH;
  float* sD = sL + Lq;
  const int seq = blockIdx.x, head = blockIdx.y;
  for (int i = threadIdx.x; i < Lq * (DH / 4); i += blockDim.x) {
    int j = i / (DH / 4), cc = i % (DH / 4);
    reinterpret_cast<float4*>(sQ)[i] = *reinterpret_cast<const float4*>(Q + (long long)seq * q_seq_stride + (long long)j * ldq + head * DH + cc * 4);
    reinterpret_cast<float4*>(sG)[i] = *reinterpret_cast<const float4*>(dO + ((long long)seq * Lq + j) * ldo + head * DH + cc * 4);
  }
  for (int i = threadIdx.x; i < Lq; i += blockDim.x) {
    sL[i] = lse[((long long)seq * heads + head) * Lq + i];
    sD[i] = Din[((long long)seq * heads + head) * Lq + i];
  }
  __syncthreads();
  const int j = threadIdx.x;
  if (j >= Lk) return;
  float k[DH], v[MODE == 2 ? 1 : DH], ak[MODE == 2 ? 1 : DH], av[MODE == 1 ? 1 : DH];
  const float* kp = Kp + ((long long)seq * Lk + j) * ldkv + head * DH;
  const float* vp = Vp + ((long long)seq * Lk + j) * ldkv + head * DH;
#pragma unroll
  for (int i = 0; i < DH; ++i) {
    k[i] = kp[i];
    if (MODE != 2) { v[i] = vp[i]; ak[i] = 0.f; }
    if (MODE != 1) av[i] = 0.f;
  }
  for (int r = 0; r < Lq; ++r) {
    const float* qr = sQ + (size_t)r * DH;
    const float* gr = sG + (size_t)r * DH;
    float s = 0.f, dp = 0.f;
#pragma unroll
    for (int i = 0; i < DH; ++i) {
      s = fmaf(qr[i], k[i], s);
      if (MODE != 2) dp = fmaf(gr[i], v[i], dp);
    }
    const float p = expf(s * c - sL[r]);
    float keep_scale = 1.f;                              // dropout on the probabilities: O = (P . mask * scale) V
    if (drop.thresh) keep_scale = drop_keep(drop, (unsigned long long)((((long long)seq * heads + head) * Lq + r) * Lk + j)) ? drop.scale : 0.f;
    if (MODE != 1) {
      const float pd = p * keep_scale;
#pragma unroll
      for (int i = 0; i < DH; ++i) av[i] = fmaf(pd, gr[i], av[i]);
    }
    if (MODE != 2) {
      dp *= keep_scale;
      const float ds = p * (dp - sD[r]) * c;
#pragma unroll
      for (int i = 0; i < DH; ++i) ak[i] = fmaf(ds, qr[i], ak[i]);
    }
  }
  if (MODE != 2) {
    float* o = dK + ((long long)seq * Lk + j) * lddkv + head * DH;
#pragma unroll
    for (int i = 0; i < DH; ++i) o[i] = ak[i];
  }
  if (MODE != 1) {
    float* o = dV + ((long long)seq * Lk + j) * lddkv + head * DH;
#pragma unroll
    for (int i = 0; i < DH; ++i) o[i] = av[i];
  }
}

// Pass 2 for head_dim 32, register-tiled against the shared-memory read path that bounds the kernel above (every Q / dO value
// read there feeds ONE key): a thread PAIR owns two key rows, each thread one half of the head dimension (16 columns of K, V,
// dK, dV for both keys), the two half dot products meet through one shuffle.  Per query a thread reads 32 floats for 128 FMAs
// (the one-key kernel: 64 floats for 128 FMAs).
__global__ void __launch_bounds__(256) attn_bwd_dkv_pair_kernel(const float* __restrict__ Q, int ldq, long long q_seq_stride, const float* __restrict__ Kp,
                                                                const float* __restrict__ Vp, int ldkv, const float* __restrict__ dO, int ldo,
                                                                const float* __restrict__ lse, const float* __restrict__ Din, int Lq, int Lk, int heads, float c,
                                                                float* __restrict__ dK, float* __restrict__ dV, int lddkv, Drop drop) {
  constexpr int DH = 32, HD = 16;
  extern __shared__ __align__(16) float smem_bwd[];
  float* sQ = smem_bwd;
  float* sG = smem_bwd + (size_t)Lq * DH;
  float* sL = sG + (size_t)Lq * DH;
  float* sD = sL + Lq;
  const int seq = blockIdx.x, head = blockIdx.y;
  for (int i = threadIdx.x; i < Lq * (DH / 4); i += blockDim.x) {
    int j = i / (DH / 4), cc = i % (DH / 4);
    reinterpret_cast<float4*>(sQ)[i] = *reinterpret_cast<const float4*>(Q + (long long)seq * q_seq_stride + (long long)j * ldq + head * DH + cc * 4);
    reinterpret_cast<float4*>(sG)[i] = *reinterpret_cast<const float4*>(dO + ((long long)seq * Lq + j) * ldo + head * DH + cc * 4);
  }
  for (int i = threadIdx.x; i < Lq; i += blockDim.x) {
    sL[i] = lse[((long long)seq * heads + head) * Lq + i];
    sD[i] = Din[((long long)seq * heads + head) * Lq + i];
  }
  __syncthreads();
  const int half = threadIdx.x & 1;
  const int j0 = 2 * (threadIdx.x >> 1);                   // keys j0, j0 + 1 (whole warps stay alive for the shuffles)
  const bool ok0 = j0 < Lk, ok1 = j0 + 1 < Lk;
  float k[2][HD], v[2][HD], ak[2][HD], av[2][HD];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const bool ok = u ? ok1 : ok0;
    const float* kp = Kp + ((long long)seq * Lk + j0 + u) * ldkv + head * DH + half * HD;
    const float* vp = Vp + ((long long)seq * Lk + j0 + u) * ldkv + head * DH + half * HD;
#pragma unroll
    for (int i = 0; i < HD; ++i) { k[u][i] = ok ? kp[i] : 0.f; v[u][i] = ok ? vp[i] : 0.f; ak[u][i] = 0.f; av[u][i] = 0.f; }
  }
  const long long pbase = ((long long)seq * heads + head) * Lq;
  for (int r = 0; r < Lq; ++r) {
    const float4* qr = reinterpret_cast<const float4*>(sQ + (size_t)r * DH + half * HD);
    const float4* gr = reinterpret_cast<const float4*>(sG + (size_t)r * DH + half * HD);
    float q[HD], g[HD];
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) {
      const float4 a = qr[i], b = gr[i];
      q[4 * i] = a.x; q[4 * i + 1] = a.y; q[4 * i + 2] = a.z; q[4 * i + 3] = a.w;
      g[4 * i] = b.x; g[4 * i + 1] = b.y; g[4 * i + 2] = b.z; g[4 * i + 3] = b.w;
    }
    float s[2] = {0.f, 0.f}, dp[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < HD; ++i) {
      s[0] = fmaf(q[i], k[0][i], s[0]); s[1] = fmaf(q[i], k[1][i], s[1]);
      dp[0] = fmaf(g[i], v[0][i], dp[0]); dp[1] = fmaf(g[i], v[1][i], dp[1]);
    }
    const float Lr = sL[r], Dr = sD[r];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      s[u] += __shfl_xor_sync(0xffffffffu, s[u], 1);
      dp[u] += __shfl_xor_sync(0xffffffffu, dp[u], 1);
      const float p = expf(s[u] * c - Lr);
      float keep_scale = 1.f;                              // dropout on the probabilities: O = (P . mask * scale) V
      if (drop.thresh) keep_scale = drop_keep(drop, (unsigned long long)((pbase + r) * Lk + j0 + u)) ? drop.scale : 0.f;
      const float pd = p * keep_scale;
      const float ds = p * (dp[u] * keep_scale - Dr) * c;
#pragma unroll
      for (int i = 0; i < HD; ++i) { av[u][i] = fmaf(pd, g[i], av[u][i]); ak[u][i] = fmaf(ds, q[i], ak[u][i]); }
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (!(u ? ok1 : ok0)) continue;
    float* ok_ = dK + ((long long)seq * Lk + j0 + u) * lddkv + head * DH + half * HD;
    float* ov_ = dV + ((long long)seq * Lk + j0 + u) * lddkv + head * DH + half * HD;
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) {
      *reinterpret_cast<float4*>(ok_ + 4 * i) = make_float4(ak[u][4 * i], ak[u][4 * i + 1], ak[u][4 * i + 2], ak[u][4 * i + 3]);
      *reinterpret_cast<float4*>(ov_ + 4 * i) = make_float4(av[u][4 * i], av[u][4 * i + 1], av[u][4 * i + 2], av[u][4 * i + 3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Loss of one head group (train.py:139-151) and its gradient with respect to the logits.
//   logits [rows, ld]: columns 0..2 onset / offset / mpe (pre-sigmoid), 3..3+V-1 velocity; rows in (b,f,n) order, or
//   (b,n,f) when time_major.  labels are [B,F,NN] in (b,f,n) order.  Every term is a mean over the rows positions:
//   BCELoss (log clamped at -100 like torch) on sigmoid(logit), CrossEntropyLoss on the velocity logits.
//   dlogits = weight * dLoss/dlogit;  loss[0] += weight * sum of the four terms.   One warp per row (V = 128).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ logits, int ld, int V, int F, int NN, long long rows, bool time_major,
                                                        const float* __restrict__ y_on, const float* __restrict__ y_off, const float* __restrict__ y_mpe,
                                                        const long long* __restrict__ y_vel, float weight, float* __restrict__ dlogits,
                                                        float* __restrict__ loss) {
  __shared__ float s_loss[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  float part = 0.f;
  if (row < rows) {
    long long lrow = row;
    if (time_major) {
      int f = (int)(row % F);
      long long bn = row / F;
      int n = (int)(bn % NN);
      lrow = ((bn / NN) * F + f) * NN + n;
    }
    const float inv_n = 1.f / (float)rows;
    const float* lp = logits + row * ld;
    float* gp = dlogits + row * ld;
    // velocity: V = 128 -> 4 logits per lane
    float v[4];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = lp[3 + lane + 32 * i]; mx = fmaxf(mx, v[i]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) se += expf(v[i] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    const float lsum = mx + logf(se);
    const int lab = (int)y_vel[lrow];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = lane + 32 * i;
      const float p = expf(v[i] - lsum);
      gp[3 + col] = weight * inv_n * (p - (col == lab ? 1.f : 0.f));
      if (col == lab) part += (lsum - v[i]) * inv_n;
    }
    if (lane < 3) {
      const float* yy = lane == 0 ? y_on : (lane == 1 ? y_off : y_mpe);
      const float y = yy[lrow];
      const float z = lp[lane];
      const float p = 1.f / (1.f + expf(-z));
      const float l1 = fmaxf(logf(p), -100.f), l0 = fmaxf(logf(1.f - p), -100.f);
      part += -(y * l1 + (1.f - y) * l0) * inv_n;
      const float pq = p * (1.f - p);
      gp[lane] = weight * inv_n * (p - y) * (pq / fmaxf(pq, 1e-12f));
    }
    for (int col = 3 + V + lane; col < ld; col += 32) gp[col] = 0.f;     // padding columns
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  }
  if (lane == 0) s_loss[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_loss[w];
    atomicAdd(loss, weight * t);
  }
}

// The eight head outputs of a train-mode forward from the packed logits [rows, ld] (columns 0..2 onset / offset / mpe, 3..3+V-1 velocity):
// sigmoid probabilities [B, F, NN] and raw velocity logits [B, F, NN, V] (model_spec2midi.py:172-175, :203-206).  One warp per row.
__global__ void __launch_bounds__(256) heads_out_kernel(const float* __restrict__ logits, int ld, int V, int F, int NN, long long rows, bool time_major,
                                                        float* __restrict__ on, float* __restrict__ off, float* __restrict__ mpe, float* __restrict__ vel) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  long long lrow = row;
  if (time_major) {
    int f = (int)(row % F);
    long long bn = row / F;
    int n = (int)(bn % NN);
    lrow = ((bn / NN) * F + f) * NN + n;
  }
  const float* lp = logits + row * ld;
  if (vel)
    for (int col = lane; col < V; col += 32) vel[lrow * V + col] = lp[3 + col];
  if (lane < 3) {
    float* dst = lane == 0 ? on : (lane == 1 ? off : mpe);
    if (dst) dst[lrow] = 1.f / (1.f + expf(-lp[lane]));
  }
}

// dlogits from the gradients of the eight outputs (autograd's loss.backward() reaching the module outputs): through the sigmoid for the three
// probability outputs (dz = g * p * (1 - p)), identity for the velocity logits.  A NULL gradient pointer means zero.
__global__ void __launch_bounds__(256) heads_outgrad_kernel(const float* __restrict__ logits, int ld, int V, int F, int NN, long long rows, bool time_major,
                                                            const float* __restrict__ g_on, const float* __restrict__ g_off, const float* __restrict__ g_mpe,
                                                            const float* __restrict__ g_vel, float* __restrict__ dlogits) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  long long lrow = row;
  if (time_major) {
    int f = (int)(row % F);
    long long bn = row / F;
    int n = (int)(bn % NN);
    lrow = ((bn / NN) * F + f) * NN + n;
  }
  const float* lp = logits + row * ld;
  float* gp = dlogits + row * ld;
  for (int col = lane; col < V; col += 32) gp[3 + col] = g_vel ? g_vel[lrow * V + col] : 0.f;
  if (lane < 3) {
    const float* gg = lane == 0 ? g_on : (lane == 1 ? g_off : g_mpe);
    const float p = 1.f / (1.f + expf(-lp[lane]));
    gp[lane] = gg ? gg[lrow] * p * (1.f - p) : 0.f;
  }
  for (int col = 3 + V + lane; col < ld; col += 32) gp[col] = 0.f;       // padding columns
}

// out[c] += sum_r in[r, c]   (embedding gradients: the same table row is added to many sequences).  grid (ceil(cols/256), splits)
__global__ void colsum_kernel(const float* __restrict__ in, long long rows, long long cols, long long rows_per_split, float* __restrict__ out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long long r0 = (long long)blockIdx.y * rows_per_split;
  const long long r1 = r0 + rows_per_split < rows ? r0 + rows_per_split : rows;
  float a = 0.f;
  for (long long r = r0; r < r1; ++r) a += in[r * cols + c];
  atomicAdd(out + c, a);
}

// gT[((b*F+f)*NN+n), h] += gU[((b*NN+n)*F+f), h] * scale     (backward of the time re-layout, model_spec2midi.py:189-191)
__global__ void time_relayout_bwd_kernel(const float* __restrict__ gU, float scale, int F, int NN, int H, long long total, float* __restrict__ gT) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int h = (int)(i % H);
  long long r = i / H;
  int f = (int)(r % F);
  long long bn = r / F;
  int n = (int)(bn % NN);
  long long b = bn / NN;
  gT[(((b * F + f) * NN + n)) * H + h] += gU[i] * scale;
}

// Front backward: dWc[h, j] += sum_{b,f} dE[(b,f,bin), h] * spec[b, bin, f + j],  dbc[h] += sum dE   (collapsed 65-tap filter).
// grid (n_bin, ceil(batch / spc)): one CTA per (bin, spc segments);
// threads = G tap groups x H (G * H = 256).  dE = dX * sqrt(H) (the embedding scale).
template <int NPROC, int G>
__global__ void __launch_bounds__(256) front_bwd_kernel(const float* __restrict__ spec, long long sb, long long sbin, long long st, const float* __restrict__ dX,
                                                        float scale, int H, int F, int NB, int B, int spc, float* __restrict__ dWc, float* __restrict__ dbc) {
  // Thread (h, g) owns the taps j = g + G a.  Frames are walked in G phases (f = ph, ph + G, ...): inside a phase the spectrogram values a
  // thread needs slide by ONE tap slot per frame, so a block of 16 frames shares a register window of NA + 15 values (that many shared-memory reads
  // for 16 x NA FMAs instead of one read per FMA) and its 16 dE values are fetched in one batch (16 loads in flight).  F % (16 G) == 0.
  __shared__ float s_row[320];
  const int bin = blockIdx.x;
  const int W = F + NPROC - 1;
  const int h = threadIdx.x % H, g = threadIdx.x / H;
  constexpr int NA = (NPROC + G - 1) / G;
  float acc[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = 0.f;
  float bsum = 0.f;
  const long long fstride = (long long)NB * H;
  for (int b = blockIdx.y * spc; b < B && b < (int)(blockIdx.y + 1) * spc; ++b) {   // spc segments per CTA: every CTA closes with NPROC * H same-address atomics
    __syncthreads();
    for (int i = threadIdx.x; i < 320; i += blockDim.x) s_row[i] = i < W ? spec[b * sb + bin * sbin + i * st] : 0.f;
    __syncthreads();
    const float* dx = dX + (((long long)b * F) * NB + bin) * H + h;
    for (int ph = 0; ph < G; ++ph) {
      for (int f0 = ph; f0 < F; f0 += 16 * G) {
        constexpr int WN = NA + 15;                      // window: taps of the first frame + one more slot per further frame
        float e[16], w[WN];
#pragma unroll
        for (int i = 0; i < 16; ++i) e[i] = __ldg(dx + (long long)(f0 + i * G) * fstride) * scale;
#pragma unroll
        for (int m = 0; m < WN; ++m) w[m] = s_row[f0 + g + G * m];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (g == 0) bsum += e[i];
#pragma unroll
          for (int a2 = 0; a2 < NA; ++a2) acc[a2] = fmaf(e[i], w[i + a2], acc[a2]);
        }
      }
    }
  }
#pragma unroll
  for (int a2 = 0; a2 < NA; ++a2) {
    const int j = g + a2 * G;
    if (j < NPROC) atomicAdd(dWc + h * NPROC + j, acc[a2]);
  }
  if (g == 0) atomicAdd(dbc + h, bsum);
}

// Chain rule through the collapse Wc[h,j] = sum_{c,i} tok_w[h, c*n_out + (j-i)] conv_w[c,i],  bc[h] = tok_b[h] + sum tok_w conv_b:
//   d tok_w[h, c*n_out+k] = sum_i dWc[h,k+i] conv_w[c,i] + dbc[h] conv_b[c];   d conv_w[c,i] = sum_{h,k} dWc[h,k+i] tok_w[h,c*n_out+k];
//   d conv_b[c] = sum_h dbc[h] sum_k tok_w[h,c*n_out+k];   d tok_b = dbc.
// Work items: one THREAD per element of d tok_w and d tok_b, one WARP per element of d conv_w / d conv_b (sums over H * n_out terms: as single
// threads those 24 elements were a 225 us serial tail).  Launch with at least n_tok + H + 32 * (C * kw + C) threads.
__global__ void front_chain_bwd_kernel(const float* __restrict__ dWc, const float* __restrict__ dbc, const float* __restrict__ tok_w,
                                       const float* __restrict__ conv_w, const float* __restrict__ conv_b, int H, int C, int kw, int n_out, int n_proc,
                                       float* __restrict__ g_tok_w, float* __restrict__ g_tok_b, float* __restrict__ g_conv_w, float* __restrict__ g_conv_b) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_tok = H * C * n_out;
  if (idx < n_tok) {
    const int h = idx / (C * n_out), c = (idx / n_out) % C, k = idx % n_out;
    float a = dbc[h] * conv_b[c];
    for (int i = 0; i < kw; ++i) a = fmaf(dWc[h * n_proc + k + i], conv_w[c * kw + i], a);
    g_tok_w[idx] += a;
    return;
  }
  if (idx < n_tok + H) {
    g_tok_b[idx - n_tok] += dbc[idx - n_tok];
    return;
  }
  // warp items (n_tok + H is a multiple of 32 whenever H is, so whole warps arrive here together; the shuffles below need that)
  const int item = (idx - n_tok - H) >> 5, lane = threadIdx.x & 31;
  if (item >= C * kw + C) return;
  float a = 0.f;
  if (item < C * kw) {
    const int c = item / kw, i = item % kw;
    for (int t = lane; t < H * n_out; t += 32) {
      const int h = t / n_out, k = t % n_out;
      a = fmaf(dWc[h * n_proc + k + i], tok_w[(long long)h * C * n_out + c * n_out + k], a);
    }
  } else {
    const int c = item - C * kw;
    for (int t = lane; t < H * n_out; t += 32) {
      const int h = t / n_out, k = t % n_out;
      a = fmaf(dbc[h], tok_w[(long long)h * C * n_out + c * n_out + k], a);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) {
    if (item < C * kw) g_conv_w[item] += a;
    else g_conv_b[item - C * kw] += a;
  }
}

// heads: fused [NP, H] weight / [NP] bias (rows 0..2 onset / offset / mpe, 3..3+V-1 velocity, rest zero padding)
__global__ void pack_heads_f32_kernel(const float* __restrict__ w_on, const float* __restrict__ w_off, const float* __restrict__ w_mpe,
                                      const float* __restrict__ w_vel, const float* __restrict__ b_on, const float* __restrict__ b_off,
                                      const float* __restrict__ b_mpe, const float* __restrict__ b_vel, int V, int H, int NP, float* __restrict__ w,
                                      float* __restrict__ bias) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NP * H) return;
  int r = i / H, h = i % H;
  float v = 0.f, b = 0.f;
  if (r == 0) { v = w_on[h]; b = b_on[0]; }
  else if (r == 1) { v = w_off[h]; b = b_off[0]; }
  else if (r == 2) { v = w_mpe[h]; b = b_mpe[0]; }
  else if (r < 3 + V) { v = w_vel[(r - 3) * H + h]; b = b_vel[r - 3]; }
  w[i] = v;
  if (h == 0) bias[r] = b;
}
__global__ void unpack_heads_grad_kernel(const float* __restrict__ gw, const float* __restrict__ gb, int V, int H, float* __restrict__ g_on,
                                         float* __restrict__ g_off, float* __restrict__ g_mpe, float* __restrict__ g_vel, float* __restrict__ gb_on,
                                         float* __restrict__ gb_off, float* __restrict__ gb_mpe, float* __restrict__ gb_vel) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (3 + V) * H) return;
  int r = i / H, h = i % H;
  float* dst = r == 0 ? g_on + h : (r == 1 ? g_off + h : (r == 2 ? g_mpe + h : g_vel + (long long)(r - 3) * H + h));
  *dst += gw[i];
  if (h == 0) {
    float* bd = r == 0 ? gb_on : (r == 1 ? gb_off : (r == 2 ? gb_mpe : gb_vel + (r - 3)));
    *bd += gb[r];
  }
}

// torch.optim.Adam (no weight decay, no amsgrad), m_training.py:146:  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).   g is multiplied by grad_scale first (1 / world size after the all-reduce).
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                            float b1, float b2, float eps, float bc1, float bc2_sqrt, float grad_scale) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * grad_scale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
}

}  // namespace
}  // namespace hft
