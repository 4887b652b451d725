// tcgen05 attention for the TRAINING step, head_dim 32 (reduced model, BASELINE configs[4]): forward with the row
// log-sum-exp, and the backward that recomputes the probabilities (reference MultiHeadAttentionLayer.forward,
// model_spec2midi.py:342-348, under loss.backward(), training/train.py:158).  Same arguments as the fp32 CUDA-core
// kernels they replace (attn_f32_r2_kernel, attn_bwd_dq_r2_kernel, attn_bwd_dkv_pair_kernel): fp32 Q | K | V, dO, O in
// global memory, fp32 results.
//
// Arithmetic: every product runs on the tensor cores as THREE fp16 MMAs with fp32 accumulation in TMEM
// (a.b = a_hi.b_hi + a_lo.b_hi + a_hi.b_lo, hi = fp16(x), lo = fp16(x - hi): 22 mantissa bits), exp2 / dropout / the
// dS formula in fp32 on the CUDA cores.  Gradients are tiny (loss means over B*128*88 positions), so dO is scaled by a
// power of two per CTA (its largest magnitude lands in [1, 2)) before the split and the results are scaled back: every
// output is linear in dO, powers of two are exact.
//
// Structure (one CTA = 256 threads = 128 TMEM lanes x 2 column halves, two CTAs per SM, 256 TMEM columns each):
//   operands are staged by the threads themselves: raw fp32 rows arrive by cp.async (every tensor of the CTA in flight at once) and
//   are converted in place to hi | lo fp16 tiles in the K-major 64-byte-swizzled UMMA layout (a [rows, 32] tile serves as A / B
//   K-major for the score-type products and as B MN-major for the accumulate-type products, so nothing is transposed);
//   forward   (seq, head, 128 queries): per 128-key unit S = Q K^T -> each thread owns 64 columns of its row with its own
//             max / sum (split softmax), writes P (hi | lo, dropout applied) in place over S, O_{unit,half} = P V with A from
//             TMEM into its own 32 columns; the partial results are merged in the epilogue, lse = max + ln(sum);
//   dQ kernel (seq, head, 128 queries), thread = query row: per 64-key chunk S = Q K^T, dP = dO V^T -> dS = P (dP - D) c
//             in place -> dQ += dS K (A from TMEM); also writes D = rowsum(dO . O);
//   dK/dV kernel (seq, head, 128 keys), thread = key row: per 64-query chunk S^T = K Q^T, dP^T = V dO^T -> P^T (dropped) and
//             dS^T in place -> dV += P^T dO, dK += dS^T Q.
//   The transposed pass recomputes S and dP on the tensor cores instead of transposing dS through shared memory: the
//   score-type products have K = 32 and cost a fifth of the accumulate-type ones.
#pragma once
#include "tc_common.cuh"
#include "tc_attn.cuh"          // ex2_approx
#include "f32_kernels.cuh"      // Drop, drop_keep

namespace hft {
namespace tc {

struct TAttnArgs {
  const float* Q; int ldq; long long q_seq_stride;    // q_seq_stride 0: the same queries for every sequence (decoder layer zero)
  const float* K; const float* V; int ldkv;
  int Lq, Lk, heads;
  float c;                                            // 1 / sqrt(head_dim)
  float* ctx;                                         // forward out / backward in (O): [S * Lq, ldo]
  int ldo;
  float* lse;                                         // [S, heads, Lq]  (forward out / backward in)
  const float* dO;                                    // [S * Lq, ldo]
  float* dQ; int lddq;
  float* dK; float* dV; int lddkv;
  float* Dbuf;                                        // [S, heads, Lq]: written by the dQ kernel, read by the dK/dV kernel
  Drop drop;
};

constexpr int kTThreads = 256;
constexpr uint32_t kTAtom = 512;                      // 8 rows x 64 bytes: one 64B-swizzle atom of a [rows, 32] fp16 tile
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

// Staging: every operand tile is fetched ONCE, by cp.async, as raw fp32 rows into the shared-memory region its hi | lo fp16 tiles will
// occupy (32 floats = 128 bytes per row = the 64 + 64 bytes of the two tiles), all tensors in flight together (one global round trip per
// CTA instead of one per tensor); the conversion then runs in place: every thread reads its 8-float chunks into registers, the CTA
// synchronises, and the chunks are written back split, in the K-major 64B-swizzled UMMA layout (row r at r * 64, 16-byte chunk ch at
// position ch ^ ((r >> 1) & 3); hi tile at the region's start, lo tile n_pad * 64 bytes further on).
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
// rows [0, n_valid) x 32 floats from src (row pitch ld) -> raw rows at region + r * 128; rows up to n_pad (a multiple of 64) are zero
__device__ __forceinline__ void traw_issue(uint8_t* region, const float* src, long long ld, int n_valid, int n_pad) {
  for (int i = threadIdx.x; i < n_pad * 8; i += kTThreads) {
    const int r = i >> 3, c = i & 7;
    if (r < n_valid) cp_async_16(region + r * 128 + c * 16, src + (long long)r * ld + c * 4);
    else *reinterpret_cast<uint4*>(region + r * 128 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
}
// this thread's chunks of a raw region: task t = chunk (t * 256 + tid) -> row = chunk >> 2, 8 floats at column (chunk & 3) * 8.  T >= n_pad / 64.
template <int T>
struct TRaw {
  float4 a[T], b[T];
  __device__ __forceinline__ void read(const uint8_t* region, int n_pad) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int i = t * kTThreads + threadIdx.x;
      if (i < n_pad * 4) {
        const float4* s4 = reinterpret_cast<const float4*>(region + (i >> 2) * 128 + (i & 3) * 32);
        a[t] = s4[0];
        b[t] = s4[1];
      }
    }
  }
  __device__ __forceinline__ float max_abs(int n_pad) const {
    float mx = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t)
      if (t * kTThreads + (int)threadIdx.x < n_pad * 4)
        mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(fabsf(a[t].x), fabsf(a[t].y)), fmaxf(fabsf(a[t].z), fabsf(a[t].w))),
                             fmaxf(fmaxf(fabsf(b[t].x), fabsf(b[t].y)), fmaxf(fabsf(b[t].z), fabsf(b[t].w)))));
    return mx;
  }
  __device__ __forceinline__ void write(uint8_t* region, int n_pad, float scale) const {
    uint8_t* lo = region + n_pad * 64;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int i = t * kTThreads + threadIdx.x, r = i >> 2, ch = i & 3;
      if (i < n_pad * 4) {
        uint32_t h[4], l[4];
        split_pack<false>(a[t].x * scale, a[t].y * scale, h[0], l[0]);
        split_pack<false>(a[t].z * scale, a[t].w * scale, h[1], l[1]);
        split_pack<false>(b[t].x * scale, b[t].y * scale, h[2], l[2]);
        split_pack<false>(b[t].z * scale, b[t].w * scale, h[3], l[3]);
        const int off = r * 64 + ((ch ^ ((r >> 1) & 3)) << 4);
        *reinterpret_cast<uint4*>(region + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
  }
};
// block-wide: the power of two that brings the largest of the threads' `mx` into [1, 2) (1 when everything is zero / denormal).
// Two barriers; s_red[9].
__device__ __forceinline__ float tpow2_block(float mx, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
  __syncthreads();
  float m = s_red[0];
#pragma unroll
  for (int w = 1; w < kTThreads / 32; ++w) m = fmaxf(m, s_red[w]);
  const uint32_t e = (__float_as_uint(m) >> 23) & 0xffu;
  return (e == 0u || e >= 253u) ? 1.f : __uint_as_float((254u - e) << 23);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// forward: ctx = dropout(softmax(Q K^T c)) V, lse = logsumexp(Q K^T c)
// ---------------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTThreads, 2) tattn_fwd_kernel(const TAttnArgs p, int lkp /* keys padded to a multiple of 128 */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_q = smem;                                   // hi at 0, lo at 8 KB
  uint8_t* s_k = s_q + 2 * 128 * 64;                     // hi, lo: lkp * 64 each
  uint8_t* s_v = s_k + 2 * lkp * 64;
  float* s_ms = reinterpret_cast<float*>(s_v + 2 * lkp * 64);   // [4 partials][128 rows]
  float* s_sum = s_ms + 4 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_sum + 4 * 128);   // [0] S ready, [1] O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int seq = blockIdx.x, head = blockIdx.y, qt = blockIdx.z;
  const int nu = lkp >> 7;                               // 128-key units

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  {
    const int qv = min(128, p.Lq - qt * 128);
    traw_issue(s_q, p.Q + (long long)seq * p.q_seq_stride + (long long)qt * 128 * p.ldq + head * 32, p.ldq, qv, 128);
    traw_issue(s_k, p.K + (long long)seq * p.Lk * p.ldkv + head * 32, p.ldkv, p.Lk, lkp);
    traw_issue(s_v, p.V + (long long)seq * p.Lk * p.ldkv + head * 32, p.ldkv, p.Lk, lkp);
    cp_async_wait_all();
    __syncthreads();
    TRaw<2> rq;
    TRaw<4> rk;
    rq.read(s_q, 128);
    rk.read(s_k, lkp);
    __syncthreads();
    rq.write(s_q, 128, 1.f);
    rk.write(s_k, lkp, 1.f);
    rk.read(s_v, lkp);
    __syncthreads();
    rk.write(s_v, lkp, 1.f);
  }
  fence_proxy_async();                                   // generic-proxy shared-memory writes -> visible to the UMMA (async proxy)
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t idesc_s = make_idesc(128, 128, false, false, false);   // S = Q K^T, both K-major
  const uint32_t idesc_o = make_idesc(128, 32, false, false, true);     // O = P V, A from TMEM, B = V MN-major
  // descriptors: constant halves + the low word of a tile (address >> 4); every MMA's descriptors are one 32-bit add away, computed by the whole
  // (converged) warp 0 so that they live in uniform registers, and the MMAs of a group are issued back to back by one elected lane
  const SDescBase kdk = sdesc_base(16, kTAtom, kSwz64), kdm = sdesc_base(kTAtom, kTAtom, kSwz64);
  const uint32_t q_lo16 = sdesc_lo(kdk, smem_u32(s_q)), k_lo16 = sdesc_lo(kdk, smem_u32(s_k)), v_lo16 = sdesc_lo(kdm, smem_u32(s_v));
  const uint32_t kv_lo = ((uint32_t)lkp * 64) >> 4;      // hi -> lo piece of K / V, in 16-byte units
  auto issue_s = [&](int u) {
    const uint32_t ku = k_lo16 + (uint32_t)(u * 128 * 64 >> 4);
    if (elect_one()) {
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        const uint32_t qp = q_lo16 + (part == 1 ? (128 * 64 >> 4) : 0), kp = ku + (part == 2 ? kv_lo : 0);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16_lohi(tmem_base, qp + 2 * k, kdk.hi, kp + 2 * k, kdk.hi, idesc_s, (part | k) ? 1u : 0u);
      }
    }
    __syncwarp();
  };
  auto issue_pv = [&](int u) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t vu = v_lo16 + (uint32_t)((u * 128 + h * 64) * 64 >> 4);
      if (elect_one()) {
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          const uint32_t vp = vu + (part == 2 ? kv_lo : 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t pcol = (uint32_t)(h * 64 + (k >> 1) * 32 + (part == 1 ? 16 : 0) + (k & 1) * 8);
            umma_f16_ts_lohi(tmem_base + 128 + (u * 2 + h) * 32, tmem_base + pcol, vp + k * (16 * 64 >> 4), kdm.hi, idesc_o, (part | k) ? 1u : 0u);
          }
        }
      }
      __syncwarp();
    }
  };

  const int r = quarter * 32 + lane;                     // query row inside the tile == TMEM lane
  const int qrow = qt * 128 + r;
  const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
  const float c2 = p.c * kLog2e;
  const long long drop_row = (((long long)seq * p.heads + head) * p.Lq + qrow) * p.Lk;

  for (int u = 0; u < nu; ++u) {
    if (warp == 0) {                                     // whole warp, uniform control flow (descriptors in uniform registers); one elected lane issues
      if (u > 0) issue_pv(u - 1);
      issue_s(u);                                        // overwrites P of the previous unit: the MMA pipe runs in issue order
      if (elect_one()) umma_commit(&bars[0]);
      __syncwarp();
    }
    mbar_wait(&bars[0], u & 1);
    fence_after_sync();
    const int key0 = u * 128 + half * 64;                // first key of this thread's 64 columns
    const int kvalid = p.Lk - key0;                      // valid keys among them (<= 0: none)
    uint32_t v[2][32];
    tmem_ld32(t_row + half * 64, v[0]);
    tmem_ld32(t_row + half * 64 + 32, v[1]);
    tmem_ld_wait();
    float mx = -INFINITY;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc)
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (cc * 32 + j < kvalid) mx = fmaxf(mx, __uint_as_float(v[cc][j]));
    const float ms = kvalid > 0 ? mx * c2 : 0.f;
    float sum = 0.f;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t pw[32];                                   // [hi 16 | lo 16]
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k0 = cc * 32 + 2 * j;
        float e0 = k0 < kvalid ? ex2_approx(fmaf(__uint_as_float(v[cc][2 * j]), c2, -ms)) : 0.f;
        float e1 = k0 + 1 < kvalid ? ex2_approx(fmaf(__uint_as_float(v[cc][2 * j + 1]), c2, -ms)) : 0.f;
        sum += e0 + e1;
        if (p.drop.thresh) {                             // dropout on the probabilities: O = (P . mask * scale) V, the sum stays undropped
          e0 = drop_keep(p.drop, (unsigned long long)(drop_row + key0 + k0)) ? e0 * p.drop.scale : 0.f;
          e1 = drop_keep(p.drop, (unsigned long long)(drop_row + key0 + k0 + 1)) ? e1 * p.drop.scale : 0.f;
        }
        split_pack<false>(e0, e1, pw[j], pw[16 + j]);
      }
      tmem_st32(t_row + half * 64 + cc * 32, pw);
    }
    s_ms[(u * 2 + half) * 128 + r] = ms;
    s_sum[(u * 2 + half) * 128 + r] = sum;
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  if (warp == 0) {
    issue_pv(nu - 1);
    if (elect_one()) umma_commit(&bars[1]);
    __syncwarp();
  }
  // merge the partial softmaxes of the row: 2 * nu (max, sum) pairs
  float M = -INFINITY;
  for (int i = 0; i < 2 * nu; ++i)
    if (s_sum[i * 128 + r] > 0.f) M = fmaxf(M, s_ms[i * 128 + r]);
  float tot = 0.f;
  for (int i = 0; i < 2 * nu; ++i)
    if (s_sum[i * 128 + r] > 0.f) tot += s_sum[i * 128 + r] * ex2_approx(s_ms[i * 128 + r] - M);
  const float inv = 1.f / tot;
  mbar_wait(&bars[1], 0);
  fence_after_sync();
  if (half == 0) {
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = 0.f;
    for (int i = 0; i < 2 * nu; ++i) {
      uint32_t t[32];
      tmem_ld32(t_row + 128 + i * 32, t);                // .aligned: every lane loads
      tmem_ld_wait();
      const float f = s_sum[i * 128 + r] > 0.f ? ex2_approx(s_ms[i * 128 + r] - M) * inv : 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = fmaf(__uint_as_float(t[j]), f, o[j]);
    }
    if (qrow < p.Lq) {
      float4* dst = reinterpret_cast<float4*>(p.ctx + ((long long)seq * p.Lq + qrow) * p.ldo + head * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
  } else if (qrow < p.Lq) {
    p.lse[((long long)seq * p.heads + head) * p.Lq + qrow] = (M + log2f(tot)) * kLn2;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// backward.  KEYMAJOR = false: dQ (and D) for one 128-query tile; KEYMAJOR = true: dK and dV for one 128-key tile.
//   P = exp(Q K^T c - lse), dP = dO V^T (. mask * scale), D_i = dO_i . O_i, dS = P (dP - D) c, dQ = dS K, dK = dS^T Q, dV = (P . mask * scale)^T dO
// ---------------------------------------------------------------------------------------------------------------------------------
template <bool KEYMAJOR>
__global__ void __launch_bounds__(kTThreads, 2) tattn_bwd_kernel(const TAttnArgs p, int lcp /* columns (keys / queries) padded to a multiple of 64 */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // row tile: X (first product) and Y (second) = Q, dO (dQ kernel) / K, V (dK/dV kernel); hi at 0, lo at +8 KB
  uint8_t* s_x = smem;
  uint8_t* s_y = s_x + 2 * 128 * 64;
  // column tensors CX, CY = K, V (dQ kernel) / Q, dO (dK/dV kernel); hi, lo: lcp * 64 each
  uint8_t* s_cx = s_y + 2 * 128 * 64;
  uint8_t* s_cy = s_cx + 2 * lcp * 64;
  float* s_lse2 = reinterpret_cast<float*>(s_cy + 2 * lcp * 64);   // [256] by query (row of the tile / column)
  float* s_D = s_lse2 + 256;
  float* s_red = s_D + 256;                                        // [32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 32);        // [0] scores ready, [1] outputs ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int seq = blockIdx.x, head = blockIdx.y, tile = blockIdx.z;
  const long long sh = (long long)seq * p.heads + head;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);

  const float* Qb = p.Q + (long long)seq * p.q_seq_stride + head * 32;
  const float* Kb = p.K + (long long)seq * p.Lk * p.ldkv + head * 32;
  const float* Vb = p.V + (long long)seq * p.Lk * p.ldkv + head * 32;
  const float* Gb = p.dO + (long long)seq * p.Lq * p.ldo + head * 32;
  const float* Ob = p.ctx + (long long)seq * p.Lq * p.ldo + head * 32;
  float sigma;
  const int ncols = KEYMAJOR ? p.Lq : p.Lk;              // valid columns
  {
    const int r0 = tile * 128;                           // first row of the tile (query / key)
    const int rv = min(128, (KEYMAJOR ? p.Lk : p.Lq) - r0);
    const float* xsrc = KEYMAJOR ? Kb + (long long)r0 * p.ldkv : Qb + (long long)r0 * p.ldq;
    const float* ysrc = KEYMAJOR ? Vb + (long long)r0 * p.ldkv : Gb + (long long)r0 * p.ldo;
    traw_issue(s_x, xsrc, KEYMAJOR ? p.ldkv : p.ldq, rv, 128);
    traw_issue(s_y, ysrc, KEYMAJOR ? p.ldkv : p.ldo, rv, 128);
    traw_issue(s_cx, KEYMAJOR ? Qb : Kb, KEYMAJOR ? p.ldq : p.ldkv, ncols, lcp);
    traw_issue(s_cy, KEYMAJOR ? Gb : Vb, KEYMAJOR ? p.ldo : p.ldkv, ncols, lcp);
    // per-row statistics travel in registers while the tiles are in flight
    const int drow = threadIdx.x >> 1, dpart = threadIdx.x & 1;      // dQ kernel: D = rowsum(dO . O), two threads per row, 16 columns each
    float4 o4[4];
    float st_l = 0.f, st_d = 0.f;
    if (!KEYMAJOR) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o4[j] = drow < rv ? __ldg(reinterpret_cast<const float4*>(Ob + (long long)(r0 + drow) * p.ldo + dpart * 16) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (dpart == 0 && drow < rv) st_l = __ldg(p.lse + sh * p.Lq + r0 + drow);
    } else if ((int)threadIdx.x < p.Lq) {
      st_l = __ldg(p.lse + sh * p.Lq + threadIdx.x);
      st_d = __ldg(p.Dbuf + sh * p.Lq + threadIdx.x);
    }
    cp_async_wait_all();
    __syncthreads();
    TRaw<2> rx, ry;
    TRaw<4> rc;
    rx.read(s_x, 128);
    ry.read(s_y, 128);
    rc.read(s_cx, lcp);
    float d = 0.f;
    if (!KEYMAJOR) {                                     // dO . O from the raw dO rows, before they are overwritten
      const float4* g4 = reinterpret_cast<const float4*>(s_y + drow * 128 + dpart * 64);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 g = g4[j], o = o4[j];
        d = fmaf(g.x, o.x, d); d = fmaf(g.y, o.y, d); d = fmaf(g.z, o.z, d); d = fmaf(g.w, o.w, d);
      }
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      sigma = tpow2_block(ry.max_abs(128), s_red);       // (barrier inside: every raw read above is complete)
    } else {
      __syncthreads();
      sigma = 1.f;
    }
    __syncthreads();
    rx.write(s_x, 128, 1.f);
    ry.write(s_y, 128, KEYMAJOR ? 1.f : sigma);
    rc.write(s_cx, lcp, 1.f);
    rc.read(s_cy, lcp);
    if (KEYMAJOR) sigma = tpow2_block(rc.max_abs(lcp), s_red + 16);
    else __syncthreads();
    __syncthreads();
    rc.write(s_cy, lcp, KEYMAJOR ? sigma : 1.f);
    if (!KEYMAJOR) {
      if (dpart == 0) {
        s_D[drow] = d * sigma;
        s_lse2[drow] = st_l * kLog2e;
        if (drow < rv) p.Dbuf[sh * p.Lq + r0 + drow] = d;
      }
    } else {
      s_lse2[threadIdx.x] = st_l * kLog2e;
      s_D[threadIdx.x] = st_d * sigma;
    }
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t idesc_1 = make_idesc(128, 64, false, false, false);    // scores: both operands K-major
  const uint32_t idesc_2 = make_idesc(128, 32, false, false, true);     // accumulations: A from TMEM, B MN-major
  // descriptors: constant halves + the low word of a tile (address >> 4) -- K-major forms for the score-type products, MN-major forms of the
  // column tensors for the accumulations; computed by the whole (converged) warp 0, the MMAs of a group issued back to back by one elected lane
  const SDescBase kdk = sdesc_base(16, kTAtom, kSwz64), kdm = sdesc_base(kTAtom, kTAtom, kSwz64);
  const uint32_t x16 = sdesc_lo(kdk, smem_u32(s_x)), y16 = sdesc_lo(kdk, smem_u32(s_y));
  const uint32_t cxk16 = sdesc_lo(kdk, smem_u32(s_cx)), cyk16 = sdesc_lo(kdk, smem_u32(s_cy));
  const uint32_t dcxm = sdesc_lo(kdm, smem_u32(s_cx)), dcym = sdesc_lo(kdm, smem_u32(s_cy));
  const uint32_t lo_r = (128 * 64) >> 4, lo_c = ((uint32_t)lcp * 64) >> 4;   // hi -> lo piece of a row tile / a column tensor, in 16-byte units
  auto issue_scores = [&](int ch) {                      // T1[128, 64] = X CX_ch^T (columns 0..63), T2 = Y CY_ch^T (columns 64..127)
    const uint32_t c0 = (uint32_t)(ch * 64 * 64 >> 4);
    if (elect_one()) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint32_t ra = t ? y16 : x16, ca = (t ? cyk16 : cxk16) + c0;
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          const uint32_t rp = ra + (part == 1 ? lo_r : 0), cp = ca + (part == 2 ? lo_c : 0);
#pragma unroll
          for (int k = 0; k < 2; ++k) umma_f16_lohi(tmem_base + t * 64, rp + 2 * k, kdk.hi, cp + 2 * k, kdk.hi, idesc_1, (part | k) ? 1u : 0u);
        }
      }
    }
    __syncwarp();
  };
  auto issue_accum = [&](int ch, uint32_t dcol, uint32_t acol, uint32_t btile16) {   // D[128, 32] (+)= A(TMEM columns acol..) B_ch
    const uint32_t acc0 = ch > 0 ? 1u : 0u;
    const uint32_t bch = btile16 + (uint32_t)(ch * 64 * 64 >> 4);
    if (elect_one()) {
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        const uint32_t bp = bch + (part == 2 ? lo_c : 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t pcol = acol + (uint32_t)((k >> 1) * 32 + (part == 1 ? 16 : 0) + (k & 1) * 8);
          umma_f16_ts_lohi(tmem_base + dcol, tmem_base + pcol, bp + k * (16 * 64 >> 4), kdm.hi, idesc_2, (part | k) ? 1u : acc0);
        }
      }
    }
    __syncwarp();
  };

  const int r = quarter * 32 + lane;                     // row of the tile == TMEM lane (query row / key row)
  const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
  const float c2 = p.c * kLog2e;
  const int nch = (ncols + 63) >> 6;
  const float my_lse2 = KEYMAJOR ? 0.f : s_lse2[r], my_D = KEYMAJOR ? 0.f : s_D[r];
  const long long drop_base = sh * p.Lq * p.Lk;

  if (warp == 0) {                                       // whole warp, uniform control flow (descriptors in uniform registers); one elected lane issues
    issue_scores(0);
    if (elect_one()) umma_commit(&bars[0]);
    __syncwarp();
  }
  for (int ch = 0; ch < nch; ++ch) {
    mbar_wait(&bars[0], ch & 1);
    fence_after_sync();
    uint32_t s[32], g[32];
    tmem_ld32(t_row + half * 32, s);
    tmem_ld32(t_row + 64 + half * 32, g);
    tmem_ld_wait();
    const int col0 = ch * 64 + half * 32;                // first column (key / query index) of this thread's 32
    uint32_t w1[32], w2[32];                             // [hi 16 | lo 16] each
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float pd[2], ds[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = col0 + 2 * j + e;
        const float l2 = KEYMAJOR ? s_lse2[col] : my_lse2, D = KEYMAJOR ? s_D[col] : my_D;
        const float pr = ex2_approx(fmaf(__uint_as_float(s[2 * j + e]), c2, -l2));
        float kf = 1.f;
        if (p.drop.thresh) {
          const long long qi = KEYMAJOR ? col : tile * 128 + r, ki = KEYMAJOR ? tile * 128 + r : col;
          kf = drop_keep(p.drop, (unsigned long long)(drop_base + qi * p.Lk + ki)) ? p.drop.scale : 0.f;
        }
        pd[e] = pr * kf;
        ds[e] = pr * (__uint_as_float(g[2 * j + e]) * kf - D) * p.c;
        if (!KEYMAJOR && col >= p.Lk) ds[e] = 0.f;       // padded keys: P is not bounded there
      }
      if (KEYMAJOR) {
        split_pack<false>(pd[0], pd[1], w1[j], w1[16 + j]);
        split_pack<false>(ds[0], ds[1], w2[j], w2[16 + j]);
      } else {
        split_pack<false>(ds[0], ds[1], w1[j], w1[16 + j]);
      }
    }
    tmem_st32(t_row + half * 32, w1);
    if (KEYMAJOR) tmem_st32(t_row + 64 + half * 32, w2);
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 0) {
      if (KEYMAJOR) {
        issue_accum(ch, 128, 0, dcym);                    // dV += (P . mask)^T dO
        issue_accum(ch, 160, 64, dcxm);                   // dK += dS^T Q
      } else {
        issue_accum(ch, 128, 0, dcxm);                    // dQ += dS K
      }
      if (ch + 1 < nch) {
        issue_scores(ch + 1);                            // overwrites T1 / T2: the MMA pipe runs in issue order
        if (elect_one()) umma_commit(&bars[0]);
      } else {
        if (elect_one()) umma_commit(&bars[1]);
      }
      __syncwarp();
    }
  }
  mbar_wait(&bars[1], 0);
  fence_after_sync();
  const float inv_sigma = 1.f / sigma;
  const int row = tile * 128 + r;
  if (KEYMAJOR || half == 0) {
    uint32_t t[32];
    tmem_ld32(t_row + 128 + (KEYMAJOR ? half * 32 : 0), t);
    tmem_ld_wait();
    float* dst = nullptr;
    if (!KEYMAJOR) { if (row < p.Lq) dst = p.dQ + ((long long)seq * p.Lq + row) * p.lddq + head * 32; }
    else if (row < p.Lk) dst = (half ? p.dK : p.dV) + ((long long)seq * p.Lk + row) * p.lddkv + head * 32;
    if (dst) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(dst)[j] = make_float4(__uint_as_float(t[4 * j]) * inv_sigma, __uint_as_float(t[4 * j + 1]) * inv_sigma,
                                                        __uint_as_float(t[4 * j + 2]) * inv_sigma, __uint_as_float(t[4 * j + 3]) * inv_sigma);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

inline size_t tattn_fwd_smem(int lkp) { return 1024 + 2 * 128 * 64 + 4 * (size_t)lkp * 64 + 2 * 4 * 128 * sizeof(float) + 64; }
inline size_t tattn_bwd_smem(int lcp) { return 1024 + 4 * 128 * 64 + 4 * (size_t)lcp * 64 + (256 + 256 + 32) * sizeof(float) + 64; }

}  // namespace tc
}  // namespace hft
