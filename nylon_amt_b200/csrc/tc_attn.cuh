// tcgen05 attention for one (sequence, head, 128-query tile):  softmax(Q K^T / sqrt(dh)) V
// (reference MultiHeadAttentionLayer.forward, model_spec2midi.py:342-348; probabilities returned at :360).
//
// Every sequence of the model is short (Lk <= 256), so a whole score row fits one TMEM accumulator: no online
// softmax.  Q, K, V tiles arrive by TMA (swizzled); S = Q K^T is one UMMA chain into TMEM; 8 warps (two per TMEM lane
// quarter, each owning half of the key columns) compute max / exp2 / sum in fp32 and write the un-normalised
// probabilities as 16-bit into shared memory in the UMMA K-major 128B-swizzled layout; O = P V is a second UMMA chain
// whose B operand is V read MN-major straight from its row-major tile (no transpose pass); the epilogue scales by
// 1/sum, stages the tile in swizzled shared memory and writes it with one TMA store (full 128-byte lines).
// Shared memory: P and the output staging overlay Q, K (dead after S is complete) so two CTAs fit per SM in the
// single-product modes; TMEM: O overlays S.
// x3 mode (split operands, see tc_gemm.cuh): S = Qh Kh + Ql Kh + Qh Kl, O = Ph Vh + Pl Vh + Ph Vl, ctx stored hi | lo.
#pragma once
#include "tc_common.cuh"

namespace hft {
namespace tc {

struct AttnParams {
  int lq;                 // valid query rows per sequence (<= 128 * q_tiles)
  int lk;                 // valid keys per sequence (<= LK)
  int q_seq_rows;         // rows between consecutive sequences in the Q tensor (0 = the same queries for every sequence)
  int q_tiles;            // ceil(lq / 128)
  int heads;
  int q_col0, k_col0, v_col0;   // first column of head 0 inside the Q / K / V tensors
  float scale_log2e;      // log2(e) / sqrt(dh)
  void* ctx;              // 16-bit [n_seq * lq, ld_ctx]
  int ld_ctx;
  float* probs;           // optional fp32 [n_seq, heads, lq, lk]
  int q_lo_off, kv_lo_off;   // x3: column distance between the hi and lo halves of the Q tensor / the K,V tensor
  int ctx_lo_off;         // x3: same for ctx
  int tma_store;          // 1: ctx tile written by TMA (needs lq % 128 == 0 and dh == 64)
  int s_single, pv_single;   // mixed-precision plan (x3 layout): scores / P V as ONE product (tc_attn2.cuh kernels only; this file's kernel ignores them)
};

template <int DH, int LK, bool X3 = false>
struct AttnSmem {
  static constexpr int parts = X3 ? 2 : 1;                                   // hi (+ lo) copies of every operand
  static constexpr int q_bytes = 128 * DH * 2;
  static constexpr int k_bytes = LK * DH * 2;
  static constexpr int v_bytes = LK * DH * 2;
  static constexpr int p_blocks = (LK + 63) / 64;
  static constexpr int p_bytes = p_blocks * 128 * 128;                       // [blocks][128 rows][64 x 16-bit]
  static constexpr int qk_al = (q_bytes + k_bytes + 1023) / 1024 * 1024;     // one Q|K pair, 1024-aligned
  static constexpr int front = (parts * qk_al > parts * p_bytes) ? parts * qk_al : parts * p_bytes;   // Q|K region, reused by P
  static constexpr int front_al = (front + 1023) / 1024 * 1024;
  static constexpr int total = 1024 + front_al + parts * v_bytes + 2048 /*row stats*/ + 64;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kAttnThreads = 256;

template <bool BF16, int DH, int LK, bool PROBS, bool X3>
__global__ void __launch_bounds__(kAttnThreads) attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                                                            const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o,
                                                            const __grid_constant__ AttnParams p) {
  using L = AttnSmem<DH, LK, X3>;
  constexpr int kParts = X3 ? 2 : 1;
  constexpr uint32_t kSwz = (DH == 64) ? kSwz128 : kSwz64;
  constexpr uint32_t kRowBytes = DH * 2;                  // 128 or 64
  constexpr uint32_t kAtom = 8 * kRowBytes;               // 8 rows of one swizzle atom: 1024 or 512 bytes
  // O overlays the first DH columns of S, except when the probabilities are returned (S is re-read after the PV MMA was issued)
  constexpr uint32_t kOCol = PROBS ? LK : 0;
  constexpr uint32_t kNeed = PROBS ? LK + DH : LK;
  constexpr uint32_t kTmemCols = (kNeed <= 32) ? 32 : (kNeed <= 64) ? 64 : (kNeed <= 128) ? 128 : (kNeed <= 256) ? 256 : 512;
  // key columns owned by the two column halves (multiples of 32): 256 -> 128|128, 128 -> 64|64, 96 -> 64|32
  constexpr int C0 = (LK == 96) ? 64 : LK / 2, C1 = LK - C0;
  static_assert(LK % 32 == 0 && LK <= 256 && DH <= LK && C0 % 32 == 0 && C1 % 32 == 0, "unsupported tile");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_q = smem;                                    // part i: Q at i*qk_al, K right after it
  uint8_t* s_k = smem + L::q_bytes;
  uint8_t* s_p = smem;                                    // overlays Q|K once S is complete; part i at i*p_bytes
  uint8_t* s_v = smem + L::front_al;                      // part i at i*v_bytes
  float* s_red = reinterpret_cast<float*>(s_v + kParts * L::v_bytes);        // [2 stats][2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_red) + 2048);   // [0] loads, [1] S ready, [2] O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int item = blockIdx.x;
  const int qt = item % p.q_tiles;
  const int head = (item / p.q_tiles) % p.heads;
  const int seq = item / (p.q_tiles * p.heads);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], kParts * (L::q_bytes + L::k_bytes + L::v_bytes));
#pragma unroll
    for (int part = 0; part < kParts; ++part) {
      tma_load_2d(s_q + part * L::qk_al, &map_q, part * p.q_lo_off + p.q_col0 + head * DH, seq * p.q_seq_rows + qt * 128, &bars[0]);
      tma_load_2d(s_k + part * L::qk_al, &map_k, part * p.kv_lo_off + p.k_col0 + head * DH, seq * p.lk, &bars[0]);
      tma_load_2d(s_v + part * L::v_bytes, &map_v, part * p.kv_lo_off + p.v_col0 + head * DH, seq * p.lk, &bars[0]);
    }
    mbar_wait(&bars[0], 0);
    fence_after_sync();
    // S[128, LK] = Q[128, DH] * K[LK, DH]^T : both operands K-major.  x3: Qh Kh + Ql Kh + Qh Kl.
    const uint32_t idesc = make_idesc(128, LK, BF16, false, false);
    const uint32_t qa = smem_u32(s_q), ka = smem_u32(s_k);
    uint32_t acc = 0;
#pragma unroll
    for (int part = 0; part < (X3 ? 3 : 1); ++part) {
      const uint32_t qp = qa + (part == 1 ? L::qk_al : 0), kp = ka + (part == 2 ? L::qk_al : 0);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) {
        umma_f16(tmem_base, make_sdesc(qp + k * 32, 16, kAtom, kSwz), make_sdesc(kp + k * 32, 16, kAtom, kSwz), idesc, acc);
        acc = 1;
      }
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  fence_after_sync();

  // ---- softmax: thread = (query row, column half) ----------------------------------------------------------------
  const int r = quarter * 32 + lane;                       // row inside the tile == TMEM lane
  const int col0 = half ? C0 : 0;                          // first key column of this thread
  const int ncol = half ? C1 : C0;
  const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + col0;
  const bool mask = p.lk < LK;                             // only the 88-key sequences are padded (to 96)
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < C0 / 32; ++c) {
    if (c * 32 < ncol) {
      uint32_t v[32];
      tmem_ld32(t_row + c * 32, v);
      tmem_ld_wait();
      if (mask) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + c * 32 + j < p.lk) mx = fmaxf(mx, __uint_as_float(v[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
    }
  }
  s_red[half * 128 + r] = mx;
  named_bar_sync(1, kAttnThreads);
  mx = fmaxf(s_red[r], s_red[128 + r]);
  float sum = 0.f;
  const float mxs = mx * p.scale_log2e;
#pragma unroll 1
  for (int c = 0; c < C0 / 32; ++c) {
    if (c * 32 < ncol) {
      uint32_t v[32];
      tmem_ld32(t_row + c * 32, v);
      tmem_ld_wait();
      uint32_t pk[16], pl[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), p.scale_log2e, -mxs));
        float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2e, -mxs));
        if (mask) {
          if (col0 + c * 32 + 2 * j >= p.lk) e0 = 0.f;
          if (col0 + c * 32 + 2 * j + 1 >= p.lk) e1 = 0.f;
        }
        sum += e0 + e1;
        pk[j] = Op16<BF16>::pack(e0, e1);
        if (X3) pl[j] = Op16<BF16>::pack(e0 - Op16<BF16>::lo(pk[j]), e1 - Op16<BF16>::hi(pk[j]));
      }
      // P[r, col .. col+31] -> K-major 128B-swizzled blocks of 64 columns (Q and K are dead: the S MMA has completed)
      const int col = col0 + c * 32;
      uint8_t* blk = s_p + (col >> 6) * (128 * 128) + r * 128;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int chunk = (((col >> 5) & 1) * 4 + q4) ^ (r & 7);
        *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
        if (X3) *reinterpret_cast<uint4*>(blk + L::p_bytes + chunk * 16) = make_uint4(pl[4 * q4], pl[4 * q4 + 1], pl[4 * q4 + 2], pl[4 * q4 + 3]);
      }
    }
  }
  s_red[256 + half * 128 + r] = sum;
  fence_proxy_async();                                     // generic-proxy smem writes -> visible to the UMMA (async proxy)
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  if (threadIdx.x == 0) {
    // O[128, DH] = P[128, LK] * V[LK, DH] : A = P K-major (128B swizzle), B = V MN-major (d contiguous)
    const uint32_t idesc = make_idesc(128, DH, BF16, false, true);
    const uint32_t pa = smem_u32(s_p), va = smem_u32(s_v);
    uint32_t acc = 0;
#pragma unroll
    for (int part = 0; part < (X3 ? 3 : 1); ++part) {                 // x3: Ph Vh + Pl Vh + Ph Vl
      const uint32_t pp = pa + (part == 1 ? L::p_bytes : 0), vp = va + (part == 2 ? L::v_bytes : 0);
#pragma unroll
      for (int k = 0; k < LK / 16; ++k) {
        const uint64_t dp = make_sdesc(pp + (k >> 2) * (128 * 128) + (k & 3) * 32, 16, 1024, kSwz128);
        const uint64_t dv = make_sdesc(vp + k * 16 * kRowBytes, kAtom, kAtom, kSwz);
        umma_f16(tmem_base + kOCol, dp, dv, idesc, acc);
        acc = 1;
      }
    }
    umma_commit(&bars[2]);
  }
  const float inv_sum = 1.f / (s_red[256 + r] + s_red[384 + r]);
  if (PROBS) {                                             // normalised probabilities (model_spec2midi.py:345,360), fp32
    const int qrow = qt * 128 + r;
    const bool live = qrow < p.lq;
    float* prow = p.probs + (((long long)seq * p.heads + head) * p.lq + (live ? qrow : 0)) * p.lk + col0;
    // S is re-read here while the PV MMA runs: with PROBS the O block sits after S, so nothing is overwritten.
#pragma unroll 1
    for (int c = 0; c < C0 / 32; ++c) {
      if (c * 32 < ncol) {
        uint32_t v[32];
        tmem_ld32(t_row + c * 32, v);                      // .aligned: every lane loads, only live rows store
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (col0 + c * 32 + 4 * j + 3 < p.lk) {
              *reinterpret_cast<float4*>(prow + c * 32 + 4 * j) =
                  make_float4(ex2_approx(fmaf(__uint_as_float(v[4 * j]), p.scale_log2e, -mxs)) * inv_sum,
                              ex2_approx(fmaf(__uint_as_float(v[4 * j + 1]), p.scale_log2e, -mxs)) * inv_sum,
                              ex2_approx(fmaf(__uint_as_float(v[4 * j + 2]), p.scale_log2e, -mxs)) * inv_sum,
                              ex2_approx(fmaf(__uint_as_float(v[4 * j + 3]), p.scale_log2e, -mxs)) * inv_sum);
            }
          }
        }
      }
    }
  }
  mbar_wait(&bars[2], 0);
  fence_after_sync();

  // ---- epilogue (column half 0 only): O / sum -> 16-bit context -------------------------------------------------------
  if (half == 0) {
    const int qrow = qt * 128 + r;
    const uint32_t t_o = tmem_base + ((uint32_t)(quarter * 32) << 16) + kOCol;
    uint8_t* stage = smem;                                 // Q/K/P are dead once the PV MMA has completed
#pragma unroll
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t o[32];
      tmem_ld32(t_o + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = __uint_as_float(o[8 * j + 2 * e]) * inv_sum, b = __uint_as_float(o[8 * j + 2 * e + 1]) * inv_sum;
          h[e] = Op16<BF16>::pack(a, b);
          if (X3) l[e] = Op16<BF16>::pack(a - Op16<BF16>::lo(h[e]), b - Op16<BF16>::hi(h[e]));
        }
        if (p.tma_store) {                                 // DH == 64: 128-byte rows, 128B swizzle
          const int chunk = (c * 4 + j) ^ (r & 7);
          *reinterpret_cast<uint4*>(stage + r * 128 + chunk * 16) = make_uint4(h[0], h[1], h[2], h[3]);
          if (X3) *reinterpret_cast<uint4*>(stage + 16384 + r * 128 + chunk * 16) = make_uint4(l[0], l[1], l[2], l[3]);
        } else if (qrow < p.lq) {
          uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.ctx) + ((long long)seq * p.lq + qrow) * p.ld_ctx + head * DH);
          *reinterpret_cast<uint4*>(dst + c * 16 + j * 4) = make_uint4(h[0], h[1], h[2], h[3]);
          if (X3) *reinterpret_cast<uint4*>(dst + p.ctx_lo_off / 2 + c * 16 + j * 4) = make_uint4(l[0], l[1], l[2], l[3]);
        }
      }
    }
    if (p.tma_store) {
      fence_proxy_async();
      named_bar_sync(2, 128);
      if (threadIdx.x == 0) {
        const int row0 = seq * p.lq + qt * 128;
#pragma unroll
        for (int part = 0; part < kParts; ++part)
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&map_o)),
                       "r"(smem_u32(stage + part * 16384)), "r"(part * p.ctx_lo_off + head * DH), "r"(row0)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace tc
}  // namespace hft
