// tcgen05 weight-gradient GEMM of the TRAINING step (reduced model, BASELINE configs[4]):
//   dW[N, K] += dY[M, N]^T X[M, K],   db[N] += colsum(dY)          (nn.Linear under loss.backward(), training/train.py:158)
// with M in the 10^5 (every row of the batch), N <= 192 output features, K = 64 / 128 input features.  Replaces dw_gemm_kernel
// (fp32 CUDA cores) -- like the other Linear kernels the shape is bound by the HBM traffic of dY and X once the products
// run on the tensor cores.
//
// The reduction runs over the ROWS, i.e. over the slow dimension of both row-major operands, so both are fed to the MMA
// MN-major straight from row-major tiles ([64 rows][128 bytes] per 64-column block, 128-byte swizzle): D[n, k] with
// M_mma = 128 output features (two 64-column blocks of dY, the second one a block of zeros when N runs out),
// N_mma = K, K_mma = 16 rows per instruction.  db comes from the same MMAs: one extra 16-column B block whose column 0 is 1.
// Arithmetic: bf16 with THREE pieces per value (x = b0 + b1 + b2, 8 mantissa bits each; fp32 exponent range, so the ~1e-6
// gradients need no scaling) and the six products that matter (b0c0, b0c1, b1c0, b1c1, b0c2, b2c0), fp32 accumulation in
// TMEM: fp32-class results.  CTAs split the rows (two CTAs per SM), accumulate their slice in TMEM and add it to dW / db
// with one set of fp32 atomics; the next stage's fp32 rows are loaded into registers while the current stage's MMAs run.
#pragma once
#include "tc_common.cuh"

namespace hft {
namespace tc {

struct TDwArgs {
  const float* dY; int ldy;
  const float* X; int ldx;
  float* dW; int ldw;
  float* db;                   // [N] or NULL
  long long M; int N, K;
  long long rows_per_cta;      // multiple of the stage height
  int tmem_cols;               // power of two >= groups * (K + 32)
};

constexpr int kDwThreads = 256;

// x = b0 + b1 + b2 for a pair of values, packed as bf16x2 words
__device__ __forceinline__ void split3_pack(float a, float b, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  p0 = Op16<true>::pack(a, b);
  const float ra = a - Op16<true>::lo(p0), rb = b - Op16<true>::hi(p0);
  p1 = Op16<true>::pack(ra, rb);
  p2 = Op16<true>::pack(ra - Op16<true>::lo(p1), rb - Op16<true>::hi(p1));
}

// KB = K / 64; kDwRows = rows per stage (64 or 32); MINB = CTAs per SM the register budget is cut for (the kernel is bound by the latency of its
// global loads: more resident CTAs = more bytes in flight)
template <int KB, int kDwRows, int MINB>
__global__ void __launch_bounds__(kDwThreads, MINB) tdw_kernel(const TDwArgs p) {
  constexpr int kDwBlk = kDwRows * 128;              // bytes of one 64-column block of a stage (one piece)
  constexpr int TMAX = kDwRows / 8;                  // 8-float chunks per thread and stage: rows * (N + K) / 8 / 256 <= rows / 8
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int N = p.N;
  const int PB = (N + 63) >> 6;                      // 64-column blocks of dY
  const int G = (PB + 1) >> 1;                       // M = 128 groups (pairs of blocks)
  const int p_piece = (PB + 1) * kDwBlk;             // one piece of dY: PB blocks + the block of zeros
  constexpr int r_piece = KB * kDwBlk;
  uint8_t* s_p = smem;                               // [3 pieces][PB + 1][rows x 128 B]
  uint8_t* s_r = s_p + 3 * p_piece;                  // [3 pieces][KB][rows x 128 B]
  uint8_t* s_one = s_r + 3 * r_piece;                // [rows x 128 B]: column 0 = 1
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_one + kDwBlk);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  // zero every dY piece once (padding columns and the zero block stay zero), build the ones block
  for (int i = tid; i < 3 * p_piece / 16; i += kDwThreads) reinterpret_cast<uint4*>(s_p)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < kDwBlk / 16; i += kDwThreads) {
    const int r = i >> 3, ch = i & 7;                // logical chunk 0 of row r sits at position (0 ^ (r & 7))
    reinterpret_cast<uint4*>(s_one)[i] = make_uint4(ch == (r & 7) ? 0x00003F80u : 0u, 0u, 0u, 0u);   // bf16 1.0 in element 0
  }

  const int chp = N >> 3, chr = KB * 8, chs = chp + chr;   // valid 8-float chunks per row: dY, X, both
  const int n_tasks = kDwRows * chs;
  const long long m_begin = (long long)blockIdx.x * p.rows_per_cta;
  const long long m_end = m_begin + p.rows_per_cta < p.M ? m_begin + p.rows_per_cta : p.M;
  float4 ra[TMAX][2];
  auto load_stage = [&](long long m0) {
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
      const int idx = t * kDwThreads + tid, r = idx / chs, c = idx % chs;
      const long long row = m0 + r;
      if (idx < n_tasks && row < m_end) {
        const float4* s4 = c < chp ? reinterpret_cast<const float4*>(p.dY + row * p.ldy + c * 8) : reinterpret_cast<const float4*>(p.X + row * p.ldx + (c - chp) * 8);
        ra[t][0] = __ldg(s4);
        ra[t][1] = __ldg(s4 + 1);
      } else {
        ra[t][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        ra[t][1] = ra[t][0];
      }
    }
  };
  if (m_begin < m_end) load_stage(m_begin);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t idesc_w = make_idesc(128, KB * 64, true, true, true);     // A = dY blocks, B = X blocks, both MN-major
  const uint32_t idesc_b = make_idesc(128, 16, true, true, true);          // B = the ones block
  const uint64_t dp = make_sdesc(smem_u32(s_p), kDwBlk, 1024, kSwz128), dr = make_sdesc(smem_u32(s_r), kDwBlk, 1024, kSwz128);   // bases: piece 0, block 0
  const uint64_t done = make_sdesc(smem_u32(s_one), kDwBlk, 1024, kSwz128);
  const uint32_t gcols = KB * 64 + 32;                                      // TMEM columns per group: dW rows | db (16 used)
  uint32_t phase = 0;
  bool first = true;

  for (long long m0 = m_begin; m0 < m_end; m0 += kDwRows) {
    if (!first) {                                    // the previous stage's MMAs still read the tiles
      mbar_wait(bar, phase);
      phase ^= 1;
      fence_after_sync();
    }
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
      const int idx = t * kDwThreads + tid, r = idx / chs, c = idx % chs;
      if (idx < n_tasks) {
        const float4 a = ra[t][0], b = ra[t][1];
        uint32_t w0[4], w1[4], w2[4];
        split3_pack(a.x, a.y, w0[0], w1[0], w2[0]);
        split3_pack(a.z, a.w, w0[1], w1[1], w2[1]);
        split3_pack(b.x, b.y, w0[2], w1[2], w2[2]);
        split3_pack(b.z, b.w, w0[3], w1[3], w2[3]);
        const int cc = c < chp ? c : c - chp;
        uint8_t* base = c < chp ? s_p : s_r;
        const int piece = c < chp ? p_piece : r_piece;
        const int off = (cc >> 3) * kDwBlk + r * 128 + (((cc & 7) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(base + off) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
        *reinterpret_cast<uint4*>(base + piece + off) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        *reinterpret_cast<uint4*>(base + 2 * piece + off) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
      }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 0) {
      for (int g = 0; g < G; ++g) {
        uint32_t acc = first ? 0u : 1u;
        // the six products that matter: (piece of dY, piece of X)
        const int pa_i[6] = {0, 0, 1, 1, 0, 2}, pb_i[6] = {0, 1, 0, 1, 2, 0};
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const uint64_t ap = sdesc_advance(dp, pa_i[q] * p_piece + 2 * g * kDwBlk), bp = sdesc_advance(dr, pb_i[q] * r_piece);
#pragma unroll
          for (int k = 0; k < kDwRows / 16; ++k) {
            umma_f16(tmem_base + g * gcols, sdesc_advance(ap, k * 2048), sdesc_advance(bp, k * 2048), idesc_w, acc);
            acc = 1;
          }
        }
        acc = first ? 0u : 1u;
#pragma unroll
        for (int q = 0; q < 3; ++q) {                // db: every piece of dY times the ones block
          const uint64_t ap = sdesc_advance(dp, q * p_piece + 2 * g * kDwBlk);
#pragma unroll
          for (int k = 0; k < kDwRows / 16; ++k) {
            umma_f16(tmem_base + g * gcols + KB * 64, sdesc_advance(ap, k * 2048), sdesc_advance(done, k * 2048), idesc_b, acc);
            acc = 1;
          }
        }
      }
      umma_commit(bar);
    }
    first = false;
    if (m0 + kDwRows < m_end) load_stage(m0 + kDwRows);          // in flight while the MMAs run
  }
  if (!first) {
    mbar_wait(bar, phase);
    fence_after_sync();
    // ---- this CTA's slice -> dW / db (fp32 atomics): thread = output feature (TMEM lane), 8 warps = 4 lane quarters x 2 column halves ----
    const int q = warp & 3, half = warp >> 2;
    for (int g = 0; g < G; ++g) {
      const int n = g * 128 + q * 32 + lane;
      for (int c = half; c < KB * 2; c += 2) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + g * gcols + c * 32, v);
        tmem_ld_wait();
        if (n < N) {
          float* dst = p.dW + (long long)n * p.ldw + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(v[j]));
        }
      }
      if (half == 0 && p.db) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + g * gcols + KB * 64)
                     : "memory");
        tmem_ld_wait();
        if (n < N) atomicAdd(p.db + n, __uint_as_float(v[0]));
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

inline size_t tdw_smem(int KB, int N, int rows) { return 1024 + (3 * ((size_t)((N + 63) / 64) + 1) + 3 * (size_t)KB + 1) * rows * 128 + 64; }

}  // namespace tc
}  // namespace hft
