// Persistent, warp-specialised tcgen05 GEMM for the hFT projections:  C[M,N] = A[M,K] * W[N,K]^T (+ epilogue).
//
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128-byte swizzle, kStages-deep mbarrier ring)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma kind::f16, fp32 accumulators in TMEM,
//                               two 256-column accumulator buffers so tile i+1's MMAs overlap tile i's epilogue)
//   warps 2..9  epilogue       (tcgen05.ld TMEM -> registers; warp w owns TMEM lanes 32*(w%4).., column half (w-2)/4;
//                               bias / ReLU / residual + LayerNorm / sigmoid heads; 16-byte global stores)
//
// A and W are 16-bit (bf16 or fp16), K-major; every GEMM of the model has K in {64..512} and N <= 768, so the
// kernel is short-K: per 128-row tile it moves 128*K*2 bytes of A against 2*128*N*K flops, i.e. it is HBM-bound
// unless fused, which is why the epilogues carry everything that follows the projection in the reference
// (model_spec2midi.py:236,242 residual + LayerNorm; :372 ReLU; :172-175 sigmoid heads).
#pragma once
#include "tc_common.cuh"

namespace hft {
namespace tc {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // 64 x 16-bit = 128 bytes = one swizzle atom row
constexpr int kStages = 4;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;

enum Epi : int { EPI_STORE = 0, EPI_RELU = 1, EPI_LN = 2, EPI_HEADS = 3 };

struct GemmParams {
  int m_tiles;            // M / 128
  int n_tiles;            // N / n_tile
  int n_tile;             // UMMA N (multiple of 64, <= 256)
  int k_chunks;           // K / 64
  const float* bias;      // [N]
  // EPI_STORE / EPI_RELU / EPI_LN: 16-bit output [M, ldc]
  void* out;
  int ldc;
  // EPI_LN
  const void* resid16;    // 16-bit residual [M, H] (or nullptr when resid32 is used)
  const float* resid32;   // fp32 residual table [resid_rows, H] indexed by row % resid_rows (constant pitch queries)
  int resid_rows;
  const float* gamma;
  const float* beta;
  // EPI_HEADS: columns 0..V-1 velocity logits, V..V+2 onset / offset / mpe logits
  float* onset;
  float* offset;
  float* mpe;
  float* velocity;
  int n_vel;
  int time_major;         // rows are (b, note, frame): permute back to [B, frame, note]
  int n_frame, n_note;
};

__host__ __device__ constexpr size_t gemm_smem_bytes(int n_tile) {
  return 1024 /*align slack*/ + (size_t)kStages * (kBlockM * kBlockK * 2 + (size_t)n_tile * kBlockK * 2) + 4096 /*LN exchange*/ + 256 /*barriers*/;
}

template <bool BF16, int EPI, int HALF_COLS>   // HALF_COLS = n_tile / 2 = columns each epilogue thread owns (multiple of 32)
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int n_tile = HALF_COLS * 2;
  const uint32_t a_bytes = kBlockM * kBlockK * 2, w_bytes = (uint32_t)n_tile * kBlockK * 2;
  uint8_t* s_a = smem;
  uint8_t* s_w = smem + kStages * a_bytes;
  float* s_ln = reinterpret_cast<float*>(s_w + kStages * w_bytes);                      // [2 stats][2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_ln) + 4096);
  uint64_t* full = bars;                 // [kStages]
  uint64_t* empty = bars + kStages;      // [kStages]
  uint64_t* tfull = bars + 2 * kStages;  // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], a_bytes + w_bytes);
          tma_load_2d(s_a + stage * a_bytes, &map_a, kc * kBlockK, mt * kBlockM, &full[stage]);
          tma_load_2d(s_w + stage * w_bytes, &map_w, kc * kBlockK, nt * n_tile, &full[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(kBlockM, n_tile, BF16, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int ab = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty[ab], aphase ^ 1);             // epilogue has drained this accumulator buffer
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + ab * 256;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&full[stage], phase);
          fence_after_sync();
          const uint32_t a_addr = smem_u32(s_a + stage * a_bytes), w_addr = smem_u32(s_w + stage * w_bytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t da = make_sdesc(a_addr + k * 32, 16, 1024, kSwz128);
            const uint64_t dw = make_sdesc(w_addr + k * 32, 16, 1024, kSwz128);
            umma_f16(d_tmem, da, dw, idesc, (kc | k) != 0);
          }
          umma_commit(&empty[stage]);                   // frees the smem stage when these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[ab]);                        // accumulator complete
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = ew >> 2;
    const int row_in_tile = quarter * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
      const int ab = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull[ab], aphase);
      fence_after_sync();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + ab * 256 + half * HALF_COLS;
      const long long row = (long long)mt * kBlockM + row_in_tile;
      const int col_base = nt * n_tile + half * HALF_COLS;      // global output column of this thread's first value

      if (EPI == EPI_STORE || EPI == EPI_RELU) {
        uint32_t* orow = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.out) + row * p.ldc + col_base);
#pragma unroll 1
        for (int c = 0; c < HALF_COLS / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v0 = __uint_as_float(r[2 * j]) + __ldg(p.bias + col_base + c * 32 + 2 * j);
            float v1 = __uint_as_float(r[2 * j + 1]) + __ldg(p.bias + col_base + c * 32 + 2 * j + 1);
            if (EPI == EPI_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            pk[j] = Op16<BF16>::pack(v0, v1);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(orow + c * 16 + j * 4) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
      } else if (EPI == EPI_LN) {
        // v = acc + bias + residual, kept in registers (HALF_COLS values); LayerNorm over the full row of 2*HALF_COLS
        float v[HALF_COLS];
        const int H = 2 * HALF_COLS;
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < HALF_COLS / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tmem_ld_wait();
          if (p.resid16) {
            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.resid16) + row * H + col_base + c * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 q = rp[j];
              uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[c * 32 + j * 8 + 2 * e] = __uint_as_float(r[j * 8 + 2 * e]) + Op16<BF16>::lo(w4[e]);
                v[c * 32 + j * 8 + 2 * e + 1] = __uint_as_float(r[j * 8 + 2 * e + 1]) + Op16<BF16>::hi(w4[e]);
              }
            }
          } else {
            const float* rp = p.resid32 + (row % p.resid_rows) * H + col_base + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[c * 32 + j] = __uint_as_float(r[j]) + __ldg(rp + j);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[c * 32 + j] += __ldg(p.bias + col_base + c * 32 + j);
            sum += v[c * 32 + j];
          }
        }
        // the accumulator is in registers now: hand the TMEM buffer back before the row statistics
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[ab]);
        s_ln[half * 128 + row_in_tile] = sum;
        named_bar_sync(1, kEpiWarps * 32);
        const float mean = (s_ln[row_in_tile] + s_ln[128 + row_in_tile]) / (float)H;
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < HALF_COLS; ++j) { float d = v[j] - mean; sq = fmaf(d, d, sq); }
        s_ln[256 + half * 128 + row_in_tile] = sq;
        named_bar_sync(1, kEpiWarps * 32);
        const float rstd = rsqrtf((s_ln[256 + row_in_tile] + s_ln[384 + row_in_tile]) / (float)H + 1e-5f);
        uint32_t* orow = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.out) + row * p.ldc + col_base);
#pragma unroll
        for (int j = 0; j < HALF_COLS / 8; ++j) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            int c0 = j * 8 + 2 * e;
            float y0 = (v[c0] - mean) * rstd * __ldg(p.gamma + col_base + c0) + __ldg(p.beta + col_base + c0);
            float y1 = (v[c0 + 1] - mean) * rstd * __ldg(p.gamma + col_base + c0 + 1) + __ldg(p.beta + col_base + c0 + 1);
            pk[e] = Op16<BF16>::pack(y0, y1);
          }
          *reinterpret_cast<uint4*>(orow + j * 4) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        named_bar_sync(1, kEpiWarps * 32);              // s_ln is reused by the next tile
        continue;                                       // tempty already signalled
      } else {                                          // EPI_HEADS
        long long orow_idx = row;
        if (p.time_major) {
          int f = (int)(row % p.n_frame);
          long long bn = row / p.n_frame;
          int n = (int)(bn % p.n_note);
          orow_idx = ((bn / p.n_note) * p.n_frame + f) * p.n_note + n;
        }
#pragma unroll 1
        for (int c = 0; c < HALF_COLS / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tmem_ld_wait();
          const int c0 = col_base + c * 32;
          if (c0 + 32 <= p.n_vel) {
            if (p.velocity) {
              float4* dst = reinterpret_cast<float4*>(p.velocity + orow_idx * p.n_vel + c0);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                dst[j] = make_float4(__uint_as_float(r[4 * j]) + __ldg(p.bias + c0 + 4 * j), __uint_as_float(r[4 * j + 1]) + __ldg(p.bias + c0 + 4 * j + 1),
                                     __uint_as_float(r[4 * j + 2]) + __ldg(p.bias + c0 + 4 * j + 2), __uint_as_float(r[4 * j + 3]) + __ldg(p.bias + c0 + 4 * j + 3));
            }
          } else if (c0 == p.n_vel) {
            float* dsts[3] = {p.onset, p.offset, p.mpe};
#pragma unroll
            for (int j = 0; j < 3; ++j)
              if (dsts[j]) dsts[j][orow_idx] = 1.f / (1.f + expf(-(__uint_as_float(r[j]) + __ldg(p.bias + c0 + j))));
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[ab]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace hft
