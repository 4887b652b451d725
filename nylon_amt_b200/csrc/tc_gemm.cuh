// Persistent, warp-specialised tcgen05 GEMM for the hFT projections:  C[M,N] = A[M,K] * W[N,K]^T (+ epilogue).
//
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128-byte swizzle, mbarrier ring of A (and W) chunks)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma kind::f16, fp32 accumulators in TMEM,
//                               two 256-column accumulator buffers so tile i+1's MMAs overlap tile i's epilogue)
//   warps 2..9  epilogue       two warpgroups; warpgroup g drains accumulator buffer g (tiles of parity g), one thread per
//                               tile row (TMEM lane) and the whole row of n_tile columns, so LayerNorm needs no cross-thread
//                               exchange.  tcgen05.ld TMEM -> registers (LayerNorm: one statistics pass and one output
//                               pass over TMEM, nothing held in registers across them); bias / gamma / beta come from
//                               shared memory; bias / ReLU / LayerNorm / sigmoid heads; every warp stages its own
//                               32-row x 64-column blocks in swizzled shared memory and writes them with its own TMA
//                               stores (no CTA-wide barrier anywhere in the epilogue)
//
// A and W are 16-bit (bf16 or fp16), K-major.  Every GEMM of the model has K in {64..512} and N <= 768: per 128-row
// tile it moves 128*K*2 bytes of A against 2*128*N*K flops, so it is bound by HBM/L2 bytes, not by the tensor pipe.
// Hence:
//   * W-resident mode: a CTA keeps its [n_tile, K] slice of W in shared memory for its whole life and only streams A
//     (re-streaming W per tile from L2 was the first bottleneck measured, profiles/r01_v1_*).
//   * the residual add of `x = LN(x + f(x))` (model_spec2midi.py:236,242) is folded into the accumulator by the
//     tensor core: the residual tile arrives through the same TMA ring and is multiplied by a 64x64 identity block
//     (exact: 1.0 * x accumulates in fp32), so the epilogue has no strided residual reads.
//   * PAIR mode (cta_group::2): two CTAs of a cluster share one UMMA tile of 256 rows.  Each CTA stages its own 128 rows
//     of A and only HALF of the W rows (the tensor cores exchange the B halves), which halves the L2 -> shared-memory
//     traffic and the shared-memory footprint of W; the even CTA issues the MMAs for both, every CTA runs its own
//     epilogue on its own 128 accumulator rows.
//   * x3 mode (split operands): A = A_hi + A_lo, W = W_hi + W_lo stored side by side ([rows, 2K]); the accumulator
//     receives A_hi W_hi + A_lo W_hi + A_hi W_lo, which carries ~22 mantissa bits with fp16 parts (fp32-class result).
#pragma once
#include "tc_common.cuh"

namespace hft {
namespace tc {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // 64 x 16-bit = 128 bytes = one swizzle atom row
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;
constexpr int kChunkA = kBlockM * kBlockK * 2;          // 16 KB
constexpr int kWarpStage = 32 * 64 * 2;                 // one warp's 32 x 64 output staging block (one TMA-store box), 4 KB
constexpr int kConstBytes = 3 * 256 * 4;                // bias | gamma | beta of this CTA's column slice

enum Epi : int { EPI_STORE = 0, EPI_RELU = 1, EPI_LN = 2, EPI_HEADS = 3 };

struct GemmParams {
  int m_tiles;            // M / 128 (PAIR: M / 256)
  int n_tiles;            // N / n_tile
  int k_chunks;           // K / 64
  int w_resident;         // 1: the CTA's W slice stays in shared memory; 0: W chunks stream through the ring with A
  int a_stages;           // depth of the A ring (16 KB slots: A chunks and residual chunks)
  int w_stages;           // depth of the W ring (n_tile x 64 slots; 0 when W is resident)
  int has_resid;          // EPI_LN: residual tile added through the identity MMA
  int resid_period;       // > 0: residual row tile index = m_tile % resid_period (constant pitch-query table)
  int x3;                 // split-operand LAYOUT (A/W/out/resid are [rows, 2*cols]: hi | lo)
  int single;             // x3 layout, but only the A_hi W_hi product is accumulated (mixed-precision plan): the lo halves of A and W are
                          // neither loaded nor multiplied; the output (and the residual) keep both halves
  int out_col0;           // first output column inside the output tensor
  int a_lo_off, w_lo_off, out_lo_off;   // x3: column distance between the hi and lo halves of A / W / out
  const float* bias;      // [N]
  const float* gamma;     // EPI_LN
  const float* beta;
  // EPI_HEADS: columns 0..V-1 velocity logits, V..V+2 onset / offset / mpe logits; fp32 outputs
  float* onset;
  float* offset;
  float* mpe;
  float* velocity;
  signed char* vel_argmax; // [rows] argmax over the n_vel velocity logits (first maximum), or NULL
  int n_vel;
  int time_major;         // rows are (b, note, frame): permute back to [B, frame, note]
  int n_frame, n_note;
  int stage_rows;         // rows of one epilogue store box (= box rows of map_o): 32 (default when 0), 16 or 8 -- smaller boxes leave
                          // shared memory to a resident W / a deeper A ring; a warp then stores its 32 rows in 32 / stage_rows passes
  int debug_flags;        // experiments only (HFT_TC_DEBUG): 1 = skip the TMA stores, 2 = skip the epilogue arithmetic
};

// n_rows_w: W rows staged per CTA (n_tile, or n_tile / 2 in PAIR mode)
__host__ __device__ constexpr size_t gemm_smem_bytes(int n_rows_w, int k_chunks, int w_resident, int a_stages, int w_stages, int x3, int has_resid,
                                                     int stage_rows = 32) {
  size_t w = w_resident ? (size_t)n_rows_w * kBlockK * 2 * k_chunks * (x3 ? 2 : 1) : (size_t)w_stages * n_rows_w * kBlockK * 2;
  return 1024 /*align*/ + w + (size_t)a_stages * kChunkA + (has_resid ? 8192 : 0) /*I64*/ + (size_t)kEpiWarps * stage_rows * 128 /*store staging*/ +
         kConstBytes + 512 /*barriers*/;
}

// NT = n_tile: UMMA N and the number of accumulator columns every epilogue thread walks.  map_o: box 64 x 32 (one warp's rows).
template <bool BF16, int EPI, int NT, bool PAIR>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_r,
            const __grid_constant__ CUtensorMap map_o, const __grid_constant__ GemmParams p) {
  constexpr int n_tile = NT;
  constexpr int n_rows_w = PAIR ? n_tile / 2 : n_tile;                   // W rows staged by this CTA
  constexpr uint32_t w_chunk = (uint32_t)n_rows_w * kBlockK * 2;
  constexpr uint32_t kCtas = PAIR ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const bool x3p = p.x3 && !p.single;                                    // three products (A_hi W_hi + A_lo W_hi + A_hi W_lo)
  const int kc_w = p.k_chunks * (x3p ? 2 : 1);                           // W chunks held when resident
  uint8_t* s_w = smem;                                                   // resident W: [kc_w][n_tile x 64]; else the W ring
  uint8_t* s_a = s_w + (p.w_resident ? (size_t)kc_w * w_chunk : (size_t)p.w_stages * w_chunk);   // A ring: a_stages x 16 KB
  uint8_t* s_i64 = s_a + (size_t)p.a_stages * kChunkA;                   // 64 x 64 identity, K-major SW128 (only with a residual)
  const int parts = p.x3 ? 2 : 1;
  uint8_t* s_out = s_i64 + (p.has_resid ? 8192 : 0);                     // [8 warps][32 x 64] staging (x3: hi, then lo through the same block)
  const int stage_rows = p.stage_rows ? p.stage_rows : 32;
  float* s_const = reinterpret_cast<float*>(s_out + (size_t)kEpiWarps * stage_rows * 128);   // bias[NT] | gamma[NT] | beta[NT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_const) + kConstBytes);
  uint64_t* full_a = bars;               // [8]
  uint64_t* empty_a = bars + 8;          // [8]
  uint64_t* full_w = bars + 16;          // [8]
  uint64_t* empty_w = bars + 24;         // [8]
  uint64_t* tfull = bars + 32;           // [2]
  uint64_t* tempty = bars + 34;          // [2]
  uint64_t* wbar = bars + 36;            // resident W landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 37);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // persistent schedule: this CTA (pair) owns n-tile `nt` and walks m-tiles mt0, mt0 + m_step, ...
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  const bool leader = rank == 0;
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int nt = unit % p.n_tiles;
  const int mt0 = unit / p.n_tiles;
  const int m_step = units / p.n_tiles;
  auto row_tile = [&](int mt) { return PAIR ? 2 * mt + (int)rank : mt; };   // this CTA's 128-row tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_w); tma_prefetch_desc(&map_o);
    if (p.has_resid) tma_prefetch_desc(&map_r);
    for (int i = 0; i < p.a_stages; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < p.w_stages; ++i) { mbar_init(&full_w[i], 1); mbar_init(&empty_w[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kCtas * (kEpiWarps / 2)); }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { if (PAIR) tmem_alloc2(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
  if (p.has_resid) {                     // identity block for the residual MMA: I[n][k] = (n == k), swizzled like a TMA tile
    const uint16_t one = BF16 ? 0x3F80 : 0x3C00;   // PAIR: this CTA holds rows [32 rank, 32 rank + 32) of the 64 x 64 identity
    const int n_off = PAIR ? 32 * (int)rank : 0;
    for (int i = threadIdx.x; i < (PAIR ? 32 : 64) * 64; i += kGemmThreads) {
      int n = i >> 6, k = i & 63;
      uint32_t off = n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
      *reinterpret_cast<uint16_t*>(s_i64 + off) = (n + n_off == k) ? one : (uint16_t)0;
    }
    fence_proxy_async();
  }
  for (int i = threadIdx.x; i < NT; i += kGemmThreads) {   // per-column constants of this CTA's slice
    s_const[i] = p.bias[nt * NT + i];
    if (EPI == EPI_LN) { s_const[NT + i] = p.gamma[i]; s_const[2 * NT + i] = p.beta[i]; }
  }
  fence_before_sync();
  __syncthreads();
  if (PAIR) cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / multicast commit
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int resid_chunks = p.has_resid ? (n_tile / 64) * (p.x3 ? 2 : 1) : 0;

  // Per k-chunk the operand schedule is (x3):  Wh, Ah, Al, Wl  ->  MMA(Ah,Wh)  MMA(Al,Wh) [free Al, Wh]  MMA(Ah,Wl) [free Ah, Wl]
  // (single product):                          W,  A           ->  MMA(A,W)   [free A, W]
  // so every A chunk is fetched once per tile and every streamed W chunk once per tile.
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // PAIR: the loads of both CTAs are counted on the LEADER's barriers (the leader alone waits on them and issues the MMAs)
      auto load = [&](void* dst, const CUtensorMap* map, int col, int row, uint64_t* bar) {
        if (PAIR) tma_load_2d_2cta(dst, map, col, row, map_to_rank(bar, 0)); else tma_load_2d(dst, map, col, row, bar);
      };
      const int w_row = nt * n_tile + (int)rank * n_rows_w;
      if (p.w_resident) {
        if (leader) mbar_expect_tx(wbar, kCtas * (uint32_t)kc_w * w_chunk);
        for (int c = 0; c < kc_w; ++c)
          load(s_w + (size_t)c * w_chunk, &map_w, (c < p.k_chunks ? c : c - p.k_chunks) * kBlockK + (c < p.k_chunks ? 0 : p.w_lo_off), w_row, wbar);
      }
      int sa = 0, sw = 0;
      uint32_t pa = 0, pw = 0;
      auto load_a = [&](const CUtensorMap* map, int col, int row) {
        mbar_wait(&empty_a[sa], pa ^ 1);
        if (leader) mbar_expect_tx(&full_a[sa], kCtas * kChunkA);
        load(s_a + (size_t)sa * kChunkA, map, col, row, &full_a[sa]);
        if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
      };
      auto load_w = [&](int col) {
        mbar_wait(&empty_w[sw], pw ^ 1);
        if (leader) mbar_expect_tx(&full_w[sw], kCtas * w_chunk);
        load(s_w + (size_t)sw * w_chunk, &map_w, col, w_row, &full_w[sw]);
        if (++sw == p.w_stages) { sw = 0; pw ^= 1; }
      };
      for (int mt = mt0; mt < p.m_tiles; mt += m_step) {
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          if (!p.w_resident) load_w(kc * kBlockK);
          load_a(&map_a, kc * kBlockK, row_tile(mt) * kBlockM);
          if (x3p) {
            load_a(&map_a, p.a_lo_off + kc * kBlockK, row_tile(mt) * kBlockM);
            if (!p.w_resident) load_w(p.w_lo_off + kc * kBlockK);
          }
        }
        const int r_row = (p.resid_period > 0 ? (row_tile(mt) % p.resid_period) : row_tile(mt)) * kBlockM;
        for (int rc = 0; rc < resid_chunks; ++rc)                        // residual tile (hi chunks, then lo chunks)
          load_a(&map_r, (rc % (n_tile / 64)) * 64 + (rc >= n_tile / 64 ? p.out_lo_off : 0), r_row);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (leader) {                                       // the whole warp runs the loop (uniform control flow: descriptors stay in uniform registers); one elected lane issues
      const uint32_t idesc = make_idesc(kCtas * kBlockM, n_tile, BF16, false, false);
      const uint32_t idesc64 = make_idesc(kCtas * kBlockM, 64, BF16, false, false);
      auto commit = [&](uint64_t* bar) { if (elect_one()) { if (PAIR) umma_commit_2cta(bar); else umma_commit(bar); } __syncwarp(); };
      const uint32_t i64_addr = smem_u32(s_i64);
      if (p.w_resident) { mbar_wait(wbar, 0); fence_after_sync(); }
      int sa = 0, sw = 0;
      uint32_t pa = 0, pw = 0;
      int it = 0;
      uint32_t first = 1;
      const SDescBase kd = sdesc_base(16, 1024, kSwz128);              // the issuing thread pays one add per descriptor (K step = 32 bytes = +2)
      auto mma4 = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t id) {
        const uint32_t a0 = sdesc_lo(kd, a_addr), b0 = sdesc_lo(kd, b_addr);
        const uint32_t acc0 = first ? 0u : 1u;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            if (PAIR) umma_f16_2cta_lohi(d, a0 + 2 * k, kd.hi, b0 + 2 * k, kd.hi, id, k ? 1u : acc0);
            else umma_f16_lohi(d, a0 + 2 * k, kd.hi, b0 + 2 * k, kd.hi, id, k ? 1u : acc0);
          }
        }
        __syncwarp();
        first = 0;
      };
      auto wait_a = [&]() -> uint32_t {                  // next A-ring slot, filled
        mbar_wait(&full_a[sa], pa);
        fence_after_sync();
        return smem_u32(s_a + (size_t)sa * kChunkA);
      };
      auto free_a = [&]() {                              // release the oldest held A slot once the MMAs issued so far retire
        commit(&empty_a[sa]);
        if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
      };
      auto wait_w = [&]() -> uint32_t {
        mbar_wait(&full_w[sw], pw);
        fence_after_sync();
        return smem_u32(s_w + (size_t)sw * w_chunk);
      };
      auto free_w = [&]() {
        commit(&empty_w[sw]);
        if (++sw == p.w_stages) { sw = 0; pw ^= 1; }
      };
      for (int mt = mt0; mt < p.m_tiles; mt += m_step, ++it) {
        const int ab = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty[ab], aphase ^ 1);             // epilogue has drained this accumulator buffer
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + ab * 256;
        first = 1;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          const uint32_t wh = p.w_resident ? smem_u32(s_w + (size_t)kc * w_chunk) : wait_w();
          const uint32_t ah = wait_a();
          mma4(d_tmem, ah, wh, idesc);
          if (!x3p) {
            free_a();
            if (!p.w_resident) free_w();
          } else {
            // the A ring is consumed in order: Ah occupies slot sa, Al slot sa+1
            const int sa_hi = sa;
            const uint32_t pa_hi = pa;
            if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
            const uint32_t al = wait_a();
            mma4(d_tmem, al, wh, idesc);
            commit(&empty_a[sa]);                  // Al done
            if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
            if (!p.w_resident) free_w();                // Wh done
            const uint32_t wl = p.w_resident ? smem_u32(s_w + (size_t)(p.k_chunks + kc) * w_chunk) : wait_w();
            mma4(d_tmem, ah, wl, idesc);
            commit(&empty_a[sa_hi]);               // Ah done
            (void)pa_hi;
            if (!p.w_resident) free_w();                // Wl done
          }
        }
        for (int rc = 0; rc < resid_chunks; ++rc) {     // acc[:, 64j .. 64j+63] += R[:, chunk] * I64
          const uint32_t a_addr = wait_a();
          const uint32_t col = (uint32_t)(rc % (n_tile / 64)) * 64;
          mma4(d_tmem + col, a_addr, i64_addr, idesc64);
          free_a();
        }
        commit(&tfull[ab]);                        // accumulator complete
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;
    const int wg = ew >> 2;                             // accumulator buffer (= tile parity) drained by this warpgroup
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int row_in_tile = quarter * 32 + lane;
    uint8_t* my_stage = s_out + (size_t)ew * stage_rows * 128;
    const float* s_bias = s_const;
    const float* s_gamma = s_const + NT;
    const float* s_beta = s_const + 2 * NT;
    const bool x3 = p.x3 != 0;
    // one 32-row x 64-column block of this warp: registers -> swizzled staging -> TMA store (hi, and lo in x3 mode)
    // (one 4 KB block per warp: in split mode the lo half follows the hi half through it, which leaves the shared memory to the
    //  operand rings / a resident W; the TMA read of the block overlaps the arithmetic of the next one)
    auto store_block = [&](const uint32_t (&pk)[32], const uint32_t (&pl)[32], int col, int row0) {
      uint8_t* dst = my_stage + (lane & (stage_rows - 1)) * 128;
      for (int part = 0; part < parts; ++part) {
        for (int pass = 0; pass * stage_rows < 32; ++pass) {   // one pass with the default 32-row boxes
          if (lane == 0) tma_store_wait_read();         // the previous store of this warp has read the staging block
          __syncwarp();
          if (lane / stage_rows == pass) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<uint4*>(dst + ((q ^ (lane & 7)) << 4)) =
                  part ? make_uint4(pl[4 * q], pl[4 * q + 1], pl[4 * q + 2], pl[4 * q + 3]) : make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && !(p.debug_flags & 1)) {
            tma_store_2d(&map_o, my_stage, col + (part ? p.out_lo_off : 0), row0 + pass * stage_rows);
            tma_store_commit();
          }
        }
      }
    };
    // accumulator buffer drained: tell the MMA issuer (PAIR: the leader CTA's barrier collects both CTAs' warps)
    auto release_acc = [&]() { if (PAIR) mbar_arrive_cluster(map_to_rank(&tempty[wg], 0)); else mbar_arrive(&tempty[wg]); };
    int it = 0;
    for (int mt = mt0; mt < p.m_tiles; mt += m_step, ++it) {
      if ((it & 1) != wg) continue;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull[wg], aphase);
      fence_after_sync();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + wg * 256;
      const long long row = (long long)row_tile(mt) * kBlockM + row_in_tile;
      const int row0 = row_tile(mt) * kBlockM + quarter * 32;   // first row of this warp's store boxes
      const int col_base = p.out_col0 + nt * NT;                // first output column of this CTA's slice

      if (EPI == EPI_STORE || EPI == EPI_RELU) {
#pragma unroll 1
        for (int blk = 0; blk < NT / 64; ++blk) {
          uint32_t pk[32], pl[32];
          uint32_t r0[32], r1[32];
          tmem_ld32(t_row + blk * 64, r0);
          tmem_ld32(t_row + blk * 64 + 32, r1);
          tmem_ld_wait();
          if (blk == NT / 64 - 1) {                             // accumulator fully read: hand the TMEM buffer back
            fence_before_sync();
            __syncwarp();
            if (lane == 0) release_acc();
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float4* bp = reinterpret_cast<const float4*>(s_bias + blk * 64 + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = bp[j];
              const uint32_t* r = c ? r1 : r0;
              float v0 = __uint_as_float(r[4 * j]) + b4.x, v1 = __uint_as_float(r[4 * j + 1]) + b4.y;
              float v2 = __uint_as_float(r[4 * j + 2]) + b4.z, v3 = __uint_as_float(r[4 * j + 3]) + b4.w;
              if (EPI == EPI_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
              if (x3) {
                split_pack<BF16>(v0, v1, pk[c * 16 + 2 * j], pl[c * 16 + 2 * j]);
                split_pack<BF16>(v2, v3, pk[c * 16 + 2 * j + 1], pl[c * 16 + 2 * j + 1]);
              } else {
                pk[c * 16 + 2 * j] = Op16<BF16>::pack(v0, v1);
                pk[c * 16 + 2 * j + 1] = Op16<BF16>::pack(v2, v3);
              }
            }
          }
          store_block(pk, pl, col_base + blk * 64, row0);
        }
        continue;                                               // tempty already signalled
      } else if (EPI == EPI_LN) {
        // v = acc (+ residual already accumulated by the identity MMA) + bias; LayerNorm over the row of NT values.
        // Pass 1 over TMEM: shifted one-pass statistics (shift = the mean of the row's first 32 values, so E[d^2] - E[d]^2 does not cancel).
        float shift = 0.f, sum = 0.f, sq = 0.f;
#pragma unroll 1
        for (int c = 0; c < NT / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tmem_ld_wait();
          const float4* bp = reinterpret_cast<const float4*>(s_bias + c * 32);
          if (c == 0) {                                       // shift = mean of the row's first 32 values (close to the row mean: no cancellation)
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) t += __uint_as_float(r[j]) + s_bias[j];
            shift = t * (1.f / 32.f);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = bp[j];
            const float d0 = (__uint_as_float(r[4 * j]) + b4.x) - shift, d1 = (__uint_as_float(r[4 * j + 1]) + b4.y) - shift;
            const float d2 = (__uint_as_float(r[4 * j + 2]) + b4.z) - shift, d3 = (__uint_as_float(r[4 * j + 3]) + b4.w) - shift;
            sum += (d0 + d1) + (d2 + d3);
            sq = fmaf(d0, d0, sq); sq = fmaf(d1, d1, sq); sq = fmaf(d2, d2, sq); sq = fmaf(d3, d3, sq);
          }
        }
        const float md = sum * (1.f / (float)NT);
        const float var = fmaxf(sq * (1.f / (float)NT) - md * md, 0.f);
        const float rstd = rsqrtf(var + 1e-5f);
        const float nmr = -(shift + md) * rstd;
        // Pass 2 over TMEM: normalise, scale, split, store
#pragma unroll 1
        for (int blk = 0; blk < NT / 64; ++blk) {
          uint32_t pk[32], pl[32];
          uint32_t r0[32], r1[32];
          tmem_ld32(t_row + blk * 64, r0);
          tmem_ld32(t_row + blk * 64 + 32, r1);
          tmem_ld_wait();
          if (blk == NT / 64 - 1) {
            fence_before_sync();
            __syncwarp();
            if (lane == 0) release_acc();
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float4* bp = reinterpret_cast<const float4*>(s_bias + blk * 64 + c * 32);
            const float4* gp = reinterpret_cast<const float4*>(s_gamma + blk * 64 + c * 32);
            const float4* ep = reinterpret_cast<const float4*>(s_beta + blk * 64 + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = bp[j], g4 = gp[j], e4 = ep[j];
              const uint32_t* r = c ? r1 : r0;
              const float y0 = fmaf(fmaf(__uint_as_float(r[4 * j]) + b4.x, rstd, nmr), g4.x, e4.x);
              const float y1 = fmaf(fmaf(__uint_as_float(r[4 * j + 1]) + b4.y, rstd, nmr), g4.y, e4.y);
              const float y2 = fmaf(fmaf(__uint_as_float(r[4 * j + 2]) + b4.z, rstd, nmr), g4.z, e4.z);
              const float y3 = fmaf(fmaf(__uint_as_float(r[4 * j + 3]) + b4.w, rstd, nmr), g4.w, e4.w);
              if (x3) {
                split_pack<BF16>(y0, y1, pk[c * 16 + 2 * j], pl[c * 16 + 2 * j]);
                split_pack<BF16>(y2, y3, pk[c * 16 + 2 * j + 1], pl[c * 16 + 2 * j + 1]);
              } else {
                pk[c * 16 + 2 * j] = Op16<BF16>::pack(y0, y1);
                pk[c * 16 + 2 * j + 1] = Op16<BF16>::pack(y2, y3);
              }
            }
          }
          store_block(pk, pl, col_base + blk * 64, row0);
        }
        continue;                                               // tempty already signalled
      } else {                                          // EPI_HEADS (direct fp32 stores)
        long long orow_idx = row;
        if (p.time_major) {
          int f = (int)(row % p.n_frame);
          long long bn = row / p.n_frame;
          int n = (int)(bn % p.n_note);
          orow_idx = ((bn / p.n_note) * p.n_frame + f) * p.n_note + n;
        }
        float best = -INFINITY;
        int best_i = 0;
#pragma unroll 1
        for (int c = 0; c < NT / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tmem_ld_wait();
          const int c0 = c * 32;
          if (c0 + 32 <= p.n_vel) {
            if (p.vel_argmax) {                                 // running argmax of the row's velocity logits (strict >: first maximum)
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(r[j]) + s_bias[c0 + j];
                if (v > best) { best = v; best_i = c0 + j; }
              }
            }
            if (p.velocity) {
              float4* dst = reinterpret_cast<float4*>(p.velocity + orow_idx * p.n_vel + c0);
              const float4* bp = reinterpret_cast<const float4*>(s_bias + c0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b4 = bp[j];
                dst[j] = make_float4(__uint_as_float(r[4 * j]) + b4.x, __uint_as_float(r[4 * j + 1]) + b4.y, __uint_as_float(r[4 * j + 2]) + b4.z,
                                     __uint_as_float(r[4 * j + 3]) + b4.w);
              }
            }
          } else if (c0 == p.n_vel) {
            float* dsts[3] = {p.onset, p.offset, p.mpe};
#pragma unroll
            for (int j = 0; j < 3; ++j)
              if (dsts[j]) dsts[j][orow_idx] = 1.f / (1.f + expf(-(__uint_as_float(r[j]) + s_bias[c0 + j])));
          }
        }
        if (p.vel_argmax) p.vel_argmax[orow_idx] = (signed char)best_i;
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) release_acc();
    }
    if (lane == 0) tma_store_wait_all();                // all stores landed before exit
  }
  fence_before_sync();
  __syncthreads();
  if (PAIR) cluster_sync_all();          // no CTA of the pair leaves while the other may still signal its barriers / read its operands
  if (warp == 1) { if (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

}  // namespace tc
}  // namespace hft
