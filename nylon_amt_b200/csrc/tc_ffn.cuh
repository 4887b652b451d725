// Fused position-wise feed-forward block on tcgen05 CTA pairs:   y = LayerNorm(x + fc_2(relu(fc_1(x))))
// (reference PositionwiseFeedforwardLayer.forward model_spec2midi.py:369-378 + the residual/LayerNorm of :242 / :271 / :305).
//
// Unfused, the block moves x in, the [rows, pf_dim] hidden activation out and back in, the residual in and y out: in split
// (hi|lo) mode 896 KB per 128-row tile.  Here the hidden activation never leaves the SM: per 128-row tile (256 rows per CTA
// pair, cta_group::2) the x tile is loaded ONCE (it is both the GEMM operand and the residual) and y is stored once: 256 KB.
//
//   for each block j of 128 hidden units:
//     S_j  = x W1_j^T                      UMMA M256 x N128, K = hid_dim, accumulator = one of two 128-column TMEM buffers
//     P_j  = split16(relu(S_j + b1_j))     epilogue warpgroup j & 1, written IN PLACE over S_j ([hi 16 | lo 16] per 32 units)
//     Y   += P_j W2[:, j]^T                UMMA M256 x N256 with A = P_j read straight from TMEM
//   Y (+= x through an identity MMA at the start of the tile) -> + b2 -> LayerNorm -> split16 -> per-warp TMA stores
//
// GEMM1 of block j+1 overlaps the epilogue of block j; the W1 / W2 chunks of a CTA (half of the rows each: the tensor cores
// of the pair exchange the B halves) stream through one ring.  Warps: 0 producer, 1 MMA issuer (leader CTA only), 2..5 and
// 6..9 two epilogue warpgroups (thread = tile row).  Built for hid_dim 256 / pf_dim 512 (the paper-size model).
#pragma once
#include "tc_gemm.cuh"

namespace hft {
namespace tc {

struct FfnParams {
  int m_tiles;              // M / 256 (pair tiles)
  int x3;                   // split (hi | lo) layout of x / y / W1 / W2
  int single;               // x3 layout, single product: only x_hi W1_hi and P_hi W2_hi are accumulated (the lo halves of W1 / W2 are not loaded);
                            // the residual still adds x_hi + x_lo and y is stored hi | lo
  int lo_off;               // column distance between the hi and lo halves of x / y (= hid_dim)
  int w1_lo_off, w2_lo_off; // same for W1 ([pf, 2 hid]) and W2 ([hid, 2 pf])
  const float* b1;          // [pf]
  const float* b2;          // [hid]
  const float* gamma;
  const float* beta;
  int slots;                // 0 = ffn_slots(x3); smaller values only (the shared-memory layout is sized for the default)
};

constexpr int kFfnH = 256, kFfnP = 512, kFfnJB = 128, kFfnNJ = kFfnP / kFfnJB, kFfnKC = kFfnH / 64;
constexpr int kFfnSlot = 16384;
__host__ __device__ constexpr int ffn_slots(int x3) { return x3 ? 4 : 8; }   // split mode: the 128 KB x tile leaves room for 4

__host__ __device__ constexpr size_t ffn_smem_bytes(int x3) {
  return 1024 + (size_t)kFfnKC * (x3 ? 2 : 1) * kChunkA /*x tile*/ + (size_t)ffn_slots(x3) * kFfnSlot /*W ring*/ + 4096 /*I64 half*/ +
         4 * kWarpStage /*store staging*/ + (kFfnP + 3 * kFfnH) * 4 /*b1 b2 gamma beta*/ + 512;
}

template <bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
ffn_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2,
           const __grid_constant__ CUtensorMap map_o, const __grid_constant__ FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int parts = p.x3 ? 2 : 1;
  const bool x3p = p.x3 && !p.single;                                    // three products
  const int wparts = x3p ? 2 : 1;                                        // W1 / W2 halves streamed
  const int kFfnSlots = p.slots > 0 ? p.slots : ffn_slots(p.x3);   // p.slots < default: ring-depth experiment (HFT_TC_FFN_SLOTS)
  uint8_t* s_x = smem;                                                   // [parts][KC] chunks of 128 x 64
  uint8_t* s_ring = s_x + (size_t)parts * kFfnKC * kChunkA;              // [slots] x 16 KB
  uint8_t* s_i64 = s_ring + (size_t)kFfnSlots * kFfnSlot;                // 32 x 64 half identity
  uint8_t* s_out = s_i64 + 4096;                                         // [4 warps] x 4 KB (hi, then lo)
  float* s_b1 = reinterpret_cast<float*>(s_out + 4 * kWarpStage);
  float* s_b2 = s_b1 + kFfnP;
  float* s_g = s_b2 + kFfnH;
  float* s_be = s_g + kFfnH;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_be + kFfnH);
  uint64_t* rfull = bars;                // [8]
  uint64_t* rempty = bars + 8;           // [8]
  uint64_t* x_full = bars + 16;
  uint64_t* x_empty = bars + 17;
  uint64_t* s_ready = bars + 18;         // [2]
  uint64_t* p_ready = bars + 20;         // [2]  (leader: 8 arrivals = 4 warps x 2 CTAs)
  uint64_t* y_ready = bars + 22;
  uint64_t* y_free = bars + 23;          // (leader: 8 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int unit = (int)(blockIdx.x >> 1), units = (int)(gridDim.x >> 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_w1); tma_prefetch_desc(&map_w2); tma_prefetch_desc(&map_o);
    for (int i = 0; i < kFfnSlots; ++i) { mbar_init(&rfull[i], 1); mbar_init(&rempty[i], 1); }
    mbar_init(x_full, 1); mbar_init(x_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_ready[i], 1); mbar_init(&p_ready[i], 8); }
    mbar_init(y_ready, 1); mbar_init(y_free, 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  {
    const uint16_t one = BF16 ? 0x3F80 : 0x3C00;
    for (int i = threadIdx.x; i < 32 * 64; i += kGemmThreads) {
      int n = i >> 6, k = i & 63;
      uint32_t off = n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
      *reinterpret_cast<uint16_t*>(s_i64 + off) = (n + 32 * (int)rank == k) ? one : (uint16_t)0;
    }
    fence_proxy_async();
  }
  for (int i = threadIdx.x; i < kFfnP; i += kGemmThreads) s_b1[i] = p.b1[i];
  for (int i = threadIdx.x; i < kFfnH; i += kGemmThreads) { s_b2[i] = p.b2[i]; s_g[i] = p.gamma[i]; s_be[i] = p.beta[i]; }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kYCol = 0, kSCol = 256;                             // Y: 256 columns; S/P buffers: 2 x 128 columns

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; bytes counted on the leader's barriers) =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0, xph = 0;
      auto ring_next = [&](uint32_t bytes) -> uint8_t* {
        mbar_wait(&rempty[slot], ph ^ 1);
        if (leader) mbar_expect_tx(&rfull[slot], 2 * bytes);
        return s_ring + (size_t)slot * kFfnSlot;
      };
      auto ring_adv = [&]() { if (++slot == kFfnSlots) { slot = 0; ph ^= 1; } };
      for (int mt = unit; mt < p.m_tiles; mt += units) {
        const int row = (2 * mt + (int)rank) * kBlockM;
        mbar_wait(x_empty, xph ^ 1);
        if (leader) mbar_expect_tx(x_full, 2u * (uint32_t)(parts * kFfnKC) * kChunkA);
        for (int part = 0; part < parts; ++part)
          for (int kc = 0; kc < kFfnKC; ++kc)
            tma_load_2d_2cta(s_x + (size_t)(part * kFfnKC + kc) * kChunkA, &map_x, part * p.lo_off + kc * 64, row, map_to_rank(x_full, 0));
        xph ^= 1;
        for (int j = 0; j < kFfnNJ; ++j) {
          // GEMM1_j operands: per pair of k-chunks W1 hi, then W1 lo (64 rows of this CTA's half, 2 x 64 columns per slot)
          for (int u = 0; u < kFfnKC / 2; ++u)
            for (int part = 0; part < wparts; ++part) {
              uint8_t* dst = ring_next(2 * 8192);
              for (int h = 0; h < 2; ++h)
                tma_load_2d_2cta(dst + h * 8192, &map_w1, part * p.w1_lo_off + (2 * u + h) * 64, j * kFfnJB + (int)rank * 64, map_to_rank(&rfull[slot], 0));
              ring_adv();
            }
          // GEMM2_{j-1} operands (issued after GEMM1_j): W2 hi / lo chunks of 64 hidden units, 128 rows of this CTA's half
          const int jj = j - 1;
          if (jj >= 0)
            for (int kc2 = 0; kc2 < 2; ++kc2)
              for (int part = 0; part < wparts; ++part) {
                uint8_t* dst = ring_next(16384);
                tma_load_2d_2cta(dst, &map_w2, part * p.w2_lo_off + jj * kFfnJB + kc2 * 64, (int)rank * 128, map_to_rank(&rfull[slot], 0));
                ring_adv();
              }
        }
        for (int kc2 = 0; kc2 < 2; ++kc2)                                 // GEMM2 of the last block
          for (int part = 0; part < wparts; ++part) {
            uint8_t* dst = ring_next(16384);
            tma_load_2d_2cta(dst, &map_w2, part * p.w2_lo_off + (kFfnNJ - 1) * kFfnJB + kc2 * 64, (int)rank * 128, map_to_rank(&rfull[slot], 0));
            ring_adv();
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (leader) {                                       // whole warp, uniform control flow; one elected lane issues MMAs and commits
      const uint32_t id1 = make_idesc(256, kFfnJB, BF16, false, false);
      const uint32_t id2 = make_idesc(256, kFfnH, BF16, false, false);
      const uint32_t id64 = make_idesc(256, 64, BF16, false, false);
      const uint32_t i64_addr = smem_u32(s_i64);
      int slot = 0;
      uint32_t ph = 0, xph = 0, pph[2] = {0, 0}, yph = 0;
      auto take = [&]() -> int {
        mbar_wait(&rfull[slot], ph);
        fence_after_sync();
        const int s = slot;
        if (++slot == kFfnSlots) { slot = 0; ph ^= 1; }
        return s;
      };
      const SDescBase kd = sdesc_base(16, 1024, kSwz128);              // the issuing thread pays one add per descriptor (K step = 32 bytes = +2)
      auto ss4 = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t id, uint32_t& acc) {   // 64 K elements: 4 K-steps, both operands in smem
        const uint32_t a0 = sdesc_lo(kd, a_addr), b0 = sdesc_lo(kd, b_addr);
        const uint32_t acc0 = acc;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_2cta_lohi(d, a0 + 2 * k, kd.hi, b0 + 2 * k, kd.hi, id, k ? 1u : acc0);
        }
        __syncwarp();
        acc = 1;
      };
      auto commit = [&](uint64_t* bar) { if (elect_one()) umma_commit_2cta(bar); __syncwarp(); };
      auto gemm2 = [&](int j) {                                          // Y += P_j W2[:, j]^T, A = P_j from TMEM
        const int b = j & 1;
        mbar_wait(&p_ready[b], pph[b]); pph[b] ^= 1;
        fence_after_sync();
        const uint32_t tp = tmem_base + kSCol + b * 128;
        for (int kc2 = 0; kc2 < 2; ++kc2) {
          const int sh = take();
          const uint32_t wh = sdesc_lo(kd, smem_u32(s_ring + (size_t)sh * kFfnSlot));
          if (elect_one()) {
            for (int part = 0; part < wparts; ++part)                    // Ph W2h, Pl W2h
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int ks = kc2 * 4 + k;
                const uint32_t pcol = p.x3 ? (uint32_t)((ks >> 1) * 32 + part * 16 + (ks & 1) * 8) : (uint32_t)(ks * 8);
                umma_f16_ts_2cta_lohi(tmem_base + kYCol, tp + pcol, wh + 2 * k, kd.hi, id2, 1u);
              }
          }
          __syncwarp();
          commit(&rempty[sh]);
          if (x3p) {
            const int sl = take();
            const uint32_t wl = sdesc_lo(kd, smem_u32(s_ring + (size_t)sl * kFfnSlot));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {                               // Ph W2l
                const int ks = kc2 * 4 + k;
                const uint32_t pcol = (uint32_t)((ks >> 1) * 32 + (ks & 1) * 8);
                umma_f16_ts_2cta_lohi(tmem_base + kYCol, tp + pcol, wl + 2 * k, kd.hi, id2, 1u);
              }
            }
            __syncwarp();
            commit(&rempty[sl]);
          }
        }
      };
      int it = 0;
      for (int mt = unit; mt < p.m_tiles; mt += units, ++it) {
        mbar_wait(x_full, xph); xph ^= 1;
        fence_after_sync();
        auto residual = [&]() {
          mbar_wait(y_free, yph ^ 1);                                     // the previous tile's LayerNorm epilogue has drained Y
          yph ^= 1;
          fence_after_sync();
          // Y = x (residual) through the identity: column block cb of Y <- x chunk cb (hi, then lo)
          for (int cb = 0; cb < kFfnKC; ++cb) {
            uint32_t acc = 0;
            for (int part = 0; part < parts; ++part)
              ss4(tmem_base + kYCol + cb * 64, smem_u32(s_x + (size_t)(part * kFfnKC + cb) * kChunkA), i64_addr, id64, acc);
          }
        };
        for (int j = 0; j < kFfnNJ; ++j) {
          const uint32_t d = tmem_base + kSCol + (j & 1) * 128;
          uint32_t acc = 0;
          for (int u = 0; u < kFfnKC / 2; ++u) {
            const int sh = take();
            const uint32_t wh = smem_u32(s_ring + (size_t)sh * kFfnSlot);
            for (int h = 0; h < 2; ++h) {
              const int kc = 2 * u + h;
              ss4(d, smem_u32(s_x + (size_t)kc * kChunkA), wh + h * 8192, id1, acc);                       // xh W1h
              if (x3p) ss4(d, smem_u32(s_x + (size_t)(kFfnKC + kc) * kChunkA), wh + h * 8192, id1, acc);   // xl W1h
            }
            commit(&rempty[sh]);
            if (x3p) {
              const int sl = take();
              const uint32_t wl = smem_u32(s_ring + (size_t)sl * kFfnSlot);
              for (int h = 0; h < 2; ++h) ss4(d, smem_u32(s_x + (size_t)(2 * u + h) * kChunkA), wl + h * 8192, id1, acc);   // xh W1l
              commit(&rempty[sl]);
            }
          }
          commit(&s_ready[j & 1]);
          if (j == kFfnNJ - 1) commit(x_empty);                 // x is dead once GEMM1 of the last block has completed
          if (j == 1) residual();                                         // after two GEMM1 blocks: the previous tile's epilogue overlaps them
          if (j >= 1) gemm2(j - 1);
        }
        gemm2(kFfnNJ - 1);
        commit(y_ready);
      }
    }
  } else {
    // ===================== epilogue warpgroups =====================
    const int ew = warp - 2;
    const int wg = ew >> 2;
    const int quarter = warp & 3;
    const bool x3 = p.x3 != 0;
    uint8_t* my_stage = s_out + (size_t)(ew & 3) * kWarpStage;
    uint32_t sph = 0;
    int it = 0;
    for (int mt = unit; mt < p.m_tiles; mt += units, ++it) {
      // hidden blocks j = wg, wg + 2: S_j + b1 -> ReLU -> 16-bit (hi | lo) in place
      for (int j = wg; j < kFfnNJ; j += 2) {
        mbar_wait(&s_ready[wg], sph); sph ^= 1;
        fence_after_sync();
        const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + kSCol + wg * 128;
#pragma unroll 1
        for (int c = 0; c < kFfnJB / 32; ++c) {
          uint32_t v[32], pw[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
          const float4* bp = reinterpret_cast<const float4*>(s_b1 + j * kFfnJB + c * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b4 = bp[q];
            const float h0 = fmaxf(__uint_as_float(v[4 * q]) + b4.x, 0.f), h1 = fmaxf(__uint_as_float(v[4 * q + 1]) + b4.y, 0.f);
            const float h2 = fmaxf(__uint_as_float(v[4 * q + 2]) + b4.z, 0.f), h3 = fmaxf(__uint_as_float(v[4 * q + 3]) + b4.w, 0.f);
            if (x3) {
              split_pack<BF16>(h0, h1, pw[2 * q], pw[16 + 2 * q]);
              split_pack<BF16>(h2, h3, pw[2 * q + 1], pw[16 + 2 * q + 1]);
            } else {
              pw[2 * q] = Op16<BF16>::pack(h0, h1);
              pw[2 * q + 1] = Op16<BF16>::pack(h2, h3);
            }
          }
          if (x3) {
            tmem_st32(t_row + c * 32, pw);
          } else {
            uint32_t ph16[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) ph16[q] = pw[q];
            tmem_st16(t_row + c * 16, ph16);
          }
        }
        tmem_st_wait();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(map_to_rank(&p_ready[wg], 0));
      }
      // final epilogue of this tile: warpgroup (it & 1).  y_ready completes once per tile; tile `it` is completion number `it`,
      // and this warpgroup cannot reach the wait before completion it - 1 happened (its hidden blocks of tile `it` come after it).
      if ((it & 1) != wg) continue;
      mbar_wait(y_ready, (uint32_t)(it & 1));
      fence_after_sync();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + kYCol;
      const int row0 = (2 * mt + (int)rank) * kBlockM + quarter * 32;
      // LayerNorm over the row of 256 values (x + fc_2(...) already summed in the accumulator): shifted one-pass statistics
      float shift = 0.f, sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int c = 0; c < kFfnH / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_ld_wait();
        const float4* bp = reinterpret_cast<const float4*>(s_b2 + c * 32);
        if (c == 0) {                                       // shift = mean of the row's first 32 values (close to the row mean: no cancellation)
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) t += __uint_as_float(r[j]) + s_b2[j];
            shift = t * (1.f / 32.f);
          }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = bp[q];
          const float d0 = (__uint_as_float(r[4 * q]) + b4.x) - shift, d1 = (__uint_as_float(r[4 * q + 1]) + b4.y) - shift;
          const float d2 = (__uint_as_float(r[4 * q + 2]) + b4.z) - shift, d3 = (__uint_as_float(r[4 * q + 3]) + b4.w) - shift;
          sum += (d0 + d1) + (d2 + d3);
          sq = fmaf(d0, d0, sq); sq = fmaf(d1, d1, sq); sq = fmaf(d2, d2, sq); sq = fmaf(d3, d3, sq);
        }
      }
      const float md = sum * (1.f / (float)kFfnH);
      const float var = fmaxf(sq * (1.f / (float)kFfnH) - md * md, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      const float nmr = -(shift + md) * rstd;
#pragma unroll 1
      for (int blk = 0; blk < kFfnH / 64; ++blk) {
        uint32_t pk[32], pl[32];
        uint32_t r0[32], r1[32];
        tmem_ld32(t_row + blk * 64, r0);
        tmem_ld32(t_row + blk * 64 + 32, r1);
        tmem_ld_wait();
        if (blk == kFfnH / 64 - 1) {                                      // Y fully read: the next tile may overwrite it
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(map_to_rank(y_free, 0));
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float4* bp = reinterpret_cast<const float4*>(s_b2 + blk * 64 + c * 32);
          const float4* gp = reinterpret_cast<const float4*>(s_g + blk * 64 + c * 32);
          const float4* ep = reinterpret_cast<const float4*>(s_be + blk * 64 + c * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b4 = bp[q], g4 = gp[q], e4 = ep[q];
            const uint32_t* r = c ? r1 : r0;
            const float y0 = fmaf(fmaf(__uint_as_float(r[4 * q]) + b4.x, rstd, nmr), g4.x, e4.x);
            const float y1 = fmaf(fmaf(__uint_as_float(r[4 * q + 1]) + b4.y, rstd, nmr), g4.y, e4.y);
            const float y2 = fmaf(fmaf(__uint_as_float(r[4 * q + 2]) + b4.z, rstd, nmr), g4.z, e4.z);
            const float y3 = fmaf(fmaf(__uint_as_float(r[4 * q + 3]) + b4.w, rstd, nmr), g4.w, e4.w);
            if (x3) {
              split_pack<BF16>(y0, y1, pk[c * 16 + 2 * q], pl[c * 16 + 2 * q]);
              split_pack<BF16>(y2, y3, pk[c * 16 + 2 * q + 1], pl[c * 16 + 2 * q + 1]);
            } else {
              pk[c * 16 + 2 * q] = Op16<BF16>::pack(y0, y1);
              pk[c * 16 + 2 * q + 1] = Op16<BF16>::pack(y2, y3);
            }
          }
        }
        for (int part = 0; part < parts; ++part) {                        // hi, then lo through the warp's one staging block
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
          uint8_t* dst = my_stage + lane * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(dst + ((q ^ (lane & 7)) << 4)) =
                part ? make_uint4(pl[4 * q], pl[4 * q + 1], pl[4 * q + 2], pl[4 * q + 3]) : make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_o, my_stage, part * p.lo_off + blk * 64, row0);
            tma_store_commit();
          }
        }
      }
      // the staging block is shared with the same warp of the other warpgroup (next tile): its reads must be over before that
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
    }
    if (lane == 0) tma_store_wait_all();
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

}  // namespace tc
}  // namespace hft
