// Front end of Model_SPEC2MIDI on the tensor cores (reference hftt_code/model/model_spec2midi.py:65-95: unfold(65) ->
// Conv2d(1, C, (1, 5)) -> Linear(C * 61, H) -> * sqrt(H) + pos_embedding_freq, collapsed to one 65-tap filter per hidden
// unit).  For one (segment b, bin) the 128 output frames are a Toeplitz product
//     out[f, h] = sum_j s[f + j] * Wc[h, j],      s = spec[b, bin, 0 .. 191]
// so the tile is D[128 frames x 256 hidden] = A[128 x 64] * W[256 x 64]^T with A[f, j] = s[f + j] built in shared memory
// straight in the UMMA K-major / 128-byte-swizzle layout (TMA cannot express the overlapping rows: a 2-byte row stride);
// tap 64 is one fp32 FMA per output in the epilogue.  Both operands are split hi + lo (three products), whatever the
// precision mode, so the front stays in the fp32 class like the CUDA-core kernel it replaces (fwd_tc.cu front16_kernel).
//
// Persistent, one CTA per SM, 13 warps:
//   warp 0      MMA issuer (12 UMMAs 128 x 256 x 16 per tile, accumulators double-buffered in TMEM)
//   warps 1-4   builders: spec row -> fp32 row + eight 16-bit copies shifted by 0..7 elements (so that every 8-tap chunk of
//               every Toeplitz row is ONE aligned 16-byte read) -> A hi / lo tiles (16 LDS.128 + 16 STS.128 per thread)
//   warps 5-12  two epilogue warpgroups (tile parity): TMEM -> (+ tap 64, bias) * sqrt(H) + pos -> 16-bit hi | lo -> TMA store
//               through a 3-D map of X viewed as [b * F + f][bin][column] (the tile's rows are 256 tensor rows apart)
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace hft {
namespace tc {

constexpr int kFrontThreads = 13 * 32;
constexpr int kFrontF = 128, kFrontH = 256, kFrontTaps = 65, kFrontRow = kFrontF + kFrontTaps - 1;   // 192 input frames per segment
constexpr int kFrontCopy = 528;                    // bytes between the shifted copies (512 + 16: conflict-free 16-byte reads)
constexpr int kFrontW = kFrontH * 64 * 2;          // one W part, 32 KB
constexpr int kFrontA = kFrontF * 64 * 2;          // one A part, 16 KB

struct FrontParams {
  const float* spec; long long sb, sbin, st;       // spec[b, bin, t] strides in floats
  const float* Wc;                                 // [H, 65] collapsed filter
  const float* bc;                                 // [H]
  const float* pos;                                // [n_bin, H]
  float scale;
  int n_bin, n_tiles;                              // tiles = B * n_bin
  int x3;                                          // output carries hi | lo blocks
  int lo_off;                                      // column offset of the lo block (= H)
};

struct FrontSmem {
  uint8_t w[2][kFrontW];                           // Wh | Wl, rows = hidden unit, K = taps 0..63
  uint8_t a[2][2][kFrontA];                        // [buffer][hi | lo]
  uint8_t stage[8][4096];                          // per epilogue warp: 32 rows x 64 columns
  uint8_t copies[2][8 * kFrontCopy];               // hi | lo shifted copies of the current row
  float row[2][256];                               // fp32 spec row per buffer (tap 64)
  float add[2][kFrontH];                           // bc * scale + pos[bin] per buffer
  float w64[kFrontH];                              // Wc[:, 64] * scale
  uint64_t a_full[2], a_empty[2], t_full[2], t_empty[2];
  uint32_t tmem_slot;
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

template <bool BF16>
__global__ void __launch_bounds__(kFrontThreads, 1) front_tc_kernel(const __grid_constant__ CUtensorMap map_o, const __grid_constant__ FrontParams p) {
  extern __shared__ uint8_t smem_raw[];
  FrontSmem& sm = *reinterpret_cast<FrontSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_o);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.a_full[i], 128);               // every builder thread
      mbar_init(&sm.a_empty[i], 1 + 4);            // MMA commit + the four warps of the epilogue warpgroup
      mbar_init(&sm.t_full[i], 1);
      mbar_init(&sm.t_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_slot, 512);
  // W hi / lo tiles, K-major SW128: element (h, j) at h * 128 + ((j / 8) ^ (h % 8)) * 16 + (j % 8) * 2
  for (int i = threadIdx.x; i < kFrontH * 64; i += kFrontThreads) {
    const int h = i >> 6, j = i & 63;
    const float v = p.Wc[h * kFrontTaps + j];
    const uint32_t hi = Op16<BF16>::pack(v, 0.f);
    const uint32_t lo = Op16<BF16>::pack(v - Op16<BF16>::lo(hi), 0.f);
    const uint32_t off = h * 128 + (((j >> 3) ^ (h & 7)) << 4) + (j & 7) * 2;
    *reinterpret_cast<uint16_t*>(sm.w[0] + off) = (uint16_t)(hi & 0xffffu);
    *reinterpret_cast<uint16_t*>(sm.w[1] + off) = (uint16_t)(lo & 0xffffu);
  }
  for (int i = threadIdx.x; i < kFrontH; i += kFrontThreads) sm.w64[i] = p.Wc[i * kFrontTaps + 64] * p.scale;
  for (int i = threadIdx.x; i < 2 * 8 * kFrontCopy / 4; i += kFrontThreads) reinterpret_cast<uint32_t*>(sm.copies)[i] = 0u;
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = sm.tmem_slot;

  if (warp == 0) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(kFrontF, kFrontH, BF16, false, false);
      const uint32_t wh = smem_u32(sm.w[0]), wl = smem_u32(sm.w[1]);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int ab = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(&sm.t_empty[ab], ph ^ 1);
        mbar_wait(&sm.a_full[ab], ph);
        fence_after_sync();
        const uint32_t d = tmem_base + ab * 256;
        const uint32_t ah = smem_u32(sm.a[ab][0]), al = smem_u32(sm.a[ab][1]);
        uint32_t acc = 0;
        const SDescBase kd = sdesc_base(16, 1024, kSwz128);
        auto mma4 = [&](uint32_t a_addr, uint32_t b_addr) {
          const uint32_t a0 = sdesc_lo(kd, a_addr), b0 = sdesc_lo(kd, b_addr);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16_lohi(d, a0 + 2 * k, kd.hi, b0 + 2 * k, kd.hi, idesc, acc);
            acc = 1;
          }
        };
        mma4(ah, wh);
        mma4(al, wh);
        mma4(ah, wl);
        umma_commit(&sm.a_empty[ab]);
        umma_commit(&sm.t_full[ab]);
      }
    }
  } else if (warp <= 4) {
    // ===================== builders =====================
    const int t = threadIdx.x - 32;                                   // 0..127 = Toeplitz row (frame) of this thread
    auto ld_spec = [&](int tile, int i) -> float {
      const int b = tile / p.n_bin, bin = tile % p.n_bin;
      return i < kFrontRow ? p.spec[b * p.sb + bin * p.sbin + i * p.st] : 0.f;
    };
    int it = 0;
    int tile = blockIdx.x;
    float s0 = 0.f, s1 = 0.f, a0 = 0.f, a1 = 0.f;                     // prefetched: spec[t], spec[t + 128], pos[2 t], pos[2 t + 1]
    auto prefetch = [&](int tl) {
      if (tl < p.n_tiles) {
        s0 = ld_spec(tl, t); s1 = ld_spec(tl, t + 128);
        const float* pp = p.pos + (long long)(tl % p.n_bin) * kFrontH;
        a0 = pp[2 * t]; a1 = pp[2 * t + 1];
      }
    };
    prefetch(tile);
    const float b0 = p.bc[2 * t] * p.scale, b1 = p.bc[2 * t + 1] * p.scale;
    for (; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const float c0 = s0, c1 = s1, e0 = a0, e1 = a1;
      prefetch(tile + (int)gridDim.x);                                // next tile's loads fly during this build
      mbar_wait(&sm.a_empty[ab], ph ^ 1);
      sm.row[ab][t] = c0;
      sm.row[ab][t + 128] = c1;
      sm.add[ab][2 * t] = b0 + e0;
      sm.add[ab][2 * t + 1] = b1 + e1;
      // shifted copies: copy c holds s16[i + c] at element i
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i = t + half * 128;
        if (i < kFrontRow) {
          const float v = half ? c1 : c0;
          const uint32_t hi = Op16<BF16>::pack(v, 0.f);
          const uint32_t lo = Op16<BF16>::pack(v - Op16<BF16>::lo(hi), 0.f);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (i >= c) {
              *reinterpret_cast<uint16_t*>(sm.copies[0] + c * kFrontCopy + (i - c) * 2) = (uint16_t)(hi & 0xffffu);
              *reinterpret_cast<uint16_t*>(sm.copies[1] + c * kFrontCopy + (i - c) * 2) = (uint16_t)(lo & 0xffffu);
            }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      {
        const int c = t & 7, base = t - c;                             // row t, chunk q = taps 8q .. 8q+7 = copy c, elements base + 8q ..
        uint8_t* dh = sm.a[ab][0] + t * 128;
        uint8_t* dl = sm.a[ab][1] + t * 128;
        const uint8_t* srch = sm.copies[0] + c * kFrontCopy + base * 2;
        const uint8_t* srcl = sm.copies[1] + c * kFrontCopy + base * 2;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t o = (uint32_t)((q ^ c) << 4);
          *reinterpret_cast<uint4*>(dh + o) = *reinterpret_cast<const uint4*>(srch + q * 16);
          *reinterpret_cast<uint4*>(dl + o) = *reinterpret_cast<const uint4*>(srcl + q * 16);
        }
      }
      fence_proxy_async();
      mbar_arrive(&sm.a_full[ab]);
      asm volatile("bar.sync 1, 128;" ::: "memory");                   // copies are rewritten by the next tile
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 5;
    const int wg = ew >> 2;
    const int quarter = warp & 3;
    const int f = quarter * 32 + lane;
    uint8_t* my_stage = sm.stage[ew];
    const int parts = p.x3 ? 2 : 1;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != wg) continue;
      const uint32_t ph = (it >> 1) & 1;
      const int b = tile / p.n_bin, bin = tile % p.n_bin;
      mbar_wait(&sm.a_full[wg], ph);                                   // row / add of this buffer are written
      mbar_wait(&sm.t_full[wg], ph);
      fence_after_sync();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + wg * 256;
      const float s64 = sm.row[wg][f + 64];
#pragma unroll 1
      for (int blk = 0; blk < kFrontH / 64; ++blk) {
        uint32_t pk[32], pl[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(t_row + blk * 64 + c * 32, r);
          tmem_ld_wait();
          if (blk == kFrontH / 64 - 1 && c == 1) {                     // accumulator fully read: hand the TMEM buffer back
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.t_empty[wg]);
          }
          const float4* wp = reinterpret_cast<const float4*>(sm.w64 + blk * 64 + c * 32);
          const float4* ap = reinterpret_cast<const float4*>(sm.add[wg] + blk * 64 + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w4 = wp[j], a4 = ap[j];
            const float v0 = fmaf(__uint_as_float(r[4 * j]), p.scale, fmaf(s64, w4.x, a4.x));
            const float v1 = fmaf(__uint_as_float(r[4 * j + 1]), p.scale, fmaf(s64, w4.y, a4.y));
            const float v2 = fmaf(__uint_as_float(r[4 * j + 2]), p.scale, fmaf(s64, w4.z, a4.z));
            const float v3 = fmaf(__uint_as_float(r[4 * j + 3]), p.scale, fmaf(s64, w4.w, a4.w));
            if (p.x3) {
              split_pack<BF16>(v0, v1, pk[c * 16 + 2 * j], pl[c * 16 + 2 * j]);
              split_pack<BF16>(v2, v3, pk[c * 16 + 2 * j + 1], pl[c * 16 + 2 * j + 1]);
            } else {
              pk[c * 16 + 2 * j] = Op16<BF16>::pack(v0, v1);
              pk[c * 16 + 2 * j + 1] = Op16<BF16>::pack(v2, v3);
            }
          }
        }
        if (blk == kFrontH / 64 - 1) {                                 // row / add of this buffer fully read
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.a_empty[wg]);
        }
        uint8_t* dst = my_stage + lane * 128;
        for (int part = 0; part < parts; ++part) {
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(dst + ((q ^ (lane & 7)) << 4)) =
                part ? make_uint4(pl[4 * q], pl[4 * q + 1], pl[4 * q + 2], pl[4 * q + 3]) : make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&map_o, my_stage, blk * 64 + (part ? p.lo_off : 0), bin, b * kFrontF + quarter * 32);
            tma_store_commit();
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace hft
