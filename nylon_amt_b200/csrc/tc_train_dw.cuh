// tcgen05 weight-gradient GEMM of the TRAINING step (reduced model, BASELINE configs[4]):
//   dW[N, K] += dY[M, N]^T X[M, K],   db[N] += colsum(dY)          (nn.Linear under loss.backward(), training/train.py:158)
// with M in the 10^5 (every row of the batch), N <= 192 output features, K = 64 / 128 input features.  Replaces dw_gemm_kernel
// (fp32 CUDA cores: 32 TFLOP/s = 45 % of the fp32 FMA peak at 1.5 TB/s of operands, i.e. bound by the FMA issue) -- on the tensor
// cores the shape is bound by the HBM traffic of dY and X.
//
// The reduction runs over the ROWS, i.e. over the slow dimension of both row-major operands, so both are fed to the MMA
// MN-major straight from row-major tiles ([32 rows][128 bytes] per 64-column block, 128-byte swizzle): D[n, k] with
// M_mma = 128 output features (two neighbouring 64-column blocks; when N runs out the neighbour is the first block of X and
// those accumulator rows are simply not stored), N_mma = K, K_mma = 16 rows per instruction.  db comes from the same MMAs:
// one extra 16-column B block whose column 0 is 1.
// Arithmetic: bf16 with THREE pieces per value (x = b0 + b1 + b2, 8 mantissa bits each; fp32 exponent range, so the ~1e-6
// gradients need no scaling) and the six products that matter (b0c0, b0c1, b1c0, b1c1, b0c2, b2c0), fp32 accumulation in
// TMEM: fp32-class results.
//
// Structure: one persistent CTA per SM splits the rows; a loader warp, 16 converter warps and an MMA warp, decoupled by two rings with
// full / empty mbarriers (no CTA-wide barrier in the loop):
//   loader:     two TMA boxes per 32-row stage ([32, N] of dY, [32, K] of X, fp32, no swizzle) land in the raw ring -- up to
//               ~120 KB in flight per SM, which is what a read stream needs to approach the HBM rate (register prefetch, two CTAs per SM
//               or more warps all stalled at 1.8-2.6 TB/s: the bytes in flight were the limit, not the arithmetic);
//   converters: wait for a raw stage, read their 8-float chunks, release the stage, split the chunks into the three bf16 pieces and write
//               them to an operand slot once the MMAs that read the slot's previous content have completed (tcgen05.commit -> empty barrier);
//   MMA thread: waits for a full slot, issues its 18 G MMAs (descriptors: one 32-bit add each) and commits.
// The accumulators stay in TMEM for the whole slice and go to dW / db with one set of fp32 atomics per CTA.
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
  long long rows_per_cta;      // multiple of 32
  int n_slots;                 // operand ring depth (2)
  int n_raw;                   // raw fp32 ring depth (2..8)
};

constexpr int kDwProducers = 512;                    // producer threads (warps 0..15: the conversion is bound by ALU latency, it wants warps); warp 16 issues the MMAs
constexpr int kDwThreads = kDwProducers + 64;           // + MMA warp + loader warp
constexpr int kDwRows = 32;                          // rows per stage
constexpr int kDwBlk = kDwRows * 128;                // bytes of one 64-column block of a stage (one piece)

// x = b0 + b1 + b2 for a pair of values, packed as bf16x2 words
__device__ __forceinline__ void split3_pack(float a, float b, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  p0 = Op16<true>::pack(a, b);
  const float ra = a - Op16<true>::lo(p0), rb = b - Op16<true>::hi(p0);
  p1 = Op16<true>::pack(ra, rb);
  p2 = Op16<true>::pack(ra - Op16<true>::lo(p1), rb - Op16<true>::hi(p1));
}

inline size_t tdw_slot_bytes(int KB, int N) { return (3 * (size_t)((N + 63) / 64) + 1 + 3 * (size_t)KB) * kDwBlk; }
inline size_t tdw_raw_bytes(int KB, int N) { return (size_t)kDwRows * (N + KB * 64) * 4; }
inline size_t tdw_smem(int KB, int N, int n_slots, int n_raw) { return 1024 + n_slots * tdw_slot_bytes(KB, N) + n_raw * tdw_raw_bytes(KB, N) + 512; }

// KB = K / 64; TPS = 8-float chunks per producer thread and stage = ceil(32 * (N + K) / 8 / 512) (1 or 2)
template <int KB, int TPS>
__global__ void __launch_bounds__(kDwThreads, 1) tdw_kernel(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_x, const TDwArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int N = p.N, NS = p.n_slots, NR = p.n_raw;
  const int PB = (N + 63) >> 6;                      // 64-column blocks of dY
  const int G = (PB + 1) >> 1;                       // M = 128 groups (pairs of neighbouring blocks)
  // operand slot: [piece 0..2][PB blocks of dY] | ones block (column 0 = 1) | [piece 0..2][KB blocks of X] -- the B side is ONE run of
  // 64-column atoms [ones, c0, c1, c2], so a dY piece meets every X piece it needs (and the ones column) in one or two wide MMAs
  const int y_piece = PB * kDwBlk, x_piece = KB * kDwBlk;
  const int b_off = 3 * y_piece;                     // first B atom (the ones block) inside a slot
  const int slot_bytes = b_off + kDwBlk + 3 * x_piece;
  uint8_t* s_ring = smem;
  const int raw_y = kDwRows * N * 4;                 // bytes of the dY box of a raw stage ([32 rows][N floats], dense)
  const int raw_bytes = raw_y + kDwRows * KB * 64 * 4;   // + the X box ([32 rows][K floats])
  uint8_t* s_raw = s_ring + (size_t)NS * slot_bytes; // [NR][dY box | X box]
  uint64_t* full = reinterpret_cast<uint64_t*>(s_raw + (size_t)NR * raw_bytes);   // [4] operand slot written
  uint64_t* empty = full + 4;                        // [4] operand slot read by its MMAs
  uint64_t* raw_full = empty + 4;                    // [8] raw stage landed
  uint64_t* raw_empty = raw_full + 8;                // [8] raw stage read by the converters
  uint64_t* done = raw_empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&full[i], kDwProducers / 32); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], kDwProducers / 32); }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == kDwProducers / 32) tmem_alloc(tmem_slot, 512);
  // zero the ring once (the padding columns of a ragged last dY block stay zero), then the ones block of every slot
  for (int i = tid; i < NS * slot_bytes / 16; i += kDwThreads) reinterpret_cast<uint4*>(s_ring)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int i = tid; i < NS * kDwRows; i += kDwThreads) {
    const int sl = i / kDwRows, r = i % kDwRows;     // logical chunk 0 of row r sits at position (0 ^ (r & 7)); bf16 1.0 in element 0
    *reinterpret_cast<uint32_t*>(s_ring + (size_t)sl * slot_bytes + b_off + r * 128 + ((r & 7) << 4)) = 0x00003F80u;
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const long long m_begin = (long long)blockIdx.x * p.rows_per_cta;
  const long long m_end = m_begin + p.rows_per_cta < p.M ? m_begin + p.rows_per_cta : p.M;
  const int S = m_begin < m_end ? (int)((m_end - m_begin + kDwRows - 1) / kDwRows) : 0;
  const uint32_t gcols = 64 + 3 * KB * 64;           // TMEM columns per group: [db | D0 | D1 | D2], one 64-column atom of B each (dW = D0 + D1 + D2)

  if (warp < kDwProducers / 32) {
    // ===================== producers =====================
    const int chp = N >> 3, chs = chp + KB * 8;      // valid 8-float chunks per row: dY, both
    const int n_tasks = kDwRows * chs;
    bool live[TPS];                                  // this thread has a chunk t
    int row_t[TPS], off_t[TPS], pstr_t[TPS], roff_t[TPS];   // its row, its offset inside piece 0, the piece stride, its offset inside a raw stage
#pragma unroll
    for (int t = 0; t < TPS; ++t) {
      const int idx = t * kDwProducers + tid, r = idx / chs, c = idx % chs;
      const bool isp = c < chp;
      const int cc = isp ? c : c - chp;
      live[t] = idx < n_tasks;
      row_t[t] = r;
      off_t[t] = (isp ? 0 : b_off + kDwBlk) + (cc >> 3) * kDwBlk + r * 128 + (((cc & 7) ^ (r & 7)) << 4);
      pstr_t[t] = isp ? y_piece : x_piece;
      roff_t[t] = isp ? r * N * 4 + cc * 32 : raw_y + r * KB * 256 + cc * 32;
    }
    for (int s = 0; s < S; ++s) {
      const int rs = s % NR, slot = s % NS, use = s / NS;
      mbar_wait(&raw_full[rs], (uint32_t)((s / NR) & 1));
      float4 ra[TPS][2];
      const int nv = (int)(m_end - m_begin - (long long)s * kDwRows);   // valid rows of the stage (>= 32: all)
#pragma unroll
      for (int t = 0; t < TPS; ++t) {
        if (live[t] && row_t[t] < nv) {
          const float4* s4 = reinterpret_cast<const float4*>(s_raw + (size_t)rs * raw_bytes + roff_t[t]);
          ra[t][0] = s4[0];
          ra[t][1] = s4[1];
        } else {
          ra[t][0] = make_float4(0.f, 0.f, 0.f, 0.f);
          ra[t][1] = ra[t][0];
        }
      }
      fence_proxy_async();                           // order these generic-proxy reads before the TMA (async-proxy) refill of the stage
      __syncwarp();
      if (lane == 0) mbar_arrive(&raw_empty[rs]);      // the stage is in registers: the loader may refill it
      if (use > 0) mbar_wait(&empty[slot], (uint32_t)((use - 1) & 1));     // the MMAs that read this slot's previous stage have completed
      uint8_t* base = s_ring + (size_t)slot * slot_bytes;
#pragma unroll
      for (int t = 0; t < TPS; ++t) {
        if (live[t]) {
          const float4 a = ra[t][0], b = ra[t][1];
          uint32_t w0[4], w1[4], w2[4];
          split3_pack(a.x, a.y, w0[0], w1[0], w2[0]);
          split3_pack(a.z, a.w, w0[1], w1[1], w2[1]);
          split3_pack(b.x, b.y, w0[2], w1[2], w2[2]);
          split3_pack(b.z, b.w, w0[3], w1[3], w2[3]);
          *reinterpret_cast<uint4*>(base + off_t[t]) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
          *reinterpret_cast<uint4*>(base + pstr_t[t] + off_t[t]) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
          *reinterpret_cast<uint4*>(base + 2 * pstr_t[t] + off_t[t]) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
        }
      }
      fence_proxy_async();                           // generic-proxy shared-memory writes -> visible to the UMMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[slot]);
    }
  } else if (warp == kDwProducers / 32 + 1) {
    // ===================== loader: two TMA boxes per stage (rows past the end of the tensor arrive as zeros) =====================
    if (lane == 0) {
      tma_prefetch_desc(&map_y);
      tma_prefetch_desc(&map_x);
      for (int s = 0; s < S; ++s) {
        const int rs = s % NR;
        if (s >= NR) mbar_wait_spin(&raw_empty[rs], (uint32_t)((s / NR - 1) & 1));
        const long long m0 = m_begin + (long long)s * kDwRows;
        uint8_t* dst = s_raw + (size_t)rs * raw_bytes;
        mbar_expect_tx(&raw_full[rs], (uint32_t)raw_bytes);
        tma_load_2d(dst, &map_y, 0, (int)m0, &raw_full[rs]);
        tma_load_2d(dst + raw_y, &map_x, 0, (int)m0, &raw_full[rs]);
      }
    }
  } else {
    // ===================== MMA issuer: the whole warp runs the loop (uniform descriptors), one elected lane issues and commits =====================
    // A = two neighbouring dY blocks of one piece (MN-major, M = 128); B = a run of up to four 64-column atoms of [ones | c0 | c1 | c2]
    // (MN-major, N <= 256).  dY piece q meets the ones column and the X pieces 0 .. 2 - q: the six products that matter
    // (b0c0, b0c1, b0c2, b1c0, b1c1, b2c0) in 3 (K = 64) or 5 (K = 128) wide MMAs per 16 rows instead of 9 narrow ones -- the narrow
    // ones were bound by the shared-memory read of the 4 KB A tile each of them repeats.
    const uint64_t d0 = make_sdesc(smem_u32(s_ring), kDwBlk, 1024, kSwz128);
    const uint32_t hi = (uint32_t)(d0 >> 32), ring_lo = (uint32_t)d0;
    const uint32_t slot16 = (uint32_t)slot_bytes >> 4;
    for (int s = 0; s < S; ++s) {
      const int slot = s % NS, use = s / NS;
      mbar_wait_spin(&full[slot], (uint32_t)(use & 1));
      fence_after_sync();
      const uint32_t base = ring_lo + slot * slot16;
      const uint32_t acc0 = s > 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (g < G) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const uint32_t a_lo = base + ((uint32_t)(q * y_piece + 2 * g * kDwBlk) >> 4);
            constexpr int kAtomsMax = 1 + 3 * KB;
            const int atoms = kAtomsMax - q * KB;           // [ones, c0 .. c(2-q)]
#pragma unroll
            for (int a0 = 0; a0 < kAtomsMax; a0 += 4) {
              if (a0 < atoms) {
                const int na = atoms - a0 < 4 ? atoms - a0 : 4;
                const uint32_t b_lo = base + ((uint32_t)(b_off + a0 * kDwBlk) >> 4);
                const uint32_t idesc = make_idesc(128, na * 64, true, true, true);
#pragma unroll
                for (int k = 0; k < kDwRows / 16; ++k)
                  umma_f16_lohi(tmem_base + g * gcols + a0 * 64, a_lo + k * 128, hi, b_lo + k * 128, hi, idesc, (q | k) ? 1u : acc0);
              }
            }
          }
        }
      }
      umma_commit(&empty[slot]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
  }
  __syncwarp();
  if (warp < 8 && S > 0) {
    mbar_wait(done, 0);
    fence_after_sync();
    // ---- this CTA's slice -> dW / db (fp32 atomics): thread = output feature (TMEM lane), 8 warps = 4 lane quarters x 2 column halves ----
    const int q = warp & 3, half = warp >> 2;
    for (int g = 0; g < G; ++g) {
      const int n = g * 128 + q * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + g * gcols;
      for (int c = half; c < KB * 2; c += 2) {
        uint32_t v0[32], v1[32], v2[32];
        tmem_ld32(t_row + 64 + c * 32, v0);
        tmem_ld32(t_row + 64 + KB * 64 + c * 32, v1);
        tmem_ld32(t_row + 64 + 2 * KB * 64 + c * 32, v2);
        tmem_ld_wait();
        if (n < N) {
          float* dst = p.dW + (long long)n * p.ldw + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(v0[j]) + (__uint_as_float(v1[j]) + __uint_as_float(v2[j])));
        }
      }
      if (half == 0 && p.db) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(t_row)
                     : "memory");
        tmem_ld_wait();
        if (n < N) atomicAdd(p.db + n, __uint_as_float(v[0]));
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kDwProducers / 32) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace hft
