// tcgen05 GEMM for the TRAINING step (reduced model, BASELINE configs[4]): the Linear layers of the forward
// (y = x W^T + b, optional ReLU; reference nn.Linear inside model_spec2midi.py:322-378) and their input gradients
// (dx = dy W, optional ReLU mask / accumulate; loss.backward(), training/train.py:158) on fp32 tensors.
// Replaces sgemm_tn_kernel / sgemm_nn_kernel (fp32 CUDA cores, 45 % of the fp32 FMA peak) for K = 64 / 128 and N <= 256:
// every shape of the model is [rows in the 10^5] x [K <= 128] x [N <= 192], i.e. bound by the HBM traffic of the
// activations once the products run on the tensor cores.
//
// Arithmetic: a.w = a_hi.w_hi + a_lo.w_hi + a_hi.w_lo in fp16 with fp32 accumulation in TMEM (22 mantissa bits); every
// row of A and every row of W (output feature) is scaled by the power of two that brings its largest magnitude into [1, 2)
// before the split (gradients are ~1e-6: fp16 would flush them); the epilogue scales back per row and column (exact).
//
// Structure: persistent CTAs (two per SM, 256 threads, 256 TMEM columns each).  W is staged once per CTA as hi | lo fp16 in
// the K-major 128-byte-swizzled UMMA layout (transposed on the way when it is stored [K, N]); per 128-row tile the threads
// load the fp32 rows (the next tile's loads are issued before the current tile's epilogue and stay in registers), split and
// store them in the same layout, one thread issues the 3 x K/16 MMAs, and the epilogue (thread = row) adds the bias,
// applies ReLU, stages 32 x 32 blocks per warp in swizzled shared memory and writes them back with full 128-byte rows
// (ReLU mask / accumulate applied there with coalesced loads).
#pragma once
#include "tc_common.cuh"

namespace hft {
namespace tc {

struct TGemmArgs {
  const float* A; int lda;
  const float* W; int ldw;
  int w_kn;                    // 0: W is [N, K] (C = A W^T);  1: W is [K, N] (C = A W)
  const float* bias;           // [N] or NULL
  float* C; int ldc;
  long long M; int N, K;
  int relu, accum;             // C = relu(.) ; C += .
  const float* mask; int ldm; float mask_scale;   // C = mask > 0 ? . * mask_scale : 0   (before the accumulate)
  int tmem_cols;               // power of two >= N (32..256)
};

constexpr int kGThreads = 256;

__device__ __forceinline__ float pow2_inv_of(float mx) {       // power of two s with mx * s in [1, 2); 1 when mx is 0 / denormal / not finite
  const uint32_t e = (__float_as_uint(mx) >> 23) & 0xffu;
  return (e == 0u || e >= 253u) ? 1.f : __uint_as_float((254u - e) << 23);
}

// KB = K / 64; MINB = CTAs per SM the register budget is cut for (3 for the narrow shapes, whose tiles are too small to hide the
// per-tile load -> convert -> MMA -> epilogue chain with two CTAs)
template <int KB, int MINB>
__global__ void __launch_bounds__(kGThreads, MINB) tgemm_kernel(const TGemmArgs p) {
  constexpr int CH = KB * 8;                 // 8-element chunks per row
  constexpr int TASKS = 128 * CH / kGThreads;   // chunks per thread per tile
  constexpr int A_PART = KB * 128 * 128;     // bytes of one part (hi or lo) of the A tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_a = smem;                                        // [hi | lo][KB][128 rows x 128 B]; the epilogue staging overlays it
  uint8_t* s_w = s_a + 2 * A_PART;                            // [hi | lo][KB][N rows x 128 B]
  const int w_part = KB * p.N * 128;
  float* s_bias = reinterpret_cast<float*>(s_w + 2 * w_part); // [256]
  float* s_winv = s_bias + 256;                               // [256] 1 / scale of W row n
  float* s_inv = s_winv + 256;                                // [128] 1 / row scale of the current tile
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_inv + 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N;

  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);

  // ---- W rows (output features) -> per-row power-of-two scale -> hi | lo tiles; every thread loads all its chunks before converting ----
  {
    constexpr int WT = 8;                                      // chunks per thread: N * CH / 256 <= 8 (N <= 256 at K = 64, N <= 128 at K = 128)
    float v[WT][8];
#pragma unroll
    for (int t = 0; t < WT; ++t) {
      const int i = t * kGThreads + tid, n = i / CH, ch = i % CH;
      if (i < N * CH) {
        if (p.w_kn) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[t][j] = __ldg(p.W + (long long)(ch * 8 + j) * p.ldw + n);
        } else {
          const float4* s4 = reinterpret_cast<const float4*>(p.W + (long long)n * p.ldw + ch * 8);
          const float4 a = __ldg(s4), b = __ldg(s4 + 1);
          v[t][0] = a.x; v[t][1] = a.y; v[t][2] = a.z; v[t][3] = a.w; v[t][4] = b.x; v[t][5] = b.y; v[t][6] = b.z; v[t][7] = b.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[t][j] = 0.f;
      }
    }
#pragma unroll
    for (int t = 0; t < WT; ++t) {
      const int i = t * kGThreads + tid, n = i / CH, ch = i % CH;
      float mx = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) mx = fmaxf(mx, fabsf(v[t][j]));
#pragma unroll
      for (int o = CH / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));   // the CH lanes that share the row
      if (i < N * CH) {
        const float sg = pow2_inv_of(mx);
        if (ch == 0) s_winv[n] = 1.f / sg;
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_pack<false>(v[t][2 * j] * sg, v[t][2 * j + 1] * sg, h[j], l[j]);
        const int off = (ch >> 3) * N * 128 + n * 128 + (((ch & 7) ^ (n & 7)) << 4);
        *reinterpret_cast<uint4*>(s_w + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(s_w + w_part + off) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
    for (int i = tid; i < 256; i += kGThreads) s_bias[i] = (p.bias && i < N) ? __ldg(p.bias + i) : 0.f;
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const long long n_tiles = (p.M + 127) >> 7;
  float4 ra[TASKS][2];
  auto load_tile = [&](long long tile) {
#pragma unroll
    for (int t = 0; t < TASKS; ++t) {
      const int idx = t * kGThreads + tid, r = idx / CH, ch = idx % CH;
      const long long row = tile * 128 + r;
      if (row < p.M) {
        const float4* s4 = reinterpret_cast<const float4*>(p.A + row * p.lda + ch * 8);
        ra[t][0] = __ldg(s4);
        ra[t][1] = __ldg(s4 + 1);
      } else {
        ra[t][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        ra[t][1] = ra[t][0];
      }
    }
  };
  const uint32_t idesc = make_idesc(128, N, false, false, false);
  const SDescBase kd = sdesc_base(16, 1024, kSwz128);
  const uint32_t a16 = sdesc_lo(kd, smem_u32(s_a)), w16 = sdesc_lo(kd, smem_u32(s_w));   // low descriptor words of the hi pieces, K block 0
  const int q = warp & 3, half = warp >> 2;
  const int n_chunks = (N + 31) >> 5;
  uint8_t* stage = s_a + warp * 4096;
  uint32_t phase = 0;

  long long tile = blockIdx.x;
  if (tile < n_tiles) load_tile(tile);
  for (; tile < n_tiles; tile += gridDim.x) {
    // ---- fp32 rows -> per-row power-of-two scale -> hi | lo tiles ----
#pragma unroll
    for (int t = 0; t < TASKS; ++t) {
      const int idx = t * kGThreads + tid, r = idx / CH, ch = idx % CH;
      const float4 a = ra[t][0], b = ra[t][1];
      float mx = fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))), fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
#pragma unroll
      for (int o = CH / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));   // the CH lanes that share the row
      const float sg = pow2_inv_of(mx);
      if (ch == 0) s_inv[r] = 1.f / sg;
      uint32_t h[4], l[4];
      split_pack<false>(a.x * sg, a.y * sg, h[0], l[0]);
      split_pack<false>(a.z * sg, a.w * sg, h[1], l[1]);
      split_pack<false>(b.x * sg, b.y * sg, h[2], l[2]);
      split_pack<false>(b.z * sg, b.w * sg, h[3], l[3]);
      const int off = (ch >> 3) * (128 * 128) + r * 128 + (((ch & 7) ^ (r & 7)) << 4);
      *reinterpret_cast<uint4*>(s_a + off) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(s_a + A_PART + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 0) {                                           // whole warp (uniform descriptors), the MMAs issued back to back by one elected lane
      if (elect_one()) {
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          const uint32_t ap = a16 + (part == 1 ? (A_PART >> 4) : 0), wp = w16 + (part == 2 ? ((uint32_t)w_part >> 4) : 0);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint32_t wk = wp + ((uint32_t)(kb * N * 128) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_lohi(tmem_base, ap + ((kb * (128 * 128)) >> 4) + 2 * k, kd.hi, wk + 2 * k, kd.hi, idesc, (part | kb | k) ? 1u : 0u);
          }
        }
        umma_commit(bar);
      }
      __syncwarp();
    }
    const long long next = tile + gridDim.x;
    if (next < n_tiles) load_tile(next);                       // in flight during the MMAs and the epilogue
    mbar_wait(bar, phase);
    phase ^= 1;
    fence_after_sync();

    // ---- epilogue: thread = row; 32-column blocks through the warp's swizzled staging block ----
    const float unscale = s_inv[q * 32 + lane];
    for (int c = half; c < n_chunks; c += 2) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          o[e] = fmaf(__uint_as_float(v[4 * g + e]) * unscale, s_winv[c * 32 + 4 * g + e], s_bias[c * 32 + 4 * g + e]);
          if (p.relu) o[e] = fmaxf(o[e], 0.f);
        }
        *reinterpret_cast<float4*>(stage + lane * 128 + ((g ^ (lane & 7)) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + (lane >> 3), g = lane & 7;
        float4 x = *reinterpret_cast<const float4*>(stage + rr * 128 + ((g ^ (rr & 7)) << 4));
        const long long grow = tile * 128 + q * 32 + rr;
        const int col = c * 32 + g * 4;
        if (grow < p.M && col < N) {
          float* cp = p.C + grow * p.ldc + col;
          if (p.mask) {
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(p.mask + grow * p.ldm + col));
            x.x = m4.x > 0.f ? x.x * p.mask_scale : 0.f; x.y = m4.y > 0.f ? x.y * p.mask_scale : 0.f;
            x.z = m4.z > 0.f ? x.z * p.mask_scale : 0.f; x.w = m4.w > 0.f ? x.w * p.mask_scale : 0.f;
          }
          if (p.accum) {
            const float4 c4 = *reinterpret_cast<const float4*>(cp);
            x.x += c4.x; x.y += c4.y; x.z += c4.z; x.w += c4.w;
          }
          *reinterpret_cast<float4*>(cp) = x;
        }
      }
      __syncwarp();
    }
    fence_before_sync();
    __syncthreads();                                           // staging (overlays the A tile) and the accumulator are free again
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

inline size_t tgemm_smem(int KB, int N) { return 1024 + 2 * (size_t)KB * 128 * 128 + 2 * (size_t)KB * N * 128 + (256 + 256 + 128) * sizeof(float) + 64; }

}  // namespace tc
}  // namespace hft
