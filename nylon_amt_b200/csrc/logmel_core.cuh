// Per-lane building blocks of the fused log-mel kernel (frames -> Hann -> 2048-point real FFT ->
// |X|^2 -> banded mel sums -> log).  Replaces torchaudio MelSpectrogram + log in AMT.wav2feature
// (reference hftt_code/model/amt.py:59-61).
//
// One warp owns one frame.  The 2048-point real FFT is a 1024-point complex FFT of z[n] = x[2n] + i x[2n+1]
// done as 32 x 32 (four-step): every lane runs a 32-point FFT in registers, multiplies by W_1024^{n1 k2},
// the warp transposes through shared memory, every lane runs a second 32-point FFT, and the real-input
// spectrum is untangled pairwise (k, 1024-k).
//
// Every function takes the lane index explicitly and is __host__ __device__, so that the index arithmetic is
// verified on the CPU by running the 32 lanes one after another (tools/logmel_hostsim.cpp).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define HFT_HD __host__ __device__ __forceinline__
#else
#define HFT_HD inline
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif
#include "fft32_gen.cuh"

namespace hft {

constexpr int kNfft = 2048;
constexpr int kHop = 256;
constexpr int kNmels = 256;
constexpr int kNfreq = 1025;
constexpr int kTStride = 33;          // float2 row stride of the 32x32 transpose tile (conflict-free)
constexpr int kMelGroups = kNmels / 32;   // lane handles mel bins lane + 32 j, j = 0..7
constexpr int kMaxMelTaps = 80;           // sum over the groups of the longest band in the group (76 for the reference's filterbank)
constexpr int kMaxMelW = kMaxMelTaps * 32;
constexpr int kMelInfo = kNmels + 2 * kMelGroups;

// Band weights in a lane-major layout: the 32 bins of group j run the SAME number of taps n_j (= the longest band of the group; shorter
// bands are zero padded), tap i of bin lane + 32 j is melw[(base_j + i) * 32 + lane] -- a conflict-free shared-memory read for every (j, i)
// (the packed [bin][tap] layout read at offsets that differ by the band length from lane to lane: 2.9 wavefronts per load on average).
// The 0.25 of the power spectrum (lm_power_regs) is folded into the weights: exact, a power of two.
// melinfo[m] = first FFT bin of band m; melinfo[kNmels + j] = base_j; melinfo[kNmels + kMelGroups + j] = n_j.
// Returns 0, or 1 when a band is longer than 31 bins / the table does not fit.  fb = [kNfreq][kNmels].
inline int lm_pack_filterbank(const float* fb, float* melw, uint32_t* melinfo) {
  int lo_[kNmels], len_[kNmels];
  for (int m = 0; m < kNmels; ++m) {
    int lo = -1, hi = -1;
    for (int k = 0; k < kNfreq; ++k)
      if (fb[(long)k * kNmels + m] != 0.f) { if (lo < 0) lo = k; hi = k; }
    lo_[m] = lo < 0 ? 0 : lo;
    len_[m] = lo < 0 ? 0 : hi - lo + 1;
    if (len_[m] > 31) return 1;
  }
  for (int i = 0; i < kMaxMelW; ++i) melw[i] = 0.f;
  int base = 0;
  for (int j = 0; j < kMelGroups; ++j) {
    int n = 0;
    for (int l = 0; l < 32; ++l) n = len_[l + 32 * j] > n ? len_[l + 32 * j] : n;
    if (base + n > kMaxMelTaps) return 1;
    for (int l = 0; l < 32; ++l) {
      const int m = l + 32 * j;
      // bins 0 and 1024 of the power spectrum are stored unscaled (lm_power_regs): their weights keep the factor 1
      for (int i = 0; i < len_[m]; ++i) {
        const int k = lo_[m] + i;
        melw[(base + i) * 32 + l] = fb[(long)k * kNmels + m] * ((k == 0 || k == kNfreq - 1) ? 1.0f : 0.25f);
      }
      melinfo[m] = (uint32_t)lo_[m];
    }
    melinfo[kNmels + j] = (uint32_t)base;
    melinfo[kNmels + kMelGroups + j] = (uint32_t)n;
    base += n;
  }
  return 0;
}

// Step 1: lane n1 loads its 32 windowed points.  xs = this frame's 2048 samples (8-byte aligned), win = Hann[2048].
HFT_HD void lm_rows_load(int lane, const float* xs, const float* win, float2 (&v)[32]) {
  const float2* x2 = reinterpret_cast<const float2*>(xs);
  const float2* w2 = reinterpret_cast<const float2*>(win);
#pragma unroll
  for (int n2 = 0; n2 < 32; ++n2) {
    float2 a = x2[lane + 32 * n2], w = w2[lane + 32 * n2];
    v[n2] = make_float2(a.x * w.x, a.y * w.y);
  }
}
// Step 3: twiddle by tw2[k2*32 + n1] = W_1024^{n1 k2} and write T[k2*kTStride + n1].
HFT_HD void lm_rows_store(int lane, const float2 (&v)[32], const float2* tw2, float2* T) {
#pragma unroll
  for (int k2 = 0; k2 < 32; ++k2) {
    float2 y = v[HFT_BITREV5(k2)], w = tw2[k2 * 32 + lane];
    T[k2 * kTStride + lane] = make_float2(y.x * w.x - y.y * w.y, y.x * w.y + y.y * w.x);
  }
}
// Step 1+2+3 in one call.
HFT_HD void lm_rows(int lane, const float* xs, const float* win, const float2* tw2, float2* T) {
  float2 v[32];
  lm_rows_load(lane, xs, win, v);
  hft_fft32(v);
  lm_rows_store(lane, v, tw2, T);
}

// Step 4a: lane k2 reads its row of the transposed tile and transforms it (result stays in registers).
HFT_HD void lm_cols_load_raw(int lane, const float2* T, float2 (&u)[32]) {
#pragma unroll
  for (int n1 = 0; n1 < 32; ++n1) u[n1] = T[lane * kTStride + n1];
}
HFT_HD void lm_cols_load(int lane, const float2* T, float2 (&u)[32]) {
  lm_cols_load_raw(lane, T, u);
  hft_fft32(u);
}

// Step 4b: Z[k2 + 32 k1] in natural order.
HFT_HD void lm_cols_store(int lane, const float2 (&u)[32], float2* Z) {
#pragma unroll
  for (int k1 = 0; k1 < 32; ++k1) Z[lane + 32 * k1] = u[HFT_BITREV5(k1)];
}

// Step 5: real-input untangle + power.  twr[k] = exp(-2 pi i k / 2048), k = 0..512.
// P[k] = 4 |X[k]|^2 for k = 1..1023 and |X[k]|^2 for k = 0, 1024 (the mel weights carry the 0.25).  The lane's 32 (+1) powers are computed into registers first (lm_power_regs) and
// stored afterwards (lm_power_store), so that P may overlay Z in shared memory (a __syncwarp() goes between the two).
HFT_HD void lm_power_regs(int lane, const float2* Z, const float2* twr, float (&plo)[16], float (&phi)[16], float& p0, float& p1024) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    int k = 1 + lane + 32 * j;                       // 1..512
    float2 zk = Z[k], zn = Z[1024 - k], w = twr[k];
    float ar = zk.x + zn.x, ai = zk.y - zn.y;        // A  = Z[k] + conj(Z[N-k])      (= 2 E[k])
    float br = zk.x - zn.x, bi = zk.y + zn.y;        // B  = Z[k] - conj(Z[N-k])
    float orr = bi, oi = -br;                        // O2 = B / i                    (= 2 O[k])
    float tr = orr * w.x - oi * w.y, ti = orr * w.y + oi * w.x;   // T2 = W^k O2
    float xr = ar + tr, xi = ai + ti, yr = ar - tr, yi = ai - ti;
    plo[j] = xr * xr + xi * xi;                      // 4 |X[k]|^2: the 0.25 lives in the mel weights (lm_pack_filterbank)
    phi[j] = yr * yr + yi * yi;
  }
  p0 = 0.f; p1024 = 0.f;
  if (lane == 0) {
    float2 z0 = Z[0];
    float s = z0.x + z0.y, d = z0.x - z0.y;
    p0 = s * s;
    p1024 = d * d;
  }
}
HFT_HD void lm_power_store(int lane, const float (&plo)[16], const float (&phi)[16], float p0, float p1024, float* P) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    int k = 1 + lane + 32 * j;
    P[k] = plo[j];
    P[1024 - k] = phi[j];
  }
  if (lane == 0) { P[0] = p0; P[1024] = p1024; }
}
HFT_HD void lm_power(int lane, const float2* Z, const float2* twr, float* P) {      // P and Z must not overlap
  float plo[16], phi[16], p0, p1024;
  lm_power_regs(lane, Z, twr, plo, phi, p0, p1024);
  lm_power_store(lane, plo, phi, p0, p1024, P);
}

// Step 6+7: banded mel sums and log.  Lane handles mel bins lane + 32 j; every bin of a group runs the group's tap count (uniform trip
// count, no divergence); reads past a band's end meet zero weights (P is followed by finite tile contents).
#ifndef HFT_LM_FASTLOG
#define HFT_LM_FASTLOG 1        // log through MUFU.LG2 (__logf, abs error ~1e-6 on the log) instead of the 20-instruction logf
#endif
#ifndef HFT_LM_UNROLL_J
#define HFT_LM_UNROLL_J 0       // 1: the eight bin groups unrolled; 0: a loop -- measured r02: 8.49 vs 8.86 ms per 10 h (the hot loop is ~35 KB of SASS;
                                // sharing one copy of the FFT between the two passes through a two-trip loop was slower: 8.91 ms)
#endif
HFT_HD void lm_mel(int lane, const float* P, const float* melw, const uint32_t* melinfo, float log_offset, float* out_row) {
#if HFT_LM_UNROLL_J
#pragma unroll
#else
#pragma unroll 1
#endif
  for (int j = 0; j < kMelGroups; ++j) {
    const int m = lane + 32 * j;
    const float* p = P + melinfo[m];
    const float* w = melw + melinfo[kNmels + j] * 32 + lane;
    const int n = (int)melinfo[kNmels + kMelGroups + j];
    float acc = 0.f;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
      acc = fmaf(p[i], w[i * 32], acc);
      acc = fmaf(p[i + 1], w[(i + 1) * 32], acc);
      acc = fmaf(p[i + 2], w[(i + 2) * 32], acc);
      acc = fmaf(p[i + 3], w[(i + 3) * 32], acc);
    }
    for (; i < n; ++i) acc = fmaf(p[i], w[i * 32], acc);
#ifdef __CUDA_ARCH__
#if HFT_LM_FASTLOG
    out_row[m] = __logf(acc + log_offset);
#else
    out_row[m] = logf(acc + log_offset);
#endif
#else
    out_row[m] = ::logf(acc + log_offset);
#endif
  }
}

}  // namespace hft
