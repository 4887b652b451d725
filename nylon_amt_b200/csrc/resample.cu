// Channel mix-down + sample-rate conversion in front of the log-mel kernel (reference hftt_code/model/amt.py:56-58:
// wave_mono = torch.mean(wave, dim=0); torchaudio.transforms.Resample(sr, 16000)(wave_mono)).
//
// torchaudio's Resample is a polyphase FIR (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99): with o = orig/gcd,
// n = new/gcd and the table K[n][2*width + o] it computes, on the signal zero-padded by (width, width + o),
//     y[i*n + p] = sum_t xpad[i*o + t] * K[p][t]          (F.conv1d with stride o, one output channel per phase p)
// truncated to ceil(n * length / o) samples.  The table is built on the host by hft_resample_build_table (this file: the published
// sinc_interp_hann formula in double precision, within 1e-7 of torchaudio's own table -- tests/test_resample.py) or handed in by
// the caller (hft_resample_create).
//
// Kernel: one CTA per 32 consecutive i.  The mono-mixed input span (32*o + kw samples) is staged once in shared memory;
// warp w walks the phases p = w, w + warps, ...; lane = i, so the coefficient K[p][t] is a broadcast load and the 32 lanes
// read x at stride o (conflict-free when o is odd; the table walk dominates otherwise).  Results go through a
// shared-memory tile so that the global stores are contiguous.
#include "common.cuh"
#include "hft_internal.h"

#include <vector>

namespace hft {

struct ResamplePlan {
  int o = 1, n = 1, width = 0, kw = 0;
  float* d_kernel = nullptr;       // [n][kw]
};

constexpr int kRsI = 32;           // i values (input strides) per CTA
constexpr int kRsWarps = 8;

__global__ void __launch_bounds__(kRsWarps * 32) resample_mono_kernel(const float* __restrict__ wav, int channels, long long n_in, const float* __restrict__ K,
                                                                    int o, int n, int width, int kw, float* __restrict__ out, long long n_out) {
  extern __shared__ __align__(16) float smem_rs[];
  const int span = kRsI * o + kw;
  float* s_x = smem_rs;                       // [span] mono-mixed, zero padded input
  float* s_y = smem_rs + span;                // [kRsI][n] outputs of this CTA, contiguous in j
  const long long i0 = (long long)blockIdx.x * kRsI;
  const long long x0 = i0 * o - width;        // first input sample of the span
  const float inv_c = 1.f / (float)channels;
  for (int t = threadIdx.x; t < span; t += blockDim.x) {
    const long long s = x0 + t;
    float v = 0.f;
    if (s >= 0 && s < n_in) {
      for (int c = 0; c < channels; ++c) v += wav[(long long)c * n_in + s];
      v *= inv_c;
    }
    s_x[t] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xs = s_x + lane * o;
  for (int p = warp; p < n; p += kRsWarps) {
    const float* kp = K + (long long)p * kw;
    float acc = 0.f;
    for (int t = 0; t < kw; ++t) acc = fmaf(xs[t], __ldg(kp + t), acc);
    s_y[lane * n + p] = acc;
  }
  __syncthreads();
  const long long j0 = i0 * n;
  for (int t = threadIdx.x; t < kRsI * n; t += blockDim.x)
    if (j0 + t < n_out) out[j0 + t] = s_y[t];
}

}  // namespace hft

using namespace hft;

#include <cmath>
#include <numeric>

// The polyphase table of torchaudio's sinc_interp_hann resampler (functional.py _get_sinc_resample_kernel with the defaults the
// reference uses: lowpass_filter_width 6, rolloff 0.99), in double precision, rounded to fp32 at the end:
//   base = min(o, n) * rolloff;  width = ceil(lpw * o / base);  idx = (-width .. width + o - 1) / o
//   t[p][k] = clamp((float(-p) / float(n) + idx[k]) * base, -lpw, lpw)    (the phase offset is rounded to fp32 first, as torchaudio does)
//   K[p][k] = sinc(pi t) * cos^2(pi t / (2 lpw)) * base / o
extern "C" int64_t hft_resample_build_table(int32_t orig_hz, int32_t new_hz, float* table_host, int64_t capacity, int32_t* orig_reduced, int32_t* new_reduced,
                                            int32_t* width_out) {
  if (orig_hz < 1 || new_hz < 1) return -1;
  const int g = std::gcd(orig_hz, new_hz);
  const int o = orig_hz / g, n = new_hz / g;
  const double lpw = 6.0, rolloff = 0.99;
  const double base = (double)(o < n ? o : n) * rolloff;
  const int width = (int)std::ceil(lpw * o / base);
  const int kw = 2 * width + o;
  if (orig_reduced) *orig_reduced = o;
  if (new_reduced) *new_reduced = n;
  if (width_out) *width_out = width;
  const int64_t need = (int64_t)n * kw;
  if (!table_host || capacity < need) return need;
  const double pi = 3.14159265358979323846;
  for (int p = 0; p < n; ++p) {
    const double phase = (double)((float)(-p) / (float)n);
    for (int k = 0; k < kw; ++k) {
      double t = (phase + (double)(k - width) / (double)o) * base;
      t = t < -lpw ? -lpw : (t > lpw ? lpw : t);
      const double c = std::cos(t * pi / lpw / 2.0);
      const double a = t * pi;
      const double sinc = a == 0.0 ? 1.0 : std::sin(a) / a;
      table_host[(int64_t)p * kw + k] = (float)(sinc * c * c * (base / (double)o));
    }
  }
  return need;
}

extern "C" int hft_resample_create(hft_resample_plan** out, const float* kernel_host, int32_t orig_reduced, int32_t new_reduced, int32_t width);

extern "C" int hft_resample_create_hz(hft_resample_plan** out, int32_t orig_hz, int32_t new_hz) {
  HFT_REQUIRE(out && orig_hz >= 1 && new_hz >= 1, HFT_ERR_ARG, "hft_resample_create_hz: bad argument");
  int32_t o = 0, n = 0, width = 0;
  const int64_t need = hft_resample_build_table(orig_hz, new_hz, nullptr, 0, &o, &n, &width);
  std::vector<float> table((size_t)need);
  hft_resample_build_table(orig_hz, new_hz, table.data(), need, &o, &n, &width);
  return hft_resample_create(out, table.data(), o, n, width);
}

extern "C" int hft_resample_create(hft_resample_plan** out, const float* kernel_host, int32_t orig_reduced, int32_t new_reduced, int32_t width) {
  HFT_REQUIRE(out && kernel_host && orig_reduced >= 1 && new_reduced >= 1 && width >= 0, HFT_ERR_ARG, "hft_resample_create: bad argument");
  ResamplePlan* pl = new ResamplePlan();
  pl->o = orig_reduced; pl->n = new_reduced; pl->width = width; pl->kw = 2 * width + orig_reduced;
  const size_t smem = ((size_t)kRsI * pl->o + pl->kw + (size_t)kRsI * pl->n) * sizeof(float);
  if (smem > 200 * 1024) {
    delete pl;
    set_error("hft_resample_create: rate ratio %d/%d needs %zu bytes of shared memory per CTA (unsupported)", orig_reduced, new_reduced, smem);
    return HFT_ERR_UNSUPPORTED;
  }
  const size_t bytes = (size_t)pl->n * pl->kw * sizeof(float);
  HFT_CHECK_CUDA(cudaMalloc(&pl->d_kernel, bytes));
  HFT_CHECK_CUDA(cudaMemcpy(pl->d_kernel, kernel_host, bytes, cudaMemcpyHostToDevice));
  HFT_CHECK_CUDA(cudaFuncSetAttribute(resample_mono_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  *out = reinterpret_cast<hft_resample_plan*>(pl);
  return HFT_OK;
}

extern "C" int hft_resample_destroy(hft_resample_plan* plan) {
  if (!plan) return HFT_OK;
  ResamplePlan* pl = reinterpret_cast<ResamplePlan*>(plan);
  cudaFree(pl->d_kernel);
  delete pl;
  return HFT_OK;
}

extern "C" int64_t hft_resample_num_samples(const hft_resample_plan* plan, int64_t n_in) {
  const ResamplePlan* pl = reinterpret_cast<const ResamplePlan*>(plan);
  if (!pl || n_in < 0) return -1;
  return (n_in * pl->n + pl->o - 1) / pl->o;             // ceil(new * length / orig), functional.py _apply_sinc_resample_kernel
}

extern "C" int hft_resample_mono_f32(hft_resample_plan* plan, const float* wav_dev, int32_t channels, int64_t n_in, float* out_dev, int64_t n_out, void* stream) {
  HFT_REQUIRE(plan && channels >= 1 && n_in >= 0, HFT_ERR_ARG, "hft_resample_mono_f32: bad argument");
  ResamplePlan* pl = reinterpret_cast<ResamplePlan*>(plan);
  HFT_REQUIRE(n_out == hft_resample_num_samples(plan, n_in), HFT_ERR_ARG, "hft_resample_mono_f32: n_out=%lld, expected %lld", (long long)n_out,
              (long long)hft_resample_num_samples(plan, n_in));
  reset_launch_count();
  if (n_out == 0) return HFT_OK;
  HFT_REQUIRE(wav_dev && out_dev, HFT_ERR_ARG, "hft_resample_mono_f32: NULL buffer");
  const long long n_i = (n_out + pl->n - 1) / pl->n;
  const long long blocks = (n_i + kRsI - 1) / kRsI;
  HFT_REQUIRE(blocks < (1ll << 31), HFT_ERR_UNSUPPORTED, "hft_resample_mono_f32: clip too long");
  const size_t smem = ((size_t)kRsI * pl->o + pl->kw + (size_t)kRsI * pl->n) * sizeof(float);
  {
    LaunchScope ls(HFT_KCLASS_LOGMEL, stream);
    resample_mono_kernel<<<(unsigned)blocks, kRsWarps * 32, smem, (cudaStream_t)stream>>>(wav_dev, channels, n_in, pl->d_kernel, pl->o, pl->n, pl->width, pl->kw, out_dev, n_out);
  }
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}
