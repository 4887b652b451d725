// Frame-level activations -> note events: the O(T x n_note) scans of AMT.mpe2note (reference hftt_code/model/amt.py:179-344) on the
// device.  The reference walks every (frame, pitch) cell in pure Python (0.2 s per 30 s clip); here three small kernels do the
// scans and the host assembles the (few) detected notes:
//   hft_note_peaks        peak flags of one activation map (amt.py:196-212: >= threshold and the first differing value on each side
//                         is lower; every point of a plateau counts)
//   hft_note_peak_times   sub-frame peak time of each detected peak from its two direct neighbours (amt.py:213-222), in the float32
//                         arithmetic numpy >= 2 performs on float32 inputs (every operation rounded separately, no FMA contraction)
//   hft_note_first_below  first frame after an onset where the mpe activation falls below its threshold (amt.py:262-271)
// Activation maps are [T][n_note] fp32 row-major, as AMT.transcript returns them.
#include "common.cuh"
#include "hft_internal.h"

namespace hft {

__global__ void note_peaks_kernel(const float* __restrict__ a, long long T, int N, float thr, uint8_t* __restrict__ flags) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * N) return;
  const long long i = idx / N;
  const int j = (int)(idx % N);
  const float v = a[idx];
  uint8_t f = 0;
  if (v >= thr) {
    bool ok = true;
    for (long long ii = i - 1; ii >= 0; --ii) {          // first differing value to the left
      const float u = a[ii * N + j];
      if (v > u) break;
      if (v < u) { ok = false; break; }
    }
    if (ok) {
      for (long long ii = i + 1; ii < T; ++ii) {         // and to the right
        const float u = a[ii * N + j];
        if (v > u) break;
        if (v < u) { ok = false; break; }
      }
    }
    f = ok ? 1 : 0;
  }
  flags[(long long)j * T + i] = f;                       // [n_note][T]: torch.nonzero then lists the peaks sorted by (pitch, frame)
}

__global__ void note_peak_times_kernel(const float* __restrict__ a, long long T, int N, const long long* __restrict__ idx2, long long n, double hop_sec,
                                       uint8_t* __restrict__ kind, float* __restrict__ t32) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int j = (int)idx2[2 * k];
  const long long i = idx2[2 * k + 1];
  uint8_t kd = 0;
  float t = 0.f;
  if (i > 0 && i < T - 1) {
    const float l = a[(i - 1) * N + j], b = a[i * N + j], r = a[(i + 1) * N + j];
    const float hop32 = __double2float_rn(hop_sec * 0.5);
    const float ih = __double2float_rn((double)i * hop_sec);
    if (l > r) {
      t = __fsub_rn(ih, __fdiv_rn(__fmul_rn(hop32, __fsub_rn(l, r)), __fsub_rn(b, r)));
      kd = 1;
    } else if (l < r) {
      t = __fadd_rn(ih, __fdiv_rn(__fmul_rn(hop32, __fsub_rn(r, l)), __fsub_rn(b, l)));
      kd = 1;
    }
  }
  kind[k] = kd;
  t32[k] = t;
}

__global__ void note_first_below_kernel(const float* __restrict__ mpe, long long T, int N, const long long* __restrict__ idx2, const long long* __restrict__ limit,
                                        long long n, float thr, long long* __restrict__ out) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int j = (int)idx2[2 * k];
  const long long lim = limit[k] < T ? limit[k] : T;
  long long r = -1;
  for (long long f = idx2[2 * k + 1] + 1; f < lim; ++f)
    if (mpe[f * N + j] < thr) { r = f; break; }
  out[k] = r;
}

}  // namespace hft

using namespace hft;

extern "C" int hft_note_peaks(const float* a_dev, int64_t T, int32_t n_note, float thr, uint8_t* flags_dev, void* stream) {
  HFT_REQUIRE(T >= 0 && n_note >= 1 && (T == 0 || (a_dev && flags_dev)), HFT_ERR_ARG, "hft_note_peaks: bad argument");
  reset_launch_count();
  if (T == 0) return HFT_OK;
  const long long n = (long long)T * n_note;
  HFT_REQUIRE(n < (1ll << 40), HFT_ERR_UNSUPPORTED, "hft_note_peaks: map too large");
  {
    LaunchScope ls(HFT_KCLASS_HEADS, stream);
    note_peaks_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a_dev, T, n_note, thr, flags_dev);
  }
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}

extern "C" int hft_note_peak_times(const float* a_dev, int64_t T, int32_t n_note, const int64_t* idx_dev, int64_t n, double hop_sec, uint8_t* kind_dev,
                                   float* t32_dev, void* stream) {
  HFT_REQUIRE(n >= 0 && (n == 0 || (a_dev && idx_dev && kind_dev && t32_dev)), HFT_ERR_ARG, "hft_note_peak_times: bad argument");
  reset_launch_count();
  if (n == 0) return HFT_OK;
  {
    LaunchScope ls(HFT_KCLASS_HEADS, stream);
    note_peak_times_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a_dev, T, n_note, reinterpret_cast<const long long*>(idx_dev), n, hop_sec,
                                                                                           kind_dev, t32_dev);
  }
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}

extern "C" int hft_note_first_below(const float* mpe_dev, int64_t T, int32_t n_note, const int64_t* idx_dev, const int64_t* limit_dev, int64_t n, float thr,
                                    int64_t* out_dev, void* stream) {
  HFT_REQUIRE(n >= 0 && (n == 0 || (mpe_dev && idx_dev && limit_dev && out_dev)), HFT_ERR_ARG, "hft_note_first_below: bad argument");
  reset_launch_count();
  if (n == 0) return HFT_OK;
  {
    LaunchScope ls(HFT_KCLASS_HEADS, stream);
    note_first_below_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mpe_dev, T, n_note, reinterpret_cast<const long long*>(idx_dev),
                                                                                            reinterpret_cast<const long long*>(limit_dev), n, thr,
                                                                                            reinterpret_cast<long long*>(out_dev));
  }
  HFT_CHECK_CUDA(cudaGetLastError());
  return HFT_OK;
}
