// Lane-serial CPU run of the per-lane log-mel building blocks (index-math check, no GPU needed).
// Build/run: see tests/test_logmel_hostsim.py.  Reads frames from stdin? No: exposes a C function.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../nylon_amt_b200/csrc/logmel_core.cuh"

extern "C" int hostsim_logmel(const float* wav, long n, const float* window, const float* fb /*[1025][256]*/,
                              float log_offset, float* out) {
  using namespace hft;
  long T = 1 + n / kHop;
  std::vector<float2> tw2(1024), twr(513);
  for (int k2 = 0; k2 < 32; ++k2)
    for (int n1 = 0; n1 < 32; ++n1) {
      double a = -2.0 * M_PI * (double)(n1 * k2) / 1024.0;
      tw2[k2 * 32 + n1] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int k = 0; k <= 512; ++k) {
    double a = -2.0 * M_PI * (double)k / 2048.0;
    twr[k] = make_float2((float)cos(a), (float)sin(a));
  }
  std::vector<float> melw(kMaxMelW, 0.f);
  std::vector<uint32_t> melinfo(kMelInfo);
  if (lm_pack_filterbank(fb, melw.data(), melinfo.data())) return 1;
  std::vector<float> xs(kNfft), P(kNfreq + 64, 0.f);   // taps past a band's end read finite padding
  std::vector<float2> Tt(32 * kTStride), Z(1024);
  for (long t = 0; t < T; ++t) {
    for (int i = 0; i < kNfft; ++i) {
      long s = t * kHop - kNfft / 2 + i;
      xs[i] = (s >= 0 && s < n) ? wav[s] : 0.f;
    }
    for (int lane = 0; lane < 32; ++lane) lm_rows(lane, xs.data(), window, tw2.data(), Tt.data());
    for (int lane = 0; lane < 32; ++lane) { float2 u[32]; lm_cols_load(lane, Tt.data(), u); lm_cols_store(lane, u, Z.data()); }
    for (int lane = 0; lane < 32; ++lane) lm_power(lane, Z.data(), twr.data(), P.data());
    for (int lane = 0; lane < 32; ++lane) lm_mel(lane, P.data(), melw.data(), melinfo.data(), log_offset, out + t * kNmels);
  }
  return 0;
}
