// Error plumbing, version and device facts of the C ABI (include/hft_sm100.h).
#include "common.cuh"
#include "hft_internal.h"
#include <stdlib.h>
#include <vector>

namespace hft {

static thread_local char g_err[1024] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
void count_launch(int n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }

struct ProfRec { cudaEvent_t a, b; };
static thread_local bool g_prof = false;
static thread_local std::vector<ProfRec>* g_recs = nullptr;   // [HFT_KCLASS_COUNT]

LaunchScope::LaunchScope(int kc, void* st) : kclass(kc), stream(st), ev0(nullptr) {
  g_launches += 1;
  if (g_prof) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, (cudaStream_t)stream);
    ev0 = e;
  }
}
// HFT_DEBUG_SYNC=1: synchronise after every launch and report the first failing one (ordinal within the call + kernel class)
static bool debug_sync() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HFT_DEBUG_SYNC"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
LaunchScope::~LaunchScope() {
  if (debug_sync()) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "[hft debug] launch #%lld (class %d) failed: %s\n", g_launches, kclass, cudaGetErrorString(e)); fflush(stderr); }
  }
  if (ev0) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, (cudaStream_t)stream);
    if (!g_recs) g_recs = new std::vector<ProfRec>[HFT_KCLASS_COUNT];
    g_recs[kclass].push_back(ProfRec{(cudaEvent_t)ev0, e});
  }
}

int num_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}

}  // namespace hft

extern "C" int hft_version(void) { return 1; }
extern "C" const char* hft_last_error(void) { return hft::get_error(); }
extern "C" int hft_device_sm_count(void) { return hft::num_sms(); }
extern "C" int64_t hft_last_launch_count(void) { return hft::g_launches; }

extern "C" int hft_profile_enable(int on) { hft::g_prof = on != 0; return 0; }
extern "C" int hft_profile_read(int kclass, double* ms, int64_t* launches) {
  using namespace hft;
  HFT_REQUIRE(kclass >= 0 && kclass < HFT_KCLASS_COUNT && ms && launches, HFT_ERR_ARG, "hft_profile_read: bad argument");
  *ms = 0.0;
  *launches = 0;
  if (!g_recs) return 0;
  for (auto& r : g_recs[kclass]) {
    float t = 0.f;
    HFT_CHECK_CUDA(cudaEventSynchronize(r.b));
    HFT_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    *ms += t;
    *launches += 1;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_recs[kclass].clear();
  return 0;
}

// ---- measurement probe: the chip's fp32 FMA issue rate (the roofline denominator of the CUDA-core kernels; MEASURED_PEAKS.json holds HBM and
// bf16 tensor figures only).  Every thread runs 16 independent FMA chains; 2 * 16 * iters flop per thread.
namespace hft {
__global__ void __launch_bounds__(256) fp32_fma_probe_kernel(int iters, float seed, float* __restrict__ out) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + (float)(threadIdx.x + i);
  const float m = 0.999f + seed * 1e-9f, c = 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;          // never true: keeps the chains alive
}
}  // namespace hft

extern "C" int hft_probe_fp32_fma(int32_t iters, float* scratch_dev, double* flop_out, void* stream) {
  HFT_REQUIRE(iters >= 1 && scratch_dev && flop_out, HFT_ERR_ARG, "hft_probe_fp32_fma: bad argument");
  const int sms = hft::num_sms(), blocks = sms * 8;
  hft::reset_launch_count();
  {
    hft::LaunchScope ls(HFT_KCLASS_NORM, stream);
    hft::fp32_fma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 0.5f, scratch_dev);
  }
  HFT_CHECK_CUDA(cudaGetLastError());
  *flop_out = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
  return 0;
}
