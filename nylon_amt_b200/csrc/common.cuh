// Shared host/device helpers for libhft_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

namespace hft {

// ---- error plumbing (C-ABI convention: int return + hft_last_error()) ---------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define HFT_CHECK_CUDA(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      hft::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                         \
    }                                                                                         \
  } while (0)

#define HFT_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      hft::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

enum : int { HFT_OK = 0, HFT_ERR_ARG = 10001, HFT_ERR_UNSUPPORTED = 10002, HFT_ERR_STATE = 10003 };

int num_sms();

// Opt a kernel in to more than 48 KB of dynamic shared memory, once per device (the attribute is per device; one process may drive
// several GPUs even though the usual deployment is one process per GPU).
#define HFT_SET_MAX_SMEM(fn, bytes)                                                                                   \
  do {                                                                                                                \
    static bool done_[64] = {};                                                                                       \
    int dev_ = 0;                                                                                                     \
    cudaGetDevice(&dev_);                                                                                             \
    if (!done_[dev_ & 63]) {                                                                                          \
      HFT_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));            \
      done_[dev_ & 63] = true;                                                                                        \
    }                                                                                                                 \
  } while (0)

#ifdef __CUDACC__
// ---- mbarrier / bulk-copy PTX (sm_90+; UBLKCP in SASS) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)      // suspend-time hint: the waiting warp sleeps until the phase completes
      : "memory");                                            // (or the hint expires) instead of spinning on the issue slots
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk async copy global -> shared (TMA engine, no tensor map).  16-byte aligned src/dst/size.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
#endif

}  // namespace hft
