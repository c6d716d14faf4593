// common.cuh -- shared helpers for the sm_100a RoI hot-path kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200det.h"

namespace b200 {

// thread-local last-error text behind b200_last_error_string()
void set_error(const char* fmt, ...);

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return B200_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return B200_ERR_CUDA;
}

#define B200_CHECK_LAUNCH(what)                                  \
  do {                                                           \
    int _rc = ::b200::check_cuda(cudaGetLastError(), what);      \
    if (_rc != B200_OK) return _rc;                              \
  } while (0)

#define B200_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::b200::set_error(__VA_ARGS__);           \
      return B200_ERR_INVALID_ARG;              \
    }                                           \
  } while (0)

// Raise a kernel's dynamic shared-memory limit only when a launch needs more than any launch
// before it (steady state issues no attribute calls, so steps can be captured in CUDA graphs).
struct SmemHighWater {
  size_t per_device[32] = {};  // the attribute is per device
};
template <typename K>
inline int ensure_dynamic_smem(K kernel, size_t bytes, SmemHighWater* hw, const char* what) {
  // (the 48 KB default limit counts static + dynamic shared memory, so small requests are
  // registered too: a kernel with 7 KB of static arrays needs the opt-in from 41 KB on)
  if (bytes <= 16 * 1024) return B200_OK;
  int dev = 0;
  cudaGetDevice(&dev);
  size_t& mark = hw->per_device[dev & 31];
  if (bytes <= mark) return B200_OK;
  int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), what);
  if (rc == B200_OK) mark = bytes;
  return rc;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA, non-tensor form) wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(phase)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  while (!mbar_try_wait(bar, phase)) {
  }
}
// global -> shared bulk copy, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace b200
