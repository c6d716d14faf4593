// mbarrier operation rate on one SM: W warps loop over { try_wait.parity on a completed phase; arrive (lane 0) } on
// per-warp or shared barriers.  Question behind it (roi_align_fwd_rows): is the consumers' per-entry hand-off
// (14 warps x (try_wait + arrive) per ring entry) bound by the latency of a warp's own chain or by a per-SM rate?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_rate mbar_rate.cu && ./mbar_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void k(int iters, long long* out) {
  __shared__ __align__(8) uint64_t bars[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (threadIdx.x == 0)
    for (int i = 0; i < 64; ++i) {
      uint32_t a = smem_u32(&bars[i]);
      // MODE 0/1: private barrier per warp (count 1); MODE 2: one barrier shared by all warps (count nw)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(MODE == 2 ? nw : 1));
    }
  __syncthreads();
  const uint32_t a = smem_u32(&bars[MODE == 2 ? 0 : warp]);
  uint32_t phase = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE != 1) {
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
    }
    if (MODE == 1) {
      // wait only: on a phase that completed long ago (parity of the phase "before" the first one)
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(a), "r"(1u) : "memory");
      if (!ok) break;
    } else {
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(phase) : "memory");
      phase ^= 1;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  long long* d;
  cudaMalloc(&d, 8 * 256);
  const int iters = 20000;
  for (int mode = 0; mode < 3; ++mode)
    for (int w : {1, 2, 4, 7, 8, 14, 16, 28, 32}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, 32 * w>>>(iters, d);
        if (mode == 1) k<1><<<148, 32 * w>>>(iters, d);
        if (mode == 2) k<2><<<148, 32 * w>>>(iters, d);
      }
      cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double c = (double)h[0] / iters;
      const char* names[3] = {"arrive + wait, private barrier", "try_wait on a completed phase only", "arrive + wait, one shared barrier"};
      printf("%-36s %2d warps: %7.1f cycles per iteration per warp, %6.1f cycles per warp-iteration per SM\n", names[mode], w, c, c / w);
    }
  return 0;
}
