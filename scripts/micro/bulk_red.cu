// Probe: throughput of cp.reduce.async.bulk (shared -> global, add.f32) against red.global.add.v4.f32 from registers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_red bulk_red.cu
// Every CTA adds chunks of NC KB (a row of NC pixels x 256 fp32 channels) at pseudo-random rows of a buffer that is
// either L2 resident (17 MB) or not (1.4 GB).  mode 0: bulk reduce (one thread issues, up to 8 groups in flight);
// mode 1: the whole CTA issues red.global.add.v4.f32 (16 B per lane, 512 B per warp instruction).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512, 1) probe(float* base, size_t n_rows, int row_bytes, int chunk_bytes, int iters, int mode) {
  extern __shared__ __align__(128) unsigned char buf[];
  for (int i = threadIdx.x; i < chunk_bytes / 4; i += blockDim.x) reinterpret_cast<float*>(buf)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  uint32_t state = blockIdx.x * 0x9E3779B9u + 4242u;
  if (mode == 0) {
    if (threadIdx.x == 0) {
      for (int it = 0; it < iters; ++it) {
        state = state * 1664525u + 1013904223u;
        char* dst = reinterpret_cast<char*>(base) + (size_t)((state >> 4) % n_rows) * row_bytes;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(buf)), "r"(chunk_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    const float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
    for (int it = 0; it < iters; ++it) {
      state = state * 1664525u + 1013904223u;
      char* dst = reinterpret_cast<char*>(base) + (size_t)((state >> 4) % n_rows) * row_bytes;
      for (int o = threadIdx.x * 16; o < chunk_bytes; o += blockDim.x * 16)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
  }
}

int main() {
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int regime = 0; regime < 2; ++regime) {
    const size_t bytes = regime == 0 ? (size_t)17 << 20 : (size_t)1400 << 20;
    float* base; cudaMalloc(&base, bytes); cudaMemset(base, 0, bytes);
    for (int nc : {4, 8, 16, 28}) for (int mode = 0; mode < 2; ++mode) {
      const int chunk = nc * 1024, row_bytes = 32 * 1024, iters = 400;
      const size_t n_rows = bytes / row_bytes;
      probe<<<148, 512, 64 * 1024>>>(base, n_rows, row_bytes, chunk, 10, mode);
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a); probe<<<148, 512, 64 * 1024>>>(base, n_rows, row_bytes, chunk, iters, mode); cudaEventRecord(b);
      cudaError_t e = cudaEventSynchronize(b);
      if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, a, b);
      const double tot = 148.0 * iters * chunk;
      printf("%-4s chunk %2d KB %-22s %7.3f ms  %6.0f GB/s of reduced bytes\n", regime == 0 ? "l2" : "dram", nc,
             mode == 0 ? "cp.reduce.async.bulk" : "red.global.add.v4.f32", ms, tot / ms / 1e6);
    }
    cudaFree(base);
  }
  return 0;
}
