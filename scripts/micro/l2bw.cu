// L2 -> SM read bandwidth probe (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2bw l2bw.cu).  Every warp instruction reads 32 x 16 B split into groups of
// `seg` lanes; each group reads seg*16 contiguous bytes at a pseudo-random, seg*16-aligned place of
// an L2-resident buffer (seg = 32: one 512 B run; 16: two 256 B runs -- the RoIAlign march
// pattern; 8: four 128 B runs).  `ilp` independent loads are issued before any is consumed.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void __launch_bounds__(256) rd(const float4* __restrict__ p, unsigned n_vec_mask, int seg_shift, int n_iter,
                                          float4* sink) {
  float4 acc = make_float4(0, 0, 0, 0);
  const unsigned gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const unsigned grp = lane >> seg_shift, within = lane & ((1u << seg_shift) - 1);
  unsigned state = gwarp * 0x9E3779B9u + 12345u;
  for (int it = 0; it < n_iter; ++it) {
    float4 v[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      state = state * 1664525u + 1013904223u;
      const unsigned h = (state ^ (grp * 0x85EBCA6Bu)) * 0xC2B2AE35u;
      const unsigned idx = (((h >> 8) << seg_shift) + within) & n_vec_mask;
      v[u] = __ldg(p + idx);
    }
#pragma unroll
    for (int u = 0; u < ILP; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  if (acc.x == 123.456f) sink[0] = acc;
}
template <int ILP>
void run(const float4* p, unsigned mask, float4* sink, size_t mb) {
  for (int seg_shift : {5, 4, 3}) for (int ctas : {148 * 3, 148 * 6, 148 * 8}) {
    const int n_iter = 2000 / ILP;
    rd<ILP><<<ctas, 256>>>(p, mask, seg_shift, 10, sink);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); rd<ILP><<<ctas, 256>>>(p, mask, seg_shift, n_iter, sink); cudaEventRecord(b);
    cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b);
    const double bytes = (double)ctas * 256 * 16 * n_iter * ILP;
    printf("buf %4zu MB ilp %d seg %3d B ctas/SM %d: %.0f GB/s\n", mb, ILP, 16 << seg_shift, ctas / 148, bytes / ms / 1e6);
  }
}
int main() {
  for (size_t mb : {64, 1024}) {
    size_t bytes = mb << 20;
    float4* p; float4* sink;
    cudaMalloc(&p, bytes); cudaMalloc(&sink, 16); cudaMemset(p, 0, bytes);
    unsigned mask = (unsigned)(bytes / 16 - 1);
    run<4>(p, mask, sink, mb);
    run<8>(p, mask, sink, mb);
    cudaFree(p); cudaFree(sink);
  }
  return 0;
}
