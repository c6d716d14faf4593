// TMA probe for the grouped row-streaming RoIAlign: how fast can one SM / the chip pull (channel slice x
// column run) boxes of an NHWC fp32 map into a shared-memory ring?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_box tma_box.cu
// mode 0/1/2: tensor-map boxes of 64 / 128 / 256 channels x NC columns x 1 row (4 / 2 / 1 passes over a patch)
// mode 3: 1-D cp.async.bulk of NC * 1 KB (all 256 channels; what roi_align_fwd_rows does today)
// Patches: R rows x NC columns.  regime "l2": patches at random places of one 100x168 map (L2 resident);
// regime "dram": patches tile 16 maps of 200x336 without overlap, CTA-strided (every byte read once).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int kSlots = 16;
constexpr int kC = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma_box(void* dst, const CUtensorMap* map, int c0, int x, int y, int b, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(x), "r"(y), "r"(b), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Params {
  const float* base;
  int B, H, W, R, NC, mode, regime, patches_per_cta, sum, nprod, rows;
};

__global__ void __launch_bounds__(256, 1) probe(const __grid_constant__ CUtensorMap map, Params p, float* sink) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ uint64_t full[kSlots], empty[kSlots];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = p.mode == 0 ? 64 : p.mode == 1 ? 128 : 256;
  const int passes = kC / cs;
  const uint32_t ebytes = (uint32_t)p.NC * cs * 4 * p.rows;
  if (tid == 0) {
    for (int s = 0; s < kSlots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 3); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nslots = min(kSlots, (int)((200u * 1024u) / ebytes));
  const int per_row = p.W / p.NC, per_col = p.H / p.R;
  if (warp < p.nprod) {
    if (lane == 0) {
      uint32_t g = 0, state = blockIdx.x * 0x9E3779B9u + 777u;
      for (int k = 0; k < p.patches_per_cta; ++k) {
        int b, y0, x0;
        if (p.regime == 0) {
          state = state * 1664525u + 1013904223u;
          b = 0; y0 = (state >> 8) % (p.H - p.R); x0 = (state >> 20) % (p.W - p.NC);
        } else {
          long long id = (long long)k * gridDim.x + blockIdx.x;
          int per_img = per_row * per_col;
          b = (int)((id / per_img) % p.B); int q = (int)(id % per_img);
          y0 = (q / per_row) * p.R; x0 = (q % per_row) * p.NC;
        }
        for (int ps = 0; ps < passes; ++ps)
          for (int r = 0; r < p.R; r += p.rows, ++g) {
            if ((int)(g % p.nprod) != warp) continue;
            const int s = g % nslots;
            if (g >= (uint32_t)nslots) mbar_wait(&empty[s], ((g / nslots) - 1) & 1);
            mbar_expect(&full[s], ebytes);
            void* dst = ring + (size_t)s * ebytes;
            if (p.mode == 3)
              bulk_1d(dst, p.base + (((size_t)b * p.H + y0 + r) * p.W + x0) * kC, ebytes, &full[s]);
            else
              tma_box(dst, &map, ps * cs, x0, y0 + r, b, &full[s]);
          }
      }
    }
  } else {
    // three "consumer" warps: wait for the entry, optionally read it (one LDS.128 per lane per 512 B), give it back
    const int total = p.patches_per_cta * passes * (p.R / p.rows);
    float acc = 0.f;
    for (int g = 0; g < total; ++g) {
      const int s = g % nslots;
      mbar_wait(&full[s], (g / nslots) & 1);
      if (p.sum) {
        const float4* e = reinterpret_cast<const float4*>(ring + (size_t)s * ebytes);
        for (uint32_t i = (warp - p.nprod) * 32 + lane; i < ebytes / 16; i += 96) { float4 v = e[i]; acc += v.x + v.y + v.z + v.w; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 123.456f) sink[0] = acc;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encode fn\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  float* sink; cudaMalloc(&sink, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int regime = 0; regime < 2; ++regime) {
    const int B = regime == 0 ? 1 : 16, H = regime == 0 ? 100 : 200, W = regime == 0 ? 168 : 336;
    size_t bytes = (size_t)B * H * W * kC * 4;
    float* base; cudaMalloc(&base, bytes); cudaMemset(base, 0, bytes);
    for (int NC : {16, 24}) for (int mode = 0; mode < 4; ++mode) for (int nprod : {1, 2, 4}) for (int rows : {1, 2, 4}) {
      const int sum = 0;
      if (mode == 3 && rows > 1) continue;
      const int cs = mode == 0 ? 64 : mode == 1 ? 128 : 256;
      CUtensorMap map;
      cuuint64_t dims[4] = {(cuuint64_t)kC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
      cuuint64_t strides[3] = {(cuuint64_t)kC * 4, (cuuint64_t)W * kC * 4, (cuuint64_t)H * W * kC * 4};
      cuuint32_t box[4] = {(cuuint32_t)cs, (cuuint32_t)NC, (cuuint32_t)rows, 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d (cs %d NC %d)\n", (int)r, cs, NC); continue; }
      Params p; p.base = base; p.B = B; p.H = H; p.W = W; p.R = 20; p.NC = NC; p.mode = mode; p.regime = regime; p.sum = sum; p.nprod = nprod; p.rows = rows;
      const int per_img = (W / NC) * (H / p.R);
      p.patches_per_cta = regime == 0 ? 200 : (B * per_img) / 148;
      const uint32_t ebytes = NC * cs * 4; (void)ebytes;
      probe<<<148, 32 * (nprod + 3), 200 * 1024>>>(map, p, sink);
      cudaEvent_t a, b2; cudaEventCreate(&a); cudaEventCreate(&b2);
      cudaEventRecord(a); probe<<<148, 32 * (nprod + 3), 200 * 1024>>>(map, p, sink); cudaEventRecord(b2);
      cudaError_t e = cudaEventSynchronize(b2);
      if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 2; }
      float ms; cudaEventElapsedTime(&ms, a, b2);
      const double tot = 148.0 * p.patches_per_cta * p.R * NC * 1024.0;
      printf("%-4s mode %d (%3d ch%s) NC %2d prod %d rows %d: %7.3f ms  %6.0f GB/s  (%5.1f B/clk/SM @1.965)\n", regime == 0 ? "l2" : "dram", mode, cs,
             mode == 3 ? " 1-D" : " box", NC, nprod, rows, ms, tot / ms / 1e6, tot / ms / 1e6 / 148 / 1.965);
      fflush(stdout);
    }
    cudaFree(base);
  }
  return 0;
}
