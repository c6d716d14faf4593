// layout.cu -- NCHW <-> NHWC staging copies (fp32), tiled through shared memory so both
// the read and the write side are coalesced.  Used when a caller's feature maps arrive
// NCHW-contiguous (the reference's layout, csrc/cuda/ROIAlign_cuda.cu:286) and it wants
// the NHWC RoIAlign path.  Pure HBM copy: 8 bytes of traffic per element.
#include "common.cuh"

namespace {

constexpr int kT = 32;

// view: src is [rows, cols] row-major per batch item, dst is [cols, rows]
__global__ void __launch_bounds__(kT * 8)
transpose2d_batched(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[kT][kT + 1];
  const size_t plane = (size_t)rows * cols;
  const float* s = src + plane * blockIdx.z;
  float* d = dst + plane * blockIdx.z;
  const int c0 = blockIdx.x * kT, r0 = blockIdx.y * kT;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int k = 0; k < kT; k += 8) {
    int r = r0 + ty + k, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + k][tx] = s[(size_t)r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kT; k += 8) {
    int c = c0 + ty + k, r = r0 + tx;
    if (r < rows && c < cols) d[(size_t)c * rows + r] = tile[tx][ty + k];
  }
}

int launch(const float* src, float* dst, int batch, int rows, int cols, void* stream) {
  using namespace b200;
  B200_REQUIRE(src && dst, "layout: null pointer");
  B200_REQUIRE(batch > 0 && rows > 0 && cols > 0, "layout: bad shape");
  B200_REQUIRE(batch <= 65535 && ceil_div(rows, kT) <= 65535, "layout: shape exceeds grid limits");
  dim3 grid(ceil_div(cols, kT), ceil_div(rows, kT), batch);
  transpose2d_batched<<<grid, dim3(kT, 8), 0, static_cast<cudaStream_t>(stream)>>>(src, dst, rows, cols);
  B200_CHECK_LAUNCH("transpose2d_batched");
  return B200_OK;
}

}  // namespace

// [B, C, H*W] -> [B, H*W, C]
extern "C" int b200_nchw_to_nhwc(const float* src, float* dst, int batch, int channels, int height, int width,
                                 void* stream) {
  return launch(src, dst, batch, channels, height * width, stream);
}

// [B, H*W, C] -> [B, C, H*W]
extern "C" int b200_nhwc_to_nchw(const float* src, float* dst, int batch, int channels, int height, int width,
                                 void* stream) {
  return launch(src, dst, batch, height * width, channels, stream);
}
