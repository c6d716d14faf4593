// roi_pool.cu -- RoIPool (max) forward / backward, NCHW.
// API compatibility with _C.roi_pool_forward / _C.roi_pool_backward (reference
// csrc/ROIPool.h:11-48; semantics of csrc/cuda/ROIPool_cuda.cu:17-108: integer-rounded
// RoI, floor/ceil bin edges, empty bin -> 0 with argmax -1).  No model of the reference
// calls it (modeling/poolers.py:48,66 hard-code ROIAlign), so it is a plain
// element-per-thread kernel with bins fastest for coalesced stores.
#include <float.h>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
roi_pool_fwd_kernel(const float* __restrict__ input, int C, int H, int W, const float* __restrict__ rois,
                    long long total, float scale, int PH, int PW, float* __restrict__ out,
                    int32_t* __restrict__ argmax) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(idx % PW), ph = (int)((idx / PW) % PH);
    const int c = (int)((idx / PW / PH) % C);
    const long long n = idx / PW / PH / C;
    const float* roi = rois + n * 5;
    const int b = (int)roi[0];
    const int sw = (int)roundf(__fmul_rn(roi[1], scale)), sh = (int)roundf(__fmul_rn(roi[2], scale));
    const int ew = (int)roundf(__fmul_rn(roi[3], scale)), eh = (int)roundf(__fmul_rn(roi[4], scale));
    const int rw = max(ew - sw + 1, 1), rh = max(eh - sh + 1, 1);
    const float bh = __fdiv_rn((float)rh, (float)PH), bw = __fdiv_rn((float)rw, (float)PW);
    int hs = (int)floorf(__fmul_rn((float)ph, bh)), ws = (int)floorf(__fmul_rn((float)pw, bw));
    int he = (int)ceilf(__fmul_rn((float)(ph + 1), bh)), we = (int)ceilf(__fmul_rn((float)(pw + 1), bw));
    hs = min(max(hs + sh, 0), H);
    he = min(max(he + sh, 0), H);
    ws = min(max(ws + sw, 0), W);
    we = min(max(we + sw, 0), W);
    const bool empty = he <= hs || we <= ws;
    float best = empty ? 0.f : -FLT_MAX;
    int besti = -1;
    const float* plane = input + ((size_t)b * C + c) * H * W;
    for (int y = hs; y < he; ++y)
      for (int x = ws; x < we; ++x) {
        float v = __ldg(plane + y * W + x);
        if (v > best) {
          best = v;
          besti = y * W + x;
        }
      }
    out[idx] = best;
    argmax[idx] = besti;
  }
}

__global__ void __launch_bounds__(256)
roi_pool_bwd_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ argmax,
                    const float* __restrict__ rois, long long total, int C, int H, int W, int NB,
                    float* __restrict__ grad_in) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int a = argmax[idx];
    if (a < 0) continue;
    const int c = (int)((idx / NB) % C);
    const long long n = idx / NB / C;
    const int b = (int)rois[n * 5];
    atomicAdd(grad_in + ((size_t)b * C + c) * H * W + a, grad_out[idx]);
  }
}

unsigned grid_for(long long total) {
  long long g = (total + 255) / 256;
  long long cap = (long long)b200::sm_count() * 32;
  return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" int b200_roi_pool_forward(const float* input, int batch, int channels, int height, int width,
                                     const float* rois, int64_t n_rois, float spatial_scale, int pooled_h,
                                     int pooled_w, float* out, int32_t* argmax, void* stream) {
  using namespace b200;
  B200_REQUIRE(batch > 0 && channels > 0 && height > 0 && width > 0 && pooled_h > 0 && pooled_w > 0 && n_rois >= 0,
               "roi_pool: bad shape");
  if (n_rois == 0) return B200_OK;
  B200_REQUIRE(input && rois && out && argmax, "roi_pool: null pointer");
  const long long total = (long long)n_rois * channels * pooled_h * pooled_w;
  roi_pool_fwd_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      input, channels, height, width, rois, total, spatial_scale, pooled_h, pooled_w, out, argmax);
  B200_CHECK_LAUNCH("roi_pool_fwd_kernel");
  return B200_OK;
}

extern "C" int b200_roi_pool_backward(const float* grad_out, const int32_t* argmax, const float* rois,
                                      int64_t n_rois, int batch, int channels, int height, int width,
                                      int pooled_h, int pooled_w, float* grad_in, void* stream) {
  using namespace b200;
  B200_REQUIRE(batch > 0 && channels > 0 && height > 0 && width > 0 && pooled_h > 0 && pooled_w > 0 && n_rois >= 0,
               "roi_pool_bwd: bad shape");
  if (n_rois == 0) return B200_OK;
  B200_REQUIRE(grad_out && argmax && rois && grad_in, "roi_pool_bwd: null pointer");
  const long long total = (long long)n_rois * channels * pooled_h * pooled_w;
  roi_pool_bwd_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      grad_out, argmax, rois, total, channels, height, width, pooled_h * pooled_w, grad_in);
  B200_CHECK_LAUNCH("roi_pool_bwd_kernel");
  return B200_OK;
}
