// roi_align_bwd.cu -- fused multi-level RoIAlign backward (gradient w.r.t. the features).
//
// Replaces _C.roi_align_backward (reference csrc/ROIAlign.h:27-45, kernel
// csrc/cuda/ROIAlign_cuda.cu:178-254) for all FPN levels in one launch.  Each addend is
// rounded exactly as the reference rounds it -- (top * w_k) / count, :239-242 -- only the
// order of the additions (atomics there, atomics here) is free.
//
//   NHWC: a thread owns (4 channels, one bin); every tap is ONE 128-bit vector reduction
//         (red.global.add.v4.f32) instead of four scalar atomics.
//   NCHW: a thread owns (channel, bin) with scalar reductions, bins fastest so that the
//         reads of grad_out are coalesced.
#include "roi_geom.cuh"

namespace b200 {
namespace {

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__device__ __forceinline__ float gterm(float top, float w, float count) {
  return __fdiv_rn(__fmul_rn(top, w), count);
}

template <int kLayout>
__global__ void __launch_bounds__(256)
roi_align_bwd_generic(const LevelGradTable lt, int C, const float* __restrict__ rois, int PH, int PW,
                      int sampling_ratio, int c_per_cta, int n_cchunks, const float* __restrict__ grad_out) {
  const long long r = blockIdx.x / n_cchunks;
  const int ck = blockIdx.x % n_cchunks;
  const float* p = rois + r * 5;
  const int batch = (int)p[0];
  const float x1 = p[1], y1 = p[2], x2 = p[3], y2 = p[4];
  const int level = lt.n_levels == 1 ? 0 : fpn_level(x1, y1, x2, y2, lt.k_min, lt.k_max);
  if (level < 0) return;
  const int H = lt.H[level], W = lt.W[level];
  float* gfeat = lt.data[level];
  const RoiGeom g = roi_geometry(x1, y1, x2, y2, lt.scale[level], PH, PW, sampling_ratio);
  const float count = (float)(g.grid_h * g.grid_w);
  const int NB = PH * PW;
  const int c_begin = ck * c_per_cta;
  const int c_count = min(c_per_cta, C - c_begin);
  const float* go_roi = grad_out + (size_t)r * C * NB;

  if (kLayout == B200_LAYOUT_NHWC) {
    const int nq = c_count >> 2;
    for (int u = threadIdx.x; u < nq * NB; u += blockDim.x) {
      const int q = u % nq, bin = u / nq;
      const int ph = bin / PW, pw = bin - ph * PW;
      const float* go = go_roi + (size_t)(c_begin + 4 * q) * NB + bin;
      const float t0 = go[0], t1 = go[NB], t2 = go[2 * NB], t3 = go[3 * NB];
      float* base = gfeat + (size_t)batch * H * W * C + c_begin + 4 * q;
      for (int iy = 0; iy < g.grid_h; ++iy) {
        bool oky;
        const AxisTap ty = axis_sample(g.start_h, ph, g.bin_h, iy, g.grid_h, H, oky);
        for (int ix = 0; ix < g.grid_w; ++ix) {
          bool okx;
          const AxisTap tx = axis_sample(g.start_w, pw, g.bin_w, ix, g.grid_w, W, okx);
          if (!(oky && okx)) continue;
          const float w[4] = {__fmul_rn(ty.h, tx.h), __fmul_rn(ty.h, tx.l), __fmul_rn(ty.l, tx.h),
                              __fmul_rn(ty.l, tx.l)};
          const size_t at[4] = {((size_t)ty.lo * W + tx.lo) * C, ((size_t)ty.lo * W + tx.hi) * C,
                                ((size_t)ty.hi * W + tx.lo) * C, ((size_t)ty.hi * W + tx.hi) * C};
#pragma unroll
          for (int k = 0; k < 4; ++k)
            red_add_v4(base + at[k], gterm(t0, w[k], count), gterm(t1, w[k], count), gterm(t2, w[k], count),
                       gterm(t3, w[k], count));
        }
      }
    }
  } else {
    for (int u = threadIdx.x; u < c_count * NB; u += blockDim.x) {
      const int c = c_begin + u / NB, bin = u % NB;
      const int ph = bin / PW, pw = bin - ph * PW;
      const float top = go_roi[(size_t)c * NB + bin];
      float* plane = gfeat + ((size_t)batch * C + c) * H * W;
      for (int iy = 0; iy < g.grid_h; ++iy) {
        bool oky;
        const AxisTap ty = axis_sample(g.start_h, ph, g.bin_h, iy, g.grid_h, H, oky);
        for (int ix = 0; ix < g.grid_w; ++ix) {
          bool okx;
          const AxisTap tx = axis_sample(g.start_w, pw, g.bin_w, ix, g.grid_w, W, okx);
          if (!(oky && okx)) continue;
          atomicAdd(plane + ty.lo * W + tx.lo, gterm(top, __fmul_rn(ty.h, tx.h), count));
          atomicAdd(plane + ty.lo * W + tx.hi, gterm(top, __fmul_rn(ty.h, tx.l), count));
          atomicAdd(plane + ty.hi * W + tx.lo, gterm(top, __fmul_rn(ty.l, tx.h), count));
          atomicAdd(plane + ty.hi * W + tx.hi, gterm(top, __fmul_rn(ty.l, tx.l), count));
        }
      }
    }
  }
}

}  // namespace
}  // namespace b200

extern "C" int b200_roi_align_backward(const b200_level_grad* levels, int n_levels, int layout, int batch,
                                       int channels, const float* rois, int64_t n_rois, int pooled_h,
                                       int pooled_w, int sampling_ratio, const float* grad_out, void* stream) {
  using namespace b200;
  B200_REQUIRE(layout == B200_LAYOUT_NCHW || layout == B200_LAYOUT_NHWC, "roi_align_bwd: bad layout %d", layout);
  B200_REQUIRE(batch > 0 && channels > 0 && pooled_h > 0 && pooled_w > 0 && n_rois >= 0 && sampling_ratio >= 0,
               "roi_align_bwd: bad shape");
  B200_REQUIRE(levels && n_levels >= 1 && n_levels <= B200_MAX_LEVELS, "roi_align_bwd: n_levels must be 1..%d",
               B200_MAX_LEVELS);
  if (n_rois == 0) return B200_OK;
  B200_REQUIRE(rois && grad_out, "roi_align_bwd: null rois / grad_out");
  LevelGradTable lt;
  for (int l = 0; l < n_levels; ++l) {
    B200_REQUIRE(levels[l].data && aligned16(levels[l].data), "roi_align_bwd: level %d data null or misaligned",
                 l);
    B200_REQUIRE(levels[l].height > 0 && levels[l].width > 0 && levels[l].spatial_scale > 0.f,
                 "roi_align_bwd: level %d has an empty shape or non-positive scale", l);
    lt.data[l] = levels[l].data;
    lt.H[l] = levels[l].height;
    lt.W[l] = levels[l].width;
    lt.scale[l] = levels[l].spatial_scale;
  }
  lt.n_levels = n_levels;
  lt.k_min = -log2f(levels[0].spatial_scale);
  lt.k_max = -log2f(levels[n_levels - 1].spatial_scale);
  const int NB = pooled_h * pooled_w;
  int c_per_cta;
  if (layout == B200_LAYOUT_NHWC) {
    B200_REQUIRE(channels % 4 == 0, "roi_align_bwd: NHWC needs channels %% 4 == 0");
    c_per_cta = channels < 64 ? channels : 64;
  } else {
    c_per_cta = (2048 + NB - 1) / NB;
    if (c_per_cta > channels) c_per_cta = channels;
  }
  const int n_cchunks = (channels + c_per_cta - 1) / c_per_cta;
  const int64_t grid = n_rois * n_cchunks;
  B200_REQUIRE(grid < (int64_t)1 << 31, "roi_align_bwd: too many RoIs for one launch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (layout == B200_LAYOUT_NHWC)
    roi_align_bwd_generic<B200_LAYOUT_NHWC><<<(unsigned)grid, 256, 0, st>>>(
        lt, channels, rois, pooled_h, pooled_w, sampling_ratio, c_per_cta, n_cchunks, grad_out);
  else
    roi_align_bwd_generic<B200_LAYOUT_NCHW><<<(unsigned)grid, 256, 0, st>>>(
        lt, channels, rois, pooled_h, pooled_w, sampling_ratio, c_per_cta, n_cchunks, grad_out);
  B200_CHECK_LAUNCH("roi_align_bwd_generic");
  return B200_OK;
}
