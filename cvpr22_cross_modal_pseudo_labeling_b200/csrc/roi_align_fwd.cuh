// roi_align_fwd.cuh -- pieces shared by the RoIAlign forward kernels (roi_align_fwd.cu: exact
// marching + generic gather; roi_align_fwd_sep.cu: separable fast-math marching).
#pragma once
#include "roi_geom.cuh"

namespace b200 {

constexpr int kChunk = 64;  // channels per output tile of the marching kernels

struct RoiHeader {
  int batch, level;
  float x1, y1, x2, y2;
};

__device__ __forceinline__ RoiHeader load_roi(const float* __restrict__ rois, long long r,
                                              const LevelTable& lt) {
  const float* p = rois + r * 5;
  RoiHeader h;
  h.batch = (int)p[0];
  h.x1 = p[1];
  h.y1 = p[2];
  h.x2 = p[3];
  h.y2 = p[4];
  h.level = lt.n_levels == 1 ? 0 : fpn_level(h.x1, h.y1, h.x2, h.y2, lt.k_min, lt.k_max);
  return h;
}

__device__ __forceinline__ void zero_fill(float* __restrict__ p, int n, int tid, int nthreads) {
  for (int i = tid; i < n; i += nthreads) p[i] = 0.f;
}

// ---------------------------------------------------------------------------------------
// Shared-memory output tile [64 channels x NB bins] of the marching kernels.  A warp stores one
// float per lane with lanes spread over channel quads, i.e. at a stride of 4*NB floats: for
// NB = 196 (14x14) that is 16 banks apart, an 8-way conflict.  When NB % 4 == 0 the row of
// channel c is therefore shifted by 4 * (c >> 3) floats (monotonic, so rows never overlap;
// keeps rows 16-byte aligned for the vectorised copy-out; leaves at most a 2-way conflict).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int tile_row(int c, int NB, bool swz) { return c * NB + (swz ? 4 * (c >> 3) : 0); }
constexpr int kTilePadFloats = 32;

__device__ __forceinline__ void tile_copy_out(const float* out_s, float* __restrict__ dst_f, int NB, bool swz,
                                              int tid, int nthreads) {
  float4* dst = reinterpret_cast<float4*>(dst_f);
  const float4* src = reinterpret_cast<const float4*>(out_s);
  if (!swz) {
    for (int i = tid; i < kChunk * NB / 4; i += nthreads) __stcs(dst + i, src[i]);
  } else {
    const int rowv = NB >> 2;  // float4 per channel row
    for (int i = tid; i < kChunk * rowv; i += nthreads) {
      const int c = i / rowv;
      __stcs(dst + i, src[i + (c >> 3)]);
    }
  }
}

// Per-channel mean over the bins of the tile (nn.AvgPool2d(kernel_size = pooled size) of the
// pooled block, roi_box_predictors.py:16-17,:62): sequential fp32 sum in bin order, then one
// division, as torch's avg_pool2d kernel does.  Threads 0..63, one channel each; bank-conflict
// free for odd NB, 2-way otherwise.
__device__ __forceinline__ void tile_mean_out(const float* tile, float* __restrict__ mean, int NB, bool swz, int tid) {
  if (tid < kChunk) {
    const float* row = tile + tile_row(tid, NB, swz);
    float s = 0.f;
    for (int i = 0; i < NB; ++i) s = __fadd_rn(s, row[i]);
    mean[tid] = __fdiv_rn(s, (float)NB);
  }
}

// the reverse: a contiguous [64 x NB] block of global memory into the (shifted) tile
__device__ __forceinline__ void tile_copy_in(float* tile, const float* __restrict__ src_f, int NB, bool swz, int tid,
                                             int nthreads) {
  const float4* src = reinterpret_cast<const float4*>(src_f);
  float4* dst = reinterpret_cast<float4*>(tile);
  if (!swz) {
    for (int i = tid; i < kChunk * NB / 4; i += nthreads) dst[i] = __ldcs(src + i);
  } else {
    const int rowv = NB >> 2;
    for (int i = tid; i < kChunk * rowv; i += nthreads) {
      const int c = i / rowv;
      dst[i + (c >> 3)] = __ldcs(src + i);
    }
  }
}

// roi_align_fwd_sep.cu
int launch_forward_sep(const LevelTable& lt, int C, const float* rois, int64_t n_rois, int PH, int PW, float* out,
                       float* out_mean, int32_t* out_levels, int variant, cudaStream_t st);

// roi_align_fwd_rows.cu
bool rows_kernel_applies(const LevelTable& lt, int C, int PH, int PW);
int launch_forward_rows(const LevelTable& lt, int C, bool bf16_maps, const float* rois, int64_t n_rois, float* out,
                        float* out_mean, int32_t* out_levels, int32_t* order_ws, int variant, cudaStream_t st);
size_t rows_order_workspace_bytes(int64_t n_rois);
int launch_roi_order(const LevelTable& lt, const float* rois, int64_t n_rois, int32_t* order, cudaStream_t st);

}  // namespace b200
