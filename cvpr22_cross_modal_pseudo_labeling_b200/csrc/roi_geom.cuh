// roi_geom.cuh -- RoIAlign / FPN-level geometry shared by the forward and backward kernels.
//
// Every expression mirrors the reference's fp32 evaluation order
// (csrc/cpu/ROIAlign_cpu.cpp:36-92,146-169; modeling/poolers.py:31-42;
// structures/bounding_box.py:226-230).  The _rn intrinsics are used so nvcc cannot
// contract a multiply and an add into one FMA: sample coordinates, bilinear weights and
// level ids come out bit-identical to the CPU reference.
#pragma once
#include "common.cuh"

namespace b200 {

struct LevelTable {
  const float* data[B200_MAX_LEVELS];
  int32_t H[B200_MAX_LEVELS];
  int32_t W[B200_MAX_LEVELS];
  float scale[B200_MAX_LEVELS];
  int32_t n_levels;
  float k_min, k_max;  // -log2(scale[0]), -log2(scale[n-1])  (poolers.py:72-75)
};

struct LevelGradTable {
  float* data[B200_MAX_LEVELS];
  int32_t H[B200_MAX_LEVELS];
  int32_t W[B200_MAX_LEVELS];
  float scale[B200_MAX_LEVELS];
  int32_t n_levels;
  float k_min, k_max;
};

// LevelMapper.__call__ (poolers.py:31-42): floor(4 + log2(sqrt(area)/224 + 1e-6)), clamped,
// minus k_min.  Returns -1 when the value is NaN (area < 0): in the reference such a RoI
// matches no level and its output rows stay zero (poolers.py:111-119).
__device__ __forceinline__ int fpn_level(float x1, float y1, float x2, float y2, float k_min, float k_max) {
  float w = __fadd_rn(__fsub_rn(x2, x1), 1.0f);
  float h = __fadd_rn(__fsub_rn(y2, y1), 1.0f);
  float s = __fsqrt_rn(__fmul_rn(w, h));
  float q = __fadd_rn(__fdiv_rn(s, 224.0f), 1e-6f);
  float l = floorf(__fadd_rn(4.0f, log2f(q)));
  if (l != l) return -1;
  l = fminf(fmaxf(l, k_min), k_max);
  return (int)l - (int)k_min;
}

struct RoiGeom {
  float start_w, start_h, bin_w, bin_h;
  int grid_w, grid_h;
};

// ROIAlign_cpu.cpp:146-169
__device__ __forceinline__ RoiGeom roi_geometry(float x1, float y1, float x2, float y2, float scale, int PH,
                                                int PW, int sampling_ratio) {
  RoiGeom g;
  g.start_w = __fmul_rn(x1, scale);
  g.start_h = __fmul_rn(y1, scale);
  float end_w = __fmul_rn(x2, scale), end_h = __fmul_rn(y2, scale);
  float rw = fmaxf(__fsub_rn(end_w, g.start_w), 1.0f);
  float rh = fmaxf(__fsub_rn(end_h, g.start_h), 1.0f);
  g.bin_h = __fdiv_rn(rh, (float)PH);
  g.bin_w = __fdiv_rn(rw, (float)PW);
  g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)PH));
  g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)PW));
  return g;
}

struct AxisTap {
  int lo, hi;  // tap indices along the axis (absolute, or patch-relative once rebased)
  float l, h;  // weight of the hi tap / lo tap; both 0 when the sample is out of range
};

// One axis of ROIAlign_cpu.cpp:36-92.  `ok` = coordinate inside [-1, extent].
__device__ __forceinline__ AxisTap axis_sample(float start, int p, float bin, int i, int grid, int extent,
                                               bool& ok) {
  float pos = __fadd_rn(start, __fmul_rn((float)p, bin));
  pos = __fadd_rn(pos, __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)grid));
  ok = !(pos < -1.0f || pos > (float)extent);
  if (pos <= 0.0f) pos = 0.0f;
  AxisTap t;
  t.lo = (int)pos;
  if (t.lo >= extent - 1) {
    t.hi = t.lo = extent - 1;
    pos = (float)t.lo;
  } else {
    t.hi = t.lo + 1;
  }
  t.l = __fsub_rn(pos, (float)t.lo);
  t.h = __fsub_rn(1.0f, t.l);
  return t;
}

}  // namespace b200
