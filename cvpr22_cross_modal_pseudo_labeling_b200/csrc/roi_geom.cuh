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


// ---------------------------------------------------------------------------------------
// Per-RoI axis tables of the marching kernels (sampling_ratio 2, NHWC).
// ---------------------------------------------------------------------------------------
struct AxisEntry {
  int lo, hi;  // element offsets: y*W*C for the row table, x*C (| column action) for the column table
  float l, h;
};

enum { kActReuse = 0, kActShift = 1, kActLoad2 = 2 };
constexpr int kMaxAxisSamples = 32;

// warp 0 fills the 2*PH row entries, warp 1 the 2*PW column entries (PH, PW <= 16); the caller
// synchronises the CTA afterwards.
__device__ __forceinline__ void build_axis_tables(const RoiGeom& g, int PH, int PW, int H, int W, int C, int warp,
                                                  int lane, AxisEntry* ytab, AxisEntry* xtab) {
  if (warp >= 2) return;
  const bool is_y = warp == 0;
  const int ns = 2 * (is_y ? PH : PW);
  bool ok = false;
  AxisTap t;
  t.lo = t.hi = 0;
  t.l = t.h = 0.f;
  if (lane < ns)
    t = is_y ? axis_sample(g.start_h, lane >> 1, g.bin_h, lane & 1, 2, H, ok)
             : axis_sample(g.start_w, lane >> 1, g.bin_w, lane & 1, 2, W, ok);
  ok = ok && lane < ns;
  AxisEntry e;
  const int stride = is_y ? W * C : C;
  // an out-of-range sample contributes nothing (ROIAlign_cpu.cpp:47-61): zero weights, taps
  // parked on element 0 of the axis
  e.lo = ok ? t.lo * stride : 0;
  e.hi = ok ? t.hi * stride : 0;
  e.l = ok ? t.l : 0.f;
  e.h = ok ? t.h : 0.f;
  // what the x-march must do to have (lo, hi) in its two register columns, given the previous
  // sample's columns.  C % 64 == 0 leaves the low bits of `lo` free for it.
  const int plo = __shfl_up_sync(0xffffffffu, e.lo, 1), phi = __shfl_up_sync(0xffffffffu, e.hi, 1);
  int act = kActLoad2;
  if (lane > 0) {
    if (e.lo == plo && e.hi == phi) act = kActReuse;
    else if (e.lo == phi) act = kActShift;
  }
  if (!is_y) e.lo |= act;
  if (lane < ns) (is_y ? ytab : xtab)[lane] = e;
}

// ---------------------------------------------------------------------------------------
// Tables of the separable marching kernels (roi_align_fwd_sep.cu, roi_align_bwd.cu).
// Rows: AxisEntry per y-sample (element offsets y*W*C).  Columns: the list of DISTINCT tap
// columns in the order a left-to-right march needs them (byte offsets x*C*4), and per x-sample
// the list index of its right tap column -- its left tap column is the previous list entry,
// or the same entry when the sample is clamped to the border (then l carries l + h and h = 0).
// ---------------------------------------------------------------------------------------
struct XSample {
  int jhi;     // index (in the column list) of the sample's right tap column
  float l, h;  // weights of the right / left tap
  int pad;
};
constexpr int kMaxCols = 2 * kMaxAxisSamples;

// warp 0 fills ytab[2*PH], warp 1 fills xs[2*PW] (+ a sentinel xs[2*PW].jhi = -1), colofs[] and
// *ncols; the caller synchronises the CTA afterwards.
__device__ __forceinline__ void build_sep_tables(const RoiGeom& g, int PH, int PW, int H, int W, int C, int warp,
                                                 int lane, AxisEntry* ytab, XSample* xs, int* colofs, int* ncols) {
  if (warp == 0) {
    bool ok = false;
    AxisTap t;
    t.lo = t.hi = 0;
    t.l = t.h = 0.f;
    if (lane < 2 * PH) t = axis_sample(g.start_h, lane >> 1, g.bin_h, lane & 1, 2, H, ok);
    ok = ok && lane < 2 * PH;
    AxisEntry e;
    e.lo = ok ? t.lo * W * C : 0;
    e.hi = ok ? t.hi * W * C : 0;
    e.l = ok ? t.l : 0.f;
    e.h = ok ? t.h : 0.f;
    if (lane < 2 * PH) ytab[lane] = e;
  } else if (warp == 1) {
    const int ns = 2 * PW;
    bool ok = false;
    AxisTap t;
    t.lo = t.hi = 0;
    t.l = t.h = 0.f;
    if (lane < ns) t = axis_sample(g.start_w, lane >> 1, g.bin_w, lane & 1, 2, W, ok);
    ok = ok && lane < ns;
    const int lo = ok ? t.lo * C : 0, hi = ok ? t.hi * C : 0;
    const int plo = __shfl_up_sync(0xffffffffu, lo, 1), phi = __shfl_up_sync(0xffffffffu, hi, 1);
    int act = kActLoad2;
    if (lane > 0) {
      if (lo == plo && hi == phi) act = kActReuse;
      else if (lo == phi) act = kActShift;
    }
    int nnew = act == kActReuse ? 0 : (act == kActShift ? 1 : (lo == hi ? 1 : 2));
    if (lane >= ns) nnew = 0;
    int scan = nnew;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, scan, d);
      if (lane >= d) scan += v;
    }
    if (lane < ns) {
      XSample e;
      e.jhi = scan - 1;
      const float wl = ok ? t.l : 0.f, wh = ok ? t.h : 0.f;
      e.l = lo == hi ? wl + wh : wl;
      e.h = lo == hi ? 0.f : wh;
      e.pad = 0;
      xs[lane] = e;
      if (nnew >= 1) colofs[scan - 1] = hi * 4;
      if (nnew == 2) colofs[scan - 2] = lo * 4;
    }
    if (lane == 31) {
      *ncols = scan;
      xs[ns].jhi = -1;  // sentinel: ends a consume loop after the last sample
    }
  }
}

// Tap rows of one output row (its two y-samples), duplicates merged: weights of equal rows add
// up and the later duplicate is dropped (use[k] = false, w[k] = 0).
__device__ __forceinline__ void merge_tap_rows(const AxisEntry& ya, const AxisEntry& yb, int (&row)[4], float (&w)[4],
                                               bool (&use)[4]) {
  row[0] = ya.lo; row[1] = ya.hi; row[2] = yb.lo; row[3] = yb.hi;
  w[0] = ya.h; w[1] = ya.l; w[2] = yb.h; w[3] = yb.l;
  use[0] = use[1] = use[2] = use[3] = true;
#pragma unroll
  for (int k = 1; k < 4; ++k) {
#pragma unroll
    for (int j = 0; j < k; ++j) {
      if (use[k] && row[k] == row[j]) {  // the earliest occurrence of a row is never merged away
        w[j] += w[k];
        w[k] = 0.f;
        use[k] = false;
      }
    }
  }
}

}  // namespace b200
