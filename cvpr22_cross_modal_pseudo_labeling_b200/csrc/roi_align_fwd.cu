// roi_align_fwd.cu -- fused multi-level RoIAlign forward for sm_100a.
//
// Replaces Pooler.forward (reference modeling/poolers.py:91-121): level assignment,
// per-level _C.roi_align_forward (csrc/cpu/ROIAlign_cpu.cpp:114-217,
// csrc/cuda/ROIAlign_cuda.cu:65-122) and the scatter back, in one launch.
//
// Three kernels here (the separable and the row-streaming fast-math kernels live in roi_align_fwd_sep.cu /
// roi_align_fwd_rows.cu):
//   roi_align_fwd_tile    NHWC features, ANY sampling ratio (the reference's shipped C4 pooler: 1024 channels,
//       14 x 14 bins, adaptive sampling): gather with per-CTA axis tables and a shared output tile, see below.
//   roi_align_fwd_march   NHWC features, sampling_ratio 2 (the FPN box / mask poolers).
//       One CTA per (RoI, 64-channel chunk).  Sample geometry is computed once per CTA
//       into two small axis tables.  A thread owns (4 channels, one output row of bins)
//       and marches along x keeping the two current tap columns of its four tap rows in
//       registers: neighbouring samples share columns, so it issues 2-4x fewer loads than
//       the 16 taps/bin of a naive gather, every load a 128-bit channel vector (16 lanes =
//       256 contiguous bytes of one pixel).  The patch is served by L1; HBM sees each
//       touched line about once per image because CTAs run in image order.  Results are
//       transposed through shared memory and leave as one contiguous [64 x PH*PW] block
//       of the NCHW output (streaming stores).
//   roi_align_fwd_generic any layout / sampling ratio: a thread per (channel[-quad], bin)
//       gathering straight from global memory with geometry evaluated on the fly.
//
// kExact=true reproduces the reference's arithmetic bit for bit (separately rounded
// multiplies and adds in its order).  kExact=false ("fast math") sends the marching shapes to
// the separable kernel of roi_align_fwd_sep.cu and lets the generic gather contract FMAs.
#include "roi_align_fwd.cuh"

namespace b200 {
namespace {


template <bool kExact>
__device__ __forceinline__ float tap4(float w1, float v1, float w2, float v2, float w3, float v3, float w4,
                                      float v4) {
  if (kExact) {
    // ROIAlign_cpu.cpp:201-204: ((w1*v1 + w2*v2) + w3*v3) + w4*v4
    float s = __fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2));
    s = __fadd_rn(s, __fmul_rn(w3, v3));
    return __fadd_rn(s, __fmul_rn(w4, v4));
  } else {
    return fmaf(w4, v4, fmaf(w3, v3, fmaf(w2, v2, w1 * v1)));
  }
}

template <bool kExact>
__device__ __forceinline__ float4 tap4v(float w1, float4 a, float w2, float4 b, float w3, float4 c, float w4,
                                        float4 d) {
  float4 r;
  r.x = tap4<kExact>(w1, a.x, w2, b.x, w3, c.x, w4, d.x);
  r.y = tap4<kExact>(w1, a.y, w2, b.y, w3, c.y, w4, d.y);
  r.z = tap4<kExact>(w1, a.z, w2, b.z, w3, c.z, w4, d.z);
  r.w = tap4<kExact>(w1, a.w, w2, b.w, w3, c.w, w4, d.w);
  return r;
}

__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}

// ---------------------------------------------------------------------------------------
// Generic gather.  NHWC: unit = (channel quad, bin); NCHW: unit = (channel, bin).
// ---------------------------------------------------------------------------------------
template <bool kExact, int kLayout>
__device__ void generic_roi(const float* __restrict__ feat, int C, int H, int W, int batch, const RoiGeom g,
                            int PH, int PW, int c_begin, int c_count, float* __restrict__ out_roi, int tid,
                            int nthreads) {
  const int NB = PH * PW;
  const float count = (float)(g.grid_h * g.grid_w);
  if (kLayout == B200_LAYOUT_NHWC) {
    const int nq = c_count >> 2;
    for (int u = tid; u < nq * NB; u += nthreads) {
      const int q = u % nq, bin = u / nq;
      const int ph = bin / PW, pw = bin - ph * PW;
      const float* base = feat + (size_t)batch * H * W * C + c_begin + 4 * q;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int iy = 0; iy < g.grid_h; ++iy) {
        bool oky;
        const AxisTap ty = axis_sample(g.start_h, ph, g.bin_h, iy, g.grid_h, H, oky);
        for (int ix = 0; ix < g.grid_w; ++ix) {
          bool okx;
          const AxisTap tx = axis_sample(g.start_w, pw, g.bin_w, ix, g.grid_w, W, okx);
          if (!(oky && okx)) continue;
          const float w1 = __fmul_rn(ty.h, tx.h), w2 = __fmul_rn(ty.h, tx.l);
          const float w3 = __fmul_rn(ty.l, tx.h), w4 = __fmul_rn(ty.l, tx.l);
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(base + ((size_t)ty.lo * W + tx.lo) * C));
          const float4 v2 = __ldg(reinterpret_cast<const float4*>(base + ((size_t)ty.lo * W + tx.hi) * C));
          const float4 v3 = __ldg(reinterpret_cast<const float4*>(base + ((size_t)ty.hi * W + tx.lo) * C));
          const float4 v4 = __ldg(reinterpret_cast<const float4*>(base + ((size_t)ty.hi * W + tx.hi) * C));
          acc = add4(acc, tap4v<kExact>(w1, v1, w2, v2, w3, v3, w4, v4));
        }
      }
      float* o = out_roi + (size_t)(c_begin + 4 * q) * NB + bin;
      o[0] = __fdiv_rn(acc.x, count);
      o[NB] = __fdiv_rn(acc.y, count);
      o[2 * NB] = __fdiv_rn(acc.z, count);
      o[3 * NB] = __fdiv_rn(acc.w, count);
    }
  } else {
    for (int u = tid; u < c_count * NB; u += nthreads) {
      const int c = c_begin + u / NB, bin = u % NB;
      const int ph = bin / PW, pw = bin - ph * PW;
      const float* plane = feat + ((size_t)batch * C + c) * H * W;
      float acc = 0.f;
      for (int iy = 0; iy < g.grid_h; ++iy) {
        bool oky;
        const AxisTap ty = axis_sample(g.start_h, ph, g.bin_h, iy, g.grid_h, H, oky);
        for (int ix = 0; ix < g.grid_w; ++ix) {
          bool okx;
          const AxisTap tx = axis_sample(g.start_w, pw, g.bin_w, ix, g.grid_w, W, okx);
          if (!(oky && okx)) continue;
          const float w1 = __fmul_rn(ty.h, tx.h), w2 = __fmul_rn(ty.h, tx.l);
          const float w3 = __fmul_rn(ty.l, tx.h), w4 = __fmul_rn(ty.l, tx.l);
          const float v1 = __ldg(plane + ty.lo * W + tx.lo), v2 = __ldg(plane + ty.lo * W + tx.hi);
          const float v3 = __ldg(plane + ty.hi * W + tx.lo), v4 = __ldg(plane + ty.hi * W + tx.hi);
          acc = __fadd_rn(acc, tap4<kExact>(w1, v1, w2, v2, w3, v3, w4, v4));
        }
      }
      out_roi[(size_t)c * NB + bin] = __fdiv_rn(acc, count);
    }
  }
}

// mean over the bins of each (RoI, channel) row of an already pooled block (generic path only:
// the marching kernels produce it from their shared-memory tile)
__global__ void __launch_bounds__(256) pooled_mean_kernel(const float* __restrict__ pooled, long long rows, int NB,
                                                          float* __restrict__ mean) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float* p = pooled + i * NB;
  float s = 0.f;
  for (int k = 0; k < NB; ++k) s = __fadd_rn(s, p[k]);
  mean[i] = __fdiv_rn(s, (float)NB);
}

// grid.x = n_rois * ceil(C / c_per_cta)
template <bool kExact, int kLayout>
__global__ void __launch_bounds__(256)
roi_align_fwd_generic(const LevelTable lt, int C, const float* __restrict__ rois, int PH, int PW,
                      int sampling_ratio, int c_per_cta, int n_cchunks, float* __restrict__ out,
                      int32_t* __restrict__ out_levels) {
  const long long r = blockIdx.x / n_cchunks;
  const int ck = blockIdx.x % n_cchunks;
  const RoiHeader h = load_roi(rois, r, lt);
  const int NB = PH * PW;
  const int c_begin = ck * c_per_cta;
  const int c_count = min(c_per_cta, C - c_begin);
  float* out_roi = out + (size_t)r * C * NB;
  if (ck == 0 && threadIdx.x == 0 && out_levels) out_levels[r] = h.level;
  if (h.level < 0) {
    zero_fill(out_roi + (size_t)c_begin * NB, c_count * NB, threadIdx.x, blockDim.x);
    return;
  }
  const RoiGeom g = roi_geometry(h.x1, h.y1, h.x2, h.y2, lt.scale[h.level], PH, PW, sampling_ratio);
  generic_roi<kExact, kLayout>(lt.data[h.level], C, lt.H[h.level], lt.W[h.level], h.batch, g, PH, PW, c_begin,
                               c_count, out_roi, threadIdx.x, blockDim.x);
}

// ---------------------------------------------------------------------------------------
// Tiled gather (NHWC, ANY sampling ratio incl. the adaptive sampling_ratio 0, C % 64 == 0, a [64 x NB] tile that
// fits in shared memory): the reference's shipped C4 pooler -- 1024 channels, 14 x 14 bins, sampling_ratio 0,
// config/defaults.py:301-305 -- where the output (803 KB per RoI) is almost all of the traffic.
// One CTA per (RoI, 64-channel chunk).  The sample geometry of both axes is evaluated once into two
// shared-memory tables (the generic gather re-evaluates it per thread and tap); a thread = (channel quad, bin),
// 16 quads side by side = 256 contiguous bytes of a pixel per tap; the arithmetic and its order are the generic
// gather's (bit-identical to ROIAlign_cpu when kExact); results leave through the shared [64 x NB] tile as one
// contiguous block of the NCHW output (128-bit streaming stores) -- the generic gather stores 4 bytes per lane at
// a stride of 4 * NB floats, a sector per element.  A RoI whose sample count exceeds the tables takes the
// generic path inside the same launch.
// ---------------------------------------------------------------------------------------
constexpr int kTileThreads = 256;
constexpr int kTileMaxAx = 128;  // samples per axis held in the tables (PH * grid_h, PW * grid_w)

struct __align__(16) TileTap {
  uint32_t lo, hi;  // BYTE offsets of the two taps inside the image's map: y * W * C * 4 (rows), x * C * 4 (columns)
  float l, h;       // their weights; both 0: the sample is out of range and contributes nothing
};

// one sample of a bin: four 128-bit taps, the reference's order ((w1 v1 + w2 v2) + w3 v3) + w4 v4 added to the bin's
// accumulator.  Two fp32 lanes per issue slot where possible (FMUL2 / FADD2 / FFMA2 round each lane like the scalar
// instruction); one 64-bit base pointer + 32-bit byte offsets (the compiler spent three integer instructions per load
// on 64-bit address arithmetic with element offsets).
template <bool kExact>
__device__ __forceinline__ void tile_sample(const char* __restrict__ base, const TileTap& ty, const TileTap& tx,
                                            float2& acc_lo, float2& acc_hi) {
  const float4 v1 = __ldg(reinterpret_cast<const float4*>(base + (ty.lo + tx.lo)));
  const float4 v2 = __ldg(reinterpret_cast<const float4*>(base + (ty.lo + tx.hi)));
  const float4 v3 = __ldg(reinterpret_cast<const float4*>(base + (ty.hi + tx.lo)));
  const float4 v4 = __ldg(reinterpret_cast<const float4*>(base + (ty.hi + tx.hi)));
  const float2 txhl = make_float2(tx.h, tx.l);
  const float2 w12 = __fmul2_rn(make_float2(ty.h, ty.h), txhl), w34 = __fmul2_rn(make_float2(ty.l, ty.l), txhl);
  const float2 w1 = make_float2(w12.x, w12.x), w2 = make_float2(w12.y, w12.y), w3 = make_float2(w34.x, w34.x),
               w4 = make_float2(w34.y, w34.y);
  if (kExact) {
    // products as FMUL2, the three sums as scalar FADDs: ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into
    // FFMA2 although both carry an explicit rounding mode (it never does for the scalar forms), which would
    // break the bit-exactness; the accumulation below has no product operand and stays a FADD2
    const float2 p1l = __fmul2_rn(w1, make_float2(v1.x, v1.y)), p1h = __fmul2_rn(w1, make_float2(v1.z, v1.w));
    const float2 p2l = __fmul2_rn(w2, make_float2(v2.x, v2.y)), p2h = __fmul2_rn(w2, make_float2(v2.z, v2.w));
    const float2 p3l = __fmul2_rn(w3, make_float2(v3.x, v3.y)), p3h = __fmul2_rn(w3, make_float2(v3.z, v3.w));
    const float2 p4l = __fmul2_rn(w4, make_float2(v4.x, v4.y)), p4h = __fmul2_rn(w4, make_float2(v4.z, v4.w));
    float2 s_lo, s_hi;
    s_lo.x = __fadd_rn(__fadd_rn(__fadd_rn(p1l.x, p2l.x), p3l.x), p4l.x);
    s_lo.y = __fadd_rn(__fadd_rn(__fadd_rn(p1l.y, p2l.y), p3l.y), p4l.y);
    s_hi.x = __fadd_rn(__fadd_rn(__fadd_rn(p1h.x, p2h.x), p3h.x), p4h.x);
    s_hi.y = __fadd_rn(__fadd_rn(__fadd_rn(p1h.y, p2h.y), p3h.y), p4h.y);
    acc_lo = __fadd2_rn(acc_lo, s_lo);
    acc_hi = __fadd2_rn(acc_hi, s_hi);
  } else {
    acc_lo = __ffma2_rn(w1, make_float2(v1.x, v1.y), acc_lo);
    acc_hi = __ffma2_rn(w1, make_float2(v1.z, v1.w), acc_hi);
    acc_lo = __ffma2_rn(w2, make_float2(v2.x, v2.y), acc_lo);
    acc_hi = __ffma2_rn(w2, make_float2(v2.z, v2.w), acc_hi);
    acc_lo = __ffma2_rn(w3, make_float2(v3.x, v3.y), acc_lo);
    acc_hi = __ffma2_rn(w3, make_float2(v3.z, v3.w), acc_hi);
    acc_lo = __ffma2_rn(w4, make_float2(v4.x, v4.y), acc_lo);
    acc_hi = __ffma2_rn(w4, make_float2(v4.z, v4.w), acc_hi);
  }
}
__device__ __forceinline__ bool tile_tap_valid(const TileTap& t) { return !(t.h == 0.f && t.l == 0.f); }  // (h = 1 - l > 0 or l > 0)

// the bins of one thread.  GH / GW > 0: samples per bin and axis known at compile time (1 x 1 for two RoIs in three of
// the C4 pooler, 2 x 2 for most of the rest): no sample loops; 0: taken from the geometry.
template <bool kExact, int GH, int GW, int kSlots>
__device__ __forceinline__ void tile_bins(const char* __restrict__ base, const TileTap* __restrict__ ytab,
                                          const TileTap* __restrict__ xtab, int grid_h, int grid_w, int PW, int NB, int slot,
                                          float* __restrict__ tile, int row0, int row1, int row2, int row3) {
  const int gh = GH > 0 ? GH : grid_h, gw = GW > 0 ? GW : grid_w;
  const int n_samples = gh * gw;
  const float count = (float)n_samples;
  // acc / count: a power-of-two count (1, 2, 4, 8, 16: nine RoIs in ten of the C4 pooler) is an exact scaling, so the
  // product with 1 / count is the correctly rounded quotient as well; other counts take the IEEE division
  const bool pow2 = (n_samples & (n_samples - 1)) == 0;
  const float inv_count = __fdiv_rn(1.0f, count);
  // (ph, pw) of the thread's bins without a division per bin
  const int step_ph = kSlots / PW, step_pw = kSlots - step_ph * PW;
  int ph = slot / PW, pw = slot - ph * PW;
  for (int bin = slot; bin < NB; bin += kSlots) {
    const TileTap* yt = ytab + ph * gh;
    const TileTap* xt = xtab + pw * gw;
    float2 acc_lo = make_float2(0.f, 0.f), acc_hi = make_float2(0.f, 0.f);
    if (GH > 0 && GW > 0) {
      TileTap ty[GH > 0 ? GH : 1], tx[GW > 0 ? GW : 1];
#pragma unroll
      for (int i = 0; i < GH; ++i) ty[i] = yt[i];
#pragma unroll
      for (int i = 0; i < GW; ++i) tx[i] = xt[i];
#pragma unroll
      for (int iy = 0; iy < GH; ++iy)
#pragma unroll
        for (int ix = 0; ix < GW; ++ix)
          if (tile_tap_valid(ty[iy]) && tile_tap_valid(tx[ix])) tile_sample<kExact>(base, ty[iy], tx[ix], acc_lo, acc_hi);
    } else {
      for (int iy = 0; iy < gh; ++iy) {
        const TileTap ty = yt[iy];
        if (!tile_tap_valid(ty)) continue;
        for (int ix = 0; ix < gw; ++ix) {
          const TileTap tx = xt[ix];
          if (!tile_tap_valid(tx)) continue;
          tile_sample<kExact>(base, ty, tx, acc_lo, acc_hi);
        }
      }
    }
    if (GH == 1 && GW == 1) {
      // (count 1: x / 1 = x)
    } else if (pow2 || !kExact) {
      const float2 ic = make_float2(inv_count, inv_count);
      acc_lo = __fmul2_rn(acc_lo, ic);
      acc_hi = __fmul2_rn(acc_hi, ic);
    } else {
      acc_lo = make_float2(__fdiv_rn(acc_lo.x, count), __fdiv_rn(acc_lo.y, count));
      acc_hi = make_float2(__fdiv_rn(acc_hi.x, count), __fdiv_rn(acc_hi.y, count));
    }
    tile[row0 + bin] = acc_lo.x;
    tile[row1 + bin] = acc_lo.y;
    tile[row2 + bin] = acc_hi.x;
    tile[row3 + bin] = acc_hi.y;
    ph += step_ph;
    pw += step_pw;
    if (pw >= PW) {
      pw -= PW;
      ++ph;
    }
  }
}

// row of channel c in the [kCh x NB] tile: shifted by 4 floats per 8 channels when NB % 4 == 0 (as tile_row)
template <int kCh>
constexpr size_t tile_kernel_smem(int NB) {
  return sizeof(float) * ((size_t)kCh * NB + kTilePadFloats) + 2 * kTileMaxAx * sizeof(TileTap);
}

// kCh: channels per CTA.  32 (a tile of 25 KB at 14 x 14: four CTAs leave ~130 KB of the SM's L1 for the taps --
// with 64-channel tiles the shared-memory carve-out left ~30 KB and every tap went to the L2) or 64.
template <bool kExact, int kCh>
__global__ void __launch_bounds__(kTileThreads)
roi_align_fwd_tile(const LevelTable lt, int C, const float* __restrict__ rois, int PH, int PW, int sampling_ratio,
                   float* __restrict__ out, float* __restrict__ out_mean, int32_t* __restrict__ out_levels) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  constexpr int kQuads = kCh / 4, kSlots = kTileThreads / kQuads;
  const int NB = PH * PW;
  const bool swz = (NB & 3) == 0;
  float* tile = reinterpret_cast<float*>(tile_smem);
  TileTap* ytab = reinterpret_cast<TileTap*>(tile_smem + sizeof(float) * (kCh * NB + kTilePadFloats));
  TileTap* xtab = ytab + kTileMaxAx;
  const int n_cchunks = C / kCh;
  const long long r = blockIdx.x / n_cchunks;
  const int c_begin = (blockIdx.x % n_cchunks) * kCh;
  const int tid = threadIdx.x;
  const RoiHeader h = load_roi(rois, r, lt);
  float* out_roi = out + (size_t)r * C * NB;
  if (c_begin == 0 && tid == 0 && out_levels) out_levels[r] = h.level;
  if (h.level < 0) {
    zero_fill(out_roi + (size_t)c_begin * NB, kCh * NB, tid, kTileThreads);
    if (out_mean) zero_fill(out_mean + (size_t)r * C + c_begin, kCh, tid, kTileThreads);
    return;
  }
  const int H = lt.H[h.level], W = lt.W[h.level];
  const RoiGeom g = roi_geometry(h.x1, h.y1, h.x2, h.y2, lt.scale[h.level], PH, PW, sampling_ratio);
  if (PH * g.grid_h > kTileMaxAx || PW * g.grid_w > kTileMaxAx) {  // (a huge RoI on a fine map)
    generic_roi<kExact, B200_LAYOUT_NHWC>(lt.data[h.level], C, H, W, h.batch, g, PH, PW, c_begin, kCh, out_roi, tid,
                                          kTileThreads);
    if (out_mean) {
      __syncthreads();  // (this CTA's own global writes)
      if (tid < kCh) {
        const float* p = out_roi + (size_t)(c_begin + tid) * NB;
        float s = 0.f;
        for (int k = 0; k < NB; ++k) s = __fadd_rn(s, p[k]);
        out_mean[(size_t)r * C + c_begin + tid] = __fdiv_rn(s, (float)NB);
      }
    }
    return;
  }
  for (int i = tid; i < PH * g.grid_h + PW * g.grid_w; i += kTileThreads) {
    const bool is_y = i < PH * g.grid_h;
    const int j = is_y ? i : i - PH * g.grid_h;
    const int grid = is_y ? g.grid_h : g.grid_w;
    const int p = j / grid, k = j - p * grid;
    bool ok;
    const AxisTap t = is_y ? axis_sample(g.start_h, p, g.bin_h, k, grid, H, ok) : axis_sample(g.start_w, p, g.bin_w, k, grid, W, ok);
    const uint32_t stride = 4u * (uint32_t)(is_y ? W * C : C);  // (the launcher checks H * W * C * 4 < 2^32)
    TileTap e;
    e.lo = ok ? (uint32_t)t.lo * stride : 0u;
    e.hi = ok ? (uint32_t)t.hi * stride : 0u;
    e.l = ok ? t.l : 0.f;
    e.h = ok ? t.h : 0.f;
    (is_y ? ytab : xtab)[j] = e;
  }
  __syncthreads();
  const int q = tid % kQuads, slot = tid / kQuads;
  const char* base = reinterpret_cast<const char*>(lt.data[h.level] + (size_t)h.batch * H * W * C + c_begin + 4 * q);
  const int row0 = tile_row(4 * q, NB, swz), row1 = tile_row(4 * q + 1, NB, swz), row2 = tile_row(4 * q + 2, NB, swz),
            row3 = tile_row(4 * q + 3, NB, swz);
  if (g.grid_h == 1 && g.grid_w == 1)
    tile_bins<kExact, 1, 1, kSlots>(base, ytab, xtab, 1, 1, PW, NB, slot, tile, row0, row1, row2, row3);
  else if (g.grid_h == 2 && g.grid_w == 2)
    tile_bins<kExact, 2, 2, kSlots>(base, ytab, xtab, 2, 2, PW, NB, slot, tile, row0, row1, row2, row3);
  else
    tile_bins<kExact, 0, 0, kSlots>(base, ytab, xtab, g.grid_h, g.grid_w, PW, NB, slot, tile, row0, row1, row2, row3);
  __syncthreads();
  // the [kCh x NB] block is contiguous in the NCHW output: 128-bit streaming stores
  {
    float4* dst = reinterpret_cast<float4*>(out_roi + (size_t)c_begin * NB);
    const float4* src = reinterpret_cast<const float4*>(tile);
    if (!swz) {
      for (int i = tid; i < kCh * NB / 4; i += kTileThreads) __stcs(dst + i, src[i]);
    } else {
      const int rowv = NB >> 2;
      for (int i = tid; i < kCh * rowv; i += kTileThreads) __stcs(dst + i, src[i + ((i / rowv) >> 3)]);
    }
  }
  if (out_mean && tid < kCh) {
    const float* row = tile + tile_row(tid, NB, swz);
    float s = 0.f;
    for (int i = 0; i < NB; ++i) s = __fadd_rn(s, row[i]);
    out_mean[(size_t)r * C + c_begin + tid] = __fdiv_rn(s, (float)NB);
  }
}

// ---------------------------------------------------------------------------------------
// Marching kernel (NHWC, sampling_ratio == 2, PH,PW <= 16, C % 64 == 0).
// One CTA per (RoI, 64-channel chunk); shared memory holds only the two axis tables and
// the [64 x NB] output tile, so many CTAs share an SM and the RoI's patch lives in L1.
// ---------------------------------------------------------------------------------------
// V channels per lane (4 -> 128-bit loads, 2 -> 64-bit loads)
template <int V>
struct Vec {
  float v[V];
};

template <int V>
__device__ __forceinline__ Vec<V> ldg_vec(const float* p) {
  Vec<V> r;
  if (V == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[V > 2 ? 2 : 0] = t.z; r.v[V > 3 ? 3 : 0] = t.w;
  } else {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r.v[0] = t.x; r.v[1] = t.y;
  }
  return r;
}

// rows[] / off are element offsets from `feat` (uniform across the CTA, < 2^31 per image)
template <int V>
__device__ __forceinline__ void load_col(Vec<V> (&dst)[4], const float* __restrict__ feat, const int (&rows)[4],
                                         int off) {
#pragma unroll
  for (int k = 0; k < 4; ++k) dst[k] = ldg_vec<V>(feat + (unsigned)(rows[k] + off));
}

// the two samples (iy = 0, 1) of one sample column: left taps in Lc, right taps in Rc
template <bool kExact, int V>
__device__ __forceinline__ void sample_pair(const Vec<V> (&Lc)[4], const Vec<V> (&Rc)[4], const AxisEntry& ya,
                                            const AxisEntry& yb, const AxisEntry& xt, Vec<V>& sa, Vec<V>& sb) {
  const float a1 = __fmul_rn(ya.h, xt.h), a2 = __fmul_rn(ya.h, xt.l), a3 = __fmul_rn(ya.l, xt.h),
              a4 = __fmul_rn(ya.l, xt.l);
  const float b1 = __fmul_rn(yb.h, xt.h), b2 = __fmul_rn(yb.h, xt.l), b3 = __fmul_rn(yb.l, xt.h),
              b4 = __fmul_rn(yb.l, xt.l);
#pragma unroll
  for (int c = 0; c < V; ++c) {
    sa.v[c] = tap4<kExact>(a1, Lc[0].v[c], a2, Rc[0].v[c], a3, Lc[1].v[c], a4, Rc[1].v[c]);
    sb.v[c] = tap4<kExact>(b1, Lc[2].v[c], b2, Rc[2].v[c], b3, Lc[3].v[c], b4, Rc[3].v[c]);
  }
}

template <bool kExact, int kThreads, int kMinBlocks, int V>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
roi_align_fwd_march(const LevelTable lt, int C, const float* __restrict__ rois, int PH, int PW,
                    float* __restrict__ out, float* __restrict__ out_mean, int32_t* __restrict__ out_levels) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NB = PH * PW;
  const bool swz = (NB & 3) == 0;
  float* out_s = reinterpret_cast<float*>(smem_raw);
  AxisEntry* ytab = reinterpret_cast<AxisEntry*>(smem_raw + sizeof(float) * (kChunk * NB + kTilePadFloats));
  AxisEntry* xtab = ytab + kMaxAxisSamples;

  const int tid = threadIdx.x;
  const int n_cchunks = C / kChunk;
  const long long r = blockIdx.x / n_cchunks;
  const int ck = blockIdx.x % n_cchunks;
  const int c_begin = ck * kChunk;
  const RoiHeader h = load_roi(rois, r, lt);
  float* out_roi = out + (size_t)r * C * NB;
  if (ck == 0 && tid == 0 && out_levels) out_levels[r] = h.level;
  if (h.level < 0) {
    zero_fill(out_roi + (size_t)c_begin * NB, kChunk * NB, tid, kThreads);
    if (out_mean) zero_fill(out_mean + (size_t)r * C + c_begin, kChunk, tid, kThreads);
    return;
  }
  const int H = lt.H[h.level], W = lt.W[h.level];
  const RoiGeom g = roi_geometry(h.x1, h.y1, h.x2, h.y2, lt.scale[h.level], PH, PW, 2);

  const int warp = tid >> 5, lane = tid & 31;
  build_axis_tables(g, PH, PW, H, W, C, warp, lane, ytab, xtab);
  __syncthreads();

  constexpr int kGroups = kChunk / V;  // lane groups (V channels each) across the chunk
  const float* feat = lt.data[h.level] + (size_t)h.batch * H * W * C + c_begin;
  // ---- march: thread = (output row ph, channel group q) ---------------------------------
  // Lane layout.  The L1 data pipe spends one wavefront per 128-byte line a request touches
  // (ncu: l1tex__data_pipe_lsu_wavefronts is this kernel's busiest unit), so the lanes of a
  // warp that read one pixel cover at least 128 contiguous bytes of it:
  // PH <= 8 (128 threads): 8 channel groups x 4 output rows per warp;
  // PH <= 16 (256 threads): 16 groups x 2 output rows per warp.
  static_assert(V == 4, "lane layout assumes 128-bit loads");
  constexpr int kWarps = kThreads / 32;
  constexpr int kLQ = kWarps <= 4 ? 8 : 16;  // groups side by side in a warp
  constexpr int kQGroups = kGroups / kLQ;              // warps needed to cover all groups
  {
    const int q = (warp % kQGroups) * kLQ + (lane % kLQ);
    const int ph = lane / kLQ + (32 / kLQ) * (warp / kQGroups);
    if (ph < PH) {
      const AxisEntry ya = ytab[2 * ph], yb = ytab[2 * ph + 1];
      const int rows[4] = {ya.lo + V * q, ya.hi + V * q, yb.lo + V * q, yb.hi + V * q};
      // Two register columns (4 tap rows x V channels each) ping-pong between the roles
      // "left tap" and "right tap": when the next sample's left column is the current right
      // one, the roles swap and only one new column is loaded -- no register moves.
      Vec<V> A[4], B[4];
      int a_is_left = 1;
      float* o = out_s + tile_row(V * q, NB, swz) + ph * PW;  // the group's rows share one shift
      for (int pw = 0; pw < PW; ++pw) {
        Vec<V> sa[2], sb[2];
#pragma unroll
        for (int ix = 0; ix < 2; ++ix) {
          const AxisEntry xt = xtab[2 * pw + ix];
          const int act = xt.lo & 3, xlo = xt.lo & ~3;
          if (act == kActShift) {
            a_is_left ^= 1;
            if (a_is_left) load_col<V>(B, feat, rows, xt.hi);
            else load_col<V>(A, feat, rows, xt.hi);
          } else if (act == kActLoad2) {
            if (a_is_left) {
              load_col<V>(A, feat, rows, xlo);
              load_col<V>(B, feat, rows, xt.hi);
            } else {
              load_col<V>(B, feat, rows, xlo);
              load_col<V>(A, feat, rows, xt.hi);
            }
          }
          if (a_is_left) sample_pair<kExact, V>(A, B, ya, yb, xt, sa[ix], sb[ix]);
          else sample_pair<kExact, V>(B, A, ya, yb, xt, sa[ix], sb[ix]);
        }
        // (iy, ix) accumulation order of the reference, then / count (count = 4, exact)
#pragma unroll
        for (int c = 0; c < V; ++c) {
          const float acc = __fadd_rn(__fadd_rn(__fadd_rn(sa[0].v[c], sa[1].v[c]), sb[0].v[c]), sb[1].v[c]);
          o[c * NB + pw] = acc * 0.25f;
        }
      }
    }
  }
  __syncthreads();
  // ---- contiguous [64 x NB] block of the NCHW output ------------------------------------
  tile_copy_out(out_s, out_roi + (size_t)c_begin * NB, NB, swz, tid, kThreads);
  if (out_mean) tile_mean_out(out_s, out_mean + (size_t)r * C + c_begin, NB, swz, tid);
}

size_t march_smem_bytes(int NB) {
  return sizeof(float) * (kChunk * NB + kTilePadFloats) + 2 * kMaxAxisSamples * sizeof(AxisEntry);
}

bool g_force_generic = false;
bool g_exact = true;
int g_variant = 0;  // tuning hook: occupancy variant of the marching kernel

template <bool kExact, int kThreads, int kMinBlocks, int V>
int launch_march(const LevelTable& lt, int C, const float* rois, int64_t n_rois, int PH, int PW, float* out,
                 float* out_mean, int32_t* out_levels, cudaStream_t st) {
  const size_t smem = march_smem_bytes(PH * PW);
  auto kern = roi_align_fwd_march<kExact, kThreads, kMinBlocks, V>;
  static SmemHighWater hw;  // one per template instantiation
  int rc = ensure_dynamic_smem(kern, smem, &hw, "roi_align: smem attribute");
  if (rc != B200_OK) return rc;
  const int64_t grid = n_rois * (C / kChunk);
  kern<<<(unsigned)grid, kThreads, smem, st>>>(lt, C, rois, PH, PW, out, out_mean, out_levels);
  B200_CHECK_LAUNCH("roi_align_fwd_march");
  return B200_OK;
}

template <bool kExact>
int launch_forward(const LevelTable& lt, int layout, int C, const float* rois, int64_t n_rois, int PH, int PW,
                   int sr, float* out, float* out_mean, int32_t* out_levels, int32_t* order_ws, cudaStream_t st) {
  const int NB = PH * PW;
  const bool march_ok = !g_force_generic && layout == B200_LAYOUT_NHWC && sr == 2 && PH <= 16 && PW <= 16 &&
                        C % kChunk == 0 && (NB * kChunk) % 4 == 0;
  B200_REQUIRE(n_rois * ((C + 15) / 16) < (int64_t)1 << 31, "roi_align: too many RoIs for one launch");
  if (march_ok) {
    if constexpr (!kExact) {
      // fast math: the row-streaming kernel for the FPN box pooler shape (roi_align_fwd_rows.cu), else
      // -- or when the tuning hook sets g_variant & 16 -- the separable marching kernel
      // (roi_align_fwd_sep.cu)
      if (!(g_variant & 16) && rows_kernel_applies(lt, C, PH, PW))
        return launch_forward_rows(lt, C, false, rois, n_rois, out, out_mean, out_levels, order_ws, g_variant, st);
      return launch_forward_sep(lt, C, rois, n_rois, PH, PW, out, out_mean, out_levels, g_variant & 15, st);
    } else {
      // g_variant (tuning hook): CTAs/SM = 6 (default) / 5 / 4 for 128 threads, 3 / 3 / 2 for 256.
      // (2 channels per lane with twice the warps was measured 35-55 % slower: V stays 4.)
      if (PH * (kChunk / 4) <= 128) {
        if (g_variant == 1) return launch_march<true, 128, 5, 4>(lt, C, rois, n_rois, PH, PW, out, out_mean, out_levels, st);
        if (g_variant == 2) return launch_march<true, 128, 4, 4>(lt, C, rois, n_rois, PH, PW, out, out_mean, out_levels, st);
        return launch_march<true, 128, 6, 4>(lt, C, rois, n_rois, PH, PW, out, out_mean, out_levels, st);
      }
      if (g_variant == 2) return launch_march<true, 256, 2, 4>(lt, C, rois, n_rois, PH, PW, out, out_mean, out_levels, st);
      return launch_march<true, 256, 3, 4>(lt, C, rois, n_rois, PH, PW, out, out_mean, out_levels, st);
    }
  }
  // NHWC, any sampling ratio: the tiled gather when the [32 x NB] tile fits (tuning: g_variant & 16384 = the plain
  // gather, & 8192 = 64 channels per CTA)
  bool maps_fit_u32 = true;  // the tap tables hold 32-bit byte offsets inside one image's map
  for (int l = 0; l < lt.n_levels; ++l) maps_fit_u32 = maps_fit_u32 && (uint64_t)lt.H[l] * lt.W[l] * C * 4 < ((uint64_t)1 << 32);
  if (!g_force_generic && !(g_variant & 16384) && layout == B200_LAYOUT_NHWC && C % kChunk == 0 && (NB * 32) % 4 == 0 &&
      tile_kernel_smem<kChunk>(NB) <= (size_t)100 * 1024 && maps_fit_u32) {
    if (g_variant & 8192) {
      auto kern = roi_align_fwd_tile<kExact, 64>;
      static SmemHighWater hw;
      int rc = ensure_dynamic_smem(kern, tile_kernel_smem<64>(NB), &hw, "roi_align tile: smem attribute");
      if (rc != B200_OK) return rc;
      kern<<<(unsigned)(n_rois * (C / 64)), kTileThreads, tile_kernel_smem<64>(NB), st>>>(lt, C, rois, PH, PW, sr, out, out_mean,
                                                                                       out_levels);
    } else {
      auto kern = roi_align_fwd_tile<kExact, 32>;
      static SmemHighWater hw;
      int rc = ensure_dynamic_smem(kern, tile_kernel_smem<32>(NB), &hw, "roi_align tile: smem attribute");
      if (rc != B200_OK) return rc;
      kern<<<(unsigned)(n_rois * (C / 32)), kTileThreads, tile_kernel_smem<32>(NB), st>>>(lt, C, rois, PH, PW, sr, out, out_mean,
                                                                                       out_levels);
    }
    B200_CHECK_LAUNCH("roi_align_fwd_tile");
    return B200_OK;
  }
  // generic: pick channels per CTA so that a CTA has >= ~2k units of work
  int c_per_cta = C;
  if (layout == B200_LAYOUT_NHWC) {
    B200_REQUIRE(C % 4 == 0, "roi_align: NHWC needs channels %% 4 == 0");
    c_per_cta = 64;
    if (c_per_cta > C) c_per_cta = C;
  } else {
    c_per_cta = (2048 + NB - 1) / NB;
    if (c_per_cta > C) c_per_cta = C;
  }
  const int n_cchunks = (C + c_per_cta - 1) / c_per_cta;
  const int64_t grid = n_rois * n_cchunks;
  if (layout == B200_LAYOUT_NHWC)
    roi_align_fwd_generic<kExact, B200_LAYOUT_NHWC>
        <<<(unsigned)grid, 256, 0, st>>>(lt, C, rois, PH, PW, sr, c_per_cta, n_cchunks, out, out_levels);
  else
    roi_align_fwd_generic<kExact, B200_LAYOUT_NCHW>
        <<<(unsigned)grid, 256, 0, st>>>(lt, C, rois, PH, PW, sr, c_per_cta, n_cchunks, out, out_levels);
  B200_CHECK_LAUNCH("roi_align_fwd_generic");
  if (out_mean) {
    const int64_t rows = n_rois * C;
    pooled_mean_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(out, rows, NB, out_mean);
    B200_CHECK_LAUNCH("pooled_mean_kernel");
  }
  return B200_OK;
}

}  // namespace

int fill_level_table(const b200_level* levels, int n_levels, LevelTable* lt) {
  B200_REQUIRE(levels && n_levels >= 1 && n_levels <= B200_MAX_LEVELS, "roi_align: n_levels must be 1..%d",
               B200_MAX_LEVELS);
  for (int l = 0; l < n_levels; ++l) {
    B200_REQUIRE(levels[l].data && aligned16(levels[l].data), "roi_align: level %d data null or misaligned", l);
    B200_REQUIRE(levels[l].height > 0 && levels[l].width > 0 && levels[l].spatial_scale > 0.f,
                 "roi_align: level %d has an empty shape or non-positive scale", l);
    lt->data[l] = levels[l].data;
    lt->H[l] = levels[l].height;
    lt->W[l] = levels[l].width;
    lt->scale[l] = levels[l].spatial_scale;
  }
  lt->n_levels = n_levels;
  lt->k_min = -log2f(levels[0].spatial_scale);
  lt->k_max = -log2f(levels[n_levels - 1].spatial_scale);
  return B200_OK;
}

}  // namespace b200

// Tuning / test hooks (not part of the reference-facing ABI; declared here only).
extern "C" void b200_debug_set(int force_generic, int exact, int variant) {
  b200::g_force_generic = force_generic != 0;
  b200::g_exact = exact != 0;
  b200::g_variant = variant;
}

namespace b200 {
namespace {
int roi_align_forward_impl(bool exact, const b200_level* levels, int n_levels, int layout, int batch, int channels,
                           const float* rois, int64_t n_rois, int pooled_h, int pooled_w, int sampling_ratio,
                           float* out, float* out_mean, int32_t* out_levels, void* workspace, size_t workspace_bytes,
                           void* stream) {
  B200_REQUIRE(layout == B200_LAYOUT_NCHW || layout == B200_LAYOUT_NHWC, "roi_align: bad layout %d", layout);
  B200_REQUIRE(batch > 0 && channels > 0 && pooled_h > 0 && pooled_w > 0 && n_rois >= 0 && sampling_ratio >= 0,
               "roi_align: bad shape");
  if (n_rois == 0) return B200_OK;
  B200_REQUIRE(rois && out, "roi_align: null rois / out");
  B200_REQUIRE(aligned16(out), "roi_align: out must be 16-byte aligned");
  LevelTable lt;
  int rc = fill_level_table(levels, n_levels, &lt);
  if (rc != B200_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // optional scratch: the RoI visiting order of the row-streaming kernel (ignored when too small)
  int32_t* order_ws = workspace && workspace_bytes >= rows_order_workspace_bytes(n_rois) &&
                              (reinterpret_cast<uintptr_t>(workspace) & 3u) == 0
                          ? static_cast<int32_t*>(workspace)
                          : nullptr;
  return exact ? launch_forward<true>(lt, layout, channels, rois, n_rois, pooled_h, pooled_w, sampling_ratio, out,
                                      out_mean, out_levels, order_ws, st)
               : launch_forward<false>(lt, layout, channels, rois, n_rois, pooled_h, pooled_w, sampling_ratio,
                                       out, out_mean, out_levels, order_ws, st);
}
}  // namespace
}  // namespace b200

extern "C" int b200_roi_align_forward(const b200_level* levels, int n_levels, int layout, int batch,
                                      int channels, const float* rois, int64_t n_rois, int pooled_h,
                                      int pooled_w, int sampling_ratio, float* out, int32_t* out_levels,
                                      void* stream) {
  // g_exact is the tuning hook of b200_debug_set; it is true unless a perf script flipped it
  return b200::roi_align_forward_impl(b200::g_exact, levels, n_levels, layout, batch, channels, rois, n_rois,
                                      pooled_h, pooled_w, sampling_ratio, out, nullptr, out_levels, nullptr, 0, stream);
}

extern "C" int b200_roi_align_forward_fast(const b200_level* levels, int n_levels, int layout, int batch,
                                           int channels, const float* rois, int64_t n_rois, int pooled_h,
                                           int pooled_w, int sampling_ratio, float* out, int32_t* out_levels,
                                           void* stream) {
  return b200::roi_align_forward_impl(false, levels, n_levels, layout, batch, channels, rois, n_rois, pooled_h,
                                      pooled_w, sampling_ratio, out, nullptr, out_levels, nullptr, 0, stream);
}

extern "C" int b200_roi_align_forward_ex(const b200_level* levels, int n_levels, int layout, int batch, int channels,
                                         const float* rois, int64_t n_rois, int pooled_h, int pooled_w,
                                         int sampling_ratio, int math, float* out, float* out_mean,
                                         int32_t* out_levels, void* stream) {
  B200_REQUIRE(math == B200_ROI_MATH_EXACT || math == B200_ROI_MATH_FAST, "roi_align: bad math mode %d", math);
  return b200::roi_align_forward_impl(math == B200_ROI_MATH_EXACT, levels, n_levels, layout, batch, channels, rois,
                                      n_rois, pooled_h, pooled_w, sampling_ratio, out, out_mean, out_levels, nullptr, 0,
                                      stream);
}

extern "C" int b200_roi_align_forward_bf16(const b200_level* levels, int n_levels, int layout, int batch, int channels,
                                           const float* rois, int64_t n_rois, int pooled_h, int pooled_w,
                                           int sampling_ratio, float* out, float* out_mean, int32_t* out_levels,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  using namespace b200;
  B200_REQUIRE(batch > 0 && channels > 0 && pooled_h > 0 && pooled_w > 0 && n_rois >= 0, "roi_align: bad shape");
  if (n_rois == 0) return B200_OK;
  B200_REQUIRE(rois && out && aligned16(out), "roi_align: null rois / out, or out not 16-byte aligned");
  LevelTable lt;
  int rc = fill_level_table(levels, n_levels, &lt);
  if (rc != B200_OK) return rc;
  if (layout != B200_LAYOUT_NHWC || sampling_ratio != 2 || !rows_kernel_applies(lt, channels, pooled_h, pooled_w)) {
    set_error("roi_align bf16: only NHWC maps with 256 channels, 7x7 bins and sampling ratio 2 have a bf16 kernel "
              "(cast other shapes to fp32, as the reference does)");
    return B200_ERR_UNSUPPORTED;
  }
  int32_t* order_ws = workspace && workspace_bytes >= rows_order_workspace_bytes(n_rois) &&
                              (reinterpret_cast<uintptr_t>(workspace) & 3u) == 0
                          ? static_cast<int32_t*>(workspace)
                          : nullptr;
  return launch_forward_rows(lt, channels, true, rois, n_rois, out, out_mean, out_levels, order_ws, g_variant,
                             static_cast<cudaStream_t>(stream));
}

extern "C" size_t b200_roi_align_workspace_bytes(int64_t n_rois) { return b200::rows_order_workspace_bytes(n_rois); }

extern "C" int b200_roi_align_forward_ws(const b200_level* levels, int n_levels, int layout, int batch, int channels,
                                         const float* rois, int64_t n_rois, int pooled_h, int pooled_w,
                                         int sampling_ratio, int math, float* out, float* out_mean,
                                         int32_t* out_levels, void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(math == B200_ROI_MATH_EXACT || math == B200_ROI_MATH_FAST, "roi_align: bad math mode %d", math);
  return b200::roi_align_forward_impl(math == B200_ROI_MATH_EXACT, levels, n_levels, layout, batch, channels, rois,
                                      n_rois, pooled_h, pooled_w, sampling_ratio, out, out_mean, out_levels, workspace,
                                      workspace_bytes, stream);
}
