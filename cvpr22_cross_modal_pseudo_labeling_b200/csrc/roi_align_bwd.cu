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

// ---------------------------------------------------------------------------------------
// Marching backward (NHWC, sampling_ratio == 2, PH,PW <= 16, C % 64 == 0): mirror of
// roi_align_fwd_march.  One CTA per (RoI, 64-channel chunk) stages its [64 x NB] tile of
// grad_out in shared memory (one coalesced read).  A thread owns (4 channels, one output row)
// and marches along x with two register columns of ACCUMULATORS (4 tap rows x float4): every
// sample adds its four (top * w) / count terms into them, and a column is written out with
// four red.global.add.v4.f32 only when the march leaves it -- 2-3x fewer reductions than one
// per tap, each 128 bits wide.
// ---------------------------------------------------------------------------------------
constexpr int kChunkB = 64;

struct Acc4 {
  float v[4];
};

// rows[1] (upper sample's high tap row) and rows[2] (lower sample's low tap row) coincide
// whenever both samples of the bin fall in adjacent cells: one reduction then carries both.
__device__ __forceinline__ void flush_col(float* __restrict__ gfeat, const int (&rows)[4], int col, Acc4 (&g)[4]) {
  const bool dup = rows[1] == rows[2];
  if (dup) {
#pragma unroll
    for (int c = 0; c < 4; ++c) g[1].v[c] += g[2].v[c];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!(k == 2 && dup))
      red_add_v4(gfeat + (unsigned)(rows[k] + col), g[k].v[0], g[k].v[1], g[k].v[2], g[k].v[3]);
    g[k].v[0] = g[k].v[1] = g[k].v[2] = g[k].v[3] = 0.f;
  }
}

// add one sample column's contributions: left taps into Lg, right taps into Rg
__device__ __forceinline__ void scatter_pair(Acc4 (&Lg)[4], Acc4 (&Rg)[4], const AxisEntry& ya, const AxisEntry& yb,
                                             const AxisEntry& xt, const float (&top)[4]) {
  const float w[8] = {__fmul_rn(ya.h, xt.h), __fmul_rn(ya.h, xt.l), __fmul_rn(ya.l, xt.h), __fmul_rn(ya.l, xt.l),
                      __fmul_rn(yb.h, xt.h), __fmul_rn(yb.h, xt.l), __fmul_rn(yb.l, xt.h), __fmul_rn(yb.l, xt.l)};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    // (top * w) / count with count = 4: the division is exact, as a multiply by 0.25
    Lg[0].v[c] += __fmul_rn(__fmul_rn(top[c], w[0]), 0.25f);
    Rg[0].v[c] += __fmul_rn(__fmul_rn(top[c], w[1]), 0.25f);
    Lg[1].v[c] += __fmul_rn(__fmul_rn(top[c], w[2]), 0.25f);
    Rg[1].v[c] += __fmul_rn(__fmul_rn(top[c], w[3]), 0.25f);
    Lg[2].v[c] += __fmul_rn(__fmul_rn(top[c], w[4]), 0.25f);
    Rg[2].v[c] += __fmul_rn(__fmul_rn(top[c], w[5]), 0.25f);
    Lg[3].v[c] += __fmul_rn(__fmul_rn(top[c], w[6]), 0.25f);
    Rg[3].v[c] += __fmul_rn(__fmul_rn(top[c], w[7]), 0.25f);
  }
}

template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
roi_align_bwd_march(const LevelGradTable lt, int C, const float* __restrict__ rois, int PH, int PW,
                    const float* __restrict__ grad_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NB = PH * PW;
  float* g_s = reinterpret_cast<float*>(smem_raw);
  AxisEntry* ytab = reinterpret_cast<AxisEntry*>(smem_raw + sizeof(float) * kChunkB * NB);
  AxisEntry* xtab = ytab + kMaxAxisSamples;

  const int tid = threadIdx.x;
  const int n_cchunks = C / kChunkB;
  const long long r = blockIdx.x / n_cchunks;
  const int c_begin = (blockIdx.x % n_cchunks) * kChunkB;
  const float* p = rois + r * 5;
  const int batch = (int)p[0];
  const float x1 = p[1], y1 = p[2], x2 = p[3], y2 = p[4];
  const int level = lt.n_levels == 1 ? 0 : fpn_level(x1, y1, x2, y2, lt.k_min, lt.k_max);
  if (level < 0) return;
  const int H = lt.H[level], W = lt.W[level];
  const RoiGeom g = roi_geometry(x1, y1, x2, y2, lt.scale[level], PH, PW, 2);

  // this CTA's contiguous [64 x NB] block of grad_out
  const float4* src = reinterpret_cast<const float4*>(grad_out + ((size_t)r * C + c_begin) * NB);
  float4* dst = reinterpret_cast<float4*>(g_s);
  for (int i = tid; i < kChunkB * NB / 4; i += kThreads) dst[i] = __ldcs(src + i);
  const int warp = tid >> 5, lane = tid & 31;
  build_axis_tables(g, PH, PW, H, W, C, warp, lane, ytab, xtab);
  __syncthreads();

  constexpr int kGroups = kChunkB / 4;
  constexpr int kWarps = kThreads / 32;
  constexpr int kLQ = kGroups / kWarps >= 4 ? 4 : 16;
  constexpr int kQGroups = kGroups / kLQ;
  const int q = (warp % kQGroups) * kLQ + (lane % kLQ);
  const int ph = lane / kLQ + (32 / kLQ) * (warp / kQGroups);
  if (ph >= PH) return;
  float* gfeat = lt.data[level] + (size_t)batch * H * W * C + c_begin;
  const AxisEntry ya = ytab[2 * ph], yb = ytab[2 * ph + 1];
  const int rows[4] = {ya.lo + 4 * q, ya.hi + 4 * q, yb.lo + 4 * q, yb.hi + 4 * q};
  Acc4 A[4], B[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int c = 0; c < 4; ++c) A[k].v[c] = B[k].v[c] = 0.f;
  int a_is_left = 1, col_l = -1, col_r = -1;  // element offsets of the columns held (-1: empty)
  const float* gt = g_s + (size_t)(4 * q) * NB + ph * PW;
  for (int pw = 0; pw < PW; ++pw) {
    const float top[4] = {gt[pw], gt[NB + pw], gt[2 * NB + pw], gt[3 * NB + pw]};
#pragma unroll
    for (int ix = 0; ix < 2; ++ix) {
      const AxisEntry xt = xtab[2 * pw + ix];
      const int act = xt.lo & 3, xlo = xt.lo & ~3;
      if (act == kActShift) {
        // the left column is finished: write it out, the right one becomes the left
        if (a_is_left) flush_col(gfeat, rows, col_l, A);
        else flush_col(gfeat, rows, col_l, B);
        a_is_left ^= 1;
        col_l = col_r;
        col_r = xt.hi;
      } else if (act == kActLoad2) {
        if (col_l >= 0) {
          if (a_is_left) {
            flush_col(gfeat, rows, col_l, A);
            flush_col(gfeat, rows, col_r, B);
          } else {
            flush_col(gfeat, rows, col_l, B);
            flush_col(gfeat, rows, col_r, A);
          }
        }
        col_l = xlo;
        col_r = xt.hi;
      }
      if (a_is_left) scatter_pair(A, B, ya, yb, xt, top);
      else scatter_pair(B, A, ya, yb, xt, top);
    }
  }
  if (col_l >= 0) {
    if (a_is_left) {
      flush_col(gfeat, rows, col_l, A);
      flush_col(gfeat, rows, col_r, B);
    } else {
      flush_col(gfeat, rows, col_l, B);
      flush_col(gfeat, rows, col_r, A);
    }
  }
}

bool g_bwd_force_generic = false;

}  // namespace
}  // namespace b200

extern "C" void b200_debug_bwd(int force_generic) { b200::g_bwd_force_generic = force_generic != 0; }

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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!g_bwd_force_generic && layout == B200_LAYOUT_NHWC && sampling_ratio == 2 && pooled_h <= 16 && pooled_w <= 16 &&
      channels % kChunkB == 0 && aligned16(grad_out)) {
    const int64_t grid = n_rois * (channels / kChunkB);
    B200_REQUIRE(grid < (int64_t)1 << 31, "roi_align_bwd: too many RoIs for one launch");
    const size_t smem = sizeof(float) * kChunkB * NB + 2 * kMaxAxisSamples * sizeof(AxisEntry);
    if (pooled_h <= 8) {
      auto kern = roi_align_bwd_march<128, 5>;
      static SmemHighWater hw;
      int rc = ensure_dynamic_smem(kern, smem, &hw, "roi_align_bwd: smem attribute");
      if (rc != B200_OK) return rc;
      kern<<<(unsigned)grid, 128, smem, st>>>(lt, channels, rois, pooled_h, pooled_w, grad_out);
    } else {
      auto kern = roi_align_bwd_march<256, 3>;
      static SmemHighWater hw;
      int rc = ensure_dynamic_smem(kern, smem, &hw, "roi_align_bwd: smem attribute");
      if (rc != B200_OK) return rc;
      kern<<<(unsigned)grid, 256, smem, st>>>(lt, channels, rois, pooled_h, pooled_w, grad_out);
    }
    B200_CHECK_LAUNCH("roi_align_bwd_march");
    return B200_OK;
  }
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
  if (layout == B200_LAYOUT_NHWC)
    roi_align_bwd_generic<B200_LAYOUT_NHWC><<<(unsigned)grid, 256, 0, st>>>(
        lt, channels, rois, pooled_h, pooled_w, sampling_ratio, c_per_cta, n_cchunks, grad_out);
  else
    roi_align_bwd_generic<B200_LAYOUT_NCHW><<<(unsigned)grid, 256, 0, st>>>(
        lt, channels, rois, pooled_h, pooled_w, sampling_ratio, c_per_cta, n_cchunks, grad_out);
  B200_CHECK_LAUNCH("roi_align_bwd_generic");
  return B200_OK;
}
