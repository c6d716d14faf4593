// roi_align_bwd.cu -- fused multi-level RoIAlign backward (gradient w.r.t. the features).
//
// Replaces _C.roi_align_backward (reference csrc/ROIAlign.h:27-45, kernel
// csrc/cuda/ROIAlign_cuda.cu:178-254) for all FPN levels in one launch.  Each addend is
// rounded exactly as the reference rounds it -- (top * w_k) / count, :239-242 -- only the
// order of the additions (atomics there, atomics here) is free.
//
//   roi_align_bwd_taprow  NHWC, sampling ratio 2 (the default for the FPN poolers): a thread owns (4 channels, one
//         distinct tap row), one red.global.add.v4.f32 per (RoI, tap pixel, quad); see the comment at the kernel.
//   roi_align_bwd_march   the round-1 formulation, a thread per output row (b200_debug_bwd(2), tests).
//   roi_align_bwd_generic any sampling ratio / layout.  NHWC: a thread owns (4 channels, one bin); every tap is ONE
//         128-bit vector reduction instead of four scalar atomics.  NCHW: a thread owns (channel, bin) with scalar
//         reductions, bins fastest so that the reads of grad_out are coalesced.
#include "roi_align_fwd.cuh"

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
// Marching kernel (NHWC, sampling_ratio == 2, PH,PW <= 16, C % 64 == 0): the transpose of the
// separable forward (roi_align_fwd_sep.cu).  A thread owns (4 channels, one output row) and
// walks the row's x-samples left to right.  Each bin's gradient (x 1/count) is spread over the
// two tap COLUMNS of each sample into two float4 column accumulators; when the march leaves a
// column its accumulator is scaled by the (up to four, duplicates merged) tap-row weights and
// written with one 128-bit red.global.add.v4.f32 per row: ~3.3 reductions per column instead
// of 16 per bin, 8 lanes side by side covering one 128-byte line.  One CTA per RoI walks all
// (up to 4) 64-channel chunks so the geometry and tables are built once per RoI.
// The addends are algebraically the reference's (top * w_y * w_x) / count
// (ROIAlign_cuda.cu:161-172, :239-242) summed per (row, column) before they reach memory;
// fp32 rounding differs from the per-tap kernel by reassociation only, and the order of the
// atomic additions is free in the reference too.
// ---------------------------------------------------------------------------------------
constexpr int kChunkB = 64;

__device__ __forceinline__ float4 lds128b(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds32b(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float ldsf(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}

__device__ __forceinline__ void red_col(char* const (&rp)[4], const bool (&use)[4], const float (&w)[4], uint32_t co,
                                        const float4& t) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (use[k])
      red_add_v4(reinterpret_cast<float*>(rp[k] + co), w[k] * t.x, w[k] * t.y, w[k] * t.z, w[k] * t.w);
}

template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
roi_align_bwd_march(const LevelGradTable lt, int C, const float* __restrict__ rois, const int32_t* __restrict__ order,
                    int PH, int PW, int chunks_per_cta, const float* __restrict__ grad_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NB = PH * PW;
  const bool swz = (NB & 3) == 0;
  float* g_s = reinterpret_cast<float*>(smem_raw);
  AxisEntry* ytab = reinterpret_cast<AxisEntry*>(smem_raw + sizeof(float) * (kChunkB * NB + kTilePadFloats));
  XSample* xs = reinterpret_cast<XSample*>(ytab + kMaxAxisSamples);
  int* colofs = reinterpret_cast<int*>(xs + kMaxAxisSamples + 1);
  __shared__ int ncols_s;

  const int tid = threadIdx.x;
  const int groups = C / (kChunkB * chunks_per_cta);
  // `order` (optional): consecutive CTAs take RoIs that are neighbours in one feature map
  const long long r = order ? (long long)order[blockIdx.x / groups] : (long long)(blockIdx.x / groups);
  const int c0 = (blockIdx.x % groups) * chunks_per_cta * kChunkB;
  const float* p = rois + r * 5;
  const int batch = (int)p[0];
  const float x1 = p[1], y1 = p[2], x2 = p[3], y2 = p[4];
  const int level = lt.n_levels == 1 ? 0 : fpn_level(x1, y1, x2, y2, lt.k_min, lt.k_max);
  if (level < 0) return;
  const int H = lt.H[level], W = lt.W[level];
  const RoiGeom g = roi_geometry(x1, y1, x2, y2, lt.scale[level], PH, PW, 2);
  const int warp = tid >> 5, lane = tid & 31;
  build_sep_tables(g, PH, PW, H, W, C, warp, lane, ytab, xs, colofs, &ncols_s);
  __syncthreads();

  const int ncols = ncols_s;
  constexpr int kGroups = kChunkB / 4;
  constexpr int kWarps = kThreads / 32;
  constexpr int kLQ = kWarps <= 4 ? 8 : 16;  // channel quads side by side: 128 / 256 B runs per reduction
  constexpr int kQGroups = kGroups / kLQ;
  const int q = (warp % kQGroups) * kLQ + (lane % kLQ);
  const int ph = lane / kLQ + (32 / kLQ) * (warp / kQGroups);
  const bool active = ph < PH;
  int row[4] = {0, 0, 0, 0};
  float w[4] = {0.f, 0.f, 0.f, 0.f};
  bool use[4] = {false, false, false, false};
  if (active) merge_tap_rows(ytab[2 * ph], ytab[2 * ph + 1], row, w, use);
  float* img = lt.data[level] + (size_t)batch * H * W * C + 4 * q;
  const uint32_t xs_a = smem_u32(xs), co_a = smem_u32(colofs);
  const uint32_t g_a = smem_u32(g_s) + 4u * (uint32_t)(tile_row(4 * q, NB, swz) + ph * PW);
  const uint32_t nb4 = 4u * (uint32_t)NB;

  for (int cc = 0; cc < chunks_per_cta; ++cc) {
    const int c_begin = c0 + cc * kChunkB;
    if (cc > 0) __syncthreads();  // everyone is done reading the previous chunk's tile
    tile_copy_in(g_s, grad_out + ((size_t)r * C + c_begin) * NB, NB, swz, tid, kThreads);
    __syncthreads();
    if (active) {
      char* rp[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        rp[k] = reinterpret_cast<char*>(img + c_begin + (unsigned)row[k]);
        asm volatile("" : "+l"(rp[k]));  // keep the row pointers in registers (see roi_align_fwd_sep.cu)
      }
      float4 t[2];
      t[0] = t[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 top = make_float4(0.f, 0.f, 0.f, 0.f);
      int s = 0;
      float4 e = lds128b(xs_a);  // (jhi, l, h, -)
      for (int j0 = 0; j0 < ncols; j0 += 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = j0 + u;
          if (j < ncols) {
            // samples whose right tap column is j: they add to columns j-1 (t[u^1]) and j (t[u])
            while (__float_as_int(e.x) == j) {
              if (!(s & 1)) {
                const uint32_t ga = g_a + 4u * (uint32_t)(s >> 1);
                top.x = 0.25f * ldsf(ga);  // 1 / count, count = 4
                top.y = 0.25f * ldsf(ga + nb4);
                top.z = 0.25f * ldsf(ga + 2u * nb4);
                top.w = 0.25f * ldsf(ga + 3u * nb4);
              }
              t[u ^ 1].x = fmaf(e.z, top.x, t[u ^ 1].x);
              t[u ^ 1].y = fmaf(e.z, top.y, t[u ^ 1].y);
              t[u ^ 1].z = fmaf(e.z, top.z, t[u ^ 1].z);
              t[u ^ 1].w = fmaf(e.z, top.w, t[u ^ 1].w);
              t[u].x = fmaf(e.y, top.x, t[u].x);
              t[u].y = fmaf(e.y, top.y, t[u].y);
              t[u].z = fmaf(e.y, top.z, t[u].z);
              t[u].w = fmaf(e.y, top.w, t[u].w);
              ++s;
              e = lds128b(xs_a + 16u * (uint32_t)s);
            }
            // column j-1 is complete: nothing right of column j's samples touches it
            if (j > 0) {
              red_col(rp, use, w, lds32b(co_a + 4u * (uint32_t)(j - 1)), t[u ^ 1]);
              t[u ^ 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
      }
      red_col(rp, use, w, lds32b(co_a + 4u * (uint32_t)(ncols - 1)), t[(ncols - 1) & 1]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Marching backward BY TAP ROW (the default for NHWC, sampling ratio 2).  roi_align_bwd_march above gives a thread
// one OUTPUT row: every output row reduces into its own (up to four) tap rows, so a tap row that feeds two or three
// output rows -- every one at 14 x 14, where samples lie 0.5-1 pixel apart -- is reduced into two or three times.  The
// kernel is bound by the L2's reduction rate ("RoIAlign backward"), so those bytes are the time.  Here a
// thread owns (4 channels, one DISTINCT tap row): the gradient of its row is first gathered over the output rows
// that touch it, top(pw) = sum_ph wy[row][ph] * g[ph][pw] (shared-memory reads of the staged gradient tile), the same
// x march spreads it over the tap columns, and every (RoI, tap pixel, channel quad) leaves as exactly ONE
// red.global.add.v4.f32.  kCA chunks of 64 channels are staged at once; the work items (chunk, tap row, quad) are
// dealt out flat, so that 16 * kCA consecutive threads reduce into 256 * kCA contiguous bytes of one pixel.
// The addends are the reference's (top * w_y * w_x) / count summed per (RoI, pixel) before they reach memory.
// ---------------------------------------------------------------------------------------
constexpr int kMaxTapRows = 2 * kMaxAxisSamples;  // distinct tap rows of a RoI, at most (two per y sample)
constexpr int kTapRowPH = 16;                     // PH <= 16

struct __align__(16) TapRowTables {
  int rowofs[kMaxTapRows];         // element offset y * W * C of distinct tap row i (ascending y)
  int phr[kMaxTapRows];            // first | last << 8 output row with a weight on row i (first > last: none)
  float wy[kMaxTapRows][kTapRowPH];  // weight x 1/count of tap row i for output row ph
  int nrows;
};

// one warp: the distinct tap rows of the RoI's 2 * PH y samples (PH <= 16) and their weights per output row
__device__ __forceinline__ void build_taprow_tables(const RoiGeom& g, int PH, int H, int W, int C, int lane,
                                                    TapRowTables* tt) {
  float4* z = reinterpret_cast<float4*>(&tt->wy[0][0]);
  for (int i = lane; i < kMaxTapRows * kTapRowPH / 4; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int ns = 2 * PH;
  bool ok = false;
  AxisTap t;
  t.lo = t.hi = 0;
  t.l = t.h = 0.f;
  if (lane < ns) t = axis_sample(g.start_h, lane >> 1, g.bin_h, lane & 1, 2, H, ok);
  ok = ok && lane < ns;
  // sample positions ascend with the lane, so do lo and hi, and the out-of-range samples sit at the two ends: a
  // sample's rows are new iff they lie above the previous valid sample's hi row
  const int phi = __shfl_up_sync(0xffffffffu, t.hi, 1);
  const bool pok = __shfl_up_sync(0xffffffffu, (int)ok, 1) != 0;
  const int pmax = (lane > 0 && pok) ? phi : -1;
  const bool new_lo = ok && t.lo > pmax, new_hi = ok && t.hi > pmax && t.hi != t.lo;
  int scan = (int)new_lo + (int)new_hi;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, scan, d);
    if (lane >= d) scan += v;
  }
  const int nrows = __shfl_sync(0xffffffffu, scan, 31);
  const int jhi = scan - 1, jlo = t.lo == t.hi ? jhi : jhi - 1;  // (hi is the highest row listed so far)
  if (new_hi) tt->rowofs[jhi] = t.hi * W * C;
  if (new_lo) tt->rowofs[jlo] = t.lo * W * C;
  __syncwarp();  // the matrix is zeroed
  // the two samples of an output row may meet in one matrix element: even samples first, then the odd ones
#pragma unroll
  for (int par = 0; par < 2; ++par) {
    if (ok && (lane & 1) == par) {
      const int ph = lane >> 1;
      tt->wy[jlo][ph] += 0.25f * t.h;  // (lo == hi, a sample clamped to the border: both weights on the one row)
      tt->wy[jhi][ph] += 0.25f * t.l;
    }
    __syncwarp();
  }
  for (int i = lane; i < nrows; i += 32) {
    int first = PH, last = -1;
    for (int ph = 0; ph < PH; ++ph)
      if (tt->wy[i][ph] != 0.f) {
        first = min(first, ph);
        last = ph;
      }
    tt->phr[i] = last < 0 ? (1 | (0 << 8)) : (first | (last << 8));
  }
  if (lane == 0) tt->nrows = nrows;
}

template <int kThreads, int kMinBlocks, int kCA>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
roi_align_bwd_taprow(const LevelGradTable lt, int C, const float* __restrict__ rois, const int32_t* __restrict__ order,
                     int PH, int PW, int groups_per_cta, const float* __restrict__ grad_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NB = PH * PW;
  const bool swz = (NB & 3) == 0;
  const int tile_floats = kChunkB * NB + kTilePadFloats;
  float* g_s = reinterpret_cast<float*>(smem_raw);
  AxisEntry* ytab = reinterpret_cast<AxisEntry*>(smem_raw + sizeof(float) * kCA * tile_floats);
  XSample* xs = reinterpret_cast<XSample*>(ytab + kMaxAxisSamples);
  int* colofs = reinterpret_cast<int*>(xs + kMaxAxisSamples + 1);
  __shared__ int ncols_s;
  __shared__ TapRowTables tt;

  const int tid = threadIdx.x;
  const int cta_channels = kChunkB * kCA * groups_per_cta;
  const int ctas_per_roi = C / cta_channels;
  // `order` (optional): consecutive CTAs take RoIs that are neighbours in one feature map
  const long long r = order ? (long long)order[blockIdx.x / ctas_per_roi] : (long long)(blockIdx.x / ctas_per_roi);
  const int c0 = (blockIdx.x % ctas_per_roi) * cta_channels;
  const float* p = rois + r * 5;
  const int batch = (int)p[0];
  const float x1 = p[1], y1 = p[2], x2 = p[3], y2 = p[4];
  const int level = lt.n_levels == 1 ? 0 : fpn_level(x1, y1, x2, y2, lt.k_min, lt.k_max);
  if (level < 0) return;
  const int H = lt.H[level], W = lt.W[level];
  const RoiGeom g = roi_geometry(x1, y1, x2, y2, lt.scale[level], PH, PW, 2);
  const int warp = tid >> 5, lane = tid & 31;
  build_sep_tables(g, PH, PW, H, W, C, warp, lane, ytab, xs, colofs, &ncols_s);  // (warp 1: the column tables)
  if (warp == 2) build_taprow_tables(g, PH, H, W, C, lane, &tt);
  __syncthreads();

  const int ncols = ncols_s, nrows = tt.nrows;
  constexpr int kQ = kChunkB / 4;  // channel quads of a chunk
  char* img = reinterpret_cast<char*>(lt.data[level] + (size_t)batch * H * W * C);
  const uint32_t xs_a = smem_u32(xs), co_a = smem_u32(colofs), g0_a = smem_u32(g_s), wy_a = smem_u32(&tt.wy[0][0]);
  const uint32_t nb4 = 4u * (uint32_t)NB;

  for (int gg = 0; gg < groups_per_cta; ++gg) {
    const int c_begin = c0 + gg * kCA * kChunkB;
    if (gg > 0) __syncthreads();  // everyone is done reading the previous group's tiles
#pragma unroll
    for (int ca = 0; ca < kCA; ++ca)
      tile_copy_in(g_s + ca * tile_floats, grad_out + ((size_t)r * C + c_begin + ca * kChunkB) * NB, NB, swz, tid, kThreads);
    __syncthreads();
    for (int u = tid; u < nrows * kQ * kCA; u += kThreads) {
      const int qc = u % (kQ * kCA), i = u / (kQ * kCA);
      const int ca = qc / kQ, q = qc % kQ;
      const int phr = tt.phr[i];
      const int ph_lo = phr & 0xff, ph_hi = phr >> 8;
      if (ph_lo > ph_hi) continue;
      char* rp = img + 4 * (size_t)(c_begin + ca * kChunkB + 4 * q + tt.rowofs[i]);
      // this thread's four channel rows of the staged gradient tile, and its row of the weight matrix
      const uint32_t g_a = g0_a + 4u * (uint32_t)(ca * tile_floats + tile_row(4 * q, NB, swz));
      const uint32_t w_a = wy_a + 4u * (uint32_t)(i * kTapRowPH);
      float4 t[2];
      t[0] = t[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 top = make_float4(0.f, 0.f, 0.f, 0.f);
      int s = 0;
      float4 e = lds128b(xs_a);  // (jhi, l, h, -)
      for (int j0 = 0; j0 < ncols; j0 += 2) {
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
          const int j = j0 + uu;
          if (j < ncols) {
            // samples whose right tap column is j: they add to columns j-1 (t[uu^1]) and j (t[uu])
            while (__float_as_int(e.x) == j) {
              if (!(s & 1)) {
                // the row's gradient for output column pw = s / 2, gathered over the output rows that touch it
                top = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int ph = ph_lo; ph <= ph_hi; ++ph) {
                  const float wv = ldsf(w_a + 4u * (uint32_t)ph);
                  const uint32_t ga = g_a + 4u * (uint32_t)(ph * PW + (s >> 1));
                  top.x = fmaf(wv, ldsf(ga), top.x);
                  top.y = fmaf(wv, ldsf(ga + nb4), top.y);
                  top.z = fmaf(wv, ldsf(ga + 2u * nb4), top.z);
                  top.w = fmaf(wv, ldsf(ga + 3u * nb4), top.w);
                }
              }
              t[uu ^ 1].x = fmaf(e.z, top.x, t[uu ^ 1].x);
              t[uu ^ 1].y = fmaf(e.z, top.y, t[uu ^ 1].y);
              t[uu ^ 1].z = fmaf(e.z, top.z, t[uu ^ 1].z);
              t[uu ^ 1].w = fmaf(e.z, top.w, t[uu ^ 1].w);
              t[uu].x = fmaf(e.y, top.x, t[uu].x);
              t[uu].y = fmaf(e.y, top.y, t[uu].y);
              t[uu].z = fmaf(e.y, top.z, t[uu].z);
              t[uu].w = fmaf(e.y, top.w, t[uu].w);
              ++s;
              e = lds128b(xs_a + 16u * (uint32_t)s);
            }
            // column j-1 is complete: nothing right of column j's samples touches it
            if (j > 0) {
              const float4 v = t[uu ^ 1];
              red_add_v4(reinterpret_cast<float*>(rp + lds32b(co_a + 4u * (uint32_t)(j - 1))), v.x, v.y, v.z, v.w);
              t[uu ^ 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
      }
      const float4 v = t[(ncols - 1) & 1];
      red_add_v4(reinterpret_cast<float*>(rp + lds32b(co_a + 4u * (uint32_t)(ncols - 1))), v.x, v.y, v.z, v.w);
    }
  }
}

int g_bwd_mode = 0;  // test / tuning hook: 0 = tap-row march, 1 = per-tap kernel, 2 = output-row march

}  // namespace
}  // namespace b200

extern "C" void b200_debug_bwd(int mode) { b200::g_bwd_mode = mode; }

extern "C" int b200_roi_align_backward(const b200_level_grad* levels, int n_levels, int layout, int batch,
                                       int channels, const float* rois, int64_t n_rois, int pooled_h,
                                       int pooled_w, int sampling_ratio, const float* grad_out, void* stream) {
  return b200_roi_align_backward_ws(levels, n_levels, layout, batch, channels, rois, n_rois, pooled_h, pooled_w,
                                    sampling_ratio, grad_out, nullptr, 0, stream);
}

extern "C" int b200_roi_align_backward_ws(const b200_level_grad* levels, int n_levels, int layout, int batch,
                                          int channels, const float* rois, int64_t n_rois, int pooled_h,
                                          int pooled_w, int sampling_ratio, const float* grad_out, void* workspace,
                                          size_t workspace_bytes, void* stream) {
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
  if (g_bwd_mode != 1 && layout == B200_LAYOUT_NHWC && sampling_ratio == 2 && pooled_h <= 16 && pooled_w <= 16 &&
      channels % kChunkB == 0 && aligned16(grad_out)) {
    const int n_chunks = channels / kChunkB;
    int cpc = n_chunks % 4 == 0 ? 4 : (n_chunks % 2 == 0 ? 2 : 1);
    if (n_rois < 2048) cpc = 1;  // few RoIs: keep the grid wide
    const int64_t grid = n_rois * (n_chunks / cpc);
    B200_REQUIRE(grid < (int64_t)1 << 31, "roi_align_bwd: too many RoIs for one launch");
    const size_t smem = sizeof(float) * (kChunkB * NB + kTilePadFloats) + kMaxAxisSamples * sizeof(AxisEntry) +
                        (kMaxAxisSamples + 1) * sizeof(XSample) + kMaxCols * sizeof(int);
    // visiting order (caller's scratch): the same (image, level, Morton cell) order as the forward
    int32_t* order = nullptr;
    if (workspace && workspace_bytes >= rows_order_workspace_bytes(n_rois) && (reinterpret_cast<uintptr_t>(workspace) & 3u) == 0 &&
        n_rois > 1) {
      LevelTable flt;
      for (int l = 0; l < n_levels; ++l) {
        flt.data[l] = lt.data[l];
        flt.H[l] = lt.H[l];
        flt.W[l] = lt.W[l];
        flt.scale[l] = lt.scale[l];
      }
      flt.n_levels = lt.n_levels;
      flt.k_min = lt.k_min;
      flt.k_max = lt.k_max;
      order = static_cast<int32_t*>(workspace);
      int rc = launch_roi_order(flt, rois, n_rois, order, st);
      if (rc != B200_OK) return rc;
    }
    if (g_bwd_mode != 2) {
      // by tap row.  7 x 7: all four chunks of a 256-channel RoI staged at once (50 KB), 14 x 14: one chunk (50 KB)
      const size_t tables = kMaxAxisSamples * sizeof(AxisEntry) + (kMaxAxisSamples + 1) * sizeof(XSample) + kMaxCols * sizeof(int);
      if (NB <= 64) {
        const int ca = n_chunks % 4 == 0 ? 4 : 1;
        const size_t sm = sizeof(float) * ca * (kChunkB * NB + kTilePadFloats) + tables;
        const int64_t grid2 = n_rois * (n_chunks / ca);
        if (ca == 4) {
          auto kern = roi_align_bwd_taprow<256, 4, 4>;
          static SmemHighWater hw;
          int rc = ensure_dynamic_smem(kern, sm, &hw, "roi_align_bwd: smem attribute");
          if (rc != B200_OK) return rc;
          kern<<<(unsigned)grid2, 256, sm, st>>>(lt, channels, rois, order, pooled_h, pooled_w, 1, grad_out);
        } else {
          auto kern = roi_align_bwd_taprow<256, 4, 1>;
          static SmemHighWater hw;
          int rc = ensure_dynamic_smem(kern, sm, &hw, "roi_align_bwd: smem attribute");
          if (rc != B200_OK) return rc;
          kern<<<(unsigned)grid2, 256, sm, st>>>(lt, channels, rois, order, pooled_h, pooled_w, 1, grad_out);
        }
      } else {
        const size_t sm = sizeof(float) * (kChunkB * NB + kTilePadFloats) + tables;
        auto kern = roi_align_bwd_taprow<256, 4, 1>;
        static SmemHighWater hw;
        int rc = ensure_dynamic_smem(kern, sm, &hw, "roi_align_bwd: smem attribute");
        if (rc != B200_OK) return rc;
        kern<<<(unsigned)grid, 256, sm, st>>>(lt, channels, rois, order, pooled_h, pooled_w, cpc, grad_out);
      }
      B200_CHECK_LAUNCH("roi_align_bwd_taprow");
      return B200_OK;
    }
    if (pooled_h <= 8) {
      auto kern = roi_align_bwd_march<128, 8>;
      static SmemHighWater hw;
      int rc = ensure_dynamic_smem(kern, smem, &hw, "roi_align_bwd: smem attribute");
      if (rc != B200_OK) return rc;
      kern<<<(unsigned)grid, 128, smem, st>>>(lt, channels, rois, order, pooled_h, pooled_w, cpc, grad_out);
    } else {
      auto kern = roi_align_bwd_march<256, 4>;
      static SmemHighWater hw;
      int rc = ensure_dynamic_smem(kern, smem, &hw, "roi_align_bwd: smem attribute");
      if (rc != B200_OK) return rc;
      kern<<<(unsigned)grid, 256, smem, st>>>(lt, channels, rois, order, pooled_h, pooled_w, cpc, grad_out);
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
