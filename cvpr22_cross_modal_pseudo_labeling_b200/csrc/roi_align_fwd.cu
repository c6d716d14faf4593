// roi_align_fwd.cu -- fused multi-level RoIAlign forward for sm_100a.
//
// Replaces Pooler.forward (reference modeling/poolers.py:91-121): level assignment,
// per-level _C.roi_align_forward (csrc/cpu/ROIAlign_cpu.cpp:114-217,
// csrc/cuda/ROIAlign_cuda.cu:65-122) and the scatter back, in one launch.
//
// Two kernels:
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
// multiplies and adds in its order); kExact=false lets the 4-tap sum use FMAs.
#include "roi_geom.cuh"

namespace b200 {
namespace {

constexpr int kChunk = 64;        // channels per CTA in the staged kernel

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

__device__ __forceinline__ void zero_fill(float* __restrict__ p, int n, int tid, int nthreads) {
  for (int i = tid; i < n; i += nthreads) p[i] = 0.f;
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
                    float* __restrict__ out, int32_t* __restrict__ out_levels) {
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
}

// ---------------------------------------------------------------------------------------
// Separable marching kernel (NHWC, sampling_ratio == 2) -- the "fast math" forward.
// Bilinear interpolation factorises: a bin's sum is
//     sum_ix [ hx * T(x_lo) + lx * T(x_hi) ],   T(x) = sum_k w_k * f(row_k, x)
// where row_k / w_k are the (up to four, duplicates merged) tap rows of the output row's two
// y-samples.  A thread (4 channels, one output row) reduces every feature column it needs to
// ONE float4 as soon as the column's tap rows arrive and never keeps raw taps, which frees
// the registers to keep kDepth columns in flight per thread.
// What bounds this kernel is the L1 data pipe (ncu: l1tex__data_pipe_lsu_wavefronts), one
// wavefront per 128-byte line a request touches -- so 8 lanes of a warp cover one 128 B run of a
// pixel (8 channel quads), duplicate tap rows are not loaded twice, and the tile stores are
// bank-conflict free.
// The result differs from the reference's summation order by fp32 reassociation only
// (<= 1e-5 relative, tests/test_gpu_roi_align.py); the exact kernel stays available.
// ---------------------------------------------------------------------------------------
struct XSample {
  int jhi;     // index (in the CTA's column list) of the sample's right tap column
  float l, h;  // weights of the right / left tap (left == right column: l = l + h, h = 0)
  int pad;
};
constexpr int kMaxCols = 2 * kMaxAxisSamples;

__device__ __forceinline__ float4 ldg4b(const char* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float4 fma4(float w, float4 v, float4 a) {
  return make_float4(fmaf(w, v.x, a.x), fmaf(w, v.y, a.y), fmaf(w, v.z, a.z), fmaf(w, v.w, a.w));
}

template <int kThreads, int kMinBlocks, int kDepth>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
roi_align_fwd_sep(const LevelTable lt, int C, const float* __restrict__ rois, int PH, int PW,
                  float* __restrict__ out, int32_t* __restrict__ out_levels) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NB = PH * PW;
  const bool swz = (NB & 3) == 0;
  float* out_s = reinterpret_cast<float*>(smem_raw);
  AxisEntry* ytab = reinterpret_cast<AxisEntry*>(smem_raw + sizeof(float) * (kChunk * NB + kTilePadFloats));
  XSample* xs = reinterpret_cast<XSample*>(ytab + kMaxAxisSamples);
  int* colofs = reinterpret_cast<int*>(xs + kMaxAxisSamples + 1);  // byte offsets
  __shared__ int ncols_s;

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
    return;
  }
  const int H = lt.H[h.level], W = lt.W[h.level];
  const RoiGeom g = roi_geometry(h.x1, h.y1, h.x2, h.y2, lt.scale[h.level], PH, PW, 2);
  const int warp = tid >> 5, lane = tid & 31;
  const int ns = 2 * PW;

  // warp 0: row table.  warp 1: sample table + the list of distinct columns in the order the
  // march needs them (a sample either reuses the previous columns, shifts by one, or starts
  // a new pair).
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
      ncols_s = scan;
      xs[ns].jhi = -1;  // sentinel: ends the consume loop after the last sample
    }
  }
  __syncthreads();

  const int ncols = ncols_s;
  const float* feat = lt.data[h.level] + (size_t)h.batch * H * W * C + c_begin;
  constexpr int kWarps = kThreads / 32;
  constexpr int kGroups = kChunk / 4;
  constexpr int kLQ = kWarps <= 4 ? 8 : 16;  // channel quads side by side in a warp (128 / 256 B runs)
  constexpr int kQGroups = kGroups / kLQ;
  {
    const int q = (warp % kQGroups) * kLQ + (lane % kLQ);
    const int ph = lane / kLQ + (32 / kLQ) * (warp / kQGroups);
    if (ph < PH) {
      const AxisEntry ya = ytab[2 * ph], yb = ytab[2 * ph + 1];
      int row[4] = {ya.lo, ya.hi, yb.lo, yb.hi};
      float w[4] = {ya.h, ya.l, yb.h, yb.l};
      bool use[4] = {true, true, true, true};
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
      const char* rp[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) rp[k] = reinterpret_cast<const char*>(feat + (unsigned)(row[k] + 4 * q));
      float4 raw[kDepth][4];
#pragma unroll
      for (int d = 0; d < kDepth; ++d) {
#pragma unroll
        for (int k = 0; k < 4; ++k) raw[d][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d < ncols) {
          const unsigned co = (unsigned)colofs[d];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (use[k]) raw[d][k] = ldg4b(rp[k] + co);
        }
      }
      float4 t[2];
      t[0] = t[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float* o0 = out_s + tile_row(4 * q, NB, swz) + ph * PW;  // the quad's 4 rows share one shift
      int s = 0;
      XSample e = xs[0];
      constexpr int kUnroll = (kDepth % 2 == 0) ? kDepth : 2 * kDepth;
      for (int j0 = 0; j0 < ncols; j0 += kUnroll) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int j = j0 + u;
          if (j < ncols) {
            float4 (&rw)[4] = raw[u % kDepth];
            float4 tv = make_float4(w[0] * rw[0].x, w[0] * rw[0].y, w[0] * rw[0].z, w[0] * rw[0].w);
            tv = fma4(w[1], rw[1], tv);
            tv = fma4(w[2], rw[2], tv);
            tv = fma4(w[3], rw[3], tv);
            t[u & 1] = tv;
            if (j + kDepth < ncols) {
              const unsigned co = (unsigned)colofs[j + kDepth];
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (use[k]) rw[k] = ldg4b(rp[k] + co);
            }
            while (e.jhi == j) {
              acc = fma4(e.h, t[(u & 1) ^ 1], acc);
              acc = fma4(e.l, t[u & 1], acc);
              if (s & 1) {
                float* o = o0 + (s >> 1);
                o[0] = acc.x * 0.25f;
                o[NB] = acc.y * 0.25f;
                o[2 * NB] = acc.z * 0.25f;
                o[3 * NB] = acc.w * 0.25f;
                acc = make_float4(0.f, 0.f, 0.f, 0.f);
              }
              ++s;
              e = xs[s];  // xs[ns] is a sentinel (jhi = -1)
            }
          }
        }
      }
    }
  }
  __syncthreads();
  tile_copy_out(out_s, out_roi + (size_t)c_begin * NB, NB, swz, tid, kThreads);
}

size_t sep_smem_bytes(int NB) {
  return sizeof(float) * (kChunk * NB + kTilePadFloats) + kMaxAxisSamples * sizeof(AxisEntry) +
         (kMaxAxisSamples + 1) * sizeof(XSample) + kMaxCols * sizeof(int);
}

size_t march_smem_bytes(int NB) {
  return sizeof(float) * (kChunk * NB + kTilePadFloats) + 2 * kMaxAxisSamples * sizeof(AxisEntry);
}

bool g_force_generic = false;
bool g_exact = true;
int g_variant = 0;  // tuning hook: occupancy variant of the marching kernel

template <bool kExact, int kThreads, int kMinBlocks, int V>
int launch_march(const LevelTable& lt, int C, const float* rois, int64_t n_rois, int PH, int PW, float* out,
                 int32_t* out_levels, cudaStream_t st) {
  const size_t smem = march_smem_bytes(PH * PW);
  auto kern = roi_align_fwd_march<kExact, kThreads, kMinBlocks, V>;
  static SmemHighWater hw;  // one per template instantiation
  int rc = ensure_dynamic_smem(kern, smem, &hw, "roi_align: smem attribute");
  if (rc != B200_OK) return rc;
  const int64_t grid = n_rois * (C / kChunk);
  kern<<<(unsigned)grid, kThreads, smem, st>>>(lt, C, rois, PH, PW, out, out_levels);
  B200_CHECK_LAUNCH("roi_align_fwd_march");
  return B200_OK;
}

template <int kThreads, int kMinBlocks, int kDepth>
int launch_sep(const LevelTable& lt, int C, const float* rois, int64_t n_rois, int PH, int PW, float* out,
               int32_t* out_levels, cudaStream_t st) {
  const size_t smem = sep_smem_bytes(PH * PW);
  auto kern = roi_align_fwd_sep<kThreads, kMinBlocks, kDepth>;
  static SmemHighWater hw;
  int rc = ensure_dynamic_smem(kern, smem, &hw, "roi_align: smem attribute");
  if (rc != B200_OK) return rc;
  const int64_t grid = n_rois * (C / kChunk);
  kern<<<(unsigned)grid, kThreads, smem, st>>>(lt, C, rois, PH, PW, out, out_levels);
  B200_CHECK_LAUNCH("roi_align_fwd_sep");
  return B200_OK;
}

template <bool kExact>
int launch_forward(const LevelTable& lt, int layout, int C, const float* rois, int64_t n_rois, int PH, int PW,
                   int sr, float* out, int32_t* out_levels, cudaStream_t st) {
  const int NB = PH * PW;
  const bool march_ok = !g_force_generic && layout == B200_LAYOUT_NHWC && sr == 2 && PH <= 16 && PW <= 16 &&
                        C % kChunk == 0 && (NB * kChunk) % 4 == 0;
  B200_REQUIRE(n_rois * ((C + 15) / 16) < (int64_t)1 << 31, "roi_align: too many RoIs for one launch");
  if (march_ok && !kExact) {
    if (PH * (kChunk / 4) <= 128) {
      if (g_variant == 1) return launch_sep<128, 6, 2>(lt, C, rois, n_rois, PH, PW, out, out_levels, st);
      if (g_variant == 2) return launch_sep<128, 6, 3>(lt, C, rois, n_rois, PH, PW, out, out_levels, st);
      return launch_sep<128, 8, 2>(lt, C, rois, n_rois, PH, PW, out, out_levels, st);
    }
    if (g_variant == 1) return launch_sep<256, 3, 2>(lt, C, rois, n_rois, PH, PW, out, out_levels, st);
    if (g_variant == 2) return launch_sep<256, 3, 3>(lt, C, rois, n_rois, PH, PW, out, out_levels, st);
    return launch_sep<256, 4, 2>(lt, C, rois, n_rois, PH, PW, out, out_levels, st);
  }
  if (march_ok) {
    // g_variant (tuning hook): CTAs/SM = 6 (default) / 5 / 4 for 128 threads, 3 / 3 / 2 for 256.
    // (2 channels per lane with twice the warps was measured 35-55 % slower: V stays 4.)
    if (PH * (kChunk / 4) <= 128) {
      if (g_variant == 1) return launch_march<kExact, 128, 5, 4>(lt, C, rois, n_rois, PH, PW, out, out_levels, st);
      if (g_variant == 2) return launch_march<kExact, 128, 4, 4>(lt, C, rois, n_rois, PH, PW, out, out_levels, st);
      return launch_march<kExact, 128, 6, 4>(lt, C, rois, n_rois, PH, PW, out, out_levels, st);
    }
    if (g_variant == 2) return launch_march<kExact, 256, 2, 4>(lt, C, rois, n_rois, PH, PW, out, out_levels, st);
    return launch_march<kExact, 256, 3, 4>(lt, C, rois, n_rois, PH, PW, out, out_levels, st);
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

extern "C" int b200_roi_align_forward(const b200_level* levels, int n_levels, int layout, int batch,
                                      int channels, const float* rois, int64_t n_rois, int pooled_h,
                                      int pooled_w, int sampling_ratio, float* out, int32_t* out_levels,
                                      void* stream) {
  using namespace b200;
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
  return g_exact ? launch_forward<true>(lt, layout, channels, rois, n_rois, pooled_h, pooled_w, sampling_ratio,
                                        out, out_levels, st)
                 : launch_forward<false>(lt, layout, channels, rois, n_rois, pooled_h, pooled_w,
                                         sampling_ratio, out, out_levels, st);
}
