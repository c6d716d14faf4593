// roi_align_fwd_sep.cu -- separable ("fast math") marching RoIAlign forward for sm_100a.
//
// Same job as roi_align_fwd_march (roi_align_fwd.cu): NHWC features, sampling_ratio 2, the
// FPN box (7x7) and mask (14x14) poolers of modeling/poolers.py:91-121.  Bilinear
// interpolation factorises, so a bin's sum is
//     sum_ix [ hx * T(x_lo) + lx * T(x_hi) ],   T(x) = sum_k w_k * f(row_k, x)
// where row_k / w_k are the (up to four, duplicates merged) tap rows of the output row's two
// y-samples.  A thread (4 channels, one output row) reduces every feature column it needs to
// ONE float4 as soon as the column's tap rows arrive; the x-samples then combine two such
// column values.  That is ~12 FMAs per output element instead of the reference order's 16
// multiplies + 12 adds, and no raw taps stay in registers.
//
// What the profiles say bounds this kernel (profiles/r01_roi_align_fwd_sep.txt): instruction
// issue and the L1 data pipe (one wavefront per 128-byte line a request touches), not HBM or
// L2 -- the time barely moves when every RoI reads one L1-resident patch.  Hence:
//   * one CTA per RoI walks all (up to 4) 64-channel chunks, so the RoI header load, level
//     assignment, geometry and axis tables are paid once per RoI, not once per chunk;
//   * 8 lanes of a warp cover one 128 B run of a pixel (8 channel quads x 4 output rows per
//     warp; 16 x 2 for 14x14), duplicate tap rows are not loaded twice;
//   * table / tile accesses go through 32-bit shared addresses, the four row pointers are
//     pinned in registers (ptxas otherwise re-derives them, 8 integer instructions per load);
//   * the output tile is bank-shifted for 14x14 (roi_align_fwd.cuh).
// Measured dead ends, kept out: two columns in flight per thread, a second output tile
// (copy-out overlapping the next chunk), cp.async.bulk.prefetch.L2 of the patch rows,
// ld.global.nc.L1::no_allocate -- all slower ("RoIAlign forward: what bounds it").
// The result differs from the reference's summation order by fp32 reassociation only
// (<= 1e-5 relative, tests/test_gpu_roi_align.py); the exact kernel stays available.
#include "roi_align_fwd.cuh"

namespace b200 {
namespace {

__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
template <int kOff>
__device__ __forceinline__ void sts32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(a), "n"(kOff), "f"(v) : "memory");
}
__device__ __forceinline__ void sts32r(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

__device__ __forceinline__ float4 ldg4b(const char* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float4 fma4(float w, float4 v, float4 a) {
  return make_float4(fmaf(w, v.x, a.x), fmaf(w, v.y, a.y), fmaf(w, v.z, a.z), fmaf(w, v.w, a.w));
}

// kPH / kPW > 0: compile-time pooled size (7x7, 14x14); 0: taken from the arguments.
template <int kThreads, int kMinBlocks, int kPH, int kPW, int kBufs, int kDepth>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
roi_align_fwd_sep(const LevelTable lt, int C, const float* __restrict__ rois, int PH_rt, int PW_rt,
                  int chunks_per_cta, float* __restrict__ out, float* __restrict__ out_mean,
                  int32_t* __restrict__ out_levels) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int PH = kPH > 0 ? kPH : PH_rt, PW = kPW > 0 ? kPW : PW_rt;
  const int NB = PH * PW;
  const bool swz = (NB & 3) == 0;
  const int tile_floats = kChunk * NB + kTilePadFloats;
  float* out_s = reinterpret_cast<float*>(smem_raw);
  AxisEntry* ytab = reinterpret_cast<AxisEntry*>(smem_raw + sizeof(float) * kBufs * tile_floats);
  XSample* xs = reinterpret_cast<XSample*>(ytab + kMaxAxisSamples);
  int* colofs = reinterpret_cast<int*>(xs + kMaxAxisSamples + 1);  // byte offsets
  __shared__ int ncols_s;

  const int tid = threadIdx.x;
  const int groups = C / (kChunk * chunks_per_cta);
  const long long r = blockIdx.x / groups;
  const int c0 = (blockIdx.x % groups) * chunks_per_cta * kChunk;
  const RoiHeader h = load_roi(rois, r, lt);
  float* out_roi = out + (size_t)r * C * NB;
  if (c0 == 0 && tid == 0 && out_levels) out_levels[r] = h.level;
  if (h.level < 0) {
    zero_fill(out_roi + (size_t)c0 * NB, chunks_per_cta * kChunk * NB, tid, kThreads);
    if (out_mean) zero_fill(out_mean + (size_t)r * C + c0, chunks_per_cta * kChunk, tid, kThreads);
    return;
  }
  const int H = lt.H[h.level], W = lt.W[h.level];
  const RoiGeom g = roi_geometry(h.x1, h.y1, h.x2, h.y2, lt.scale[h.level], PH, PW, 2);
  const int warp = tid >> 5, lane = tid & 31;

  build_sep_tables(g, PH, PW, H, W, C, warp, lane, ytab, xs, colofs, &ncols_s);
  __syncthreads();

  const int ncols = ncols_s;
  constexpr int kWarps = kThreads / 32;
  constexpr int kGroups = kChunk / 4;
  constexpr int kLQ = kWarps <= 4 ? 8 : 16;  // channel quads side by side in a warp (128 / 256 B runs)
  constexpr int kQGroups = kGroups / kLQ;
  const int q = (warp % kQGroups) * kLQ + (lane % kLQ);
  const int ph = lane / kLQ + (32 / kLQ) * (warp / kQGroups);
  const bool active = ph < PH;

  // tap rows of this thread's output row, duplicates merged (weights add up)
  int row[4] = {0, 0, 0, 0};
  float w[4] = {0.f, 0.f, 0.f, 0.f};
  bool use[4] = {false, false, false, false};
  if (active) merge_tap_rows(ytab[2 * ph], ytab[2 * ph + 1], row, w, use);
  const float* img = lt.data[h.level] + (size_t)h.batch * H * W * C + 4 * q;
  uint32_t xs_a = smem_u32(xs), co_a = smem_u32(colofs);
  uint32_t tile_a = smem_u32(out_s) + 4u * (uint32_t)(tile_row(4 * q, NB, swz) + ph * PW);
  // opaque to the optimiser: otherwise ptxas re-derives these shared addresses (S2UR / UMOV /
  // ULEA chains, ~25 % of the issued instructions) at every use instead of keeping 3 registers
  asm volatile("" : "+r"(xs_a), "+r"(co_a), "+r"(tile_a));
  float4 raw[kDepth][4];  // kDepth columns in flight
#pragma unroll
  for (int d = 0; d < kDepth; ++d)
#pragma unroll
    for (int k = 0; k < 4; ++k) raw[d][k] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int cc = 0; cc < chunks_per_cta; ++cc) {
    const int c_begin = c0 + cc * kChunk;
    const int buf = kBufs == 2 ? (cc & 1) : 0;
    if (active) {
      const char* rp[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        rp[k] = reinterpret_cast<const char*>(img + c_begin + (unsigned)row[k]);
        // opaque to the optimiser: keep the four row pointers in registers instead of
        // re-deriving them (8 integer instructions per load) inside the column loop
        asm volatile("" : "+l"(rp[k]));
      }
      const uint32_t o_a = tile_a + 4u * (uint32_t)(buf * tile_floats);
#pragma unroll
      for (int d = 0; d < kDepth; ++d) {
        if (d < ncols) {
          const uint32_t co = lds32(co_a + 4u * d);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (use[k]) raw[d][k] = ldg4b(rp[k] + co);
        }
      }
      float4 t[2];
      t[0] = t[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int s = 0;
      float4 e = lds128(xs_a);  // (jhi, l, h, -)
      constexpr int kUnroll = (kDepth % 2 == 0) ? kDepth : 2 * kDepth;
      for (int j0 = 0; j0 < ncols; j0 += kUnroll) {
#pragma unroll
        for (int uu = 0; uu < kUnroll; ++uu) {
          const int j = j0 + uu;
          constexpr int kDummy = 0;
          (void)kDummy;
          const int u = uu & 1;
          if (j < ncols) {
            float4 (&rw)[4] = raw[uu % kDepth];
            float4 tv = make_float4(w[0] * rw[0].x, w[0] * rw[0].y, w[0] * rw[0].z, w[0] * rw[0].w);
            tv = fma4(w[1], rw[1], tv);
            tv = fma4(w[2], rw[2], tv);
            tv = fma4(w[3], rw[3], tv);
            t[u] = tv;
            if (j + kDepth < ncols) {
              const uint32_t co = lds32(co_a + 4u * (uint32_t)(j + kDepth));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (use[k]) rw[k] = ldg4b(rp[k] + co);
            }
            while (__float_as_int(e.x) == j) {
              acc = fma4(e.z, t[u ^ 1], acc);
              acc = fma4(e.y, t[u], acc);
              if (s & 1) {
                const uint32_t oa = o_a + 4u * (uint32_t)(s >> 1);
                if (kPH > 0) {
                  constexpr int kNB4 = 4 * kPH * kPW;
                  sts32<0>(oa, acc.x * 0.25f);
                  sts32<kNB4>(oa, acc.y * 0.25f);
                  sts32<2 * kNB4>(oa, acc.z * 0.25f);
                  sts32<3 * kNB4>(oa, acc.w * 0.25f);
                } else {
                  sts32r(oa, acc.x * 0.25f);
                  sts32r(oa + 4u * NB, acc.y * 0.25f);
                  sts32r(oa + 8u * NB, acc.z * 0.25f);
                  sts32r(oa + 12u * NB, acc.w * 0.25f);
                }
                acc = make_float4(0.f, 0.f, 0.f, 0.f);
              }
              ++s;
              e = lds128(xs_a + 16u * (uint32_t)s);
            }
          }
        }
      }
    }
    __syncthreads();
    tile_copy_out(out_s + buf * tile_floats, out_roi + (size_t)c_begin * NB, NB, swz, tid, kThreads);
    if (out_mean) tile_mean_out(out_s + buf * tile_floats, out_mean + (size_t)r * C + c_begin, NB, swz, tid);
    // one tile: everyone must be done reading it before the next chunk's stores.  Two tiles:
    // the next chunk writes the other one, and the barrier after THAT chunk orders its
    // copy-out before this tile is written again.
    if (kBufs == 1) __syncthreads();
  }
}

size_t sep_smem_bytes(int NB, int bufs) {
  return sizeof(float) * bufs * (kChunk * NB + kTilePadFloats) + kMaxAxisSamples * sizeof(AxisEntry) +
         (kMaxAxisSamples + 1) * sizeof(XSample) + kMaxCols * sizeof(int);
}

template <int kThreads, int kMinBlocks, int kPH, int kPW, int kBufs, int kDepth>
int launch_sep(const LevelTable& lt, int C, const float* rois, int64_t n_rois, int PH, int PW, int chunks_per_cta,
               float* out, float* out_mean, int32_t* out_levels, cudaStream_t st) {
  const size_t smem = sep_smem_bytes(PH * PW, kBufs);
  auto kern = roi_align_fwd_sep<kThreads, kMinBlocks, kPH, kPW, kBufs, kDepth>;
  static SmemHighWater hw;  // one per template instantiation
  int rc = ensure_dynamic_smem(kern, smem, &hw, "roi_align: smem attribute");
  if (rc != B200_OK) return rc;
  const int64_t grid = n_rois * (C / (kChunk * chunks_per_cta));
  kern<<<(unsigned)grid, kThreads, smem, st>>>(lt, C, rois, PH, PW, chunks_per_cta, out, out_mean, out_levels);
  B200_CHECK_LAUNCH("roi_align_fwd_sep");
  return B200_OK;
}

}  // namespace

// Preconditions (checked by the caller): NHWC, sampling_ratio 2, PH, PW <= 16, C % 64 == 0,
// (PH * PW * 64) % 4 == 0.  `variant` is the tuning hook of b200_debug_set.
int launch_forward_sep(const LevelTable& lt, int C, const float* rois, int64_t n_rois, int PH, int PW, float* out,
                       float* out_mean, int32_t* out_levels, int variant, cudaStream_t st) {
  const int n_chunks = C / kChunk;
  int cpc = n_chunks % 4 == 0 ? 4 : (n_chunks % 2 == 0 ? 2 : 1);
  // few RoIs (the mask pooler on the kept detections): one chunk per CTA keeps the grid wide
  if (n_rois < 2048) cpc = 1;
  if (variant & 8) cpc = 1;
  variant &= 7;
#define B200_SEP(T, MB, KPH, KPW, BUFS, D) \
  return launch_sep<T, MB, KPH, KPW, BUFS, D>(lt, C, rois, n_rois, PH, PW, cpc, out, out_mean, out_levels, st)
  // Measured on B200 (scripts/probe_sep.py, B=16 x 1000 RoIs, C=256): one column in flight per
  // thread at 8 CTAs/SM (7x7) / 4 CTAs/SM (14x14) is fastest; two columns in flight or a second
  // output tile cost registers / shared memory, i.e. resident warps, and lose 5-15 %.
  if (PH == 7 && PW == 7) {
    if (variant == 1) B200_SEP(128, 8, 7, 7, 1, 2);
    if (variant == 2) B200_SEP(128, 10, 7, 7, 1, 1);
    B200_SEP(128, 8, 7, 7, 1, 1);
  }
  if (PH == 14 && PW == 14) {
    if (variant == 1) B200_SEP(256, 4, 14, 14, 1, 2);
    if (variant == 2) B200_SEP(256, 5, 14, 14, 1, 1);
    B200_SEP(256, 4, 14, 14, 1, 1);
  }
  if (PH <= 8) B200_SEP(128, 8, 0, 0, 1, 1);
  B200_SEP(256, 4, 0, 0, 1, 1);
#undef B200_SEP
}

}  // namespace b200
