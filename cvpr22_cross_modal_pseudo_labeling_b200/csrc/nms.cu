// nms.cu -- batched (segmented) greedy NMS for sm_100a.
//
// Replaces _C.nms (reference csrc/nms.h:10-28; CPU kernel csrc/cpu/nms_cpu.cpp:6-65,
// CUDA kernel csrc/cuda/nms.cu:23-130) and the per-(image,level) / per-(image,class)
// Python loops around it.  Semantics follow the CPU file: legacy "+1" extents,
// suppress when IoU >= thresh, every fp32 operation rounded on its own
// (the _rn intrinsics below are never contracted into FMAs).
//
// Three kernels, all segments of a batch per launch, nothing returns to the host:
//   1. nms_sort_kernel   one CTA per segment: (score desc, index asc) order by an
//                        in-shared-memory bitonic sort of 64-bit keys; skipped when the
//                        segment is already ordered (RPN top-k output is).
//   2. nms_mask_kernel   64x64 IoU tiles of the upper triangle -> one 64-bit suppression
//                        word per (row box, column tile).
//   3. nms_sweep_kernel  one CTA per segment walks the tiles in order: the diagonal word
//                        chain is resolved serially, the kept rows' words are OR-reduced
//                        into the running "removed" words by the whole CTA; then the kept
//                        ORIGINAL indices are compacted in ascending order.
#include "common.cuh"

namespace {

constexpr int kTile = 64;
constexpr int kSortThreads = 1024;
constexpr int kSweepThreads = 512;
constexpr int kMaxTiles = B200_NMS_MAX_SEG / kTile;  // 256

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t desc_score_bits(float s) {
  s = s + 0.0f;  // -0 -> +0 so both zeros tie
  uint32_t u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-orderable
  return ~u;                                       // descending
}

// (x2 - x1 + 1) * (y2 - y1 + 1), three separately rounded ops per factor (nms_cpu.cpp:22)
__device__ __forceinline__ float legacy_area(const float4 b) {
  float w = __fadd_rn(__fsub_rn(b.z, b.x), 1.0f);
  float h = __fadd_rn(__fsub_rn(b.w, b.y), 1.0f);
  return __fmul_rn(w, h);
}

// nms_cpu.cpp:50-60 for one pair; `a` is the earlier-visited box.
__device__ __forceinline__ bool suppresses(const float4 a, float area_a, const float4 b, float area_b,
                                           float thresh) {
  float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  float w = fmaxf(__fadd_rn(__fsub_rn(xx2, xx1), 1.0f), 0.0f);
  float h = fmaxf(__fadd_rn(__fsub_rn(yy2, yy1), 1.0f), 0.0f);
  float inter = __fmul_rn(w, h);
  // 0/x is 0 or NaN: never >= a positive threshold, so the division can be skipped.
  if (inter == 0.0f && thresh > 0.0f) return false;
  float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  return __fdiv_rn(inter, uni) >= thresh;
}

// Necessary condition for suppresses(a, b): with ow the x-overlap, inter = ow * oh >= t * union
// >= t * w * h of either box and oh <= h, so ow >= t * max(w_a, w_b); and ow <= (w_a + w_b) / 2 -
// |cx_a - cx_b|.  Hence |cx_a - cx_b| <= (1 - t) / 2 * (w_a + w_b) = reach(a) + reach(b).
// reach() is padded (1 % + 0.01 px) against fp32 rounding; thresh <= 0 disables the test.
__device__ __forceinline__ float2 x_reach(const float4 b, float thresh) {
  const float w = b.z - b.x + 1.0f;
  const float r = thresh > 0.0f ? (1.0f - thresh) * 0.5f * w * 1.01f + 0.01f : 3.0e38f;
  return make_float2(0.5f * (b.x + b.z), r);
}
__device__ __forceinline__ uint32_t orderable_f(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Is box b (area_b, x centre / reach cb) suppressed by one of `count` kept boxes?  The three
// lists sit in shared memory (32-bit addresses: float2 centre/reach, float4 box, float area).
// The 32 lanes of a warp hold boxes that sit side by side in the image (the chunk is handed out in
// x order), so the warp first filters the list against the x interval of ITS alive boxes -- 32 kept
// boxes per step, one per lane, merged with a ballot -- and only walks the survivors (a few per
// cent of the list); for those a lane that is out of x reach (see x_reach) skips the IoU test.
// Both filters are necessary conditions only (the warp interval is widened by a pixel against
// fp32 rounding): verdicts are exactly those of testing every pair.
// The thread's own box is only fetched (own_a: its shared address) when a kept box is near.
__device__ __forceinline__ bool survives_list(uint32_t c_a, uint32_t b_a, uint32_t a_a, int count, bool alive,
                                              uint32_t own_a, const float2 cb, float thresh, int lane) {
  // [lo, hi]: union of the alive lanes' reach intervals (signed-orderable ints through redux)
  const uint32_t olo = orderable_f(cb.x - cb.y) ^ 0x80000000u, ohi = orderable_f(cb.x + cb.y) ^ 0x80000000u;
  // (a box with NaN coordinates is never near anything -- every comparison fails, as in nms_cpu -- and must
  // not poison the window of its neighbours)
  const bool win = alive && cb.x - cb.y <= cb.x + cb.y;
  const int ilo = __reduce_min_sync(0xffffffffu, win ? (int)olo : 0x7fffffff);
  const int ihi = __reduce_max_sync(0xffffffffu, win ? (int)ohi : (int)0x80000000);
  if (ilo == 0x7fffffff) return alive;  // nobody alive (uniform)
  const uint32_t ulo = (uint32_t)ilo ^ 0x80000000u, uhi = (uint32_t)ihi ^ 0x80000000u;
  const float lo = __uint_as_float((ulo & 0x80000000u) ? (ulo & 0x7fffffffu) : ~ulo) - 1.0f;
  const float hi = __uint_as_float((uhi & 0x80000000u) ? (uhi & 0x7fffffffu) : ~uhi) + 1.0f;
  for (int r0 = 0; r0 < count; r0 += 32) {
    const int k = r0 + lane;
    float kx = 0.f, kr = -1.0e30f;  // (beyond the list: an empty interval)
    if (k < count) asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(kx), "=f"(kr) : "r"(c_a + 8u * (uint32_t)k));
    unsigned rel = __ballot_sync(0xffffffffu, kx + kr >= lo && kx - kr <= hi);
    while (rel) {
      const int r = r0 + __ffs(rel) - 1;
      rel &= rel - 1;
      float cx, cr;
      asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(cx), "=f"(cr) : "r"(c_a + 8u * (uint32_t)r));
      const bool near = alive && fabsf(cx - cb.x) - cr <= cb.y;
      if (__any_sync(0xffffffffu, near)) {
        if (near) {
          float4 kk, b;
          float ak;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(kk.x), "=f"(kk.y), "=f"(kk.z), "=f"(kk.w) : "r"(b_a + 16u * (uint32_t)r));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ak) : "r"(a_a + 4u * (uint32_t)r));
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(own_a));
          if (suppresses(kk, ak, b, legacy_area(b), thresh)) alive = false;
        }
      }
    }
    if (!__any_sync(0xffffffffu, alive)) break;
  }
  return alive;
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads)
nms_sort_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores,
                const int32_t* __restrict__ seg_off, int max_seg_len, float4* __restrict__ sboxes,
                int32_t* __restrict__ order, int32_t* __restrict__ sorted_flag) {
  extern __shared__ u64 keys[];
  __shared__ int s_unsorted;
  const int s = blockIdx.x, tid = threadIdx.x;
  const int off = seg_off[s];
  int n = seg_off[s + 1] - off;
  if (n > max_seg_len) n = max_seg_len;
  if (n <= 0) {
    if (tid == 0) sorted_flag[s] = 1;
    return;
  }
  int npad = 2;
  while (npad < n) npad <<= 1;
  if (tid == 0) s_unsorted = 0;
  for (int i = tid; i < npad; i += kSortThreads)
    keys[i] = i < n ? ((u64)desc_score_bits(scores[off + i]) << 32) | (uint32_t)i : ~0ull;
  __syncthreads();
  for (int i = tid; i + 1 < n; i += kSortThreads)
    if (keys[i] > keys[i + 1]) s_unsorted = 1;
  __syncthreads();
  const bool unsorted = s_unsorted != 0;
  if (unsorted) {
    for (int k = 2; k <= npad; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < (npad >> 1); t += kSortThreads) {
          int lo = 2 * t - (t & (j - 1));
          int hi = lo + j;
          u64 a = keys[lo], b = keys[hi];
          bool up = (lo & k) == 0;
          if ((a > b) == up) {
            keys[lo] = b;
            keys[hi] = a;
          }
        }
        __syncthreads();
      }
    }
  }
  for (int i = tid; i < n; i += kSortThreads) {
    int src = (int)(uint32_t)keys[i];
    order[off + i] = src;
    sboxes[off + i] = boxes[off + src];
  }
  if (tid == 0) sorted_flag[s] = unsorted ? 0 : 1;
}

// ---------------------------------------------------------------------------------------
// grid (S, MB, MB): blockIdx.y = row tile i, blockIdx.z = column tile j >= i.
__global__ void __launch_bounds__(kTile)
nms_mask_kernel(const float4* __restrict__ sboxes, const int32_t* __restrict__ seg_off, int max_seg_len,
                int MB, float thresh, u64* __restrict__ mask) {
  const int s = blockIdx.x, i = blockIdx.y, j = blockIdx.z;
  if (j < i) return;
  const int off = seg_off[s];
  int n = seg_off[s + 1] - off;
  if (n > max_seg_len) n = max_seg_len;
  if (j * kTile >= n) return;
  __shared__ float4 cbox[kTile];
  __shared__ float carea[kTile];
  const int t = threadIdx.x;
  const int ncol = min(kTile, n - j * kTile);
  if (t < ncol) {
    float4 b = sboxes[off + j * kTile + t];
    cbox[t] = b;
    carea[t] = legacy_area(b);
  }
  __syncthreads();
  const int row = i * kTile + t;
  if (row >= n) return;
  const float4 a = sboxes[off + row];
  const float area_a = legacy_area(a);
  u64 bits = 0;
  for (int c = (i == j) ? t + 1 : 0; c < ncol; ++c)
    if (suppresses(a, area_a, cbox[c], carea[c], thresh)) bits |= 1ull << c;
  mask[(size_t)(off + row) * MB + j] = bits;
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSweepThreads)
nms_sweep_kernel(const u64* __restrict__ mask, const int32_t* __restrict__ order,
                 const int32_t* __restrict__ sorted_flag, const int32_t* __restrict__ seg_off,
                 int max_seg_len, int MB, long long max_keep, long long* __restrict__ keep_idx,
                 int32_t* __restrict__ keep_cnt) {
  __shared__ u64 remv[kMaxTiles];      // bit set = sorted position already suppressed
  __shared__ u64 keepbits[kMaxTiles];  // bit set = ORIGINAL index kept
  __shared__ u64 diag[2][kTile];
  __shared__ int scan[kMaxTiles];
  __shared__ u64 s_kept;
  __shared__ int s_nkept;
  const int s = blockIdx.x, tid = threadIdx.x;
  const int off = seg_off[s];
  int n = seg_off[s + 1] - off;
  if (n > max_seg_len) n = max_seg_len;
  if (n <= 0) {
    if (tid == 0) keep_cnt[s] = 0;
    return;
  }
  const int nb = (n + kTile - 1) / kTile;
  for (int w = tid; w < kMaxTiles; w += kSweepThreads) {
    remv[w] = 0;
    keepbits[w] = 0;
  }
  if (tid < kTile) diag[0][tid] = tid < n ? mask[(size_t)(off + tid) * MB] : 0;
  if (tid == 0) s_nkept = 0;
  const bool can_stop = max_keep > 0 && sorted_flag[s] != 0;
  __syncthreads();

  for (int i = 0; i < nb; ++i) {
    const int cur = i & 1;
    if (tid == 0) {
      // serial resolve of tile i: a row survives if no earlier kept row removed it
      u64 gone = remv[i], kept = 0;
      const int nrows = min(kTile, n - i * kTile);
#pragma unroll 8
      for (int r = 0; r < nrows; ++r) {
        u64 d = diag[cur][r];
        if (!((gone >> r) & 1ull)) {
          kept |= 1ull << r;
          gone |= d;
        }
      }
      s_kept = kept;
      s_nkept += __popcll(kept);
    } else if (tid >= 64 && tid < 64 + kTile && i + 1 < nb) {
      // prefetch the next tile's diagonal words while thread 0 works
      int row = (i + 1) * kTile + (tid - 64);
      diag[cur ^ 1][tid - 64] = row < n ? mask[(size_t)(off + row) * MB + (i + 1)] : 0;
    }
    __syncthreads();
    const u64 kept = s_kept;
    if (tid < kTile && ((kept >> tid) & 1ull)) {
      int o = order[off + i * kTile + tid];
      atomicOr(&keepbits[o >> 6], 1ull << (o & 63));
    }
    if (can_stop && (long long)s_nkept >= max_keep) break;  // uniform: shared values
    const int ncols = nb - (i + 1);
    if (ncols > 0 && kept != 0) {
      // OR the kept rows' words into remv[i+1 ..].  Thread (rg, jj) takes rows rg, rg+RG, ...
      // of column word jj; the loads are independent of each other, so they are all issued
      // before the first one is consumed (one memory round trip per tile, not one per row).
      const int RG = kSweepThreads / ncols;  // >= 2 since ncols <= 255
      const int rg = tid / ncols, jj = tid - rg * ncols;
      if (rg < RG) {
        const u64* base = mask + (size_t)(off + i * kTile) * MB + (i + 1) + jj;
        u64 acc = 0;
        for (int r0 = rg; r0 < kTile; r0 += 8 * RG) {
          u64 v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int r = r0 + k * RG;
            v[k] = (r < kTile && ((kept >> r) & 1ull)) ? __ldg(base + (size_t)r * MB) : 0ull;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) acc |= v[k];
        }
        if (acc) atomicOr(&remv[i + 1 + jj], acc);
      }
    }
    __syncthreads();
  }
  __syncthreads();

  // ascending compaction of the kept original indices
  const int nw = nb;  // original indices also span ceil(n/64) words
  if (tid < kMaxTiles) scan[tid] = tid < nw ? __popcll(keepbits[tid]) : 0;
  __syncthreads();
  for (int d = 1; d < kMaxTiles; d <<= 1) {
    int v = 0;
    if (tid < kMaxTiles && tid >= d) v = scan[tid - d];
    __syncthreads();
    if (tid < kMaxTiles) scan[tid] += v;
    __syncthreads();
  }
  const int total = scan[kMaxTiles - 1];
  const int limit = (max_keep > 0 && max_keep < (long long)total) ? (int)max_keep : total;
  if (tid < nw) {
    u64 bits = keepbits[tid];
    int pos = scan[tid] - __popcll(bits);
    while (bits && pos < limit) {
      int b = __ffsll((long long)bits) - 1;
      bits &= bits - 1;
      keep_idx[off + pos] = (long long)tid * 64 + b;
      ++pos;
    }
  }
  for (int p = limit + tid; p < n; p += kSweepThreads) keep_idx[off + p] = -1;
  if (tid == 0) keep_cnt[s] = limit;
}

// ---------------------------------------------------------------------------------------
// Fused single-kernel path (segments up to kFusedMaxSeg boxes): one CTA per segment keeps
// the segment's boxes in shared memory and never materialises the N x N/64 bitmask.
// Tiles of 64 boxes are visited in order; for tile i
//   (a) the 64x64 diagonal IoU words are computed on the fly (rows already removed skip),
//   (b) one thread resolves the in-tile chain,
//   (c) every still-alive LATER box is tested only against the tile's KEPT boxes and the
//       per-warp verdicts are merged with a ballot into the "removed" bit words.
// Work is sum_i kept_i * alive_later_i pair tests instead of N^2/2, and stops as soon as
// max_keep boxes are kept when the segment arrived sorted (RPN top-k output does).
// ---------------------------------------------------------------------------------------
constexpr int kFusedMaxSeg = 12288;  // 16 B * n of dynamic shared memory + kept list next to ~8 KB static
constexpr int kChunkBoxes = 1024;    // lazy-update granularity of the early-stopping sweep

// phase timers of the fused kernel (build with -DB200_NMS_STATS; scripts/probe_nms_step.py): cycles of thread 0
#ifdef B200_NMS_STATS
__device__ long long g_nms_stats[256][16];
#define NMS_T(slot)                                 \
  do {                                              \
    if (tid == 0) {                                 \
      const long long t_now = clock64();            \
      g_nms_stats[blockIdx.x][slot] += t_now - t_last; \
      t_last = t_now;                               \
    }                                               \
  } while (0)
#else
#define NMS_T(slot) do {} while (0)
#endif

template <int kThreads>
__global__ void __launch_bounds__(kThreads, 1)
nms_fused_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores,
                 const int32_t* __restrict__ seg_off, int max_seg_len, float thresh, long long max_keep,
                 int kl_offset_boxes, int kl_capacity, int32_t* __restrict__ order,
                 long long* __restrict__ keep_idx, int32_t* __restrict__ keep_cnt) {
  extern __shared__ __align__(16) unsigned char fused_smem[];
  float4* sb = reinterpret_cast<float4*>(fused_smem);  // boxes in visiting order
  u64* keys = reinterpret_cast<u64*>(fused_smem);      // aliases sb during the sort
  __shared__ uint32_t remv[B200_NMS_MAX_SEG / 32];
  __shared__ u64 keepbits[kMaxTiles];
  __shared__ u64 diag[2][kTile];  // double-buffered: tile i+1's words are computed during tile i's chain
  __shared__ float4 kb[kTile];
  __shared__ float ka[kTile];
  __shared__ float2 kbc[kTile];  // (x centre, x reach) of the tile's kept boxes
  __shared__ int scan[kMaxTiles];
  __shared__ u64 s_kept;
  __shared__ int s_nkept, s_unsorted;

  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int off = seg_off[s];
  int n = seg_off[s + 1] - off;
  if (n > max_seg_len) n = max_seg_len;
  if (n <= 0) {
    if (tid == 0) keep_cnt[s] = 0;
    return;
  }
  const int nb = (n + kTile - 1) / kTile;
#ifdef B200_NMS_STATS
  long long t_last = clock64();
  if (tid < 16) g_nms_stats[blockIdx.x][tid] = 0;
  __syncthreads();
#endif

  // ---- visiting order -------------------------------------------------------------------
  if (tid == 0) {
    s_unsorted = 0;
    s_nkept = 0;
  }
  for (int w = tid; w < nb * 2; w += kThreads) remv[w] = 0;
  for (int w = tid; w < nb; w += kThreads) keepbits[w] = 0;
  __syncthreads();
  for (int i = tid; i + 1 < n; i += kThreads) {
    const u64 a = ((u64)desc_score_bits(scores[off + i]) << 32) | (uint32_t)i;
    const u64 b = ((u64)desc_score_bits(scores[off + i + 1]) << 32) | (uint32_t)(i + 1);
    if (a > b) s_unsorted = 1;
  }
  __syncthreads();
  const bool unsorted = s_unsorted != 0;
  if (unsorted) {
    int npad = 2;
    while (npad < n) npad <<= 1;
    for (int i = tid; i < npad; i += kThreads)
      keys[i] = i < n ? ((u64)desc_score_bits(scores[off + i]) << 32) | (uint32_t)i : ~0ull;
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < (npad >> 1); t += kThreads) {
          const int lo = 2 * t - (t & (j - 1)), hi = lo + j;
          const u64 a = keys[lo], b = keys[hi];
          if ((a > b) == ((lo & k) == 0)) {
            keys[lo] = b;
            keys[hi] = a;
          }
        }
        __syncthreads();
      }
    }
    for (int i = tid; i < n; i += kThreads) order[off + i] = (int)(uint32_t)keys[i];
    __syncthreads();  // keys are dead from here; the same bytes become sb
    for (int i = tid; i < n; i += kThreads) sb[i] = boxes[off + order[off + i]];
  } else {
    for (int i = tid; i < n; i += kThreads) sb[i] = boxes[off + i];
  }
  const bool can_stop = max_keep > 0 && !unsorted;
  __syncthreads();
  NMS_T(0);  // load / order

  constexpr int G = kThreads / kTile;  // threads cooperating on one diagonal row
  constexpr int CPT = kTile / G;       // columns per thread
  // When the sweep may stop early (sorted input + max_keep), later boxes are only brought up
  // to date in chunks of kChunkBoxes: a chunk is first tested against the list of ALL boxes
  // kept so far (kl), then swept tile by tile with step (c) confined to the chunk.  Boxes
  // beyond the chunk in which max_keep is reached are never touched.
  float4* kl = reinterpret_cast<float4*>(fused_smem + sizeof(float4) * (size_t)kl_offset_boxes);
  float2* klc = reinterpret_cast<float2*>(kl + kl_capacity);  // (x centre, x reach) per kept box
  float* kla = reinterpret_cast<float*>(klc + kl_capacity);
  // In that mode the boxes of a chunk are handed to the threads in order of their x centre
  // (perm): the 32 lanes of a warp then sit side by side in the image, and a kept box that is
  // out of x reach of all of them (see x_reach) is rejected with 5 warp-uniform instructions
  // instead of the 25 of the IoU test.  Verdicts go to the "removed" bits with atomicOr.
  u64* pkeys = reinterpret_cast<u64*>(kla + kl_capacity + (kl_capacity & 1));
  int* perm = reinterpret_cast<int*>(pkeys + kChunkBoxes);
  constexpr int kPer = (kChunkBoxes + kThreads - 1) / kThreads;  // chunk boxes per thread
  // per-thread state of the chunk's boxes this thread owns (early-stopping mode): position, alive, x centre / reach
  int own_p[kPer];
  bool own_alive[kPer];
  float2 own_c[kPer];
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    own_p[k] = 0;
    own_alive[k] = false;
    own_c[k] = make_float2(0.f, 0.f);
  }
  int nk = 0;  // boxes in kl (uniform)
  bool done = false;
  int diag_tile = -1;  // tile whose diagonal words sit in diag[diag_tile & 1] (uniform)
  // work item (row r, column group cg) of a tile's 64 x 64 diagonal block: CPT columns > r, the G
  // lanes of a row OR their bits together (they are consecutive lanes of one warp)
  auto diag_item = [&](int tbase, int trows, u64 tgone, int item, u64* dbuf) {
    const int r = item / G, cg = item % G;
    u64 bits = 0;
    if (r < trows && !((tgone >> r) & 1ull)) {
      const float4 a = sb[tbase + r];
      const float area_a = legacy_area(a);
#pragma unroll 4
      for (int k = 0; k < CPT; ++k) {
        const int c = cg * CPT + k;
        if (c > r && c < trows) {
          const float4 b = sb[tbase + c];
          if (suppresses(a, area_a, b, legacy_area(b), thresh)) bits |= 1ull << c;
        }
      }
    }
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, d);
    if (cg == 0) dbuf[r] = bits;
  };
  for (int c0 = 0, c1 = 0; c0 < n && !done; c0 = c1) {
    // chunk length: everything when the sweep cannot stop early; else up to kChunkBoxes, but only about as many
    // boxes as the sweep can still need (1.5 x the boxes still to be kept + a tile): the work of a chunk --
    // the pass against the kept list and step (c) -- grows with its length, and boxes behind the stopping
    // point are wasted (s_nkept is uniform here: read between barriers)
    int chunk_boxes = nb * kTile;
    if (can_stop) {
      const long long rem = max_keep - (long long)s_nkept;
      const long long want = (rem + rem / 2 + 2 * kTile - 1) / kTile * kTile;
      chunk_boxes = (int)(want < (long long)kChunkBoxes ? (want > 2 * kTile ? want : 2 * kTile) : kChunkBoxes);
    }
    c1 = min(n, c0 + chunk_boxes);
    if (can_stop) {
      // this chunk's boxes in x-centre order: a counting sort into 256 x buckets between the chunk's
      // smallest and largest centre (histogram with shared atomics, one scan, one scatter: 4 barriers; the
      // bitonic sort of the 1024 keys it replaces took 55).  The order inside a bucket is whatever the
      // atomics give -- it only decides which lane tests which box, never a verdict.
      const int clen = c1 - c0;
      int* bcnt = reinterpret_cast<int*>(pkeys);  // 256 counters, then their exclusive scan (pkeys is free here)
      int* brange = bcnt + 256;                    // orderable(min cx), orderable(max cx)
      if (tid < 256) bcnt[tid] = 0;
      if (tid == 0) {
        brange[0] = 0x7fffffff;
        brange[1] = (int)0x80000000;
      }
      __syncthreads();
      float cxs[kPer];
      {
        int mn = 0x7fffffff, mx = (int)0x80000000;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
          const int t = tid + k * kThreads;
          cxs[k] = 0.f;
          if (t < clen) {
            const float4 b = sb[c0 + t];
            cxs[k] = 0.5f * (b.x + b.z);
            const int o = (int)(orderable_f(cxs[k]) ^ 0x80000000u);  // signed-orderable
            mn = min(mn, o);
            mx = max(mx, o);
          }
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) {
          atomicMin(&brange[0], mn);
          atomicMax(&brange[1], mx);
        }
      }
      __syncthreads();
      int bkt[kPer], slot[kPer];
      {
        const uint32_t umn = (uint32_t)brange[0] ^ 0x80000000u, umx = (uint32_t)brange[1] ^ 0x80000000u;
        const float lo = __uint_as_float((umn & 0x80000000u) ? (umn & 0x7fffffffu) : ~umn);
        const float hi = __uint_as_float((umx & 0x80000000u) ? (umx & 0x7fffffffu) : ~umx);
        const float scale = hi > lo ? 255.999f / (hi - lo) : 0.f;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
          const int t = tid + k * kThreads;
          bkt[k] = -1;
          if (t < clen) {
            bkt[k] = min(255, max(0, (int)((cxs[k] - lo) * scale)));
            slot[k] = atomicAdd(&bcnt[bkt[k]], 1);
          }
        }
      }
      __syncthreads();
      if (tid < 32) {  // exclusive scan of the 256 counters by one warp (8 per lane)
        int c[8], run = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          c[j] = bcnt[8 * tid + j];
          run += c[j];
        }
        int incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += v;
        }
        int ex = incl - run;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          bcnt[8 * tid + j] = ex;
          ex += c[j];
        }
      }
      for (int t = clen + tid; t < kChunkBoxes; t += kThreads) perm[t] = -1;  // padding
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kPer; ++k)
        if (bkt[k] >= 0) perm[bcnt[bkt[k]] + slot[k]] = tid + k * kThreads;
      __syncthreads();
      NMS_T(1);  // chunk x order
      // the boxes this thread owns for the whole chunk (x order: the lanes of a warp are neighbours in the image)
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int pl = perm[tid + k * kThreads];
        own_p[k] = c0 + (pl >= 0 ? pl : 0);
        own_alive[k] = pl >= 0 && !((remv[own_p[k] >> 5] >> (own_p[k] & 31)) & 1u);
        own_c[k] = x_reach(sb[own_p[k]], thresh);
      }
      if (c0 > 0 && nk > 0) {
        // chunk vs everything kept so far
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
          const bool alive = survives_list(b200::smem_u32(klc), b200::smem_u32(kl), b200::smem_u32(kla), nk, own_alive[k],
                                           b200::smem_u32(sb + own_p[k]), own_c[k], thresh, lane);
          if (own_alive[k] && !alive) atomicOr(&remv[own_p[k] >> 5], 1u << (own_p[k] & 31));
          own_alive[k] = alive;
        }
        __syncthreads();
        NMS_T(2);  // chunk vs kept list
      }
    } else if (c0 > 0 && nk > 0) {
      // chunk vs everything kept so far
      for (int j0 = c0 + (tid & ~31); j0 < c1; j0 += kThreads) {
        const int j = j0 + lane;
        // warp-uniform trip count, predicated body: the lanes of a warp stay converged
        // (a per-lane `break` made ptxas serialise the 32 lanes of this loop)
        const uint32_t dead = remv[j0 >> 5];
        bool alive = j < c1 && !((dead >> lane) & 1u);
        const float4 b = sb[alive ? j : c0];
        const float area_b = legacy_area(b);
        for (int r = 0; r < nk; ++r)
          if (alive && suppresses(kl[r], kla[r], b, area_b, thresh)) alive = false;
        const uint32_t v = __ballot_sync(0xffffffffu, j < c1 && !((dead >> lane) & 1u) && !alive);
        if (lane == 0 && v) remv[j0 >> 5] = dead | v;
      }
      __syncthreads();
    }
    for (int i = c0 / kTile; i * kTile < c1; ++i) {
      const int base = i * kTile;
      const int nrows = min(kTile, n - base);
      const u64 gone0 = ((u64)remv[2 * i + 1] << 32) | remv[2 * i];
      const u64 live_mask = nrows == 64 ? ~0ull : ((1ull << nrows) - 1);
      if ((gone0 & live_mask) == live_mask) continue;  // whole tile already removed (uniform)

      // (a) diagonal words of tile i -- unless they were already produced during the chain of
      // the previous tile (below)
      if (diag_tile != i) {
        diag_item(base, nrows, gone0, tid, diag[i & 1]);
        __syncthreads();
      }
      NMS_T(3);  // diagonal words not precomputed
      // (b) in-tile chain on thread 0; meanwhile warps >= 1 compute the diagonal words of tile
      // i+1 (pure pairwise facts: rows that turn out to be removed are simply not used)
      const int nbase = base + kTile;
      const bool pre = nbase < n;  // uniform
      if (tid < 32) {
        // The greedy chain "row r is kept iff it is alive and no kept row before it suppresses it" is the unique
        // fixed point of K -> alive & ~OR{d[j] : j in K} (d[j] only has bits above j, so bit r of the image
        // depends on bits below r only: after t rounds the lowest t bits are final).  Warp 0 iterates it with
        // two rows per lane and a 64-bit OR reduction per round; with few suppressions inside a tile it
        // settles in 2-4 rounds instead of 64 dependent steps of one thread.
        const u64* dg = diag[i & 1];
        const u64 alive0 = ~gone0 & live_mask;
        const u64 d0 = lane < nrows ? dg[lane] : 0ull, d1 = lane + 32 < nrows ? dg[lane + 32] : 0ull;
        u64 kept = alive0;
        for (int round = 0; round < kTile; ++round) {
          const u64 mine = (((kept >> lane) & 1ull) ? d0 : 0ull) | (((kept >> (lane + 32)) & 1ull) ? d1 : 0ull);
          const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)mine);
          const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(mine >> 32));
          const u64 next = alive0 & ~(((u64)hi << 32) | lo);
          if (next == kept) break;
          kept = next;
        }
        if (tid == 0) {
          s_kept = kept;
          s_nkept += __popcll(kept);
        }
      } else if (pre) {
        const int nrows2 = min(kTile, n - nbase);
        for (int item = tid - 32; item < kThreads; item += kThreads - 32)
          diag_item(nbase, nrows2, 0ull, item, diag[(i + 1) & 1]);
      }
      if (pre) diag_tile = i + 1;
      NMS_T(4);  // chain (thread 0)
      __syncthreads();
      NMS_T(5);  // ... waiting for the next tile's diagonal words
      const u64 kept = s_kept;
      const int m = __popcll(kept);
      if (tid < kTile && ((kept >> tid) & 1ull)) {
        const int pos = __popcll(kept & ((1ull << tid) - 1));
        const float4 b = sb[base + tid];
        const float ab = legacy_area(b);
        kb[pos] = b;
        ka[pos] = ab;
        kbc[pos] = x_reach(b, thresh);
        if (can_stop && nk + pos < kl_capacity) {
          kl[nk + pos] = b;
          kla[nk + pos] = ab;
          klc[nk + pos] = x_reach(b, thresh);
        }
        // (original positions; a 64-bit shared atomicOr is a CAS loop -- 64 threads on one word took 5.8 k cycles
        // per tile -- so sorted input stores the tile's word directly and unsorted input uses 32-bit atomics)
        if (unsorted) {
          const int o = order[off + base + tid];
          atomicOr(reinterpret_cast<uint32_t*>(keepbits) + (o >> 5), 1u << (o & 31));
        }
      }
      if (!unsorted && tid == 0) keepbits[i] = kept;
      nk += m;
      if (can_stop && (long long)s_nkept >= max_keep) {  // uniform: shared value
        done = true;
        break;
      }
      __syncthreads();
      NMS_T(6);  // kept list update
      // (c) later boxes of this chunk vs this tile's kept boxes
      if (can_stop) {
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
          const bool cand = own_alive[k] && own_p[k] >= base + kTile;
          const bool alive = survives_list(b200::smem_u32(kbc), b200::smem_u32(kb), b200::smem_u32(ka), m, cand,
                                           b200::smem_u32(sb + own_p[k]), own_c[k], thresh, lane);
          if (cand && !alive) {
            atomicOr(&remv[own_p[k] >> 5], 1u << (own_p[k] & 31));
            own_alive[k] = false;
          }
        }
      } else
      for (int j0 = base + kTile + (tid & ~31); j0 < c1; j0 += kThreads) {
        const int j = j0 + lane;
        const uint32_t dead = remv[j0 >> 5];
        const bool cand = j < c1 && !((dead >> lane) & 1u);
        bool alive = cand;
        const float4 b = sb[cand ? j : base];
        const float area_b = legacy_area(b);
        for (int r = 0; r < m; ++r)  // warp-uniform trip count, predicated body
          if (alive && suppresses(kb[r], ka[r], b, area_b, thresh)) alive = false;
        const uint32_t v = __ballot_sync(0xffffffffu, cand && !alive);
        if (lane == 0 && v) remv[j0 >> 5] = dead | v;
      }
      __syncthreads();
      NMS_T(7);  // (c)
#ifdef B200_NMS_STATS
      if (tid == 0) g_nms_stats[blockIdx.x][9] += 1;
#endif
    }
  }
  __syncthreads();

  // ---- ascending compaction of the kept original indices --------------------------------
  for (int w = tid; w < kMaxTiles; w += kThreads) scan[w] = w < nb ? __popcll(keepbits[w]) : 0;
  __syncthreads();
  for (int d = 1; d < kMaxTiles; d <<= 1) {
    int v[(kMaxTiles + kThreads - 1) / kThreads];
    int k = 0;
    for (int w = tid; w < kMaxTiles; w += kThreads) v[k++] = w >= d ? scan[w - d] : 0;
    __syncthreads();
    k = 0;
    for (int w = tid; w < kMaxTiles; w += kThreads) scan[w] += v[k++];
    __syncthreads();
  }
  const int total = scan[kMaxTiles - 1];
  const int limit = (max_keep > 0 && max_keep < (long long)total) ? (int)max_keep : total;
  for (int w = tid; w < nb; w += kThreads) {
    u64 bits = keepbits[w];
    int pos = scan[w] - __popcll(bits);
    while (bits && pos < limit) {
      const int b = __ffsll((long long)bits) - 1;
      bits &= bits - 1;
      keep_idx[off + pos] = (long long)w * 64 + b;
      ++pos;
    }
  }
  for (int p = limit + tid; p < n; p += kThreads) keep_idx[off + p] = -1;
  if (tid == 0) keep_cnt[s] = limit;
  NMS_T(8);  // compaction
}

bool g_nms_force_bitmask = false;

// dynamic shared memory of the fused kernel: boxes, and for the early-stopping mode the kept list
// (box, area, x centre/reach) plus the per-chunk x-order (sort keys + permutation)
size_t fused_smem_bytes(size_t n_cap, size_t kl_cap) {
  size_t b = sizeof(float4) * n_cap;
  if (kl_cap > 0)
    b += (sizeof(float4) + sizeof(float2) + sizeof(float)) * kl_cap + sizeof(float) * (kl_cap & 1) +
         (sizeof(u64) + sizeof(int)) * (size_t)kChunkBoxes;
  return b;
}

// the fused kernel needs the segment's boxes (+ the kept list) in one CTA's shared memory
bool fused_applies(int64_t max_seg_len, int64_t max_keep) {
  if (g_nms_force_bitmask || max_seg_len > kFusedMaxSeg) return false;
  const size_t n_cap = (size_t)b200::ceil_div<int64_t>(max_seg_len, kTile) * kTile;
  const size_t kl_cap = max_keep > 0 ? (size_t)(max_keep < max_seg_len ? max_keep : max_seg_len) + kTile : 0;
  return fused_smem_bytes(n_cap, kl_cap) + 9 * 1024 <= 227 * 1024;
}

struct Workspace {
  float4* sboxes;
  int32_t* order;
  int32_t* sorted_flag;
  u64* mask;
  size_t bytes;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Workspace carve(void* base, int64_t n_total, int64_t n_segments, int64_t max_seg_len) {
  Workspace w;
  size_t MB = (size_t)b200::ceil_div<int64_t>(max_seg_len > 0 ? max_seg_len : 1, kTile);
  size_t o = 0;
  char* p = static_cast<char*>(base);
  w.sboxes = reinterpret_cast<float4*>(p + o);
  o = align_up(o + sizeof(float4) * (size_t)n_total, 256);
  w.order = reinterpret_cast<int32_t*>(p + o);
  o = align_up(o + sizeof(int32_t) * (size_t)n_total, 256);
  w.sorted_flag = reinterpret_cast<int32_t*>(p + o);
  o = align_up(o + sizeof(int32_t) * (size_t)n_segments, 256);
  w.mask = reinterpret_cast<u64*>(p + o);
  // the bitmask exists only on the three-kernel path
  o = align_up(o + sizeof(u64) * (size_t)n_total * MB, 256);
  w.bytes = o;
  return w;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// Cross-level proposal selection: per image, the top_n highest-scoring boxes among the boxes
// its NMS segments kept (reference modeling/rpn/inference.py:173-180, select_over_all_levels
// in test mode), emitted directly in RoI format (batch, x1, y1, x2, y2) for the pooler.
// One CTA per image: gather (score, global index) keys of the kept boxes, bitonic sort in
// shared memory, write the first top_n.  Equal scores keep ascending global index order.
// ---------------------------------------------------------------------------------------
constexpr int kSelectThreads = 1024;
constexpr int kSelectMax = 16384;

__global__ void __launch_bounds__(kSelectThreads)
select_topk_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores,
                   const int32_t* __restrict__ seg_off, const long long* __restrict__ keep_idx,
                   const int32_t* __restrict__ keep_cnt, int segs_per_image, int top_n, int npad_max,
                   float* __restrict__ rois_out, float* __restrict__ scores_out, int32_t* __restrict__ count_out) {
  extern __shared__ u64 skeys[];
  __shared__ int s_base[65];
  const int img = blockIdx.x, tid = threadIdx.x;
  const int s0 = img * segs_per_image;
  if (tid == 0) {
    int acc = 0;
    for (int l = 0; l < segs_per_image; ++l) {
      s_base[l] = acc;
      acc += keep_cnt[s0 + l];
    }
    s_base[segs_per_image] = acc < npad_max ? acc : npad_max;
  }
  __syncthreads();
  const int total = s_base[segs_per_image];
  int npad = 2;
  while (npad < total) npad <<= 1;
  for (int i = tid; i < npad; i += kSelectThreads) skeys[i] = ~0ull;
  __syncthreads();
  for (int l = 0; l < segs_per_image; ++l) {
    const int off = seg_off[s0 + l], base = s_base[l];
    const int cnt = min(keep_cnt[s0 + l], total - base);
    for (int j = tid; j < cnt; j += kSelectThreads) {
      const int gi = off + (int)keep_idx[off + j];
      skeys[base + j] = ((u64)desc_score_bits(scores[gi]) << 32) | (uint32_t)gi;
    }
  }
  __syncthreads();
  // The kept list of a segment that arrived sorted (the RPN's top-k output) is itself in key
  // order: then no sort is needed -- an entry's global rank is its index in its own run plus,
  // for every other run, the number of smaller keys (binary search; keys are unique), and the
  // entry goes straight to its output row.
  int unsorted = 0;
  for (int l = 0; l < segs_per_image; ++l) {
    const int base = s_base[l], cnt = min(s_base[l + 1], total) - base;
    for (int j = tid; j + 1 < cnt; j += kSelectThreads) unsorted |= skeys[base + j] > skeys[base + j + 1];
  }
  const int n_out = min(total, top_n);
  if (!__syncthreads_or(unsorted)) {
    for (int l = 0; l < segs_per_image; ++l) {
      const int base = s_base[l], cnt = min(s_base[l + 1], total) - base;
      for (int j = tid; j < cnt; j += kSelectThreads) {
        const u64 key = skeys[base + j];
        int rank = j;
        for (int o = 0; o < segs_per_image; ++o) {
          if (o == l) continue;
          int lo = s_base[o], hi = min(s_base[o + 1], total);  // first index in [lo, hi) with key' > key
          const int lo0 = lo;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (skeys[mid] < key) lo = mid + 1;
            else hi = mid;
          }
          rank += lo - lo0;
        }
        if (rank < top_n) {
          const int gi = (int)(uint32_t)key;
          const float4 b = boxes[gi];
          float* r = rois_out + ((size_t)img * top_n + rank) * 5;
          r[0] = (float)img;
          r[1] = b.x;
          r[2] = b.y;
          r[3] = b.z;
          r[4] = b.w;
          if (scores_out) scores_out[(size_t)img * top_n + rank] = scores[gi];
        }
      }
    }
    for (int i = n_out + tid; i < top_n; i += kSelectThreads) {
      float* r = rois_out + ((size_t)img * top_n + i) * 5;
      r[0] = (float)img;
      r[1] = r[2] = r[3] = r[4] = 0.f;
      if (scores_out) scores_out[(size_t)img * top_n + i] = 0.f;
    }
    if (tid == 0) count_out[img] = n_out;
    return;
  }
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (npad >> 1); t += kSelectThreads) {
        const int lo = 2 * t - (t & (j - 1)), hi = lo + j;
        const u64 a = skeys[lo], b = skeys[hi];
        if ((a > b) == ((lo & k) == 0)) {
          skeys[lo] = b;
          skeys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < top_n; i += kSelectThreads) {
    float* r = rois_out + ((size_t)img * top_n + i) * 5;
    if (i < n_out) {
      const int gi = (int)(uint32_t)skeys[i];
      const float4 b = boxes[gi];
      r[0] = (float)img;
      r[1] = b.x;
      r[2] = b.y;
      r[3] = b.z;
      r[4] = b.w;
      if (scores_out) scores_out[(size_t)img * top_n + i] = scores[gi];
    } else {
      r[0] = (float)img;
      r[1] = r[2] = r[3] = r[4] = 0.f;
      if (scores_out) scores_out[(size_t)img * top_n + i] = 0.f;
    }
  }
  if (tid == 0) count_out[img] = n_out;
}

#ifdef B200_NMS_STATS
extern "C" int b200_debug_nms_stats(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_nms_stats, sizeof(g_nms_stats)) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" void b200_debug_nms(int force_bitmask) { g_nms_force_bitmask = force_bitmask != 0; }

extern "C" size_t b200_nms_workspace_bytes(int64_t n_total, int64_t n_segments, int64_t max_seg_len) {
  if (n_total < 0 || n_segments < 0 || max_seg_len < 0) return 0;
  return carve(nullptr, n_total, n_segments, max_seg_len).bytes;
}

extern "C" int b200_nms_batched(const float* boxes, const float* scores, const int32_t* seg_offsets,
                                int64_t n_total, int64_t n_segments, int64_t max_seg_len, float thresh,
                                int64_t max_keep, int64_t* keep_idx, int32_t* keep_cnt, void* workspace,
                                size_t workspace_bytes, void* stream) {
  using namespace b200;
  B200_REQUIRE(n_total >= 0 && n_segments >= 0 && max_seg_len >= 0, "nms: negative size");
  if (n_segments == 0) return B200_OK;
  B200_REQUIRE(seg_offsets && keep_cnt, "nms: null seg_offsets / keep_cnt");
  B200_REQUIRE(n_total == 0 || (boxes && scores && keep_idx && workspace), "nms: null pointer");
  B200_REQUIRE(n_total == 0 || aligned16(boxes), "nms: boxes must be 16-byte aligned");
  if (max_seg_len > B200_NMS_MAX_SEG) {
    set_error("nms: max_seg_len %lld exceeds B200_NMS_MAX_SEG (%d)", (long long)max_seg_len, B200_NMS_MAX_SEG);
    return B200_ERR_UNSUPPORTED;
  }
  B200_REQUIRE(n_total < (int64_t)1 << 31, "nms: n_total must fit int32");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (max_seg_len == 0) max_seg_len = 1;
  Workspace w = carve(workspace, n_total, n_segments, max_seg_len);
  if (w.bytes > workspace_bytes) {
    set_error("nms: workspace %zu bytes < required %zu", workspace_bytes, w.bytes);
    return B200_ERR_WORKSPACE;
  }
  const int64_t kMaxGridX = 2147483647LL;
  B200_REQUIRE(n_segments <= kMaxGridX, "nms: too many segments");
  const int n_cap = (int)ceil_div<int64_t>(max_seg_len, kTile) * kTile;
  // kept-box list for the early-stopping sweep (only when max_keep is given)
  const int kl_cap = max_keep > 0 ? (int)(max_keep < max_seg_len ? max_keep : max_seg_len) + kTile : 0;
  const size_t smem = fused_smem_bytes((size_t)n_cap, (size_t)kl_cap);
  if (fused_applies(max_seg_len, max_keep)) {
    if (max_seg_len <= 1024) {
      auto kern = nms_fused_kernel<256>;
      static SmemHighWater hw;
      int rc = ensure_dynamic_smem(kern, smem, &hw, "nms: cudaFuncSetAttribute");
      if (rc != B200_OK) return rc;
      kern<<<(unsigned)n_segments, 256, smem, st>>>(reinterpret_cast<const float4*>(boxes), scores, seg_offsets,
                                                    (int)max_seg_len, thresh, (long long)max_keep, n_cap, kl_cap,
                                                    w.order, reinterpret_cast<long long*>(keep_idx), keep_cnt);
    } else {
      auto kern = nms_fused_kernel<1024>;
      static SmemHighWater hw;
      int rc = ensure_dynamic_smem(kern, smem, &hw, "nms: cudaFuncSetAttribute");
      if (rc != B200_OK) return rc;
      kern<<<(unsigned)n_segments, 1024, smem, st>>>(reinterpret_cast<const float4*>(boxes), scores, seg_offsets,
                                                     (int)max_seg_len, thresh, (long long)max_keep, n_cap, kl_cap,
                                                     w.order, reinterpret_cast<long long*>(keep_idx), keep_cnt);
    }
    B200_CHECK_LAUNCH("nms_fused_kernel");
    return B200_OK;
  }
  const int MB = (int)ceil_div<int64_t>(max_seg_len, kTile);
  int npad = 2;
  while (npad < max_seg_len) npad <<= 1;
  const size_t sort_smem = sizeof(u64) * (size_t)npad;
  static bool attr_set = false;
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(u64) * B200_NMS_MAX_SEG)),
                        "nms: cudaFuncSetAttribute");
    if (rc != B200_OK) return rc;
    attr_set = true;
  }
  nms_sort_kernel<<<(unsigned)n_segments, kSortThreads, sort_smem, st>>>(
      reinterpret_cast<const float4*>(boxes), scores, seg_offsets, (int)max_seg_len, w.sboxes, w.order,
      w.sorted_flag);
  B200_CHECK_LAUNCH("nms_sort_kernel");
  nms_mask_kernel<<<dim3((unsigned)n_segments, MB, MB), kTile, 0, st>>>(w.sboxes, seg_offsets, (int)max_seg_len,
                                                                        MB, thresh, w.mask);
  B200_CHECK_LAUNCH("nms_mask_kernel");
  nms_sweep_kernel<<<(unsigned)n_segments, kSweepThreads, 0, st>>>(
      w.mask, w.order, w.sorted_flag, seg_offsets, (int)max_seg_len, MB, (long long)max_keep,
      reinterpret_cast<long long*>(keep_idx), keep_cnt);
  B200_CHECK_LAUNCH("nms_sweep_kernel");
  return B200_OK;
}

extern "C" int b200_select_topk(const float* boxes, const float* scores, const int32_t* seg_offsets,
                                const int64_t* keep_idx, const int32_t* keep_cnt, int n_images, int segs_per_image,
                                int64_t max_kept_per_image, int top_n, float* rois_out, float* scores_out,
                                int32_t* count_out, void* stream) {
  using namespace b200;
  B200_REQUIRE(n_images >= 0 && segs_per_image >= 1 && segs_per_image <= 64 && top_n >= 1 && max_kept_per_image >= 0,
               "select_topk: bad shape");
  if (n_images == 0) return B200_OK;
  B200_REQUIRE(boxes && scores && seg_offsets && keep_idx && keep_cnt && rois_out && count_out,
               "select_topk: null pointer");
  B200_REQUIRE(aligned16(boxes), "select_topk: boxes must be 16-byte aligned");
  if (max_kept_per_image > kSelectMax) {
    set_error("select_topk: %lld kept boxes per image exceed %d", (long long)max_kept_per_image, kSelectMax);
    return B200_ERR_UNSUPPORTED;
  }
  int npad = 2;
  while (npad < max_kept_per_image) npad <<= 1;
  const size_t smem = sizeof(u64) * (size_t)npad;
  static SmemHighWater hw;
  int rc = ensure_dynamic_smem(select_topk_kernel, smem, &hw, "select_topk: smem attribute");
  if (rc != B200_OK) return rc;
  select_topk_kernel<<<n_images, kSelectThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(boxes), scores, seg_offsets, reinterpret_cast<const long long*>(keep_idx),
      keep_cnt, segs_per_image, top_n, npad, rois_out, scores_out, count_out);
  B200_CHECK_LAUNCH("select_topk_kernel");
  return B200_OK;
}
