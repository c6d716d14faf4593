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
// reach() is padded (1 % + 0.01 px) against fp32 rounding; thresh <= 0 disables the test.  The same holds
// in y with the heights.
// the same necessary condition in both axes: (x centre, x reach, y centre, y reach)
__device__ __forceinline__ float4 xy_reach(const float4 b, float thresh) {
  const float w = b.z - b.x + 1.0f, h = b.w - b.y + 1.0f;
  const float k = (1.0f - thresh) * 0.5f * 1.01f;
  const bool on = thresh > 0.0f;
  return make_float4(0.5f * (b.x + b.z), on ? k * w + 0.01f : 3.0e38f, 0.5f * (b.y + b.w), on ? k * h + 0.01f : 3.0e38f);
}
__device__ __forceinline__ uint32_t orderable_f(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
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
// Tiles of 64 boxes are visited in order.  Two sweeps:
//   PUSH (no early stop, or unsorted input): for tile i (a) the 64x64 diagonal IoU words (rows
//     already removed skip), (b) the in-tile chain on warp 0, (c) every still-alive LATER box is
//     tested against the tile's KEPT boxes, warp verdicts merged with a ballot into the "removed"
//     words.  Work is sum_i kept_i * alive_later_i pair tests instead of N^2/2.
//   PULL (sorted input + max_keep: RPN top-k output): a tile is tested against the boxes kept so
//     far -- held as a list sorted by x centre, of which a box only sees the window its x reach
//     allows, with a y test before the IoU -- then chained and merged into the list; nothing behind
//     the tile in which max_keep is reached is touched (details at the loop).
// ---------------------------------------------------------------------------------------
constexpr int kFusedMaxSeg = 12288;  // 16 B * n of dynamic shared memory + kept list next to ~9.5 KB static

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

  constexpr int G = kThreads / kTile;  // threads cooperating on one row of a tile (consecutive lanes of one warp)
  constexpr int CPT = kTile / G;       // columns per thread
  // work item (row r, column group cg) of a tile's 64 x 64 diagonal block: CPT columns > r, the G
  // lanes of a row OR their bits together
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
  // The greedy chain "row r is kept iff it is alive and no kept row before it suppresses it" is the unique
  // fixed point of K -> alive & ~OR{d[j] : j in K} (d[j] only has bits above j, so bit r of the image
  // depends on bits below r only: after t rounds the lowest t bits are final).  Warp 0 iterates it with
  // two rows per lane and a 64-bit OR reduction per round; with few suppressions inside a tile it
  // settles in 2-4 rounds instead of 64 dependent steps of one thread.
  auto chain = [&](const u64* dg, int nrows, u64 alive0) -> u64 {
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
    return kept;
  };

  if (can_stop) {
    // ---- early-stopping sweep (sorted input + max_keep: the RPN case): PULL ----------------------------
    // A box is kept iff no box kept BEFORE it suppresses it (nms_cpu.cpp:36-61 read from the side of the later
    // box), so a tile only has to be tested against the boxes kept so far, and nothing behind the tile in which
    // max_keep is reached is ever touched.  The kept boxes are held as a list of (x centre, position) SORTED BY
    // x CENTRE: IoU >= t needs |cx_a - cx_b| <= reach_a + reach_b (xy_reach), so a box only looks at the window
    // [cx - reach - rmax, cx + reach + rmax] of that list (two binary searches; rmax = largest reach kept so
    // far) -- a few per cent of it -- with the G lanes of its row striding over the window.  Per tile:
    //   A  pull (row r vs its window of the kept list) and the tile's 64 x 64 diagonal words, all threads;
    //   B  warp 0: the in-tile chain from the rows that survived the pull;
    //   C  the tile's newly kept boxes ranked among themselves by (x centre, position);
    //   D  merge into the sorted list (every old entry moves up by the number of new entries before it).
    // Boxes with NaN centre / reach never suppress anything (every comparison of nms_cpu fails) and are not
    // listed.  Both filters are necessary conditions only (+1 px against fp32 rounding): the verdicts are
    // those of testing every pair.
    // list entry: (x centre, x reach, y centre, y reach) + the box's position, two buffers (the merge copies)
    float4* xs_a = reinterpret_cast<float4*>(fused_smem + sizeof(float4) * (size_t)kl_offset_boxes);
    float4* xs_b = xs_a + kl_capacity;
    int* xp_a = reinterpret_cast<int*>(xs_b + kl_capacity);
    int* xp_b = xp_a + kl_capacity;
    __shared__ float4 rowc[kTile];  // (x centre, x reach, y centre, y reach) of the tile's rows
    __shared__ int newk[kTile];     // the tile's listed kept rows, sorted by (x centre, row)
    __shared__ uint32_t s_gone[2];
    __shared__ u64 s_listed;
    __shared__ float s_rmax;
    if (tid == 0) {
      s_gone[0] = s_gone[1] = 0u;
      s_rmax = -3.0e38f;
    }
    __syncthreads();
    const int r = tid / G, cg = tid % G;
    const int row_lane0 = (lane / G) * G;  // first lane of this row's group
    int nk = 0;                            // entries of the sorted list (uniform)
    float4* xs = xs_a;                     // current list; the merge writes the other buffer
    int* xp = xp_a;
    for (int i = 0; i < nb; ++i) {
      const int base = i * kTile;
      const int nrows = min(kTile, n - base);
      const u64 live_mask = nrows == 64 ? ~0ull : ((1ull << nrows) - 1);
      // ---- A: pull + diagonal words ----
      {
        bool dead = false;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 cb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrows) {
          b = sb[base + r];
          cb = xy_reach(b, thresh);
          if (cg == 0) rowc[r] = cb;
        }
        if (nk > 0) {
          const float R = cb.y + s_rmax + 1.0f;
          int bound = 0;
          if (r < nrows && cg < 2) {  // lane 0 of the row: first entry >= cx - R; lane 1: first entry > cx + R
            const float key = cg == 0 ? cb.x - R : cb.x + R;
            int lo = 0, hi = nk;
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              const float v = xs[mid].x;
              if (cg == 0 ? (v < key) : (v <= key)) lo = mid + 1;
              else hi = mid;
            }
            bound = lo;
          }
          const int jlo = __shfl_sync(0xffffffffu, bound, row_lane0), jhi = __shfl_sync(0xffffffffu, bound, row_lane0 + 1);
          const float area_b = legacy_area(b);
          for (int j = jlo + cg; j < jhi; j += G) {
            const float4 kc = xs[j];
            if (fabsf(kc.x - cb.x) - kc.y <= cb.y && fabsf(kc.z - cb.z) - kc.w <= cb.w) {  // near in x and in y
              const float4 kbx = sb[xp[j]];
              if (suppresses(kbx, legacy_area(kbx), b, area_b, thresh)) dead = true;
            }
          }
#pragma unroll
          for (int d = G / 2; d > 0; d >>= 1) dead |= __shfl_xor_sync(0xffffffffu, (int)dead, d) != 0;
          if (cg == 0 && dead) atomicOr(&s_gone[r >> 5], 1u << (r & 31));
        }
        // (the diagonal words of tiles > 0 were computed by warps 1.. during the previous tile's chain)
        if (i == 0) diag_item(base, nrows, 0ull, tid, diag[0]);
      }
      __syncthreads();
      NMS_T(2);
      // ---- B: chain, which of the kept rows are listed, largest reach ----
      if (tid < 32) {
        const u64 gone0 = ((u64)s_gone[1] << 32) | s_gone[0];
        const u64 kept = chain(diag[i & 1], nrows, ~gone0 & live_mask);
        const float4 c0v = rowc[lane], c1v = rowc[lane + 32];
        const bool l0 = ((kept >> lane) & 1ull) && c0v.x == c0v.x && c0v.y == c0v.y;
        const bool l1 = ((kept >> (lane + 32)) & 1ull) && c1v.x == c1v.x && c1v.y == c1v.y;
        const u64 listed = ((u64)__ballot_sync(0xffffffffu, l1) << 32) | __ballot_sync(0xffffffffu, l0);
        float rm = fmaxf(l0 ? c0v.y : -3.0e38f, l1 ? c1v.y : -3.0e38f);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) rm = fmaxf(rm, __shfl_xor_sync(0xffffffffu, rm, d));
        if (tid == 0) {
          s_kept = kept;
          s_listed = listed;
          s_nkept += __popcll(kept);
          s_rmax = fmaxf(s_rmax, rm);
          s_gone[0] = s_gone[1] = 0u;
          keepbits[i] = kept;  // (sorted input: positions are original indices)
        }
      } else if (base + kTile < n) {
        // meanwhile: the diagonal words of the next tile (pure pairwise facts, all rows)
        const int nbase = base + kTile, nrows2 = min(kTile, n - nbase);
        for (int item = tid - 32; item < kThreads; item += kThreads - 32)
          diag_item(nbase, nrows2, 0ull, item, diag[(i + 1) & 1]);
      }
      __syncthreads();
      NMS_T(4);
      if ((long long)s_nkept >= max_keep) break;  // uniform: shared value
      // ---- C: rank of every listed row among the listed rows of this tile, by (x centre, row) ----
      const u64 listed = s_listed;
      const int mv = __popcll(listed);
      {
        int less = 0;
        if (r < nrows && ((listed >> r) & 1ull)) {
          const float cx = rowc[r].x;
#pragma unroll
          for (int k = 0; k < CPT; ++k) {
            const int q = cg * CPT + k;
            if (q < nrows && ((listed >> q) & 1ull)) {
              const float cq = rowc[q].x;
              less += (cq < cx || (cq == cx && q < r)) ? 1 : 0;
            }
          }
        }
#pragma unroll
        for (int d = G / 2; d > 0; d >>= 1) less += __shfl_xor_sync(0xffffffffu, less, d);
        if (cg == 0 && r < nrows && ((listed >> r) & 1ull)) newk[less] = r;
      }
      __syncthreads();
      NMS_T(6);
      // ---- D: merge newk into the sorted list ----
      {
        // (keys (x centre, position) are distinct; old positions are all below this tile's)
        float4* xn = xs == xs_a ? xs_b : xs_a;
        int* pn = xp == xp_a ? xp_b : xp_a;
        for (int t = tid; t < nk + mv; t += kThreads) {
          if (t < nk) {  // old entry: moves up by the number of new entries before it (x centre strictly smaller)
            const float4 e = xs[t];
            int lo = 0, hi = mv;
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (rowc[newk[mid]].x < e.x) lo = mid + 1;
              else hi = mid;
            }
            xn[t + lo] = e;
            pn[t + lo] = xp[t];
          } else {  // new entry q: behind the q new entries before it and the old entries with centre <= its own
            const int q = t - nk, row = newk[q];
            const float4 e = rowc[row];
            int lo = 0, hi = nk;
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (xs[mid].x <= e.x) lo = mid + 1;
              else hi = mid;
            }
            xn[q + lo] = e;
            pn[q + lo] = base + row;
          }
        }
        xs = xn;
        xp = pn;
        nk += mv;
      }
      __syncthreads();
      NMS_T(7);
#ifdef B200_NMS_STATS
      if (tid == 0) g_nms_stats[blockIdx.x][9] += 1;
#endif
    }
  } else {
    // ---- full sweep (no early stop, or unsorted input): PUSH ------------------------------------------------
    // Tiles of 64 boxes in order: (a) diagonal words, (b) chain, (c) every still-alive LATER box is tested
    // against the tile's KEPT boxes, warp verdicts merged by a ballot into the "removed" words.
    int diag_tile = -1;  // tile whose diagonal words sit in diag[diag_tile & 1] (uniform)
    for (int i = 0; i < nb; ++i) {
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
      // (b) in-tile chain on warp 0; meanwhile the other warps compute the diagonal words of tile
      // i+1 (pure pairwise facts: rows that turn out to be removed are simply not used)
      const int nbase = base + kTile;
      const bool pre = nbase < n;  // uniform
      if (tid < 32) {
        const u64 kept = chain(diag[i & 1], nrows, ~gone0 & live_mask);
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
      __syncthreads();
      const u64 kept = s_kept;
      const int m = __popcll(kept);
      if (tid < kTile && ((kept >> tid) & 1ull)) {
        const int pos = __popcll(kept & ((1ull << tid) - 1));
        const float4 b = sb[base + tid];
        kb[pos] = b;
        ka[pos] = legacy_area(b);
        // (original positions; a 64-bit shared atomicOr is a CAS loop -- 64 threads on one word took 5.8 k cycles
        // per tile -- so sorted input stores the tile's word directly and unsorted input uses 32-bit atomics)
        if (unsorted) {
          const int o = order[off + base + tid];
          atomicOr(reinterpret_cast<uint32_t*>(keepbits) + (o >> 5), 1u << (o & 31));
        }
      }
      if (!unsorted && tid == 0) keepbits[i] = kept;
      __syncthreads();
      // (c) later boxes vs this tile's kept boxes
      for (int j0 = base + kTile + (tid & ~31); j0 < n; j0 += kThreads) {
        const int j = j0 + lane;
        // warp-uniform trip count, predicated body: the lanes of a warp stay converged
        // (a per-lane `break` made ptxas serialise the 32 lanes of this loop)
        const uint32_t dead = remv[j0 >> 5];
        const bool cand = j < n && !((dead >> lane) & 1u);
        bool alive = cand;
        const float4 b = sb[cand ? j : base];
        const float area_b = legacy_area(b);
        for (int rr = 0; rr < m; ++rr)
          if (alive && suppresses(kb[rr], ka[rr], b, area_b, thresh)) alive = false;
        const uint32_t v = __ballot_sync(0xffffffffu, cand && !alive);
        if (lane == 0 && v) remv[j0 >> 5] = dead | v;
      }
      __syncthreads();
    }
  }
  __syncthreads();

  // ---- ascending compaction of the kept original indices --------------------------------
  // inclusive scan of the per-word counts by warp 0: kMaxTiles / 32 consecutive words per lane, one warp scan
  if (tid < 32) {
    constexpr int kW = kMaxTiles / 32;
    static_assert(kMaxTiles % 32 == 0, "words per lane");
    int c[kW], run = 0;
#pragma unroll
    for (int j = 0; j < kW; ++j) {
      const int w = lane * kW + j;
      c[j] = w < nb ? __popcll(keepbits[w]) : 0;
      run += c[j];
    }
    int incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    int acc = incl - run;
#pragma unroll
    for (int j = 0; j < kW; ++j) {
      acc += c[j];
      scan[lane * kW + j] = acc;
    }
  }
  __syncthreads();
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
int g_nms_force_fused = 0;

// dynamic shared memory of the fused kernel: boxes, and for the early-stopping mode the kept list
// (box, area, x centre/reach) plus the per-chunk x-order (sort keys + permutation)
size_t fused_smem_bytes(size_t n_cap, size_t kl_cap) {
  size_t b = sizeof(float4) * n_cap;
  if (kl_cap > 0)
    b += 2 * (sizeof(float4) + sizeof(int)) * kl_cap;  // two buffers of (centre / reach in x and y) + position
  return b;
}

// the fused kernel needs the segment's boxes (+ the kept list) in one CTA's shared memory
bool fused_applies(int64_t max_seg_len, int64_t max_keep, int64_t n_total, int64_t n_segments) {
  if (g_nms_force_bitmask || max_seg_len > kFusedMaxSeg) return false;
  // keep-all over long segments (layers.nms / boxlist_nms without max_proposals on RPN-sized inputs): the push sweep
  // of the fused kernel tests every kept box against every later live box from ONE CTA per segment, the bitmask path
  // spreads the same pair tests over (segment, 64 x 64 tile) CTAs -- measured 1.5-2.4 x faster from 500 to 6000 boxes
  // per segment (scripts/perf_nms_keepall.py).  Many short segments (the box head's per-class lists) stay fused: the
  // bitmask grid is sized by the LONGEST segment.
  if (g_nms_force_fused == 0 && max_keep <= 0 && max_seg_len >= 512 && n_total >= 384 * n_segments) return false;
  const size_t n_cap = (size_t)b200::ceil_div<int64_t>(max_seg_len, kTile) * kTile;
  const size_t kl_cap = max_keep > 0 ? (size_t)(max_keep < max_seg_len ? max_keep : max_seg_len) + kTile : 0;
  return fused_smem_bytes(n_cap, kl_cap) + 10 * 1024 <= 227 * 1024;
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

extern "C" void b200_debug_nms(int mode) {  // 0 = default choice, 1 = bitmask path, 2 = fused kernel wherever it fits
  g_nms_force_bitmask = mode == 1;
  g_nms_force_fused = mode == 2 ? 1 : 0;
}

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
  if (fused_applies(max_seg_len, max_keep, n_total, n_segments)) {
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
