// roi_align_fwd_rows.cu -- row-streaming ("fast math") RoIAlign forward for the FPN box pooler on
// sm_100a: 7x7 bins, sampling_ratio 2, NHWC features with 256 channels
// (modeling/poolers.py:91-121 + csrc/cpu/ROIAlign_cpu.cpp:18-218 for all levels in one launch).
//
// Why a second fast kernel: roi_align_fwd_sep keeps the taps of ONE feature column per thread in
// flight (registers), and its column step waits out a full L2 / HBM round trip -- ncu shows ~11
// warps stalled on the long scoreboard per issued instruction and 44 % issue utilisation
// (profiles/r01_roi_align_fwd_sep.txt).  Here the memory system is driven by the copy engine
// instead of by registers:
//   * persistent CTAs, one per SM; CTA b walks RoIs b, b + gridDim, ... (image-major order is kept,
//     so an image's pyramid is still pulled from HBM about once);
//   * two planner warps (alternating RoIs, four RoIs ahead) build a RoI's tables: the distinct tap
//     rows / columns, the merged x weights of every bin, the ring entries -- one per (tap row, pair
//     of output rows it feeds), grouped by output row -- and, holding a placement turn, where each
//     entry goes in a 168 KB byte ring and which earlier entry's release its copy must wait for;
//   * two copy warps (entry g belongs to warp g % 2) stream every entry -- the run(s) of tapped
//     columns x all 256 channels, contiguous in NHWC -- into the ring with cp.async.bulk (1-D TMA,
//     up to 28 KB per request, mbarrier complete_tx); up to 16 entries (~9 at the mean patch size,
//     ~160 KB per SM) are in flight, across RoI boundaries;
//   * 14 consumer warps: thread = (channel quad, output column pw).  Per arriving entry a thread
//     reduces the (<= 4, duplicates merged) tap columns of its bin to one value per channel (the x
//     pass, FMUL2 / FFMA2: two fp32 lanes per issue slot), gives the entry back, prefetches the next
//     entry's columns into the same registers, then adds into the accumulators of the two output
//     rows of the entry's group (the y pass; group loops are unrolled, so accumulator indices are
//     compile-time).  All 49 bins x 4 channels of a thread stay in registers;
//   * epilogue: accumulators -> [256 x 49] shared tile (bank-conflict-free through a per-lane
//     channel rotation) -> ONE cp.async.bulk store of the contiguous 50 KB block of the NCHW output
//     (evict-first), which drains while the consumers are in the next RoI; optional fused channel
//     mean as in roi_align_fwd_sep.
// Bilinear interpolation is separable, bin = 1/4 * sum_y sum_x wy * wx * f(y, x); the 1/4 is folded
// into the row weights (exact: a power of two).  The result differs from the reference's
// summation order by fp32 reassociation only (<= 1e-5 relative, tests/test_gpu_roi_align_rows.py);
// the table / ring-placement algorithm is also checked on the CPU (tests/rows_tables_emulation.py).
// Measured steps and dead ends (one producer warp doing everything: 0.77 ms; fixed 28 KB slots;
// unconditional 4-column loads; 28 consumer warps with 2 channels each; L2 evict-last loads;
// spatially sorted RoIs): DESIGN.md, "RoIAlign forward".
#include <cstddef>

#include "roi_align_fwd.cuh"

namespace b200 {
namespace {

constexpr int kP = 7;                            // pooled height = width
constexpr int kNS = 2 * kP;                      // samples per axis
constexpr int kBins = kP * kP;                   // 49
constexpr int kMaxList = 2 * kNS;                // distinct tap rows / columns of a RoI, at most
constexpr int kRC = 256;                         // channels
constexpr int kPxBytes = kRC * 4;                // one pixel, all channels
constexpr int kRingPx = 168;                     // byte ring of tap rows, in pixels (1 KB units)
constexpr int kRingBytes = kRingPx * kPxBytes;
constexpr int kNBar = 16;                        // ring entries in flight, at most (barrier pairs)
constexpr int kHist = 64;                        // placement history the planner keeps (> kNBar + kMaxList)
constexpr int kTabs = 4;                         // RoI tables in flight
// consumers: thread = (kV channels, output column pw); kV = 4: 14 warps, kV = 2: 28 warps
constexpr int cons_warps(int v) { return kP * (kRC / v / 32); }
constexpr int kPlanWarps = 2;  // planner warps, alternating RoIs (one warp's ~900 instructions per RoI bounded the kernel)
constexpr int kTileFloats = kRC * kBins;

// A ring entry = one copy of a tap row, feeding output rows b and b + 1 with weights w0, w1 (x 1/4).
// Entries are grouped by b (ascending), so the consumers' accumulator indices are compile-time
// constants; a tap row feeding more than two output rows (bins narrower than ~1.3 pixels) is simply
// listed -- and copied -- once per pair.
// `place`: where the planner put the entry in the byte ring; `dep`: the copy may be issued once every
// ring entry with a sequence number < dep has been given back (entries are given back in order)
struct __align__(16) RingEntry {
  float w0, w1;
  uint32_t place, dep;
};
struct __align__(16) RoiTab {
  int nent, ncols, nruns, level;
  int batch, width, pad0, pad1;
  int cx[kP][4];    // byte offset inside a slot of the k-th merged tap column of bin pw
  float wx[kP][4];  // its weight (0 for unused entries)
  int nk[8];        // merged tap columns per bin
  int gend[8];      // entries [gend[b-1], gend[b]) feed output rows b, b + 1
  RingEntry ent[kMaxList];
  int ent_y[kMaxList];    // feature row of entry e
  int run_pos[kMaxList], run_col[kMaxList], run_len[kMaxList];  // runs of consecutive tapped columns
};

constexpr size_t kRowsSmem = (size_t)kRingBytes + sizeof(float) * kTileFloats + kTabs * sizeof(RoiTab);

// packed fp32 pairs: FFMA2 / FMUL2 on sm_100a do two lanes of fma.rn per issue slot, same rounding
struct V4 {
  float2 lo, hi;  // channels (0, 1) and (2, 3) of a quad
};
__device__ __forceinline__ V4 lds_v4(uint32_t a) {
  V4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.lo.x), "=f"(v.lo.y), "=f"(v.hi.x), "=f"(v.hi.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ int lds_i32(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t phase) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(phase)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t phase) {
  while (!mbar_try_wait_a(bar, phase)) {
  }
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// shared -> global bulk store (TMA, one thread), tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, uint32_t src_smem, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
               "r"(src_smem), "r"(bytes), "l"(policy)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
template <int kThreads>
__device__ __forceinline__ void bar_consumers() {
  asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
}
__device__ __forceinline__ float2 lds_v2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}

struct RoiPlace {
  int level, batch, H, W;
};

// Producer warp: axis tables of RoI r.  Lanes 0-13 own the y samples, lanes 16-29 the x samples.
// planner-private scratch (one copy; the tables above are what the other warps read)
struct __align__(16) PlanScratch {
  float wy[kMaxList][8];  // weight (x 1/4) of list row i for output row ph
  int rowy[kMaxList];     // feature row of list entry i
  int colx[kMaxList];     // feature column of list entry j
};
// ring placement state, handed from planner warp to planner warp with the placement turn
struct __align__(16) RingState {
  uint32_t g0;    // ring entries of the RoIs placed so far
  uint32_t head;  // next free pixel of the ring
  uint32_t pad[2];
  uint32_t hist_place[kHist], hist_size[kHist];  // placement of the last kHist entries (pixels)
};

__device__ __forceinline__ RoiPlace build_rows_tab(const LevelTable& lt, const float* __restrict__ rois, long long r,
                                                   RoiTab* tb, int lane, PlanScratch* sc) {
  const RoiHeader h = load_roi(rois, r, lt);
  RoiPlace pl;
  pl.level = h.level;
  pl.batch = h.batch;
  pl.H = pl.W = 1;
  if (h.level < 0) {  // matches no level: the output rows stay zero (poolers.py:111-119)
    if (lane < 8) tb->gend[lane] = 0;
    if (lane == 0) {
      tb->batch = 0;
      tb->width = 1;
      tb->nent = 0;
      tb->ncols = 0;
      tb->nruns = 0;
      tb->level = h.level;
    }
    return pl;
  }
  const int H = lt.H[h.level], W = lt.W[h.level];
  pl.H = H;
  pl.W = W;
  const RoiGeom g = roi_geometry(h.x1, h.y1, h.x2, h.y2, lt.scale[h.level], kP, kP, 2);
  const int hl = lane & 15;
  const bool isx = lane >= 16;
  bool ok = false;
  AxisTap t;
  t.lo = t.hi = 0;
  t.l = t.h = 0.f;
  if (hl < kNS)
    t = isx ? axis_sample(g.start_w, hl >> 1, g.bin_w, hl & 1, 2, W, ok)
            : axis_sample(g.start_h, hl >> 1, g.bin_h, hl & 1, 2, H, ok);
  ok = ok && hl < kNS;
  // an out-of-range sample contributes nothing (ROIAlign_cpu.cpp:47-61): zero weights, parked on 0
  const int lo = ok ? t.lo : 0, hi = ok ? t.hi : 0;
  const float wl = ok ? t.l : 0.f, wh = ok ? t.h : 0.f;
  // the list of distinct coordinates in marching order (same rule as build_sep_tables)
  const int plo = __shfl_up_sync(0xffffffffu, lo, 1, 16), phi = __shfl_up_sync(0xffffffffu, hi, 1, 16);
  int nnew = lo == hi ? 1 : 2;
  if (hl > 0) {
    if (lo == plo && hi == phi) nnew = 0;
    else if (lo == phi) nnew = 1;
  }
  if (hl >= kNS) nnew = 0;
  int scan = nnew;
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, scan, d, 16);
    if (hl >= d) scan += v;
  }
  const int nrows = __shfl_sync(0xffffffffu, scan, 15), ncols = __shfl_sync(0xffffffffu, scan, 31);
  int* list = isx ? sc->colx : sc->rowy;
  if (hl < kNS) {
    if (nnew >= 1) list[scan - 1] = hi;
    if (nnew == 2) list[scan - 2] = lo;
  }
  // zero the row-weight table
  float* wyf = &sc->wy[0][0];
  for (int i = lane; i < kMaxList * 8; i += 32) wyf[i] = 0.f;
  // the two taps of this sample as (list index, weight)
  const int jhi = scan - 1, jlo = lo == hi ? jhi : jhi - 1;
  const float w_lo = lo == hi ? 0.f : wh, w_hi = lo == hi ? wl + wh : wl;
  // bin b = samples 2b, 2b+1 of the same half; lanes 0-6 / 16-22 merge the four taps of bin hl
  const int sa = (lane & 16) + ((2 * hl) & 15), sb = (lane & 16) + ((2 * hl + 1) & 15);
  int idx[4];
  float w[4];
  idx[0] = __shfl_sync(0xffffffffu, jlo, sa);
  idx[1] = __shfl_sync(0xffffffffu, jhi, sa);
  idx[2] = __shfl_sync(0xffffffffu, jlo, sb);
  idx[3] = __shfl_sync(0xffffffffu, jhi, sb);
  w[0] = __shfl_sync(0xffffffffu, w_lo, sa);
  w[1] = __shfl_sync(0xffffffffu, w_hi, sa);
  w[2] = __shfl_sync(0xffffffffu, w_lo, sb);
  w[3] = __shfl_sync(0xffffffffu, w_hi, sb);
#pragma unroll
  for (int k = 1; k < 4; ++k) {
#pragma unroll
    for (int j = 0; j < k; ++j) {
      if (idx[k] == idx[j]) {  // weights of equal coordinates add up in the earliest entry
        w[j] += w[k];
        w[k] = 0.f;
      }
    }
  }
  __syncwarp();  // wy zeroed, lists written
  if (hl < kP) {
    if (isx) {
      int m = 0;
      int cxo[4] = {0, 0, 0, 0};
      float wxo[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (w[k] != 0.f) {
#pragma unroll
          for (int s = 0; s < 4; ++s)
            if (s == m) {
              cxo[s] = idx[k] * kPxBytes;
              wxo[s] = w[k];
            }
          ++m;
        }
      }
#pragma unroll
      for (int s = 1; s < 4; ++s)
        if (s >= m) cxo[s] = cxo[0];  // unused entries: a valid address, weight 0
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        tb->cx[hl][s] = cxo[s];
        tb->wx[hl][s] = wxo[s];
      }
      tb->nk[hl] = m;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (w[k] != 0.f) sc->wy[idx[k]][hl] = 0.25f * w[k];
    }
  }
  __syncwarp();
  // ring entries: list row i feeds the output rows in `mask`; pair them greedily from the lowest
  // bit (b, b + 1), then order the pairs of all rows by b (stable counting sort over the warp)
  int nent = 0;
  {
    unsigned mask = 0;
    if (lane < nrows) {
#pragma unroll
      for (int ph = 0; ph < kP; ++ph)
        if (sc->wy[lane][ph] != 0.f) mask |= 1u << ph;
    }
    unsigned pk = 0;
#pragma unroll
    for (int b = 0; b < kP; ++b)
      if ((mask >> b) & 1u) {
        pk |= 1u << b;
        mask &= ~(3u << b);
      }
    const int y = lane < nrows ? sc->rowy[lane] : 0;
#pragma unroll
    for (int b = 0; b < kP; ++b) {
      const bool f = (pk >> b) & 1u;
      const unsigned bal = __ballot_sync(0xffffffffu, f);
      if (f) {
        const int dst = nent + __popc(bal & ((1u << lane) - 1u));
        tb->ent_y[dst] = y;
        tb->ent[dst].w0 = sc->wy[lane][b];
        tb->ent[dst].w1 = b + 1 < kP ? sc->wy[lane][b + 1] : 0.f;
      }
      nent += __popc(bal);
      if (lane == 0) tb->gend[b] = nent;
    }
  }
  // runs of consecutive tapped columns: one bulk copy each
  {
    const int x = lane < ncols ? sc->colx[lane] : 0;
    const int xp = (lane > 0 && lane < ncols) ? sc->colx[lane - 1] : 0;
    const bool start = lane < ncols && (lane == 0 || x != xp + 1);
    const unsigned smask = __ballot_sync(0xffffffffu, start);
    if (start) {
      const int id = __popc(smask & ((1u << lane) - 1u));
      const unsigned higher = lane < 31 ? (smask >> (lane + 1)) : 0u;
      const int next = higher ? lane + __ffs(higher) : ncols;
      tb->run_pos[id] = lane;
      tb->run_col[id] = x;
      tb->run_len[id] = next - lane;
    }
    if (lane == 0) {
      tb->batch = h.batch;
      tb->width = W;
      tb->nent = nent;
      tb->ncols = ncols;
      tb->nruns = __popc(smask);
      tb->level = h.level;
    }
  }
  return pl;
}

// Ring placement of the entries of one RoI (the planner warp that holds the placement turn): the
// entries all span ncols pixels and are laid one after the other, wrapping to the start of the ring
// when the next one would not fit; an entry's copy depends on the newest earlier entry whose bytes it
// overwrites (found in the placement history) and on the entry that last used its barrier pair.
__device__ __forceinline__ void place_rows(RoiTab* tb, int lane, RingState* rs) {
  const int nent = tb->nent;
  if (nent == 0) return;
  const uint32_t s = (uint32_t)tb->ncols;  // >= 1
  const uint32_t g0 = rs->g0, head = rs->head;
  // small non-negative integers: floor((a + 0.5) / s) through the fp32 reciprocal is exact
  const float inv = __frcp_rn((float)s);
  const uint32_t k0 = (uint32_t)(((float)(kRingPx - (int)head) + 0.5f) * inv);
  const uint32_t kfit = (uint32_t)(((float)kRingPx + 0.5f) * inv);
  const uint32_t i = (uint32_t)lane;
  uint32_t j = i - k0;
  if (i >= k0)
    while (j >= kfit) j -= kfit;
  const uint32_t place = i < k0 ? head + i * s : j * s;
  const uint32_t g = g0 + i;
  __syncwarp();  // everyone has read g0 / head
  if (lane < nent) {
    rs->hist_place[g % kHist] = place;
    rs->hist_size[g % kHist] = s;
  }
  __syncwarp();
  if (lane < nent) {
    uint32_t dep = g >= (uint32_t)kNBar ? g - (uint32_t)kNBar + 1u : 0u;
#pragma unroll 5
    for (uint32_t d = 1; d < (uint32_t)kNBar; ++d) {
      if (d > g) break;
      const uint32_t e = g - d;
      const uint32_t pe = rs->hist_place[e % kHist], se = rs->hist_size[e % kHist];
      if (pe < place + s && place < pe + se) {
        dep = dep > e + 1u ? dep : e + 1u;
        break;  // older overlapping entries only give smaller values
      }
    }
    tb->ent[lane].place = place * (uint32_t)kPxBytes;
    tb->ent[lane].dep = dep;
  }
  if (lane == nent - 1) {
    rs->head = place + s;
    rs->g0 = g0 + (uint32_t)nent;
  }
}

// kCopyWarps: warps issuing the bulk copies (ring entry g belongs to warp g % kCopyWarps -- a single
// warp's dependent instruction stream, ~40 instructions per entry, was what bounded the first version);
// kProbe != 0: consumers skip the arithmetic (copy-engine throughput probe)
template <int kCopyWarps, int kV, int kProbe>
__global__ void __launch_bounds__(32 * cons_warps(kV) + 32 * kPlanWarps + 32 * kCopyWarps, 1)
roi_align_fwd_rows(const LevelTable lt, const float* __restrict__ rois, long long n_rois, float* __restrict__ out,
                   float* __restrict__ out_mean, int32_t* __restrict__ out_levels) {
  // (no integer round trip on this pointer: the compiler must keep seeing shared-space addresses,
  // or every access below turns into a generic LD.E / ST.E)
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  __shared__ __align__(128) uint64_t full_bar[kNBar], empty_bar[kNBar], tab_full[kTabs], tab_empty[kTabs];
  __shared__ PlanScratch plan_scratch[kPlanWarps];
  __shared__ RingState ring_state;
  __shared__ __align__(8) uint64_t place_turn[kPlanWarps];
  unsigned char* ring = smem_dyn;
  float* tile = reinterpret_cast<float*>(smem_dyn + (size_t)kRingBytes);
  RoiTab* tabs = reinterpret_cast<RoiTab*>(tile + kTileFloats);

  constexpr int kWarpsPerBin = kRC / kV / 32;
  constexpr int kConsWarps = cons_warps(kV), kConsThreads = 32 * kConsWarps;
  constexpr int kPlanWarp0 = kConsWarps;               // first of the warps that build the RoI tables
  constexpr int kCopyWarp0 = kConsWarps + kPlanWarps;  // first of the warps that issue the bulk copies
  constexpr int kNP = kV / 2;                 // channel pairs per thread
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kNBar; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsWarps);
    }
    for (int s = 0; s < kPlanWarps; ++s) mbar_init(&place_turn[s], 1);
    ring_state.g0 = 0;
    ring_state.head = 0;
    for (int s = 0; s < kTabs; ++s) {
      mbar_init(&tab_full[s], 1);
      mbar_init(&tab_empty[s], kConsWarps + kCopyWarps);  // every reader of a table
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp >= kPlanWarp0 && warp < kCopyWarp0) {
    // ------------------------------- planners: RoI tables --------------------------------
    // planner p builds the tables of RoIs n = p, p + kPlanWarps, ...; only the ring placement is
    // sequential across RoIs: it is done holding the placement turn, passed round-robin
    const int p = warp - kPlanWarp0;
    int n = p, k = 0;
    for (long long r = blockIdx.x + (long long)p * gridDim.x; r < n_rois;
         r += (long long)kPlanWarps * gridDim.x, n += kPlanWarps, ++k) {
      const int ti = n % kTabs;
      mbar_wait(&tab_empty[ti], (uint32_t)(((n / kTabs) & 1) ^ 1));  // (a fresh barrier passes)
      const RoiPlace pl = build_rows_tab(lt, rois, r, tabs + ti, lane, &plan_scratch[p]);
      if (lane == 0 && out_levels) out_levels[r] = pl.level;
      __syncwarp();
      // my turn: after planner p - 1 placed RoI n - 1 (planner 0's first turn is free)
      mbar_wait(&place_turn[p], (uint32_t)((k & 1) ^ (p == 0 ? 1 : 0)));
      place_rows(tabs + ti, lane, &ring_state);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&place_turn[(p + 1) % kPlanWarps]);
        mbar_arrive(&tab_full[ti]);
      }
    }
    return;
  }

  if (warp >= kCopyWarp0) {
    // ------------------------------- copy warps: the row ring ----------------------------
    const int cw = warp - kCopyWarp0;
    uint32_t g0 = 0;  // ring entries of the RoIs before this one
    int n = 0;
    for (long long r = blockIdx.x; r < n_rois; r += gridDim.x, ++n) {
      const int ti = n % kTabs;
      const RoiTab* tb = tabs + ti;
      mbar_wait(&tab_full[ti], (uint32_t)((n / kTabs) & 1));
      const int nent = tb->nent, ncols = tb->ncols, nruns = tb->nruns;
      if (nent > 0) {
        const int level = tb->level, width = tb->width;
        const char* gbase = reinterpret_cast<const char*>(lt.data[level]) +
                            (size_t)tb->batch * lt.H[level] * width * kPxBytes;
        uint32_t my_pos = 0, my_len = 0;
        if (lane < nruns) {
          my_pos = (uint32_t)tb->run_pos[lane] * kPxBytes;
          my_len = (uint32_t)tb->run_len[lane] * kPxBytes;
          gbase += (size_t)tb->run_col[lane] * kPxBytes;
        }
        const uint32_t size = (uint32_t)ncols * kPxBytes;
        // first entry of this RoI that belongs to this warp
        int i = (int)((cw + kCopyWarps - g0 % kCopyWarps) % kCopyWarps);
        for (; i < nent; i += kCopyWarps) {
          const uint32_t g = g0 + (uint32_t)i, bi = g % kNBar;
          const uint32_t place = tb->ent[i].place, dep = tb->ent[i].dep;
          // (dep - 1 >= g - kNBar, so the parity below names one phase unambiguously)
          if (dep > 0) mbar_wait(&empty_bar[(dep - 1u) % kNBar], ((dep - 1u) / kNBar) & 1u);
          if (lane == 0) mbar_arrive_expect_tx(&full_bar[bi], size);
          if (lane < nruns)
            bulk_g2s(ring + place + my_pos, gbase + (size_t)tb->ent_y[i] * width * kPxBytes, my_len, &full_bar[bi]);
        }
        g0 += (uint32_t)nent;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&tab_empty[ti]);
    }
    return;
  }

  // --------------------------------- consumers --------------------------------------------
  const int pw = warp / kWarpsPerBin;
  const int q = (warp % kWarpsPerBin) * 32 + lane;
  const int rot = kV == 4 ? (lane >> 3) & 3 : (lane >> 4) & 1;
  // 32-bit shared addresses, pinned in registers (ptxas otherwise re-derives them from the CTA's
  // shared window at every use)
  uint32_t ring_a = smem_u32(ring) + q * (4 * kV), full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);
  uint32_t tabs_a = smem_u32(tabs);
  asm volatile("" : "+r"(ring_a), "+r"(full_a), "+r"(empty_a), "+r"(tabs_a));
  const uint64_t store_policy = policy_evict_first();
  uint32_t g = 0;   // ring entry to load into registers next
  uint32_t rg = 0;  // ring entry to give back next
  int n = 0;
  for (long long r = blockIdx.x; r < n_rois; r += gridDim.x, ++n) {
    const int ti = n % kTabs;
    const RoiTab* tb = tabs + ti;
    const uint32_t tb_a = tabs_a + (uint32_t)ti * (uint32_t)sizeof(RoiTab);
    mbar_wait(&tab_full[ti], (uint32_t)((n / kTabs) & 1));
    const int nent = tb->nent;
    const int4 cxv = *reinterpret_cast<const int4*>(tb->cx[pw]);
    const float4 wxv = *reinterpret_cast<const float4*>(tb->wx[pw]);
    const int nk = tb->nk[pw];
    const uint32_t c0 = ring_a + cxv.x, c1 = ring_a + cxv.y, c2 = ring_a + cxv.z, c3 = ring_a + cxv.w;
    const float2 wx0 = make_float2(wxv.x, wxv.x), wx1 = make_float2(wxv.y, wxv.y), wx2 = make_float2(wxv.z, wxv.z),
                 wx3 = make_float2(wxv.w, wxv.w);
    float2 acc2[kP][kNP];  // [output row][channel pair]
#pragma unroll
    for (int ph = 0; ph < kP; ++ph)
#pragma unroll
      for (int c = 0; c < kNP; ++c) acc2[ph][c] = make_float2(0.f, 0.f);
    float2 v[4][kNP];  // [tap column][channel pair]
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < kNP; ++c) v[k][c] = make_float2(0.f, 0.f);
    auto lds_col = [&](int k, uint32_t a) {
      if (kV == 4) {
        const V4 t = lds_v4(a);
        v[k][0] = t.lo;
        v[k][kNP - 1] = t.hi;
      } else {
        v[k][0] = lds_v2(a);
      }
    };
    // the tap columns of bin pw of the next ring entry -> registers
    auto load_row = [&](uint32_t off) {
      const uint32_t bi = g % kNBar;
      mbar_wait_a(full_a + 8u * bi, (g / kNBar) & 1u);
      if (kProbe == 0) {
        // (loading all four columns unconditionally -- unused ones repeat the first with weight 0 --
        // was measured 15 % slower: shared-memory wavefronts matter more than the two branches)
        lds_col(0, c0 + off);
        lds_col(1, c1 + off);
        if (nk > 2) lds_col(2, c2 + off);
        if (nk > 3) lds_col(3, c3 + off);
      }
      ++g;
    };
    uint32_t ent_a = tb_a + (uint32_t)offsetof(RoiTab, ent);
    const uint32_t ent_last = ent_a + 16u * (uint32_t)(nent - 1);
    if (nent > 0) load_row((uint32_t)lds_i32(ent_a + 8u));
#pragma unroll
    for (int b = 0; b < kP; ++b) {
      const uint32_t end_a = tb_a + (uint32_t)offsetof(RoiTab, ent) + 16u * (uint32_t)lds_i32(tb_a + (uint32_t)offsetof(RoiTab, gend) + 4u * b);
      for (; ent_a < end_a; ent_a += 16) {
        const float2 w = lds_f2(ent_a);
        const uint32_t next_off = (uint32_t)lds_i32(ent_a + 16u + 8u);  // (past the last entry: unused)
        // x pass: the tap columns reduced to one value per channel
        float2 u[kNP];
#pragma unroll
        for (int c = 0; c < kNP; ++c) u[c] = make_float2(0.f, 0.f);
        if (kProbe == 0) {
#pragma unroll
          for (int c = 0; c < kNP; ++c) {
            u[c] = __fmul2_rn(wx0, v[0][c]);
            u[c] = __ffma2_rn(wx1, v[1][c], u[c]);
          }
          if (nk > 2) {
#pragma unroll
            for (int c = 0; c < kNP; ++c) u[c] = __ffma2_rn(wx2, v[2][c], u[c]);
          }
          if (nk > 3) {
#pragma unroll
            for (int c = 0; c < kNP; ++c) u[c] = __ffma2_rn(wx3, v[3][c], u[c]);
          }
        }
        // this entry has been read: give it back, then fetch the next row while this one is accumulated
        __syncwarp();
        if (lane == 0) mbar_arrive_a(empty_a + 8u * (rg % kNBar));
        ++rg;
        if (ent_a < ent_last) load_row(next_off);
        // y pass
        const float2 w0 = make_float2(w.x, w.x);
#pragma unroll
        for (int c = 0; c < kNP; ++c) acc2[b][c] = __ffma2_rn(w0, u[c], acc2[b][c]);
        if (b + 1 < kP) {
          const float2 w1 = make_float2(w.y, w.y);
#pragma unroll
          for (int c = 0; c < kNP; ++c) acc2[b + 1][c] = __ffma2_rn(w1, u[c], acc2[b + 1][c]);
        }
      }
    }
    // the tables of this RoI are no longer needed
    __syncwarp();
    if (lane == 0) mbar_arrive(&tab_empty[ti]);

    // ---- epilogue: registers -> [256 x 49] tile -> one contiguous block of the output -------
    if (tid == 0) bulk_wait_read();  // the bulk store of the previous RoI has read the tile
    bar_consumers<kConsThreads>();
    {
      // lane groups write different channels of their thread's kV in one instruction (rotation by
      // lane / 8 for kV = 4, lane / 16 for kV = 2): the 32 lanes then hit 32 different banks
      // (thread stride 49 * kV floats = kV mod 32, channel stride 49 = 17 mod 32)
      float* t0 = tile + (kV * q) * kBins + pw;
      float* tc[kV];
#pragma unroll
      for (int j = 0; j < kV; ++j) tc[j] = t0 + ((j + rot) % kV) * kBins;
#pragma unroll
      for (int ph = 0; ph < kP; ++ph) {
        if (kV == 4) {
          float a = acc2[ph][0].x, b = acc2[ph][0].y, c = acc2[ph][kNP - 1].x, d = acc2[ph][kNP - 1].y;
          if (rot & 1) {
            const float t = a;
            a = b;
            b = c;
            c = d;
            d = t;
          }
          if (rot & 2) {
            float t = a;
            a = c;
            c = t;
            t = b;
            b = d;
            d = t;
          }
          tc[0][ph * kP] = a;
          tc[1][ph * kP] = b;
          tc[kV - 2][ph * kP] = c;
          tc[kV - 1][ph * kP] = d;
        } else {
          float a = acc2[ph][0].x, b = acc2[ph][0].y;
          if (rot & 1) {
            const float t = a;
            a = b;
            b = t;
          }
          tc[0][ph * kP] = a;
          tc[1][ph * kP] = b;
        }
      }
    }
    // one contiguous 50 KB block of the NCHW output: a single bulk store (TMA) drains the tile while
    // the consumers are already in the next RoI's rows
    fence_proxy_async();  // this thread's tile writes -> visible to the async proxy
    bar_consumers<kConsThreads>();
    {
      if (tid == 0) bulk_s2g(out + (size_t)r * kTileFloats, smem_u32(tile), kTileFloats * 4, store_policy);
      if (out_mean && tid < kRC) {
        const float* row = tile + tid * kBins;
        float s = 0.f;
#pragma unroll 7
        for (int i = 0; i < kBins; ++i) s = __fadd_rn(s, row[i]);
        out_mean[(size_t)r * kRC + tid] = __fdiv_rn(s, (float)kBins);
      }
    }
  }
  if (tid == 0) bulk_wait_all();  // the last store must have left shared memory before the CTA exits
}

}  // namespace

bool rows_kernel_applies(const LevelTable& lt, int C, int PH, int PW) {
  (void)lt;
  return C == kRC && PH == kP && PW == kP;
}

// Preconditions (checked by the caller): NHWC, sampling_ratio 2; rows_kernel_applies().
int launch_forward_rows(const LevelTable& lt, int C, const float* rois, int64_t n_rois, float* out, float* out_mean,
                        int32_t* out_levels, int variant, cudaStream_t st) {
  B200_REQUIRE(C == kRC, "roi_align rows kernel: %d channels", C);
  const int64_t grid = n_rois < sm_count() ? n_rois : sm_count();
#define B200_ROWS(CW, V, PROBE)                                                                                   \
  do {                                                                                                            \
    static SmemHighWater hw;                                                                                      \
    int rc = ensure_dynamic_smem(roi_align_fwd_rows<CW, V, PROBE>, kRowsSmem, &hw, "roi_align rows: smem");       \
    if (rc != B200_OK) return rc;                                                                                 \
    roi_align_fwd_rows<CW, V, PROBE><<<(unsigned)grid, 32 * cons_warps(V) + 32 * kPlanWarps + 32 * CW, kRowsSmem, st>>>(       \
        lt, rois, (long long)n_rois, out, out_mean, out_levels);                                                  \
  } while (0)
  // variant (tuning hook of b200_debug_set): bit 5 = two channels per consumer thread (28 consumer warps),
  // bit 6 = copy-engine probe (no arithmetic)
  if (variant & 64) B200_ROWS(2, 4, 1);
  else if (variant & 32) B200_ROWS(2, 2, 0);
  else B200_ROWS(2, 4, 0);
#undef B200_ROWS
  B200_CHECK_LAUNCH("roi_align_fwd_rows");
  return B200_OK;
}

}  // namespace b200
