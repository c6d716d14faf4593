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
//   * a producer warp builds the RoI's axis tables and streams every DISTINCT tap row of the RoI
//     -- the run(s) of tapped columns x all 256 channels, contiguous in NHWC -- into a ring of
//     shared-memory slots with cp.async.bulk (1-D TMA, up to 28 KB per request, mbarrier
//     complete_tx); it runs ahead across RoI boundaries, so ~100 KB per SM are in flight;
//   * 14 consumer warps: thread = (channel quad, output column pw).  Per arriving row a thread
//     reduces the (<= 4, duplicates merged) tap columns of its bin to one float4 (the x pass),
//     then adds it into the accumulators of the (<= 2, else generic path) output rows that tap
//     this row (the y pass).  All 49 bins x 4 channels of a thread stay in registers; nothing
//     waits on global memory.
//   * epilogue: accumulators -> [256 x 49] shared tile (bank-conflict-free through a per-lane
//     channel rotation) -> one contiguous 50 KB block of the NCHW output with streaming stores;
//     optional fused channel mean as in roi_align_fwd_sep.
// Bilinear interpolation is separable, bin = 1/4 * sum_y sum_x wy * wx * f(y, x); the 1/4 is folded
// into the row weights (exact: a power of two).  The result differs from the reference's
// summation order by fp32 reassociation only (<= 1e-5 relative, tests/test_gpu_roi_align.py).
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
constexpr int kWarpsPerBin = kRC / 4 / 32;       // 2
constexpr int kConsWarps = kP * kWarpsPerBin;    // 14
constexpr int kConsThreads = kConsWarps * 32;    // 448
constexpr int kPlanWarp = kConsWarps;            // builds the RoI tables, kTabs RoIs ahead
constexpr int kCopyWarp0 = kConsWarps + 1;       // first of the warps that issue the bulk copies
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
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(kConsThreads) : "memory"); }

struct RoiPlace {
  int level, batch, H, W;
};

// Producer warp: axis tables of RoI r.  Lanes 0-13 own the y samples, lanes 16-29 the x samples.
// planner-private scratch (one copy; the tables above are what the other warps read)
struct __align__(16) PlanScratch {
  float wy[kMaxList][8];  // weight (x 1/4) of list row i for output row ph
  int rowy[kMaxList];     // feature row of list entry i
  int colx[kMaxList];     // feature column of list entry j
  uint32_t hist_place[kHist], hist_size[kHist];  // ring placement of the last kHist entries (pixels)
};

struct RingPlanner {
  uint32_t g0;    // ring entries of the RoIs planned so far
  uint32_t head;  // next free pixel of the ring
};

__device__ __forceinline__ RoiPlace build_rows_tab(const LevelTable& lt, const float* __restrict__ rois, long long r,
                                                   RoiTab* tb, int lane, RingPlanner& rp, PlanScratch* sc) {
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
  // ring placement: the entries of a RoI all span ncols pixels and are laid one after the other,
  // wrapping to the start of the ring when the next one would not fit; an entry's copy depends on the
  // newest earlier entry whose bytes it overwrites (found in the placement history) and on the
  // entry that last used its barrier pair
  {
    const uint32_t s = (uint32_t)ncols;  // >= 1
    const uint32_t k0 = ((uint32_t)kRingPx - rp.head) / s, kfit = (uint32_t)kRingPx / s;
    const uint32_t i = (uint32_t)lane;
    const uint32_t place = i < k0 ? rp.head + i * s : ((i - k0) % kfit) * s;
    const uint32_t g = rp.g0 + i;
    if (lane < nent) {
      sc->hist_place[g % kHist] = place;
      sc->hist_size[g % kHist] = s;
    }
    __syncwarp();
    if (lane < nent) {
      uint32_t dep = g >= (uint32_t)kNBar ? g - (uint32_t)kNBar + 1u : 0u;
#pragma unroll 5
      for (uint32_t d = 1; d < (uint32_t)kNBar; ++d) {
        if (d > g) break;
        const uint32_t e = g - d;
        const uint32_t pe = sc->hist_place[e % kHist], se = sc->hist_size[e % kHist];
        if (pe < place + s && place < pe + se) {
          dep = dep > e + 1u ? dep : e + 1u;
          break;  // older overlapping entries only give smaller values
        }
      }
      tb->ent[lane].place = place * (uint32_t)kPxBytes;
      tb->ent[lane].dep = dep;
    }
    if (nent > 0) rp.head = __shfl_sync(0xffffffffu, place, nent - 1) + s;
    rp.g0 += (uint32_t)nent;
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

// kCopyWarps: warps issuing the bulk copies (ring entry g belongs to warp g % kCopyWarps -- a single
// warp's dependent instruction stream, ~40 instructions per entry, was what bounded the first version);
// kProbe != 0: consumers skip the arithmetic (copy-engine throughput probe)
template <int kCopyWarps, int kProbe>
__global__ void __launch_bounds__(kConsThreads + 32 + 32 * kCopyWarps, 1)
roi_align_fwd_rows(const LevelTable lt, const float* __restrict__ rois, long long n_rois, float* __restrict__ out,
                   float* __restrict__ out_mean, int32_t* __restrict__ out_levels) {
  // (no integer round trip on this pointer: the compiler must keep seeing shared-space addresses,
  // or every access below turns into a generic LD.E / ST.E)
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  __shared__ __align__(128) uint64_t full_bar[kNBar], empty_bar[kNBar], tab_full[kTabs], tab_empty[kTabs];
  __shared__ PlanScratch plan_scratch;
  unsigned char* ring = smem_dyn;
  float* tile = reinterpret_cast<float*>(smem_dyn + (size_t)kRingBytes);
  RoiTab* tabs = reinterpret_cast<RoiTab*>(tile + kTileFloats);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kNBar; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsWarps);
    }
    for (int s = 0; s < kTabs; ++s) {
      mbar_init(&tab_full[s], 1);
      mbar_init(&tab_empty[s], kConsWarps + kCopyWarps);  // every reader of a table
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kPlanWarp) {
    // ------------------------------- planner: RoI tables ---------------------------------
    RingPlanner rp;
    rp.g0 = 0;
    rp.head = 0;
    int n = 0;
    for (long long r = blockIdx.x; r < n_rois; r += gridDim.x, ++n) {
      const int ti = n % kTabs;
      mbar_wait(&tab_empty[ti], (uint32_t)(((n / kTabs) & 1) ^ 1));  // (a fresh barrier passes)
      const RoiPlace pl = build_rows_tab(lt, rois, r, tabs + ti, lane, rp, &plan_scratch);
      if (lane == 0 && out_levels) out_levels[r] = pl.level;
      __syncwarp();
      if (lane == 0) mbar_arrive(&tab_full[ti]);
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
  const int rot = (lane >> 3) & 3;
  // 32-bit shared addresses, pinned in registers (ptxas otherwise re-derives them from the CTA's
  // shared window at every use)
  uint32_t ring_a = smem_u32(ring) + q * 16, full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);
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
    float2 alo[kP], ahi[kP];
#pragma unroll
    for (int ph = 0; ph < kP; ++ph) alo[ph] = ahi[ph] = make_float2(0.f, 0.f);
    V4 v0, v1, v2, v3;
    v0.lo = v0.hi = make_float2(0.f, 0.f);
    v1 = v2 = v3 = v0;
    // the tap columns of bin pw of the next ring entry -> registers
    auto load_row = [&](uint32_t off) {
      const uint32_t bi = g % kNBar;
      mbar_wait_a(full_a + 8u * bi, (g / kNBar) & 1u);
      if (kProbe == 0) {
        // (loading all four columns unconditionally -- unused ones repeat the first with weight 0 --
        // was measured 15 % slower: shared-memory wavefronts matter more than the two branches)
        v0 = lds_v4(c0 + off);
        v1 = lds_v4(c1 + off);
        if (nk > 2) v2 = lds_v4(c2 + off);
        if (nk > 3) v3 = lds_v4(c3 + off);
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
        // x pass: the tap columns reduced to one value per channel
        const float2 w = lds_f2(ent_a);
        const uint32_t next_off = (uint32_t)lds_i32(ent_a + 16u + 8u);  // (past the last entry: unused)
        float2 ulo = make_float2(0.f, 0.f), uhi = ulo;
        if (kProbe == 0) {
          ulo = __fmul2_rn(wx0, v0.lo), uhi = __fmul2_rn(wx0, v0.hi);
          ulo = __ffma2_rn(wx1, v1.lo, ulo);
          uhi = __ffma2_rn(wx1, v1.hi, uhi);
          if (nk > 2) {
            ulo = __ffma2_rn(wx2, v2.lo, ulo);
            uhi = __ffma2_rn(wx2, v2.hi, uhi);
          }
          if (nk > 3) {
            ulo = __ffma2_rn(wx3, v3.lo, ulo);
            uhi = __ffma2_rn(wx3, v3.hi, uhi);
          }
        }
        // this entry has been read: give it back, then fetch the next row while this one is accumulated
        __syncwarp();
        if (lane == 0) mbar_arrive_a(empty_a + 8u * (rg % kNBar));
        ++rg;
        if (ent_a < ent_last) load_row(next_off);
        // y pass
        const float2 w0 = make_float2(w.x, w.x);
        alo[b] = __ffma2_rn(w0, ulo, alo[b]);
        ahi[b] = __ffma2_rn(w0, uhi, ahi[b]);
        if (b + 1 < kP) {
          const float2 w1 = make_float2(w.y, w.y);
          alo[b + 1] = __ffma2_rn(w1, ulo, alo[b + 1]);
          ahi[b + 1] = __ffma2_rn(w1, uhi, ahi[b + 1]);
        }
      }
    }
    float4 acc[kP];
#pragma unroll
    for (int ph = 0; ph < kP; ++ph) acc[ph] = make_float4(alo[ph].x, alo[ph].y, ahi[ph].x, ahi[ph].y);
    // the tables of this RoI are no longer needed
    __syncwarp();
    if (lane == 0) mbar_arrive(&tab_empty[ti]);

    // ---- epilogue: registers -> [256 x 49] tile -> one contiguous block of the output -------
    if (tid == 0) bulk_wait_read();  // the bulk store of the previous RoI has read the tile
    bar_consumers();
    {
      // lane groups of 8 write different channels of their quad in one instruction (rotation by
      // lane / 8): 32 lanes then hit 32 different banks (quad stride 196 = 4 mod 32, channel 49 = 17)
      float* t0 = tile + (4 * q) * kBins + pw;
      float* tc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) tc[j] = t0 + ((j + rot) & 3) * kBins;
#pragma unroll
      for (int ph = 0; ph < kP; ++ph) {
        float a = acc[ph].x, b = acc[ph].y, c = acc[ph].z, d = acc[ph].w;
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
        tc[2][ph * kP] = c;
        tc[3][ph * kP] = d;
      }
    }
    // one contiguous 50 KB block of the NCHW output: a single bulk store (TMA) drains the tile while
    // the consumers are already in the next RoI's rows
    fence_proxy_async();  // this thread's tile writes -> visible to the async proxy
    bar_consumers();
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
  static_assert(kTileFloats / 4 % kConsThreads == 0, "copy-out assumes whole passes");
  B200_REQUIRE(C == kRC, "roi_align rows kernel: %d channels", C);
  const int64_t grid = n_rois < sm_count() ? n_rois : sm_count();
#define B200_ROWS(CW, PROBE)                                                                                      \
  do {                                                                                                            \
    static SmemHighWater hw;                                                                                      \
    int rc = ensure_dynamic_smem(roi_align_fwd_rows<CW, PROBE>, kRowsSmem, &hw, "roi_align rows: smem");          \
    if (rc != B200_OK) return rc;                                                                                 \
    roi_align_fwd_rows<CW, PROBE><<<(unsigned)grid, kConsThreads + 32 + 32 * CW, kRowsSmem, st>>>(                \
        lt, rois, (long long)n_rois, out, out_mean, out_levels);                                                  \
  } while (0)
  // variant (tuning hook of b200_debug_set): bit 5 = four copy warps instead of two, bit 6 = copy-engine probe
  if (variant & 64) B200_ROWS(2, 1);
  else if (variant & 32) B200_ROWS(4, 0);
  else B200_ROWS(2, 0);
#undef B200_ROWS
  B200_CHECK_LAUNCH("roi_align_fwd_rows");
  return B200_OK;
}

}  // namespace b200
