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
constexpr int kRingBytes = 167 * 1024;           // byte ring of tap rows (with the tables: all of the 227 KB)
// kPxB (template parameter below): bytes of one pixel, all channels -- 1024 for fp32 maps, 512 for bf16 maps
// (bf16 in, fp32 arithmetic and output: the reference's amp.float_function semantics on bf16-rounded inputs,
// layers/roi_align.py:57); the ring then holds twice as many pixels
constexpr int kNBar = 32;                        // ring entries in flight, at most (barrier pairs)
constexpr int kHist = 64;                        // placement history the planner keeps (> kNBar + kMaxList)
constexpr int kTabs = 4;                         // RoI tables in flight
constexpr int kConsWarps = kP * (kRC / 4 / 32);  // consumers: thread = (4 channels, output column pw): 14 warps
constexpr int kPlanWarps = 2;  // planner warps, alternating RoIs (one warp's ~900 instructions per RoI bounded the kernel)
constexpr int kTileFloats = kRC * kBins;

// A ring entry = one copy of a tap row, feeding output rows b and b + 1 with weights w0, w1 (x 1/4).
// Entries are grouped by b (ascending), so the consumers' accumulator indices are compile-time
// constants; a tap row feeding more than two output rows (bins narrower than ~1.3 pixels) is simply
// listed -- and copied -- once per pair.
// `place`: where the planner put the entry in the byte ring; `bar`: shared address of the entry's "full"
// barrier, bit 31 = the phase parity to wait for (the planner tracks it, so the consumers do no barrier
// arithmetic), bit 30 = the entry starts a WAIT GROUP: the copies of the group's entries all complete on this
// barrier and the consumers wait for it before loading the entry's columns (other entries: no wait); RoiTab::dep: the copy may be issued once every ring entry with a
// sequence number < dep has been given back (entries are given back in order)
struct __align__(16) RingEntry {
  float w0, w1;
  uint32_t place, bar;
};
struct __align__(16) RoiTab {
  int nent, ncols, nruns, level;
  int batch, width, roi, pad1;  // roi: the RoI's index (= its output slot)
  int cx[kP][4];    // byte offset inside a slot of the k-th merged tap column of bin pw
  float wx[kP][4];  // its weight (0 for unused entries)
  int nk[8];        // merged tap columns per bin
  int gend[8];      // entries [gend[b-1], gend[b]) feed output rows b, b + 1
  RingEntry ent[kMaxList + 1];  // (+1: the consumers fetch headers one entry ahead)
  uint32_t dep[kMaxList];
  int egrp[kMaxList];     // first entry of the group entry e belongs to (the group's "full" barrier is that entry's)
                          // | entries in the group << 8
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
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
// four channels of a pixel: 16 bytes of fp32, or 8 bytes of bf16 widened to fp32 (a shift / a mask per value)
template <int kPxB>
__device__ __forceinline__ V4 lds_px(uint32_t a) {
  if (kPxB == kRC * 4) return lds_v4(a);
  uint32_t lo, hi;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a));
  V4 v;
  v.lo = make_float2(__uint_as_float(lo << 16), __uint_as_float(lo & 0xffff0000u));
  v.hi = make_float2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
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

// (kProbe & 2): cycle counters of who waits for whom (b200_debug_rows_stats; scripts/probe_rows_stats.py)
__device__ unsigned long long g_rows_stats[160][16];
struct RowsStats {
  long long full_wait = 0, tab_wait = 0, bar0 = 0, bar1 = 0, nent = 0, full_miss = 0;
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
  uint32_t full_parity;  // bit i: the phase parity the next user of full_bar[i] waits for
  uint32_t pad;
  uint32_t hist_place[kHist], hist_size[kHist];  // placement of the last kHist entries (pixels)
};

template <int kPxB>
__device__ __forceinline__ RoiPlace build_rows_tab(const LevelTable& lt, const float* __restrict__ rois, long long r,
                                                   RoiTab* tb, int lane, PlanScratch* sc, int gmax) {
  const RoiHeader h = load_roi(rois, r, lt);
  RoiPlace pl;
  pl.level = h.level;
  pl.batch = h.batch;
  pl.H = pl.W = 1;
  if (h.level < 0) {  // matches no level: the output rows stay zero (poolers.py:111-119)
    if (lane < 8) tb->gend[lane] = 0;
    if (lane < 8) tb->nk[lane] = 0;
    if (lane < kP * 4) {  // (the consumers still fetch the sentinel entry's columns: valid addresses)
      (&tb->cx[0][0])[lane] = 0;
      (&tb->wx[0][0])[lane] = 0.f;
    }
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
              cxo[s] = idx[k] * kPxB;
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
        // wait groups: runs of at most gmax entries of the output-row group (gmax 1: every entry waits on a
        // barrier of its own -- the round-1 scheme; larger: fewer waits, but the first entry of a run is only
        // consumed when the whole run has landed)
        const int j = dst - nent, j0 = j - j % gmax, cnt = min(gmax, __popc(bal) - j0);
        tb->egrp[dst] = (nent + j0) | (cnt << 8);
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
template <int kPxB>
__device__ __forceinline__ void place_rows(RoiTab* tb, int lane, RingState* rs, uint32_t full_a, uint32_t idle_a) {
  constexpr int kRingPx = kRingBytes / kPxB;
  const int nent = tb->nent;
  // sentinel after the last entry: weights 0, a valid ring address, and the parity of a barrier that is
  // never armed (waiting for the phase "before" the first one passes at once)
  if (lane == 0) {
    tb->ent[nent].w0 = tb->ent[nent].w1 = 0.f;
    tb->ent[nent].place = 0u;
    tb->ent[nent].bar = idle_a | 0xc0000000u;
  }
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
    tb->ent[lane].place = place * (uint32_t)kPxB;
    tb->dep[lane] = dep;
  }
  // "full" barriers: one per GROUP of entries (the entries feeding output rows b, b + 1), namely the barrier of the
  // group's first entry -- the copies of the whole group complete on it and the consumers wait once per group.
  // A barrier is therefore not used on every lap of the sequence numbers: its phase parity is tracked here.
  const uint32_t bit = 1u << (g % (uint32_t)kNBar);
  const bool first = lane < nent && (tb->egrp[lane] & 0xff) == lane;
  const uint32_t par = rs->full_parity;
  const uint32_t used = __reduce_or_sync(0xffffffffu, first ? bit : 0u);
  if (lane < nent)
    tb->ent[lane].bar = (full_a + 8u * (g % (uint32_t)kNBar)) | ((par & bit) ? 0x80000000u : 0u) | (first ? 0x40000000u : 0u);
  __syncwarp();
  if (lane == nent - 1) {
    rs->head = place + s;
    rs->g0 = g0 + (uint32_t)nent;
    rs->full_parity = par ^ used;
  }
}

// One RoI's worth of consumer work for a warp whose bin has NK (2..4) merged tap columns.  The entry
// headers {w0, w1, place, bar} are fetched one entry ahead, the tap columns of the next entry are
// loaded right after the current one has been reduced along x and given back, and only the two
// output rows an entry group can touch are kept as accumulators: row b is complete when group b ends
// and goes to the shared tile at once (`flush`), so the loop over b is a real loop (no unrolling,
// small code, ~70 registers).
template <int NK, int kProbe, int kPxB, typename Flush>
__device__ __forceinline__ void consume_roi(uint32_t ent_a, uint32_t gend_a, uint32_t c0, uint32_t c1, uint32_t c2,
                                            uint32_t c3, float4 wxv, uint32_t empty_off, int lane, Flush&& flush, RowsStats& st) {
  float2 acc0[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, acc1[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  float2 v[4][2];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k][0] = v[k][1] = make_float2(0.f, 0.f);
  auto wait_group = [&](uint32_t bar) {
    if ((kProbe & 2)) {
      const long long t0 = clock64();
      uint32_t ok;
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar & 0x3fffffffu), "r"(bar >> 31)
          : "memory");
      if (!ok) {
        ++st.full_miss;
        mbar_wait_a(bar & 0x3fffffffu, bar >> 31);
      }
      st.full_wait += clock64() - t0;
      ++st.nent;
    } else {
      mbar_wait_a(bar & 0x3fffffffu, bar >> 31);
    }
  };
  auto load_cols = [&](uint32_t off) {
    if (!(kProbe & 1)) {
      V4 t = lds_px<kPxB>(c0 + off);
      v[0][0] = t.lo, v[0][1] = t.hi;
      t = lds_px<kPxB>(c1 + off);
      v[1][0] = t.lo, v[1][1] = t.hi;
      if (NK > 2) {
        t = lds_px<kPxB>(c2 + off);
        v[2][0] = t.lo, v[2][1] = t.hi;
      }
      if (NK > 3) {
        t = lds_px<kPxB>(c3 + off);
        v[3][0] = t.lo, v[3][1] = t.hi;
      }
    }
  };
  uint4 hdr;
  // one entry: x pass over the columns held in v, give the entry back, fetch the next entry's columns (after the
  // wait for its group when it starts one: kWait) while this one is accumulated into the two output rows
  auto entry = [&](uint32_t ea) {
    const uint4 nh = lds_u4(ea + 16u);
    float2 u0 = make_float2(0.f, 0.f), u1 = make_float2(0.f, 0.f);
    if (!(kProbe & 1)) {
      const float2 wx0 = make_float2(wxv.x, wxv.x), wx1 = make_float2(wxv.y, wxv.y);
      u0 = __fmul2_rn(wx0, v[0][0]);
      u1 = __fmul2_rn(wx0, v[0][1]);
      u0 = __ffma2_rn(wx1, v[1][0], u0);
      u1 = __ffma2_rn(wx1, v[1][1], u1);
      if (NK > 2) {
        const float2 wx2 = make_float2(wxv.z, wxv.z);
        u0 = __ffma2_rn(wx2, v[2][0], u0);
        u1 = __ffma2_rn(wx2, v[2][1], u1);
      }
      if (NK > 3) {
        const float2 wx3 = make_float2(wxv.w, wxv.w);
        u0 = __ffma2_rn(wx3, v[3][0], u0);
        u1 = __ffma2_rn(wx3, v[3][1], u1);
      }
    }
    // this entry has been read: give it back, then fetch the next row while this one is accumulated
    __syncwarp();
    if (lane == 0) mbar_arrive_a((hdr.w & 0x3fffffffu) + empty_off);
    if (nh.w & 0x40000000u) wait_group(nh.w);  // the next entry starts a wait group (or is the sentinel)
    load_cols(nh.z);
    const float w0 = __uint_as_float(hdr.x), w1 = __uint_as_float(hdr.y);
    const float2 w0v = make_float2(w0, w0), w1v = make_float2(w1, w1);
    acc0[0] = __ffma2_rn(w0v, u0, acc0[0]);
    acc0[1] = __ffma2_rn(w0v, u1, acc0[1]);
    acc1[0] = __ffma2_rn(w1v, u0, acc1[0]);
    acc1[1] = __ffma2_rn(w1v, u1, acc1[1]);
    hdr = nh;
  };
  // (the entry after the last one is a sentinel whose barrier always passes: no "is there a next entry"
  // branch around the loads, so ptxas keeps the tap columns in place instead of copying them)
  hdr = lds_u4(ent_a);
  wait_group(hdr.w);
  load_cols(hdr.z);
  uint32_t ea = ent_a;
#pragma unroll 1
  for (int b = 0; b < kP; ++b) {
    const uint32_t end_a = ent_a + 16u * (uint32_t)lds_i32(gend_a + 4u * b);
    // the copies of a wait group complete on ONE barrier (its first entry's), waited for before the group's
    // first columns are loaded: inside a group the loop has no barrier in its dependency chain
#pragma unroll 1
    for (; ea < end_a; ea += 16u) entry(ea);
    // (holding completed rows back for two groups, so that the first flush of a RoI never waits for the bulk
    // store of the previous one, was measured 2 % slower: registers and moves cost more than the wait)
    flush(b, acc0[0], acc0[1]);
    acc0[0] = acc1[0];
    acc0[1] = acc1[1];
    acc1[0] = acc1[1] = make_float2(0.f, 0.f);
  }
}

// kCopyWarps: warps issuing the bulk copies (ring entry g belongs to warp g % kCopyWarps -- a single
// warp's dependent instruction stream, ~40 instructions per entry, was what bounded the first version);
// kProbe: 1 = consumers skip the arithmetic (copy-engine throughput probe), 2 = cycle counters (g_rows_stats)
template <int kCopyWarps, int kProbe, int kPxB>
__global__ void __launch_bounds__(32 * kConsWarps + 32 * kPlanWarps + 32 * kCopyWarps, 1)
roi_align_fwd_rows(const LevelTable lt, const float* __restrict__ rois, const int32_t* __restrict__ order,
                   long long n_rois, float* __restrict__ out, float* __restrict__ out_mean,
                   int32_t* __restrict__ out_levels, int flags, int gmax) {
  // flags (probe, b200_debug_set variant bit 8): 1 = the output tile is not stored
  // (no integer round trip on this pointer: the compiler must keep seeing shared-space addresses,
  // or every access below turns into a generic LD.E / ST.E)
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  __shared__ __align__(128) uint64_t full_bar[kNBar], empty_bar[kNBar], tab_full[kTabs], tab_empty[kTabs];
  __shared__ PlanScratch plan_scratch[kPlanWarps];
  __shared__ RingState ring_state;
  __shared__ __align__(8) uint64_t place_turn[kPlanWarps], idle_bar;
  unsigned char* ring = smem_dyn;
  float* tile = reinterpret_cast<float*>(smem_dyn + (size_t)kRingBytes);
  RoiTab* tabs = reinterpret_cast<RoiTab*>(tile + kTileFloats);

  constexpr int kConsThreads = 32 * kConsWarps;
  constexpr int kPlanWarp0 = kConsWarps;               // first of the warps that build the RoI tables
  constexpr int kCopyWarp0 = kConsWarps + kPlanWarps;  // first of the warps that issue the bulk copies
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kNBar; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsWarps);
    }
    for (int s = 0; s < kPlanWarps; ++s) mbar_init(&place_turn[s], 1);
    mbar_init(&idle_bar, 1);
    ring_state.g0 = 0;
    ring_state.head = 0;
    ring_state.full_parity = 0;
    for (int s = 0; s < kTabs; ++s) {
      mbar_init(&tab_full[s], 1);
      mbar_init(&tab_empty[s], kConsWarps + kCopyWarps);  // every reader of a table
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp >= kPlanWarp0 && warp < kCopyWarp0) {
    // ------------------------------- planners: RoI tables --------------------------------
    // planner p builds the tables of the CTA's RoIs n = p, p + kPlanWarps, ...; only the ring placement
    // is sequential across RoIs: it is done holding the placement turn, passed round-robin.
    // `order` (optional): the RoIs are visited in that order (b200_roi_order: image, level, Morton
    // code of the centre), so the 148 RoIs in flight at any time are neighbours in one feature map
    // and share its rows through the L2; outputs still go to the RoI's own slot.
    const int p = warp - kPlanWarp0;
    const uint32_t full_a = smem_u32(full_bar), idle_a = smem_u32(&idle_bar);
    int n = p, k = 0;
    long long s_tab = 0, s_turn = 0;
    const long long s_t0 = (kProbe & 2) ? clock64() : 0;
    for (long long i = blockIdx.x + (long long)p * gridDim.x; i < n_rois;
         i += (long long)kPlanWarps * gridDim.x, n += kPlanWarps, ++k) {
      const long long r = order ? (long long)order[i] : i;
      const int ti = n % kTabs;
      long long t0 = (kProbe & 2) ? clock64() : 0;
      mbar_wait(&tab_empty[ti], (uint32_t)(((n / kTabs) & 1) ^ 1));  // (a fresh barrier passes)
      if ((kProbe & 2)) s_tab += clock64() - t0;
      const RoiPlace pl = build_rows_tab<kPxB>(lt, rois, r, tabs + ti, lane, &plan_scratch[p], gmax);
      if (lane == 0) {
        tabs[ti].roi = (int)r;
        if (out_levels) out_levels[r] = pl.level;
      }
      __syncwarp();
      // my turn: after planner p - 1 placed RoI n - 1 (planner 0's first turn is free)
      t0 = (kProbe & 2) ? clock64() : 0;
      if (p > 0 || k > 0) mbar_wait(&place_turn[p], (uint32_t)((k & 1) ^ (p == 0 ? 1 : 0)));
      if ((kProbe & 2)) s_turn += clock64() - t0;
      place_rows<kPxB>(tabs + ti, lane, &ring_state, full_a, idle_a);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&place_turn[(p + 1) % kPlanWarps]);
        mbar_arrive(&tab_full[ti]);
      }
    }
    if ((kProbe & 2) && p == 0 && lane == 0) {
      g_rows_stats[blockIdx.x][10] = (unsigned long long)s_tab;
      g_rows_stats[blockIdx.x][11] = (unsigned long long)s_turn;
      g_rows_stats[blockIdx.x][12] = (unsigned long long)(clock64() - s_t0);
    }
    return;
  }

  if (warp >= kCopyWarp0) {
    // ------------------------------- copy warps: the row ring ----------------------------
    const int cw = warp - kCopyWarp0;
    uint32_t g0 = 0;  // ring entries of the RoIs before this one
    int n = 0;
    long long s_dep = 0, s_tab = 0, s_ndep = 0;
    const long long s_t0 = (kProbe & 2) ? clock64() : 0;
    for (long long i = blockIdx.x; i < n_rois; i += gridDim.x, ++n) {
      const int ti = n % kTabs;
      const RoiTab* tb = tabs + ti;
      long long t0 = (kProbe & 2) ? clock64() : 0;
      mbar_wait(&tab_full[ti], (uint32_t)((n / kTabs) & 1));
      if ((kProbe & 2)) s_tab += clock64() - t0;
      const int nent = tb->nent, ncols = tb->ncols, nruns = tb->nruns;
      if (nent > 0) {
        const int level = tb->level, width = tb->width;
        const char* gbase = reinterpret_cast<const char*>(lt.data[level]) +
                            (size_t)tb->batch * lt.H[level] * width * kPxB;
        uint32_t my_pos = 0, my_len = 0;
        if (lane < nruns) {
          my_pos = (uint32_t)tb->run_pos[lane] * kPxB;
          my_len = (uint32_t)tb->run_len[lane] * kPxB;
          gbase += (size_t)tb->run_col[lane] * kPxB;
        }

        const uint32_t size = (uint32_t)ncols * kPxB;
        // Entry g belongs to warp g % kCopyWarps, each warp taking its entries in order: a warp is then never a
        // whole lap of the barriers ahead of the releases, which the parity waits below rely on.  The copies of
        // a GROUP of entries complete on one "full" barrier -- its first entry's, armed by that entry's warp with
        // the bytes of the whole group; a later entry of the group may complete before the barrier is armed
        // (the transaction count is then transiently negative while the arrival is still pending).
        // first entry of this RoI that belongs to this warp
        int e = (int)((cw + kCopyWarps - g0 % kCopyWarps) % kCopyWarps);
        for (; e < nent; e += kCopyWarps) {
          const int eg = tb->egrp[e], first = eg & 0xff;
          const uint32_t bi = (g0 + (uint32_t)first) % kNBar;
          const uint32_t place = tb->ent[e].place, dep = tb->dep[e];
          // (dep - 1 >= g - kNBar, so the parity below names one phase unambiguously)
          if ((kProbe & 2)) {
            t0 = clock64();
            if (dep > 0 && !mbar_try_wait_a(smem_u32(&empty_bar[(dep - 1u) % kNBar]), ((dep - 1u) / kNBar) & 1u)) {
              ++s_ndep;
              mbar_wait(&empty_bar[(dep - 1u) % kNBar], ((dep - 1u) / kNBar) & 1u);
            }
            s_dep += clock64() - t0;
          } else if (dep > 0) mbar_wait(&empty_bar[(dep - 1u) % kNBar], ((dep - 1u) / kNBar) & 1u);
          if (lane == 0 && e == first) mbar_arrive_expect_tx(&full_bar[bi], size * (uint32_t)(eg >> 8));
          if (lane < nruns)
            bulk_g2s(ring + place + my_pos, gbase + (size_t)tb->ent_y[e] * width * kPxB, my_len, &full_bar[bi]);
        }
        g0 += (uint32_t)nent;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&tab_empty[ti]);
    }
    if ((kProbe & 2) && cw == 0 && lane == 0) {
      g_rows_stats[blockIdx.x][7] = (unsigned long long)s_dep;
      g_rows_stats[blockIdx.x][8] = (unsigned long long)s_tab;
      g_rows_stats[blockIdx.x][9] = (unsigned long long)(clock64() - s_t0);
      g_rows_stats[blockIdx.x][13] = (unsigned long long)s_ndep;
    }
    return;
  }

  // --------------------------------- consumers --------------------------------------------
  // thread = (channel quad q, output column pw): two warps per pw, a warp's lanes read 512 contiguous
  // bytes of a pixel (conflict-free LDS.128)
  const int pw = warp >> 1;
  const int q = (warp & 1) * 32 + lane;
  const int rot = (lane >> 3) & 3;
  // 32-bit shared addresses, pinned in registers (ptxas otherwise re-derives them from the CTA's
  // shared window at every use)
  uint32_t ring_a = smem_u32(ring) + q * (kPxB / 64), tabs_a = smem_u32(tabs);
  uint32_t empty_off = smem_u32(empty_bar) - smem_u32(full_bar);
  // tile slots of the thread's four channels, rotated by lane / 8 so that the 32 lanes of a store hit
  // 32 different banks (thread stride 4 * 49 floats = 4 mod 32, channel stride 49 = 17 mod 32)
  uint32_t tc0, tc1, tc2, tc3;
  {
    const uint32_t t0 = smem_u32(tile) + 4u * (uint32_t)((4 * q) * kBins + pw);
    tc0 = t0 + 4u * (uint32_t)(((0 + rot) & 3) * kBins);
    tc1 = t0 + 4u * (uint32_t)(((1 + rot) & 3) * kBins);
    tc2 = t0 + 4u * (uint32_t)(((2 + rot) & 3) * kBins);
    tc3 = t0 + 4u * (uint32_t)(((3 + rot) & 3) * kBins);
  }
  asm volatile("" : "+r"(ring_a), "+r"(tabs_a), "+r"(tc0), "+r"(tc1), "+r"(tc2), "+r"(tc3), "+r"(empty_off));
  const uint64_t store_policy = policy_evict_first();
  // output row b of this thread's (4 channels, pw) is complete: into the tile.  Before the first row
  // of a RoI the bulk store of the previous RoI must have read the tile.
  RowsStats st;
  const long long s_t0 = (kProbe & 2) ? clock64() : 0;
  auto flush = [&](int b, float2 lo, float2 hi) {
    if (b == 0) {
      const long long t0 = (kProbe & 2) ? clock64() : 0;
      if (tid == 0) bulk_wait_read();
      bar_consumers<kConsThreads>();
      if ((kProbe & 2)) st.bar0 += clock64() - t0;
    }
    float a = lo.x, bb = lo.y, c = hi.x, d = hi.y;
    if (rot & 1) {
      const float t = a;
      a = bb;
      bb = c;
      c = d;
      d = t;
    }
    if (rot & 2) {
      float t = a;
      a = c;
      c = t;
      t = bb;
      bb = d;
      d = t;
    }
    const uint32_t o = 4u * (uint32_t)(b * kP);
    sts_f32(tc0 + o, a);
    sts_f32(tc1 + o, bb);
    sts_f32(tc2 + o, c);
    sts_f32(tc3 + o, d);
  };
  int n = 0;
  for (long long i = blockIdx.x; i < n_rois; i += gridDim.x, ++n) {
    const int ti = n % kTabs;
    const RoiTab* tb = tabs + ti;
    const uint32_t tb_a = tabs_a + (uint32_t)ti * (uint32_t)sizeof(RoiTab);
    long long t0 = (kProbe & 2) ? clock64() : 0;
    mbar_wait(&tab_full[ti], (uint32_t)((n / kTabs) & 1));
    if ((kProbe & 2)) st.tab_wait += clock64() - t0;
    const long long r = tb->roi;
    const int4 cxv = *reinterpret_cast<const int4*>(tb->cx[pw]);
    const float4 wxv = *reinterpret_cast<const float4*>(tb->wx[pw]);
    const int nk = tb->nk[pw];
    const uint32_t c0 = ring_a + cxv.x, c1 = ring_a + cxv.y, c2 = ring_a + cxv.z, c3 = ring_a + cxv.w;
    const uint32_t ent_a = tb_a + (uint32_t)offsetof(RoiTab, ent), gend_a = tb_a + (uint32_t)offsetof(RoiTab, gend);
    // (nk is the same for the whole warp: one specialised loop per RoI, no per-entry branches)
    if (nk > 3) consume_roi<4, kProbe, kPxB>(ent_a, gend_a, c0, c1, c2, c3, wxv, empty_off, lane, flush, st);
    else if (nk > 2) consume_roi<3, kProbe, kPxB>(ent_a, gend_a, c0, c1, c2, c3, wxv, empty_off, lane, flush, st);
    else consume_roi<2, kProbe, kPxB>(ent_a, gend_a, c0, c1, c2, c3, wxv, empty_off, lane, flush, st);
    // the tables of this RoI are no longer needed
    __syncwarp();
    if (lane == 0) mbar_arrive(&tab_empty[ti]);

    // ---- the [256 x 49] tile is complete: one contiguous 50 KB block of the NCHW output, a single
    // bulk store (TMA) that drains while the consumers are already in the next RoI's rows ----
    fence_proxy_async();  // this thread's tile writes -> visible to the async proxy
    t0 = (kProbe & 2) ? clock64() : 0;
    bar_consumers<kConsThreads>();
    if ((kProbe & 2)) st.bar1 += clock64() - t0;
    if (tid == 0 && !(flags & 1)) bulk_s2g(out + (size_t)r * kTileFloats, smem_u32(tile), kTileFloats * 4, store_policy);
    if (out_mean && tid < kRC) {
      // seven independent row sums, then their sum: a 49-long dependent chain of additions held eight of the
      // fourteen consumer warps back for ~300 cycles per RoI while the others waited at the next RoI's first
      // flush (this kernel is the fast-math path: the order of an fp32 sum is free within its 1e-5)
      const float* row = tile + tid * kBins;
      float rs[kP];
#pragma unroll
      for (int b = 0; b < kP; ++b) rs[b] = row[b * kP];
#pragma unroll
      for (int j = 1; j < kP; ++j)
#pragma unroll
        for (int b = 0; b < kP; ++b) rs[b] += row[b * kP + j];
      const float s = ((rs[0] + rs[1]) + (rs[2] + rs[3])) + ((rs[4] + rs[5]) + rs[6]);
      out_mean[(size_t)r * kRC + tid] = s * (1.0f / (float)kBins);
    }
  }
  if (tid == 0) bulk_wait_all();  // the last store must have left shared memory before the CTA exits
  if ((kProbe & 2) && lane == 0 && (warp == 0 || warp == kConsWarps - 1)) {
    unsigned long long* o = g_rows_stats[blockIdx.x + (warp == 0 ? 0 : 0)];
    if (warp == 0) {
      o[0] = (unsigned long long)(clock64() - s_t0);
      o[1] = (unsigned long long)st.full_wait;
      o[2] = (unsigned long long)st.tab_wait;
      o[3] = (unsigned long long)st.bar0;
      o[4] = (unsigned long long)st.bar1;
      o[5] = (unsigned long long)st.nent;
      o[6] = (unsigned long long)st.full_miss;
    } else {
      o[14] = (unsigned long long)st.full_wait;
      o[15] = (unsigned long long)(st.bar0 + st.bar1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Visiting order of the RoIs for the row-streaming kernel: chunks of kOrderChunk consecutive RoIs are
// bucketed by (image, FPN level, cell of a 16 x 16 grid over the level's map in Morton order) with a
// counting sort in shared memory -- histogram with shared atomics, one scan, one scatter (~3 us; a
// bitonic sort of the same chunk took 42 us).  The order inside a bucket is whatever the atomics give:
// it only decides which SM takes which RoI when, never a result.
// ---------------------------------------------------------------------------------------------
constexpr int kOrderChunk = 2048, kOrderThreads = 1024, kOrderBuckets = 4096;

__global__ void __launch_bounds__(kOrderThreads) roi_order_kernel(const LevelTable lt, const float* __restrict__ rois,
                                                                  long long n_rois, int32_t* __restrict__ order) {
  __shared__ int hist[kOrderBuckets];
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long base = (long long)blockIdx.x * kOrderChunk;
  const int n = (int)min((long long)kOrderChunk, n_rois - base);
  for (int i = tid; i < kOrderBuckets; i += kOrderThreads) hist[i] = 0;
  __syncthreads();
  const int img0 = (int)rois[base * 5];  // buckets are relative to the chunk's first image
  int bucket[kOrderChunk / kOrderThreads], slot[kOrderChunk / kOrderThreads];
#pragma unroll
  for (int k = 0; k < kOrderChunk / kOrderThreads; ++k) {
    const int i = tid + k * kOrderThreads;
    bucket[k] = -1;
    if (i < n) {
      const RoiHeader h = load_roi(rois, base + i, lt);
      const int l = h.level < 0 ? 0 : h.level;
      const float sx = lt.scale[l] * 16.f / (float)lt.W[l], sy = lt.scale[l] * 16.f / (float)lt.H[l];
      const int cx = (int)fminf(fmaxf((h.x1 + h.x2) * 0.5f * sx, 0.f), 15.f), cy = (int)fminf(fmaxf((h.y1 + h.y2) * 0.5f * sy, 0.f), 15.f);
      int m = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) m |= (((cx >> b) & 1) << (2 * b)) | (((cy >> b) & 1) << (2 * b + 1));
      bucket[k] = ((((h.batch - img0) & 3) << 2 | ((h.level + 1) & 3)) << 8) | m;
      slot[k] = atomicAdd(&hist[bucket[k]], 1);
    }
  }
  __syncthreads();
  // exclusive scan of the 4096 counters: 4 per thread, warp scan, scan of the warp totals
  int c[4], run = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    c[j] = hist[4 * tid + j];
    run += c[j];
  }
  int incl = run;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += t;
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  int ex = incl - run + (warp > 0 ? warp_tot[warp - 1] : 0);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    hist[4 * tid + j] = ex;
    ex += c[j];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kOrderChunk / kOrderThreads; ++k)
    if (bucket[k] >= 0) order[base + hist[bucket[k]] + slot[k]] = (int32_t)(base + tid + k * kOrderThreads);
}

}  // namespace

bool rows_kernel_applies(const LevelTable& lt, int C, int PH, int PW) {
  (void)lt;
  return C == kRC && PH == kP && PW == kP;
}

// Preconditions (checked by the caller): NHWC, sampling_ratio 2; rows_kernel_applies().
int launch_forward_rows(const LevelTable& lt, int C, bool bf16_maps, const float* rois, int64_t n_rois, float* out,
                        float* out_mean, int32_t* out_levels, int32_t* order_ws, int variant, cudaStream_t st) {
  B200_REQUIRE(C == kRC, "roi_align rows kernel: %d channels", C);
  B200_REQUIRE(n_rois < ((int64_t)1 << 31), "roi_align rows kernel: too many RoIs");
  // variant (tuning hook of b200_debug_set): bit 5 = visit the RoIs in the order given (no b200 ordering pass),
  // bit 6 = copy-engine probe (no arithmetic)
  const int32_t* order = nullptr;
  if (order_ws && !(variant & 32) && n_rois > 1) {
    roi_order_kernel<<<(unsigned)ceil_div<int64_t>(n_rois, kOrderChunk), kOrderThreads, 0, st>>>(lt, rois, (long long)n_rois,
                                                                                              order_ws);
    B200_CHECK_LAUNCH("roi_order_kernel");
    order = order_ws;
  }
  const int64_t grid = n_rois < sm_count() ? n_rois : sm_count();
  // entries per wait group (build_rows_tab): measured best 1 for fp32 maps (the ring holds ~9 rows: waiting for
  // several to land stalls the consumers), whole output-row groups for bf16 maps (twice the rows in flight);
  // variant bits 10-12 override (tuning)
  const int gmax = ((variant >> 10) & 7) ? ((variant >> 10) & 7) : (bf16_maps ? 4 : 1);
#define B200_ROWS(CW, PROBE, PXB)                                                                               \
  do {                                                                                                          \
    static SmemHighWater hw;                                                                                    \
    static int static_smem = -1;                                                                                \
    if (static_smem < 0) {                                                                                      \
      cudaFuncAttributes fa;                                                                                    \
      int rc0 = check_cuda(cudaFuncGetAttributes(&fa, roi_align_fwd_rows<CW, PROBE, PXB>), "roi_align rows");   \
      if (rc0 != B200_OK) return rc0;                                                                           \
      static_smem = (int)fa.sharedSizeBytes;                                                                    \
    }                                                                                                           \
    B200_REQUIRE((size_t)static_smem + kRowsSmem <= (size_t)227 * 1024,                                         \
                 "roi_align rows kernel: %d + %d bytes of shared memory", static_smem, (int)kRowsSmem);         \
    int rc = ensure_dynamic_smem(roi_align_fwd_rows<CW, PROBE, PXB>, kRowsSmem, &hw, "roi_align rows: smem");   \
    if (rc != B200_OK) return rc;                                                                               \
    roi_align_fwd_rows<CW, PROBE, PXB><<<(unsigned)grid, 32 * kConsWarps + 32 * kPlanWarps + 32 * CW, kRowsSmem, st>>>( \
        lt, rois, order, (long long)n_rois, out, out_mean, out_levels, (variant >> 8) & 1, gmax);               \
  } while (0)
  // (L2 prefetch of the planned rows -- bulk prefetches from the planner or one RoI ahead from a copy warp,
  // and LSU-side prefetch.global.L2 -- was measured 10-45 % SLOWER in all three forms: DESIGN.md)
  // bit 7 (tuning): two copy warps instead of four
  if (bf16_maps) {
    if (variant & 128) B200_ROWS(2, 0, kRC * 2);
    else B200_ROWS(4, 0, kRC * 2);
  } else if (variant & 64) B200_ROWS(4, 1, kRC * 4);
  else if (variant & 512) B200_ROWS(4, 2, kRC * 4);
  else if (variant & 128) B200_ROWS(2, 0, kRC * 4);
  else B200_ROWS(4, 0, kRC * 4);
#undef B200_ROWS
  B200_CHECK_LAUNCH("roi_align_fwd_rows");
  return B200_OK;
}

// the visiting order alone (also used by the backward: CTAs in this order scatter into neighbouring pixels,
// so the reductions mostly meet lines that are still in the L2)
int launch_roi_order(const LevelTable& lt, const float* rois, int64_t n_rois, int32_t* order, cudaStream_t st) {
  B200_REQUIRE(n_rois < ((int64_t)1 << 31), "roi order: too many RoIs");
  if (n_rois <= 0) return B200_OK;
  roi_order_kernel<<<(unsigned)ceil_div<int64_t>(n_rois, kOrderChunk), kOrderThreads, 0, st>>>(lt, rois, (long long)n_rois, order);
  B200_CHECK_LAUNCH("roi_order_kernel");
  return B200_OK;
}

extern "C" int b200_debug_rows_stats(unsigned long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_rows_stats, sizeof(g_rows_stats)) == cudaSuccess ? 0 : 1;
}

size_t rows_order_workspace_bytes(int64_t n_rois) { return n_rois > 0 ? sizeof(int32_t) * (size_t)n_rois : 0; }

}  // namespace b200
