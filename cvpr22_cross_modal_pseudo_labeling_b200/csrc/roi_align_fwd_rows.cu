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
constexpr int kSlotBytes = kMaxList * kPxBytes;  // one tap row: the tapped columns, compacted
constexpr int kSlots = 6;                        // ring of tap rows
constexpr int kRingBytes = kSlots * kSlotBytes;
constexpr int kTabs = 4;                         // RoI tables in flight
constexpr int kMaxCopyWarps = 4;
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
struct __align__(16) RoiTab {
  int nent, ncols, nruns, level;
  int batch, width, pad0, pad1;
  int cx[kP][4];    // byte offset inside a slot of the k-th merged tap column of bin pw
  float wx[kP][4];  // its weight (0 for unused entries)
  int nk[8];        // merged tap columns per bin
  int gend[8];      // entries [gend[b-1], gend[b]) feed output rows b, b + 1
  float2 ent_w[kMaxList];
  int ent_y[kMaxList];    // feature row of entry e
  float wy[kMaxList][8];  // weight (x 1/4) of list row i for output row ph
  int rowy[kMaxList];     // feature row of list entry i
  int colx[kMaxList];     // feature column of list entry j
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
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(kConsThreads) : "memory"); }

struct RoiPlace {
  int level, batch, H, W;
};

// Producer warp: axis tables of RoI r.  Lanes 0-13 own the y samples, lanes 16-29 the x samples.
__device__ __forceinline__ RoiPlace build_rows_tab(const LevelTable& lt, const float* __restrict__ rois, long long r,
                                                   RoiTab* tb, int lane) {
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
  int* list = isx ? tb->colx : tb->rowy;
  if (hl < kNS) {
    if (nnew >= 1) list[scan - 1] = hi;
    if (nnew == 2) list[scan - 2] = lo;
  }
  // zero the row-weight table
  float* wyf = &tb->wy[0][0];
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
        if (w[k] != 0.f) tb->wy[idx[k]][hl] = 0.25f * w[k];
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
        if (tb->wy[lane][ph] != 0.f) mask |= 1u << ph;
    }
    unsigned pk = 0;
#pragma unroll
    for (int b = 0; b < kP; ++b)
      if ((mask >> b) & 1u) {
        pk |= 1u << b;
        mask &= ~(3u << b);
      }
    const int y = lane < nrows ? tb->rowy[lane] : 0;
#pragma unroll
    for (int b = 0; b < kP; ++b) {
      const bool f = (pk >> b) & 1u;
      const unsigned bal = __ballot_sync(0xffffffffu, f);
      if (f) {
        const int dst = nent + __popc(bal & ((1u << lane) - 1u));
        tb->ent_y[dst] = y;
        tb->ent_w[dst] = make_float2(tb->wy[lane][b], b + 1 < kP ? tb->wy[lane][b + 1] : 0.f);
      }
      nent += __popc(bal);
      if (lane == 0) tb->gend[b] = nent;
    }
  }
  // runs of consecutive tapped columns: one bulk copy each
  {
    const int x = lane < ncols ? tb->colx[lane] : 0;
    const int xp = (lane > 0 && lane < ncols) ? tb->colx[lane - 1] : 0;
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
  __shared__ __align__(128) uint64_t full_bar[kSlots], empty_bar[kSlots], tab_full[kTabs], tab_empty[kTabs];
  unsigned char* ring = smem_dyn;
  float* tile = reinterpret_cast<float*>(smem_dyn + (size_t)kRingBytes);
  RoiTab* tabs = reinterpret_cast<RoiTab*>(tile + kTileFloats);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kSlots; ++s) {
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
    int n = 0;
    for (long long r = blockIdx.x; r < n_rois; r += gridDim.x, ++n) {
      const int ti = n % kTabs;
      mbar_wait(&tab_empty[ti], (uint32_t)(((n / kTabs) & 1) ^ 1));  // (a fresh barrier passes)
      const RoiPlace pl = build_rows_tab(lt, rois, r, tabs + ti, lane);
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
          const uint32_t g = g0 + (uint32_t)i, slot = g % kSlots;
          mbar_wait(&empty_bar[slot], ((g / kSlots) & 1u) ^ 1u);
          if (lane == 0) mbar_arrive_expect_tx(&full_bar[slot], size);
          if (lane < nruns)
            bulk_g2s(ring + slot * kSlotBytes + my_pos, gbase + (size_t)tb->ent_y[i] * width * kPxBytes, my_len,
                     &full_bar[slot]);
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
    auto load_row = [&]() {
      const uint32_t bi = g % kSlots;
      mbar_wait_a(full_a + 8u * bi, (g / kSlots) & 1u);
      const uint32_t off = bi * kSlotBytes;
      if (kProbe == 0) {
        v0 = lds_v4(c0 + off);
        v1 = lds_v4(c1 + off);
        if (nk > 2) v2 = lds_v4(c2 + off);
        if (nk > 3) v3 = lds_v4(c3 + off);
      }
      ++g;
    };
    if (nent > 0) load_row();
    uint32_t ent_a = tb_a + (uint32_t)offsetof(RoiTab, ent_w);
    const uint32_t ent_last = ent_a + 8u * (uint32_t)(nent - 1);
#pragma unroll
    for (int b = 0; b < kP; ++b) {
      const uint32_t end_a = tb_a + (uint32_t)offsetof(RoiTab, ent_w) + 8u * (uint32_t)lds_i32(tb_a + (uint32_t)offsetof(RoiTab, gend) + 4u * b);
      for (; ent_a < end_a; ent_a += 8) {
        // x pass: the tap columns reduced to one value per channel
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
        const float2 w = lds_f2(ent_a);
        // the registers are free again: fetch the next row while this one is accumulated
        if (ent_a < ent_last) load_row();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(empty_a + 8u * (rg % kSlots));  // this entry has been read: give it back
        ++rg;
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
    bar_consumers();  // the copy-out of the previous RoI has left the tile
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
    bar_consumers();
    {
      const float4* src = reinterpret_cast<const float4*>(tile);
      float4* dst = reinterpret_cast<float4*>(out + (size_t)r * kTileFloats);
#pragma unroll
      for (int i = 0; i < kTileFloats / 4 / kConsThreads; ++i) __stcs(dst + tid + i * kConsThreads, src[tid + i * kConsThreads]);
      if (out_mean && tid < kRC) {
        const float* row = tile + tid * kBins;
        float s = 0.f;
#pragma unroll 7
        for (int i = 0; i < kBins; ++i) s = __fadd_rn(s, row[i]);
        out_mean[(size_t)r * kRC + tid] = __fdiv_rn(s, (float)kBins);
      }
    }
  }
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
    int rc = ensure_dynamic_smem(roi_align_fwd_rows<CW, PROBE>, kRowsSmem, &hw, "roi_align rows: smem attribute"); \
    if (rc != B200_OK) return rc;                                                                                 \
    roi_align_fwd_rows<CW, PROBE><<<(unsigned)grid, kConsThreads + 32 + 32 * CW, kRowsSmem, st>>>(                \
        lt, rois, (long long)n_rois, out, out_mean, out_levels);                                                  \
  } while (0)
  // variant (tuning hook of b200_debug_set): bit 5 = two copy warps instead of four, bit 6 = copy-engine probe
  if (variant & 64) B200_ROWS(4, 1);
  else if (variant & 32) B200_ROWS(2, 0);
  else B200_ROWS(4, 0);
#undef B200_ROWS
  B200_CHECK_LAUNCH("roi_align_fwd_rows");
  return B200_OK;
}

}  // namespace b200
