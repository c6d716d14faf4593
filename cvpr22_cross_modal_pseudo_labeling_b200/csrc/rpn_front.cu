// rpn_front.cu -- the RPN post-processor's per-level front end as one launch.
//
// Replaces, for every (image, level) pair at once, RPNPostProcessor.forward_for_single_feature_map
// up to the NMS (reference modeling/rpn/inference.py:76-110): permute_and_flatten
// (modeling/rpn/utils.py:10-14), sigmoid, topk(pre_nms_top_n, sorted), the gathers of the
// regression and the anchors, BoxCoder.decode (modeling/box_coder.py:52-95) and
// clip_to_image(remove_empty=False) (structures/bounding_box.py:214-219).  Output = the
// candidate boxes / scores of b200_nms_batched, segment (image n, level l), descending score.
//
// One CTA per (level, image).  Goal: the k entries with the largest (logit, -flattened index),
// sorted descending -- i.e. a stable descending sort of the flattened logits cut at k.
//   1. Candidates: a strided sample of 4096 logits is sorted in shared memory and its quantile
//      that should leave about (k + capacity)/2 elements above it becomes a lower bound; ONE
//      coalesced pass over the level's A*H*W logits appends every element >= the bound, packed
//      as (orderable(logit) << 32) | ~flattened index, to a shared-memory list (warp-aggregated).
//      If the list holds < k or > capacity entries (a badly skewed sample, masses of equal
//      logits) the exact fallback runs instead: bisection on the order-preserving integer image
//      of the floats for the k-th largest logit (32 counting passes, no atomics), and, if the
//      ties at that value do not fit, a second bisection on the flattened index among them.
//   2. The list is bitonic-sorted descending in shared memory; its first k entries are the result.
//   3. Each of them is decoded from its 4 regression planes and its anchor, clipped and written
//      with sigmoid(logit).
// sigmoid is monotonic, so selecting on logits equals selecting on probabilities except where
// two different logits round to the same probability (then the order inside that tie may
// differ from torch.topk's, whose tie order is unspecified anyway).
#include "common.cuh"

namespace b200 {
namespace {

constexpr float kBboxXformClip = 4.135166556742356f;  // log(1000 / 16), box_coder.py:22
constexpr int kRpnThreads = 1024;
constexpr int kRpnMaxLevels = B200_MAX_LEVELS;

struct RpnLevels {
  const float* obj[kRpnMaxLevels];      // [N, A, H, W]
  const float* reg[kRpnMaxLevels];      // [N, 4A, H, W]
  const float* anchors[kRpnMaxLevels];  // [M, 4] shared by the images, or [N, M, 4]
  long long anchor_stride[kRpnMaxLevels];  // floats between images (0: shared)
  int A[kRpnMaxLevels], HW[kRpnMaxLevels];
  int k[kRpnMaxLevels];      // min(pre_nms_top_n, A*H*W)
  int start[kRpnMaxLevels];  // first slot of the level inside an image's K slots
  int K;                     // slots per image
};

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t orderable_bits(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ int block_sum(int v, int* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();  // scratch free
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  int t = lane < kRpnThreads / 32 ? scratch[lane] : 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  return t;  // same value in every thread
}

constexpr int kRpnSample = 4096;

// Ordered, atomic-free collection of the logits that satisfy `take(key, flat)` into cand[]:
// pass A counts per warp, a scan over the 32 warp totals gives every warp its base, pass B
// re-reads the (L2-resident) logits and writes.  A single shared counter with warp-aggregated
// atomics serialised the 32 warps (6 k atomics on one address per level-2 CTA: 0.2 ms).
// Returns the total; nothing is written when it exceeds `cap`.
constexpr int kRpnUnroll = 8;  // logits fetched per thread before any is consumed (hides the L2 latency)

template <class Take>
__device__ __forceinline__ int collect(const float* __restrict__ obj, int M, int HW, int A, u64* cand, int cap,
                                       int* scratch, Take take) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kStep = kRpnThreads * kRpnUnroll;
  const int m_round = (M + kStep - 1) / kStep * kStep;
  int c = 0;
  for (int i0 = 0; i0 < m_round; i0 += kStep) {
    float v[kRpnUnroll];
#pragma unroll
    for (int u = 0; u < kRpnUnroll; ++u) {
      const int i = i0 + u * kRpnThreads + tid;
      v[u] = i < M ? obj[i] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kRpnUnroll; ++u) {
      const int i = i0 + u * kRpnThreads + tid;
      bool t = false;
      if (i < M) {
        const int a = i / HW, p = i - a * HW;
        t = take(orderable_bits(v[u]), (uint32_t)(p * A + a));
      }
      c += __popc(__ballot_sync(0xffffffffu, t));
    }
  }
  __syncthreads();  // scratch free
  if (lane == 0) scratch[warp] = c;
  __syncthreads();
  int w = lane < kRpnThreads / 32 ? scratch[lane] : 0, incl = w;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  int base = __shfl_sync(0xffffffffu, incl - w, warp);
  if (total > cap) return total;
  for (int i0 = 0; i0 < m_round; i0 += kStep) {
    float v[kRpnUnroll];
#pragma unroll
    for (int u = 0; u < kRpnUnroll; ++u) {
      const int i = i0 + u * kRpnThreads + tid;
      v[u] = i < M ? obj[i] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kRpnUnroll; ++u) {
      const int i = i0 + u * kRpnThreads + tid;
      bool t = false;
      u64 pk = 0ull;
      if (i < M) {
        const int a = i / HW, p = i - a * HW;
        const uint32_t key = orderable_bits(v[u]), flat = (uint32_t)(p * A + a);  // (h*W + w)*A + a
        t = take(key, flat);
        pk = ((u64)key << 32) | (u64)(0xffffffffu - flat);
      }
      const uint32_t tm = __ballot_sync(0xffffffffu, t);
      if (t) cand[base + __popc(tm & ((1u << lane) - 1))] = pk;
      base += __popc(tm);
    }
  }
  return total;
}

__device__ __forceinline__ void bitonic_desc(u64* keys, int n) {  // n: power of two
  for (int kk = 2; kk <= n; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += kRpnThreads) {
        const int lo_i = 2 * t - (t & (j - 1)), hi_i = lo_i + j;
        const u64 x = keys[lo_i], y = keys[hi_i];
        if ((x < y) == ((lo_i & kk) == 0)) {
          keys[lo_i] = y;
          keys[hi_i] = x;
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kRpnThreads)
rpn_front_kernel(const RpnLevels lv, const float* __restrict__ im_sizes, float wx, float wy, float ww, float wh,
                 int cap, int force_exact, float4* __restrict__ boxes, float* __restrict__ scores) {
  extern __shared__ __align__(16) unsigned char rpn_smem[];
  u64* cand = reinterpret_cast<u64*>(rpn_smem);  // [cap] (key << 32) | (0xffffffff - flattened index)
  __shared__ int scratch[32];
  const int l = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  const int A = lv.A[l], HW = lv.HW[l], M = A * HW, k = lv.k[l];
  const float* obj = lv.obj[l] + (size_t)n * M;

  // ---- 1. candidates ----------------------------------------------------------------------
  int nc = -1;  // entries in cand[] (uniform); -1: not collected yet
  if (M <= cap) {  // small level: everything is a candidate
    nc = collect(obj, M, HW, A, cand, cap, scratch, [](uint32_t, uint32_t) { return true; });
  } else if (!force_exact) {
    // sample -> lower bound that should leave ~ (k + cap) / 2 elements above it
    u64* sk = cand + kRpnSample;  // sort the samples in the upper part of the buffer (cap >= 8192)
    for (int t = tid; t < kRpnSample; t += kRpnThreads)
      sk[t] = orderable_bits(obj[(int)(((long long)t * M) / kRpnSample)]);
    __syncthreads();
    bitonic_desc(sk, kRpnSample);
    const long long target = ((long long)k + cap) / 2;
    int q = (int)((target * kRpnSample + M - 1) / M);
    if (q > kRpnSample - 1) q = kRpnSample - 1;
    const uint32_t lo0 = (uint32_t)sk[q];
    __syncthreads();
    nc = collect(obj, M, HW, A, cand, cap, scratch, [lo0](uint32_t key, uint32_t) { return key >= lo0; });
    if (nc < k || nc > cap) nc = -1;  // uniform
  }
  if (nc < 0) {
    // exact fallback: k-th largest key by bisection (invariant: count(key >= lo) >= k)
    uint32_t lo = 0u, hi = 0xffffffffu;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1) + 1u;
      int c = 0;
      for (int i = tid; i < M; i += kRpnThreads) c += orderable_bits(obj[i]) >= mid;
      if (block_sum(c, scratch) >= k) lo = mid;
      else hi = mid - 1u;
    }
    const uint32_t thr = lo;
    int cg = 0, ce = 0;
    for (int i = tid; i < M; i += kRpnThreads) {
      const uint32_t key = orderable_bits(obj[i]);
      cg += key > thr;
      ce += key == thr;
    }
    const int count_gt = block_sum(cg, scratch), count_eq = block_sum(ce, scratch);
    const int need_eq = k - count_gt;
    uint32_t flat_max = 0xffffffffu;  // ties at thr are taken while their flattened index <= flat_max
    if (count_gt + count_eq > cap) {
      // too many equal logits for the list: take the need_eq of them with the lowest flattened
      // index (second bisection; flattened indices are unique, so the count lands exactly)
      uint32_t flo = 0u, fhi = (uint32_t)M - 1u;  // invariant: count(eq && flat <= fhi) >= need_eq
      while (flo < fhi) {
        const uint32_t fmid = flo + ((fhi - flo) >> 1);
        int c = 0;
        for (int i = tid; i < M; i += kRpnThreads) {
          const int a = i / HW, p = i - a * HW;
          c += orderable_bits(obj[i]) == thr && (uint32_t)(p * A + a) <= fmid;
        }
        if (block_sum(c, scratch) >= need_eq) fhi = fmid;
        else flo = fmid + 1u;
      }
      flat_max = flo;
    }
    nc = collect(obj, M, HW, A, cand, cap, scratch,
                 [thr, flat_max](uint32_t key, uint32_t flat) { return key > thr || (key == thr && flat <= flat_max); });
  }
  __syncthreads();
  // ---- 2. trim the list to exactly k entries, then sort them ------------------------------
  // The bitonic sort is shared-memory-bandwidth bound (log^2 passes over the list), so it runs
  // on the k winners only: the k-th largest packed key (all keys are distinct) is found by
  // bisection over the list in shared memory, and the entries >= it are compacted in place.
  if (nc > k) {
    u64 lo = 0ull, hi = ~0ull;  // invariant: count(cand >= lo) >= k
    while (lo < hi) {
      const u64 mid = lo + ((hi - lo) >> 1) + 1ull;
      int c = 0;
      for (int i = tid; i < nc; i += kRpnThreads) c += cand[i] >= mid;
      if (block_sum(c, scratch) >= k) lo = mid;
      else hi = mid - 1ull;
    }
    const u64 kth = lo;
    int written = 0;  // uniform
    for (int i0 = 0; i0 < nc; i0 += kRpnThreads) {
      const int i = i0 + tid;
      const u64 e = i < nc ? cand[i] : 0ull;
      const bool t = i < nc && e >= kth;
      const uint32_t tm = __ballot_sync(0xffffffffu, t);
      __syncthreads();  // every entry of this chunk is in a register; scratch free
      if ((tid & 31) == 0) scratch[tid >> 5] = __popc(tm);
      __syncthreads();
      int before = 0, chunk = 0;
      for (int w = 0; w < kRpnThreads / 32; ++w) {
        const int cw = scratch[w];
        before += w < (tid >> 5) ? cw : 0;
        chunk += cw;
      }
      if (t) cand[written + before + __popc(tm & ((1u << (tid & 31)) - 1))] = e;  // dst <= src
      written += chunk;
    }
    __syncthreads();
    nc = k;
  }
  int npad = 2;
  while (npad < nc) npad <<= 1;
  for (int i = nc + tid; i < npad; i += kRpnThreads) cand[i] = 0ull;  // padding sorts last
  __syncthreads();
  bitonic_desc(cand, npad);
  const u64* keys = cand;
  // ---- 3. decode + clip + sigmoid --------------------------------------------------------
  const float img_w = im_sizes[2 * n], img_h = im_sizes[2 * n + 1];
  const float mx = __fsub_rn(img_w, 1.0f), my = __fsub_rn(img_h, 1.0f);
  const float* reg = lv.reg[l] + (size_t)n * 4 * M;
  const float4* anc = reinterpret_cast<const float4*>(lv.anchors[l] + (size_t)n * lv.anchor_stride[l]);
  const size_t out0 = (size_t)n * lv.K + lv.start[l];
  for (int j = tid; j < k; j += kRpnThreads) {
    const u64 e = keys[j];
    const uint32_t flat = 0xffffffffu - (uint32_t)e;
    const float logit = from_orderable((uint32_t)(e >> 32));
    const int p = flat / A, a = flat - p * A;
    const float* r = reg + (size_t)(4 * a) * HW + p;  // channel a*4 + c (rpn/utils.py:11)
    const float4 d = make_float4(r[0], r[HW], r[2 * HW], r[3 * HW]);
    const float4 an = anc[flat];
    const float w = __fadd_rn(__fsub_rn(an.z, an.x), 1.0f), h = __fadd_rn(__fsub_rn(an.w, an.y), 1.0f);
    const float cx = __fadd_rn(an.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(an.y, __fmul_rn(0.5f, h));
    const float dx = __fdiv_rn(d.x, wx), dy = __fdiv_rn(d.y, wy);
    const float dw = fminf(__fdiv_rn(d.z, ww), kBboxXformClip), dh = fminf(__fdiv_rn(d.w, wh), kBboxXformClip);
    const float px = __fadd_rn(__fmul_rn(dx, w), cx), py = __fadd_rn(__fmul_rn(dy, h), cy);
    const float hpw = __fmul_rn(0.5f, __fmul_rn(expf(dw), w)), hph = __fmul_rn(0.5f, __fmul_rn(expf(dh), h));
    float x1 = __fsub_rn(px, hpw), y1 = __fsub_rn(py, hph);
    float x2 = __fsub_rn(__fadd_rn(px, hpw), 1.0f), y2 = __fsub_rn(__fadd_rn(py, hph), 1.0f);
    x1 = fminf(fmaxf(x1, 0.f), mx);
    y1 = fminf(fmaxf(y1, 0.f), my);
    x2 = fminf(fmaxf(x2, 0.f), mx);
    y2 = fminf(fmaxf(y2, 0.f), my);
    boxes[out0 + j] = make_float4(x1, y1, x2, y2);
    scores[out0 + j] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-logit)));  // torch.sigmoid
  }
}

bool g_rpn_force_exact = false;  // test hook: skip the sampled lower bound, always run the exact fallback

}  // namespace
}  // namespace b200

// test hook (not part of the reference-facing ABI)
extern "C" void b200_debug_rpn(int force_exact) { b200::g_rpn_force_exact = force_exact != 0; }

extern "C" int b200_rpn_candidates(const b200_rpn_level* levels, int n_levels, int n_images,
                                   const float* image_sizes, int pre_nms_top_n, float wx, float wy, float ww,
                                   float wh, float* boxes, float* scores, void* stream) {
  using namespace b200;
  B200_REQUIRE(levels && n_levels >= 1 && n_levels <= kRpnMaxLevels, "rpn_candidates: n_levels must be 1..%d",
               kRpnMaxLevels);
  B200_REQUIRE(n_images >= 0 && pre_nms_top_n > 0, "rpn_candidates: bad shape");
  if (n_images == 0) return B200_OK;
  B200_REQUIRE(image_sizes && boxes && scores, "rpn_candidates: null pointer");
  B200_REQUIRE(aligned16(boxes), "rpn_candidates: boxes must be 16-byte aligned");
  B200_REQUIRE(wx != 0.f && wy != 0.f && ww != 0.f && wh != 0.f, "rpn_candidates: zero box-coder weight");
  RpnLevels lv;
  int K = 0, kmax = 0;
  for (int l = 0; l < n_levels; ++l) {
    const b200_rpn_level& s = levels[l];
    B200_REQUIRE(s.objectness && s.box_regression && s.anchors && aligned16(s.anchors),
                 "rpn_candidates: level %d has a null or misaligned pointer", l);
    B200_REQUIRE(s.num_anchors > 0 && s.height > 0 && s.width > 0, "rpn_candidates: level %d has an empty shape", l);
    const long long M = (long long)s.num_anchors * s.height * s.width;
    B200_REQUIRE(M < (1ll << 30), "rpn_candidates: level %d too large", l);
    lv.obj[l] = s.objectness;
    lv.reg[l] = s.box_regression;
    lv.anchors[l] = s.anchors;
    lv.anchor_stride[l] = s.anchors_per_image ? 4 * M : 0;
    lv.A[l] = s.num_anchors;
    lv.HW[l] = s.height * s.width;
    lv.k[l] = (int)(M < pre_nms_top_n ? M : pre_nms_top_n);
    lv.start[l] = K;
    K += lv.k[l];
    if (lv.k[l] > kmax) kmax = lv.k[l];
  }
  lv.K = K;
  // candidate list capacity: a power of two >= 2 * kmax where shared memory allows (<= 16384
  // entries = 128 KB), never below 8192 (the sample sort borrows the upper half)
  int cap = 8192;
  while (cap < 2 * kmax && cap < 16384) cap <<= 1;
  if (cap < kmax) {
    set_error("rpn_candidates: pre_nms_top_n %d exceeds the %d candidates the kernel holds", pre_nms_top_n, cap);
    return B200_ERR_UNSUPPORTED;
  }
  const size_t smem = sizeof(u64) * (size_t)cap;
  static SmemHighWater hw;
  int rc = ensure_dynamic_smem(rpn_front_kernel, smem, &hw, "rpn_candidates: smem attribute");
  if (rc != B200_OK) return rc;
  rpn_front_kernel<<<dim3(n_levels, n_images), kRpnThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      lv, image_sizes, wx, wy, ww, wh, cap, g_rpn_force_exact ? 1 : 0, reinterpret_cast<float4*>(boxes), scores);
  B200_CHECK_LAUNCH("rpn_front_kernel");
  return B200_OK;
}
