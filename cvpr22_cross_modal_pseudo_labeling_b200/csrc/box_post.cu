// box_post.cu -- the steps either side of the box head's batched NMS.
//
//   b200_box_candidates     replaces, for all images and classes at once, the front half of
//       PostProcessor.forward / filter_results (reference
//       modeling/roi_heads/box_head/inference.py:69-76, :96, :134-141):
//       BoxCoder.decode (modeling/box_coder.py:52-95) -> clip_to_image
//       (structures/bounding_box.py:214-224) -> `scores > score_thresh` -> per-class nonzero/gather.
//       The R x C x 4 decoded-box tensor (`proposals.repeat(1, C)`, :75) is never materialised: a
//       candidate's box is decoded when the candidate is written.  Output = the candidate list in
//       the reference's enumeration order (image, class 1..C-1, RoI ascending) as NMS segments.
//   b200_select_detections  replaces the back half (:143-163): concatenation of the per-class NMS
//       results and the kthvalue rule that keeps the `detections_per_img` best (ties kept).
//
// Nothing here synchronises the host: counts stay on the device (`status`, `det_count`).
#include "common.cuh"

namespace b200 {
namespace {

constexpr float kBboxXformClip = 4.135166556742356f;  // log(1000 / 16), box_coder.py:22

struct DecodeParams {
  float wx, wy, ww, wh;
};

// box_coder.py:64-93 + bounding_box.py:214-219, every operation rounded separately as the
// elementwise torch ops round them.
__device__ __forceinline__ float4 decode_clip(const float4 a, const float4 d, const DecodeParams& p, float img_w,
                                              float img_h) {
  const float w = __fadd_rn(__fsub_rn(a.z, a.x), 1.0f), h = __fadd_rn(__fsub_rn(a.w, a.y), 1.0f);
  const float cx = __fadd_rn(a.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(a.y, __fmul_rn(0.5f, h));
  const float dx = __fdiv_rn(d.x, p.wx), dy = __fdiv_rn(d.y, p.wy);
  const float dw = fminf(__fdiv_rn(d.z, p.ww), kBboxXformClip), dh = fminf(__fdiv_rn(d.w, p.wh), kBboxXformClip);
  const float px = __fadd_rn(__fmul_rn(dx, w), cx), py = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
  const float hpw = __fmul_rn(0.5f, pw), hph = __fmul_rn(0.5f, ph);
  float x1 = __fsub_rn(px, hpw), y1 = __fsub_rn(py, hph);
  float x2 = __fsub_rn(__fadd_rn(px, hpw), 1.0f), y2 = __fsub_rn(__fadd_rn(py, hph), 1.0f);
  const float mx = __fsub_rn(img_w, 1.0f), my = __fsub_rn(img_h, 1.0f);
  x1 = fminf(fmaxf(x1, 0.f), mx);
  y1 = fminf(fmaxf(y1, 0.f), my);
  x2 = fminf(fmaxf(x2, 0.f), mx);
  y2 = fminf(fmaxf(y2, 0.f), my);
  return make_float4(x1, y1, x2, y2);
}

constexpr int kCandWarps = 8;

// One warp per segment (image i, class j = 1..C-1): counts (kFill = false) or writes
// (kFill = true) the RoIs of the image whose probability of class j exceeds the threshold, in
// ascending RoI order.  The probability matrix is small (R x C fp32, L2-resident), so the
// column-strided reads cost sectors, not HBM.
template <bool kFill>
__global__ void __launch_bounds__(kCandWarps * 32)
cand_kernel(const float* __restrict__ probs, const float* __restrict__ reg, const float4* __restrict__ boxes,
            const int32_t* __restrict__ roi_off, const float* __restrict__ im_sizes, int n_images, int C,
            int reg_stride, int class_agnostic, DecodeParams dp, float thresh, long long capacity,
            int32_t* __restrict__ seg_len, const int32_t* __restrict__ seg_off, float4* __restrict__ cand_boxes,
            float* __restrict__ cand_scores, int32_t* __restrict__ cand_roi) {
  const int lane = threadIdx.x & 31;
  const long long seg = (long long)blockIdx.x * kCandWarps + (threadIdx.x >> 5);
  const int cfg = C - 1;
  if (seg >= (long long)n_images * cfg) return;
  const int img = (int)(seg / cfg), cls = (int)(seg % cfg) + 1;
  const int r0 = roi_off[img], r1 = roi_off[img + 1];
  const float img_w = kFill ? im_sizes[2 * img] : 0.f, img_h = kFill ? im_sizes[2 * img + 1] : 0.f;
  long long base = kFill ? seg_off[seg] : 0;
  int count = 0;
  for (int rb = r0; rb < r1; rb += 32) {
    const int r = rb + lane;
    const float p = r < r1 ? probs[(size_t)r * C + cls] : 0.f;
    const bool hit = r < r1 && p > thresh;
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    if (kFill) {
      if (hit) {
        const long long pos = base + __popc(m & ((1u << lane) - 1));
        if (pos < capacity) {
          const float* d = reg + (size_t)r * reg_stride + (class_agnostic ? reg_stride - 4 : 4 * cls);
          cand_boxes[pos] = decode_clip(boxes[r], make_float4(d[0], d[1], d[2], d[3]), dp, img_w, img_h);
          cand_scores[pos] = p;
          cand_roi[pos] = r;
        }
      }
      base += __popc(m);
    } else {
      count += __popc(m);
    }
  }
  if (!kFill && lane == 0) seg_len[seg] = count;
}

// exclusive scan of seg_len[n] in place into seg_off[n + 1] (seg_len aliases seg_off + 1 is NOT
// assumed: separate arrays); one CTA, chunked with a running carry.  status[0] = total,
// status[1] = 1 when the total exceeds the candidate capacity.
__global__ void __launch_bounds__(1024)
seg_scan_kernel(const int32_t* __restrict__ seg_len, long long n, long long capacity, int32_t* __restrict__ seg_off,
                int32_t* __restrict__ status) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (long long b = 0; b < n; b += 1024) {
    const long long i = b + tid;
    const int v = i < n ? seg_len[i] : 0;
    int s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += t;
    }
    if (lane == 31) warp_sum[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int ws = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, ws, d);
        if (lane >= d) ws += t;
      }
      warp_sum[lane] = ws;  // inclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    const int incl = carry + s + (warp > 0 ? warp_sum[warp - 1] : 0);
    // offsets are clamped to the capacity: on overflow (status[1], raised by the host after the chain) the
    // NMS / selection kernels that follow still see in-bounds -- truncated or empty -- segments
    if (i < n) seg_off[i] = (int)min((long long)(incl - v), capacity);
    __syncthreads();
    if (tid == 1023) carry_s = incl;
    __syncthreads();
  }
  if (tid == 0) {
    const int total = carry_s;
    seg_off[n] = (int)min((long long)total, capacity);
    status[0] = total;
    status[1] = (long long)total > capacity ? 1 : 0;
  }
}

__device__ __forceinline__ uint32_t orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int kSelThreads = 1024;
constexpr int kSelMaxClasses = 2048;  // foreground classes per image the selection kernel supports

// One CTA per image.  Phase 1 lists the image's kept candidates in the reference's order (class
// ascending; within a class the NMS keep order = ascending candidate index) into det_* at the
// image's base offset seg_off[img * cfg].  Phase 2 (only when more than `max_det` are kept):
// finds the max_det-th largest score by bisection on the order-preserving integer image of the
// scores -- torch.kthvalue(scores, n - max_det + 1), inference.py:156-158 -- and keeps
// `score >= that`, ties included, compacting in place and in order.
__global__ void __launch_bounds__(kSelThreads)
select_detections_kernel(const float4* __restrict__ cand_boxes, const float* __restrict__ cand_scores,
                         const int32_t* __restrict__ seg_off, const long long* __restrict__ keep_idx,
                         const int32_t* __restrict__ keep_cnt, int cfg, int max_det, float4* __restrict__ det_boxes,
                         float* __restrict__ det_scores, long long* __restrict__ det_labels,
                         int32_t* __restrict__ det_count) {
  __shared__ int warp_sum[32];
  __shared__ int carry_s;
  __shared__ unsigned int cnt_s;
  __shared__ int pre[kSelMaxClasses];
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long s0 = (long long)img * cfg;
  const long long out0 = seg_off[s0];
  // ---- phase 1: exclusive prefix of the per-segment kept counts, then warp per segment copies
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int sb = 0; sb < cfg; sb += kSelThreads) {
    const int sj = sb + tid;
    const int v = sj < cfg ? keep_cnt[s0 + sj] : 0;
    int s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += t;
    }
    if (lane == 31) warp_sum[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int ws = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, ws, d);
        if (lane >= d) ws += t;
      }
      warp_sum[lane] = ws;
    }
    __syncthreads();
    const int incl = carry_s + s + (warp > 0 ? warp_sum[warp - 1] : 0);
    if (sj < cfg) pre[sj] = incl - v;
    __syncthreads();
    if (tid == kSelThreads - 1) carry_s = incl;
    __syncthreads();
  }
  for (int sj = warp; sj < cfg; sj += kSelThreads / 32) {
    const int v = keep_cnt[s0 + sj];
    const long long src0 = seg_off[s0 + sj];
    const long long dst0 = out0 + pre[sj];
    for (int k = lane; k < v; k += 32) {
      const long long c = src0 + keep_idx[src0 + k];
      det_boxes[dst0 + k] = cand_boxes[c];
      det_scores[dst0 + k] = cand_scores[c];
      det_labels[dst0 + k] = sj + 1;
    }
  }
  __syncthreads();
  const int n = carry_s;
  if (!(max_det > 0 && n > max_det)) {
    if (tid == 0) det_count[img] = n;
    return;
  }
  // ---- phase 2: largest key t with count(key >= t) >= max_det  (= the max_det-th largest score)
  __threadfence_block();
  __syncthreads();
  uint32_t lo = 0u, hi = 0xffffffffu;  // invariant: count(>= lo) >= max_det
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1) + 1u;  // upper middle, > lo
    if (tid == 0) cnt_s = 0;
    __syncthreads();
    int c = 0;
    for (int i = tid; i < n; i += kSelThreads) c += orderable(det_scores[out0 + i]) >= mid;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane == 0 && c) atomicAdd(&cnt_s, (unsigned)c);
    __syncthreads();
    const unsigned total = cnt_s;
    __syncthreads();
    if (total >= (unsigned)max_det) lo = mid;
    else hi = mid - 1u;
  }
  const uint32_t thr = lo;
  // ordered in-place compaction, one chunk of kSelThreads entries at a time
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int b = 0; b < n; b += kSelThreads) {
    const int i = b + tid;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    float sc = 0.f;
    long long lb = 0;
    bool keep = false;
    if (i < n) {
      sc = det_scores[out0 + i];
      keep = orderable(sc) >= thr;
      if (keep) {
        bx = det_boxes[out0 + i];
        lb = det_labels[out0 + i];
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_sum[warp] = __popc(m);
    __syncthreads();  // also: every thread has read its source entry
    if (warp == 0) {
      int ws = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, ws, d);
        if (lane >= d) ws += t;
      }
      warp_sum[lane] = ws;
    }
    __syncthreads();
    const int base = carry_s + (warp > 0 ? warp_sum[warp - 1] : 0);
    if (keep) {
      const long long dst = out0 + base + __popc(m & ((1u << lane) - 1));
      det_boxes[dst] = bx;
      det_scores[dst] = sc;
      det_labels[dst] = lb;
    }
    __syncthreads();
    if (tid == 0) carry_s += warp_sum[31];
    __syncthreads();
  }
  if (tid == 0) det_count[img] = carry_s;
}

}  // namespace
}  // namespace b200

extern "C" int b200_box_candidates(const float* probs, const float* box_regression, const float* boxes,
                                   const int32_t* roi_offsets, const float* image_sizes, int n_images,
                                   int64_t n_rois, int n_classes, int reg_stride, int class_agnostic, float wx,
                                   float wy, float ww, float wh, float score_thresh, int64_t capacity,
                                   int32_t* seg_len, int32_t* seg_offsets, float* cand_boxes, float* cand_scores,
                                   int32_t* cand_roi, int32_t* status, void* stream) {
  using namespace b200;
  B200_REQUIRE(n_images >= 0 && n_rois >= 0 && n_classes >= 1 && capacity >= 0, "box_candidates: bad shape");
  B200_REQUIRE(reg_stride >= 4 && (class_agnostic || reg_stride >= 4 * n_classes),
               "box_candidates: box_regression has %d columns, needs %d", reg_stride,
               class_agnostic ? 4 : 4 * n_classes);
  B200_REQUIRE(wx != 0.f && wy != 0.f && ww != 0.f && wh != 0.f, "box_candidates: zero box-coder weight");
  B200_REQUIRE(seg_offsets && status, "box_candidates: null seg_offsets / status");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n_seg = (int64_t)n_images * (n_classes - 1);
  if (n_seg == 0 || n_rois == 0) {
    int rc = check_cuda(cudaMemsetAsync(seg_offsets, 0, sizeof(int32_t) * (size_t)(n_seg + 1), st),
                        "box_candidates: memset");
    if (rc != B200_OK) return rc;
    return check_cuda(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st), "box_candidates: memset");
  }
  B200_REQUIRE(probs && box_regression && boxes && roi_offsets && image_sizes && seg_len && cand_boxes &&
                   cand_scores && cand_roi,
               "box_candidates: null pointer");
  B200_REQUIRE(aligned16(boxes) && aligned16(cand_boxes), "box_candidates: boxes must be 16-byte aligned");
  B200_REQUIRE(n_rois * (int64_t)(n_classes - 1) < ((int64_t)1 << 31), "box_candidates: R x C exceeds int32");
  const DecodeParams dp{wx, wy, ww, wh};
  const unsigned grid = (unsigned)((n_seg + kCandWarps - 1) / kCandWarps);
  cand_kernel<false><<<grid, kCandWarps * 32, 0, st>>>(
      probs, box_regression, reinterpret_cast<const float4*>(boxes), roi_offsets, image_sizes, n_images, n_classes,
      reg_stride, class_agnostic, dp, score_thresh, (long long)capacity, seg_len, nullptr, nullptr, nullptr,
      nullptr);
  B200_CHECK_LAUNCH("cand_kernel<count>");
  seg_scan_kernel<<<1, 1024, 0, st>>>(seg_len, (long long)n_seg, (long long)capacity, seg_offsets, status);
  B200_CHECK_LAUNCH("seg_scan_kernel");
  cand_kernel<true><<<grid, kCandWarps * 32, 0, st>>>(
      probs, box_regression, reinterpret_cast<const float4*>(boxes), roi_offsets, image_sizes, n_images, n_classes,
      reg_stride, class_agnostic, dp, score_thresh, (long long)capacity, nullptr, seg_offsets,
      reinterpret_cast<float4*>(cand_boxes), cand_scores, cand_roi);
  B200_CHECK_LAUNCH("cand_kernel<fill>");
  return B200_OK;
}

extern "C" int b200_select_detections(const float* cand_boxes, const float* cand_scores, const int32_t* seg_offsets,
                                      const int64_t* keep_idx, const int32_t* keep_cnt, int n_images,
                                      int classes_minus_1, int detections_per_img, float* det_boxes,
                                      float* det_scores, int64_t* det_labels, int32_t* det_count, void* stream) {
  using namespace b200;
  B200_REQUIRE(n_images >= 0 && classes_minus_1 >= 0, "select_detections: bad shape");
  if (n_images == 0) return B200_OK;
  B200_REQUIRE(det_count, "select_detections: null det_count");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (classes_minus_1 == 0)
    return check_cuda(cudaMemsetAsync(det_count, 0, sizeof(int32_t) * (size_t)n_images, st),
                      "select_detections: memset");
  B200_REQUIRE(cand_boxes && cand_scores && seg_offsets && keep_idx && keep_cnt && det_boxes && det_scores &&
                   det_labels,
               "select_detections: null pointer");
  B200_REQUIRE(aligned16(cand_boxes) && aligned16(det_boxes), "select_detections: boxes must be 16-byte aligned");
  if (classes_minus_1 > kSelMaxClasses) {
    set_error("select_detections: %d foreground classes exceed %d", classes_minus_1, kSelMaxClasses);
    return B200_ERR_UNSUPPORTED;
  }
  select_detections_kernel<<<n_images, kSelThreads, 0, st>>>(
      reinterpret_cast<const float4*>(cand_boxes), cand_scores, seg_offsets,
      reinterpret_cast<const long long*>(keep_idx), keep_cnt, classes_minus_1, detections_per_img,
      reinterpret_cast<float4*>(det_boxes), det_scores, reinterpret_cast<long long*>(det_labels), det_count);
  B200_CHECK_LAUNCH("select_detections_kernel");
  return B200_OK;
}
