/*
 * b200det.h -- C ABI of the B200-native RoI hot path (libb200det.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.
 * Every entry point names the reference interface it replaces
 * (paths relative to /root/reference/maskrcnn_benchmark).  INTEGRATION.md shows
 * the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers unless a parameter says "host";
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     every kernel is launched on it and no call synchronises the host;
 *   - nothing is allocated: the caller owns outputs and workspaces
 *     (`*_workspace_bytes` tells how much);
 *   - return value: B200_OK (0) or a negative b200_status; a human-readable
 *     message for the calling thread's last failure is in b200_last_error_string();
 *   - there is no CPU implementation behind any symbol.
 */
#ifndef B200DET_H_
#define B200DET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DET_VERSION 100 /* major*1000 + minor*100 + patch */

typedef enum {
  B200_OK = 0,
  B200_ERR_INVALID_ARG = -1, /* bad shape / null pointer / misaligned pointer */
  B200_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed          */
  B200_ERR_WORKSPACE = -3,   /* workspace too small                          */
  B200_ERR_UNSUPPORTED = -4  /* size or option outside what the kernels cover */
} b200_status;

int b200_version(void);
const char* b200_last_error_string(void);

/* ------------------------------------------------------------------------
 * Feature layout of a level: logical shape is always [B, C, H, W] fp32.
 *   B200_LAYOUT_NCHW : contiguous as the reference requires
 *                      (csrc/cuda/ROIAlign_cuda.cu:286 `.contiguous()`).
 *   B200_LAYOUT_NHWC : torch.channels_last strides (memory [B, H, W, C]); the
 *                      fast path -- every tap is a 128-bit channel-vector load.
 * ------------------------------------------------------------------------ */
#define B200_LAYOUT_NCHW 0
#define B200_LAYOUT_NHWC 1
#define B200_MAX_LEVELS 8

typedef struct {
  const float* data;   /* device, 16-byte aligned                           */
  int32_t height;
  int32_t width;
  float spatial_scale; /* 1/stride of this level                            */
} b200_level;

typedef struct {
  float* data; /* gradient buffer of a level, same layout as the forward input */
  int32_t height;
  int32_t width;
  float spatial_scale;
} b200_level_grad;

/*
 * Fused multi-level RoIAlign forward.
 * Replaces, in one launch:  Pooler.forward  (modeling/poolers.py:91-121)
 *   = convert_to_roi_format + LevelMapper (poolers.py:31-42) + per-level
 *     _C.roi_align_forward (csrc/ROIAlign.h:11-24; kernels csrc/cpu/ROIAlign_cpu.cpp:221,
 *     csrc/cuda/ROIAlign_cuda.cu:257) + the result[idx] = ... scatter.
 * With n_levels == 1 it is exactly _C.roi_align_forward.
 *
 *   levels[n_levels]  host array; level l has scale spatial_scale[l]; for
 *                     n_levels > 1 the FPN level of each RoI is
 *                     floor(4 + log2(sqrt(area)/224 + 1e-6)) clamped to
 *                     [-log2(scale[0]), -log2(scale[n-1])], evaluated op by op in fp32.
 *   rois  [n_rois,5]  fp32 (batch_index, x1, y1, x2, y2) in image pixels
 *   out   [n_rois, channels, pooled_h, pooled_w] fp32, contiguous (always NCHW order)
 *   out_levels        optional int32 [n_rois]: the level each RoI used (may be NULL)
 */
int b200_roi_align_forward(const b200_level* levels, int n_levels, int layout,
                           int batch, int channels, const float* rois,
                           int64_t n_rois, int pooled_h, int pooled_w,
                           int sampling_ratio, float* out, int32_t* out_levels,
                           void* stream);

/*
 * Same operation and arguments as b200_roi_align_forward, evaluated separably ("fast math"):
 * every feature column a row of bins needs is first reduced over its tap rows, then the
 * x-samples combine two such column values, all with FMAs.  Algebraically identical to
 * csrc/cpu/ROIAlign_cpu.cpp:190-207; the fp32 result differs from the reference's summation
 * order by reassociation only (<= 1e-5 relative, the tolerance BASELINE.json states), while
 * b200_roi_align_forward is bit-identical to it.  ~12-35 % faster on the FPN poolers
 * (NHWC, sampling_ratio 2); the FPN box pooler shape (256 channels, 7x7) runs a row-streaming
 * kernel (cp.async.bulk row ring, ~40 % faster than the exact kernel); other shapes run the
 * generic gather with FMA contraction allowed.
 */
int b200_roi_align_forward_fast(const b200_level* levels, int n_levels, int layout,
                                int batch, int channels, const float* rois,
                                int64_t n_rois, int pooled_h, int pooled_w,
                                int sampling_ratio, float* out,
                                int32_t* out_levels, void* stream);

/*
 * Extended forward: arithmetic selected per call, plus an optional second output
 *   out_mean [n_rois, channels] fp32 (may be NULL): the mean over the pooled_h*pooled_w bins of
 *   every (RoI, channel) -- nn.AvgPool2d(kernel_size = pooled size) of the pooled block, the first
 *   step of FastRCNNPredictor.forward (modeling/roi_heads/box_head/roi_box_predictors.py:16-17, :62)
 *   -- taken from the output tile while it is still in shared memory, so the pooled block is not
 *   read back from HBM for it.  Sequential fp32 sum in bin order, one division (torch's order).
 */
#define B200_ROI_MATH_EXACT 0
#define B200_ROI_MATH_FAST 1
int b200_roi_align_forward_ex(const b200_level* levels, int n_levels, int layout,
                              int batch, int channels, const float* rois,
                              int64_t n_rois, int pooled_h, int pooled_w,
                              int sampling_ratio, int math, float* out,
                              float* out_mean, int32_t* out_levels, void* stream);

/*
 * b200_roi_align_forward_ex with caller-provided scratch (b200_roi_align_workspace_bytes(n_rois) bytes,
 * 4-byte aligned; NULL / too small = b200_roi_align_forward_ex).  The row-streaming kernel (fast math,
 * 256 NHWC channels, 7x7, sampling ratio 2) uses it for the order in which it visits the RoIs: sorted by
 * (image, FPN level, Morton code of the box centre), so the RoIs in flight on the 148 SMs at any time
 * are neighbours in one feature map and share its rows through the L2 (DRAM traffic ~ the touched
 * features once).  Results are written to every RoI's own slot: the output does not depend on it.
 * The reference allocates inside the op (csrc/cuda/ROIAlign_cuda.cu:271); here the caller owns all memory.
 */
size_t b200_roi_align_workspace_bytes(int64_t n_rois);
int b200_roi_align_forward_ws(const b200_level* levels, int n_levels, int layout,
                              int batch, int channels, const float* rois,
                              int64_t n_rois, int pooled_h, int pooled_w,
                              int sampling_ratio, int math, float* out,
                              float* out_mean, int32_t* out_levels, void* workspace,
                              size_t workspace_bytes, void* stream);

/*
 * The same forward on bf16 feature maps (levels[l].data points at bf16 values; BASELINE config #4: the
 * bf16 training step): the maps are read as bf16 -- half the bytes through the copy ring -- widened to fp32
 * in registers, pooled with the fast-math arithmetic, and written as fp32.  That is the reference's
 * amp.float_function (layers/roi_align.py:57: the op computes in fp32 whatever the input precision) on
 * bf16-rounded inputs.  Only the row-streaming shape has a bf16 kernel (NHWC, 256 channels, 7x7 bins,
 * sampling ratio 2); anything else returns B200_ERR_UNSUPPORTED and the caller casts to fp32.
 */
int b200_roi_align_forward_bf16(const b200_level* levels, int n_levels, int layout,
                                int batch, int channels, const float* rois,
                                int64_t n_rois, int pooled_h, int pooled_w,
                                int sampling_ratio, float* out, float* out_mean,
                                int32_t* out_levels, void* workspace,
                                size_t workspace_bytes, void* stream);

/*
 * Fused multi-level RoIAlign backward (gradient w.r.t. the features).
 * Replaces _C.roi_align_backward (csrc/ROIAlign.h:27-45; kernel
 * csrc/cuda/ROIAlign_cuda.cu:178-254, host :302-346) for every level at once.
 * The level gradient buffers must be zero-filled by the caller (the reference
 * allocates them with at::zeros, ROIAlign_cuda.cu:316); contributions are
 * ADDED, so one buffer can collect several calls.
 *   grad_out [n_rois, channels, pooled_h, pooled_w] fp32 contiguous
 */
int b200_roi_align_backward(const b200_level_grad* levels, int n_levels,
                            int layout, int batch, int channels,
                            const float* rois, int64_t n_rois, int pooled_h,
                            int pooled_w, int sampling_ratio,
                            const float* grad_out, void* stream);
/*
 * The same with caller-provided scratch (b200_roi_align_workspace_bytes(n_rois) bytes; NULL = the call above):
 * the marching kernel then visits the RoIs in the forward's (image, level, Morton cell) order, so that the
 * reductions of neighbouring RoIs meet lines that are still in the L2.  The gradient is a sum of atomic
 * additions in either order (as in the reference, csrc/cuda/ROIAlign_cuda.cu:239-250).
 */
int b200_roi_align_backward_ws(const b200_level_grad* levels, int n_levels,
                               int layout, int batch, int channels,
                               const float* rois, int64_t n_rois, int pooled_h,
                               int pooled_w, int sampling_ratio,
                               const float* grad_out, void* workspace,
                               size_t workspace_bytes, void* stream);

/*
 * NCHW -> NHWC staging copy of one level ([B,C,H,W] contiguous -> [B,H,W,C]).
 * No reference counterpart: it is what a caller uses when its backbone emits
 * NCHW-contiguous maps and it wants the NHWC fast path ("layouts").
 */
int b200_nchw_to_nhwc(const float* src, float* dst, int batch, int channels,
                      int height, int width, void* stream);
int b200_nhwc_to_nchw(const float* src, float* dst, int batch, int channels,
                      int height, int width, void* stream);

/*
 * Batched (segmented) greedy NMS.
 * Replaces _C.nms (csrc/nms.h:10-28 -> csrc/cpu/nms_cpu.cpp:67, csrc/cuda/nms.cu:70)
 * and the Python loops that call it once per (image, level) and per
 * (image, class): modeling/rpn/inference.py:111-122,
 * modeling/roi_heads/box_head/inference.py:135-149,
 * structures/boxlist_ops.py:9-31.
 * Semantics are the CPU reference's: legacy +1 widths, suppress when
 * IoU >= thresh, fp32 with separately rounded operations; visiting order is
 * descending score, equal scores visited in ascending index order.
 *
 *   boxes  [n_total,4] fp32 xyxy        scores [n_total] fp32
 *   seg_offsets [n_segments+1] int32 device, ascending, seg_offsets[0] == 0 and
 *               seg_offsets[n_segments] == n_total; segment s owns
 *               boxes[seg_offsets[s] : seg_offsets[s+1]]
 *   max_seg_len host upper bound on any segment's length (sizes grids and the
 *               workspace; <= B200_NMS_MAX_SEG)
 *   max_keep    > 0: only the first max_keep kept indices (ascending index
 *               order, as boxlist_nms's keep[:max_proposals]) are reported
 *   keep_idx [n_total] int64: for segment s the kept indices, LOCAL to the
 *               segment, ascending, at keep_idx[seg_offsets[s] ...]; the rest
 *               of the segment's slots are set to -1
 *   keep_cnt [n_segments] int32: number of kept entries per segment
 */
#define B200_NMS_MAX_SEG 16384
size_t b200_nms_workspace_bytes(int64_t n_total, int64_t n_segments,
                                int64_t max_seg_len);
int b200_nms_batched(const float* boxes, const float* scores,
                     const int32_t* seg_offsets, int64_t n_total,
                     int64_t n_segments, int64_t max_seg_len, float thresh,
                     int64_t max_keep, int64_t* keep_idx, int32_t* keep_cnt,
                     void* workspace, size_t workspace_bytes, void* stream);

/*
 * Cross-level proposal selection after the batched NMS: per image, the top_n highest-scoring
 * boxes among those its `segs_per_image` consecutive segments kept, written in RoI format.
 * Replaces RPNPostProcessor.select_over_all_levels in test mode
 * (modeling/rpn/inference.py:173-180) and the cat_boxlist / convert_to_roi_format glue that
 * feeds the pooler (modeling/rpn/inference.py:145-146, modeling/poolers.py:78-89).
 *   keep_idx / keep_cnt     outputs of b200_nms_batched over n_images * segs_per_image segments
 *   max_kept_per_image      host upper bound on sum of keep_cnt per image (<= 16384)
 *   rois_out  [n_images*top_n, 5] fp32 (image index, x1, y1, x2, y2), descending score per
 *             image (equal scores: ascending box index); rows past count_out[i] are zero boxes
 *   scores_out [n_images*top_n] fp32 (may be NULL), count_out [n_images] int32
 */
int b200_select_topk(const float* boxes, const float* scores,
                     const int32_t* seg_offsets, const int64_t* keep_idx,
                     const int32_t* keep_cnt, int n_images, int segs_per_image,
                     int64_t max_kept_per_image, int top_n, float* rois_out,
                     float* scores_out, int32_t* count_out, void* stream);

/*
 * Region -> class-embedding scoring: logits = A . E^T on the tcgen05 tensor
 * cores (bf16 operands, fp32 accumulation in TMEM) with the consumer fused
 * into the epilogue.  Replaces
 *   roi_box_predictors.py:67   einsum('pe,ce->pc', cls_emb, cls_score)
 *   box_head/inference.py:62   F.softmax(class_logits, -1)  (+ the `> score_thresh`
 *                              candidate test of :134)
 *   detector/st_generalized_rcnn.py:245-255  einsum('pd,wd->pw') + max over
 *                              regions + sigmoid (caption alignment)
 *
 *   A [n_rows, dim] bf16 row-major (projected RoI embeddings), 16-byte aligned,
 *     dim % 8 == 0;  E [n_cols, dim] bf16 row-major (class / word embeddings).
 *
 * B200_MATCH_SOFTMAX (n_cols <= 512):
 *   probs  [n_rows, n_cols] fp32  row softmax (may be NULL)
 *   logits [n_rows, n_cols] fp32  raw scores  (may be NULL)
 *   top_label [n_rows] int32 / top_prob [n_rows] fp32: arg-max over columns
 *     >= 1 (column 0 is the background row) and its probability; top_label is
 *     0 when that probability is <= score_thresh (may be NULL)
 * B200_MATCH_COLMAX:
 *   row_seg [n_rows] int32 : image index of each row (ascending)
 *   col_seg [n_cols] int32 : image index of each column (word)
 *   col_best [n_cols] uint64, caller-initialised to 0: receives per column the
 *     packed maximum over the rows of ITS image,
 *       (orderable(score) << 32) | (0xFFFFFFFF - local_row)
 *     so equal scores resolve to the first row, as torch.max does
 *     (decode with b200_colmax_decode).
 *   logits may be non-NULL to also get the raw [n_rows, n_cols] scores.
 */
#define B200_MATCH_SOFTMAX 0
#define B200_MATCH_COLMAX 1
int b200_embed_match(const void* A_bf16, const void* E_bf16, int64_t n_rows,
                     int n_cols, int dim, int mode, float score_thresh,
                     float* probs, float* logits, int32_t* top_label,
                     float* top_prob, const int32_t* row_seg,
                     const int32_t* col_seg, const int32_t* row_seg_start,
                     uint64_t* col_best, void* stream);
/*
 * y = x . W^T + b on the tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM, persistent
 * warp-specialised kernel with double-buffered accumulators, csrc/tc_gemm.cu): the `emb_pred` projection
 * nn.Linear(in, EMB_DIM) of FastRCNNPredictor that runs in front of the scoring product
 * (modeling/roi_heads/box_head/roi_box_predictors.py:63-66; called on its own by
 * detector/st_generalized_rcnn.py:226-228), and -- with transposed operands -- its two gradient GEMMs.
 *   A [n_rows, dim] bf16 row-major, W [n_out, dim] bf16 row-major (nn.Linear's weight layout), both
 *   16-byte aligned, dim % 8 == 0; bias [n_out] fp32 or NULL;
 *   out_bf16 [n_rows, n_out] bf16 and / or out_f32 [n_rows, n_out] fp32 (at least one).
 */
int b200_linear_bf16(const void* A_bf16, const void* W_bf16, const float* bias,
                     int64_t n_rows, int n_out, int dim, void* out_bf16,
                     float* out_f32, void* stream);

/*
 * SOFTMAX scoring against a class matrix of ANY width (e.g. the 1203-word LVIS vocabulary the
 * student scores caption images against, detector/st_generalized_rcnn.py:71-75,:191): one persistent
 * launch multiplies every 128-row tile twice -- a statistics pass (row max, sum, top-1 label), then a pass
 * that recomputes the column blocks and writes the probabilities -- so the [n_rows, n_cols] logits make no
 * HBM round trip.  probs / logits / top_label / top_prob are each optional, with the semantics of
 * b200_embed_match (logits is a plain output here, not scratch).
 */
int b200_embed_match_wide(const void* A_bf16, const void* E_bf16, int64_t n_rows,
                          int n_cols, int dim, float score_thresh, float* probs,
                          float* logits, int32_t* top_label, float* top_prob,
                          void* stream);
/* col_best -> (row index local to the image, max score, sigmoid(max score)) */
int b200_colmax_decode(const uint64_t* col_best, int n_cols, int32_t* row_idx,
                       float* max_score, float* sigmoid_score, void* stream);

/*
 * RPN candidates: the per-level front end of the RPN post-processor, all (image, level) pairs in
 * one launch.  Replaces RPNPostProcessor.forward_for_single_feature_map up to the
 * NMS (modeling/rpn/inference.py:76-110): permute_and_flatten (modeling/rpn/utils.py:10-14),
 * sigmoid, topk(min(pre_nms_top_n, A*H*W), sorted), gather of regression and anchors,
 * BoxCoder.decode (modeling/box_coder.py:52-95), clip_to_image(remove_empty=False)
 * (structures/bounding_box.py:214-219).
 *   levels[n_levels] host array; anchors in the reference's flattened order (h*W + w)*A + a
 *   image_sizes [n_images, 2] fp32 (width, height), device;  wx..wh: BoxCoder weights
 * Outputs, K = sum_l min(pre_nms_top_n, A_l*H_l*W_l) slots per image, image-major, levels in
 * order, descending objectness inside each (image, level) run (equal logits: ascending
 * flattened index) -- the segment layout b200_nms_batched / b200_select_topk take:
 *   boxes [n_images*K, 4] fp32 xyxy clipped, scores [n_images*K] fp32 = sigmoid(logit)
 */
typedef struct {
  const float* objectness;     /* [N, A, H, W] fp32 contiguous logits              */
  const float* box_regression; /* [N, 4A, H, W] fp32 contiguous                    */
  const float* anchors;        /* [A*H*W, 4] or [N, A*H*W, 4] xyxy, 16-byte aligned */
  int32_t num_anchors;         /* A                                                */
  int32_t height, width;
  int32_t anchors_per_image;   /* 0: one anchor set shared by all images           */
} b200_rpn_level;
int b200_rpn_candidates(const b200_rpn_level* levels, int n_levels, int n_images,
                        const float* image_sizes, int pre_nms_top_n, float wx,
                        float wy, float ww, float wh, float* boxes, float* scores,
                        void* stream);

/*
 * Box-head candidates: decode + clip + score threshold + per-class compaction, all images and
 * classes at once, feeding b200_nms_batched.  Replaces the front half of
 * PostProcessor.forward / filter_results (modeling/roi_heads/box_head/inference.py:69-76, :96,
 * :134-141): BoxCoder.decode (modeling/box_coder.py:52-95, legacy +1/-1, dw/dh clamped at
 * log(1000/16)), BoxList.clip_to_image (structures/bounding_box.py:214-224), `scores > thresh`,
 * and the per-class nonzero / gather.  The R x C x 4 decoded boxes are never materialised.
 *
 *   probs [n_rois, n_classes] fp32 (column 0 = background, skipped)
 *   box_regression [n_rois, reg_stride] fp32; class_agnostic != 0: the LAST four columns are the
 *       deltas of every class (inference.py:70), else columns 4j..4j+3 are class j's
 *   boxes [n_rois, 4] fp32 xyxy proposals, 16-byte aligned; RoIs of image i are rows
 *       roi_offsets[i] .. roi_offsets[i+1]-1 (int32 [n_images+1], device)
 *   image_sizes [n_images, 2] fp32 (width, height), device;  wx..wh: BoxCoder weights
 *   capacity: rows available in cand_*; with softmax scores at most
 *       n_rois * min(n_classes-1, ceil(1/score_thresh)-1) candidates exist
 * Outputs (segment s = image * (n_classes-1) + (class-1); inside a segment RoIs ascend -- the
 * order the reference enumerates):
 *   seg_len [n_segments] int32 scratch, seg_offsets [n_segments+1] int32
 *   cand_boxes [capacity,4], cand_scores [capacity], cand_roi [capacity] int32 (row of the RoI)
 *   status [2] int32: {total candidates, 1 if total > capacity (rows beyond capacity dropped)}
 */
int b200_box_candidates(const float* probs, const float* box_regression,
                        const float* boxes, const int32_t* roi_offsets,
                        const float* image_sizes, int n_images, int64_t n_rois,
                        int n_classes, int reg_stride, int class_agnostic,
                        float wx, float wy, float ww, float wh,
                        float score_thresh, int64_t capacity, int32_t* seg_len,
                        int32_t* seg_offsets, float* cand_boxes, float* cand_scores,
                        int32_t* cand_roi, int32_t* status, void* stream);

/*
 * Detections per image after the per-class NMS.  Replaces the back half of filter_results
 * (inference.py:143-163): concatenation of the classes' kept boxes (class ascending, NMS keep
 * order inside a class) and, when more than detections_per_img are kept, the kthvalue rule
 * `score >= (n - detections_per_img + 1)-th smallest score` (ties kept).
 *   cand_*, seg_offsets: outputs of b200_box_candidates;  keep_idx / keep_cnt: outputs of
 *   b200_nms_batched over the same segments;  classes_minus_1 <= 2048
 *   det_boxes [capacity,4], det_scores [capacity], det_labels [capacity] int64: image i's
 *   detections are rows seg_offsets[i*classes_minus_1] .. + det_count[i]
 */
int b200_select_detections(const float* cand_boxes, const float* cand_scores,
                           const int32_t* seg_offsets, const int64_t* keep_idx,
                           const int32_t* keep_cnt, int n_images, int classes_minus_1,
                           int detections_per_img, float* det_boxes, float* det_scores,
                           int64_t* det_labels, int32_t* det_count, void* stream);

/*
 * Mask targets of the student straight from the pseudo-labels' mask_size x mask_size masks, without the
 * full-image boolean masks in between: the composition of Masker's paste
 * (modeling/roi_heads/mask_head/inference.py:96-186, as b200_paste_masks) with project_masks_on_boxes
 * (modeling/roi_heads/mask_head/loss.py:11-42 = BinaryMaskList.crop at the proposal rounded half-to-even,
 * bilinear resize to target_size x target_size with align_corners = False, `.type_as(bool)`).
 *   masks [n_labels, mask_size, mask_size] fp32 probabilities; label_boxes [n_labels, 4] xyxy;
 *   match [n_proposals] int32: the label each proposal was matched to (< 0: all-zero target);
 *   proposals [n_proposals, 4] xyxy;  out [n_proposals, target_size, target_size] fp32 in {0, 1}.
 */
int b200_mask_targets(const float* masks, const float* label_boxes, const int32_t* match,
                      const float* proposals, int64_t n_proposals, int mask_size, int padding,
                      int im_h, int im_w, float thresh, int target_size, float* out,
                      void* stream);

/*
 * Mask paste: Masker.forward_single_image / paste_mask_in_image for all boxes of an
 * image in one launch (modeling/roi_heads/mask_head/inference.py:96-186): zero border of
 * `padding` pixels around each M x M mask probability map, the box grown by the same factor and
 * truncated to int32, bilinear resize to the box size (F.interpolate, align_corners=False),
 * `> thresh`, paste into a zero image.
 *   masks [n_boxes, M, M] fp32 probabilities;  boxes [n_boxes, 4] fp32 xyxy (image pixels)
 *   out   [n_boxes, im_h, im_w] uint8 (torch.bool), every byte written (0 outside the boxes)
 */
int b200_paste_masks(const float* masks, const float* boxes, int64_t n_boxes,
                     int mask_size, int padding, int im_h, int im_w, float thresh,
                     uint8_t* out, void* stream);

/*
 * RoIPool (max) forward / backward -- API compatibility with
 * _C.roi_pool_forward / _C.roi_pool_backward (csrc/ROIPool.h:11-48, kernels
 * csrc/cuda/ROIPool_cuda.cu:17-108).  NCHW only; no model in the reference uses it.
 */
int b200_roi_pool_forward(const float* input, int batch, int channels,
                          int height, int width, const float* rois,
                          int64_t n_rois, float spatial_scale, int pooled_h,
                          int pooled_w, float* out, int32_t* argmax,
                          void* stream);
int b200_roi_pool_backward(const float* grad_out, const int32_t* argmax,
                           const float* rois, int64_t n_rois, int batch,
                           int channels, int height, int width, int pooled_h,
                           int pooled_w, float* grad_in, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DET_H_ */
