"""oracle -- CPU restatement of the reference RoI hot path.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl
reference) may import this package; the product package never does.

Parity pin: the reference has no tests or golden vectors of its own
(SURVEY.md section 4).  The restatement is pinned instead against the reference's own
CPU sources compiled in place (oracle/_ref/libref_cpu.so, `ref_*` below) and
against fixtures generated from them (tests/golden/make_golden.py).

numpy in / numpy out; all arithmetic fp32 in the reference's operation order.
Citations are relative to /root/reference/maskrcnn_benchmark.
"""
import ctypes
import os

import numpy as np

from . import build_oracle as _b

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_ref = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        try:
            path = _b.build_oracle()
        except Exception:
            if not os.path.exists(path):
                raise
        _lib = ctypes.CDLL(path)
        _lib.oracle_nms.restype = ctypes.c_int64
    return _lib


def ref_lib():
    """The reference's own compiled CPU code, or None if unavailable."""
    global _ref
    if _ref is None:
        path = None
        try:
            path = _b.build_ref()
        except Exception:
            p = os.path.join(_HERE, "_ref", "libref_cpu.so")
            path = p if os.path.exists(p) else None
        if path is None:
            return None
        import torch  # noqa: F401  (libtorch must be resident before dlopen)
        _ref = ctypes.CDLL(path)
        _ref.ref_nms.restype = ctypes.c_long
    return _ref


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# --------------------------------------------------------------------------
# RoIAlign
# --------------------------------------------------------------------------
def roi_align_forward(inp, rois, scale, PH, PW, sr):
    """csrc/cpu/ROIAlign_cpu.cpp:221-257 (ROIAlign_forward_cpu)."""
    inp, rois = _f32(inp), _f32(rois).reshape(-1, 5)
    B, C, H, W = inp.shape
    R = rois.shape[0]
    out = np.empty((R, C, PH, PW), np.float32)
    if out.size:
        rc = lib().oracle_roi_align_forward(_p(inp, _f32p), B, C, H, W, _p(rois, _f32p), R,
                                            ctypes.c_float(scale), PH, PW, sr, _p(out, _f32p))
        assert rc == 0
    return out


def ref_roi_align_forward(inp, rois, scale, PH, PW, sr):
    """The compiled reference itself (oracle/_ref)."""
    r = ref_lib()
    assert r is not None, "oracle/_ref not built"
    inp, rois = _f32(inp), _f32(rois).reshape(-1, 5)
    B, C, H, W = inp.shape
    R = rois.shape[0]
    out = np.empty((R, C, PH, PW), np.float32)
    if out.size:
        rc = r.ref_roi_align_forward(_p(inp, _f32p), B, C, H, W, _p(rois, _f32p), R,
                                     ctypes.c_float(scale), PH, PW, sr, _p(out, _f32p))
        assert rc == 0
    return out


def roi_align_backward(grad, rois, scale, PH, PW, B, C, H, W, sr, fp64=True):
    """csrc/cuda/ROIAlign_cuda.cu:178-254.  Returns (grad32, grad64)."""
    grad, rois = _f32(grad), _f32(rois).reshape(-1, 5)
    R = rois.shape[0]
    g32 = np.zeros((B, C, H, W), np.float32)
    g64 = np.zeros((B, C, H, W), np.float64) if fp64 else None
    if R:
        rc = lib().oracle_roi_align_backward(_p(grad, _f32p), _p(rois, _f32p), R, ctypes.c_float(scale),
                                             PH, PW, B, C, H, W, sr, _p(g32, _f32p), _p(g64, _f64p))
        assert rc == 0
    return g32, g64


def level_map(rois, k_min, k_max, s0=224.0, lvl0=4.0, eps=1e-6):
    """modeling/poolers.py:31-42 (LevelMapper.__call__) on [R,5] rois."""
    rois = _f32(rois).reshape(-1, 5)
    lv = np.empty((rois.shape[0],), np.int32)
    lib().oracle_level_map(_p(rois, _f32p), ctypes.c_int64(rois.shape[0]), ctypes.c_float(k_min),
                           ctypes.c_float(k_max), ctypes.c_float(s0), ctypes.c_float(lvl0),
                           ctypes.c_float(eps), _p(lv, _i32p))
    return lv


def pooler_forward(feats, rois, scales, PH, PW, sr):
    """modeling/poolers.py:91-121 (Pooler.forward): feats = list of [B,C,H_l,W_l]."""
    rois = _f32(rois).reshape(-1, 5)
    if len(feats) == 1:
        return roi_align_forward(feats[0], rois, scales[0], PH, PW, sr), np.zeros(len(rois), np.int32)
    k_min = -np.log2(np.float32(scales[0]))
    k_max = -np.log2(np.float32(scales[-1]))
    lv = level_map(rois, float(k_min), float(k_max))
    C = feats[0].shape[1]
    out = np.zeros((rois.shape[0], C, PH, PW), np.float32)
    for l, (f, s) in enumerate(zip(feats, scales)):
        idx = np.nonzero(lv == l)[0]
        if len(idx):
            out[idx] = roi_align_forward(f, rois[idx], s, PH, PW, sr)
    return out, lv


# --------------------------------------------------------------------------
# RoIPool
# --------------------------------------------------------------------------
def roi_pool_forward(inp, rois, scale, PH, PW):
    """csrc/cuda/ROIPool_cuda.cu:17-77."""
    inp, rois = _f32(inp), _f32(rois).reshape(-1, 5)
    B, C, H, W = inp.shape
    R = rois.shape[0]
    out = np.empty((R, C, PH, PW), np.float32)
    arg = np.empty((R, C, PH, PW), np.int32)
    if out.size:
        lib().oracle_roi_pool_forward(_p(inp, _f32p), B, C, H, W, _p(rois, _f32p), R,
                                      ctypes.c_float(scale), PH, PW, _p(out, _f32p), _p(arg, _i32p))
    return out, arg


def roi_pool_backward(grad, argmax, rois, PH, PW, B, C, H, W):
    """csrc/cuda/ROIPool_cuda.cu:80-108."""
    grad, rois = _f32(grad), _f32(rois).reshape(-1, 5)
    argmax = np.ascontiguousarray(argmax, np.int32)
    g = np.zeros((B, C, H, W), np.float32)
    if rois.shape[0]:
        lib().oracle_roi_pool_backward(_p(grad, _f32p), _p(argmax, _i32p), _p(rois, _f32p), rois.shape[0],
                                       PH, PW, B, C, H, W, _p(g, _f32p))
    return g


# --------------------------------------------------------------------------
# NMS
# --------------------------------------------------------------------------
def nms(dets, scores, thresh, order=None):
    """csrc/cpu/nms_cpu.cpp:67-75.  Ties broken toward the lower index."""
    dets, scores = _f32(dets).reshape(-1, 4), _f32(scores).reshape(-1)
    N = dets.shape[0]
    keep = np.empty((N,), np.int64)
    o = np.ascontiguousarray(order, np.int64) if order is not None else None
    k = lib().oracle_nms(_p(dets, _f32p), _p(scores, _f32p), ctypes.c_int64(N), ctypes.c_float(thresh),
                         _p(o, _i64p), _p(keep, _i64p))
    assert k >= 0
    return keep[:k].copy()


def ref_nms(dets, scores, thresh):
    r = ref_lib()
    assert r is not None, "oracle/_ref not built"
    dets, scores = _f32(dets).reshape(-1, 4), _f32(scores).reshape(-1)
    N = dets.shape[0]
    keep = np.empty((max(N, 1),), np.int64)
    k = r.ref_nms(_p(dets, _f32p), _p(scores, _f32p), ctypes.c_long(N), ctypes.c_float(thresh),
                  keep.ctypes.data_as(ctypes.POINTER(ctypes.c_long)))
    assert k >= 0
    return keep[:k].copy()


def nms_batched(dets, scores, seg_off, thresh, max_keep=-1):
    """Per-segment nms; returns (keep_local [N] padded, keep_cnt [S])."""
    dets, scores = _f32(dets).reshape(-1, 4), _f32(scores).reshape(-1)
    seg_off = np.ascontiguousarray(seg_off, np.int64)
    S = len(seg_off) - 1
    keep = np.full((dets.shape[0],), -1, np.int64)
    cnt = np.zeros((S,), np.int64)
    rc = lib().oracle_nms_batched(_p(dets, _f32p), _p(scores, _f32p), _p(seg_off, _i64p), ctypes.c_int64(S),
                                  ctypes.c_float(thresh), ctypes.c_int64(max_keep), _p(keep, _i64p), _p(cnt, _i64p))
    assert rc == 0
    return keep, cnt


# --------------------------------------------------------------------------
# box decode / clip
# --------------------------------------------------------------------------
BBOX_XFORM_CLIP = float(np.log(1000.0 / 16))


def box_decode(codes, anchors, weights, img_size=None, clip=BBOX_XFORM_CLIP):
    """modeling/box_coder.py:52-95 then BoxList.clip_to_image (bounding_box.py:214-224).
    Note: torch.exp on CPU may differ from libm expf in the last ulp; callers
    compare decoded boxes with a 1e-4 px tolerance, not bit-exactly."""
    codes, anchors = _f32(codes).reshape(-1, 4), _f32(anchors).reshape(-1, 4)
    out = np.empty_like(codes)
    w, h = (img_size if img_size is not None else (-1.0, -1.0))
    lib().oracle_box_decode(_p(codes, _f32p), _p(anchors, _f32p), ctypes.c_int64(codes.shape[0]),
                            *[ctypes.c_float(x) for x in weights], ctypes.c_float(clip),
                            ctypes.c_float(w), ctypes.c_float(h), _p(out, _f32p))
    return out


def box_candidates(probs, box_regression, boxes, roi_offsets, image_sizes, weights, score_thresh,
                   class_agnostic):
    """Front half of PostProcessor.forward / filter_results for all images
    (modeling/roi_heads/box_head/inference.py:69-76, :96, :134-141): decode + clip +
    `scores > thresh`, candidates enumerated per (image, class 1..C-1) with RoIs ascending.
    Returns (seg_offsets int32 [B*(C-1)+1], cand_boxes [N,4], cand_scores [N], cand_roi int32 [N])."""
    probs, reg, boxes = _f32(probs), _f32(box_regression), _f32(boxes)
    n_img, C = len(roi_offsets) - 1, probs.shape[1]
    seg_off, cb, cs, cr = [0], [], [], []
    for i in range(n_img):
        r0, r1 = int(roi_offsets[i]), int(roi_offsets[i + 1])
        w, h = image_sizes[i]
        for j in range(1, C):
            inds = np.nonzero(probs[r0:r1, j] > np.float32(score_thresh))[0] + r0          # :137
            codes = reg[inds, -4:] if class_agnostic else reg[inds, 4 * j:4 * j + 4]          # :70, :139
            cb.append(box_decode(codes, boxes[inds], weights, (float(w), float(h))))
            cs.append(probs[inds, j])
            cr.append(inds.astype(np.int32))
            seg_off.append(seg_off[-1] + len(inds))
    cat = lambda xs, shp, dt: np.concatenate(xs) if xs else np.zeros(shp, dt)
    return (np.asarray(seg_off, np.int32), cat(cb, (0, 4), np.float32), cat(cs, (0,), np.float32),
            cat(cr, (0,), np.int32))


def select_detections(cand_boxes, cand_scores, seg_offsets, keep_idx, keep_cnt, n_images, detections_per_img):
    """Back half of filter_results (inference.py:143-163): concatenate the classes' NMS results
    (class ascending, keep order inside a class), then keep `score >= kthvalue(scores,
    n - detections_per_img + 1)` when more than detections_per_img survive (ties kept).
    Returns a list of (boxes, scores, labels int64) per image."""
    cfg = (len(seg_offsets) - 1) // max(n_images, 1)
    out = []
    for i in range(n_images):
        b, s, l = [], [], []
        for j in range(cfg):
            seg = i * cfg + j
            o, k = int(seg_offsets[seg]), int(keep_cnt[seg])
            idx = o + np.asarray(keep_idx[o:o + k], np.int64)
            b.append(cand_boxes[idx]); s.append(cand_scores[idx]); l.append(np.full(k, j + 1, np.int64))
        b = np.concatenate(b) if b else np.zeros((0, 4), np.float32)
        s = np.concatenate(s) if s else np.zeros((0,), np.float32)
        l = np.concatenate(l) if l else np.zeros((0,), np.int64)
        n = len(s)
        if n > detections_per_img > 0:
            thr = np.sort(s, kind="stable")[n - detections_per_img]      # kthvalue(k = n - d + 1), 1-based
            m = s >= thr
            b, s, l = b[m], s[m], l[m]
        out.append((b, s, l))
    return out


def paste_masks(masks, boxes, im_h, im_w, thresh=0.5, padding=1):
    """Masker.forward_single_image (modeling/roi_heads/mask_head/inference.py:96-186) in numpy
    float32: expand_masks / expand_boxes, int32 truncation, bilinear resize with the index and
    weight rules of F.interpolate(align_corners=False) (at::native area_pixel_compute_source_index
    + guard_index_and_lambda), `> thresh`, paste.  masks [N, M, M] -> bool [N, im_h, im_w]."""
    f = np.float32
    masks, boxes = _f32(masks), _f32(boxes)
    N, M = masks.shape[0], masks.shape[-1]
    Mp = M + 2 * padding
    out = np.zeros((N, im_h, im_w), np.bool_)

    def axis(scale, n_out, n_in):
        s = scale * (np.arange(n_out, dtype=f) + f(0.5)) - f(0.5)
        s = np.where(s < 0, f(0), s).astype(f)
        i0 = np.minimum(np.floor(s).astype(np.int64), n_in - 1)
        i1 = i0 + (i0 < n_in - 1)
        l1 = np.clip(s - i0.astype(f), f(0), f(1)).astype(f)
        return i0, i1, (f(1) - l1).astype(f), l1

    for n in range(N):
        pm = np.zeros((Mp, Mp), f)
        pm[padding:Mp - padding, padding:Mp - padding] = masks[n].reshape(M, M)
        scale = f(Mp) / f(M)
        x1, y1, x2, y2 = boxes[n]
        w_half, h_half = (x2 - x1) * f(0.5) * scale, (y2 - y1) * f(0.5) * scale
        x_c, y_c = (x2 + x1) * f(0.5), (y2 + y1) * f(0.5)
        b = [int(np.trunc(v)) for v in (x_c - w_half, y_c - h_half, x_c + w_half, y_c + h_half)]
        w, h = max(b[2] - b[0] + 1, 1), max(b[3] - b[1] + 1, 1)
        yi0, yi1, yl0, yl1 = axis(f(Mp) / f(h), h, Mp)
        xi0, xi1, xl0, xl1 = axis(f(Mp) / f(w), w, Mp)
        top = xl0[None] * pm[yi0][:, xi0] + xl1[None] * pm[yi0][:, xi1]
        bot = xl0[None] * pm[yi1][:, xi0] + xl1[None] * pm[yi1][:, xi1]
        m = (yl0[:, None] * top + yl1[:, None] * bot) > f(thresh)
        x_0, x_1 = max(b[0], 0), min(b[2] + 1, im_w)
        y_0, y_1 = max(b[1], 0), min(b[3] + 1, im_h)
        if x_1 > x_0 and y_1 > y_0:
            out[n, y_0:y_1, x_0:x_1] = m[y_0 - b[1]:y_1 - b[1], x_0 - b[0]:x_1 - b[0]]
    return out


def mask_targets(masks, label_boxes, match, proposals, im_h, im_w, M, thresh=0.5, padding=1):
    """project_masks_on_boxes (modeling/roi_heads/mask_head/loss.py:11-42) applied to the masks Masker pastes
    (paste_masks above): per proposal, BinaryMaskList.crop of its matched label's full-image mask
    (structures/segmentation_mask.py:118-137: Python round(), clamps), bilinear resize to M x M with
    align_corners = False (:139-158) and `.type_as(bool)` (any non-zero blend is True).  -> fp32 [P, M, M]."""
    f = np.float32
    full = paste_masks(masks, label_boxes, im_h, im_w, thresh, padding)
    proposals = _f32(proposals)
    out = np.zeros((len(proposals), M, M), f)

    def axis(n_in):
        s = (f(n_in) / f(M)) * (np.arange(M, dtype=f) + f(0.5)) - f(0.5)
        s = np.where(s < 0, f(0), s).astype(f)
        i0 = np.minimum(np.floor(s).astype(np.int64), n_in - 1)
        i1 = i0 + (i0 < n_in - 1)
        l1 = np.clip(s - i0.astype(f), f(0), f(1)).astype(f)
        return i0, i1, (f(1) - l1).astype(f), l1

    for p, (box, k) in enumerate(zip(proposals, match)):
        if k < 0:
            continue
        xmin, ymin, xmax, ymax = [round(float(b)) for b in box]
        xmin = min(max(xmin, 0), im_w - 1)
        ymin = min(max(ymin, 0), im_h - 1)
        xmax = max(min(max(xmax, 0), im_w), xmin + 1)
        ymax = max(min(max(ymax, 0), im_h), ymin + 1)
        crop = full[k, ymin:ymax, xmin:xmax].astype(f)
        yi0, yi1, yl0, yl1 = axis(crop.shape[0])
        xi0, xi1, xl0, xl1 = axis(crop.shape[1])
        top = xl0[None] * crop[yi0][:, xi0] + xl1[None] * crop[yi0][:, xi1]
        bot = xl0[None] * crop[yi1][:, xi0] + xl1[None] * crop[yi1][:, xi1]
        out[p] = ((yl0[:, None] * top + yl1[:, None] * bot) != 0).astype(f)
    return out


# --------------------------------------------------------------------------
# Region -> class-embedding scoring (plain torch ops in the reference; numpy here)
# --------------------------------------------------------------------------
def embed_logits(A, E, dtype=np.float32):
    """roi_box_predictors.py:67  einsum('pe,ce->pc', cls_emb, cls_score)."""
    return (np.asarray(A, dtype) @ np.asarray(E, dtype).T).astype(dtype)


def softmax_rows(logits):
    """box_head/inference.py:62  F.softmax(class_logits, -1)."""
    z = np.asarray(logits, np.float64)
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def embed_match_softmax(A, E, score_thresh):
    """logits -> softmax -> (probs, top-1 foreground label, its prob, thresholded mask).
    Class 0 is background (all-zero row, data/datasets/coco.py:85-89); the
    per-class candidate mask is `prob > score_thresh` for j >= 1
    (box_head/inference.py:134-136)."""
    logits = embed_logits(A, E, np.float64)
    probs = softmax_rows(logits)
    fg = probs[:, 1:]
    if fg.shape[1]:
        top = fg.argmax(axis=1) + 1
        top_p = probs[np.arange(len(probs)), top]
    else:
        top = np.zeros(len(probs), np.int64)
        top_p = np.zeros(len(probs), np.float32)
    cand = probs > np.float32(score_thresh)
    cand[:, 0] = False
    return probs, top.astype(np.int32), top_p.astype(np.float32), cand


def caption_align(emb_img, w_cap):
    """modeling/detector/st_generalized_rcnn.py:245-255 for ONE image:
    region_scores = emb_img @ w_cap.T ; per word max/argmax over regions
    (first index on ties, as torch.max) ; score = sigmoid(max)."""
    s = embed_logits(emb_img, w_cap, np.float64)            # [P, W]
    idx = s.argmax(axis=0)
    mx = s[idx, np.arange(s.shape[1])]
    return idx.astype(np.int32), mx.astype(np.float32), (1.0 / (1.0 + np.exp(-mx))).astype(np.float32)
