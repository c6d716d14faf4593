"""PostProcessor: drop-in for maskrcnn_benchmark.modeling.roi_heads.box_head.inference
(reference modeling/roi_heads/box_head/inference.py:12-193).

The per-class Python loop of filter_results (reference :135-149, one boxlist_nms call per
(image, class)) becomes ONE batched NMS over all (image, class) segments; candidates are
listed class-major / RoI-ascending exactly as the reference enumerates them, so results
come out in the same order.  Decode + clip + threshold + compaction before the NMS and the
top-`detections_per_img` selection after it are one library call each (b200_box_candidates,
b200_select_detections): the forward issues 5 kernels and synchronises the host once.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .... import _ext
from ....layers import nms_batched
from ....structures import BoxList
from ...box_coder import BoxCoder


class PostProcessor(nn.Module):
    def __init__(self, score_thresh=0.05, nms=0.5, detections_per_img=100, box_coder=None,
                 cls_agnostic_bbox_reg=False, bbox_aug_enabled=False, is_teacher=False, gt_box_eval=False):
        super(PostProcessor, self).__init__()
        self.score_thresh = score_thresh
        self.nms = nms
        self.detections_per_img = detections_per_img
        self.box_coder = box_coder if box_coder is not None else BoxCoder(weights=(10., 10., 5., 5.))
        self.cls_agnostic_bbox_reg = cls_agnostic_bbox_reg
        self.bbox_aug_enabled = bbox_aug_enabled
        self.is_teacher = is_teacher
        self.gt_box_eval = gt_box_eval

    def forward(self, x, boxes):
        """x = (class_logits [R,C], box_regression [R,4k]); boxes: list[image] BoxList.
        Returns list[image] BoxList with fields scores / labels (reference :49-100).
        If the predictor attached fused softmax output (`class_logits.b200_probs`), it is used
        instead of recomputing the softmax (reference :62)."""
        class_logits, box_regression = x
        class_prob = getattr(class_logits, "b200_probs", None)
        if class_prob is None:
            class_prob = F.softmax(class_logits, -1)
        image_shapes = [b.size for b in boxes]
        boxes_per_image = [len(b) for b in boxes]
        concat_boxes = torch.cat([b.bbox for b in boxes], dim=0)
        num_classes = class_prob.shape[1]
        teacher_path = self.bbox_aug_enabled or self.is_teacher
        if not teacher_path and not self.gt_box_eval and self._fused_ok(class_prob, box_regression, num_classes):
            return self._filter_fused(class_prob, box_regression, concat_boxes, boxes_per_image, image_shapes,
                                      num_classes)
        if self.cls_agnostic_bbox_reg:
            box_regression = box_regression[:, -4:]
        proposals = self.box_coder.decode(box_regression.view(sum(boxes_per_image), -1), concat_boxes)
        # clip_to_image per image (reference :96, bounding_box.py:214-224)
        lim = torch.cat([torch.tensor([[w - 1, h - 1, w - 1, h - 1]], dtype=proposals.dtype).expand(n, 4)
                         for (w, h), n in zip(image_shapes, boxes_per_image)], dim=0).to(proposals.device)
        k = proposals.shape[1] // 4
        proposals = torch.min(proposals.clamp(min=0), lim.repeat(1, k))

        if self.gt_box_eval and not self.is_teacher:
            labels = [b.get_field("labels").long() if b.has_field("labels") else None for b in boxes]
            new_prob = torch.zeros_like(class_prob)
            o = 0
            for lab, n in zip(labels, boxes_per_image):
                if lab is not None:
                    r = torch.arange(o, o + n, device=class_prob.device)
                    new_prob[r, lab] = class_prob[r, lab] + 1.1
                o += n
            class_prob = new_prob

        if self.bbox_aug_enabled or self.is_teacher:
            # unfiltered: every (RoI, class) pair, boxes repeated per class (reference :75,:97)
            if self.cls_agnostic_bbox_reg:
                proposals = proposals.repeat(1, num_classes)
            results = []
            for p, s, shape in zip(proposals.split(boxes_per_image, 0), class_prob.split(boxes_per_image, 0),
                                   image_shapes):
                bl = BoxList(p.reshape(-1, 4), shape, mode="xyxy")
                bl.add_field("scores", s.reshape(-1))
                results.append(bl)
            return results
        return self._filter_batched(proposals, class_prob, boxes_per_image, image_shapes, num_classes)

    def _fused_ok(self, class_prob, box_regression, num_classes):
        """The library path covers fp32 CUDA inputs, the stock dw/dh clamp and <= 2048 foreground classes."""
        clip = getattr(self.box_coder, "bbox_xform_clip", math.log(1000. / 16))
        return (class_prob.is_cuda and class_prob.dtype == torch.float32 and box_regression.dtype == torch.float32
                and 2 <= num_classes <= 2049 and abs(clip - math.log(1000. / 16)) < 1e-9
                and (self.cls_agnostic_bbox_reg or box_regression.shape[1] >= 4 * num_classes))

    def _filter_fused(self, class_prob, box_regression, concat_boxes, boxes_per_image, image_shapes, num_classes):
        """decode + clip + `score > thresh` + per-class compaction (b200_box_candidates) -> batched NMS ->
        per-image top detections_per_img with the kthvalue tie rule (b200_select_detections).
        Reference :69-76, :96, :121-163.  One host synchronisation, at the end, sizes the results."""
        device = class_prob.device
        n_img, r_total, cfg = len(boxes_per_image), sum(boxes_per_image), num_classes - 1
        if r_total == 0:
            return [self._empty(shape, device) for shape in image_shapes]
        probs = class_prob.contiguous()
        reg = box_regression.reshape(r_total, -1).contiguous()
        rois = concat_boxes.to(torch.float32).contiguous()
        offs = [0]
        for n in boxes_per_image:
            offs.append(offs[-1] + n)
        roi_off = torch.tensor(offs, dtype=torch.int32).to(device, non_blocking=True)
        sizes = torch.tensor([[float(w), float(h)] for (w, h) in image_shapes], dtype=torch.float32).to(
            device, non_blocking=True)
        # softmax rows sum to 1: at most ceil(1/thresh) - 1 classes of a RoI can exceed thresh
        per_roi = cfg if self.score_thresh <= 0 else min(cfg, max(1, int(math.ceil(1.0 / self.score_thresh)) - 1))
        cap = r_total * per_roi
        n_seg = n_img * cfg
        seg_len = torch.empty((n_seg,), dtype=torch.int32, device=device)
        seg_off = torch.empty((n_seg + 1,), dtype=torch.int32, device=device)
        cand_boxes = torch.empty((cap, 4), dtype=torch.float32, device=device)
        cand_scores = torch.empty((cap,), dtype=torch.float32, device=device)
        cand_roi = torch.empty((cap,), dtype=torch.int32, device=device)
        status = torch.empty((2,), dtype=torch.int32, device=device)
        wx, wy, ww, wh = [float(v) for v in self.box_coder.weights]
        lib = _ext.lib()
        with torch.cuda.device(device):
            st = _ext.stream_ptr(device)
            rc = lib.b200_box_candidates(_ext.ptr(probs), _ext.ptr(reg), _ext.ptr(rois), _ext.ptr(roi_off),
                                         _ext.ptr(sizes), n_img, r_total, num_classes, reg.shape[1],
                                         int(bool(self.cls_agnostic_bbox_reg)), wx, wy, ww, wh,
                                         float(self.score_thresh), cap, _ext.ptr(seg_len), _ext.ptr(seg_off),
                                         _ext.ptr(cand_boxes), _ext.ptr(cand_scores), _ext.ptr(cand_roi),
                                         _ext.ptr(status), st)
            _ext.check(rc, "b200_box_candidates")
            if self.nms > 0:
                keep_idx, keep_cnt = nms_batched(cand_boxes, cand_scores, seg_off, self.nms, -1, max(boxes_per_image))
            else:  # "nms_thresh <= 0: no-op" (boxlist_ops.py:20-21): every candidate is kept
                keep_cnt = seg_len
                pos = torch.arange(cap, device=device, dtype=torch.int64)
                seg_of = torch.searchsorted(seg_off[1:].long(), pos, right=True).clamp(max=n_seg - 1)
                keep_idx = pos - seg_off.long()[seg_of]
            det_boxes = torch.empty((cap, 4), dtype=torch.float32, device=device)
            det_scores = torch.empty((cap,), dtype=torch.float32, device=device)
            det_labels = torch.empty((cap,), dtype=torch.int64, device=device)
            det_count = torch.empty((n_img,), dtype=torch.int32, device=device)
            rc = lib.b200_select_detections(_ext.ptr(cand_boxes), _ext.ptr(cand_scores), _ext.ptr(seg_off),
                                            _ext.ptr(keep_idx), _ext.ptr(keep_cnt), n_img, cfg,
                                            int(self.detections_per_img), _ext.ptr(det_boxes), _ext.ptr(det_scores),
                                            _ext.ptr(det_labels), _ext.ptr(det_count), st)
            _ext.check(rc, "b200_select_detections")
        bases = seg_off[torch.arange(n_img, device=device) * cfg]
        host = torch.cat([det_count, bases, status]).tolist()        # the one host sync
        counts, base, (total, overflow) = host[:n_img], host[n_img:2 * n_img], host[2 * n_img:]
        if overflow:
            raise RuntimeError("PostProcessor: %d candidates exceed the capacity %d (scores are not a softmax?)"
                               % (total, cap))
        results = []
        for i in range(n_img):
            o, n = base[i], counts[i]
            bl = BoxList(det_boxes[o:o + n], image_shapes[i], mode="xyxy")
            s = det_scores[o:o + n]
            bl.add_field("scores", s)
            bl.add_field("objectness", s)
            bl.add_field("labels", det_labels[o:o + n])
            results.append(bl)
        return results

    def _filter_batched(self, proposals, class_prob, boxes_per_image, image_shapes, num_classes):
        """score > thresh, per-class NMS, top detections_per_img (reference :121-163), all
        images and classes at once."""
        device = class_prob.device
        n_img = len(boxes_per_image)
        cfg = num_classes - 1
        if cfg <= 0 or sum(boxes_per_image) == 0:
            return [self._empty(shape, device) for shape in image_shapes]
        rmax = max(boxes_per_image)
        # candidate mask laid out [image, class(1..C-1), roi] so that nonzero() enumerates
        # class-major, RoI-ascending -- the order the reference builds its result in
        mask = torch.zeros((n_img, cfg, rmax), dtype=torch.bool, device=device)
        o = 0
        for i, n in enumerate(boxes_per_image):
            if n:
                mask[i, :, :n] = (class_prob[o:o + n, 1:] > self.score_thresh).t()
            o += n
        img_start = torch.tensor([0] + boxes_per_image[:-1], device=device).cumsum(0)
        ii, jj, rr = mask.nonzero(as_tuple=True)                 # host sync (sizes the candidate list)
        roi = img_start[ii] + rr
        cls = jj + 1
        cand_scores = class_prob[roi, cls].contiguous()
        if proposals.shape[1] == 4:
            cand_boxes = proposals[roi].contiguous()
        else:
            cand_boxes = proposals.view(proposals.shape[0], -1, 4)[roi, cls].contiguous()
        seg_len = mask.sum(dim=2).view(-1)
        seg_off = torch.zeros(n_img * cfg + 1, dtype=torch.int32, device=device)
        seg_off[1:] = torch.cumsum(seg_len, 0).to(torch.int32)
        n_cand = cand_scores.shape[0]
        if n_cand == 0:
            return [self._empty(shape, device) for shape in image_shapes]
        if self.nms > 0:
            keep_idx, keep_cnt = nms_batched(cand_boxes, cand_scores, seg_off, self.nms, -1, rmax)
            seg_id = ii * cfg + jj
            start = seg_off[:-1].long()[seg_id]
            pos = torch.arange(n_cand, device=device)
            valid = (pos - start) < keep_cnt.long()[seg_id]
            tgt = torch.where(valid, start + keep_idx.clamp(min=0), torch.full_like(pos, n_cand))
            kept = torch.zeros(n_cand + 1, dtype=torch.bool, device=device)
            kept[tgt] = True
            kept = kept[:-1]
        else:
            kept = torch.ones(n_cand, dtype=torch.bool, device=device)
        # top detections_per_img per image with the reference's kthvalue rule (ties kept, :155-162)
        det_img = ii[kept]
        det_scores = cand_scores[kept]
        det_boxes = cand_boxes[kept]
        det_labels = cls[kept]
        counts = torch.bincount(det_img, minlength=n_img).tolist()   # host sync
        results = []
        o = 0
        for i in range(n_img):
            n = counts[i]
            s, b, l = det_scores[o:o + n], det_boxes[o:o + n], det_labels[o:o + n]
            o += n
            if n > self.detections_per_img > 0:
                thr, _ = torch.kthvalue(s, n - self.detections_per_img + 1)
                m = s >= thr
                s, b, l = s[m], b[m], l[m]
            bl = BoxList(b, image_shapes[i], mode="xyxy")
            bl.add_field("scores", s)
            bl.add_field("objectness", s)
            bl.add_field("labels", l.to(torch.int64))
            results.append(bl)
        return results

    @staticmethod
    def _empty(shape, device):
        bl = BoxList(torch.zeros((0, 4), device=device), shape, mode="xyxy")
        bl.add_field("scores", torch.zeros((0,), device=device))
        bl.add_field("objectness", torch.zeros((0,), device=device))
        bl.add_field("labels", torch.zeros((0,), dtype=torch.int64, device=device))
        return bl


def make_roi_box_post_processor(cfg, is_teacher=False):
    """Same factory as the reference (:166-193)."""
    score_thresh = cfg.MODEL.ROI_HEADS.SCORE_THRESH
    nms_thresh = cfg.MODEL.ROI_HEADS.NMS
    gt_box_eval = getattr(cfg.MODEL, "GT_BOX_EVAL", False)
    if gt_box_eval:
        score_thresh, nms_thresh = 1.0, 1.0
    return PostProcessor(score_thresh, nms_thresh, cfg.MODEL.ROI_HEADS.DETECTIONS_PER_IMG,
                         BoxCoder(weights=cfg.MODEL.ROI_HEADS.BBOX_REG_WEIGHTS), cfg.MODEL.CLS_AGNOSTIC_BBOX_REG,
                         cfg.TEST.BBOX_AUG.ENABLED, is_teacher, gt_box_eval)
