"""RPNPostProcessor: drop-in for maskrcnn_benchmark.modeling.rpn.inference (reference
modeling/rpn/inference.py:13-206) with ONE batched NMS launch for all (image, level) pairs
instead of B x L calls of boxlist_nms (reference :111-122), and one host sync per forward.
"""
import math

import torch

from ... import _ext
from ...layers import nms_batched, select_topk
from ...structures import BoxList, cat_boxlist
from ..box_coder import BoxCoder


def permute_and_flatten(layer, N, A, C, H, W):
    """[N, A*C, H, W] -> [N, H*W*A, C] (reference modeling/rpn/utils.py:10-14)."""
    return layer.view(N, -1, C, H, W).permute(0, 3, 4, 1, 2).reshape(N, -1, C)


class RPNPostProcessor(torch.nn.Module):
    def __init__(self, pre_nms_top_n, post_nms_top_n, nms_thresh, min_size, box_coder=None,
                 fpn_post_nms_top_n=None, fpn_post_nms_per_batch=True):
        super(RPNPostProcessor, self).__init__()
        self.pre_nms_top_n = pre_nms_top_n
        self.post_nms_top_n = post_nms_top_n
        self.nms_thresh = nms_thresh
        self.min_size = min_size
        self.box_coder = box_coder if box_coder is not None else BoxCoder(weights=(1.0, 1.0, 1.0, 1.0))
        self.fpn_post_nms_top_n = post_nms_top_n if fpn_post_nms_top_n is None else fpn_post_nms_top_n
        self.fpn_post_nms_per_batch = fpn_post_nms_per_batch
        self.fused_select = True   # test hook: False sends the selection through the torch-op formulation
        self._seg_cache = {}

    # ---- reference :53-74 -------------------------------------------------------------
    def add_gt_proposals(self, proposals, targets):
        device = proposals[0].bbox.device
        out = []
        for proposal, target in zip(proposals, targets):
            gt = target.copy_with_fields([])
            gt.add_field("objectness", torch.ones(len(gt), device=device))
            out.append(cat_boxlist((proposal, gt)))
        return out

    def _segments(self, n_img, ks, device):
        """Cached per-shape index helpers: segment offsets (image-major: image n, level l),
        and for every slot of the [n_img, K] layout its segment id and segment start."""
        key = (n_img, tuple(ks), str(device))
        hit = self._seg_cache.get(key)
        if hit is None:
            K = sum(ks)
            lv_start = [0]
            for k in ks:
                lv_start.append(lv_start[-1] + k)
            offs = [n * K + s for n in range(n_img) for s in lv_start[:-1]] + [n_img * K]
            seg_off = torch.tensor(offs, dtype=torch.int32, device=device)
            seg_len = seg_off[1:] - seg_off[:-1]
            seg_id = torch.repeat_interleave(torch.arange(len(offs) - 1, device=device), seg_len.long(),
                                             output_size=n_img * K)
            slot_start = seg_off[:-1].long()[seg_id]
            hit = (seg_off, seg_id, slot_start)
            if len(self._seg_cache) > 32:
                self._seg_cache.clear()
            self._seg_cache[key] = hit
        return hit

    def _decode_level(self, anchors, objectness, box_regression):
        """Per level: sigmoid, top-k, decode, clip (reference :76-115).  Returns
        proposals [N, k, 4], scores [N, k] (descending per image)."""
        N, A, H, W = objectness.shape
        obj = permute_and_flatten(objectness, N, A, 1, H, W).view(N, -1).sigmoid()
        reg = permute_and_flatten(box_regression, N, A, 4, H, W)
        k = min(self.pre_nms_top_n, A * H * W)
        obj, idx = obj.topk(k, dim=1, sorted=True)
        bidx = torch.arange(N, device=obj.device)[:, None]
        reg = reg[bidx, idx]
        anc = torch.cat([a.bbox for a in anchors], dim=0).reshape(N, -1, 4)[bidx, idx]
        props = self.box_coder.decode(reg.view(-1, 4), anc.view(-1, 4)).view(N, k, 4)
        # BoxList.clip_to_image(remove_empty=False): clamp to [0, w-1] x [0, h-1] per image
        wh = torch.tensor([[a.size[0] - 1, a.size[1] - 1] for a in anchors], dtype=props.dtype, device=props.device)
        lim = torch.cat([wh, wh], dim=1)[:, None, :]
        props = torch.min(props.clamp(min=0), lim)
        return props, obj

    def _front_fused_ok(self, objectness, box_regression):
        clip = getattr(self.box_coder, "bbox_xform_clip", math.log(1000. / 16))
        return (all(o.is_cuda and o.dtype == torch.float32 for o in objectness) and
                all(b.is_cuda and b.dtype == torch.float32 for b in box_regression) and
                len(objectness) <= _ext.B200_MAX_LEVELS and abs(clip - math.log(1000. / 16)) < 1e-9 and
                min(self.pre_nms_top_n, max(o[0].numel() for o in objectness)) <= 16384)

    def _decode_all_fused(self, per_level, objectness, box_regression, sizes):
        """All levels, all images: sigmoid + top-k + gather + decode + clip in ONE launch
        (b200_rpn_candidates; reference :76-110).  Returns boxes [N*K, 4], scores [N*K] in the
        (image, level) segment layout, and the per-level k."""
        n_img = objectness[0].shape[0]
        device = objectness[0].device
        arr = (_ext.b200_rpn_level * len(objectness))()
        hold, ks = [], []
        for l, (a, o, b) in enumerate(zip(per_level, objectness, box_regression)):
            N, A, H, W = o.shape
            o, b = o.contiguous(), b.contiguous()
            shared = all(x.bbox.data_ptr() == a[0].bbox.data_ptr() for x in a)
            anc = a[0].bbox if shared else torch.cat([x.bbox for x in a], dim=0)
            anc = anc.to(torch.float32).contiguous()
            hold += [o, b, anc]
            arr[l].objectness, arr[l].box_regression, arr[l].anchors = o.data_ptr(), b.data_ptr(), anc.data_ptr()
            arr[l].num_anchors, arr[l].height, arr[l].width = A, H, W
            arr[l].anchors_per_image = 0 if shared else 1
            ks.append(min(self.pre_nms_top_n, A * H * W))
        K = sum(ks)
        im = torch.tensor([[float(s[0]), float(s[1])] for s in sizes], dtype=torch.float32).to(device)
        boxes = torch.empty((n_img * K, 4), dtype=torch.float32, device=device)
        score = torch.empty((n_img * K,), dtype=torch.float32, device=device)
        wx, wy, ww, wh = [float(v) for v in self.box_coder.weights]
        with torch.cuda.device(device):
            rc = _ext.lib().b200_rpn_candidates(arr, len(objectness), n_img, _ext.ptr(im), int(self.pre_nms_top_n),
                                                wx, wy, ww, wh, _ext.ptr(boxes), _ext.ptr(score),
                                                _ext.stream_ptr(device))
        _ext.check(rc, "b200_rpn_candidates")
        return boxes, score, ks

    def forward(self, anchors, objectness, box_regression, targets=None):
        """anchors: list[image] of list[level] BoxList; objectness / box_regression:
        list[level] of [N, A, H, W] / [N, 4A, H, W].  Returns list[image] BoxList with
        field "objectness" (reference :125-152)."""
        num_levels = len(objectness)
        n_img = objectness[0].shape[0]
        device = objectness[0].device
        per_level = list(zip(*anchors))
        sizes = [a[0].size for a in anchors]
        if self._front_fused_ok(objectness, box_regression):
            boxes, score, ks = self._decode_all_fused(per_level, objectness, box_regression, sizes)
            K = sum(ks)
        else:
            props, scores = [], []
            for a, o, b in zip(per_level, objectness, box_regression):
                p, s = self._decode_level(a, o, b)
                props.append(p)
                scores.append(s)
            ks = [p.shape[1] for p in props]
            K = sum(ks)
            boxes = torch.cat(props, dim=1).reshape(n_img * K, 4).contiguous()
            score = torch.cat(scores, dim=1).reshape(n_img * K).contiguous()
        seg_off, seg_id, slot_start = self._segments(n_img, ks, device)

        if self.min_size > 0:
            # remove_small_boxes (reference :115): boxes that fail are given the lowest score and
            # zero extent so they neither suppress nor survive; they are dropped after the NMS
            w = boxes[:, 2] - boxes[:, 0] + 1
            h = boxes[:, 3] - boxes[:, 1] + 1
            small = (w < self.min_size) | (h < self.min_size)
        else:
            small = None

        # the batched NMS + b200_select_topk serve every case but the per-batch top-k of FPN training (reference
        # :161-172) and min_size > 0: several levels in test mode (:173-180), and the single feature map of the
        # reference's shipped C4 configs in both modes, where the result is simply each image's kept boxes in
        # score order (:111-122 -- select_over_all_levels is skipped, :148-150)
        fused_select = (self.fused_select and self.nms_thresh > 0 and small is None and
                        not (num_levels > 1 and self.training and self.fpn_post_nms_per_batch))
        if fused_select:
            per_img = sum(min(k, self.post_nms_top_n) if self.post_nms_top_n > 0 else k for k in ks)
            fused_select = per_img <= 16384
        if fused_select:
            # NMS for all (image, level) segments, then the per-image top-k over levels
            # (reference :173-180) in one more kernel; one host sync for the counts
            keep_idx, keep_cnt = nms_batched(boxes, score, seg_off, self.nms_thresh, self.post_nms_top_n, max(ks))
            k2 = min(self.fpn_post_nms_top_n, K) if num_levels > 1 else per_img
            rois, sc, cnt = select_topk(boxes, score, seg_off, keep_idx, keep_cnt, n_img, k2, per_img)
            n_keep = cnt.tolist()
            results = []
            for i in range(n_img):
                bl = BoxList(rois[i * k2:i * k2 + n_keep[i], 1:], sizes[i], mode="xyxy")
                bl.add_field("objectness", sc[i * k2:i * k2 + n_keep[i]])
                results.append(bl)
            if self.training and targets is not None:
                results = self.add_gt_proposals(results, targets)
            return results

        if self.nms_thresh > 0:
            nms_scores = score
            nms_boxes = boxes
            if small is not None:
                # park small boxes far outside the image, each on its own spot
                park = -1e6 - 10.0 * torch.arange(boxes.shape[0], device=device, dtype=boxes.dtype)
                nms_boxes = torch.where(small[:, None], torch.stack([park, park, park, park], 1), boxes)
                nms_scores = torch.where(small, torch.full_like(score, -1.0), score)
            # small boxes sort last, so the first kept entries are exactly the reference's
            keep_idx, keep_cnt = nms_batched(nms_boxes, nms_scores, seg_off, self.nms_thresh,
                                             -1 if small is not None else self.post_nms_top_n, max(ks))
            pos = torch.arange(n_img * K, device=device)
            valid = (pos - slot_start) < keep_cnt.long()[seg_id]
            tgt = torch.where(valid, slot_start + keep_idx.clamp(min=0), torch.full_like(pos, n_img * K))
            kept = torch.zeros(n_img * K + 1, dtype=torch.bool, device=device)
            kept[tgt] = True
            kept = kept[:-1]
            if small is not None:
                kept &= ~small
                if self.post_nms_top_n > 0:  # keep[:post_nms_top_n] per segment (slots are in score order)
                    csum = torch.cumsum(kept.long(), dim=0)
                    base = torch.where(slot_start > 0, csum[(slot_start - 1).clamp(min=0)], torch.zeros_like(csum))
                    kept &= (csum - base) <= self.post_nms_top_n
        else:
            kept = torch.ones(n_img * K, dtype=torch.bool, device=device)
            if small is not None:
                kept &= ~small
        kept = kept.view(n_img, K)
        boxes = boxes.view(n_img, K, 4)
        score = score.view(n_img, K)

        results = []
        if num_levels > 1 and not (self.training and self.fpn_post_nms_per_batch):
            # per image top fpn_post_nms_top_n over all levels (reference :173-180)
            k2 = min(self.fpn_post_nms_top_n, K)
            masked = torch.where(kept, score, torch.full_like(score, -1.0))
            top_s, top_i = masked.topk(k2, dim=1, sorted=True)
            n_keep = torch.clamp(kept.sum(dim=1), max=k2).tolist()  # the one host sync
            for i in range(n_img):
                sel = top_i[i, : n_keep[i]]
                bl = BoxList(boxes[i, sel], sizes[i], mode="xyxy")
                bl.add_field("objectness", top_s[i, : n_keep[i]])
                results.append(bl)
        else:
            if num_levels > 1:
                # training: one top-k over the whole batch (reference :161-172)
                flat = torch.where(kept, score, torch.full_like(score, -1.0)).view(-1)
                total = int(kept.sum().item())
                k2 = min(self.fpn_post_nms_top_n, total)
                _, inds = flat.topk(k2, dim=0, sorted=True)
                sel = torch.zeros_like(flat, dtype=torch.bool)
                sel[inds] = True
                kept = sel.view(n_img, K)
            for i in range(n_img):
                m = kept[i]
                bl = BoxList(boxes[i][m], sizes[i], mode="xyxy")
                bl.add_field("objectness", score[i][m])
                results.append(bl)

        if self.training and targets is not None:
            results = self.add_gt_proposals(results, targets)
        return results


def make_rpn_postprocessor(config, rpn_box_coder, is_train):
    """Same factory as the reference (:184-206); config is any attribute-style config."""
    rpn = config.MODEL.RPN
    return RPNPostProcessor(
        pre_nms_top_n=rpn.PRE_NMS_TOP_N_TRAIN if is_train else rpn.PRE_NMS_TOP_N_TEST,
        post_nms_top_n=rpn.POST_NMS_TOP_N_TRAIN if is_train else rpn.POST_NMS_TOP_N_TEST,
        nms_thresh=rpn.NMS_THRESH,
        min_size=rpn.MIN_SIZE,
        box_coder=rpn_box_coder,
        fpn_post_nms_top_n=rpn.FPN_POST_NMS_TOP_N_TRAIN if is_train else rpn.FPN_POST_NMS_TOP_N_TEST,
        fpn_post_nms_per_batch=rpn.FPN_POST_NMS_PER_BATCH,
    )
