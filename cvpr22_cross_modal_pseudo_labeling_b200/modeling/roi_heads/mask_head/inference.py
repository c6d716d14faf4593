"""MaskPostProcessor / Masker: drop-ins for maskrcnn_benchmark.modeling.roi_heads.mask_head.inference
(reference modeling/roi_heads/mask_head/inference.py:11-66, :124-205).

The reference pastes masks one box at a time on the CPU (paste_mask_in_image, :124-165: pad,
expand the box, bilinear resize, threshold, paste).  Here all boxes of an image go through ONE
launch of b200_paste_masks; the result stays on the device as the same [N, 1, H, W] bool tensor.
"""
import torch
from torch import nn

from .... import _ext
from ....structures import BoxList


class Masker(object):
    """Projects a set of masks in an image on the locations specified by the bounding boxes
    (reference :168-205)."""

    def __init__(self, threshold=0.5, padding=1):
        if threshold < 0:
            raise ValueError("the un-thresholded debug mode of the reference (threshold < 0) is not provided")
        self.threshold = threshold
        self.padding = padding

    def forward_single_image(self, masks, boxes):
        """masks [N, 1, M, M] probabilities, boxes BoxList -> [N, 1, im_h, im_w] bool."""
        boxes = boxes.convert("xyxy")
        im_w, im_h = boxes.size
        n = masks.shape[0]
        if n == 0:
            return masks.new_empty((0, 1, masks.shape[-2], masks.shape[-1]))
        _ext.require_cuda(masks, "masks")
        if masks.shape[-1] != masks.shape[-2]:
            raise ValueError("masks must be square")
        m = masks.reshape(n, masks.shape[-2], masks.shape[-1]).float().contiguous()
        bb = boxes.bbox.float().contiguous()
        out = torch.empty((n, 1, im_h, im_w), dtype=torch.bool, device=masks.device)
        with torch.cuda.device(masks.device):
            rc = _ext.lib().b200_paste_masks(_ext.ptr(m), _ext.ptr(bb), n, m.shape[-1], int(self.padding), int(im_h),
                                             int(im_w), float(self.threshold), _ext.ptr(out),
                                             _ext.stream_ptr(masks.device))
        _ext.check(rc, "b200_paste_masks")
        return out

    def __call__(self, masks, boxes):
        if isinstance(boxes, BoxList):
            boxes = [boxes]
        assert len(boxes) == len(masks), "Masks and boxes should have the same length."
        results = []
        for mask, box in zip(masks, boxes):
            assert mask.shape[0] == len(box), "Number of objects should be the same."
            results.append(self.forward_single_image(mask, box))
        return results


class MaskPostProcessor(nn.Module):
    """From the mask logits, take the mask of the predicted class (or channel 1 when class
    agnostic), as probabilities, into the field "mask"; with a masker, pasted into the image
    (reference :11-66)."""

    def __init__(self, masker=None, cls_agnostic_mask=False):
        super(MaskPostProcessor, self).__init__()
        self.masker = masker
        self.cls_agnostic_mask = cls_agnostic_mask

    def forward(self, x, boxes):
        mask_prob = x.sigmoid()
        if not self.cls_agnostic_mask:
            labels = torch.cat([bbox.get_field("labels") for bbox in boxes])
            index = torch.arange(x.shape[0], device=labels.device)
            mask_prob = mask_prob[index, labels][:, None]
        else:
            mask_prob = mask_prob[:, 1][:, None]
        mask_prob = mask_prob.split([len(box) for box in boxes], dim=0)
        if self.masker:
            mask_prob = self.masker(mask_prob, boxes)
        results = []
        for prob, box in zip(mask_prob, boxes):
            bbox = BoxList(box.bbox, box.size, mode="xyxy")
            for field in box.fields():
                bbox.add_field(field, box.get_field(field))
            bbox.add_field("mask", prob)
            results.append(bbox)
        return results


def make_roi_mask_post_processor(cfg):
    """Same factory as the reference (:208-216)."""
    head = cfg.MODEL.ROI_MASK_HEAD
    masker = Masker(threshold=head.POSTPROCESS_MASKS_THRESHOLD, padding=1) if head.POSTPROCESS_MASKS else None
    return MaskPostProcessor(masker, getattr(cfg.MODEL, "CLS_AGNOSTIC_MASK", False))
