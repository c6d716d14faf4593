"""Mask targets of the student from the teacher's pseudo-label masks, without full-image masks in between
(the "paste o crop" fusion).

The reference pastes every pseudo-label's M0 x M0 mask into a full-image boolean mask on the way out of
generate_pseudo_label (Masker, modeling/detector/st_generalized_rcnn.py:267-272) and the mask loss then crops
that image at each positive proposal and resizes the crop to M x M on the CPU, one proposal at a time
(project_masks_on_boxes, modeling/roi_heads/mask_head/loss.py:11-42, with its own "FIXME: CPU computation
bottleneck").  `project_masks_on_boxes` here has the reference's name and meaning but takes the M0 x M0 mask
probabilities and the label boxes instead of a SegmentationMask: one launch of b200_mask_targets evaluates the
composition (16 taps per target pixel) for all proposals.  Results equal the reference's bit for bit
(tests/golden/mask_targets.npz is produced by the reference's own Masker + project_masks_on_boxes).
"""
import torch

from .... import _ext


def project_masks_on_boxes(mask_probs, label_boxes, matched_idxs, proposals, discretization_size, threshold=0.5,
                           padding=1):
    """
    mask_probs   [K, M0, M0] or [K, 1, M0, M0] fp32: mask probabilities of the K pseudo-labels of ONE image
                 (MaskPostProcessor output before the Masker, mask_head/inference.py:28-66)
    label_boxes  BoxList of the K pseudo-labels (the boxes the Masker would paste into)
    matched_idxs int [P]: the pseudo-label each proposal is matched to (Matcher output; < 0 = none -> zero target)
    proposals    BoxList of the P proposals (same image size)
    -> fp32 [P, M, M] in {0, 1}, on the proposals' device (reference loss.py:40-42)
    """
    assert label_boxes.size == proposals.size, "{}, {}".format(label_boxes, proposals)
    m = int(discretization_size)
    lb = label_boxes.convert("xyxy").bbox
    pb = proposals.convert("xyxy").bbox
    _ext.require_cuda(pb, "proposals")
    _ext.require_cuda(mask_probs, "mask_probs")
    dev = pb.device
    p = pb.shape[0]
    if p == 0:
        return torch.empty(0, dtype=torch.float32, device=dev)
    probs = mask_probs.reshape(mask_probs.shape[0], mask_probs.shape[-2], mask_probs.shape[-1]).float().contiguous()
    if probs.shape[0] != lb.shape[0] or probs.shape[1] != probs.shape[2]:
        raise ValueError("mask_probs must be [K, M0, M0] with one mask per label box")
    match = torch.as_tensor(matched_idxs, device=dev).to(torch.int32).contiguous()
    if match.numel() != p:
        raise ValueError("matched_idxs must have one entry per proposal")
    if probs.shape[0] == 0:
        return torch.zeros((p, m, m), dtype=torch.float32, device=dev)
    out = torch.empty((p, m, m), dtype=torch.float32, device=dev)
    im_w, im_h = proposals.size
    with torch.cuda.device(dev):
        rc = _ext.lib().b200_mask_targets(_ext.ptr(probs), _ext.ptr(lb.float().contiguous()), _ext.ptr(match),
                                          _ext.ptr(pb.float().contiguous()), p, probs.shape[1], int(padding), int(im_h),
                                          int(im_w), float(threshold), m, _ext.ptr(out), _ext.stream_ptr(dev))
    _ext.check(rc, "b200_mask_targets")
    return out
