"""BoxList operations on the hot path (reference structures/boxlist_ops.py)."""
import torch

from ..layers import nms as _box_nms
from .bounding_box import BoxList


def boxlist_nms(boxlist, nms_thresh, max_proposals=-1, score_field="scores"):
    """Same contract as the reference boxlist_nms (structures/boxlist_ops.py:9-31):
    no-op for nms_thresh <= 0, keep[:max_proposals] of the ascending keep list."""
    if nms_thresh <= 0:
        return boxlist
    mode = boxlist.mode
    boxlist = boxlist.convert("xyxy")
    keep = _box_nms(boxlist.bbox, boxlist.get_field(score_field), nms_thresh)
    if max_proposals > 0:
        keep = keep[:max_proposals]
    return boxlist[keep].convert(mode)


def remove_small_boxes(boxlist, min_size):
    """Keep boxes whose legacy (+1) width and height are both >= min_size (:34-48)."""
    wh = boxlist.convert("xywh").bbox[:, 2:]
    keep = ((wh[:, 0] >= min_size) & (wh[:, 1] >= min_size)).nonzero().squeeze(1)
    return boxlist[keep]


def boxlist_iou(boxlist1, boxlist2):
    """[N,M] IoU with the legacy +1 convention (:53-91)."""
    if boxlist1.size != boxlist2.size:
        raise RuntimeError("boxlists should have same image size, got {}, {}".format(boxlist1, boxlist2))
    a, b = boxlist1.convert("xyxy"), boxlist2.convert("xyxy")
    lt = torch.max(a.bbox[:, None, :2], b.bbox[:, :2])
    rb = torch.min(a.bbox[:, None, 2:], b.bbox[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    return inter / (a.area()[:, None] + b.area() - inter)


def cat_boxlist(bboxes):
    """Concatenate BoxLists of one image (same size, mode and field set) (:103-129)."""
    assert isinstance(bboxes, (list, tuple)) and len(bboxes) > 0
    size, mode = bboxes[0].size, bboxes[0].mode
    fields = set(bboxes[0].fields())
    for b in bboxes:
        assert b.size == size and b.mode == mode and set(b.fields()) == fields
    if len(bboxes) == 1:
        return bboxes[0]
    out = BoxList(torch.cat([b.bbox for b in bboxes], dim=0), size, mode)
    for f in fields:
        out.add_field(f, torch.cat([b.get_field(f) for b in bboxes], dim=0))
    return out
