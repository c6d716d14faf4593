"""NMS entry points.  `nms` mirrors maskrcnn_benchmark.layers.nms (reference
layers/nms.py:8 -> _C.nms, csrc/nms.h:10-28); `nms_batched` is the segmented form the
post-processors use instead of one call per (image, level) / (image, class)."""
import torch

from .. import _ext

_ws_cache = {}


def _workspace(nbytes, device):
    """Grow-only scratch buffer per device (torch caching allocator owns the memory)."""
    buf = _ws_cache.get(device)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[device] = buf
    return buf


def nms_batched(boxes, scores, seg_offsets, thresh, max_keep=-1, max_seg_len=None):
    """Greedy NMS over S independent segments in one call, no host sync.

    boxes [N,4] fp32 xyxy, scores [N] fp32, seg_offsets [S+1] int32 (device) with
    seg_offsets[0]=0, seg_offsets[S]=N.  `max_seg_len` is a host upper bound on
    the longest segment (defaults to N).  Returns (keep_idx int64 [N], keep_cnt
    int32 [S]): segment s's kept indices (local to the segment, ascending) sit at
    keep_idx[seg_offsets[s] : seg_offsets[s] + keep_cnt[s]], the rest is -1.
    """
    _ext.require_cuda(boxes, "boxes")
    _ext.require_cuda(scores, "scores")
    _ext.require_cuda(seg_offsets, "seg_offsets")
    if boxes.dim() != 2 or boxes.size(1) != 4:
        raise ValueError("boxes must be [N,4]")
    n = boxes.size(0)
    if scores.numel() != n:
        raise ValueError("scores must have one entry per box")
    s = seg_offsets.numel() - 1
    if s < 0:
        raise ValueError("seg_offsets must have at least one entry")
    boxes = boxes.float().contiguous()
    scores = scores.float().contiguous()
    seg_offsets = seg_offsets.to(torch.int32).contiguous()
    if max_seg_len is None:
        max_seg_len = n
    keep_idx = torch.empty((n,), dtype=torch.int64, device=boxes.device)
    keep_cnt = torch.empty((s,), dtype=torch.int32, device=boxes.device)
    if s == 0:
        return keep_idx, keep_cnt
    lib = _ext.lib()
    nbytes = lib.b200_nms_workspace_bytes(n, s, max_seg_len)
    ws = _workspace(nbytes, boxes.device)
    with torch.cuda.device(boxes.device):
        rc = lib.b200_nms_batched(_ext.ptr(boxes), _ext.ptr(scores), _ext.ptr(seg_offsets), n, s,
                                  int(max_seg_len), float(thresh), int(max_keep), _ext.ptr(keep_idx),
                                  _ext.ptr(keep_cnt), _ext.ptr(ws), ws.numel(), _ext.stream_ptr(boxes.device))
    _ext.check(rc, "b200_nms_batched")
    return keep_idx, keep_cnt


_MAX_SEG = 16384  # B200_NMS_MAX_SEG (include/b200det.h): longest segment one b200_nms_batched call takes


def _suppressed_by(rest, kept, threshold, chunk=1 << 22):
    """rest [M,4], kept [K,4] -> bool [M]: IoU(rest_i, kept_j) >= threshold for some j, every operation
    rounded separately in fp32 in the order of csrc/cpu/nms_cpu.cpp:45-57 (legacy +1 extents)."""
    out = torch.zeros((rest.size(0),), dtype=torch.bool, device=rest.device)
    if kept.size(0) == 0 or rest.size(0) == 0:
        return out
    ka = (kept[:, 2] - kept[:, 0] + 1) * (kept[:, 3] - kept[:, 1] + 1)
    step = max(1, chunk // kept.size(0))
    for o in range(0, rest.size(0), step):
        r = rest[o:o + step]
        ra = (r[:, 2] - r[:, 0] + 1) * (r[:, 3] - r[:, 1] + 1)
        w = (torch.minimum(kept[None, :, 2], r[:, None, 2]) - torch.maximum(kept[None, :, 0], r[:, None, 0]) + 1).clamp_(min=0)
        h = (torch.minimum(kept[None, :, 3], r[:, None, 3]) - torch.maximum(kept[None, :, 1], r[:, None, 1]) + 1).clamp_(min=0)
        inter = w * h
        out[o:o + step] = (inter / (ka[None, :] + ra[:, None] - inter) >= threshold).any(1)
    return out


def _nms_long(dets, scores, threshold):
    """Segments longer than B200_NMS_MAX_SEG (the reference handles any N): greedy NMS is block-recursive on
    the score-sorted list -- the kept boxes of the first block do not depend on later boxes, later boxes
    are first filtered by them, and the survivors are the next problem.  Each block runs through
    b200_nms_batched; the cross-block filter is elementwise torch in the reference's operation order."""
    order = torch.sort(scores.float(), descending=True, stable=True).indices   # ties: ascending index
    boxes = dets.float()[order].contiguous()
    kept = []
    while boxes.size(0) > 0:
        blk = boxes[:_MAX_SEG].contiguous()
        m = blk.size(0)
        rank = torch.arange(m, 0, -1, dtype=torch.float32, device=dets.device)  # already sorted: keep that order
        off = torch.tensor([0, m], dtype=torch.int32, device=dets.device)
        ki, kc = nms_batched(blk, rank, off, threshold, -1, m)
        k = ki[: int(kc.item())]
        kept.append(order[:m][k])
        rest, rest_order = boxes[m:], order[m:]
        if rest.size(0) == 0:
            break
        alive = ~_suppressed_by(rest, blk[k], threshold)
        boxes, order = rest[alive].contiguous(), rest_order[alive]
    return torch.sort(torch.cat(kept)).values


def nms(dets, scores, threshold):
    """_C.nms(dets[N,4], scores[N], threshold) -> int64 keep indices, ascending
    (reference csrc/cpu/nms_cpu.cpp:64).  Semantics of the CPU reference (`>=`,
    +1 extents).  One host sync to size the result, as the reference's CUDA path
    has (csrc/cuda/nms.cu:99-103)."""
    _ext.require_cuda(dets, "dets")
    _ext.require_cuda(scores, "scores")
    if dets.numel() == 0:
        # reference returns an empty int64 CPU tensor (csrc/nms.h:17-18); keep the device
        return torch.empty((0,), dtype=torch.int64, device=dets.device)
    n = dets.size(0)
    if n > _MAX_SEG:
        return _nms_long(dets, scores, threshold)
    off = torch.tensor([0, n], dtype=torch.int32, device=dets.device)
    keep_idx, keep_cnt = nms_batched(dets, scores, off, threshold, -1, n)
    return keep_idx[: int(keep_cnt.item())]


def select_topk(boxes, scores, seg_offsets, keep_idx, keep_cnt, n_images, top_n, max_kept_per_image):
    """Per image, the top_n best boxes among those kept by `nms_batched` in the image's
    consecutive segments (the RPN's cross-level selection, reference rpn/inference.py:173-180).
    Returns (rois [n_images*top_n, 5], scores [n_images*top_n], count int32 [n_images]); no host sync."""
    _ext.require_cuda(boxes, "boxes")
    s = keep_cnt.numel()
    if n_images <= 0 or s % n_images:
        raise ValueError("segments must divide evenly among images")
    dev = boxes.device
    rois = torch.empty((n_images * top_n, 5), dtype=torch.float32, device=dev)
    sc = torch.empty((n_images * top_n,), dtype=torch.float32, device=dev)
    cnt = torch.empty((n_images,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _ext.lib().b200_select_topk(_ext.ptr(boxes), _ext.ptr(scores), _ext.ptr(seg_offsets), _ext.ptr(keep_idx),
                                         _ext.ptr(keep_cnt), n_images, s // n_images, int(max_kept_per_image),
                                         int(top_n), _ext.ptr(rois), _ext.ptr(sc), _ext.ptr(cnt), _ext.stream_ptr(dev))
    _ext.check(rc, "b200_select_topk")
    return rois, sc, cnt
