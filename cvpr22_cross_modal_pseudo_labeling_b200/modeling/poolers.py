"""Pooler: drop-in for maskrcnn_benchmark.modeling.poolers (reference modeling/poolers.py).

`Pooler(output_size, scales, sampling_ratio).forward(x, boxes)` keeps the reference's
call signature, but the level assignment (LevelMapper, :31-42), the per-level
RoIAlign launches and the scatter back (:111-119) are ONE kernel launch
(b200_roi_align_forward), with autograd through b200_roi_align_backward.
"""
import weakref

import torch
from torch import nn

from .. import _ext
from ..layers.roi_align import ROIAlign, roi_align_multilevel


class LevelMapper(object):
    """FPN eqn.(1) level assignment (reference poolers.py:11-42), kept for API
    compatibility; the fused kernel evaluates the same fp32 expression on the device."""

    def __init__(self, k_min, k_max, canonical_scale=224, canonical_level=4, eps=1e-6):
        self.k_min = k_min
        self.k_max = k_max
        self.s0 = canonical_scale
        self.lvl0 = canonical_level
        self.eps = eps

    def __call__(self, boxlists):
        s = torch.sqrt(torch.cat([b.area() for b in boxlists]))
        lv = torch.floor(self.lvl0 + torch.log2(s / self.s0 + self.eps))
        lv = torch.clamp(lv, min=self.k_min, max=self.k_max)
        return lv.to(torch.int64) - self.k_min


class _NhwcCache(object):
    """NCHW-contiguous feature maps are re-laid out to NHWC once per tensor and reused by
    every pooler that sees the same tensor (box 7x7, mask 14x14, forward and backward)."""

    def __init__(self, capacity=16):
        self.capacity = capacity
        self.entries = {}

    def get(self, x):
        key = id(x)
        e = self.entries.get(key)
        if e is not None:
            ref, version, y = e
            if ref() is x and x._version == version:
                return y
        b, c, h, w = x.shape
        y = torch.empty((b, c, h, w), dtype=x.dtype, device=x.device).contiguous(memory_format=torch.channels_last)
        with torch.cuda.device(x.device):
            rc = _ext.lib().b200_nchw_to_nhwc(_ext.ptr(x), _ext.ptr(y), b, c, h, w, _ext.stream_ptr(x.device))
        _ext.check(rc, "b200_nchw_to_nhwc")
        if len(self.entries) >= self.capacity:
            self.entries.clear()
        self.entries[key] = (weakref.ref(x), x._version, y)
        return y


_nhwc_cache = _NhwcCache()


class Pooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio, stage_nhwc=True):
        """
        output_size (tuple[int] or int), scales (list[float]), sampling_ratio (int):
        as the reference (poolers.py:55-76).
        stage_nhwc: re-lay NCHW-contiguous inference inputs out to NHWC (cached per tensor)
        so the staged kernel runs; channels_last inputs are always used in place.
        """
        super(Pooler, self).__init__()
        self.poolers = nn.ModuleList(
            [ROIAlign(output_size, spatial_scale=s, sampling_ratio=sampling_ratio) for s in scales])
        self.output_size = output_size if isinstance(output_size, (tuple, list)) else (output_size, output_size)
        self.scales = tuple(float(s) for s in scales)
        self.sampling_ratio = sampling_ratio
        self.stage_nhwc = stage_nhwc
        lvl_min = -torch.log2(torch.tensor(scales[0], dtype=torch.float32)).item()
        lvl_max = -torch.log2(torch.tensor(scales[-1], dtype=torch.float32)).item()
        self.map_levels = LevelMapper(lvl_min, lvl_max)

    def convert_to_roi_format(self, boxes):
        """list[BoxList] -> [R,5] (batch_index, x1, y1, x2, y2) (reference poolers.py:78-89)."""
        bb = torch.cat([b.bbox for b in boxes], dim=0)
        counts = [len(b) for b in boxes]
        ids = torch.repeat_interleave(torch.arange(len(boxes), dtype=bb.dtype, device=bb.device),
                                      torch.tensor(counts, device=bb.device), output_size=sum(counts))
        return torch.cat([ids[:, None], bb], dim=1)

    def forward(self, x, boxes):
        """x: list[Tensor [B,C,H_l,W_l]], boxes: list[BoxList] -> [R,C,PH,PW] in RoI order."""
        rois = self.convert_to_roi_format(boxes)
        feats = list(x)[: len(self.scales)]
        if self.stage_nhwc and self.sampling_ratio == 2 and feats[0].size(1) % 64 == 0:
            staged = []
            for f in feats:
                needs_grad = f.requires_grad and torch.is_grad_enabled()
                if f.is_cuda and f.is_contiguous() and not needs_grad and f.size(1) > 1 and \
                        not f.is_contiguous(memory_format=torch.channels_last):
                    f = _nhwc_cache.get(f)
                staged.append(f)
            feats = staged
        return roi_align_multilevel(feats, rois, self.output_size, self.scales, self.sampling_ratio)


def make_pooler(cfg, head_name):
    """Same factory as the reference (poolers.py:124-133); cfg is any attribute-style config."""
    head = getattr(cfg.MODEL, head_name) if not isinstance(cfg.MODEL, dict) else cfg.MODEL[head_name]
    res = head.POOLER_RESOLUTION
    return Pooler(output_size=(res, res), scales=head.POOLER_SCALES, sampling_ratio=head.POOLER_SAMPLING_RATIO)
