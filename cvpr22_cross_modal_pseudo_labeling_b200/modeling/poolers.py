"""Pooler: drop-in for maskrcnn_benchmark.modeling.poolers (reference modeling/poolers.py).

`Pooler(output_size, scales, sampling_ratio).forward(x, boxes)` keeps the reference's
call signature, but the level assignment (LevelMapper, :31-42), the per-level
RoIAlign launches and the scatter back (:111-119) are ONE kernel launch
(b200_roi_align_forward), with autograd through b200_roi_align_backward.
"""
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


class Pooler(nn.Module):
    def __init__(self, output_size, scales, sampling_ratio, stage_nhwc=True, math=None):
        """
        output_size (tuple[int] or int), scales (list[float]), sampling_ratio (int):
        as the reference (poolers.py:55-76).
        stage_nhwc: run NCHW-contiguous maps through an NHWC copy cached per tensor (and an NHWC
        gradient buffer laid back out to NCHW in the backward) so the marching kernels are used;
        channels_last inputs are always used in place.
        math: "exact" (bit-identical to ROIAlign_forward_cpu) | "fast" (separable FMA evaluation,
        <= 1e-5 relative) | None = the process default (layers.roi_align.set_roi_align_math).
        """
        super(Pooler, self).__init__()
        self.poolers = nn.ModuleList(
            [ROIAlign(output_size, spatial_scale=s, sampling_ratio=sampling_ratio) for s in scales])
        self.output_size = output_size if isinstance(output_size, (tuple, list)) else (output_size, output_size)
        self.scales = tuple(float(s) for s in scales)
        self.sampling_ratio = sampling_ratio
        self.stage_nhwc = stage_nhwc
        self.math = math
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
        # half / bf16 maps: the op computes in fp32 (the reference's amp.float_function, layers/roi_align.py:57)
        # and the result comes back in the input dtype (poolers.py:104-109); layers.roi_align handles both,
        # keeping bf16 channels_last maps of the box-pooler shape bf16 all the way into the kernel
        return roi_align_multilevel(feats, rois.float(), self.output_size, self.scales, self.sampling_ratio,
                                    stage_nhwc=self.stage_nhwc, math=self.math)


def make_pooler(cfg, head_name):
    """Same factory as the reference (poolers.py:124-133); cfg is any attribute-style config."""
    head = getattr(cfg.MODEL, head_name) if not isinstance(cfg.MODEL, dict) else cfg.MODEL[head_name]
    res = head.POOLER_RESOLUTION
    return Pooler(output_size=(res, res), scales=head.POOLER_SCALES, sampling_ratio=head.POOLER_SAMPLING_RATIO)
