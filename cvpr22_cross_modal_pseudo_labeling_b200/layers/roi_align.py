"""RoIAlign entry points mirroring maskrcnn_benchmark.layers.roi_align (reference
layers/roi_align.py:12-61) plus the fused multi-level form the Pooler uses."""
import ctypes

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from .. import _ext


def _layout_of(x):
    """(layout flag, tensor whose memory matches it).  channels_last tensors are used in
    place; anything else is made NCHW-contiguous like the reference does
    (csrc/cuda/ROIAlign_cuda.cu:286)."""
    if x.dim() != 4:
        raise ValueError("feature map must be [B,C,H,W]")
    if x.size(1) > 1 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
        return _ext.B200_LAYOUT_NHWC, x
    return _ext.B200_LAYOUT_NCHW, x.contiguous()


def _levels_array(tensors, scales):
    arr = (_ext.b200_level * len(tensors))()
    for i, (t, s) in enumerate(zip(tensors, scales)):
        arr[i].data = t.data_ptr()
        arr[i].height = t.size(2)
        arr[i].width = t.size(3)
        arr[i].spatial_scale = float(s)
    return arr


def _check_inputs(feats, rois):
    if len(feats) == 0 or len(feats) > _ext.B200_MAX_LEVELS:
        raise ValueError("need 1..%d feature levels" % _ext.B200_MAX_LEVELS)
    for f in feats:
        _ext.require_cuda(f, "input")
        if f.dtype != torch.float32:
            raise TypeError("input must be float32 (the reference forces fp32: layers/roi_align.py:57)")
    _ext.require_cuda(rois, "rois")
    if rois.dim() != 2 or rois.size(1) != 5:
        raise ValueError("rois must be [R,5] (batch_index, x1, y1, x2, y2)")
    b, c = feats[0].size(0), feats[0].size(1)
    for f in feats:
        if f.size(0) != b or f.size(1) != c:
            raise ValueError("all levels must share batch and channel sizes")


def _forward(feats, scales, rois, output_size, sampling_ratio, want_levels=False):
    _check_inputs(feats, rois)
    ph, pw = output_size
    rois = rois.float().contiguous()
    lay = [_layout_of(f) for f in feats]
    layouts = {l for l, _ in lay}
    if len(layouts) > 1:  # mixed: fall back to the reference's layout for all
        lay = [(_ext.B200_LAYOUT_NCHW, f.contiguous()) for f in feats]
    layout = lay[0][0]
    tensors = [t for _, t in lay]
    r, c = rois.size(0), feats[0].size(1)
    dev = feats[0].device
    out = torch.empty((r, c, ph, pw), dtype=torch.float32, device=dev)
    levels_out = torch.empty((r,), dtype=torch.int32, device=dev) if want_levels else None
    if r > 0:
        arr = _levels_array(tensors, scales)
        with torch.cuda.device(dev):
            rc = _ext.lib().b200_roi_align_forward(arr, len(tensors), layout, feats[0].size(0), c, _ext.ptr(rois),
                                                   r, ph, pw, int(sampling_ratio), _ext.ptr(out),
                                                   _ext.ptr(levels_out), _ext.stream_ptr(dev))
        _ext.check(rc, "b200_roi_align_forward")
    return out, levels_out


def _backward(grad_out, rois, shapes, layouts_nhwc, scales, output_size, sampling_ratio):
    """shapes: list of (B,C,H,W); returns one gradient tensor per level, in the memory
    format of the forward input."""
    ph, pw = output_size
    grad_out = grad_out.float().contiguous()
    rois = rois.float().contiguous()
    dev = grad_out.device
    fmt = torch.channels_last if layouts_nhwc else torch.contiguous_format
    grads = [torch.empty(s, dtype=torch.float32, device=dev, memory_format=fmt).zero_() for s in shapes]
    r = rois.size(0)
    if r > 0:
        arr = _levels_array(grads, scales)
        with torch.cuda.device(dev):
            rc = _ext.lib().b200_roi_align_backward(
                arr, len(grads), _ext.B200_LAYOUT_NHWC if layouts_nhwc else _ext.B200_LAYOUT_NCHW, shapes[0][0],
                shapes[0][1], _ext.ptr(rois), r, ph, pw, int(sampling_ratio), _ext.ptr(grad_out),
                _ext.stream_ptr(dev))
        _ext.check(rc, "b200_roi_align_backward")
    return grads


class _ROIAlignMulti(Function):
    """Fused level-assignment + RoIAlign over a feature pyramid (reference
    modeling/poolers.py:91-121 + layers/roi_align.py:12-45)."""

    @staticmethod
    def forward(ctx, rois, output_size, scales, sampling_ratio, *feats):
        output_size = _pair(output_size)
        out, _ = _forward(list(feats), scales, rois, output_size, sampling_ratio)
        ctx.save_for_backward(rois.float().contiguous())
        ctx.output_size = output_size
        ctx.scales = tuple(scales)
        ctx.sampling_ratio = sampling_ratio
        ctx.shapes = [tuple(f.shape) for f in feats]
        lay = {_layout_of(f)[0] for f in feats}
        ctx.nhwc = lay == {_ext.B200_LAYOUT_NHWC}
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        grads = _backward(grad_output, rois, ctx.shapes, ctx.nhwc, ctx.scales, ctx.output_size,
                          ctx.sampling_ratio)
        return (None, None, None, None, *grads)


def roi_align_multilevel(feats, rois, output_size, scales, sampling_ratio):
    """feats: list of [B,C,H_l,W_l]; rois [R,5] -> [R,C,PH,PW] in RoI order."""
    return _ROIAlignMulti.apply(rois, output_size, tuple(scales), sampling_ratio, *feats)


def roi_align(input, roi, output_size, spatial_scale, sampling_ratio):
    """Same call as the reference's `roi_align = _ROIAlign.apply` (layers/roi_align.py:48)."""
    return _ROIAlignMulti.apply(roi, output_size, (spatial_scale,), sampling_ratio, input)


class ROIAlign(nn.Module):
    """Drop-in for maskrcnn_benchmark.layers.ROIAlign (reference layers/roi_align.py:50-69)."""

    def __init__(self, output_size, spatial_scale, sampling_ratio):
        super(ROIAlign, self).__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    def forward(self, input, rois):
        # the reference wraps this in amp.float_function: fp32 in, fp32 out
        return roi_align(input.float(), rois.float(), self.output_size, self.spatial_scale, self.sampling_ratio)

    def __repr__(self):
        return "%s(output_size=%s, spatial_scale=%s, sampling_ratio=%s)" % (
            self.__class__.__name__, self.output_size, self.spatial_scale, self.sampling_ratio)
