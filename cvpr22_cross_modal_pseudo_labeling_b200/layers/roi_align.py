"""RoIAlign entry points mirroring maskrcnn_benchmark.layers.roi_align (reference
layers/roi_align.py:12-61) plus the fused multi-level form the Pooler uses."""
import ctypes
import os
import weakref

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from .. import _ext


MATH_MODES = ("exact", "fast")
_default_math = os.environ.get("B200_ROI_ALIGN_MATH", "exact")
if _default_math not in MATH_MODES:
    raise ValueError("B200_ROI_ALIGN_MATH must be one of %s" % (MATH_MODES,))


def set_roi_align_math(mode):
    """Process-wide default arithmetic of the RoIAlign forward: "exact" reproduces the reference's
    operation order bit for bit (b200_roi_align_forward); "fast" evaluates the same bilinear sums
    separably with FMAs (b200_roi_align_forward_fast; <= 1e-5 relative).  Returns the old mode."""
    global _default_math
    if mode not in MATH_MODES:
        raise ValueError("math must be one of %s" % (MATH_MODES,))
    old, _default_math = _default_math, mode
    return old


def get_roi_align_math():
    return _default_math


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
        if f.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            raise TypeError("input must be float32, bfloat16 or float16")
        if f.dtype != feats[0].dtype:
            raise TypeError("all levels must share one dtype")
    _ext.require_cuda(rois, "rois")
    if rois.dim() != 2 or rois.size(1) != 5:
        raise ValueError("rois must be [R,5] (batch_index, x1, y1, x2, y2)")
    b, c = feats[0].size(0), feats[0].size(1)
    for f in feats:
        if f.size(0) != b or f.size(1) != c:
            raise ValueError("all levels must share batch and channel sizes")


_ws_cache = {}


def _workspace(nbytes, device):
    """Grow-only scratch buffer per (device, stream); the torch caching allocator owns the memory."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _forward(feats, scales, rois, output_size, sampling_ratio, want_levels=False, math=None, mean_out=None):
    """-> (pooled [R,C,PH,PW], levels or None).  mean_out: optional preallocated [R,C] fp32 tensor that
    receives the per-channel mean over the bins (b200_roi_align_forward_ex, fused AvgPool2d)."""
    _check_inputs(feats, rois)
    math = _default_math if math is None else math
    if math not in MATH_MODES:
        raise ValueError("math must be one of %s" % (MATH_MODES,))
    ph, pw = output_size
    rois = rois.float().contiguous()
    if feats[0].dtype != torch.float32:
        # The reference computes RoIAlign in fp32 whatever the input precision (amp.float_function,
        # layers/roi_align.py:57).  bf16 channels_last maps of the FPN box-pooler shape stay bf16 all the way
        # into the kernel (b200_roi_align_forward_bf16: half the bytes, fp32 arithmetic and output); every other
        # reduced-precision input is widened first, as the reference does.
        out = _forward_bf16(feats, scales, rois, (ph, pw), sampling_ratio, want_levels, mean_out) \
            if math == "fast" else None
        if out is not None:
            return out
        feats = [f.float() for f in feats]
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
        if mean_out is not None and (mean_out.shape != (r, c) or mean_out.dtype != torch.float32 or
                                     not mean_out.is_contiguous() or mean_out.device != dev):
            raise ValueError("mean_out must be a contiguous float32 [R,C] tensor on the input's device")
        lib = _ext.lib()
        # scratch for the RoI visiting order of the row-streaming kernel (grow-only, per device and stream)
        ws = _workspace(lib.b200_roi_align_workspace_bytes(r), dev) if math == "fast" else None
        with torch.cuda.device(dev):
            rc = lib.b200_roi_align_forward_ws(
                arr, len(tensors), layout, feats[0].size(0), c, _ext.ptr(rois), r, ph, pw, int(sampling_ratio),
                MATH_MODES.index(math), _ext.ptr(out), _ext.ptr(mean_out), _ext.ptr(levels_out), _ext.ptr(ws),
                ws.numel() if ws is not None else 0, _ext.stream_ptr(dev))
        _ext.check(rc, "b200_roi_align_forward")
    return out, levels_out


def _forward_bf16(feats, scales, rois, output_size, sampling_ratio, want_levels, mean_out):
    """bf16 maps straight into the row-streaming kernel, or None when the shape has no bf16 kernel."""
    ph, pw = output_size
    if feats[0].dtype != torch.bfloat16 or (ph, pw) != (7, 7) or sampling_ratio != 2 or feats[0].size(1) != 256:
        return None
    if any(_layout_of(f)[0] != _ext.B200_LAYOUT_NHWC for f in feats):
        return None
    r, c = rois.size(0), 256
    dev = feats[0].device
    out = torch.empty((r, c, ph, pw), dtype=torch.float32, device=dev)
    levels_out = torch.empty((r,), dtype=torch.int32, device=dev) if want_levels else None
    if r > 0:
        if mean_out is not None and (mean_out.shape != (r, c) or mean_out.dtype != torch.float32 or
                                     not mean_out.is_contiguous() or mean_out.device != dev):
            raise ValueError("mean_out must be a contiguous float32 [R,C] tensor on the input's device")
        lib = _ext.lib()
        arr = _levels_array(feats, scales)
        ws = _workspace(lib.b200_roi_align_workspace_bytes(r), dev)
        with torch.cuda.device(dev):
            rc = lib.b200_roi_align_forward_bf16(arr, len(feats), _ext.B200_LAYOUT_NHWC, feats[0].size(0), c, _ext.ptr(rois), r,
                                                 ph, pw, 2, _ext.ptr(out), _ext.ptr(mean_out), _ext.ptr(levels_out),
                                                 _ext.ptr(ws), ws.numel(), _ext.stream_ptr(dev))
        _ext.check(rc, "b200_roi_align_forward_bf16")
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
        lib = _ext.lib()
        ws = _workspace(lib.b200_roi_align_workspace_bytes(r), dev)   # the RoI visiting order of the marching kernel
        with torch.cuda.device(dev):
            rc = lib.b200_roi_align_backward_ws(
                arr, len(grads), _ext.B200_LAYOUT_NHWC if layouts_nhwc else _ext.B200_LAYOUT_NCHW, shapes[0][0],
                shapes[0][1], _ext.ptr(rois), r, ph, pw, int(sampling_ratio), _ext.ptr(grad_out), _ext.ptr(ws), ws.numel(),
                _ext.stream_ptr(dev))
        _ext.check(rc, "b200_roi_align_backward")
    return grads


class _NhwcCache(object):
    """NCHW-contiguous feature maps are re-laid out to NHWC once per tensor (b200_nchw_to_nhwc)
    and reused by every pooler call that sees the same tensor object at the same version: the
    box pooler, the mask pooler, and the backward pass."""

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
        # drop copies whose source tensor is gone, so the cache never pins dead feature maps
        for k in [k for k, (ref, _, _) in self.entries.items() if ref() is None]:
            del self.entries[k]
        b, c, h, w = x.shape
        src = x.detach()
        y = torch.empty((b, c, h, w), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        with torch.cuda.device(x.device):
            rc = _ext.lib().b200_nchw_to_nhwc(_ext.ptr(src), _ext.ptr(y), b, c, h, w, _ext.stream_ptr(x.device))
        _ext.check(rc, "b200_nchw_to_nhwc")
        if len(self.entries) >= self.capacity:
            self.entries.clear()
        self.entries[key] = (weakref.ref(x), x._version, y)
        return y


nhwc_cache = _NhwcCache()


def _stageable(f, sampling_ratio):
    """NCHW-contiguous CUDA map that the NHWC kernels could take if it were NHWC: the marching / row-streaming
    kernels at sampling_ratio 2, the tiled gather (roi_align_fwd_tile) at any other ratio -- e.g. the
    reference's shipped C4 pooler, sampling_ratio 0 on a [1, 1024, H/16, W/16] map (config/defaults.py:301-305)."""
    return (f.is_cuda and f.dtype == torch.float32 and f.dim() == 4 and sampling_ratio >= 0 and f.size(1) % 64 == 0 and
            f.size(1) > 1 and f.is_contiguous() and not f.is_contiguous(memory_format=torch.channels_last))


class _ROIAlignMulti(Function):
    """Fused level-assignment + RoIAlign over a feature pyramid (reference
    modeling/poolers.py:91-121 + layers/roi_align.py:12-45).  With `stage`, NCHW-contiguous
    maps run through a cached NHWC copy (forward) and an NHWC gradient buffer that is laid
    back out to NCHW (backward), so both directions use the marching kernels."""

    @staticmethod
    def forward(ctx, rois, output_size, scales, sampling_ratio, stage, math, *feats):
        output_size = _pair(output_size)
        staged = [bool(stage) and _stageable(f, sampling_ratio) for f in feats]
        run = [nhwc_cache.get(f) if s else f for f, s in zip(feats, staged)]
        out, _ = _forward(run, scales, rois, output_size, sampling_ratio, math=math)
        ctx.in_dtype = feats[0].dtype if feats else torch.float32
        if ctx.in_dtype != torch.float32:
            out = out.to(ctx.in_dtype)     # the Pooler hands back the input dtype (reference poolers.py:104-109)
        ctx.save_for_backward(rois.float().contiguous())
        ctx.output_size = output_size
        ctx.scales = tuple(scales)
        ctx.sampling_ratio = sampling_ratio
        ctx.shapes = [tuple(f.shape) for f in feats]
        lay = {_layout_of(f)[0] for f in run}
        ctx.nhwc = lay == {_ext.B200_LAYOUT_NHWC}
        ctx.staged = staged
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        # (gradients are accumulated in fp32 -- atomics on the fp32 buffer -- and rounded once at the end)
        grads = _backward(grad_output, rois, ctx.shapes, ctx.nhwc, ctx.scales, ctx.output_size,
                          ctx.sampling_ratio)
        if ctx.nhwc and any(ctx.staged):
            # the caller's maps are NCHW-contiguous: hand back gradients in that layout
            lib = _ext.lib()
            out = []
            for g, s in zip(grads, ctx.staged):
                if s:
                    b, c, h, w = g.shape
                    n = torch.empty((b, c, h, w), dtype=g.dtype, device=g.device)
                    with torch.cuda.device(g.device):
                        rc = lib.b200_nhwc_to_nchw(_ext.ptr(g), _ext.ptr(n), b, c, h, w, _ext.stream_ptr(g.device))
                    _ext.check(rc, "b200_nhwc_to_nchw")
                    g = n
                out.append(g)
            grads = out
        if ctx.in_dtype != torch.float32:
            grads = [g.to(ctx.in_dtype) for g in grads]
        return (None, None, None, None, None, None, *grads)


def roi_align_multilevel(feats, rois, output_size, scales, sampling_ratio, stage_nhwc=False, math=None):
    """feats: list of [B,C,H_l,W_l]; rois [R,5] -> [R,C,PH,PW] in RoI order.
    stage_nhwc: run NCHW-contiguous maps through a cached NHWC copy (see _ROIAlignMulti).
    math: "exact" | "fast" | None (the process default, see set_roi_align_math)."""
    return _ROIAlignMulti.apply(rois, output_size, tuple(scales), sampling_ratio, stage_nhwc, math, *feats)


def roi_align_with_mean(feats, rois, output_size, scales, sampling_ratio, math=None):
    """Inference-only fused pooler + AvgPool2d(pooled size): -> (pooled [R,C,PH,PW], mean [R,C]).
    The mean is what FastRCNNPredictor.forward's avgpool (roi_box_predictors.py:62) computes from
    the pooled block; here it leaves the pooler's shared-memory tile directly."""
    feats = list(feats)
    mean = torch.empty((rois.size(0), feats[0].size(1)), dtype=torch.float32, device=feats[0].device)
    out, _ = _forward(feats, scales, rois, _pair(output_size), sampling_ratio, math=math, mean_out=mean)
    return out, mean


def roi_align(input, roi, output_size, spatial_scale, sampling_ratio):
    """Same call as the reference's `roi_align = _ROIAlign.apply` (layers/roi_align.py:48)."""
    return _ROIAlignMulti.apply(roi, output_size, (spatial_scale,), sampling_ratio, False, None, input)


class ROIAlign(nn.Module):
    """Drop-in for maskrcnn_benchmark.layers.ROIAlign (reference layers/roi_align.py:50-69)."""

    def __init__(self, output_size, spatial_scale, sampling_ratio, math=None):
        super(ROIAlign, self).__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.math = math  # None: process default (set_roi_align_math)

    def forward(self, input, rois):
        # the reference wraps this in amp.float_function: fp32 in, fp32 out
        return _ROIAlignMulti.apply(rois.float(), self.output_size, (self.spatial_scale,), self.sampling_ratio,
                                    False, self.math, input.float())

    def __repr__(self):
        return "%s(output_size=%s, spatial_scale=%s, sampling_ratio=%s)" % (
            self.__class__.__name__, self.output_size, self.spatial_scale, self.sampling_ratio)
