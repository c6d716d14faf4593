"""RoIPool (max) mirroring maskrcnn_benchmark.layers.roi_pool (reference layers/roi_pool.py:12-58)."""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from .. import _ext


class _ROIPool(Function):
    @staticmethod
    def forward(ctx, input, roi, output_size, spatial_scale):
        _ext.require_cuda(input, "input")
        _ext.require_cuda(roi, "rois")
        ph, pw = _pair(output_size)
        x = input.float().contiguous()
        roi = roi.float().contiguous()
        r = roi.size(0)
        b, c, h, w = x.shape
        out = torch.empty((r, c, ph, pw), dtype=torch.float32, device=x.device)
        argmax = torch.empty((r, c, ph, pw), dtype=torch.int32, device=x.device)
        if r > 0:
            with torch.cuda.device(x.device):
                rc = _ext.lib().b200_roi_pool_forward(_ext.ptr(x), b, c, h, w, _ext.ptr(roi), r, float(spatial_scale),
                                                      ph, pw, _ext.ptr(out), _ext.ptr(argmax),
                                                      _ext.stream_ptr(x.device))
            _ext.check(rc, "b200_roi_pool_forward")
        ctx.save_for_backward(roi, argmax)
        ctx.output_size = (ph, pw)
        ctx.input_shape = (b, c, h, w)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        roi, argmax = ctx.saved_tensors
        ph, pw = ctx.output_size
        b, c, h, w = ctx.input_shape
        g = grad_output.float().contiguous()
        grad_in = torch.zeros((b, c, h, w), dtype=torch.float32, device=g.device)
        if roi.size(0) > 0:
            with torch.cuda.device(g.device):
                rc = _ext.lib().b200_roi_pool_backward(_ext.ptr(g), _ext.ptr(argmax), _ext.ptr(roi), roi.size(0), b, c,
                                                       h, w, ph, pw, _ext.ptr(grad_in), _ext.stream_ptr(g.device))
            _ext.check(rc, "b200_roi_pool_backward")
        return grad_in, None, None, None


roi_pool = _ROIPool.apply


class ROIPool(nn.Module):
    def __init__(self, output_size, spatial_scale):
        super(ROIPool, self).__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale

    def forward(self, input, rois):
        return roi_pool(input, rois, self.output_size, self.spatial_scale)

    def __repr__(self):
        return "%s(output_size=%s, spatial_scale=%s)" % (self.__class__.__name__, self.output_size, self.spatial_scale)
