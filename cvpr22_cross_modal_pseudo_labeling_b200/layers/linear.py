"""nn.Linear on the tcgen05 tensor cores (C ABI b200_linear_bf16, csrc/tc_gemm.cu).

Replaces, in the reference, the `emb_pred` projection that runs in front of the class-embedding
scoring: `cls_emb = self.emb_pred(x)` (modeling/roi_heads/box_head/roi_box_predictors.py:63-66) and
the direct call `predictor.emb_pred(f_regions)` of generate_pseudo_label
(modeling/detector/st_generalized_rcnn.py:226-228) -- SURVEY 8f-2.  bf16 operands, fp32 accumulation,
fp32 bias; the gradient GEMMs (grad_x = g . W, grad_W = g^T . x) run on the same kernel with
transposed operands.
"""
import torch
from torch import nn
from torch.autograd import Function

from .. import _ext


def _as_bf16(t):
    return t if (t.dtype == torch.bfloat16 and t.is_contiguous()) else t.to(torch.bfloat16).contiguous()


def linear_bf16(x, weight, bias=None, want_f32=True, want_bf16=False):
    """x [R, in], weight [out, in] (nn.Linear layout), bias [out] or None -> (y fp32 or None, y bf16 or None).
    Operands are rounded to bf16 (no copy when they already are), products accumulate in fp32."""
    _ext.require_cuda(x, "x")
    _ext.require_cuda(weight, "weight")
    if x.dim() != 2 or weight.dim() != 2 or x.size(1) != weight.size(1):
        raise ValueError("x must be [R, in] and weight [out, in]")
    if x.size(1) % 8:
        raise ValueError("the input dimension must be a multiple of 8 (16-byte rows for TMA)")
    if not (want_f32 or want_bf16):
        raise ValueError("no output requested")
    a, w = _as_bf16(x), _as_bf16(weight)
    b = None if bias is None else bias.detach().float().contiguous()
    r, n = a.size(0), w.size(0)
    dev = a.device
    y32 = torch.empty((r, n), dtype=torch.float32, device=dev) if want_f32 else None
    y16 = torch.empty((r, n), dtype=torch.bfloat16, device=dev) if want_bf16 else None
    if r > 0 and n > 0:
        with torch.cuda.device(dev):
            rc = _ext.lib().b200_linear_bf16(_ext.ptr(a), _ext.ptr(w), _ext.ptr(b), r, n, a.size(1), _ext.ptr(y16),
                                             _ext.ptr(y32), _ext.stream_ptr(dev))
        _ext.check(rc, "b200_linear_bf16")
    return y32, y16


class _LinearTC(Function):
    """Differentiable y = x . W^T + b with all three GEMMs on the tensor cores."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        y32, _ = linear_bf16(x, weight, bias)
        return y32 if x.dtype == torch.float32 else y32.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g16 = _as_bf16(g)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:   # [R, out] . [out, in]: the "weight" operand is W^T laid out [in, out]
            gx = linear_bf16(g16, weight.detach().t(), None)[0].to(x.dtype)
        if ctx.needs_input_grad[1]:   # [out, R] . [R, in]: operands g^T [out, R] and x^T [in, R]
            r = g16.size(0)
            pad = (-r) % 8            # the contraction runs over the R rows: 16-byte row pitch for TMA
            gt, xt = g16.t(), _as_bf16(x.detach()).t()
            if pad:
                gt = torch.nn.functional.pad(gt, (0, pad))
                xt = torch.nn.functional.pad(xt, (0, pad))
            gw = linear_bf16(gt.contiguous(), xt.contiguous(), None)[0].to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g.float().sum(0)
        return gx, gw, gb


class TensorCoreLinear(nn.Linear):
    """Drop-in nn.Linear (same parameters / state dict) whose product runs on tcgen05 when the input is a
    CUDA tensor: always without autograd (the teacher, eval mode); under autograd only for bf16 inputs
    (the AMP / config #4 path -- dtype follows the input, fp32 training keeps torch's fp32 GEMM and its exact
    semantics).  The bf16 copy of the weight is cached per parameter version."""

    def __init__(self, *args, **kwargs):
        super(TensorCoreLinear, self).__init__(*args, **kwargs)
        self._w16 = None

    def _weight_bf16(self):
        w = self.weight
        key = (w.data_ptr(), w._version, w.device)
        if self._w16 is None or self._w16[0] != key:
            self._w16 = (key, w.detach().to(torch.bfloat16).contiguous())
        return self._w16[1]

    def forward(self, x):
        lead = x.shape[:-1]
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad or
                                                  (self.bias is not None and self.bias.requires_grad))
        if not x.is_cuda or self.in_features % 8 or (needs_grad and x.dtype != torch.bfloat16):
            if not x.is_cuda:
                raise RuntimeError("input must be a CUDA tensor (the B200 path has no CPU implementation)")
            return nn.functional.linear(x, self.weight, self.bias)
        x2 = x.reshape(-1, self.in_features)
        if needs_grad:
            y = _LinearTC.apply(x2, self.weight, self.bias)
        else:
            y32, y16 = linear_bf16(x2, self._weight_bf16(), self.bias, want_f32=True, want_bf16=True)
            y = y32 if x.dtype == torch.float32 else y32.to(x.dtype)
            y.b200_bf16 = y16   # the scoring kernel takes this as is (no second rounding pass)
        return y.reshape(*lead, self.out_features) if len(lead) != 1 else y
