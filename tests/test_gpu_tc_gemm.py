"""GPU parity of the persistent tcgen05 GEMM (csrc/tc_gemm.cu) through the C ABI: the emb_pred projection
(b200_linear_bf16; reference roi_box_predictors.py:63-66) against an fp64 product of the bf16-rounded
operands, its autograd form, and the SOFTMAX scoring epilogue against the one-CTA-per-tile kernel and the
numpy oracle (bars of BASELINE.json: 2e-2 absolute, top-1 >= 99.9 %)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _rand(g, shape, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("shape", [(1000, 768, 1024), (16000, 768, 256), (130, 100, 72), (1, 16, 8), (513, 257, 200),
                                   (4096, 2048, 768)])
@pytest.mark.parametrize("bias", [True, False])
def test_linear_bf16_matches_fp64(shape, bias):
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import linear_bf16
    r, n, k = shape
    g = torch.Generator().manual_seed(r + n + k)
    x, w = _rand(g, (r, k)), _rand(g, (n, k), 1.0 / k ** 0.5)
    b = torch.randn((n,), generator=g) if bias else None
    want = x.double() @ w.double().t() + (b.double() if bias else 0.0)
    y32, y16 = linear_bf16(x.cuda(), w.cuda(), b.cuda() if bias else None, want_f32=True, want_bf16=True)
    assert y32.shape == (r, n) and y16.dtype == torch.bfloat16
    # fp32 accumulation of exact bf16 products: error ~ K * 2^-24 relative to the magnitude of the terms
    assert float((y32.cpu().double() - want).abs().max()) <= 2e-4 * float(want.abs().max() + 1)
    assert torch.equal(y16.cpu(), y32.cpu().to(torch.bfloat16))


def test_linear_rejects_cpu_and_bad_shapes():
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import TensorCoreLinear, linear_bf16
    with pytest.raises(RuntimeError):
        linear_bf16(torch.zeros((4, 8)), torch.zeros((2, 8)))
    with pytest.raises(ValueError):
        linear_bf16(torch.zeros((4, 12), device="cuda"), torch.zeros((2, 12), device="cuda"))
    with pytest.raises(RuntimeError):
        TensorCoreLinear(8, 4)(torch.zeros((2, 8)))


def test_tensor_core_linear_module_and_autograd():
    """Same parameters / state dict as nn.Linear; no-grad calls and bf16 autograd run on tcgen05 (forward and
    both gradient GEMMs), fp32 autograd keeps torch's GEMM."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import TensorCoreLinear
    g = torch.Generator().manual_seed(3)
    lin = TensorCoreLinear(1024, 768).cuda()
    ref = torch.nn.Linear(1024, 768).cuda()
    ref.load_state_dict(lin.state_dict())
    x = torch.randn((515, 1024), generator=g).cuda()
    with torch.no_grad():
        y = lin(x)
        want = ref(x)
    assert y.dtype == torch.float32 and hasattr(y, "b200_bf16")
    assert float((y - want).abs().max()) < 3e-2 * float(want.abs().max())        # bf16 operand rounding
    xr = x.clone().requires_grad_(True)
    assert torch.equal(lin(xr), ref(xr))                                         # fp32 autograd: torch's own GEMM
    # bf16 autograd: all three GEMMs on the tensor cores
    xb = x.to(torch.bfloat16).requires_grad_(True)
    yb = lin(xb)
    w = torch.randn(yb.shape, generator=g).cuda().to(torch.bfloat16)
    (yb.float() * w.float()).sum().backward()
    x64, w64 = xb.detach().double(), lin.weight.detach().to(torch.bfloat16).double()
    gx = w.double() @ w64
    gw = w.double().t() @ x64
    assert float((xb.grad.double() - gx).abs().max()) <= 2e-2 * float(gx.abs().max())
    assert float((lin.weight.grad.double() - gw).abs().max()) <= 2e-2 * float(gw.abs().max())
    assert float((lin.bias.grad.double() - w.double().sum(0)).abs().max()) <= 1e-3 * float(w.double().sum(0).abs().max())


@pytest.mark.parametrize("shape", [(1000, 66, 768), (4096, 501, 512), (700, 512, 64), (300, 257, 128), (129, 300, 256),
                                   (5, 2, 8), (2000, 256, 768), (640, 17, 72)])
def test_softmax_epilogue_vs_legacy_kernel_and_oracle(shape):
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import embed_match_softmax
    r, c, d = shape
    g = torch.Generator().manual_seed(17 * r + c)
    A = (torch.randn((r, d), generator=g) * (3.0 / d ** 0.5)).to(torch.bfloat16)
    E = torch.nn.functional.normalize(torch.randn((c, d), generator=g), dim=-1) * 3.0
    E[0] = 0
    E = E.to(torch.bfloat16)
    new = embed_match_softmax(A.cuda(), E.cuda(), 0.05, want_probs=True, want_logits=True)
    _ext.debug_match(True)
    try:
        old = embed_match_softmax(A.cuda(), E.cuda(), 0.05, want_probs=True, want_logits=True)
    finally:
        _ext.debug_match(False)
    logits = A.double() @ E.double().t()
    want = torch.softmax(logits, -1)
    assert float((new["logits"].cpu().double() - logits).abs().max()) < 1e-3 * float(logits.abs().max() + 1)
    assert float((new["probs"].cpu().double() - want).abs().max()) < 2e-3          # fp16 stash / fp32 TMEM values
    assert float((new["probs"] - old["probs"]).abs().max()) < 2e-3
    assert float((new["probs"].sum(1) - 1).abs().max()) < 2e-3
    if c > 1:
        top = want[:, 1:].argmax(1) + 1
        tp = want.gather(1, top[:, None])[:, 0]
        agree = (new["top_label"].cpu().long() == torch.where(tp > 0.05, top, torch.zeros_like(top)))
        second = want[:, 1:].topk(min(2, c - 1), dim=1).values
        clear = ((second[:, 0] - second[:, -1]) > 1e-3) | (c == 2)
        clear &= (tp - 0.05).abs() > 1e-3
        assert bool(agree[clear].all())
        assert float((new["top_prob"].cpu().double() - tp).abs().max()) < 2e-3
    assert torch.equal(new["top_label"], old["top_label"]) or float((new["top_label"] != old["top_label"]).float().mean()) < 1e-3
