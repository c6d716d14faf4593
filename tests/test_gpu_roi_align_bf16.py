"""GPU parity of the bf16 RoIAlign path (BASELINE config #4; reference layers/roi_align.py:57: the op computes in
fp32 whatever the input precision): bf16 feature maps against the fp32 CPU oracle on the bf16-ROUNDED inputs, as
SURVEY section 5 ("Mixed precision") prescribes.  Tolerance of the fast-math kernels: rtol 1e-5 (+ atol 1e-6)."""
import numpy as np
import pytest
import torch

import oracle
from tests import synth

pytestmark = pytest.mark.gpu


def _pyramid_bf16(rng, b, c, img_h=800, img_w=1333):
    return [torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32)).cuda().to(torch.bfloat16)
            .contiguous(memory_format=torch.channels_last) for (h, w) in synth.fpn_shapes(img_h, img_w)]


@pytest.mark.timeout(120)
def test_bf16_maps_box_pooler_matches_oracle_on_rounded_inputs():
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward
    rng = np.random.default_rng(4401)
    feats = _pyramid_bf16(rng, 2, 256)
    rois = synth.make_rois(rng, 300, 2, smin=6.0)
    rois[3, 1:] += 9000            # outside
    rois[4, 3], rois[4, 4] = rois[4, 1] - 2, rois[4, 2] - 3    # malformed (x2 < x1, y2 < y1)
    want, wl = oracle.pooler_forward([f.float().cpu().contiguous().numpy() for f in feats], rois, synth.FPN_SCALES, 7, 7, 2)
    mean = torch.empty((len(rois), 256), device="cuda")
    got, lv = _forward(feats, synth.FPN_SCALES, torch.from_numpy(rois).cuda(), (7, 7), 2, want_levels=True, math="fast",
                       mean_out=mean)
    assert got.dtype == torch.float32
    assert np.array_equal(lv.cpu().numpy(), wl)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-6)
    assert np.all(got[3].cpu().numpy() == 0)
    assert torch.allclose(mean, got.mean(dim=(2, 3)), rtol=1e-5, atol=1e-6)
    # the same maps widened to fp32 go through the fp32 kernel: same numbers up to the fast-math tolerance
    wide, _ = _forward([f.float() for f in feats], synth.FPN_SCALES, torch.from_numpy(rois).cuda(), (7, 7), 2, math="fast")
    assert torch.allclose(got, wide, rtol=2e-5, atol=2e-6)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("res", [7, 14])
def test_pooler_takes_bf16_maps_and_returns_their_dtype(res):
    """Pooler under AMP: bf16 in -> bf16 out (reference poolers.py:104-109), gradients in bf16, values those of
    the fp32 op on the rounded inputs.  14x14 has no bf16 kernel and runs widened, like the reference."""
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import Pooler
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    rng = np.random.default_rng(4402 + res)
    feats = [f.requires_grad_(True) for f in _pyramid_bf16(rng, 2, 256, 400, 672)]
    rois = synth.make_rois(rng, 120, 2, 672, 400, smin=8.0, smax=500.0)
    boxes = [BoxList(torch.from_numpy(rois[i * 120:(i + 1) * 120, 1:].copy()).cuda(), (672, 400)) for i in range(2)]
    pool = Pooler((res, res), synth.FPN_SCALES, 2, math="fast")
    y = pool(feats, boxes)
    assert y.dtype == torch.bfloat16 and y.shape == (240, 256, res, res)
    want, _ = oracle.pooler_forward([f.detach().float().cpu().contiguous().numpy() for f in feats], rois, synth.FPN_SCALES, res, res, 2)
    np.testing.assert_allclose(y.detach().float().cpu().numpy(), want, rtol=1e-2, atol=1e-2)     # bf16 rounding of the output
    g = torch.from_numpy(rng.standard_normal(y.shape).astype(np.float32)).cuda().to(torch.bfloat16)
    y.backward(g)
    f32 = [f.detach().float().requires_grad_(True) for f in feats]
    pool(f32, boxes).backward(g.float())
    for a, b in zip(feats, f32):
        assert a.grad is not None and a.grad.dtype == torch.bfloat16
        assert torch.allclose(a.grad.float(), b.grad, rtol=1e-2, atol=1e-2 * float(b.grad.abs().max()))


def test_bf16_entry_rejects_shapes_without_a_kernel():
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward_bf16, _levels_array
    x = torch.zeros((1, 128, 20, 20), device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    rois = torch.tensor([[0, 1, 1, 30, 30]], device="cuda", dtype=torch.float32)
    assert _forward_bf16([x], (0.25,), rois, (7, 7), 2, False, None) is None          # 128 channels: widened instead
    out = torch.empty((1, 128, 7, 7), device="cuda")
    rc = _ext.lib().b200_roi_align_forward_bf16(_levels_array([x], (0.25,)), 1, _ext.B200_LAYOUT_NHWC, 1, 128, _ext.ptr(rois),
                                                1, 7, 7, 2, _ext.ptr(out), None, None, None, 0, _ext.stream_ptr(x.device))
    assert rc == _ext.B200_ERR_UNSUPPORTED
