"""GPU parity: the tiled gather (roi_align_fwd_tile: NHWC maps, any sampling ratio, C % 64 == 0) -- the kernel behind
the reference's shipped C4 pooler (14 x 14 bins, sampling_ratio 0, config/defaults.py:301-305) -- against the CPU
oracle.  Exact mode: bit-identical; fast mode (FMA contraction): rtol 1e-5 + atol 1e-6."""
import numpy as np
import pytest
import torch

import oracle
from tests import synth

pytestmark = pytest.mark.gpu


def _ext():
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    return _ext


def _feat(rng, b, c, h, w):
    x = torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32)).cuda()
    return x.contiguous(memory_format=torch.channels_last)


CASES = [
    # (B, C, H, W, scale, PH, PW, sr, n_rois)
    (1, 128, 50, 84, 1 / 16, 14, 14, 0, 120),   # config #1 geometry, fewer channels
    (2, 64, 50, 84, 1 / 16, 7, 7, 0, 80),
    (2, 64, 38, 50, 1 / 16, 14, 14, 1, 60),
    (1, 192, 25, 42, 1 / 32, 7, 7, 3, 60),
    (1, 64, 30, 40, 1 / 16, 3, 5, 4, 40),
    (1, 64, 20, 20, 1 / 4, 1, 1, 0, 20),
    (1, 64, 60, 60, 1 / 8, 16, 16, 0, 30),
]


@pytest.mark.parametrize("case", CASES)
def test_tile_kernel_exact(case):
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import roi_align
    b, c, h, w, scale, ph, pw, sr, n = case
    rng = np.random.default_rng(abs(hash(case)) % 2**31)
    x = _feat(rng, b, c, h, w)
    rois = synth.make_rois(rng, n, b, int(w / scale), int(h / scale), smin=4.0, smax=float(max(w, h) / scale))
    rois[:3, 3] = rois[:3, 1] - 7           # malformed (x2 < x1): forced to 1x1 (ROIAlign_cpu.cpp:156)
    rois[3:6, 1:] += 5000                   # fully outside: zeros (ROIAlign_cpu.cpp:47)
    rois[6, 1:] = [-40.0, -40.0, 30.0, 25.0]   # straddles the corner: some samples out of range
    want = oracle.roi_align_forward(x.cpu().contiguous().numpy(), rois, scale, ph, pw, sr)
    _ext().debug_set(False, True, 0)
    got = roi_align(x, torch.from_numpy(rois).cuda(), (ph, pw), scale, sr)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.all(want[3:6] == 0)
    # the plain gather (variant bit 14) is the same arithmetic
    _ext().debug_set(False, True, 16384)
    got2 = roi_align(x, torch.from_numpy(rois).cuda(), (ph, pw), scale, sr)
    _ext().debug_set(False, True, 0)
    assert np.array_equal(got2.cpu().numpy(), want)


def test_tile_kernel_fast_math_and_mean():
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import roi_align_with_mean
    rng = np.random.default_rng(5)
    x = _feat(rng, 1, 128, 50, 84)
    rois = synth.make_rois(rng, 150, 1, 1333, 800, smin=8.0, smax=800.0)
    want = oracle.roi_align_forward(x.cpu().contiguous().numpy(), rois, 1 / 16, 14, 14, 0)
    for math in ("exact", "fast"):
        got, mean = roi_align_with_mean([x], torch.from_numpy(rois).cuda(), (14, 14), (1 / 16,), 0, math=math)
        got, mean = got.cpu().numpy(), mean.cpu().numpy()
        if math == "exact":
            assert np.array_equal(got, want)
        else:
            assert np.allclose(got, want, rtol=1e-5, atol=1e-6)
        # fused channel mean == AvgPool2d(14) of the pooled block (sequential fp32 sum, one division)
        ref = torch.nn.functional.avg_pool2d(torch.from_numpy(got), 14).reshape(got.shape[0], -1).numpy()
        assert np.allclose(mean, ref, rtol=1e-6, atol=1e-7)


def test_tile_kernel_oversized_roi_takes_the_generic_path_inside_the_launch():
    """PH * grid_h beyond the axis tables (a huge RoI on a fine map): that RoI is gathered without tables, the
    others through the tile; all bit-identical to the oracle."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import roi_align_with_mean
    rng = np.random.default_rng(11)
    x = _feat(rng, 1, 64, 200, 260)
    rois = synth.make_rois(rng, 12, 1, 260, 200, smin=4.0, smax=60.0)
    rois[0, 1:] = [3.0, 5.0, 250.0, 190.0]      # 185 rows / 14 bins -> grid 14 -> 196 samples > 128
    rois[1, 1:] = [10.0, 2.0, 40.0, 198.0]      # tall: only the y axis overflows
    want = oracle.roi_align_forward(x.cpu().contiguous().numpy(), rois, 1.0, 14, 14, 0)
    got, mean = roi_align_with_mean([x], torch.from_numpy(rois).cuda(), (14, 14), (1.0,), 0, math="exact")
    assert np.array_equal(got.cpu().numpy(), want)
    ref = torch.nn.functional.avg_pool2d(got.cpu(), 14).reshape(got.shape[0], -1).numpy()
    assert np.allclose(mean.cpu().numpy(), ref, rtol=1e-6, atol=1e-7)


def test_pooler_stages_nchw_maps_for_adaptive_sampling():
    """The reference's C4 pooler on an NCHW-contiguous map (its own layout): the Pooler runs it through the cached
    NHWC copy + the tiled gather and returns exactly the reference's numbers; gradients come back NCHW."""
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import Pooler
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    rng = np.random.default_rng(21)
    x = torch.from_numpy(rng.standard_normal((2, 128, 50, 84)).astype(np.float32)).cuda().requires_grad_(True)
    rois = synth.make_rois(rng, 40, 2, 1333, 800, smin=16.0, smax=700.0)
    boxes = [BoxList(torch.from_numpy(rois[rois[:, 0] == i, 1:]).cuda(), (1333, 800), mode="xyxy") for i in range(2)]
    pooler = Pooler((14, 14), (1 / 16,), 0)
    got = pooler([x], boxes)
    want = oracle.roi_align_forward(x.detach().cpu().numpy(), rois, 1 / 16, 14, 14, 0)
    assert np.array_equal(got.detach().cpu().numpy(), want)
    g = torch.from_numpy(rng.standard_normal(got.shape).astype(np.float32)).cuda()
    got.backward(g)
    assert x.grad is not None and x.grad.is_contiguous() and x.grad.shape == x.shape
    _, want64 = oracle.roi_align_backward(g.cpu().numpy(), rois, 1 / 16, 14, 14, 2, 128, 50, 84, 0)
    _, mag = oracle.roi_align_backward(np.abs(g.cpu().numpy()), rois, 1 / 16, 14, 14, 2, 128, 50, 84, 0)
    err = np.abs(x.grad.cpu().numpy().astype(np.float64) - want64)
    assert np.all(err <= 1e-5 * mag + 1e-30), float((err / (mag + 1e-30)).max())
