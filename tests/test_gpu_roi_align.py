"""GPU parity: RoIAlign forward/backward, fused multi-level Pooler, layouts, RoIPool -- through
the C ABI, against the CPU oracle.  Forward tolerance: bit-exact in the default (exact)
mode; rtol 1e-5 (+ atol 1e-6 on unit-variance features) in the fast mode
(b200_roi_align_forward_fast: separable FMA evaluation).  Backward: rtol 1e-5 of the fp64-accumulated
oracle, scaled per element by the sum of |addends| (atomics reorder fp32 sums)."""
import numpy as np
import pytest
import torch

import oracle
from tests import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _ext():
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    return _ext


def _feat(rng, b, c, h, w, nhwc):
    x = torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32)).cuda()
    return x.contiguous(memory_format=torch.channels_last) if nhwc else x


def _rand_rois(rng, n, b, img_w, img_h):
    return synth.make_rois(rng, n, b, img_w, img_h, smin=4.0, smax=float(max(img_w, img_h)))


CASES = [
    # (B, C, H, W, scale, PH, PW, sr, n_rois)
    (2, 64, 50, 84, 1 / 16, 7, 7, 2, 60),
    (2, 128, 25, 42, 1 / 32, 14, 14, 2, 40),
    (1, 64, 100, 168, 1 / 8, 7, 7, 2, 80),
    (1, 32, 50, 84, 1 / 16, 14, 14, 0, 50),   # C4-style adaptive sampling (config #1 shape, fewer channels)
    (2, 16, 30, 40, 1 / 16, 3, 5, 3, 30),
    (1, 8, 20, 20, 1 / 4, 1, 1, 1, 10),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("nhwc", [False, True])
def test_roi_align_forward_exact(case, nhwc):
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import roi_align
    b, c, h, w, scale, ph, pw, sr, n = case
    rng = np.random.default_rng(hash(case) % 2**31)
    x = _feat(rng, b, c, h, w, nhwc)
    rois = _rand_rois(rng, n, b, int(w / scale), int(h / scale))
    rois[:3, 3] = rois[:3, 1] - 7           # malformed (x2 < x1): forced to 1x1 (ROIAlign_cpu.cpp:156)
    rois[3:6, 1:] += 5000                   # fully outside: zeros (ROIAlign_cpu.cpp:47)
    want = oracle.roi_align_forward(x.cpu().contiguous().numpy(), rois, scale, ph, pw, sr)
    _ext().debug_set(False, True, 0)
    got = roi_align(x, torch.from_numpy(rois).cuda(), (ph, pw), scale, sr)
    assert got.is_contiguous() and got.shape == want.shape
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.all(want[3:6] == 0)
    # generic kernel must agree with the staged one too
    _ext().debug_set(True, True, 0)
    got2 = roi_align(x, torch.from_numpy(rois).cuda(), (ph, pw), scale, sr)
    _ext().debug_set(False, True, 0)
    assert np.array_equal(got2.cpu().numpy(), want)


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("res", [7, 14])
def test_roi_align_forward_occupancy_variants(variant, res):
    """Every register/occupancy variant of the marching kernel must stay bit-exact."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import roi_align
    rng = np.random.default_rng(variant)
    x = _feat(rng, 2, 128, 50, 84, True)
    rois = _rand_rois(rng, 100, 2, 1333, 800)
    want = oracle.roi_align_forward(x.cpu().contiguous().numpy(), rois, 1 / 16, res, res, 2)
    _ext().debug_set(False, True, variant)
    got = roi_align(x, torch.from_numpy(rois).cuda(), (res, res), 1 / 16, 2)
    _ext().debug_set(False, True, 0)
    assert np.array_equal(got.cpu().numpy(), want)


def test_roi_align_forward_fma_mode_tolerance():
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import roi_align
    rng = np.random.default_rng(3)
    x = _feat(rng, 2, 64, 50, 84, True)
    rois = _rand_rois(rng, 200, 2, 1333, 800)
    want = oracle.roi_align_forward(x.cpu().contiguous().numpy(), rois, 1 / 16, 7, 7, 2)
    _ext().debug_set(False, False, 0)
    got = roi_align(x, torch.from_numpy(rois).cuda(), (7, 7), 1 / 16, 2).cpu().numpy()
    _ext().debug_set(False, True, 0)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("nhwc", [False, True])
def test_roi_align_forward_fast_math(case, nhwc):
    """b200_roi_align_forward_fast on every shape class: the separable marching kernel (NHWC,
    sampling_ratio 2) and the FMA-contracted generic gather (everything else)."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward
    b, c, h, w, scale, ph, pw, sr, n = case
    rng = np.random.default_rng(hash(case) % 2**31 + 1)
    x = _feat(rng, b, c, h, w, nhwc)
    rois = _rand_rois(rng, n, b, int(w / scale), int(h / scale))
    rois[:3, 3] = rois[:3, 1] - 7
    rois[3:6, 1:] += 5000
    rois[6, 1:] = [-30.0, -30.0, 40.0, 50.0]                # straddles the top-left corner (clamped taps)
    rois[7, 1:] = [w / scale - 20, h / scale - 20, w / scale + 60, h / scale + 60]   # bottom-right corner
    want = oracle.roi_align_forward(x.cpu().contiguous().numpy(), rois, scale, ph, pw, sr)
    got, _ = _forward([x], (scale,), torch.from_numpy(rois).cuda(), (ph, pw), sr, math="fast")
    got = got.cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-6)
    assert np.all(got[3:6] == 0)


@pytest.mark.parametrize("res", [7, 14])
@pytest.mark.parametrize("channels", [64, 128, 192, 256])
def test_pooler_multilevel_fast_math(res, channels):
    """Pooler(math="fast") over the pyramid: levels identical, values within 1e-5 of the oracle;
    192 channels = 3 chunks exercises the one-chunk-per-CTA grid, 128 / 256 the 2- / 4-chunk CTAs."""
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import Pooler
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    rng = np.random.default_rng(170 + res + channels)
    b, n = 2, 120
    feats = _pyramid(rng, b, channels, True)
    rois = synth.make_rois(rng, n, b)
    boxes = [BoxList(torch.from_numpy(rois[i * n:(i + 1) * n, 1:]).cuda(), (synth.IMG_W, synth.IMG_H)) for i in range(b)]
    got = Pooler((res, res), synth.FPN_SCALES, 2, math="fast")(feats, boxes).cpu().numpy()
    want, _ = oracle.pooler_forward([f.cpu().contiguous().numpy() for f in feats], rois, synth.FPN_SCALES, res, res, 2)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-6)


def test_roi_align_math_default_and_autograd():
    """set_roi_align_math switches the process default; the fast forward feeds the same backward."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import ROIAlign, get_roi_align_math, roi_align, set_roi_align_math
    rng = np.random.default_rng(5)
    x = _feat(rng, 1, 64, 50, 84, True)
    rois = torch.from_numpy(_rand_rois(rng, 50, 1, 1333, 800)).cuda()
    assert get_roi_align_math() == "exact"
    exact = roi_align(x, rois, (7, 7), 1 / 16, 2)
    old = set_roi_align_math("fast")
    try:
        fast = roi_align(x, rois, (7, 7), 1 / 16, 2)
    finally:
        set_roi_align_math(old)
    assert get_roi_align_math() == "exact"
    assert not torch.equal(exact, fast)                     # different summation order ...
    assert torch.allclose(exact, fast, rtol=RTOL, atol=1e-6)   # ... same numbers
    assert torch.equal(ROIAlign((7, 7), 1 / 16, 2, math="fast")(x, rois), fast)
    with pytest.raises(ValueError):
        set_roi_align_math("sloppy")
    xg = x.clone().requires_grad_(True)
    ROIAlign((7, 7), 1 / 16, 2, math="fast")(xg, rois).sum().backward()
    xe = x.clone().requires_grad_(True)
    ROIAlign((7, 7), 1 / 16, 2, math="exact")(xe, rois).sum().backward()
    assert torch.allclose(xg.grad, xe.grad, rtol=1e-5, atol=1e-5)   # same kernel; atomics reorder the sums


@pytest.mark.parametrize("math", ["exact", "fast"])
@pytest.mark.parametrize("res", [7, 14])
@pytest.mark.parametrize("nhwc", [False, True])
def test_roi_align_fused_channel_mean(math, res, nhwc):
    """b200_roi_align_forward_ex's second output equals AvgPool2d(res) of the pooled block
    (roi_box_predictors.py:16-17,:62) -- marching kernels (NHWC) and generic path (NCHW)."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import roi_align_with_mean
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward
    rng = np.random.default_rng(31 + res)
    feats = _pyramid(rng, 2, 128, nhwc)
    rois = torch.from_numpy(synth.make_rois(rng, 90, 2)).cuda()
    pooled, mean = roi_align_with_mean(feats, rois, (res, res), synth.FPN_SCALES, 2, math=math)
    plain, _ = _forward(feats, synth.FPN_SCALES, rois, (res, res), 2, math=math)
    assert torch.equal(pooled, plain)                       # the extra output does not disturb the pooled block
    want = torch.nn.functional.avg_pool2d(pooled, res).flatten(1)
    assert mean.shape == want.shape
    assert torch.allclose(mean, want, rtol=1e-6, atol=1e-7)
    with pytest.raises(ValueError):
        _forward(feats, synth.FPN_SCALES, rois, (res, res), 2, mean_out=torch.empty((3, 3), device="cuda"))


def test_roi_align_known_answers():
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import roi_align
    x = torch.arange(2 * 6 * 6, dtype=torch.float32, device="cuda").reshape(1, 2, 6, 6)
    # RoI covering exactly pixel (y=2,x=3) cell centre: samples all inside one bilinear cell of a linear ramp
    rois = torch.tensor([[0, 3.0, 2.0, 4.0, 3.0]], device="cuda")
    out = roi_align(x, rois, (1, 1), 1.0, 2)
    # ramp is linear -> average of samples = value at the RoI centre (3.5, 2.5)
    assert torch.allclose(out[0, :, 0, 0].cpu(), torch.tensor([2.5 * 6 + 3.5, 36 + 2.5 * 6 + 3.5]))
    e = roi_align(x, torch.zeros((0, 5), device="cuda"), (7, 7), 1.0, 2)
    assert e.shape == (0, 2, 7, 7)
    with pytest.raises(RuntimeError):
        roi_align(x.cpu(), rois.cpu(), (1, 1), 1.0, 2)


def _pyramid(rng, b, c, nhwc, img_h=800, img_w=1333):
    shapes = synth.fpn_shapes(img_h, img_w)
    return [_feat(rng, b, c, h, w, nhwc) for (h, w) in shapes]


@pytest.mark.parametrize("nhwc", [False, True])
@pytest.mark.parametrize("res", [7, 14])
def test_pooler_multilevel_matches_oracle(nhwc, res):
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import Pooler
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward
    rng = np.random.default_rng(17 + res)
    b, c, n = 2, 64, 150
    feats = _pyramid(rng, b, c, nhwc)
    rois = synth.make_rois(rng, n, b)
    boxes = [BoxList(torch.from_numpy(rois[i * n:(i + 1) * n, 1:]).cuda(), (synth.IMG_W, synth.IMG_H)) for i in range(b)]
    pooler = Pooler((res, res), synth.FPN_SCALES, 2)
    got = pooler(feats, boxes)
    want, want_lv = oracle.pooler_forward([f.cpu().contiguous().numpy() for f in feats], rois, synth.FPN_SCALES, res, res, 2)
    _, lv = _forward(feats, synth.FPN_SCALES, torch.from_numpy(rois).cuda(), (res, res), 2, want_levels=True)
    lv = lv.cpu().numpy()
    mism = int((lv != want_lv).sum())
    assert mism == 0, "level mismatches: %d" % mism     # reported separately from value parity
    assert len(np.unique(want_lv)) == 4                  # all four levels populated
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.fixture(params=["march", "generic", "march_by_output_row"])
def bwd_path(request):
    """Backward runs on the marching kernel by tap row (NHWC, sampling_ratio 2: the default), on the per-tap kernel
    and on the round-1 marching kernel by output row."""
    _ext().debug_bwd({"march": 0, "generic": 1, "march_by_output_row": 2}[request.param])
    yield request.param
    _ext().debug_bwd(0)


@pytest.mark.parametrize("nhwc", [False, True])
@pytest.mark.parametrize("cfg", [(7, 7, 2), (14, 14, 2), (5, 3, 0)])
def test_roi_align_backward(nhwc, cfg, bwd_path):
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import roi_align
    ph, pw, sr = cfg
    rng = np.random.default_rng(23 + ph)
    b, c, h, w, scale, n = 2, 64, 25, 42, 1 / 32, 120
    x = _feat(rng, b, c, h, w, nhwc).requires_grad_(True)
    rois = _rand_rois(rng, n, b, 1333, 800)
    g = rng.standard_normal((n * b, c, ph, pw)).astype(np.float32)
    out = roi_align(x, torch.from_numpy(rois).cuda(), (ph, pw), scale, sr)
    out.backward(torch.from_numpy(g).cuda())
    got = x.grad.cpu().contiguous().numpy().astype(np.float64)
    _, want64 = oracle.roi_align_backward(g, rois, scale, ph, pw, b, c, h, w, sr)
    _, mag = oracle.roi_align_backward(np.abs(g), rois, scale, ph, pw, b, c, h, w, sr)   # sum of |addends|
    err = np.abs(got - want64)
    assert np.all(err <= RTOL * mag + 1e-30), float((err / (mag + 1e-30)).max())
    assert (want64 != 0).any()


def test_pooler_backward_multilevel(bwd_path):
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import Pooler
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    rng = np.random.default_rng(29)
    b, c, n = 1, 64, 200
    feats = [f.requires_grad_(True) for f in _pyramid(rng, b, c, True, 400, 672)]
    rois = synth.make_rois(rng, n, b, 672, 400, smin=8, smax=400)
    boxes = [BoxList(torch.from_numpy(rois[:, 1:]).cuda(), (672, 400))]
    pooler = Pooler((7, 7), synth.FPN_SCALES, 2)
    out = pooler(feats, boxes)
    g = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    out.backward(torch.from_numpy(g).cuda())
    lv = oracle.level_map(rois, 2.0, 5.0)
    for l, f in enumerate(feats):
        idx = np.nonzero(lv == l)[0]
        _, want = oracle.roi_align_backward(g[idx], rois[idx], synth.FPN_SCALES[l], 7, 7, b, c, f.shape[2], f.shape[3], 2)
        _, mag = oracle.roi_align_backward(np.abs(g[idx]), rois[idx], synth.FPN_SCALES[l], 7, 7, b, c, f.shape[2], f.shape[3], 2)
        err = np.abs(f.grad.cpu().contiguous().numpy() - want)
        assert np.all(err <= RTOL * mag + 1e-30), l


def test_layout_roundtrip():
    e = _ext()
    x = torch.randn(3, 70, 37, 53, device="cuda")
    y = torch.empty_like(x).contiguous(memory_format=torch.channels_last)
    e.check(e.lib().b200_nchw_to_nhwc(e.ptr(x), e.ptr(y), 3, 70, 37, 53, e.stream_ptr()), "to_nhwc")
    assert torch.equal(y, x)  # logical equality: y is channels_last memory
    z = torch.empty_like(x)
    e.check(e.lib().b200_nhwc_to_nchw(e.ptr(y), e.ptr(z), 3, 70, 37, 53, e.stream_ptr()), "to_nchw")
    assert torch.equal(z, x)


def test_roi_pool_matches_oracle():
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import ROIPool
    rng = np.random.default_rng(31)
    x = _feat(rng, 2, 16, 30, 40, False).requires_grad_(True)
    rois = _rand_rois(rng, 50, 2, 640, 480)
    out = ROIPool((7, 7), 1 / 16)(x, torch.from_numpy(rois).cuda())
    want, arg = oracle.roi_pool_forward(x.detach().cpu().numpy(), rois, 1 / 16, 7, 7)
    assert np.array_equal(out.detach().cpu().numpy(), want)
    g = rng.standard_normal(want.shape).astype(np.float32)
    out.backward(torch.from_numpy(g).cuda())
    wg = oracle.roi_pool_backward(g, arg, rois, 7, 7, 2, 16, 30, 40)
    np.testing.assert_allclose(x.grad.cpu().numpy(), wg, rtol=1e-5, atol=1e-6)


def test_pooler_nchw_training_path_is_staged_and_exact():
    """NCHW-contiguous maps that require grad: the forward runs on the cached NHWC copy (still
    bit-exact), the backward returns NCHW-contiguous gradients within tolerance."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import nhwc_cache
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import Pooler
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    rng = np.random.default_rng(41)
    b, c, n = 2, 64, 100
    feats = [f.requires_grad_(True) for f in _pyramid(rng, b, c, False, 400, 672)]
    rois = synth.make_rois(rng, n, b, 672, 400, smin=8, smax=400)
    boxes = [BoxList(torch.from_numpy(rois[i * n:(i + 1) * n, 1:]).cuda(), (672, 400)) for i in range(b)]
    pooler = Pooler((7, 7), synth.FPN_SCALES, 2)
    out = pooler(feats, boxes)
    assert any(e[0]() is feats[0] for e in nhwc_cache.entries.values())      # staged
    want, _ = oracle.pooler_forward([f.detach().cpu().numpy() for f in feats], rois, synth.FPN_SCALES, 7, 7, 2)
    assert np.array_equal(out.detach().cpu().numpy(), want)
    g = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    out.backward(torch.from_numpy(g).cuda())
    lv = oracle.level_map(rois, 2.0, 5.0)
    for l, f in enumerate(feats):
        assert f.grad.is_contiguous()
        idx = np.nonzero(lv == l)[0]
        _, w64 = oracle.roi_align_backward(g[idx], rois[idx], synth.FPN_SCALES[l], 7, 7, b, c, f.shape[2], f.shape[3], 2)
        _, mag = oracle.roi_align_backward(np.abs(g[idx]), rois[idx], synth.FPN_SCALES[l], 7, 7, b, c, f.shape[2], f.shape[3], 2)
        assert np.all(np.abs(f.grad.cpu().numpy() - w64) <= RTOL * mag + 1e-30), l
    # an in-place update invalidates the cached copy
    with torch.no_grad():
        feats[0].mul_(2.0)
    out2 = pooler(feats, boxes)
    want2, _ = oracle.pooler_forward([f.detach().cpu().numpy() for f in feats], rois, synth.FPN_SCALES, 7, 7, 2)
    assert np.array_equal(out2.detach().cpu().numpy(), want2)
