"""GPU parity of the row-streaming RoIAlign forward (csrc/roi_align_fwd_rows.cu: cp.async.bulk row
ring + register accumulators, the FPN box pooler shape 7x7 / sampling_ratio 2 / 256 NHWC channels)
through the C ABI, against the CPU oracle.  Fast-math tolerance: rtol 1e-5 (+ atol 1e-6 on
unit-variance features); FPN levels identical; RoIs outside the image exactly zero."""
import numpy as np
import pytest
import torch

import oracle
from tests import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-5
ROWS = 0   # the default fast path for this shape
SEP = 16   # b200_debug_set variant bit that forces the separable marching kernel instead


def _ext():
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    return _ext


def _pyramid(rng, b, c, img_h=800, img_w=1333):
    return [torch.from_numpy(rng.standard_normal((b, c, h, w)).astype(np.float32)).cuda()
            .contiguous(memory_format=torch.channels_last) for (h, w) in synth.fpn_shapes(img_h, img_w)]


def _fwd(feats, scales, rois, variant, want_levels=False, mean=False):
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import roi_align_with_mean
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward
    _ext().debug_set(False, True, variant)
    try:
        if mean:
            return roi_align_with_mean(feats, rois, (7, 7), scales, 2, math="fast")
        return _forward(feats, scales, rois, (7, 7), 2, want_levels=want_levels, math="fast")
    finally:
        _ext().debug_set(False, True, 0)


EDGE = np.array([
    [0, -50, -50, -20, -20], [0, 1400, 900, 1500, 1000], [1, 1300, 780, 1332, 799], [0, -30, 100, 40, 180],
    [1, 100, 100, 100.5, 100.5], [0, 0, 0, 1332, 799], [1, 300, 200, 290, 190], [0, 10.3, 20.7, 12.1, 22.2],
    [0, -200, -200, 1500, 1000], [1, 1280, 0, 1600, 30], [0, 0, 780, 30, 1200], [1, 5, 5, 9, 300],
    [0, 5000, 5000, 5100, 5100],
], dtype=np.float32)


# 0: RoIs visited in the library's own order (image, level, Morton code of the centre); 32: in the order given
LAYOUTS = [0, 32]


@pytest.mark.timeout(120)
@pytest.mark.parametrize("layout", LAYOUTS)
def test_rows_kernel_matches_oracle_multilevel(layout):
    rng = np.random.default_rng(2601)
    b, n = 2, 150
    feats = _pyramid(rng, b, 256)
    rois = np.concatenate([synth.make_rois(rng, n, b, smin=6.0), EDGE]).astype(np.float32)
    want, wl = oracle.pooler_forward([f.cpu().contiguous().numpy() for f in feats], rois, synth.FPN_SCALES, 7, 7, 2)
    got, lv = _fwd(feats, synth.FPN_SCALES, torch.from_numpy(rois).cuda(), ROWS | layout, want_levels=True)
    got = got.cpu().numpy()
    assert np.array_equal(lv.cpu().numpy(), wl)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-6)
    assert np.all(got[-1] == 0) and np.all(got[len(rois) - len(EDGE)] == 0)      # fully outside
    # and really a different kernel from the marching one: bitwise differences are expected somewhere
    ref, _ = _fwd(feats, synth.FPN_SCALES, torch.from_numpy(rois).cuda(), SEP)
    assert not np.array_equal(got, ref.cpu().numpy())
    np.testing.assert_allclose(ref.cpu().numpy(), want, rtol=RTOL, atol=1e-6)      # the fallback stays covered at 256 channels
    np.testing.assert_allclose(got, ref.cpu().numpy(), rtol=2 * RTOL, atol=2e-6)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("layout", LAYOUTS)
def test_rows_kernel_single_level_sparse_taps(layout):
    """One level, RoIs much wider than 28 feature pixels: the tapped columns split into many runs."""
    rng = np.random.default_rng(2602)
    x = torch.from_numpy(rng.standard_normal((2, 256, 100, 168)).astype(np.float32)).cuda() \
        .contiguous(memory_format=torch.channels_last)
    rois = synth.make_rois(rng, 80, 2, smin=150.0, smax=1300.0, degenerate=0.0)
    want = oracle.roi_align_forward(x.cpu().contiguous().numpy(), rois, 1 / 8, 7, 7, 2)
    got, _ = _fwd([x], (1 / 8,), torch.from_numpy(rois).cuda(), ROWS | layout)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=RTOL, atol=1e-6)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("n", [1, 3, 149, 700])
@pytest.mark.parametrize("layout", LAYOUTS)
def test_rows_kernel_small_grids_and_mean(n, layout):
    """Fewer RoIs than SMs, one more than a multiple, several per CTA; fused channel mean; run-to-run
    bit-identical (no atomics, fixed summation order)."""
    rng = np.random.default_rng(2603 + n)
    feats = _pyramid(rng, 1, 256, 400, 672)
    rois = torch.from_numpy(synth.make_rois(rng, n, 1, 672, 400, smin=8.0, smax=600.0)).cuda()
    want, _ = oracle.pooler_forward([f.cpu().contiguous().numpy() for f in feats], rois.cpu().numpy(),
                                    synth.FPN_SCALES, 7, 7, 2)
    pooled, mean = _fwd(feats, synth.FPN_SCALES, rois, ROWS | layout, mean=True)
    np.testing.assert_allclose(pooled.cpu().numpy(), want, rtol=RTOL, atol=1e-6)
    assert torch.allclose(mean, pooled.mean(dim=(2, 3)), rtol=1e-5, atol=1e-6)
    again, _ = _fwd(feats, synth.FPN_SCALES, rois, ROWS | layout)
    assert torch.equal(again, pooled)


@pytest.mark.timeout(120)
def test_rows_kernel_visiting_order_does_not_change_results():
    """The ordering pass only decides which SM takes which RoI when; every RoI is computed from its own
    tables and written to its own slot, so results are bit-identical with and without it, for RoIs given
    in any order (shuffled across images, more than one sort chunk of 4096)."""
    rng = np.random.default_rng(2607)
    feats = _pyramid(rng, 3, 256)
    rois = synth.make_rois(rng, 1700, 3)
    rois = torch.from_numpy(rois[rng.permutation(len(rois))]).cuda()
    a, la = _fwd(feats, synth.FPN_SCALES, rois, ROWS, want_levels=True)
    b, lb = _fwd(feats, synth.FPN_SCALES, rois, ROWS | 32, want_levels=True)
    assert torch.equal(a, b) and torch.equal(la, lb)
    want, wl = oracle.pooler_forward([f.cpu().contiguous().numpy() for f in feats], rois[:300].cpu().numpy(),
                                     synth.FPN_SCALES, 7, 7, 2)
    np.testing.assert_allclose(a[:300].cpu().numpy(), want, rtol=RTOL, atol=1e-6)
    assert np.array_equal(la[:300].cpu().numpy(), wl)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("bf16", [False, True])
def test_rows_kernel_wait_group_size_does_not_change_results(bf16):
    """The copies of up to k consecutive ring entries may complete on one barrier (one consumer wait per group;
    default 1 for fp32 maps, 4 for bf16 maps): scheduling only -- the results are bit-identical for every k."""
    rng = np.random.default_rng(2607)
    b, n = 2, 400
    feats = _pyramid(rng, b, 256)
    if bf16:
        feats = [f.to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for f in feats]
    rois = torch.from_numpy(np.concatenate([synth.make_rois(rng, n, b, smin=6.0), EDGE]).astype(np.float32)).cuda()
    outs = [_fwd(feats, synth.FPN_SCALES, rois, ROWS | (k << 10))[0] for k in (0, 1, 2, 3, 4, 7)]
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


@pytest.mark.timeout(180)
def test_rows_kernel_full_size_against_exact():
    """BASELINE config #2 size (16 images x 1000 RoIs, 5.8 GB of taps): within 1e-5 of the
    bit-exact marching kernel; constant features stay constant (weights sum to one)."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward
    rng = np.random.default_rng(2604)
    g = torch.Generator(device="cuda").manual_seed(2604)
    b = 16
    feats = [torch.randn((b, 256, h, w), device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
             for (h, w) in synth.fpn_shapes()]
    rois = torch.from_numpy(synth.make_rois(rng, 1000, b)).cuda()
    got, lv = _fwd(feats, synth.FPN_SCALES, rois, ROWS, want_levels=True)
    exact, le = _forward(feats, synth.FPN_SCALES, rois, (7, 7), 2, want_levels=True, math="exact")
    assert torch.equal(lv, le)
    err = (got - exact).abs()
    assert bool((err <= RTOL * exact.abs() + 1e-6).all()), float(err.max())
    # ... and the bench's configuration touches the oracle directly: 200 random RoIs of the full-size batch
    pick = np.sort(rng.choice(rois.shape[0], 200, replace=False))
    want, wl = oracle.pooler_forward([f.cpu().contiguous().numpy() for f in feats], rois[pick].cpu().numpy(),
                                     synth.FPN_SCALES, 7, 7, 2)
    assert np.array_equal(lv[pick].cpu().numpy(), wl)
    np.testing.assert_allclose(got[pick].cpu().numpy(), want, rtol=RTOL, atol=1e-6)
    assert np.array_equal(exact[pick].cpu().numpy(), want)   # the exact kernel: bit-identical
    ones = [torch.full_like(f, 3.25) for f in feats]
    del feats
    inside = (rois[:, 1] >= 0) & (rois[:, 2] >= 0) & (rois[:, 3] <= synth.IMG_W - 1) & (rois[:, 4] <= synth.IMG_H - 1)
    c, _ = _fwd(ones, synth.FPN_SCALES, rois, ROWS)
    assert torch.allclose(c[inside], torch.full_like(c[inside], 3.25), rtol=1e-6, atol=0)
