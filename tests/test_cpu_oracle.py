"""CPU suite: the oracle against the committed golden vectors (generated from the reference
itself, tests/golden/make_golden.py) and, when oracle/_ref is present, against the compiled
reference live."""
import os

import numpy as np
import pytest

import oracle
from tests import synth

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


def test_roi_align_oracle_matches_golden():
    g = _load("roi_align.npz")
    for i in range(3):
        scale, ph, pw, sr = g["ra%d_cfg" % i]
        got = oracle.roi_align_forward(g["ra%d_x" % i], g["ra%d_rois" % i], float(scale), int(ph), int(pw), int(sr))
        assert np.array_equal(got, g["ra%d_out" % i]), i


def test_nms_oracle_matches_golden():
    g = _load("nms.npz")
    for i in range(4):
        got = oracle.nms(g["nms%d_boxes" % i], g["nms%d_scores" % i], float(g["nms%d_thr" % i][0]))
        assert np.array_equal(got, g["nms%d_keep" % i]), i


def test_pooler_oracle_matches_golden():
    g = _load("pooler.npz")
    feats = [g["pool_f%d" % l] for l in range(4)]
    for res in (7, 14):
        got, lv = oracle.pooler_forward(feats, g["pool_rois"], synth.FPN_SCALES, res, res, 2)
        assert np.array_equal(got, g["pool_out%d" % res])
        assert len(np.unique(lv)) >= 3


def test_box_decode_oracle_matches_golden_teacher_boxes():
    g = _load("box_head.npz")
    rois, reg = g["bh_rois"], g["bh_reg"]
    n = 300
    dec = oracle.box_decode(reg[:n, -4:], rois[:n, 1:], (10., 10., 5., 5.), (640, 480))
    want = g["bh_teacher_boxes0"].reshape(n, -1, 4)[:, 0]
    np.testing.assert_allclose(dec, want, atol=1e-3)


def test_scoring_oracle_matches_golden():
    g = _load("box_head.npz")
    logits = oracle.embed_logits(g["bh_A"], g["bh_E"])
    np.testing.assert_allclose(logits, g["bh_logits"], atol=1e-4)
    probs = oracle.softmax_rows(g["bh_logits"])
    want = g["bh_teacher_scores0"].reshape(300, -1)
    np.testing.assert_allclose(probs[:300], want, atol=1e-6)
    idx, mx, sig = oracle.caption_align(g["bh_A"][:300], g["cap_W"])
    assert np.array_equal(idx, g["cap_idx"])
    np.testing.assert_allclose(mx, g["cap_max"], atol=1e-4)
    np.testing.assert_allclose(sig, g["cap_sig"], atol=1e-6)


def test_nms_known_answers():
    b = np.array([[10, 10, 50, 50], [10, 10, 50, 50]], np.float32)
    assert oracle.nms(b, np.array([0.9, 0.8], np.float32), 1.0).tolist() == [0]       # `>=` (nms_cpu.cpp:60)
    b = np.array([[0, 0, 9, 9], [5, 0, 14, 9]], np.float32)                           # IoU 1/3 with +1 extents
    s = np.array([0.9, 0.8], np.float32)
    assert oracle.nms(b, s, 0.34).tolist() == [0, 1]
    assert oracle.nms(b, s, 0.33).tolist() == [0]
    assert oracle.nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.5).tolist() == []
    # ties: lower index first
    b = np.array([[0, 0, 9, 9], [0, 0, 9, 9], [100, 100, 120, 120]], np.float32)
    assert oracle.nms(b, np.array([0.5, 0.5, 0.5], np.float32), 0.5).tolist() == [0, 2]


def test_roi_align_known_answers():
    x = np.arange(72, dtype=np.float32).reshape(1, 2, 6, 6)
    out = oracle.roi_align_forward(x, np.array([[0, 3, 2, 4, 3]], np.float32), 1.0, 1, 1, 2)
    np.testing.assert_allclose(out[0, :, 0, 0], [2.5 * 6 + 3.5, 36 + 2.5 * 6 + 3.5])
    out = oracle.roi_align_forward(x, np.array([[0, 50, 50, 60, 60]], np.float32), 1.0, 2, 2, 2)   # outside
    assert np.all(out == 0)
    assert oracle.roi_align_forward(x, np.zeros((0, 5), np.float32), 1.0, 7, 7, 2).shape == (0, 2, 7, 7)


def test_roi_align_backward_is_adjoint_of_forward():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((1, 3, 12, 16)).astype(np.float32)
    rois = synth.make_rois(rng, 20, 1, 128, 96, smin=4, smax=100)
    g = rng.standard_normal((20, 3, 4, 4)).astype(np.float32)
    y = oracle.roi_align_forward(x, rois, 1 / 8, 4, 4, 2)
    _, gx = oracle.roi_align_backward(g, rois, 1 / 8, 4, 4, 1, 3, 12, 16, 2)
    assert abs((y.astype(np.float64) * g).sum() - (gx * x).sum()) < 1e-3   # <A x, g> == <x, A^T g>


def test_backward_oracle_matches_torchvision_autograd():
    tv = pytest.importorskip("torchvision")
    import torch
    rng = np.random.default_rng(4)
    x = torch.from_numpy(rng.standard_normal((2, 4, 20, 24)).astype(np.float32)).requires_grad_(True)
    rois = synth.make_rois(rng, 30, 2, 384, 320, smin=8, smax=300)
    g = rng.standard_normal((60, 4, 7, 7)).astype(np.float32)
    y = tv.ops.roi_align(x, torch.from_numpy(rois), (7, 7), 1 / 16, 2, aligned=False)
    y.backward(torch.from_numpy(g))
    _, want = oracle.roi_align_backward(g, rois, 1 / 16, 7, 7, 2, 4, 20, 24, 2)
    np.testing.assert_allclose(x.grad.numpy(), want, rtol=1e-4, atol=1e-5)


@pytest.mark.skipif(oracle.ref_lib() is None, reason="oracle/_ref not built on this machine")
def test_oracle_equals_compiled_reference_live():
    rng = np.random.default_rng(8)
    x = rng.standard_normal((2, 8, 50, 84)).astype(np.float32)
    rois = synth.make_rois(rng, 150, 2, smin=4, smax=1300)
    for (ph, pw, sr) in ((7, 7, 2), (14, 14, 0)):
        assert np.array_equal(oracle.roi_align_forward(x, rois, 1 / 16, ph, pw, sr),
                              oracle.ref_roi_align_forward(x, rois, 1 / 16, ph, pw, sr))
    b, s = synth.make_nms_boxes(rng, 2000)
    for thr in (0.3, 0.5, 0.7):
        assert np.array_equal(oracle.nms(b, s, thr), oracle.ref_nms(b, s, thr))


def test_level_map_matches_torch():
    import torch
    rng = np.random.default_rng(9)
    rois = synth.make_rois(rng, 50000, 1, 3000, 3000, smin=1, smax=2500, degenerate=0)
    lv = oracle.level_map(rois, 2.0, 5.0)
    b = torch.from_numpy(rois[:, 1:])
    area = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    t = torch.floor(4 + torch.log2(torch.sqrt(area) / 224 + 1e-6)).clamp(min=2.0, max=5.0).to(torch.int64) - 2
    assert int((t.numpy() != lv).sum()) == 0


def test_paste_masks_oracle_matches_reference_golden():
    """oracle.paste_masks == the reference's MaskPostProcessor + Masker outputs (tests/golden/masks.npz,
    generated by tests/golden/make_golden.py from /root/reference)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "masks.npz"))
    prob = 1.0 / (1.0 + np.exp(-g["mk_logits"].astype(np.float64)))
    o = 0
    for i in range(2):
        boxes, labels = g["mk_boxes%d" % i], g["mk_labels%d" % i]
        w, h = [int(v) for v in g["mk_size%d" % i]]
        n = len(boxes)
        want = np.unpackbits(g["mk_mask%d_packed" % i])[:n * h * w].reshape(n, h, w).astype(bool)
        # probabilities exactly as torch computes them are needed for bit parity at the threshold; the
        # fp64 sigmoid rounds to the same fp32 value except in rare last-ulp cases: allow 1e-5 of the pixels
        sel = prob[np.arange(o, o + n), labels].astype(np.float32)
        got = oracle.paste_masks(sel, boxes, h, w, 0.5, 1)
        assert (got != want).mean() <= 1e-5
        assert got[-1 if i == 0 else 0].any() == want[-1 if i == 0 else 0].any()
        o += n


def test_mask_targets_oracle_equals_the_reference_golden():
    """oracle.mask_targets against the output of the reference's own Masker + project_masks_on_boxes
    (tests/golden/make_golden.py:gen_mask_targets)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mask_targets.npz"))
    im_w, im_h, m = [int(v) for v in g["mt_size"]]
    n = len(g["mt_proposals"])
    want = np.unpackbits(g["mt_targets_packed"])[: n * m * m].reshape(n, m, m).astype(np.float32)
    got = oracle.mask_targets(g["mt_probs"], g["mt_label_boxes"], g["mt_match"], g["mt_proposals"], im_h, im_w, m)
    assert np.array_equal(got, want) and 0.05 < want.mean() < 0.95
