"""Full-size checks (BASELINE.json configs): config #1 (the reference's CPU-runnable C4 chain)
directly against the oracle, config #2 through size-independent properties."""
import numpy as np
import pytest
import torch

import oracle
from tests import synth

pytestmark = pytest.mark.gpu


def test_config1_c4_teacher_chain_matches_oracle():
    """configs[0]: one 800x1333 image, R-50-C4 shapes: RPN NMS on 6000 candidates (thr 0.7, keep
    1000) -> Pooler 14x14, sampling_ratio 0 (adaptive), 1024 channels, stride 16 -> scoring."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import embed_match_softmax
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import Pooler
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList, boxlist_nms
    rng = np.random.default_rng(1235)
    boxes, scores = synth.make_nms_boxes(rng, 6000)
    order = np.argsort(-scores, kind="stable")
    boxes, scores = boxes[order], scores[order]
    bl = BoxList(torch.from_numpy(boxes).cuda(), (1333, 800))
    bl.add_field("objectness", torch.from_numpy(scores).cuda())
    props = boxlist_nms(bl, 0.7, max_proposals=1000, score_field="objectness")
    want_keep = oracle.nms(boxes, scores, 0.7)[:1000]
    assert np.array_equal(props.bbox.cpu().numpy(), boxes[want_keep])
    feat = torch.from_numpy(rng.standard_normal((1, 1024, 50, 84)).astype(np.float32)).cuda()
    pooled = Pooler((14, 14), (1.0 / 16,), 0)([feat], [props])
    rois = np.concatenate([np.zeros((len(want_keep), 1), np.float32), boxes[want_keep]], 1)
    want = oracle.roi_align_forward(feat.cpu().numpy(), rois, 1.0 / 16, 14, 14, 0)
    assert np.array_equal(pooled.cpu().numpy(), want)
    # scoring on the pooled features (mean over bins, first 768 channels as a stand-in embedding)
    emb = pooled.mean(dim=(2, 3))[:, :768].to(torch.bfloat16)
    g = torch.Generator().manual_seed(1)
    E = torch.nn.functional.normalize(torch.randn((66, 768), generator=g), dim=-1)
    E[0] = 0
    E = E.to(torch.bfloat16)
    out = embed_match_softmax(emb, E.cuda(), 0.05)
    probs, top, _, _ = oracle.embed_match_softmax(emb.float().cpu().numpy(), E.float().numpy(), 0.05)
    assert np.abs(out["probs"].cpu().numpy() - probs).max() < 2e-2


@pytest.fixture(scope="module")
def config2():
    rng = np.random.default_rng(1236)
    g = torch.Generator(device="cuda").manual_seed(1236)
    feats = [torch.randn((16, 256, h, w), device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
             for (h, w) in synth.fpn_shapes()]
    rois = torch.from_numpy(synth.make_rois(rng, 1000, 16)).cuda()
    return feats, rois


def test_config2_pooler_properties(config2):
    """16 images x 1000 RoIs x 256 ch x 4 levels: (i) a random subset equals the oracle bit for
    bit, (ii) constant features pool to the constant (weights of a sample sum to 1) for RoIs
    strictly inside the image, (iii) linearity in the features."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward
    feats, rois = config2
    out, lv = _forward(feats, synth.FPN_SCALES, rois, (7, 7), 2, want_levels=True)
    assert out.shape == (16000, 256, 7, 7) and torch.isfinite(out).all()
    lv = lv.cpu().numpy()
    want_lv = oracle.level_map(rois.cpu().numpy(), 2.0, 5.0)
    assert int((lv != want_lv).sum()) == 0 and len(np.unique(lv)) == 4
    rng = np.random.default_rng(0)
    pick = rng.choice(16000, 40, replace=False)
    r = rois[pick].cpu().numpy()
    for l in range(4):
        sel = np.nonzero(want_lv[pick] == l)[0]
        if len(sel) == 0:
            continue
        f = feats[l].cpu().contiguous().numpy()
        want = oracle.roi_align_forward(f, r[sel], synth.FPN_SCALES[l], 7, 7, 2)
        assert np.array_equal(out[pick[sel]].cpu().numpy(), want), l
    # (ii) constant maps
    ones = [torch.full_like(f, 3.0) for f in feats]
    o1, _ = _forward(ones, synth.FPN_SCALES, rois, (7, 7), 2)
    inside = (rois[:, 1] > 40) & (rois[:, 2] > 40) & (rois[:, 3] < 1200) & (rois[:, 4] < 700)
    assert inside.sum() > 1000
    assert torch.allclose(o1[inside], torch.full_like(o1[inside], 3.0), rtol=1e-6, atol=0)
    # (iii) linearity: pool(2*x + y) == 2*pool(x) + pool(y) up to fp32 rounding
    mix = [2.0 * f + o for f, o in zip(feats, ones)]
    o2, _ = _forward(mix, synth.FPN_SCALES, rois, (7, 7), 2)
    assert torch.allclose(o2, 2.0 * out + o1, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("res", [7, 14])
def test_config2_fast_math_within_tolerance(config2, res):
    """Full size, fast math against the bit-exact kernel (itself pinned to the oracle above):
    every element within rtol 1e-5 (+ atol 1e-6 on unit-variance features); constants pool to
    the constant; same level assignment."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward
    feats, rois = config2
    sub = rois if res == 7 else rois[::4]
    exact, lv_e = _forward(feats, synth.FPN_SCALES, sub, (res, res), 2, want_levels=True, math="exact")
    fast, lv_f = _forward(feats, synth.FPN_SCALES, sub, (res, res), 2, want_levels=True, math="fast")
    assert torch.equal(lv_e, lv_f)
    err = (fast - exact).abs()
    bound = 1e-5 * exact.abs() + 1e-6
    assert bool((err <= bound).all()), float((err - bound).max())
    del exact, fast, err, bound
    ones = [torch.full_like(f, 3.0) for f in feats]
    o1, _ = _forward(ones, synth.FPN_SCALES, sub, (res, res), 2, math="fast")
    inside = (sub[:, 1] > 40) & (sub[:, 2] > 40) & (sub[:, 3] < 1200) & (sub[:, 4] < 700)
    assert torch.allclose(o1[inside], torch.full_like(o1[inside], 3.0), rtol=1e-6, atol=0)


def test_config2_mask_pooler_and_backward_adjoint(config2):
    """<Pool(x), g> == <x, Pool^T(g)> at full size (the backward is the adjoint of the forward)."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _backward, _forward
    feats, rois = config2
    sub = rois[::8]                                     # 2000 RoIs keep the fp32 dot products tame
    for res in (7, 14):
        out, _ = _forward(feats, synth.FPN_SCALES, sub, (res, res), 2)
        g = torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(res))
        grads = _backward(g, sub, [tuple(f.shape) for f in feats], True, synth.FPN_SCALES, (res, res), 2)
        lhs = (out.double() * g.double()).sum()
        rhs = sum((f.double() * gr.double()).sum() for f, gr in zip(feats, grads))
        scale = float((out.double().abs() * g.double().abs()).sum())
        assert abs(float(lhs - rhs)) <= 1e-6 * scale, (res, float(lhs), float(rhs), scale)


def test_config2_rpn_nms_properties():
    """80 segments of up to 6000 boxes with keep <= 1000: counts, ascending order, kept boxes
    mutually below the threshold, and every kept prefix equals the oracle on sampled segments."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms_batched
    rng = np.random.default_rng(2)
    lens = [6000, 6000, 6000, 3150, 819] * 16
    off = np.concatenate([[0], np.cumsum(lens)])
    bs, ss = [], []
    for L in lens:
        b, s = synth.make_nms_boxes(rng, L)
        o = np.argsort(-s, kind="stable")
        bs.append(b[o]); ss.append(s[o])
    boxes, scores = np.concatenate(bs), np.concatenate(ss)
    ki, kc = nms_batched(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(),
                         torch.from_numpy(off.astype(np.int32)).cuda(), 0.7, 1000, 6000)
    ki, kc = ki.cpu().numpy(), kc.cpu().numpy()
    assert kc.max() <= 1000 and kc.min() > 0
    for s in (0, 2, 3, 4, 77, 79):
        a = off[s]
        want = oracle.nms(boxes[a:off[s + 1]], scores[a:off[s + 1]], 0.7)[:1000]
        assert kc[s] == len(want) and np.array_equal(ki[a:a + kc[s]], want), s
        assert np.all(ki[a + kc[s]:off[s + 1]] == -1)
