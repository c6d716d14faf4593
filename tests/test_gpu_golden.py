"""GPU parity against the committed golden fixtures, i.e. against outputs of the REFERENCE's
own Python classes driving the reference's own CPU kernels (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from tests import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("nhwc", [False, True])
def test_roi_align_golden(nhwc):
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import ROIAlign
    g = _load("roi_align.npz")
    for i in range(3):
        scale, ph, pw, sr = g["ra%d_cfg" % i]
        x = _cu(g["ra%d_x" % i])
        if nhwc:
            x = x.contiguous(memory_format=torch.channels_last)
        out = ROIAlign((int(ph), int(pw)), float(scale), int(sr))(x, _cu(g["ra%d_rois" % i]))
        assert np.array_equal(out.cpu().numpy(), g["ra%d_out" % i]), i


def test_nms_golden():
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms
    g = _load("nms.npz")
    for i in range(4):
        keep = nms(_cu(g["nms%d_boxes" % i]), _cu(g["nms%d_scores" % i]), float(g["nms%d_thr" % i][0]))
        assert np.array_equal(keep.cpu().numpy(), g["nms%d_keep" % i]), i


@pytest.mark.parametrize("nhwc", [False, True])
def test_pooler_golden(nhwc):
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import Pooler
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    g = _load("pooler.npz")
    feats = [_cu(g["pool_f%d" % l]) for l in range(4)]
    if nhwc:
        feats = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    rois = g["pool_rois"]
    n = rois.shape[0] // 2
    boxes = [BoxList(_cu(rois[i * n:(i + 1) * n, 1:]), (416, 256)) for i in range(2)]
    for res in (7, 14):
        out = Pooler((res, res), synth.FPN_SCALES, 2)(feats, boxes)
        assert np.array_equal(out.cpu().numpy(), g["pool_out%d" % res]), res


def _match_boxes(got_b, got_s, want_b, want_s, tol=2e-3):
    """Same detections up to order among exactly-equal scores and 1-ulp exp() differences."""
    assert got_b.shape == want_b.shape, (got_b.shape, want_b.shape)
    og, ow = np.lexsort((got_b[:, 0], -got_s)), np.lexsort((want_b[:, 0], -want_s))
    np.testing.assert_allclose(got_s[og], want_s[ow], atol=1e-6)
    np.testing.assert_allclose(got_b[og], want_b[ow], atol=tol)


def test_rpn_postprocessor_golden():
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import BoxCoder, RPNPostProcessor
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    g = _load("rpn.npz")
    w, h = [int(v) for v in g["rpn_imsize"]]
    anchors = [[BoxList(_cu(g["rpn_anchors_%d_%d" % (i, l)]), (w, h)) for l in range(3)] for i in range(2)]
    obj = [_cu(g["rpn_obj%d" % l]) for l in range(3)]
    reg = [_cu(g["rpn_reg%d" % l]) for l in range(3)]
    pp = RPNPostProcessor(600, 100, 0.7, 0, BoxCoder((1., 1., 1., 1.)), fpn_post_nms_top_n=150).eval()
    res = pp(anchors, obj, reg)
    for i, bl in enumerate(res):
        _match_boxes(bl.bbox.cpu().numpy(), bl.get_field("objectness").cpu().numpy(),
                     g["rpn_test_boxes%d" % i], g["rpn_test_obj%d" % i])
    pp1 = RPNPostProcessor(500, 80, 0.7, 0, BoxCoder((1., 1., 1., 1.))).eval()
    res = pp1([[a[1]] for a in anchors], obj[1:2], reg[1:2])
    for i, bl in enumerate(res):
        # single level: order is the NMS keep order (descending score), compare positionally
        np.testing.assert_allclose(bl.get_field("objectness").cpu().numpy(), g["rpn_single_obj%d" % i], atol=1e-6)
        np.testing.assert_allclose(bl.bbox.cpu().numpy(), g["rpn_single_boxes%d" % i], atol=2e-3)


def test_box_postprocessor_golden():
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import BoxCoder, PostProcessor
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    g = _load("box_head.npz")
    rois = g["bh_rois"]
    n = rois.shape[0] // 2
    boxes = [BoxList(_cu(rois[i * n:(i + 1) * n, 1:]), (640, 480)) for i in range(2)]
    pp = PostProcessor(0.05, 0.5, 100, BoxCoder((10., 10., 5., 5.)), cls_agnostic_bbox_reg=True)
    res = pp((_cu(g["bh_logits"]), _cu(g["bh_reg"])), boxes)
    for i, bl in enumerate(res):
        # same order as the reference: class-major, RoI ascending
        assert np.array_equal(bl.get_field("labels").cpu().numpy(), g["bh_labels%d" % i])
        np.testing.assert_allclose(bl.get_field("scores").cpu().numpy(), g["bh_scores%d" % i], atol=1e-6)
        np.testing.assert_allclose(bl.bbox.cpu().numpy(), g["bh_boxes%d" % i], atol=2e-3)
    boxes = [BoxList(_cu(rois[i * n:(i + 1) * n, 1:]), (640, 480)) for i in range(2)]
    tp = PostProcessor(0.05, 0.5, 100, BoxCoder((10., 10., 5., 5.)), cls_agnostic_bbox_reg=True, is_teacher=True)
    res = tp((_cu(g["bh_logits"]), _cu(g["bh_reg"])), boxes)
    np.testing.assert_allclose(res[0].bbox.cpu().numpy(), g["bh_teacher_boxes0"], atol=2e-3)
    np.testing.assert_allclose(res[0].get_field("scores").cpu().numpy(), g["bh_teacher_scores0"], atol=1e-6)


def test_predictor_and_pseudo_labels_golden():
    """Embedding predictor (tcgen05 scoring) + PostProcessor + caption alignment vs the
    reference chain on bf16-representable inputs: scores <= 2e-2, identical detections' labels."""
    from types import SimpleNamespace as NS
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import caption_align, embed_match_softmax
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import FastRCNNPredictor
    g = _load("box_head.npz")
    A, E = _cu(g["bh_A"]), _cu(g["bh_E"])
    out = embed_match_softmax(A, E, 0.05, want_probs=True, want_logits=True)
    np.testing.assert_allclose(out["logits"].cpu().numpy(), g["bh_logits"], atol=2e-3 * np.abs(g["bh_logits"]).max())
    want_p = g["bh_teacher_scores0"].reshape(300, -1)
    np.testing.assert_allclose(out["probs"][:300].cpu().numpy(), want_p, atol=2e-2)
    assert (out["probs"][:300, 1:].argmax(1).cpu().numpy() == want_p[:, 1:].argmax(1)).mean() >= 0.999
    idx, mx, sig = caption_align(A[:300], [300], [_cu(g["cap_W"])])[0]
    assert np.array_equal(idx.cpu().numpy(), g["cap_idx"])
    np.testing.assert_allclose(sig.cpu().numpy(), g["cap_sig"], atol=2e-2)
    # the module wrapper: identity projection so cls_emb == A
    cfg = NS(MODEL=NS(ROI_BOX_HEAD=NS(EMBEDDING_BASED=True, EMB_DIM=64, FREEZE_EMB_PRED=False),
                      CLS_AGNOSTIC_BBOX_REG=True, ROI_HEADS=NS(SCORE_THRESH=0.05)))
    pred = FastRCNNPredictor(cfg, 64).cuda().eval()
    with torch.no_grad():
        pred.emb_pred.weight.copy_(torch.eye(64))
        pred.emb_pred.bias.zero_()
    pred.set_class_embeddings(E)
    with torch.no_grad():
        logit, reg = pred(A.view(-1, 64, 1, 1))
    assert reg.shape == (600, 8) and hasattr(logit, "b200_probs")
    np.testing.assert_allclose(logit.cpu().numpy(), g["bh_logits"], atol=2e-3 * np.abs(g["bh_logits"]).max())
