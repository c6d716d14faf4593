"""GPU parity of b200_rpn_candidates (the RPN post-processor's fused front end) against the
torch-op formulation of the same steps (RPNPostProcessor._decode_level: permute_and_flatten, sigmoid,
topk, gather, BoxCoder.decode, clip) and, end to end, of the two RPNPostProcessor paths."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _canon(boxes, scores):
    """Rows ordered by (score desc, box): two DIFFERENT logits can round to the same sigmoid, and
    the order inside such a score tie is unspecified for torch.topk (the kernel orders by logit)."""
    b, s = boxes.cpu().numpy().astype(np.float64), scores.cpu().numpy().astype(np.float64)
    order = np.lexsort((b[:, 3], b[:, 2], b[:, 1], b[:, 0], -s))
    return b[order], s[order]


def _levels(seed, n_img, shapes, A=3, quant=None, shared=True):
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    g = torch.Generator(device="cuda").manual_seed(seed)
    obj, reg, anchors = [], [], [[] for _ in range(n_img)]
    sizes = [(1333, 800), (1200, 750), (640, 480)][:n_img] + [(1000, 600)] * max(0, n_img - 3)
    for li, (h, w) in enumerate(shapes):
        o = torch.randn((n_img, A, h, w), device="cuda", generator=g) * 2.0
        if quant:
            o = torch.round(o * quant) / quant              # many exactly equal logits
        obj.append(o)
        reg.append(torch.randn((n_img, 4 * A, h, w), device="cuda", generator=g) * 0.3)
        stride = 4 * 2 ** li
        ys, xs = torch.meshgrid(torch.arange(h, device="cuda"), torch.arange(w, device="cuda"), indexing="ij")
        ctr = torch.stack([xs, ys, xs, ys], -1).reshape(-1, 1, 4).float() * stride
        half = torch.tensor([[-1, -0.5, 1, 0.5], [-0.75, -0.75, 0.75, 0.75], [-0.5, -1, 0.5, 1]], device="cuda") * 8 * stride
        base = (ctr + half[None]).reshape(-1, 4).contiguous()          # (h, w, a) order
        for i in range(n_img):
            anchors[i].append(BoxList(base if shared else base.clone(), sizes[i], mode="xyxy"))
    return anchors, obj, reg


@pytest.fixture(params=["sampled", "exact"])
def select_path(request):
    """The top-k selection runs with the sampled lower bound (one pass) and with the exact bisection fallback."""
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    _ext.debug_rpn(request.param == "exact")
    yield request.param
    _ext.debug_rpn(False)


@pytest.mark.parametrize("shared", [True, False])
@pytest.mark.parametrize("pre_nms", [300, 6000, 12000])
def test_rpn_front_matches_torch_path(shared, pre_nms, select_path):
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import RPNPostProcessor
    anchors, obj, reg = _levels(1, 2, [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)], shared=shared)
    pp = RPNPostProcessor(pre_nms, 1000, 0.7, 0)
    per_level = list(zip(*anchors))
    sizes = [a[0].size for a in anchors]
    assert pp._front_fused_ok(obj, reg)
    boxes, score, ks = pp._decode_all_fused(per_level, obj, reg, sizes)
    K = sum(ks)
    boxes, score = boxes.view(2, K, 4), score.view(2, K)
    o0 = 0
    for a, o, b, k in zip(per_level, obj, reg, ks):
        p, s = pp._decode_level(a, o, b)
        assert p.shape[1] == k
        assert torch.allclose(score[:, o0:o0 + k], s, atol=1e-6, rtol=0)
        assert bool((score[:, o0:o0 + k][:, 1:] <= score[:, o0:o0 + k][:, :-1]).all())     # descending
        for n in range(2):
            gb, gs = _canon(boxes[n, o0:o0 + k], score[n, o0:o0 + k])
            wb, ws = _canon(p[n], s[n])
            np.testing.assert_allclose(gb, wb, atol=1e-3, rtol=0)
            # outside score ties the positions agree exactly
            uniq = torch.ones(k, dtype=torch.bool, device="cuda")
            uniq[1:] &= s[n, 1:] != s[n, :-1]
            uniq[:-1] &= s[n, :-1] != s[n, 1:]
            assert torch.allclose(boxes[n, o0:o0 + k][uniq], p[n][uniq], atol=1e-3, rtol=0)
        o0 += k


@pytest.mark.parametrize("case", [(40, 60, 500, 4), (120, 160, 3000, 2), (120, 160, 700, 0)])
def test_rpn_front_ties_at_the_threshold(case, select_path):
    """Quantised logits: hundreds (or, with quant 0, ALL) of equal values straddle the k-th place.
    Rule of b200_rpn_candidates: a stable descending sort of the flattened logits cut at k, i.e.
    order (logit descending, flattened anchor index (h*W + w)*A + a ascending), also at the cut.
    torch.topk leaves tie order unspecified, so the expectation is built explicitly."""
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import RPNPostProcessor
    H, W, k, quant = case
    A = 3
    anchors, obj, reg = _levels(2, 1, [(H, W)], quant=quant if quant else None)
    if quant == 0:
        obj = [torch.zeros_like(o) for o in obj]           # an untrained head: every logit equal
    pp = RPNPostProcessor(k, 1000, 0.7, 0)
    per_level = list(zip(*anchors))
    boxes, score, ks = pp._decode_all_fused(per_level, obj, reg, [anchors[0][0].size])
    assert ks == [k]
    _, s = pp._decode_level(per_level[0], obj[0], reg[0])
    assert torch.allclose(score, s.view(-1), atol=1e-6, rtol=0)             # same score multiset, sorted
    logit_flat = obj[0].permute(0, 2, 3, 1).reshape(-1).cpu().numpy().astype(np.float64)   # (h, w, a) order
    order = np.lexsort((np.arange(A * H * W), -logit_flat))[:k]               # stable descending, cut at k
    thr = logit_flat[order[-1]]
    assert (logit_flat == thr).sum() > (logit_flat[order] == thr).sum() > 0   # the tie really straddles the cut
    flat = torch.from_numpy(order).cuda()
    all_boxes = pp.box_coder.decode(reg[0].view(1, A, 4, H, W).permute(0, 3, 4, 1, 2).reshape(-1, 4),
                                    per_level[0][0].bbox)
    w, h = anchors[0][0].size
    lim = torch.tensor([w - 1, h - 1, w - 1, h - 1], device="cuda", dtype=torch.float32)
    want_boxes = torch.min(all_boxes.clamp(min=0), lim)[flat]
    assert torch.allclose(boxes, want_boxes, atol=1e-3, rtol=0)


@pytest.mark.parametrize("min_size", [0, 12])
def test_rpn_postprocessor_fused_front_equals_torch_front(min_size):
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import RPNPostProcessor
    anchors, obj, reg = _levels(3, 3, [(50, 84), (25, 42), (13, 21)])
    pp = RPNPostProcessor(600, 200, 0.7, min_size, fpn_post_nms_top_n=300).eval()
    fused = pp(anchors, obj, reg)
    pp._front_fused_ok = lambda *a, **k: False
    plain = pp(anchors, obj, reg)
    for f, p in zip(fused, plain):
        assert len(f) == len(p) > 0
        assert torch.allclose(f.get_field("objectness"), p.get_field("objectness"), atol=1e-6, rtol=0)
        gb, _ = _canon(f.bbox, f.get_field("objectness"))
        wb, _ = _canon(p.bbox, p.get_field("objectness"))
        np.testing.assert_allclose(gb, wb, atol=1e-3, rtol=0)


@pytest.mark.parametrize("training", [False, True])
@pytest.mark.parametrize("post_nms", [150, 2000])
def test_rpn_postprocessor_single_level_selection_is_fused(training, post_nms):
    """The reference's shipped configs are C4: ONE feature map, no cross-level top-k (rpn/inference.py:148-150).  The
    batched NMS + b200_select_topk path must return what the torch-op formulation returns: each image's kept boxes
    in score order, at most post_nms_top_n of them; in training the ground-truth boxes are appended (:53-74)."""
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import RPNPostProcessor
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    anchors, obj, reg = _levels(2, 3, [(50, 84)])
    pp = RPNPostProcessor(3000, post_nms, 0.7, 0)
    pp.train(training)
    targets = None
    if training:
        targets = [BoxList(torch.tensor([[10., 20., 200., 300.], [50., 60., 400., 450.]], device="cuda"), a[0].size)
                   for a in anchors]
    fused = pp(anchors, obj, reg, targets)
    pp.fused_select = False
    plain = pp(anchors, obj, reg, targets)
    for f, p in zip(fused, plain):
        assert len(f) == len(p) > 0
        assert len(f) <= post_nms + (2 if training else 0)
        assert torch.equal(f.get_field("objectness"), p.get_field("objectness"))
        assert torch.equal(f.bbox, p.bbox)
        s = f.get_field("objectness")[: len(f) - (2 if training else 0)]
        assert bool((s[:-1] >= s[1:]).all())
