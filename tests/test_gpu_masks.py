"""GPU parity of b200_paste_masks / Masker / MaskPostProcessor against the reference's own
outputs (tests/golden/masks.npz) and the numpy restatement on random boxes.  The result is a boolean
image: a pixel can only differ where the interpolated probability sits within rounding of the
threshold, so the bar is "<= 1e-5 of the pixels differ"; on the golden case 0 differ."""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def test_mask_postprocessor_matches_reference_golden():
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import Masker, MaskPostProcessor
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "masks.npz"))
    bl = []
    for i in range(2):
        w, h = [int(v) for v in g["mk_size%d" % i]]
        t = BoxList(torch.from_numpy(g["mk_boxes%d" % i]).cuda(), (w, h), mode="xyxy")
        t.add_field("labels", torch.from_numpy(g["mk_labels%d" % i]).cuda())
        bl.append(t)
    res = MaskPostProcessor(masker=Masker(threshold=0.5, padding=1))(torch.from_numpy(g["mk_logits"]).cuda(), bl)
    for i, r in enumerate(res):
        w, h = [int(v) for v in g["mk_size%d" % i]]
        n = len(g["mk_boxes%d" % i])
        want = np.unpackbits(g["mk_mask%d_packed" % i])[:n * h * w].reshape(n, 1, h, w).astype(bool)
        got = r.get_field("mask")
        assert got.dtype == torch.bool and tuple(got.shape) == (n, 1, h, w)
        assert r.has_field("labels") and torch.equal(r.bbox.cpu(), torch.from_numpy(g["mk_boxes%d" % i]))
        assert int((got.cpu().numpy() != want).sum()) == 0
    # without a masker the field holds the selected probabilities
    res = MaskPostProcessor()(torch.from_numpy(g["mk_logits"]).cuda(), bl)
    prob = torch.from_numpy(g["mk_logits"]).sigmoid()
    n0 = len(g["mk_boxes0"])
    want0 = prob[torch.arange(n0), torch.from_numpy(g["mk_labels0"])][:, None]
    assert torch.allclose(res[0].get_field("mask").cpu(), want0, atol=1e-6)


@pytest.mark.parametrize("case", [(28, 1, 800, 1333, 40), (14, 1, 64, 48, 25), (7, 0, 333, 500, 12), (14, 2, 90, 1001, 9)])
def test_paste_masks_matches_oracle(case):
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import Masker
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    M, pad, im_h, im_w, n = case
    rng = np.random.default_rng(sum(case))
    masks = rng.random((n, 1, M, M)).astype(np.float32)
    x1 = rng.uniform(-0.2 * im_w, im_w, n); y1 = rng.uniform(-0.2 * im_h, im_h, n)
    boxes = np.stack([x1, y1, x1 + rng.uniform(0.2, 0.7 * im_w, n), y1 + rng.uniform(0.2, 0.7 * im_h, n)], 1).astype(np.float32)
    boxes[0] = [0, 0, im_w - 1, im_h - 1]
    boxes[1] = [im_w + 5, im_h + 5, im_w + 50, im_h + 60]            # fully outside: empty mask
    got = Masker(0.5, pad).forward_single_image(torch.from_numpy(masks).cuda(),
                                                BoxList(torch.from_numpy(boxes).cuda(), (im_w, im_h)))
    want = oracle.paste_masks(masks[:, 0], boxes, im_h, im_w, 0.5, pad)
    got = got[:, 0].cpu().numpy()
    assert got.shape == want.shape and not got[1].any()
    assert (got != want).mean() <= 1e-5, float((got != want).mean())
    assert want.sum() > 0


def test_masker_edge_cases():
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import Masker
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    e = Masker(0.5, 1).forward_single_image(torch.zeros((0, 1, 14, 14), device="cuda"),
                                            BoxList(torch.zeros((0, 4), device="cuda"), (64, 48)))
    assert e.shape == (0, 1, 14, 14)
    with pytest.raises(RuntimeError):
        Masker(0.5, 1).forward_single_image(torch.zeros((1, 1, 14, 14)), BoxList(torch.zeros((1, 4)), (64, 48)))
    with pytest.raises(ValueError):
        Masker(-1.0, 1)


def _boxlist(b, size):
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    return BoxList(torch.from_numpy(np.asarray(b, np.float32)).cuda(), size, mode="xyxy")


def test_mask_targets_equal_the_reference_golden():
    """b200_mask_targets (paste o crop) == the reference's Masker + project_masks_on_boxes, 0 pixels off."""
    import os
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling.roi_heads.mask_head.loss import project_masks_on_boxes
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mask_targets.npz"))
    im_w, im_h, m = [int(v) for v in g["mt_size"]]
    n = len(g["mt_proposals"])
    want = np.unpackbits(g["mt_targets_packed"])[: n * m * m].reshape(n, m, m).astype(np.float32)
    got = project_masks_on_boxes(torch.from_numpy(g["mt_probs"]).cuda(), _boxlist(g["mt_label_boxes"], (im_w, im_h)),
                                 torch.from_numpy(g["mt_match"]), _boxlist(g["mt_proposals"], (im_w, im_h)), m)
    assert got.dtype == torch.float32 and got.is_cuda
    assert np.array_equal(got.cpu().numpy(), want)


def test_mask_targets_random_boxes_vs_oracle_and_vs_paste_then_crop():
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling.roi_heads.mask_head.loss import project_masks_on_boxes
    rng = np.random.default_rng(31)
    im_w, im_h, k, p, m = 640, 427, 12, 300, 28
    probs = 1 / (1 + np.exp(-(rng.standard_normal((k, 14, 14)) * 2))).astype(np.float32)
    x1, y1 = rng.uniform(-20, im_w - 30, k), rng.uniform(-20, im_h - 30, k)
    lb = np.stack([x1, y1, x1 + rng.uniform(2, 300, k), y1 + rng.uniform(2, 250, k)], 1).astype(np.float32)
    match = rng.integers(-1, k, p).astype(np.int32)
    props = (lb[np.maximum(match, 0)] + rng.normal(0, 10, (p, 4))).astype(np.float32)
    props[:, 2:] = np.maximum(props[:, 2:], props[:, :2])
    props[::7] = np.round(props[::7]) + 0.5                    # half-integer corners: round-half-to-even
    want = oracle.mask_targets(probs, lb, match, props, im_h, im_w, m)
    got = project_masks_on_boxes(torch.from_numpy(probs).cuda(), _boxlist(lb, (im_w, im_h)), torch.from_numpy(match),
                                 _boxlist(props, (im_w, im_h)), m).cpu().numpy()
    # the kernel's interpolated mask value may sit within rounding of the threshold where numpy's does not
    assert (got != want).mean() <= 1e-4
    assert np.all(got[match < 0] == 0)
    assert project_masks_on_boxes(torch.zeros((0, 14, 14)).cuda(), _boxlist(np.zeros((0, 4)), (im_w, im_h)), torch.zeros((0,)),
                                  _boxlist(np.zeros((0, 4)), (im_w, im_h)), m).numel() == 0
