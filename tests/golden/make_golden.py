"""Generate the committed golden fixtures from the REFERENCE itself (run in the build
container, where /root/reference exists; the GPU box only reads the .npz files).

The reference's own CPU kernels (oracle/_ref, compiled in place from
/root/reference/maskrcnn_benchmark/csrc/cpu) are plugged in as `maskrcnn_benchmark._C`, and
the reference's own Python classes (Pooler, RPNPostProcessor, PostProcessor,
FastRCNNPredictor) are imported from /root/reference and driven on seeded synthetic inputs.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from tests import synth  # noqa: E402

REF = "/root/reference"


def install_reference():
    """apex.amp / maskrcnn_benchmark._C stubs so the reference's Python imports on CPU."""
    np.float = float  # rpn/anchor_generator.py uses the removed alias
    amp = types.ModuleType("apex.amp")
    amp.float_function = lambda f: f
    apex = types.ModuleType("apex")
    apex.amp = amp
    sys.modules["apex"], sys.modules["apex.amp"] = apex, amp
    sys.path.insert(0, REF)
    import maskrcnn_benchmark
    c = types.ModuleType("maskrcnn_benchmark._C")

    def nms(dets, scores, thr):
        return torch.from_numpy(oracle.ref_nms(dets.numpy(), scores.numpy(), float(thr)))

    def roi_align_forward(inp, rois, scale, ph, pw, sr):
        return torch.from_numpy(oracle.ref_roi_align_forward(inp.contiguous().numpy(), rois.contiguous().numpy(),
                                                             float(scale), int(ph), int(pw), int(sr)))
    c.nms = nms
    c.roi_align_forward = roi_align_forward
    sys.modules["maskrcnn_benchmark._C"] = c
    maskrcnn_benchmark._C = c


def gen_roi_align(rng):
    out = {}
    cases = [(2, 16, 25, 42, 1 / 32, 7, 7, 2, 40), (1, 8, 50, 84, 1 / 16, 14, 14, 0, 30), (1, 8, 30, 30, 1 / 8, 3, 5, 3, 20)]
    for i, (b, c, h, w, scale, ph, pw, sr, n) in enumerate(cases):
        x = rng.standard_normal((b, c, h, w)).astype(np.float32)
        rois = synth.make_rois(rng, n, b, int(w / scale), int(h / scale), smin=4, smax=float(max(w, h) / scale))
        rois[0, 3] = rois[0, 1] - 3          # malformed
        rois[1, 1:] += 9000                  # outside
        out["ra%d_x" % i] = x
        out["ra%d_rois" % i] = rois
        out["ra%d_cfg" % i] = np.array([scale, ph, pw, sr], np.float64)
        out["ra%d_out" % i] = oracle.ref_roi_align_forward(x, rois, scale, ph, pw, sr)
    return out


def gen_nms(rng):
    out = {}
    for i, (n, thr) in enumerate([(1, 0.5), (65, 0.5), (700, 0.7), (3000, 0.7)]):
        b, s = synth.make_nms_boxes(rng, n)
        out["nms%d_boxes" % i], out["nms%d_scores" % i] = b, s
        out["nms%d_thr" % i] = np.array([thr])
        out["nms%d_keep" % i] = oracle.ref_nms(b, s, thr)
    return out


def gen_pooler(rng):
    from maskrcnn_benchmark.modeling.poolers import Pooler
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    b, c, n = 2, 8, 60
    shapes = synth.fpn_shapes(256, 416)
    feats = [rng.standard_normal((b, c, h, w)).astype(np.float32) for (h, w) in shapes]
    rois = synth.make_rois(rng, n, b, 416, 256, smin=8, smax=300)
    boxes = [BoxList(torch.from_numpy(rois[i * n:(i + 1) * n, 1:].copy()), (416, 256)) for i in range(b)]
    out = {}
    for res in (7, 14):
        y = Pooler((res, res), synth.FPN_SCALES, 2)([torch.from_numpy(f) for f in feats], boxes)
        out["pool_out%d" % res] = y.numpy()
    for l, f in enumerate(feats):
        out["pool_f%d" % l] = f
    out["pool_rois"] = rois
    return out


def gen_rpn(rng):
    from maskrcnn_benchmark.modeling.box_coder import BoxCoder
    from maskrcnn_benchmark.modeling.rpn.anchor_generator import AnchorGenerator
    from maskrcnn_benchmark.modeling.rpn.inference import RPNPostProcessor
    from maskrcnn_benchmark.structures.image_list import ImageList
    n_img, img_h, img_w = 2, 256, 416
    strides = (4, 8, 16)
    sizes = (32, 64, 128)
    ag = AnchorGenerator(sizes=sizes, aspect_ratios=(0.5, 1.0, 2.0), anchor_strides=strides, straddle_thresh=0)
    feats = [torch.zeros(n_img, 1, img_h // s, img_w // s) for s in strides]
    images = ImageList(torch.zeros(n_img, 3, img_h, img_w), [(img_h, img_w)] * n_img)
    anchors = ag(images, feats)
    # objectness with unique values per (image, level) so top-k has no ties
    obj, reg = [], []
    for f in feats:
        h, w = f.shape[-2:]
        a = 3
        vals = rng.permutation(n_img * a * h * w).astype(np.float32) / (n_img * a * h * w) * 8 - 4
        obj.append(torch.from_numpy(vals.reshape(n_img, a, h, w)))
        reg.append(torch.from_numpy((rng.standard_normal((n_img, 4 * a, h, w)) * 0.3).astype(np.float32)))
    out = {}
    for tag, kw in (("test", dict(pre_nms_top_n=600, post_nms_top_n=100, fpn_post_nms_top_n=150)),
                    ("single", dict(pre_nms_top_n=500, post_nms_top_n=80))):
        pp = RPNPostProcessor(nms_thresh=0.7, min_size=0, box_coder=BoxCoder((1., 1., 1., 1.)), **kw)
        pp.eval()
        if tag == "single":
            res = pp([[a[1]] for a in anchors], obj[1:2], reg[1:2])
        else:
            res = pp(anchors, obj, reg)
        for i, bl in enumerate(res):
            out["rpn_%s_boxes%d" % (tag, i)] = bl.bbox.numpy()
            out["rpn_%s_obj%d" % (tag, i)] = bl.get_field("objectness").numpy()
    for l in range(len(strides)):
        out["rpn_obj%d" % l] = obj[l].numpy()
        out["rpn_reg%d" % l] = reg[l].numpy()
        for i in range(n_img):
            out["rpn_anchors_%d_%d" % (i, l)] = anchors[i][l].bbox.numpy()
    out["rpn_imsize"] = np.array([img_w, img_h])
    return out


def gen_box_head(rng):
    from maskrcnn_benchmark.modeling.box_coder import BoxCoder
    from maskrcnn_benchmark.modeling.roi_heads.box_head.inference import PostProcessor
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    n_img, r, c, d = 2, 300, 20, 64
    g = torch.Generator().manual_seed(77)
    A = (torch.randn((n_img * r, d), generator=g) * 4).to(torch.bfloat16).float()
    E = torch.nn.functional.normalize(torch.randn((c, d), generator=g), dim=-1)
    E[0] = 0
    E = E.to(torch.bfloat16).float()
    logits = torch.einsum("pe,ce->pc", A, E)              # roi_box_predictors.py:67
    reg = torch.randn((n_img * r, 8), generator=g) * 0.5
    rois = synth.make_rois(rng, r, n_img, 640, 480, smin=16, smax=300)
    boxes = [BoxList(torch.from_numpy(rois[i * r:(i + 1) * r, 1:].copy()), (640, 480)) for i in range(n_img)]
    out = dict(bh_A=A.numpy(), bh_E=E.numpy(), bh_logits=logits.numpy(), bh_reg=reg.numpy(), bh_rois=rois)
    pp = PostProcessor(0.05, 0.5, 100, BoxCoder((10., 10., 5., 5.)), cls_agnostic_bbox_reg=True)
    res = pp((logits, reg), boxes)
    for i, bl in enumerate(res):
        out["bh_boxes%d" % i] = bl.bbox.numpy()
        out["bh_scores%d" % i] = bl.get_field("scores").numpy()
        out["bh_labels%d" % i] = bl.get_field("labels").numpy()
    boxes = [BoxList(torch.from_numpy(rois[i * r:(i + 1) * r, 1:].copy()), (640, 480)) for i in range(n_img)]
    tp = PostProcessor(0.05, 0.5, 100, BoxCoder((10., 10., 5., 5.)), cls_agnostic_bbox_reg=True, is_teacher=True)
    res = tp((logits, reg), boxes)
    out["bh_teacher_boxes0"] = res[0].bbox.numpy()
    out["bh_teacher_scores0"] = res[0].get_field("scores").numpy()
    # caption alignment, st_generalized_rcnn.py:245-255 (plain torch ops there)
    W = torch.nn.functional.normalize(torch.randn((5, d), generator=g), dim=-1).to(torch.bfloat16).float()
    rs = torch.einsum("pd,wd->pw", A[:r], W)
    mx, idx = torch.max(rs, dim=0)
    out.update(cap_W=W.numpy(), cap_idx=idx.numpy(), cap_max=mx.numpy(), cap_sig=torch.sigmoid(mx).numpy())
    return out


def gen_masks(rng):
    """The reference's MaskPostProcessor + Masker (mask_head/inference.py:11-66, :168-205) on two
    images: boxes inside, straddling every border, tiny (sub-pixel) and one fully outside."""
    from maskrcnn_benchmark.modeling.roi_heads.mask_head.inference import Masker, MaskPostProcessor
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    out = {}
    sizes = [(320, 200), (171, 243)]
    boxes = [np.array([[10.3, 20.7, 80.2, 90.9], [-25.0, -13.5, 60.0, 40.0], [250.0, 150.0, 400.0, 260.0],
                       [100.2, 50.9, 100.6, 51.3], [5.0, 3.0, 318.0, 198.0], [500.0, 500.0, 600.0, 650.0]], np.float32),
             np.array([[0.0, 0.0, 170.0, 242.0], [33.3, 77.7, 44.4, 201.1], [160.5, 230.5, 175.0, 250.0],
                       [60.0, 10.0, 61.0, 240.0]], np.float32)]
    labels = [rng.integers(1, 5, len(b)) for b in boxes]
    n = sum(len(b) for b in boxes)
    x = (rng.standard_normal((n, 5, 14, 14)) * 2).astype(np.float32)
    # smooth the logits a little so that masks have blobs, not salt and pepper
    x = (x + np.roll(x, 1, 2) + np.roll(x, 1, 3) + np.roll(x, (1, 1), (2, 3))) / 2
    bl = []
    for b, lab, sz in zip(boxes, labels, sizes):
        t = BoxList(torch.from_numpy(b), sz, mode="xyxy")
        t.add_field("labels", torch.from_numpy(lab.astype(np.int64)))
        bl.append(t)
    res = MaskPostProcessor(masker=Masker(threshold=0.5, padding=1))(torch.from_numpy(x), bl)
    out["mk_logits"] = x
    for i, (b, lab, sz, r) in enumerate(zip(boxes, labels, sizes, res)):
        out["mk_boxes%d" % i], out["mk_labels%d" % i], out["mk_size%d" % i] = b, lab.astype(np.int64), np.array(sz)
        m = r.get_field("mask").numpy()
        assert m.dtype == np.bool_ and m.shape == (len(b), 1, sz[1], sz[0])
        out["mk_mask%d_packed" % i] = np.packbits(m.reshape(-1))
        # pin the numpy restatement against the reference right here
        prob = torch.from_numpy(x).sigmoid().numpy()
        o = sum(len(bb) for bb in boxes[:i])
        sel = prob[np.arange(o, o + len(b)), lab]
        assert np.array_equal(oracle.paste_masks(sel, b, sz[1], sz[0], 0.5, 1), m[:, 0]), "oracle.paste_masks != reference"
    return out


def gen_mask_targets(rng):
    """The student's mask targets through the reference's own code: Masker pastes the pseudo-labels' 14x14 masks
    into full-image boolean masks (st_generalized_rcnn.py:267-272), SegmentationMask(mode='mask') wraps them, and
    project_masks_on_boxes (mask_head/loss.py:11-42) crops / resizes them at the proposals.  pycocotools and cv2
    are only imported by segmentation_mask.py, never called on this path: empty stub modules."""
    for name in ("pycocotools", "pycocotools.mask", "cv2"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
    from maskrcnn_benchmark.modeling.roi_heads.mask_head.inference import Masker
    from maskrcnn_benchmark.modeling.roi_heads.mask_head.loss import project_masks_on_boxes
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.segmentation_mask import SegmentationMask
    im_w, im_h, M = 333, 250, 14
    label_boxes = np.array([[20.4, 30.6, 150.2, 200.9], [-12.0, -8.5, 80.0, 60.0], [200.0, 100.0, 340.0, 260.0],
                            [100.2, 50.9, 101.6, 52.3], [5.0, 3.0, 330.0, 248.0]], np.float32)
    k = len(label_boxes)
    x = (rng.standard_normal((k, 1, 14, 14)) * 2).astype(np.float32)
    x = (x + np.roll(x, 1, 2) + np.roll(x, 1, 3) + np.roll(x, (1, 1), (2, 3))) / 2
    probs = torch.from_numpy(x).sigmoid()
    labels_bl = BoxList(torch.from_numpy(label_boxes), (im_w, im_h), mode="xyxy")
    full = Masker(threshold=0.5, padding=1)([probs], [labels_bl])[0]                 # [k, 1, H, W] bool
    # proposals: jittered copies of the label boxes (positives), half-integer corners (round-half-to-even),
    # boxes leaving the image, a sub-pixel one
    match, props = [], []
    for i in range(k):
        for _ in range(6):
            match.append(i)
            props.append(label_boxes[i] + rng.normal(0, 6, 4).astype(np.float32))
    props += [np.array([10.5, 20.5, 60.5, 90.5], np.float32), np.array([11.5, 21.5, 61.5, 91.5], np.float32),
              np.array([-30.0, -20.0, 40.0, 50.0], np.float32), np.array([300.0, 200.0, 400.0, 300.0], np.float32),
              np.array([120.2, 80.1, 120.4, 80.3], np.float32)]
    match += [0, 0, 1, 2, 0]
    props = np.stack(props).astype(np.float32)
    props[:, 2] = np.maximum(props[:, 2], props[:, 0])
    props[:, 3] = np.maximum(props[:, 3], props[:, 1])
    match = np.array(match, np.int32)
    seg = SegmentationMask(full[:, 0][torch.from_numpy(match).long()], (im_w, im_h), mode="mask")
    want = project_masks_on_boxes(seg, BoxList(torch.from_numpy(props), (im_w, im_h), mode="xyxy"), M).numpy()
    assert want.shape == (len(props), M, M) and set(np.unique(want)) <= {0.0, 1.0}
    got = oracle.mask_targets(probs[:, 0].numpy(), label_boxes, match, props, im_h, im_w, M, 0.5, 1)
    assert np.array_equal(got, want), "oracle.mask_targets != reference (%d pixels differ)" % int((got != want).sum())
    return {"mt_probs": probs[:, 0].numpy(), "mt_label_boxes": label_boxes, "mt_match": match, "mt_proposals": props,
            "mt_size": np.array([im_w, im_h, M]), "mt_targets_packed": np.packbits(want.astype(np.bool_).reshape(-1))}


def main():
    install_reference()
    rng = np.random.default_rng(20221017)
    np.savez_compressed(os.path.join(HERE, "roi_align.npz"), **gen_roi_align(rng))
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **gen_nms(rng))
    np.savez_compressed(os.path.join(HERE, "pooler.npz"), **gen_pooler(rng))
    np.savez_compressed(os.path.join(HERE, "rpn.npz"), **gen_rpn(rng))
    np.savez_compressed(os.path.join(HERE, "box_head.npz"), **gen_box_head(rng))
    np.savez_compressed(os.path.join(HERE, "masks.npz"), **gen_masks(np.random.default_rng(99)))
    np.savez_compressed(os.path.join(HERE, "mask_targets.npz"), **gen_mask_targets(np.random.default_rng(100)))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KB")


if __name__ == "__main__":
    main()
