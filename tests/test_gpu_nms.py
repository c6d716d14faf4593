"""GPU parity: batched NMS (C ABI b200_nms_batched) vs the CPU oracle -- keep indices bit-exact."""
import numpy as np
import pytest
import torch

import oracle
from tests import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["fused", "bitmask"], autouse=True)
def nms_path(request):
    """Every test runs on both device paths: the fused shared-memory kernel and the
    three-kernel 64x64 bitmask + sweep path (used for segments > 13952 boxes)."""
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    _ext.debug_nms(1 if request.param == "bitmask" else 2)   # (2: the fused kernel wherever it fits)
    yield request.param
    _ext.debug_nms(0)


def _gpu_nms(boxes, scores, thr):
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms
    k = nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), thr)
    assert k.dtype == torch.int64 and k.is_cuda
    return k.cpu().numpy()


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 127, 129, 1000, 2000, 6000, 12000, 16384])
@pytest.mark.parametrize("thr", [0.5, 0.7])
def test_nms_matches_oracle(n, thr):
    rng = np.random.default_rng(100 + n)
    boxes, scores = synth.make_nms_boxes(rng, n)
    assert len(np.unique(scores)) == n
    got = _gpu_nms(boxes, scores, thr)
    want = oracle.nms(boxes, scores, thr)
    assert np.array_equal(got, want)
    assert np.all(np.diff(got) > 0)  # ascending original indices (nms_cpu.cpp:64)


def test_nms_longer_than_one_segment_call(nms_path):
    """layers.nms is installed over _C.nms for every caller, and the reference takes any N: segments beyond
    B200_NMS_MAX_SEG run block-recursively (layers/nms.py:_nms_long), keep list still bit-exact, with ties."""
    if nms_path == "bitmask":
        pytest.skip("one device path is enough for the 40k-box case")
    rng = np.random.default_rng(4242)
    boxes, scores = synth.make_nms_boxes(rng, 40000)
    for s in (scores, (np.round(scores * 64) / 64).astype(np.float32)):
        assert np.array_equal(_gpu_nms(boxes, s, 0.5), oracle.nms(boxes, s, 0.5))


def test_nms_sorted_input_and_ties():
    rng = np.random.default_rng(7)
    boxes, scores = synth.make_nms_boxes(rng, 3000)
    order = np.argsort(-scores, kind="stable")
    b, s = boxes[order], scores[order]
    assert np.array_equal(_gpu_nms(b, s, 0.7), oracle.nms(b, s, 0.7))
    # saturated objectness: many exactly-equal scores; defined order = lower index first
    s2 = np.round(s * 8) / 8
    assert len(np.unique(s2)) < 20
    assert np.array_equal(_gpu_nms(b, s2.astype(np.float32), 0.7), oracle.nms(b, s2.astype(np.float32), 0.7))


def test_nms_known_answers():
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms
    # identical boxes at thr=1.0: IoU == 1 >= 1 -> second suppressed (CPU `>=`; the CUDA reference's `>` keeps both)
    b = torch.tensor([[10., 10., 50., 50.], [10., 10., 50., 50.]], device="cuda")
    s = torch.tensor([0.9, 0.8], device="cuda")
    assert nms(b, s, 1.0).tolist() == [0]
    assert nms(b, s.flip(0), 1.0).tolist() == [1]
    # disjoint boxes all survive, ascending order
    b = torch.tensor([[0., 0., 9., 9.], [20., 20., 29., 29.], [40., 40., 49., 49.]], device="cuda")
    s = torch.tensor([0.1, 0.9, 0.5], device="cuda")
    assert nms(b, s, 0.5).tolist() == [0, 1, 2]
    # empty input
    e = nms(torch.zeros((0, 4), device="cuda"), torch.zeros((0,), device="cuda"), 0.5)
    assert e.numel() == 0 and e.dtype == torch.int64
    # legacy +1: [0,0,9,9] vs [5,0,14,9] -> inter 5*10=50, union 150 -> IoU 1/3
    b = torch.tensor([[0., 0., 9., 9.], [5., 0., 14., 9.]], device="cuda")
    s = torch.tensor([0.9, 0.8], device="cuda")
    assert nms(b, s, 0.34).tolist() == [0, 1]
    assert nms(b, s, 0.33).tolist() == [0]


def test_nms_cpu_tensor_is_an_error():
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms
    with pytest.raises(RuntimeError):
        nms(torch.zeros((4, 4)), torch.zeros((4,)), 0.5)


@pytest.mark.parametrize("max_keep", [-1, 50])
def test_nms_batched_ragged_segments(max_keep):
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms_batched
    rng = np.random.default_rng(11)
    lens = [0, 1, 700, 64, 0, 1500, 65, 3, 2048, 0]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    bs, ss = [], []
    for i, n in enumerate(lens):
        b, s = synth.make_nms_boxes(rng, max(n, 1))
        if i % 2 == 0:  # half of the segments arrive sorted (RPN-style) -> early-stop path
            o = np.argsort(-s, kind="stable")
            b, s = b[o], s[o]
        bs.append(b[:n]); ss.append(s[:n])
    boxes, scores = np.concatenate(bs), np.concatenate(ss)
    want_idx, want_cnt = oracle.nms_batched(boxes, scores, off, 0.6, max_keep)
    gi, gc = nms_batched(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(),
                         torch.from_numpy(off.astype(np.int32)).cuda(), 0.6, max_keep, max(lens))
    gi, gc = gi.cpu().numpy(), gc.cpu().numpy()
    assert np.array_equal(gc, want_cnt)
    for s in range(len(lens)):
        a, k = off[s], want_cnt[s]
        assert np.array_equal(gi[a:a + k], want_idx[a:a + k]), s
        assert np.all(gi[a + k:off[s + 1]] == -1)


def test_nms_full_size_properties():
    """BASELINE config #2 size (16 images x 5 levels, up to 6000 boxes): invariants that
    need no oracle run -- kept boxes pairwise IoU < thr; NMS(NMS(x)) == NMS(x)."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms_batched
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList, boxlist_iou
    rng = np.random.default_rng(5)
    lens = [6000, 6000, 6000, 3150, 819] * 16
    off = np.concatenate([[0], np.cumsum(lens)])
    bs, ss = zip(*[synth.make_nms_boxes(rng, n) for n in lens])
    boxes = torch.from_numpy(np.concatenate(bs)).cuda()
    scores = torch.from_numpy(np.concatenate(ss)).cuda()
    offs = torch.from_numpy(off.astype(np.int32)).cuda()
    ki, kc = nms_batched(boxes, scores, offs, 0.7, -1, 6000)
    kc_h = kc.cpu().numpy()
    assert kc_h.min() > 0
    for s in (0, 3, 4, 79):
        a = int(off[s]); k = int(kc_h[s])
        kept = ki[a:a + k] + a
        bl = BoxList(boxes[kept], (synth.IMG_W, synth.IMG_H))
        iou = boxlist_iou(bl, bl)
        iou.fill_diagonal_(0)
        assert float(iou.max()) < 0.7
        # idempotence on the kept set
        k2, c2 = nms_batched(boxes[kept], scores[kept], torch.tensor([0, k], dtype=torch.int32, device="cuda"), 0.7)
        assert int(c2.item()) == k
        # spot-check one full segment against the oracle
    s = 4
    a, b = int(off[s]), int(off[s + 1])
    want = oracle.nms(boxes[a:b].cpu().numpy(), scores[a:b].cpu().numpy(), 0.7)
    assert np.array_equal(ki[a:a + int(kc_h[s])].cpu().numpy(), want)


@pytest.mark.parametrize("sorted_input", [True, False])
def test_select_topk_matches_numpy(sorted_input):
    """Per-image cross-level top-k after the batched NMS (reference rpn/inference.py:173-180):
    sorted segments take the rank-merge path, unsorted ones the shared-memory sort; scores are
    quantised so that equal scores occur across levels (ties: ascending box index)."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms_batched, select_topk
    rng = np.random.default_rng(21)
    n_img, lens = 3, [700, 300, 64, 5]
    all_lens = lens * n_img
    off = np.concatenate([[0], np.cumsum(all_lens)])
    bs, ss = [], []
    for L in all_lens:
        b, s = synth.make_nms_boxes(rng, L)
        s = (np.round(s * rng.uniform(0.5, 1.0) * 512) / 512).astype(np.float32)   # levels: different ranges, ties
        o = np.argsort(-s, kind="stable") if sorted_input else rng.permutation(L)
        bs.append(b[o]); ss.append(s[o])
    boxes, scores = np.concatenate(bs), np.concatenate(ss)
    tb, ts = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
    to = torch.from_numpy(off.astype(np.int32)).cuda()
    for post, top_n in ((100, 150), (50, 400)):
        ki, kc = nms_batched(tb, ts, to, 0.7, post, max(lens))
        rois, sc, cnt = select_topk(tb, ts, to, ki, kc, n_img, top_n, len(lens) * post)
        ki_h, kc_h = ki.cpu().numpy(), kc.cpu().numpy()
        for i in range(n_img):
            gidx = np.concatenate([off[s] + ki_h[off[s]:off[s] + kc_h[s]] for s in range(i * 4, i * 4 + 4)])
            order = np.lexsort((gidx, -scores[gidx]))[:top_n]
            want = gidx[order]
            n = int(cnt[i])
            assert n == len(want)
            got = rois[i * top_n:i * top_n + n].cpu().numpy()
            assert np.all(got[:, 0] == i)
            assert np.array_equal(got[:, 1:], boxes[want])
            assert np.array_equal(sc[i * top_n:i * top_n + n].cpu().numpy(), scores[want])
            assert np.all(rois[i * top_n + n:(i + 1) * top_n, 1:].cpu().numpy() == 0)


@pytest.mark.parametrize("lens", [[1500, 900, 2048], [40] * 60 + [700], [600, 600]])
def test_nms_default_dispatch_keep_all(lens):
    """The library's own choice of path (test hook mode 0): keep-all calls over long segments go to the bitmask
    kernels, many short segments (or max_keep) to the fused kernel -- the keep lists are the oracle's either way."""
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms_batched
    rng = np.random.default_rng(len(lens))
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    bs, ss = zip(*[synth.make_nms_boxes(rng, n) for n in lens])
    boxes, scores = np.concatenate(bs), np.concatenate(ss)
    _ext.debug_nms(0)
    for max_keep in (-1, 100):
        want_idx, want_cnt = oracle.nms_batched(boxes, scores, off, 0.6, max_keep)
        gi, gc = nms_batched(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(),
                             torch.from_numpy(off.astype(np.int32)).cuda(), 0.6, max_keep, max(lens))
        gi, gc = gi.cpu().numpy(), gc.cpu().numpy()
        assert np.array_equal(gc, want_cnt)
        for s in range(len(lens)):
            a, k = off[s], want_cnt[s]
            assert np.array_equal(gi[a:a + k], want_idx[a:a + k]), (s, max_keep)
