import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms_batched
from tests import synth
rng = np.random.default_rng(0)
def t(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
for nseg in (1, 80):
    lens = [6000] * nseg
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    bs, ss = [], []
    for L in lens:
        b_, s_ = synth.make_nms_boxes(rng, L)
        o = np.argsort(-s_, kind="stable"); bs.append(b_[o]); ss.append(s_[o])
    boxes = torch.from_numpy(np.concatenate(bs)).cuda(); scores = torch.from_numpy(np.concatenate(ss)).cuda()
    offs = torch.from_numpy(off).cuda()
    for mk in (-1, 64, 200, 500, 1000, 2000, 4000):
        ms = t(lambda: nms_batched(boxes, scores, offs, 0.7, mk, 6000))
        ki, kc = nms_batched(boxes, scores, offs, 0.7, mk, 6000)
        print("nseg", nseg, "max_keep", mk, "ms %.3f" % ms, "kept[0]", int(kc[0]), flush=True)
