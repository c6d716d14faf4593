import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms_batched
from tests import synth
rng = np.random.default_rng(0)
mk = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nseg = int(sys.argv[2]) if len(sys.argv) > 2 else 8
lens = [6000] * nseg
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
bs, ss = [], []
for L in lens:
    b_, s_ = synth.make_nms_boxes(rng, L)
    o = np.argsort(-s_, kind="stable"); bs.append(b_[o]); ss.append(s_[o])
boxes = torch.from_numpy(np.concatenate(bs)).cuda(); scores = torch.from_numpy(np.concatenate(ss)).cuda()
offs = torch.from_numpy(off).cuda()
for _ in range(2):
    nms_batched(boxes, scores, offs, 0.7, mk, 6000)
torch.cuda.synchronize()
