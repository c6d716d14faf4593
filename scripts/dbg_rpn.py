import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_rpn_front import _levels
from cvpr22_cross_modal_pseudo_labeling_b200.modeling import RPNPostProcessor
anchors, obj, reg = _levels(1, 2, [(50, 84), (25, 42), (13, 21), (7, 11)], shared=True)
pp = RPNPostProcessor(300, 1000, 0.7, 0)
per_level = list(zip(*anchors)); sizes = [a[0].size for a in anchors]
boxes, score, ks = pp._decode_all_fused(per_level, obj, reg, sizes)
K = sum(ks); boxes, score = boxes.view(2, K, 4), score.view(2, K)
o0 = 0
for a, o, b, k in zip(per_level, obj, reg, ks):
    p, s = pp._decode_level(a, o, b)
    d = (boxes[:, o0:o0+k] - p).abs().amax(-1)
    bad = (d > 1e-3).nonzero()
    print("level k", k, "bad rows", bad.shape[0], "max", float(d.max()))
    for n, j in bad[:6].tolist():
        print("  img", n, "pos", j, "score", float(score[n, o0+j]), float(s[n, j]), "neighbors", s[n, max(0,j-1):j+2].tolist(), boxes[n, o0+j].tolist(), p[n, j].tolist())
    o0 += k
