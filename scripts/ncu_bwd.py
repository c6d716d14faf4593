"""One launch of the RoIAlign backward (box 7x7 / mask 14x14, B=16) for an ncu capture."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _backward
from tests import synth
B, n, C = 16, 1000, 256
res = int(sys.argv[1]) if len(sys.argv) > 1 else 7
rng = np.random.default_rng(1236)
rois = torch.from_numpy(synth.make_rois(rng, n, B)).cuda()
shapes = [(B, C, h, w) for (h, w) in synth.fpn_shapes()]
g = torch.randn((rois.shape[0], C, res, res), device="cuda")
for _ in range(2):
    _backward(g, rois, shapes, True, synth.FPN_SCALES, (res, res), 2)
torch.cuda.synchronize()
