"""One launch of the fast RoIAlign forward with every RoI on the same patch (issue-bound floor) for ncu."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward
from tests import synth
B, n, C = 16, 1000, 256
res = int(sys.argv[1]); v = int(sys.argv[2]); exact = sys.argv[3] != "fast"
rng = np.random.default_rng(1236)
g = torch.Generator(device="cuda").manual_seed(1236)
feats = [torch.randn((B, C, h, w), device="cuda", generator=g).contiguous(memory_format=torch.channels_last) for (h, w) in synth.fpn_shapes()]
rois = torch.from_numpy(synth.make_rois(rng, n, B)).cuda()
rois[:, 0] = 0; rois[:, 1] = 200; rois[:, 2] = 200; rois[:, 3] = 264; rois[:, 4] = 264
_ext.debug_set(False, exact, v)
for _ in range(2):
    _forward(feats, synth.FPN_SCALES, rois, (res, res), 2)
torch.cuda.synchronize()
