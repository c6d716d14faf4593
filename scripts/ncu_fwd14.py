"""ncu target: the mask pooler shape at R = 16 000 (16 images x 1000 RoIs, 256 channels, 14 x 14, fast math):
roi_align_fwd_sep<256, 4, 14, 14, 1, 1>; two launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth
from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward
B, C = 16, 256
rng = np.random.default_rng(1236)
g = torch.Generator(device="cuda").manual_seed(1236)
feats = [torch.randn((B, C, h, w), device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
         for (h, w) in synth.fpn_shapes()]
rois = torch.from_numpy(synth.make_rois(rng, 1000, B)).cuda()
for _ in range(2):
    _forward(feats, synth.FPN_SCALES, rois, (14, 14), 2, math="fast")
torch.cuda.synchronize()
