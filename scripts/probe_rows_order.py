"""Probe: does the order of the RoIs inside an image matter for the row-streaming RoIAlign forward?
(image-major as given | sorted by (image, FPN level, y centre) | by (image, level, y band of 64 px, x))"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import synth
from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))


B, C = 16, 256
rng = np.random.default_rng(1236)
g = torch.Generator(device="cuda").manual_seed(1236)
feats = [torch.randn((B, C, h, w), device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
         for (h, w) in synth.fpn_shapes()]
rois = torch.from_numpy(synth.make_rois(rng, 1000, B)).cuda()
_, lv = _forward(feats, synth.FPN_SCALES, rois, (7, 7), 2, want_levels=True)
cy = (rois[:, 2] + rois[:, 4]) * 0.5
cx = (rois[:, 1] + rois[:, 3]) * 0.5
img = rois[:, 0].long()
k1 = (img * 8 + lv.long()) * 4096 + cy.long().clamp(0, 4095)
k2 = ((img * 8 + lv.long()) * 64 + (cy / 64).long().clamp(0, 63)) * 4096 + cx.long().clamp(0, 4095)
orders = {"as given": rois, "by (image, level, y)": rois[torch.argsort(k1)].contiguous(),
          "by (image, level, y band, x)": rois[torch.argsort(k2)].contiguous()}
for v in (0, 16):
    _ext.debug_set(False, False, v)
    for name, r in orders.items():
        print("variant %d, RoIs %-30s %.3f ms" % (v, name + ":", timeit(lambda: _forward(feats, synth.FPN_SCALES, r, (7, 7), 2))),
              flush=True)
_ext.debug_set(False, True, 0)
