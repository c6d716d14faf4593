"""Probe: row-streaming RoIAlign forward (variant 0, the default) against the separable marching kernel
(variant 16) on the box-pooler microbench (16 images x 1000 RoIs, 256 channels, 7x7)."""
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
    return float(np.median(ts)), float(np.min(ts))


B, C = 16, 256
rng = np.random.default_rng(1236)
g = torch.Generator(device="cuda").manual_seed(1236)
feats = [torch.randn((B, C, h, w), device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
         for (h, w) in synth.fpn_shapes()]
rois = torch.from_numpy(synth.make_rois(rng, 1000, B)).cuda()
r1 = rois.clone(); r1[:, 0] = 0; r1[:, 1] = 200; r1[:, 2] = 200; r1[:, 3] = 264; r1[:, 4] = 264
r0 = rois.clone(); r0[:, 0] = 0
variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [16, 0]
for v in variants:
    _ext.debug_set(False, False, v)
    t, tmin = timeit(lambda: _forward(feats, synth.FPN_SCALES, rois, (7, 7), 2))
    t1, _ = timeit(lambda: _forward(feats, synth.FPN_SCALES, r1, (7, 7), 2))
    t0, _ = timeit(lambda: _forward(feats, synth.FPN_SCALES, r0, (7, 7), 2))
    print("variant %d: %.3f ms (min %.3f) | same patch %.3f ms | all on image 0 %.3f ms" % (v, t, tmin, t1, t0),
          flush=True)
_ext.debug_set(False, True, 0)
