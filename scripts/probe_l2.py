"""Probe: is the RoIAlign forward bound by DRAM or by the L2->SM path?
Runs the box pooler with (a) RoIs spread over 16 images (1.46 GB of features, DRAM-resident),
(b) the same RoIs all pointed at image 0 (91 MB, L2-resident)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import synth
from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward

def timeit(fn, iters=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))

B, C = 16, 256
rng = np.random.default_rng(1)
g = torch.Generator(device="cuda").manual_seed(1)
feats = [torch.randn((B, C, h, w), device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
         for (h, w) in synth.fpn_shapes()]
rois_np = synth.make_rois(rng, 1000, B)
rois = torch.from_numpy(rois_np).cuda()
r0 = rois.clone(); r0[:, 0] = 0
for res in (7, 14):
    for exact in (True, False):
        _ext.debug_set(False, exact, 0)
        t_all = timeit(lambda: _forward(feats, synth.FPN_SCALES, rois, (res, res), 2))
        t_one = timeit(lambda: _forward(feats, synth.FPN_SCALES, r0, (res, res), 2))
        print("res %d exact %d: 16 images %.3f ms | all on image 0 (L2-resident) %.3f ms" % (res, exact, t_all, t_one), flush=True)
_ext.debug_set(False, True, 0)
