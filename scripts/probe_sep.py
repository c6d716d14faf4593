"""Probe: occupancy / buffering variants of the separable forward kernel (and the issue-bound floor:
every RoI reading one L1-resident patch)."""
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
rng = np.random.default_rng(1236)
g = torch.Generator(device="cuda").manual_seed(1236)
feats = [torch.randn((B, C, h, w), device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
         for (h, w) in synth.fpn_shapes()]
rois = torch.from_numpy(synth.make_rois(rng, 1000, B)).cuda()
variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0]
r1 = rois.clone(); r1[:, 0] = 0; r1[:, 1] = 200; r1[:, 2] = 200; r1[:, 3] = 264; r1[:, 4] = 264
for res in (7, 14):
    for v in variants:
        _ext.debug_set(False, False, v)
        t = timeit(lambda: _forward(feats, synth.FPN_SCALES, rois, (res, res), 2))
        t1 = timeit(lambda: _forward(feats, synth.FPN_SCALES, r1, (res, res), 2))
        print("res %d fast variant %d: %.3f ms   (same-patch floor %.3f ms)" % (res, v, t, t1), flush=True)
_ext.debug_set(False, True, 0)
# hot-line contention probe: same RoIs, order shuffled across images
perm = torch.randperm(rois.shape[0], device="cuda")
rs = rois[perm].contiguous()
# and: RoIs interleaved by image (r0 of img0, r0 of img1, ...)
ri = rois.view(B, -1, 5).transpose(0, 1).reshape(-1, 5).contiguous()
for res in (7, 14):
    _ext.debug_set(False, False, 0)
    t = timeit(lambda: _forward(feats, synth.FPN_SCALES, rois, (res, res), 2))
    ts_ = timeit(lambda: _forward(feats, synth.FPN_SCALES, rs, (res, res), 2))
    ti = timeit(lambda: _forward(feats, synth.FPN_SCALES, ri, (res, res), 2))
    print("res %d: image-ordered %.3f ms | shuffled %.3f ms | interleaved %.3f ms" % (res, t, ts_, ti), flush=True)
_ext.debug_set(False, True, 0)
r0 = rois.clone(); r0[:, 0] = 0
for res in (7, 14):
    _ext.debug_set(False, False, 0)
    t0 = timeit(lambda: _forward(feats, synth.FPN_SCALES, r0, (res, res), 2))
    print("res %d: all RoIs on image 0 (features L2-resident) %.3f ms" % (res, t0), flush=True)
_ext.debug_set(False, True, 0)
