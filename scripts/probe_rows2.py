"""Probe: row-streaming RoIAlign forward on (a) the microbench RoIs and (b) the bench step's RoIs (NMS output),
with / without the visiting-order pass; the separable kernel and the copy-only probe for reference."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import synth
import bench
from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms_batched, select_topk
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
shapes = synth.fpn_shapes()
feats = [torch.randn((B, C, h, w), device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
         for (h, w) in shapes]
micro = torch.from_numpy(synth.make_rois(rng, 1000, B)).cuda()
cb, cs = bench.make_rpn_candidates(rng, B)
cb, cs = torch.from_numpy(cb).cuda(), torch.from_numpy(cs).cuda()
seg_off = torch.from_numpy(np.concatenate([[0], np.cumsum(bench.RPN_LENS * B)]).astype(np.int32)).cuda()
ki, kc = nms_batched(cb, cs, seg_off, 0.7, 1000, max(bench.RPN_LENS))
step_rois, _, _ = select_topk(cb, cs, seg_off, ki, kc, B, 1000, 5000)
variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 32, 64, 16]
for name, rois in (("microbench", micro), ("bench step", step_rois)):
    _, lv = _forward(feats, synth.FPN_SCALES, rois, (7, 7), 2, want_levels=True, math="fast")
    ft = bench.f_touched_bytes(rois, lv, shapes, synth.FPN_SCALES, C)
    algo = ft + rois.shape[0] * (C * 49 * 4 + 20)
    mean = torch.empty((rois.shape[0], C), device="cuda")
    for v in variants:
        _ext.debug_set(False, True, v)
        t, tmin = timeit(lambda: _forward(feats, synth.FPN_SCALES, rois, (7, 7), 2, math="fast"))
        tm, _ = timeit(lambda: _forward(feats, synth.FPN_SCALES, rois, (7, 7), 2, math="fast", mean_out=mean))
        print("%-10s variant %2d: %.3f ms (min %.3f; with mean %.3f)  %.0f GB/s = %.3f of 6548.5" %
              (name, v, t, tmin, tm, algo / t / 1e6, algo / t / 1e6 / 6548.5), flush=True)
_ext.debug_set(False, True, 0)
# bf16 maps (config #4): same RoIs, half the bytes through the ring
fb = [f.to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for f in feats]
for name, rois in (("microbench", micro), ("bench step", step_rois)):
    _, lv = _forward(fb, synth.FPN_SCALES, rois, (7, 7), 2, want_levels=True, math="fast")
    ft = bench.f_touched_bytes(rois, lv, shapes, synth.FPN_SCALES, C) // 2
    algo = ft + rois.shape[0] * (C * 49 * 4 + 20)
    for v in variants:
        _ext.debug_set(False, True, v)
        t, tmin = timeit(lambda: _forward(fb, synth.FPN_SCALES, rois, (7, 7), 2, math="fast"))
        print("%-10s bf16 maps variant %4d: %.3f ms (min %.3f)  %.0f GB/s = %.3f of 6548.5" % (name, v, t, tmin, algo / t / 1e6, algo / t / 1e6 / 6548.5), flush=True)
_ext.debug_set(False, True, 0)
