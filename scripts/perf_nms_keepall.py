"""Keep-all NMS (no max_keep): fused one-CTA-per-segment kernel against the three-kernel bitmask path, by segment length."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms_batched


def timeit(fn, iters=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))


rng = np.random.default_rng(1236)
cb, cs = bench.make_rpn_candidates(rng, 16)           # 16 images x {6000, 6000, 6000, 3150, 819}, score-sorted
lens_img = bench.RPN_LENS
for n in (500, 1000, 2000, 3000, 4000, 6000):
    # 48 segments of n boxes: the first n of every 6000-box segment
    segs, off = [], 0
    for i in range(16):
        for l in lens_img:
            if l >= 6000:
                segs.append((off, n))
            off += l
    boxes = torch.from_numpy(np.concatenate([cb[o:o + k] for o, k in segs])).cuda()
    scores = torch.from_numpy(np.concatenate([cs[o:o + k] for o, k in segs])).cuda()
    seg_off = torch.from_numpy(np.arange(len(segs) + 1, dtype=np.int32) * n).cuda()
    out = {}
    for name, force in (("fused", 2), ("bitmask", 1)):
        _ext.debug_nms(force)
        out[name] = timeit(lambda: nms_batched(boxes, scores, seg_off, 0.7, -1, n))
        ki, kc = nms_batched(boxes, scores, seg_off, 0.7, -1, n)
        out[name + "_kept"] = int(kc.sum().item())
    _ext.debug_nms(0)
    auto = timeit(lambda: nms_batched(boxes, scores, seg_off, 0.7, -1, n))
    print("keep-all, %d segments x %5d boxes: fused %.3f ms, bitmask %.3f ms, default %.3f ms (kept %d / %d)" %
          (len(segs), n, out["fused"], out["bitmask"], auto, out["fused_kept"], out["bitmask_kept"]), flush=True)
