"""ncu target: the row-streaming box pooler on the bench step's RoIs (NMS output), variants from argv
(0 = ordered, 32 = order given, 64 = copy-only probe); two launches per variant."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth, bench
from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms_batched, select_topk
from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward
B, C = 16, 256
rng = np.random.default_rng(1236)
g = torch.Generator(device="cuda").manual_seed(1236)
feats = [torch.randn((B, C, h, w), device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
         for (h, w) in synth.fpn_shapes()]
which = sys.argv[2] if len(sys.argv) > 2 else "step"
if which == "micro":
    rois = torch.from_numpy(synth.make_rois(rng, 1000, B)).cuda()
else:
    cb, cs = bench.make_rpn_candidates(rng, B)
    cb, cs = torch.from_numpy(cb).cuda(), torch.from_numpy(cs).cuda()
    seg_off = torch.from_numpy(np.concatenate([[0], np.cumsum(bench.RPN_LENS * B)]).astype(np.int32)).cuda()
    ki, kc = nms_batched(cb, cs, seg_off, 0.7, 1000, max(bench.RPN_LENS))
    rois, _, _ = select_topk(cb, cs, seg_off, ki, kc, B, 1000, 5000)
if len(sys.argv) > 3 and sys.argv[3] == "bf16":
    feats = [f.to(torch.bfloat16).contiguous(memory_format=torch.channels_last) for f in feats]
mean = torch.empty((rois.shape[0], C), device="cuda")
for v in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0").split(",")]:
    _ext.debug_set(False, True, v)
    for _ in range(2):
        _forward(feats, synth.FPN_SCALES, rois, (7, 7), 2, math="fast", mean_out=mean)
torch.cuda.synchronize()
_ext.debug_set(False, True, 0)
