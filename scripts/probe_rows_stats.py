"""Probe: who waits for whom in roi_align_fwd_rows (variant 512 = cycle counters, b200_debug_rows_stats)."""
import ctypes
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

B, C = 16, 256
rng = np.random.default_rng(1236)
g = torch.Generator(device="cuda").manual_seed(1236)
shapes = synth.fpn_shapes()
feats = [torch.randn((B, C, h, w), device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
         for (h, w) in shapes]
rois = torch.from_numpy(synth.make_rois(rng, 1000, B)).cuda()
names = ["cons total", "cons full_wait", "cons tab_wait", "cons bar0", "cons bar1", "entries", "full misses",
         "copy dep_wait", "copy tab_wait", "copy total", "plan tab_empty wait", "plan turn wait", "plan total",
         "copy dep misses", "cons13 full_wait", "cons13 bars"]
for v in (512, 1536):
    _ext.debug_set(False, True, v)
    for _ in range(3):
        _forward(feats, synth.FPN_SCALES, rois, (7, 7), 2, math="fast")
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (160 * 16))()
    l = _ext.lib()
    l.b200_debug_rows_stats.restype = ctypes.c_int
    l.b200_debug_rows_stats.argtypes = [ctypes.c_void_p]
    assert l.b200_debug_rows_stats(buf) == 0
    a = np.frombuffer(buf, dtype=np.uint64).reshape(160, 16)[:148].astype(np.float64)
    print("variant", v)
    for i, n in enumerate(names):
        print("  %-22s mean %12.0f  min %12.0f  max %12.0f" % (n, a[:, i].mean(), a[:, i].min(), a[:, i].max()))
    print("  per entry: total %.0f cycles, full_wait %.0f" % ((a[:, 0] / a[:, 5]).mean(), (a[:, 1] / a[:, 5]).mean()))
_ext.debug_set(False, True, 0)
