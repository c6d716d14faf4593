"""Probe: the bench step's RPN NMS (80 segments, thr 0.7, keep 1000) alone, CUDA events, median of 20."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms_batched

B = 16
rng = np.random.default_rng(1236)
cb, cs = bench.make_rpn_candidates(rng, B)
cb, cs = torch.from_numpy(cb).cuda(), torch.from_numpy(cs).cuda()
seg_off = torch.from_numpy(np.concatenate([[0], np.cumsum(bench.RPN_LENS * B)]).astype(np.int32)).cuda()
for mk in (1000, 300, -1):
    fn = lambda: nms_batched(cb, cs, seg_off, 0.7, mk, max(bench.RPN_LENS))
    for _ in range(3):
        fn()
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ki, kc = fn()
    print("max_keep %5d: %.3f ms (min %.3f)  kept per segment: mean %.0f max %d" %
          (mk, float(np.median(ts)), min(ts), float(kc.float().mean()), int(kc.max())), flush=True)

# phase timers (library built with -DB200_NMS_STATS): cycles of thread 0 per phase, per segment class
import ctypes
from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
l = _ext.lib()
if hasattr(l, "b200_debug_nms_stats"):
    nms_batched(cb, cs, seg_off, 0.7, 1000, max(bench.RPN_LENS))
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (256 * 16))()
    l.b200_debug_nms_stats.argtypes = [ctypes.c_void_p]
    assert l.b200_debug_nms_stats(buf) == 0
    a = np.frombuffer(buf, dtype=np.int64).reshape(256, 16)[:80].astype(np.float64)
    names = ["load/order", "-", "A pull + diagonal words", "-", "B chain", "-", "C rank new", "D merge", "compaction", "tiles"]
    lens = np.repeat(np.asarray(bench.RPN_LENS), B) if len(bench.RPN_LENS) * B == 80 else np.tile(np.asarray(bench.RPN_LENS), B)
    seg_len = (seg_off[1:] - seg_off[:-1]).cpu().numpy()
    for L in sorted(set(seg_len.tolist()), reverse=True):
        m = seg_len == L
        print("segments of %d boxes (%d): total %.0f cycles" % (L, m.sum(), a[m][:, :9].sum(axis=1).mean()))
        for i, nme in enumerate(names):
            print("   %-26s %10.0f" % (nme, a[m][:, i].mean()))
