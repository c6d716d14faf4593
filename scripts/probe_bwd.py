import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch, synth
from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _backward, _levels_array, _workspace
B, C = 16, 256
rng = np.random.default_rng(1236)
rois = torch.from_numpy(synth.make_rois(rng, 1000, B)).cuda()
shapes = [(B, C, h, w) for (h, w) in synth.fpn_shapes()]
lib = _ext.lib()
for res in (7, 14):
    g = torch.randn((rois.shape[0], C, res, res), device="cuda")
    grads = [torch.zeros(s, device="cuda").contiguous(memory_format=torch.channels_last) for s in shapes]
    arr = _levels_array(grads, synth.FPN_SCALES)
    ws = _workspace(lib.b200_roi_align_workspace_bytes(rois.shape[0]), rois.device)
    for name, w, wb in (("order given", None, 0), ("visiting order", ws, ws.numel())):
        ts = []
        for _ in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = lib.b200_roi_align_backward_ws(arr, 4, _ext.B200_LAYOUT_NHWC, B, C, _ext.ptr(rois), rois.shape[0], res, res, 2,
                                                _ext.ptr(g), _ext.ptr(w), wb, _ext.stream_ptr(rois.device))
            b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
        print("bwd %dx%d kernel only, %-14s %.3f ms" % (res, res, name + ":", float(np.median(ts[1:]))), flush=True)
    t = []
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _backward(g, rois, shapes, True, synth.FPN_SCALES, (res, res), 2); b.record(); b.synchronize(); t.append(a.elapsed_time(b))
    print("bwd %dx%d _backward incl. zero fill: %.3f ms" % (res, res, float(np.median(t[1:]))), flush=True)
