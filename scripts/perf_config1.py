"""BASELINE config #1 (the reference's shipped pooler: C4 map [1, 1024, 50, 84], 14 x 14 bins, sampling_ratio 0, 1000
RoIs): the tiled gather against the plain gather.  Variants = b200_debug_set bits (0: 32 channels per CTA, 8192: 64,
16384: the plain gather)."""
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

dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(1235)
c4 = torch.randn((1, 1024, 50, 84), device=dev, generator=gen).contiguous(memory_format=torch.channels_last)
r1 = torch.from_numpy(synth.make_rois(np.random.default_rng(1235), 1000, 1)).to(dev)
flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
algo = c4.numel() * 4 + 1000 * (1024 * 196 * 4 + 20)
variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 8192, 16384]
for math in ("exact", "fast"):
    for v in variants:
        _ext.debug_set(False, True, v)
        ts = []
        for _ in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _forward([c4], (1.0 / 16,), r1, (14, 14), 0, math=math)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts[2:]))
        print("config #1 %-5s variant %5d: %.3f ms  %.0f GB/s = %.3f of 6548.5" % (math, v, ms, algo / ms / 1e6, algo / ms / 1e6 / 6548.5), flush=True)
_ext.debug_set(False, True, 0)
