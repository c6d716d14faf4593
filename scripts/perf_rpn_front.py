"""Device time of b200_rpn_candidates at the RPN test settings (B=16, 5 FPN levels, 3 anchors/cell, top 6000)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvpr22_cross_modal_pseudo_labeling_b200.modeling import BoxCoder, RPNPostProcessor
from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
dev = "cuda"; g = torch.Generator(device=dev).manual_seed(3); B, W, H = 16, 1344, 800
anchors_per_level, obj, reg = [], [], []
for s, sz in zip((4, 8, 16, 32, 64), (32, 64, 128, 256, 512)):
    h, w = -(-H // s), -(-W // s)
    ys, xs = torch.meshgrid(torch.arange(h, device=dev) * s, torch.arange(w, device=dev) * s, indexing="ij")
    cells = torch.stack([xs, ys, xs, ys], -1).reshape(-1, 1, 4).float()
    ar = torch.tensor([0.5, 1.0, 2.0], device=dev)
    ws_, hs_ = sz * torch.sqrt(1 / ar), sz * torch.sqrt(ar)
    anchors_per_level.append((cells + torch.stack([-ws_ / 2, -hs_ / 2, ws_ / 2, hs_ / 2], -1)[None]).reshape(-1, 4))
    obj.append(torch.randn((B, 3, h, w), device=dev, generator=g) * 2)
    reg.append(torch.randn((B, 12, h, w), device=dev, generator=g) * 0.3)
anchors = [[BoxList(a, (1333, 800)) for a in anchors_per_level] for _ in range(B)]
rpn = RPNPostProcessor(6000, 1000, 0.7, 0, BoxCoder((1., 1., 1., 1.)), fpn_post_nms_top_n=1000).eval()
per_level = list(zip(*anchors)); sizes = [a[0].size for a in anchors]
def ev(fn, it=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it
print("fused front (device+launch) %.3f ms" % ev(lambda: rpn._decode_all_fused(per_level, obj, reg, sizes)))
print("torch front %.3f ms" % ev(lambda: [rpn._decode_level(a, o, b) for a, o, b in zip(per_level, obj, reg)]))
print("whole forward %.3f ms" % ev(lambda: rpn(anchors, obj, reg)))
