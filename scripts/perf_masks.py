"""Device time of Masker.forward_single_image (b200_paste_masks): 100 detections on an 800x1333 image."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvpr22_cross_modal_pseudo_labeling_b200.modeling import Masker
from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
rng = np.random.default_rng(0)
n, im_w, im_h = 100, 1333, 800
masks = torch.rand((n, 1, 28, 28), device="cuda")
x1 = rng.uniform(0, im_w - 50, n); y1 = rng.uniform(0, im_h - 50, n)
boxes = torch.from_numpy(np.stack([x1, y1, x1 + rng.uniform(20, 400, n), y1 + rng.uniform(20, 400, n)], 1).astype(np.float32)).cuda()
bl = BoxList(boxes, (im_w, im_h))
mk = Masker(0.5, 1)
for _ in range(3): mk.forward_single_image(masks, bl)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10): out = mk.forward_single_image(masks, bl)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print("paste 100 masks 28x28 -> 800x1333: %.3f ms, %.0f GB/s of output" % (ms, out.numel() / ms / 1e6))
