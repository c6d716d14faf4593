"""Class-level timings of the drop-in modules at BASELINE sizes (development aid)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvpr22_cross_modal_pseudo_labeling_b200.modeling import BoxCoder, Pooler, PostProcessor, RPNPostProcessor
from cvpr22_cross_modal_pseudo_labeling_b200.layers import embed_logits
from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
from tests import synth

def timeit(fn, warm=2, iters=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(iters): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3

res = {}
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(3)
B, W, H = 16, 1344, 800
# ---- RPNPostProcessor: 5 FPN levels, 3 anchors/cell, test settings 6000 -> 1000 -> 1000/img ----
strides = (4, 8, 16, 32, 64)
sizes = (32, 64, 128, 256, 512)
anchors_per_level, obj, reg = [], [], []
for s, sz in zip(strides, sizes):
    h, w = -(-H // s), -(-W // s)
    ys, xs = torch.meshgrid(torch.arange(h, device=dev) * s, torch.arange(w, device=dev) * s, indexing="ij")
    cells = torch.stack([xs, ys, xs, ys], -1).reshape(-1, 1, 4).float()
    ar = torch.tensor([0.5, 1.0, 2.0], device=dev)
    ws_, hs_ = sz * torch.sqrt(1 / ar), sz * torch.sqrt(ar)
    base = torch.stack([-ws_ / 2, -hs_ / 2, ws_ / 2, hs_ / 2], -1)[None]
    anchors_per_level.append((cells + base).reshape(-1, 4))
    obj.append(torch.randn((B, 3, h, w), device=dev, generator=g) * 2)
    reg.append(torch.randn((B, 12, h, w), device=dev, generator=g) * 0.3)
anchors = [[BoxList(a, (1333, 800)) for a in anchors_per_level] for _ in range(B)]
rpn = RPNPostProcessor(6000, 1000, 0.7, 0, BoxCoder((1., 1., 1., 1.)), fpn_post_nms_top_n=1000).eval()
props = rpn(anchors, obj, reg)
res["rpn_postprocessor_B16_ms"] = timeit(lambda: rpn(anchors, obj, reg))
print("RPNPostProcessor B=16:", res["rpn_postprocessor_B16_ms"], "ms; proposals/img", len(props[0]), flush=True)
# ---- PostProcessor: 16 x 1000 RoIs, 66 classes ----
R, C = 1000, 66
logits = torch.randn((B * R, C), device=dev, generator=g) * 3
regs = torch.randn((B * R, 8), device=dev, generator=g) * 0.3
pp = PostProcessor(0.05, 0.5, 100, BoxCoder((10., 10., 5., 5.)), cls_agnostic_bbox_reg=True)
dets = pp((logits, regs), props)
res["box_postprocessor_B16_C66_ms"] = timeit(lambda: pp((logits, regs), props))
print("PostProcessor B=16 C=66:", res["box_postprocessor_B16_C66_ms"], "ms; dets/img", len(dets[0]), flush=True)
# ---- config #4 slice: Pooler fwd+bwd on 512 RoIs/img (B=16 here) + embedding logits fwd/bwd ----
feats = [torch.randn((B, 256, h, w), device=dev, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
         for (h, w) in synth.fpn_shapes()]
sub = [p[:512] for p in props]
pooler = Pooler((7, 7), synth.FPN_SCALES, 2)
E = torch.nn.functional.normalize(torch.randn((49, 768), device=dev, generator=g), dim=-1).to(torch.bfloat16)
Wp = torch.randn((768, 256), device=dev, generator=g) * 0.05
def train_slice():
    for f in feats: f.grad = None
    x = pooler(feats, sub)
    emb = torch.nn.functional.linear(x.mean(dim=(2, 3)), Wp)
    logit = embed_logits(emb, E)
    loss = torch.nn.functional.cross_entropy(logit, torch.zeros(logit.shape[0], dtype=torch.long, device=dev))
    loss.backward()
res["config4_slice_pooler_fwd_bwd_embed_B16x512_ms"] = timeit(train_slice)
print("config #4 slice (B=16 x 512 RoIs):", res["config4_slice_pooler_fwd_bwd_embed_B16x512_ms"], "ms", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/perf_modules.json", "w"), indent=1)
