"""Embedding-match kernel timing: BASELINE configs #3 (64k x 768 x 66) and #5 (262k x 512 x 501)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvpr22_cross_modal_pseudo_labeling_b200.layers import embed_match_softmax, caption_align

def timeit(fn, warm=3, iters=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]

res = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, (r, c, d) in {"cfg3_64000x66x768": (64000, 66, 768), "cfg5_262144x501x512": (262144, 501, 512),
                        "lvis_like_8192x512x768": (8192, 512, 768)}.items():
    A = (torch.randn((r, d), device="cuda") * 3).to(torch.bfloat16)
    E = torch.nn.functional.normalize(torch.randn((c, d), device="cuda"), dim=-1).to(torch.bfloat16)
    for mode, kw in (("top_only", dict(want_probs=False)), ("probs", dict(want_probs=True))):
        ms = timeit(lambda: embed_match_softmax(A, E, 0.05, **kw))
        flops = 2.0 * r * c * d
        byts = r * d * 2 + c * d * 2 + r * 8 + (r * c * 4 if mode == "probs" else 0)
        res["%s_%s" % (name, mode)] = dict(ms=ms, tflops=flops / ms / 1e9, gbs=byts / ms / 1e6)
        print(name, mode, res["%s_%s" % (name, mode)], flush=True)
    ref = timeit(lambda: torch.softmax(A @ E.t(), -1), iters=5)
    res[name + "_torch_bf16_matmul_softmax"] = dict(ms=ref)
    print(name, "torch bf16 matmul+softmax ms", ref, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/perf_match.json", "w"), indent=1)
