"""Embedding-head kernel timing: scoring at BASELINE configs #3 (64k x 768 x 66, HBM-bound) and #5
(262k x 512 x 501, tensor-bound), the LVIS-wide case, and the emb_pred projection (b200_linear_bf16);
persistent kernel (tc_gemm.cu) next to the one-CTA-per-tile kernel (embed_match.cu) and torch/cuBLAS."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
from cvpr22_cross_modal_pseudo_labeling_b200.layers import embed_match_softmax, linear_bf16

PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
TF, GB = float(PEAKS.get("bf16_tflops_sustained", 1402.8)), float(PEAKS.get("hbm_gbs", 6548.5))


def timeit(fn, warm=3, iters=10, reps=8):
    """Median device time of one call: `reps` calls captured in a CUDA graph (no launch gaps), replayed."""
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); g.replay(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) / reps for a, b in evs)
    return ts[len(ts) // 2]

res = {}
for name, (r, c, d) in {"cfg3_16000x66x768": (16000, 66, 768), "cfg3_64000x66x768": (64000, 66, 768),
                        "cfg5_262144x501x512": (262144, 501, 512), "lvis_32768x1203x768": (32768, 1203, 768)}.items():
    A = (torch.randn((r, d), device="cuda") * 3).to(torch.bfloat16)
    E = torch.nn.functional.normalize(torch.randn((c, d), device="cuda"), dim=-1).to(torch.bfloat16)
    for mode, kw in (("top_only", dict(want_probs=False)), ("probs", dict(want_probs=True))):
        for legacy in (False, True):
            _ext.debug_match(legacy)
            if legacy and c > 512:
                kw = dict(kw, want_logits=True)   # the legacy wide path needs the logits buffer as scratch
            ms = timeit(lambda: embed_match_softmax(A, E, 0.05, **kw))
            _ext.debug_match(False)
            flops = 2.0 * r * c * d
            byts = r * d * 2 + c * d * 2 + r * 8 + (r * c * 4 if mode == "probs" else 0)
            key = "%s_%s_%s" % (name, mode, "legacy" if legacy else "persistent")
            res[key] = dict(ms=ms, tflops=flops / ms / 1e9, frac_tensor=flops / ms / 1e9 / TF, gbs=byts / ms / 1e6, frac_hbm=byts / ms / 1e6 / GB)
            print(key, json.dumps(res[key]), flush=True)
    ref = timeit(lambda: torch.softmax(A @ E.t(), -1), iters=5, reps=2)
    res[name + "_torch_bf16_matmul_softmax"] = dict(ms=ref)
    print(name, "torch bf16 matmul+softmax ms", ref, flush=True)
for name, (r, n, k) in {"emb_pred_16000x768x1024": (16000, 768, 1024), "emb_pred_64000x768x1024": (64000, 768, 1024),
                        "emb_pred_64000x768x2048": (64000, 768, 2048), "head_stub_16000x768x256": (16000, 768, 256)}.items():
    x = torch.randn((r, k), device="cuda").to(torch.bfloat16)
    w = (torch.randn((n, k), device="cuda") * 0.02).to(torch.bfloat16)
    b = torch.randn((n,), device="cuda")
    ms = timeit(lambda: linear_bf16(x, w, b, want_f32=False, want_bf16=True))
    ref = timeit(lambda: torch.nn.functional.linear(x, w, b.to(torch.bfloat16)))
    flops = 2.0 * r * n * k
    res[name] = dict(ms=ms, tflops=flops / ms / 1e9, frac_tensor=flops / ms / 1e9 / TF, torch_ms=ref, torch_tflops=flops / ref / 1e9)
    print(name, json.dumps(res[name]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/perf_match.json", "w"), indent=1)
