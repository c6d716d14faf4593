"""ncu target: the tcgen05 kernels at the BASELINE sizes.  argv[1] = cfg5 | cfg3 | lvis | linear, argv[2] = probs | top."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cvpr22_cross_modal_pseudo_labeling_b200.layers import embed_match_softmax, linear_bf16
which = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
probs = len(sys.argv) > 2 and sys.argv[2] == "probs"
if which == "linear":
    x = torch.randn((64000, 1024), device="cuda").to(torch.bfloat16)
    w = (torch.randn((768, 1024), device="cuda") * 0.02).to(torch.bfloat16)
    b = torch.zeros((768,), device="cuda")
    for _ in range(3):
        linear_bf16(x, w, b, want_f32=False, want_bf16=True)
else:
    r, c, d = {"cfg5": (262144, 501, 512), "cfg3": (64000, 66, 768), "lvis": (32768, 1203, 768)}[which]
    A = (torch.randn((r, d), device="cuda") * 3).to(torch.bfloat16)
    E = torch.nn.functional.normalize(torch.randn((c, d), device="cuda"), dim=-1).to(torch.bfloat16)
    for _ in range(3):
        embed_match_softmax(A, E, 0.05, want_probs=probs)
torch.cuda.synchronize()
