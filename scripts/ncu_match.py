"""ncu target: scoring at config #5 (262144 x 501 x 512) / config #3 (64000 x 66 x 768), argv[1] = cfg5|cfg3, argv[2] = probs|top."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cvpr22_cross_modal_pseudo_labeling_b200.layers import embed_match_softmax
r, c, d = (262144, 501, 512) if (len(sys.argv) < 2 or sys.argv[1] == "cfg5") else (64000, 66, 768)
probs = len(sys.argv) > 2 and sys.argv[2] == "probs"
A = (torch.randn((r, d), device="cuda") * 3).to(torch.bfloat16)
E = torch.nn.functional.normalize(torch.randn((c, d), device="cuda"), dim=-1).to(torch.bfloat16)
for _ in range(3):
    embed_match_softmax(A, E, 0.05, want_probs=probs)
torch.cuda.synchronize()
