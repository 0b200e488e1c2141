"""One blockmix call at the headline shape for ncu captures (run under ncu on the GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhla_b200  # noqa: E402

normalize = "--no-normalize" not in sys.argv
kw = {}
if "--fused" in sys.argv:
    kw["fused"] = True
if "--three" in sys.argv:
    kw["three_launch"] = True
B, H, M, w, D = 2, 16, 128, 256, 64
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
q = (torch.relu(torch.randn(B, H, M, w, D, generator=g, device=dev)) + 1e-6).bfloat16()
k = (torch.relu(torch.randn(B, H, M, w, D, generator=g, device=dev)) + 1e-6).bfloat16()
v = torch.randn(B, H, M, w, D, generator=g, device=dev).bfloat16()
W = torch.rand(M, M, device=dev) / M
out = torch.empty_like(q)
for _ in range(int(os.environ.get("NCU_CALLS", "2"))):
    mhla_b200.mhla(q, k, v, W, normalize=normalize, out=out, **kw)
torch.cuda.synchronize()
print("done")
