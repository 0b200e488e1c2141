"""One call of a named kernel case for ncu captures (run under ncu on the GPU box):
   headline | smalln | causal | wan3d | wanprep"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhla_b200  # noqa: E402

case = sys.argv[1]
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
calls = int(os.environ.get("NCU_CALLS", "2"))
rnd = lambda *s: torch.randn(*s, generator=g, device=dev)  # noqa: E731
if case == "headline" or case == "smalln":
    B, H, M, w, D = (2, 16, 128, 256, 64) if case == "headline" else (64, 6, 16, 16, 64)
    q, k, v = (rnd(B, H, M, w, D).relu() + 1e-6).bfloat16(), (rnd(B, H, M, w, D).relu() + 1e-6).bfloat16(), rnd(B, H, M, w, D).bfloat16()
    W = torch.rand(M, M, device=dev) / M
    out = torch.empty_like(q)
    fn = lambda: mhla_b200.mhla(q, k, v, W, normalize=True, out=out)  # noqa: E731
elif case == "causal":
    B, T, H, K, V = 8, 2048, 4, 128, 256
    q, k, v = rnd(B, T, H, K).bfloat16(), rnd(B, T, H, K).bfloat16(), rnd(B, T, H, V).bfloat16()
    mm = torch.clamp(torch.rand(32, 32, device=dev), 1e-5, 1).tril()
    fn = lambda: mhla_b200.naive_chunk_simple_mhla_fixed(q, k, v, mm)  # noqa: E731
elif case == "wan3d":
    B, nh, D, grid, layout = 1, 12, 128, (21, 30, 50), (3, 5, 10)
    N = grid[0] * grid[1] * grid[2]
    q, k, v = rnd(B, N, nh, D).bfloat16(), rnd(B, N, nh, D).bfloat16(), rnd(B, N, nh, D).bfloat16()
    W = torch.rand(150, 150, device=dev) / 150
    fn = lambda: mhla_b200.mhla_blockmix_grid(q, k, v, W, grid, layout, normalize=False)  # noqa: E731
elif case == "wanprep":
    from mhla_b200.modules.wan import _rope_tables
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    B, N, C, D = 1, 31500, 1536, 128
    xq, xk = rnd(B, N, C).bfloat16(), rnd(B, N, C).bfloat16()
    w = torch.ones(C, device=dev)
    cos, sin = torch.rand(N, D // 2, device=dev), torch.rand(N, D // 2, device=dev)
    fn = lambda: mhla_b200.wan_prep(xq, xk, w, w, cos, sin, D)  # noqa: E731
else:
    raise SystemExit(case)
for _ in range(calls):
    fn()
torch.cuda.synchronize()
print("done", case)
