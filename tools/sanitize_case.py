"""A few calls of small shapes for compute-sanitizer (memcheck / synccheck / racecheck): fused and three-launch paths,
rope + normaliser included.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhla_b200  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
cases = [(1, 2, 6, 210, 128, True, True), (1, 2, 4, 256, 64, True, False), (2, 3, 16, 16, 64, True, False),
         (1, 1, 150, 48, 64, True, False)]
if len(sys.argv) > 1:
    cases = [cases[int(sys.argv[1])]]
for (B, H, M, w, D, norm, rope) in cases:
    mk = lambda: torch.randn(B, H, M, w, D, generator=g, device=dev).bfloat16()  # noqa: E731
    q, k, v = mk().relu() + 1e-6, mk().relu() + 1e-6, mk()
    qr, kr = (mk(), mk()) if rope else (None, None)
    W = torch.rand(M, M, device=dev) / M
    for kw in ({}, {"three_launch": True}):
        for _ in range(2):
            o = mhla_b200.mhla(q, k, v, W, q_rope=qr, k_rope=kr, normalize=norm, **kw)
        torch.cuda.synchronize()
        print("ok", (B, H, M, w, D, norm, rope), kw or "fused", float(o.float().abs().mean()), flush=True)
if "--causal" in sys.argv:
    q = torch.randn(1, 256, 2, 64, generator=g, device=dev).bfloat16()
    mm = torch.clamp(torch.rand(32, 32, device=dev), 1e-5, 1).tril()
    o = mhla_b200.mhla_causal(q, q, q, mm)
    torch.cuda.synchronize()
    print("ok causal", float(o.float().abs().mean()))
