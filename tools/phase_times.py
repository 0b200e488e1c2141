"""Per-phase CUDA-event timings of the blockmix path at the headline shape (run on the GPU box).

    python tools/phase_times.py [--no-normalize] [--reps 30]

Each phase is launched alone through the debugging flags of the C ABI (STOP_AFTER_P1 / ONLY_P2 / ONLY_P3) on a
workspace that a full run has filled, with the inputs (537 MB) far larger than L2, so every number is an
HBM-cold steady-state figure.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhla_b200  # noqa: E402
from mhla_b200 import _capi  # noqa: E402

reps = 30
if "--reps" in sys.argv:
    reps = int(sys.argv[sys.argv.index("--reps") + 1])
B, H, M, w, D = 2, 16, 128, 256, 64
if "--shape" in sys.argv:
    B, H, M, w, D = [int(x) for x in sys.argv[sys.argv.index("--shape") + 1].split(",")]
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
q = (torch.relu(torch.randn(B, H, M, w, D, generator=g, device=dev)) + 1e-6).bfloat16()
k = (torch.relu(torch.randn(B, H, M, w, D, generator=g, device=dev)) + 1e-6).bfloat16()
v = torch.randn(B, H, M, w, D, generator=g, device=dev).bfloat16()
W = torch.rand(M, M, device=dev) / M
out = torch.empty_like(q)
alg_bytes = 4 * q.numel() * 2


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for normalize in ([False] if "--no-normalize" in sys.argv else [True, False]):
    res = {}
    for name, kw in (("full", {}), ("three", {"three_launch": True}), ("P1", {"debug_flags": _capi.FLAG_STOP_AFTER_P1}),
                     ("P2", {"debug_flags": _capi.FLAG_ONLY_P2}), ("P3", {"debug_flags": _capi.FLAG_ONLY_P3}),
                     ):
        try:
            res[name] = timed(lambda: mhla_b200.mhla(q, k, v, W, normalize=normalize, out=out, **kw))
        except Exception as e:  # noqa: BLE001
            res[name] = float("nan")
            print("phase", name, "failed:", e)
    print(f"normalize={int(normalize)} shape={B},{H},{M},{w},{D}: " +
          "  ".join(f"{n}={t:.1f}us" for n, t in res.items()) +
          f"   full: {alg_bytes / res['full'] / 1e3:.0f} GB/s algorithmic")
