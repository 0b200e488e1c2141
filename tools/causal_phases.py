"""Per-phase device time of the causal path (three phase launches): summaries, + mixing, + readout (CUDA-graph timing)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mhla_b200 import _capi, ops  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, B, T, H, K, V in [("cfg3 NLP 340M B=8 T=2048 H=4 K=128 V=256", 8, 2048, 4, 128, 256),
                            ("B=8 T=2048 H=16 K=64 V=64", 8, 2048, 16, 64, 64),
                            ("cfg3 B=2", 2, 2048, 4, 128, 256)]:
    q = torch.randn(B, T, H, K, generator=g, device=dev).bfloat16()
    k = torch.randn(B, T, H, K, generator=g, device=dev).bfloat16()
    v = torch.randn(B, T, H, V, generator=g, device=dev).bfloat16()
    mm = torch.clamp(torch.rand(32, 32, device=dev), 1e-5, 1).tril()
    t1 = timed(lambda: ops._causal_fwd(q, k, v, mm, debug_flags=_capi.FLAG_STOP_AFTER_P1))
    t2 = timed(lambda: ops._causal_fwd(q, k, v, mm, debug_flags=_capi.FLAG_STOP_AFTER_P2))
    t3 = timed(lambda: ops._causal_fwd(q, k, v, mm, unfused=True))
    tf = timed(lambda: ops._causal_fwd(q, k, v, mm, unfused=False))
    alg = (2 * q.numel() + 2 * v.numel()) * 2
    print(f"{name}: prologue+P1 {t1:.1f} us, +P2 {t2:.1f} us, +P3 {t3:.1f} us (three launches), single kernel {tf:.1f} us; "
          f"algorithmic {alg / 1e6:.0f} MB -> {alg / t3 / 1e3:.0f} GB/s")
