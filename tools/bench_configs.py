"""Kernel-only timings (CUDA events, inputs resident in HBM) of the BASELINE configs other than the headline - parity for
these shapes is in tests/; this prints algorithmic GB/s next to the measured copy peak.  GPU box only."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhla_b200  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)


def timed(fn, reps=20):
    """Device time per call: the calls are captured into ONE CUDA graph and the graph is replayed, so the Python / ctypes
    enqueue cost (~20 us per call, longer than the short kernels) is not in the number.  Falls back to eager enqueue."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(reps):
                fn()
        gr.replay()
        torch.cuda.synchronize()
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3
    except Exception as e:   # noqa: BLE001
        print("graph capture failed, eager timing:", repr(e)[:200], file=sys.stderr)
        torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


rows = []
for name, B, H, M, w, D, normalize, rope in [
    ("cfg1 B=1 H=4 N=1024 D=64", 1, 4, 16, 64, 64, True, False),
    ("cfg2 DiT-S/2 N=256 batch 2", 2, 6, 16, 16, 64, True, False),
    ("cfg2 DiT-S/2 N=256 batch 64", 64, 6, 16, 16, 64, True, False),
    ("cfg2 DiT-S/2 N=256 batch 256", 256, 6, 16, 16, 64, True, False),
    ("ViT-S 14x14 tokens (M=4, w=49) batch 256", 256, 6, 4, 49, 64, True, False),
    ("cfg4 Wan2.1-1.3B N=31500 B=1 (shipped: no normaliser)", 1, 12, 150, 210, 128, False, True),
    ("cfg4 Wan2.1-1.3B N=31500 B=2 normaliser on", 2, 12, 150, 210, 128, True, True),
    ("cfg5 N=8192 B=2 H=16 D=64", 2, 16, 32, 256, 64, True, False),
    ("cfg5 N=32768 B=2 H=16 D=64 (headline)", 2, 16, 128, 256, 64, True, False),
    ("cfg5 N=131072 B=2 H=16 D=64", 2, 16, 512, 256, 64, True, False),
]:
    mk = lambda: torch.randn(B, H, M, w, D, generator=g, device=dev).bfloat16()  # noqa: E731
    q, k, v = mk().relu() + 1e-6, mk().relu() + 1e-6, mk()
    qr, kr = (mk(), mk()) if rope else (None, None)
    W = torch.rand(M, M, device=dev) / M
    out = torch.empty_like(q)
    t = timed(lambda: mhla_b200.mhla(q, k, v, W, q_rope=qr, k_rope=kr, normalize=normalize, out=out))
    nbytes = 4 * q.numel() * 2
    rows.append({"config": name, "us": t * 1e6, "tokens_per_s": B * M * w / t, "algorithmic_GBps": nbytes / t / 1e9})
    print(json.dumps(rows[-1]))
for name, B, T, H, K, V in [("cfg3 NLP 340M T=2048 K=128 V=256 B=8", 8, 2048, 4, 128, 256),
                            ("causal T=2048 K=64 V=64 B=8 H=16", 8, 2048, 16, 64, 64)]:
    q = torch.randn(B, T, H, K, generator=g, device=dev).bfloat16()
    k = torch.randn(B, T, H, K, generator=g, device=dev).bfloat16()
    v = torch.randn(B, T, H, V, generator=g, device=dev).bfloat16()
    mm = torch.clamp(torch.rand(32, 32, device=dev), 1e-5, 1).tril()
    t = timed(lambda: mhla_b200.naive_chunk_simple_mhla_fixed(q, k, v, mm))
    nbytes = (2 * q.numel() + 2 * v.numel()) * 2
    rows.append({"config": name, "us": t * 1e6, "tokens_per_s": B * T / t, "algorithmic_GBps": nbytes / t / 1e9})
    print(json.dumps(rows[-1]))
