"""Fused vs three-launch timing of the causal kernel at the NLP shapes (GPU box)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import mhla_b200
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (B, T, H, K, V) in [(8, 2048, 4, 128, 256), (8, 2048, 16, 64, 64), (2, 2048, 4, 128, 256), (32, 2048, 4, 128, 256)]:
    q = torch.randn(B, T, H, K, generator=g, device=dev).bfloat16(); k = torch.randn(B, T, H, K, generator=g, device=dev).bfloat16()
    v = torch.randn(B, T, H, V, generator=g, device=dev).bfloat16()
    mm = torch.clamp(torch.rand(32, 32, device=dev), 1e-5, 1).tril()
    t0 = timed(lambda: mhla_b200.mhla_causal(q, k, v, mm, unfused=False))
    t1 = timed(lambda: mhla_b200.mhla_causal(q, k, v, mm, unfused=True))
    nb = (2 * q.numel() + 2 * v.numel()) * 2
    print(f"B={B} T={T} H={H} K={K} V={V}: fused {t0:.1f} us ({nb / t0 / 1e3:.0f} GB/s)  three-launch {t1:.1f} us ({nb / t1 / 1e3:.0f} GB/s)")
