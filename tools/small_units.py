"""Per-rank kernel time of the strong-scaling shards on ONE GPU: 32 / 16 / 8 / 4 (b,h) units of the headline shape, eager
back-to-back launches vs one CUDA graph of 20 launches (PDL edges kept) vs 20 single-launch graph replays."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhla_b200  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
M, w, D = 128, 256, 64
W = mhla_b200.block_distance_matrix((M, 1, 1), "linear").to(dev)
for U in (32, 16, 8, 4):
    nsets = max(1, -(-int(380e6) // (4 * U * M * w * D * 2)))
    sets = []
    for _ in range(nsets):
        q = (torch.randn(U, M, w, D, generator=g, device=dev).relu() + 1e-6).bfloat16()
        k = (torch.randn(U, M, w, D, generator=g, device=dev).relu() + 1e-6).bfloat16()
        v = torch.randn(U, M, w, D, generator=g, device=dev).bfloat16()
        sets.append((q, k, v, torch.empty_like(q)))
    call = lambda i: mhla_b200.mhla(sets[i % nsets][0], sets[i % nsets][1], sets[i % nsets][2], W, out=sets[i % nsets][3])  # noqa: E731
    for i in range(nsets + 3):
        call(i)
    torch.cuda.synchronize()
    K = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        call(i)
    e1.record()
    torch.cuda.synchronize()
    t_eager = e0.elapsed_time(e1) / K * 1e3
    big = torch.cuda.CUDAGraph()
    with torch.cuda.graph(big):
        for i in range(K):
            call(i)
    big.replay()
    torch.cuda.synchronize()
    e0.record()
    big.replay()
    e1.record()
    torch.cuda.synchronize()
    t_big = e0.elapsed_time(e1) / K * 1e3
    singles = []
    for i in range(nsets):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            call(i)
        singles.append(gr)
    torch.cuda.synchronize()
    e0.record()
    for i in range(K):
        singles[i % nsets].replay()
    e1.record()
    torch.cuda.synchronize()
    t_single = e0.elapsed_time(e1) / K * 1e3
    alg = 4 * U * M * w * D * 2
    print(f"{U:2d} units ({nsets} rotating sets): eager {t_eager:6.1f} us | one graph of {K} launches {t_big:6.1f} us "
          f"({alg / t_big / 1e3:.0f} GB/s) | {K} single-launch graph replays {t_single:6.1f} us")
