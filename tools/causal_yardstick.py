"""Yardstick for the causal tolerance: how far is the REFERENCE's own chunk operator
(mhla_nlp/fla/ops/mhla/naive.py:10-83, loaded from /root/reference - build container only) from the fp32 oracle when it
runs the way the trainers run it, i.e. under bf16 autocast, on the same bf16-rounded inputs the CUDA kernel gets?
Prints RMS error ratio and max-abs error / max|ref| per test shape -> profiles/r02_causal_yardstick.log."""
import importlib.util
import os
import sys

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_naive", "/root/reference/mhla_nlp/fla/ops/mhla/naive.py")
nv = importlib.util.module_from_spec(spec)
spec.loader.exec_module(nv)
torch.set_grad_enabled(False)

for (B, T, H, K, V, signed, init) in [(1, 1024, 4, 64, 64, True, False), (2, 2048, 4, 128, 256, False, True),
                                      (1, 256, 2, 64, 128, True, False), (2, 200, 2, 64, 64, False, False),
                                      (1, 48, 2, 128, 128, True, True)]:
    g = torch.Generator().manual_seed(7)
    q, k = torch.randn(B, T, H, K, generator=g), torch.randn(B, T, H, K, generator=g)
    if not signed:
        q, k = torch.relu(q), torch.relu(k)
    v = torch.randn(B, T, H, V, generator=g)
    q, k, v = q.bfloat16(), k.bfloat16(), v.bfloat16()
    L = 32
    mm = (torch.tril(torch.ones(L, L)) / (torch.arange(L, dtype=torch.float32).unsqueeze(1) + 1.0)) if init else \
        torch.clamp(torch.rand(L, L, generator=torch.Generator().manual_seed(3)), 1e-5, 1).tril()
    ref = oracle.causal_chunk_fwd(q.float(), k.float(), v.float(), mm)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        o = nv.naive_chunk_simple_mhla_fixed(q, k, v, mm.view(L, L, 1, 1, 1, 1)).float()
    rms = oracle.err_ratio(ref, o)
    mx = float((ref - o).abs().max() / ref.abs().max())
    print(f"B={B} T={T} H={H} K={K} V={V} signed={signed} init_mm={init}: reference under bf16 autocast vs fp32 oracle: "
          f"rms {rms:.3e}  max-abs/max|ref| {mx:.3e}")
