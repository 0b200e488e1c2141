"""Per-phase cycle breakdown of the short-sequence kernel's warpgroup 0 (debug instrumentation, GPU box)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhla_b200  # noqa: E402
from mhla_b200 import _capi  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H, M, w, D = 6, 16, 16, 64
dev = torch.device("cuda")
q = (torch.rand(B, H, M, w, D, device=dev) + 1e-6).bfloat16()
k, v = torch.rand_like(q), torch.randn(B, H, M, w, D, device=dev).bfloat16()
W = torch.rand(M, M, device=dev) / M
out = torch.empty_like(q)
L = _capi.lib()
L.mhla_debug_set_profile_buffer.argtypes = [C.c_void_p]
for _ in range(3):
    mhla_b200.mhla(q, k, v, W, out=out)
torch.cuda.synchronize()
prof = torch.zeros(148, 16, dtype=torch.int64, device=dev)
L.mhla_debug_set_profile_buffer(prof.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
mhla_b200.mhla(q, k, v, W, out=out)
e1.record()
torch.cuda.synchronize()
L.mhla_debug_set_profile_buffer(None)
p = prof.cpu().double()
p = p[p[:, 6] > 0]
names = ["wait scores (load + QK^T)", "masking epilogue", "tile barrier + normaliser", "wait O = A V", "readout epilogue", "store hand-off"]
print(f"batch {B}: {B * H} units, step {e0.elapsed_time(e1) * 1e3:.1f} us, units per CTA mean {p[:, 6].mean():.2f}")
for i, n in enumerate(names):
    print(f"  {n:28s} {p[:, i].sum() / p[:, 6].sum():9.0f} cycles per unit")
print(f"  {'total':28s} {p[:, :6].sum() / p[:, 6].sum():9.0f} cycles per unit")
