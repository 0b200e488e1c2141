"""Per-role wait-cycle breakdown of the fused blockmix kernel (debug instrumentation, run on the GPU box)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhla_b200  # noqa: E402
from mhla_b200 import _capi  # noqa: E402

normalize = "--no-normalize" not in sys.argv
KW = {}
for m, f in (("--p1only", _capi.FLAG_STOP_AFTER_P1), ("--p2only", _capi.FLAG_ONLY_P2), ("--p3only", _capi.FLAG_ONLY_P3)):
    if m in sys.argv:
        KW = dict(debug_flags=f)
B, H, M, w, D = 2, 16, 128, 256, 64
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
q = (torch.relu(torch.randn(B, H, M, w, D, generator=g, device=dev)) + 1e-6).bfloat16()
k = (torch.relu(torch.randn(B, H, M, w, D, generator=g, device=dev)) + 1e-6).bfloat16()
v = torch.randn(B, H, M, w, D, generator=g, device=dev).bfloat16()
W = torch.rand(M, M, device=dev) / M
out = torch.empty_like(q)
L = _capi.lib()
L.mhla_debug_set_profile_buffer.argtypes = [C.c_void_p]
for _ in range(3):
    mhla_b200.mhla(q, k, v, W, normalize=normalize, out=out, **KW)
torch.cuda.synchronize()
prof = torch.zeros(148, 16, dtype=torch.int64, device=dev)
L.mhla_debug_set_profile_buffer(prof.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
mhla_b200.mhla(q, k, v, W, normalize=normalize, out=out, **KW)
e1.record()
torch.cuda.synchronize()
L.mhla_debug_set_profile_buffer(None)
p = prof.cpu().double()
names = ["prod.wait_empty", "prod.wait_dep", "prod.total", "mma.wait_full", "mma.wait_tempty", "epi.wait_tfull",
         "epi.wait_sfree", "epi.t_ld+pack", "epi.t_P1", "epi.t_P2", "epi.t_P3", "epi.items", "epi.t_tmem_ld",
         "epi.t_flush", "gt", "epi.t_stage"]
print(f"normalize={normalize}  step (events) = {e0.elapsed_time(e1) * 1e3:.1f} us")
print("SM clock (GHz) from clock64/globaltimer over the producer lifetime: mean %.3f min %.3f max %.3f ; lifetime us mean %.1f" % ((p[:,2]/p[:,14]).mean(), (p[:,2]/p[:,14]).min(), (p[:,2]/p[:,14]).max(), p[:,14].mean()/1e3))
tot = p[:, 2].mean()
for i, n in enumerate(names):
    col = p[:, i]
    print(f"{n:18s} mean {col.mean():12.0f}  min {col.min():12.0f}  max {col.max():12.0f}   ({100 * col.mean() / tot:5.1f}% of producer lifetime)")
if "--per-cta" in sys.argv:
    # one row per CTA, sorted by the number of items its warpgroup 0 handled (the dynamic tickets give faster SMs more)
    order = sorted(range(148), key=lambda c: float(p[c, 11]))
    print("cta items(wg0) prod.wait_empty prod.wait_dep mma.wait_full mma.wait_tempty epi.wait_tfull  (cycles)")
    for c in order:
        print(f"{c:3d} {int(p[c, 11]):3d} {int(p[c, 0]):8d} {int(p[c, 1]):8d} {int(p[c, 3]):8d} {int(p[c, 4]):8d} {int(p[c, 5]):8d}")
    slow, fast = order[0], order[-1]
    print(f"slowest CTA {slow}, fastest CTA {fast}")
    with open(os.path.join(ROOT, "gpurun_out", "slow_fast_cta.txt"), "w") as f:
        f.write(f"{slow} {fast}\n")
