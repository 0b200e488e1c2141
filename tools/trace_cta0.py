"""Dump the event trace of CTA 0 of the fused blockmix kernel (debug instrumentation) -> gpurun_out/trace_cta0*.csv"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhla_b200  # noqa: E402
from mhla_b200 import _capi  # noqa: E402

normalize = "--no-normalize" not in sys.argv
tag = "norm" if normalize else "nonorm"
if os.environ.get("MHLA_TRACE_CTA"):
    tag += "_cta" + os.environ["MHLA_TRACE_CTA"]
kw = {}
if "--p1only" in sys.argv:
    kw = dict(debug_flags=_capi.FLAG_STOP_AFTER_P1)
    tag += "_p1only"
if "--p2only" in sys.argv:
    kw = dict(debug_flags=_capi.FLAG_ONLY_P2)
    tag += "_p2only"
if "--p3only" in sys.argv:
    kw = dict(debug_flags=_capi.FLAG_ONLY_P3)
    tag += "_p3only"
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
    mhla_b200.mhla(q, k, v, W, normalize=normalize, out=out, **kw)
torch.cuda.synchronize()
prof = torch.zeros(148 * 16 + 4 * 256 * 4, dtype=torch.int64, device=dev)
L.mhla_debug_set_profile_buffer(prof.data_ptr())
mhla_b200.mhla(q, k, v, W, normalize=normalize, out=out, **kw)
torch.cuda.synchronize()
L.mhla_debug_set_profile_buffer(None)
tr = prof[148 * 16:].cpu().view(4, 256, 4)
t0 = int(tr[0, 0, 0])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"trace_cta0_{tag}.csv"), "w") as f:
    f.write("item,type,prod_start,prod_dep_ok,prod_end,mma_start,mma_tempty_ok,mma_commit,epi_start,epi_tfull,epi_end,st_start,st_issued,st_done\n")
    for i in range(256):
        if int(tr[0, i, 0]) == 0:
            break
        rel = lambda x: (int(x) - t0) if int(x) else -1  # noqa: E731
        f.write(",".join(str(x) for x in [
            i, int(tr[0, i, 3]), rel(tr[0, i, 0]), rel(tr[0, i, 2]), rel(tr[0, i, 1]),
            rel(tr[1, i, 0]), rel(tr[1, i, 1]), rel(tr[1, i, 2]),
            rel(tr[2, i, 0]), rel(tr[2, i, 1]), rel(tr[2, i, 2]),
            rel(tr[3, i, 0]), rel(tr[3, i, 1]), rel(tr[3, i, 2])]) + "\n")
print("wrote trace", tag)
