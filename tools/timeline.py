"""Per-launch timeline of the blockmix path (debug instrumentation, GPU box): when does each CTA of each launch start and
finish (globaltimer), i.e. how long are the ramps, tails and gaps between the PDL-chained launches."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhla_b200  # noqa: E402
from mhla_b200 import _capi  # noqa: E402

normalize = "--no-normalize" not in sys.argv
kw = {}
if "--three" in sys.argv:
    kw = dict(three_launch=True)
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
NTL = 6 * 148 * 4
prof = torch.zeros(148 * 16 + 4 * 256 * 4 + NTL, dtype=torch.int64, device=dev)
L.mhla_debug_set_profile_buffer(prof.data_ptr())
mhla_b200.mhla(q, k, v, W, normalize=normalize, out=out, **kw)
torch.cuda.synchronize()
L.mhla_debug_set_profile_buffer(None)
tl = prof[148 * 16 + 4 * 256 * 4:].cpu().view(6, 148, 4)
t0 = int(tl[tl > 0].min())
print(f"normalize={normalize} {kw}")
for mode in range(6):
    st, en = tl[mode, :, 0], tl[mode, :, 1]
    m = st > 0
    if not m.any():
        continue
    st, en = (st[m] - t0).double() / 1e3, (en[m] - t0).double() / 1e3
    x1, x2 = tl[mode, :, 2][m], tl[mode, :, 3][m]
    if (x1 > 0).any():
        x1 = (x1[x1 > 0] - t0).double() / 1e3
        print(f"        P1 tickets exhausted: min {x1.min():7.1f} median {x1.median():7.1f} max {x1.max():7.1f} us")
    if (x2 > 0).any():
        x2 = (x2[x2 > 0] - t0).double() / 1e3
        print(f"        P2 tickets exhausted: min {x2.min():7.1f} median {x2.median():7.1f} max {x2.max():7.1f} us")
    print(f"mode {mode}: CTAs {int(m.sum()):3d}  start min {st.min():7.1f} max {st.max():7.1f} us | end min {en.min():7.1f} "
          f"median {en.median():7.1f} max {en.max():7.1f} us | busy mean {(en - st).mean():6.1f} us")
