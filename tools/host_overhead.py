"""Host-side cost of one mhla() call: a tiny problem whose kernel takes ~10 us, timed over a long loop (GPU box)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import mhla_b200
B, H, M, w, D = 1, 4, 16, 64, 64
dev = "cuda"
q = torch.rand(B, H, M, w, D, device=dev).bfloat16(); k = torch.rand_like(q); v = torch.rand_like(q)
W = torch.rand(M, M, device=dev) / M
out = torch.empty_like(q)
for _ in range(50):
    mhla_b200.mhla(q, k, v, W, out=out)
torch.cuda.synchronize()
n = 2000
t0 = time.perf_counter()
for _ in range(n):
    mhla_b200.mhla(q, k, v, W, out=out)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e6 * (t1 - t0) / n:.1f} us/call, incl. drain {1e6 * (t2 - t0) / n:.1f} us/call")
