"""Repeat the Wan-shaped (rope + normaliser) call many times and count mismatches against the first result and the oracle."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import mhla_b200, oracle
B, H, M, w, D = 2, 12, 150, 210, 128
g = torch.Generator(device="cuda").manual_seed(3)
mk = lambda relu: (torch.relu(torch.randn(B, H, M, w, D, generator=g, device="cuda")) + 1e-6 if relu else torch.randn(B, H, M, w, D, generator=g, device="cuda")).bfloat16()
q, k, v, qr, kr = mk(True), mk(True), mk(False), mk(False), mk(False)
W = (torch.rand(M, M, generator=torch.Generator().manual_seed(4)) / M + 0.5 * torch.eye(M) / M).cuda()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
kw = {sys.argv[2]: True} if len(sys.argv) > 2 and sys.argv[2] != "fused" else {"force_fused": True}
norm = not (len(sys.argv) > 3 and sys.argv[3] == "nonorm")
first = None; bad = 0
for i in range(n):
    out = mhla_b200.mhla(q, k, v, W, q_rope=qr, k_rope=kr, normalize=norm, **kw)
    if first is None:
        first = out.clone()
        ref = oracle.blockmix_fwd(q[0, 0][None].cpu(), k[0, 0][None].cpu(), v[0, 0][None].cpu(), W.cpu(), normalize=norm, q_rope=qr[0, 0][None].cpu(), k_rope=kr[0, 0][None].cpu())
        print("err vs oracle (unit 0,0):", oracle.err_ratio(ref[0], out[0, 0].float().cpu()))
    elif not torch.equal(out, first):
        bad += 1
        if bad <= 3:
            d = (out.float() - first.float()).abs()
            print("mismatch at iter", i, "max", float(d.max()), "n", int((d > 0).sum()), "units", sorted(set((d.flatten(2).amax(-1) > 0).nonzero()[:, :2].flatten().tolist()))[:8])
torch.cuda.synchronize()
print(f"{kw} norm={norm}: {bad} of {n - 1} repeats differ from the first result")
