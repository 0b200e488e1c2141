import os, sys, torch
sys.path.insert(0, "/root/repo")
import mhla_b200, oracle
g = torch.Generator().manual_seed(0)
for (B,T,H,K,V) in [(2,2048,4,64,64),(8,2048,4,128,256),(8,2048,16,64,64)]:
    q = torch.randn(B,T,H,K,generator=g).bfloat16(); k = torch.randn(B,T,H,K,generator=g).bfloat16(); v = torch.randn(B,T,H,V,generator=g).bfloat16()
    mm = torch.clamp(torch.rand(32,32,generator=g),1e-5,1).tril()
    for unf in (True, False):
        o = mhla_b200.mhla_causal(q.cuda(),k.cuda(),v.cuda(),mm.cuda(), unfused=unf)
        torch.cuda.synchronize()
        ref = oracle.causal_chunk_fwd(q[:1].float(),k[:1].float(),v[:1].float(),mm)
        print((B,T,H,K,V), "unfused" if unf else "fused", "err", oracle.err_ratio(ref, o[:1].float().cpu()))
