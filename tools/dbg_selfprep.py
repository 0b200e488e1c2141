import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import mhla_b200, oracle
from mhla_b200 import ops, _capi
B,H,M,w,D = 1,4,16,64,64
g = torch.Generator().manual_seed(0)
q = (torch.relu(torch.randn(B,H,M,w,D,generator=g))+1e-6).bfloat16(); k = (torch.relu(torch.randn(B,H,M,w,D,generator=g))+1e-6).bfloat16(); v = torch.randn(B,H,M,w,D,generator=g).bfloat16()
W = torch.rand(M,M,generator=g)/M
ref = oracle.blockmix_fwd(q,k,v,W,normalize=True)
for it in range(3):
    out = mhla_b200.mhla(q.cuda(),k.cuda(),v.cuda(),W.cuda(),normalize=True)
    torch.cuda.synchronize()
    print("call", it, "err", oracle.err_ratio(ref, out.float().cpu()), "nan", bool(torch.isnan(out).any()))
    for key, ws in ops._WS_CACHE.items():
        nbytes = key[2]
        d = _capi.BlockmixDesc(); d.B,d.H,d.M,d.w,d.D = B,H,M,w,D; d.dtype=0; d.flags=1
        lay = (C.c_size_t*8)(); _capi.lib().mhla_blockmix_workspace_layout(C.byref(d), C.byref(lay))
        offW, offC, Mp = lay[3], lay[4], lay[7]
        base = (ws.data_ptr()+1023)//1024*1024 - ws.data_ptr()
        wsv = ws[base:]
        Wp = wsv[offW:offW+2*M*Mp*2].view(torch.bfloat16).view(2,M,Mp).float().sum(0).cpu()
        print("   Wp err", oracle.err_ratio(W, Wp[:,:M]), "ctrl nonzero words:", int((wsv[offC:offC+(2*B*H*16+128)*4].view(torch.int32)!=0).sum()))
