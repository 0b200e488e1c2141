"""Bring-up diagnostics for the blockmix kernel (run on the GPU box):  python tools/gpu_diag.py [case ...]

Each case runs in its own subprocess with a timeout, so a trap or a hang in one configuration does not take
the others down.  Intermediate workspace tensors (S, S~, den) are compared with the oracle phase by phase.
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (B, H, M, w, D, normalize, rope, dtype, stop_phase, unfused)
    "p1_tiny": (1, 1, 2, 128, 64, False, False, "bf16", 1, True),
    "p1_norm": (1, 1, 2, 128, 64, True, False, "bf16", 1, True),
    "p2_tiny": (1, 1, 2, 128, 64, True, False, "bf16", 2, True),
    "p3_tiny": (1, 1, 2, 128, 64, True, False, "bf16", 3, True),
    "p3_w256": (1, 2, 4, 256, 64, True, False, "bf16", 3, True),
    "p3_w16": (2, 3, 16, 16, 64, True, False, "bf16", 3, True),
    "p3_w210_d128": (1, 2, 6, 210, 128, True, False, "bf16", 3, True),
    "p3_d128_nonorm": (1, 2, 5, 256, 128, False, False, "bf16", 3, True),
    "p3_m150": (1, 1, 150, 64, 64, True, False, "bf16", 3, True),
    "p3_rope": (1, 2, 4, 256, 64, True, True, "bf16", 3, True),
    "p3_rope_d128": (1, 1, 3, 128, 128, True, True, "fp16", 3, True),
    "only_p3_tiny": (1, 1, 2, 128, 64, True, False, "bf16", 4, True),
    "only_p3_w256": (1, 2, 4, 256, 64, True, False, "bf16", 4, True),
    "only_p3_d128": (1, 2, 6, 210, 128, True, True, "bf16", 4, True),
    "fused_tiny": (1, 1, 2, 128, 64, True, False, "bf16", 3, False),
    "fused_mid": (2, 4, 32, 256, 64, True, False, "bf16", 3, False),
    "fused_d128": (1, 3, 20, 210, 128, True, True, "bf16", 3, False),
    "fused_big": (2, 16, 128, 256, 64, True, False, "bf16", 3, False),
    "fused_big_nonorm": (2, 16, 128, 256, 64, False, False, "bf16", 3, False),
}


def run_case(name):
    import torch
    import oracle
    from mhla_b200 import _capi
    from mhla_b200.ops import _t5

    B, H, M, w, D, normalize, rope, dt, stop, unfused = CASES[name]
    dtype = torch.bfloat16 if dt == "bf16" else torch.float16
    g = torch.Generator().manual_seed(0)
    q = (torch.relu(torch.randn(B, H, M, w, D, generator=g)) + 1e-6).to(dtype)
    k = (torch.relu(torch.randn(B, H, M, w, D, generator=g)) + 1e-6).to(dtype)
    v = torch.randn(B, H, M, w, D, generator=g).to(dtype)
    qr = kr = None
    if rope:
        qr = torch.randn(B, H, M, w, D, generator=g).to(dtype)
        kr = torch.randn(B, H, M, w, D, generator=g).to(dtype)
    W = torch.rand(M, M, generator=g) / M + 0.5 * torch.eye(M) / M
    dev = torch.device("cuda")
    tq, tk, tv = q.to(dev), k.to(dev), v.to(dev)
    tqr, tkr = (qr.to(dev), kr.to(dev)) if rope else (None, None)
    tW = W.to(dev)
    out = torch.full((B, H, M, w, D), float("nan"), dtype=dtype, device=dev)

    d = _capi.BlockmixDesc()
    d.B, d.H, d.M, d.w, d.D = B, H, M, w, D
    d.dtype = 0 if dt == "bf16" else 1
    flags = _capi.FLAG_NORMALIZE if normalize else 0
    if unfused:
        flags |= _capi.FLAG_UNFUSED
        if stop == 1:
            flags |= _capi.FLAG_STOP_AFTER_P1
        elif stop == 2:
            flags |= _capi.FLAG_STOP_AFTER_P2
        elif stop == 4:
            flags |= _capi.FLAG_ONLY_P3
    d.flags, d.eps = flags, 1e-6
    d.q, d.k, d.v, d.out = _t5(tq), _t5(tk), _t5(tv), _t5(out)
    d.q_rope, d.k_rope = _t5(tqr), _t5(tkr)
    d.mix, d.mix_ld = tW.data_ptr(), M
    L = _capi.lib()
    nbytes = L.mhla_blockmix_workspace_bytes(C.byref(d))
    lay = (C.c_size_t * 8)()
    assert L.mhla_blockmix_workspace_layout(C.byref(d), C.byref(lay)) == 0
    ws = torch.full((nbytes + 1024,), 0xFF, dtype=torch.uint8, device=dev)   # NaN-poisoned workspace
    base = (ws.data_ptr() + 1023) // 1024 * 1024
    shift = base - ws.data_ptr()
    ws = ws
    d.workspace, d.workspace_bytes = base, nbytes
    offS, offSt, offDen, offW, offC, ncols, wpad, Mp = [int(x) for x in lay]
    G = B * H
    f32 = torch.float32
    qf, kf, vf = q.to(f32), k.to(f32), v.to(f32)
    knum = kr.to(f32) if rope else kf
    if stop == 4:   # readout only: fill S~ and den with oracle values
        S_ref = oracle.blockmix_summaries(knum, vf).reshape(G, M, D * D)
        St_ref = torch.einsum("ij,gjc->gic", W, S_ref).to(dtype)
        ws[shift + offSt: shift + offSt + G * M * D * D * 2] = St_ref.contiguous().view(torch.uint8).flatten().to(dev)
        nloc_ref = torch.einsum("bhjtd,bhjd->bhjt", qf, kf.sum(-2)).reshape(G, M, w)
        den_ref = torch.zeros(G, M, 2 * wpad)
        den_ref[:, :, :w] = torch.einsum("ij,gjt->git", W, nloc_ref)
        ws[shift + offDen: shift + offDen + G * M * 2 * wpad * 4] = den_ref.contiguous().view(torch.uint8).flatten().to(dev)
    torch.cuda.synchronize()
    rc = L.mhla_fwd_blockmix(C.byref(d), torch.cuda.current_stream().cuda_stream)
    res = {"case": name, "rc": rc, "launches": L.mhla_last_launch_count()}
    if rc != 0:
        res["err"] = L.mhla_strerror(rc).decode() + " / " + L.mhla_last_cuda_error().decode()
        return res
    torch.cuda.synchronize()
    wsv = ws[shift:]
    if stop == 4:
        ref = oracle.blockmix_fwd(qf, kf, vf, W, eps=1e-6, normalize=normalize, q_rope=qr, k_rope=kr)
        o = out.cpu().to(f32)
        res["out_err"] = oracle.err_ratio(ref, o)
        res["out_nan"] = int(torch.isnan(o).sum())
        return res
    Wp = wsv[offW:offW + 2 * M * Mp * 2].view(dtype).view(2, M, Mp).cpu().float().sum(0)
    res["Wp_err"] = oracle.err_ratio(W, Wp[:, :M])
    S_all = wsv[offS:offS + G * M * ncols * 2].view(dtype).view(G, M, ncols).cpu().float()
    S_ref = oracle.blockmix_summaries(knum, vf).reshape(G, M, D * D)
    res["S_err"] = oracle.err_ratio(S_ref, S_all[:, :, :D * D])
    res["S_nan"] = int(torch.isnan(S_all[:, :, :D * D]).sum())
    if normalize:
        nloc_ref = torch.einsum("bhjtd,bhjd->bhjt", qf, kf.sum(-2)).reshape(G, M, w)
        res["nloc_err"] = oracle.err_ratio(nloc_ref, S_all[:, :, D * D:D * D + w] + S_all[:, :, D * D + wpad:D * D + wpad + w])
    if stop >= 2:
        St = wsv[offSt:offSt + G * M * D * D * 2].view(dtype).view(G, M, D * D).cpu().to(f32)
        St_ref = torch.einsum("ij,gjc->gic", W, S_ref)
        res["St_err"] = oracle.err_ratio(St_ref, St)
        if normalize:
            den = wsv[offDen:offDen + G * M * 2 * wpad * 4].view(f32).view(G, M, 2 * wpad).cpu()
            den_ref = torch.einsum("ij,gjt->git", W, nloc_ref)
            res["den_err"] = oracle.err_ratio(den_ref, den[:, :, :w] + den[:, :, wpad:wpad + w])
    if stop >= 3:
        ref = oracle.blockmix_fwd(qf, kf, vf, W, eps=1e-6, normalize=normalize, q_rope=qr, k_rope=kr)
        o = out.cpu().to(f32)
        res["out_err"] = oracle.err_ratio(ref, o)
        res["out_nan"] = int(torch.isnan(o).sum())
        res["out_maxabs"] = float((ref - o).abs().max() / ref.abs().max())
    if os.environ.get("MHLA_DIAG_DUMP"):
        import numpy as np
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"diag_{name}.npz"), S=S_all.numpy(), S_ref=S_ref.numpy())
    return res


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        print("RESULT " + json.dumps(run_case(sys.argv[2])))
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    for n in names:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True, text=True,
                               timeout=120, env=dict(os.environ))
            lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                print(lines[-1][7:])
            else:
                print(json.dumps({"case": n, "exit": p.returncode, "stdout": p.stdout[-800:], "stderr": p.stderr[-1200:]}))
        except subprocess.TimeoutExpired:
            print(json.dumps({"case": n, "timeout": True}))
        sys.stdout.flush()
