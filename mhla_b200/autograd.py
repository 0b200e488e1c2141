"""Autograd support for the MHLA operators (trainers: mhla_dit/train.py:298-310,
mhla_image_classification timm_train.py:1137-1170, mhla_videogen train_wan.py:717, the fla HF Trainer).

Forward = the hand-written CUDA kernels (``ops._blockmix_fwd`` / ``ops._causal_fwd``).  Backward = the analytic gradient
of the same formulas, evaluated with batched fp32 matmuls on the tensors' own device (cuBLAS on the GPU) - interim
library code until the forward kernels' P1/P2/P3 items are re-instantiated for the gradient contractions (every one of
them has the shape of a forward phase: dS~_i = Q_i^T dO~_i is a P1, dS = W^T dS~ a P2, dQ_i = dO~_i S~_i^T a P3 ...).
The point of this file is that training with the drop-in modules is CORRECT: the reference's trainable mixing
matrices (``piece_attn.conv.weight``, ``block_attn.conv.weight``, ``mixing_matrix``) and the q/k/v projections receive
the gradients of the reference operator.  The math below is pure torch and device-agnostic, so the CPU test-suite
checks it against ``torch.autograd`` of the oracle (tests/test_autograd_cpu.py).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


# ------------------------------------------------------------------------------------------------ variants A / B
def blockmix_backward(q, k, v, W, do, *, q_rope=None, k_rope=None, eps: float = 1e-6, normalize: bool = True,
                      needs: Tuple[bool, ...] = (True,) * 6):
    """Gradients of out = blockmix(q, k, v, W[, q_rope, k_rope]) (mhla.py:262-268 / mhla_utils.py:328-341) for an upstream
    gradient ``do``; tensors are [..., M, w, D] (any leading dims), W is [M, M].  Returns (dq, dk, dv, dW, dq_rope,
    dk_rope) in fp32; entries whose ``needs`` flag is False may be None."""
    f = torch.float32
    qn_, kn_ = (q if q_rope is None else q_rope), (k if k_rope is None else k_rope)
    Qn, Kn, V, Wf, dO = qn_.to(f), kn_.to(f), v.to(f), W.to(f), do.to(f)
    S = torch.matmul(Kn.transpose(-2, -1), V)                       # [..., M, D, D]
    St = torch.einsum("ij,...jab->...iab", Wf, S)
    dnl = None
    if normalize:
        Q, K = q.to(f), k.to(f)
        ksum = K.sum(dim=-2)                                        # [..., M, D]
        nl = torch.einsum("...jtd,...jd->...jt", Q, ksum)           # [..., M, w]
        den = torch.einsum("ij,...jt->...it", Wf, nl) + eps
        num = torch.matmul(Qn, St)
        dnum = dO / den.unsqueeze(-1)
        dden = -(dnum * num).sum(dim=-1) / den                      # d/dden of num/den, per token
        dnl = torch.einsum("ij,...it->...jt", Wf, dden)
    else:
        dnum = dO
    dQn = torch.matmul(dnum, St.transpose(-2, -1))
    dSt = torch.matmul(Qn.transpose(-2, -1), dnum)                  # [..., M, D, D]
    dS = torch.einsum("ij,...iab->...jab", Wf, dSt)
    dKn = torch.matmul(V, dS.transpose(-2, -1))
    dV = torch.matmul(Kn, dS)
    dW = None
    if needs[3]:
        lead = tuple(range(dSt.dim() - 3))
        dW = torch.einsum("...iab,...jab->...ij", dSt, S).sum(dim=lead) if lead else torch.einsum("iab,jab->ij", dSt, S)
        if normalize:
            t = torch.einsum("...it,...jt->...ij", dden, nl)
            dW = dW + (t.sum(dim=lead) if lead else t)
    dq = dk = None
    if normalize:
        dq = dnl.unsqueeze(-1) * ksum.unsqueeze(-2)                 # [..., M, w, D]
        dksum = torch.einsum("...jt,...jtd->...jd", dnl, Q)
        dk = dksum.unsqueeze(-2).expand_as(K)
    if q_rope is None:
        dq = dQn if dq is None else dq + dQn
        dk = dKn if dk is None else dk + dKn
        return dq, dk, dV, dW, None, None
    if dq is None:
        dq, dk = torch.zeros_like(Qn), torch.zeros_like(Kn)
    return dq, dk, dV, dW, dQn, dKn


class BlockmixFunction(torch.autograd.Function):
    """forward: CUDA kernel; backward: ``blockmix_backward``."""

    @staticmethod
    def forward(ctx, q, k, v, mix, q_rope, k_rope, eps, normalize, kw):
        from . import ops
        out = ops._blockmix_fwd(q, k, v, mix, q_rope=q_rope, k_rope=k_rope, eps=eps, normalize=normalize, **kw)
        ctx.save_for_backward(q, k, v, mix, q_rope, k_rope)
        ctx.eps, ctx.normalize = eps, normalize
        return out

    @staticmethod
    def backward(ctx, do):
        q, k, v, mix, q_rope, k_rope = ctx.saved_tensors
        M = q.shape[-3]
        dq, dk, dv, dW, dqr, dkr = blockmix_backward(
            q, k, v, mix.reshape(M, M), do, q_rope=q_rope, k_rope=k_rope, eps=ctx.eps, normalize=ctx.normalize,
            needs=tuple(ctx.needs_input_grad[:6]))
        cast = lambda g, ref: None if (g is None or ref is None) else g.to(ref.dtype)   # noqa: E731
        return (cast(dq, q), cast(dk, k), cast(dv, v), None if dW is None else dW.reshape(mix.shape).to(mix.dtype),
                cast(dqr, q_rope), cast(dkr, k_rope), None, None, None)


# ------------------------------------------------------------------------------------------------ variant C
def causal_backward(q, k, v, mm, do, chunk_size: int = 64, scale: Optional[float] = None):
    """Gradients of o = causal_chunk(q, k, v, mm) (naive.py:10-83): q,k [B,T,H,K], v,do [B,T,H,V], mm [L,L].
    o_i = scale (q_i sum_{j<i} mm_ij S_j + mm_ii tril(q_i k_i^T) v_i),  S_j = k_j^T v_j.  Returns (dq, dk, dv, dmm) fp32."""
    f = torch.float32
    B, T, H, K = q.shape
    V = v.shape[-1]
    c = chunk_size
    sc = float(K ** -0.5 if scale is None else scale)
    pad = (c - T % c) % c
    prep = lambda t: torch.nn.functional.pad(t.to(f).transpose(1, 2), (0, 0, 0, pad))   # noqa: E731  b h t d
    n = (T + pad) // c
    Q, Kk, Vv, dO = (prep(t).reshape(B, H, n, c, -1) for t in (q, k, v, do))
    L = mm.shape[0]
    mmf = mm.reshape(L, mm.shape[1]).to(f)[:n, :n]
    lower = torch.tril(mmf, -1) * sc                       # strictly-lower mixing, scale folded
    diag = torch.diagonal(mmf) * sc                        # [n]
    tril = torch.tril(torch.ones(c, c, dtype=f, device=q.device))
    S = torch.matmul(Kk.transpose(-2, -1), Vv)             # [B,H,n,K,V]
    P = torch.einsum("ij,bhjkv->bhikv", lower, S)          # prefix state seen by chunk i
    A = torch.matmul(Q, Kk.transpose(-2, -1)) * tril       # [B,H,n,c,c]
    G = torch.matmul(dO, Vv.transpose(-2, -1)) * tril      # dA (before the diagonal weight)
    dg = diag.view(1, 1, n, 1, 1)
    dP = torch.matmul(Q.transpose(-2, -1), dO)             # [B,H,n,K,V]
    dS = torch.einsum("ij,bhikv->bhjkv", lower, dP)
    dQ = torch.matmul(dO, P.transpose(-2, -1)) + dg * torch.matmul(G, Kk)
    dK = dg * torch.matmul(G.transpose(-2, -1), Q) + torch.matmul(Vv, dS.transpose(-2, -1))
    dV = dg * torch.matmul(A.transpose(-2, -1), dO) + torch.matmul(Kk, dS)
    dmm_low = torch.einsum("bhikv,bhjkv->ij", dP, S) * sc
    dmm = torch.tril(dmm_low, -1) + torch.diag_embed((dO * torch.matmul(A, Vv)).sum(dim=(0, 1, 3, 4)) * sc)
    full = torch.zeros(L, mm.shape[1], dtype=f, device=q.device)
    full[:n, :n] = dmm
    unprep = lambda t: t.reshape(B, H, n * c, -1)[:, :, :T].transpose(1, 2)   # noqa: E731
    return unprep(dQ), unprep(dK), unprep(dV), full


class CausalFunction(torch.autograd.Function):
    """forward: CUDA kernel; backward: ``causal_backward``."""

    @staticmethod
    def forward(ctx, q, k, v, mixing_matrix, chunk_size, scale, unfused):
        from . import ops
        out = ops._causal_fwd(q, k, v, mixing_matrix, chunk_size=chunk_size, scale=scale, unfused=unfused)
        ctx.save_for_backward(q, k, v, mixing_matrix)
        ctx.chunk_size, ctx.scale = chunk_size, scale
        return out

    @staticmethod
    def backward(ctx, do):
        q, k, v, mm = ctx.saved_tensors
        dq, dk, dv, dmm = causal_backward(q, k, v, mm, do, ctx.chunk_size, ctx.scale)
        return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), dmm.reshape(mm.shape).to(mm.dtype), None, None, None
