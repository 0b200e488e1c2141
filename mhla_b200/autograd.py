"""Autograd support for the MHLA operators (trainers: mhla_dit/train.py:298-310,
mhla_image_classification timm_train.py:1137-1170, mhla_videogen train_wan.py:717, the fla HF Trainer).

Forward = the hand-written CUDA kernels (``ops._blockmix_fwd`` / ``ops._causal_fwd``).  Backward on the GPU = the SAME
kernels: every gradient contraction of the operator has the shape of a forward call with permuted operands,

    dQ = mhla(q=dO~, k=V,   v=K, W)          (dO~_i S~_i^T,  S~_i^T = sum_j W_ij V_j^T K_j)
    dV = mhla(q=K,   k=Q,   v=dO~, W^T)      (K_j dS_j,      dS_j   = sum_i W_ij Q_i^T dO~_i)
    dK = mhla(q=V,   k=dO~, v=Q, W^T)        (V_j dS_j^T)

with dO~ = dO / den (un-normalised: dO~ = dO), so ``blockmix_backward_native`` / ``causal_backward_native`` are three
(causal: five or more, see there) launches of the forward kernels - TMA, tcgen05 and the run-time scheduler included -
on 16-bit operands with fp32 accumulation.  The gradient of the mixing matrix contracts the block summaries the kernels
themselves produced (read back from the launches' workspaces) in one small GEMM; the normaliser's own terms (rank-1 per
block) are elementwise torch.  ``blockmix_backward`` / ``causal_backward`` state the same gradients in plain fp32 torch:
device-agnostic, checked on CPU against ``torch.autograd`` of the oracle (tests/test_autograd_cpu.py), and the yardstick
the GPU tests hold the native backward to.  The point of this file is that training with the drop-in modules is
CORRECT: the reference's trainable mixing matrices (``piece_attn.conv.weight``, ``block_attn.conv.weight``,
``mixing_matrix``) and the q/k/v projections receive the gradients of the reference operator.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

_TORCH_BACKWARD = os.environ.get("MHLA_TORCH_BACKWARD", "0") == "1"


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


def _summary_dot(ws_a: dict, ws_b: dict) -> torch.Tensor:
    """sum_g  A_g B_g^T  over the block summaries [G, M, D*D] two launches left in their workspaces -> [M, M] fp32
    (one batched 16-bit GEMM with fp32 accumulation per unit, the units summed in fp32)."""
    dd = ws_a["D"] * ws_a["D"]
    dt = ws_a["dtype"]
    a = ws_a["S"][..., :dd].view(dt)
    b = ws_b["S"][..., :dd].view(dt)
    return torch.bmm(a, b.transpose(1, 2)).float().sum(dim=0)


def blockmix_backward_native(q, k, v, W, do, out, *, q_rope=None, k_rope=None, eps: float = 1e-6, normalize: bool = True,
                             needs: Tuple[bool, ...] = (True,) * 6, nl=None, den=None):
    """``blockmix_backward`` on the forward CUDA kernels (see the module docstring): CUDA tensors [..., M, w, D] with 4
    or 5 dims, ``out`` = the saved forward output (num = out * den).  ``nl`` / ``den`` [..., M, w]: the forward launch's own
    n_loc and normaliser (read from its workspace by ``BlockmixFunction.forward``); recomputed here when absent.
    Returns (dq, dk, dv, dW, dq_rope, dk_rope) in the 16-bit compute dtype (dW fp32)."""
    from . import ops
    f = torch.float32
    cd = q.dtype if q.dtype in ops._DT else torch.bfloat16
    c = lambda t: t if t.dtype == cd else t.to(cd)   # noqa: E731
    Qn, Kn = c(q if q_rope is None else q_rope), c(k if k_rope is None else k_rope)
    Vc = c(v)
    D = q.shape[-1]
    Wf = W.detach().to(f)
    Wt = Wf.t().contiguous()
    aux = q.is_cuda and D in (64, 128) and cd in (torch.bfloat16, torch.float16)   # csrc/bwd_aux_kernel.cuh
    ksum = dden = None
    if normalize:
        ksum = ops.block_wsum(c(k)) if aux else k.sum(dim=-2, dtype=f)      # [..., M, D]
        if den is None or nl is None:
            nl = torch.einsum("...jtd,...jd->...jt", q.to(f), ksum)         # [..., M, w]
            den = torch.einsum("ij,...jt->...it", Wf, nl) + eps
        if aux:
            dnum, dden = ops.bwd_prep(c(do), c(out), den)
        else:
            dden = -(do.to(f) * out.to(f)).sum(dim=-1) / den                # num = out * den
            dnum = (do.to(f) / den.unsqueeze(-1)).to(cd)
    else:
        dnum = c(do)
    fwd = lambda q_, k_, v_, w_, ws: ops._blockmix_fwd(q_, k_, v_, w_, normalize=False, ws_out=ws)   # noqa: E731
    want_w = bool(needs[3])
    ws_a, ws_c = ({}, {}) if want_w else (None, None)
    dQn = fwd(dnum, Vc, Kn, Wf, ws_a)            # dO~_i S~_i^T;   its summaries: V_j^T K_j = S_j^T
    dV = fwd(Kn, Qn, dnum, Wt, None)             # K_j dS_j
    dKn = fwd(Vc, dnum, Qn, Wt, ws_c)            # V_j dS_j^T;     its summaries: dO~_i^T Q_i = dS~_i^T
    dW = None
    if want_w:
        if ws_a and ws_c and ws_a["D"] == D:
            ws_a["dtype"] = ws_c["dtype"] = cd
            dW = _summary_dot(ws_c, ws_a)        # <dS~_i, S_j> summed over the (b,h) units
        else:   # short-sequence kernel (no workspace) / zero-padded head dim: the summaries are tiny, take them from cuBLAS
            S = torch.matmul(Kn.transpose(-2, -1), Vc).to(f)
            dSt = torch.matmul(Qn.transpose(-2, -1), dnum).to(f)
            dW = torch.einsum("...iab,...jab->ij", dSt, S)
        if normalize:
            dW = dW + torch.einsum("...it,...jt->ij", dden, nl.to(f))
    dq = dk = None
    if normalize:
        dnl = torch.einsum("ij,...it->...jt", Wf, dden)                     # [..., M, w]
        if aux:
            dksum = ops.block_wsum(c(q), dnl)                                # sum_t dnl[j, t] q[j, t, :]
            rope = q_rope is not None
            dq, dk = ops.bwd_post(None if rope else dQn, None if rope else dKn, dnl, ksum, dksum, cd)
            return (dq, dk, dV, dW, dQn, dKn) if rope else (dq, dk, dV, dW, None, None)
        dq = dnl.unsqueeze(-1) * ksum.unsqueeze(-2)
        dk = torch.einsum("...jt,...jtd->...jd", dnl, q.to(f)).unsqueeze(-2).expand(q.shape)
    if q_rope is None:
        dq = dQn if dq is None else dq + dQn
        dk = dKn if dk is None else dk + dKn
        return dq, dk, dV, dW, None, None
    if dq is None:
        dq, dk = torch.zeros_like(Qn), torch.zeros_like(Kn)
    return dq, dk, dV, dW, dQn, dKn


class BlockmixFunction(torch.autograd.Function):
    """forward: CUDA kernel; backward: the same kernel with permuted operands (``blockmix_backward_native``;
    ``MHLA_TORCH_BACKWARD=1`` selects the plain-torch statement ``blockmix_backward`` instead)."""

    @staticmethod
    def forward(ctx, q, k, v, mix, q_rope, k_rope, eps, normalize, kw):
        from . import ops
        ws = {} if (normalize and not _TORCH_BACKWARD) else None
        out = ops._blockmix_fwd(q, k, v, mix, q_rope=q_rope, k_rope=k_rope, eps=eps, normalize=normalize, ws_out=ws, **kw)
        nl = den = None
        if ws:   # the general kernel ran: keep ITS n_loc and normaliser (G*M*w floats each) for the backward pass
            wpad, dd, w = ws["wpad"], ws["D"] * ws["D"], q.shape[-2]
            dt = out.dtype if out.dtype in ops._DT else torch.bfloat16
            S = ws["S"]
            nl = (S[..., dd:dd + wpad].view(dt).float() + S[..., dd + wpad:dd + 2 * wpad].view(dt).float())[..., :w]
            den = (ws["den"][..., :wpad] + ws["den"][..., wpad:])[..., :w] + eps
            nl, den = nl.reshape(q.shape[:-1]), den.reshape(q.shape[:-1])
        ctx.save_for_backward(q, k, v, mix, q_rope, k_rope, out, nl, den)
        ctx.eps, ctx.normalize = eps, normalize
        return out

    @staticmethod
    def backward(ctx, do):
        q, k, v, mix, q_rope, k_rope, out, nl, den = ctx.saved_tensors
        M = q.shape[-3]
        fn = blockmix_backward if _TORCH_BACKWARD else (
            lambda *a, **kw_: blockmix_backward_native(*a, out, nl=nl, den=den, **kw_))
        dq, dk, dv, dW, dqr, dkr = fn(
            q, k, v, mix.reshape(M, M), do, q_rope=q_rope, k_rope=k_rope, eps=ctx.eps, normalize=ctx.normalize,
            needs=tuple(ctx.needs_input_grad[:6]))
        cast = lambda g, ref: None if (g is None or ref is None) else g.to(ref.dtype)   # noqa: E731
        return (cast(dq, q), cast(dk, k), cast(dv, v), None if dW is None else dW.reshape(mix.shape).to(mix.dtype),
                cast(dqr, q_rope), cast(dkr, k_rope), None, None, None)


class BlockmixGridFunction(torch.autograd.Function):
    """Token-major 3-D block view (Wan, ``ops.mhla_blockmix_grid``) without the normaliser - the shipped Wan configuration:
    forward and the three gradient launches all gather / scatter the blocks by TMA, so a training step makes no layout
    copy either.  q, k are the (roped) numerator operands, [B, N, heads, D]."""

    @staticmethod
    def forward(ctx, q, k, v, mix, grid, layout, eps):
        from . import ops
        out = ops._blockmix_grid_fwd(q, k, v, mix, grid, layout, eps=eps, normalize=False)
        ctx.save_for_backward(q, k, v, mix)
        ctx.grid, ctx.layout, ctx.eps = grid, layout, eps
        return out

    @staticmethod
    def backward(ctx, do):
        from . import ops
        q, k, v, mix = ctx.saved_tensors
        cd = q.dtype if q.dtype in ops._DT else torch.bfloat16
        c = lambda t: t if t.dtype == cd else t.to(cd)   # noqa: E731
        M = mix.shape[0]
        Wf = mix.detach().reshape(M, M).float()
        Wt = Wf.t().contiguous()
        want_w = ctx.needs_input_grad[3]
        ws_a, ws_c = ({}, {}) if want_w else (None, None)
        fwd = lambda q_, k_, v_, w_, ws: ops._blockmix_grid_fwd(c(q_), c(k_), c(v_), w_, ctx.grid, ctx.layout,   # noqa: E731
                                                                eps=ctx.eps, normalize=False, ws_out=ws)
        dq = fwd(do, v, k, Wf, ws_a)
        dv = fwd(k, q, do, Wt, None)
        dk = fwd(v, do, q, Wt, ws_c)
        dW = None
        if want_w:
            ws_a["dtype"] = ws_c["dtype"] = cd
            dW = _summary_dot(ws_c, ws_a).reshape(mix.shape).to(mix.dtype)
        return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), dW, None, None, None


# ------------------------------------------------------------------------------------------------ Wan pre-processing
def wan_prep_reference(x, w, cos, sin, head_dim: int, eps_norm: float = 1e-6, eps: float = 1e-6):
    """Differentiable torch restatement of what ``ops.wan_prep`` computes for ONE of its two inputs (mhla_utils.py:267-276,
    :127-156): y = relu(x * rsqrt(mean_C(x^2) + eps_norm) * w) + eps, then the interleaved-pair rotation of every head by
    the token's angles.  x [B, N, C], w [C] or None, cos / sin [N, D/2] -> [B, N, C // D, D] in x's float type.  The
    checker of ``wan_prep_backward`` (tests/test_autograd_cpu.py) and the formula the CUDA kernel is tested against."""
    B, N, Cc = x.shape
    D = int(head_dim)
    r = torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + eps_norm)
    y = x * r
    if w is not None:
        y = y * w
    y = torch.relu(y) + eps
    yp = y.reshape(B, N, Cc // D, D // 2, 2)
    a, b = yp[..., 0], yp[..., 1]
    c, s_ = cos[None, :, None, :].to(y.dtype), sin[None, :, None, :].to(y.dtype)
    return torch.stack((a * c - b * s_, a * s_ + b * c), dim=-1).reshape(B, N, Cc // D, D)


def wan_prep_backward(x, w, cos, sin, head_dim: int, eps_norm: float, g_rope, g_plain=None):
    """Gradients of ``wan_prep_reference`` w.r.t. x and w given the gradient of the roped output (and, optionally, of the
    un-roped output relu(.) + eps): returns (gx [B, N, C], gw [C] or None) in fp32 (fp64 for fp64 inputs).
      rotation^T:  g_a = g0 c + g1 s,  g_b = -g0 s + g1 c
      relu:        g_y = g * [y > 0]
      RMSNorm:     u = x r, y = u w:  gw = sum_rows g_y u,  gx = r (g_u - u mean_C(g_u u)),  g_u = g_y w"""
    f = torch.float64 if x.dtype == torch.float64 else torch.float32
    B, N, Cc = x.shape
    D = int(head_dim)
    xf = x.to(f)
    r = torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps_norm)
    u = xf * r
    wf = None if w is None else w.to(f)
    y = u if wf is None else u * wf
    g = g_rope.to(f).reshape(B, N, Cc // D, D // 2, 2)
    g0, g1 = g[..., 0], g[..., 1]
    c, s_ = cos[None, :, None, :].to(f), sin[None, :, None, :].to(f)
    gy = torch.stack((g0 * c + g1 * s_, g1 * c - g0 * s_), dim=-1).reshape(B, N, Cc)
    if g_plain is not None:
        gy = gy + g_plain.to(f).reshape(B, N, Cc)
    gy = gy * (y > 0).to(f)
    gw = None if wf is None else (gy * u).sum(dim=(0, 1))
    gu = gy if wf is None else gy * wf
    gx = r * (gu - u * (gu * u).mean(dim=-1, keepdim=True))
    return gx, gw


class WanPrepFunction(torch.autograd.Function):
    """Training through the fused Wan pre-processing: forward = the ONE CUDA launch of ``ops.wan_prep`` (RMSNorm over C,
    relu + eps, RoPE, 16-bit token-major outputs), backward = ``wan_prep_backward`` (a dozen fp32 elementwise / row
    reductions per input) - instead of autograd through ~10 fp32 / complex temporaries per input in both directions
    (mhla_utils.py:267-276, :127-156, :303-316).  Returns (q_rope, k_rope) as [B, N, heads, D]."""

    @staticmethod
    def forward(ctx, xq, xk, wq, wk, cos, sin, head_dim, eps_norm, eps):
        from . import ops
        qr, kr, _, _ = ops.wan_prep(xq, xk, wq, wk, cos, sin, head_dim, eps_norm=eps_norm, eps=eps, want_plain=False)
        ctx.save_for_backward(xq, xk, wq, wk, cos, sin)
        ctx.head_dim, ctx.eps_norm = int(head_dim), float(eps_norm)
        return qr, kr

    @staticmethod
    def backward(ctx, gq, gk):
        xq, xk, wq, wk, cos, sin = ctx.saved_tensors
        gxq = gxk = gwq = gwk = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[2]:
            gxq, gwq = wan_prep_backward(xq, wq, cos, sin, ctx.head_dim, ctx.eps_norm, gq)
            gxq = gxq.to(xq.dtype)
            gwq = None if (gwq is None or not ctx.needs_input_grad[2]) else gwq.to(wq.dtype)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[3]:
            gxk, gwk = wan_prep_backward(xk, wk, cos, sin, ctx.head_dim, ctx.eps_norm, gk)
            gxk = gxk.to(xk.dtype)
            gwk = None if (gwk is None or not ctx.needs_input_grad[3]) else gwk.to(wk.dtype)
        return gxq, gxk, gwq, gwk, None, None, None, None, None


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


def causal_backward_native(q, k, v, mm, do, chunk_size: int = 64, scale: Optional[float] = None):
    """``causal_backward`` on the forward CUDA kernel.  With G = tril(dO V^T), A = tril(Q K^T), dP_i = Q_i^T dO_i:
        dQ_i = dO_i P_i^T + mm_ii G_i K_i          = causal(q=dO, k=V, v=K; mm)                     (forward in time)
        dV_j = K_j dS_j  + mm_jj A_j^T dO_j        = causal(q=K, k=Q, v=dO; mm^T) on the time-REVERSED sequence
        dK_j = V_j dS_j^T + mm_jj G_j^T Q_j        = causal(q=V, k=dO, v=Q; mm^T) on the time-reversed sequence
    (reversing the token order turns the anti-causal sums over later chunks / later tokens into causal ones; the mixing
    matrix becomes its transpose flipped along both axes).  The kernel's key dim is at most 128, so calls whose key
    operand is the value tensor (dQ, dK) run once per 128-wide slice of the value dim and are summed.  d mm: the diagonal
    is <dO, causal(q, k, v; I)> per chunk (the kernel with an identity mixing matrix = the masked intra-chunk product);
    the strictly-lower part contracts the chunk states Q_i^T dO_i and K_j^T V_j (one batched GEMM each, then [n, n])."""
    from . import ops
    f = torch.float32
    B, T, H, K = q.shape
    V = v.shape[-1]
    c = chunk_size
    sc = float(K ** -0.5 if scale is None else scale)
    cd = q.dtype if q.dtype in ops._DT else torch.bfloat16
    pad = (c - T % c) % c
    prep = lambda t: torch.nn.functional.pad(t.to(cd), (0, 0, 0, 0, 0, pad)) if pad else t.to(cd)   # noqa: E731
    Q, Kk, Vv, dO = prep(q), prep(k), prep(v), prep(do)
    n = (T + pad) // c
    L = mm.shape[0]
    mmf = mm.detach().reshape(L, mm.shape[1]).to(f)[:n, :n].contiguous()
    mmr = mmf.t().flip(0, 1).contiguous()
    rev = lambda t: t.flip(1)   # noqa: E731
    fwd = lambda q_, k_, v_, m_: ops._causal_fwd(q_, k_, v_, m_, chunk_size=c, scale=sc)   # noqa: E731
    Qr, dOr = rev(Q), rev(dO)
    dQ = torch.zeros(Q.shape, dtype=f, device=q.device)
    dK = torch.zeros(Q.shape, dtype=f, device=q.device)
    for a in range(0, V, 128):
        sl = slice(a, min(V, a + 128))
        dQ += fwd(dO[..., sl], Vv[..., sl], Kk, mmf)
        dK += fwd(rev(Vv[..., sl]), dOr[..., sl], Qr, mmr)
    dV = fwd(rev(Kk), Qr, dOr, mmr).flip(1)
    dK = dK.flip(1)
    # d mm
    eye = torch.eye(n, dtype=f, device=q.device)
    o_diag = fwd(Q, Kk, Vv, eye)                                      # sc * tril(q_i k_i^T) v_i
    diag = (dO.to(f) * o_diag.to(f)).view(B, n, c, H, V).sum(dim=(0, 2, 3, 4))
    Q5, K5, V5, dO5 = (t.view(B, n, c, H, -1) for t in (Q, Kk, Vv, dO))
    S = torch.einsum("bnchk,bnchv->bhnkv", K5, V5)
    dP = torch.einsum("bnchk,bnchv->bhnkv", Q5, dO5)
    low = torch.einsum("bhikv,bhjkv->ij", dP.to(f), S.to(f)) * sc
    dmm = torch.tril(low, -1) + torch.diag_embed(diag)
    full = torch.zeros(L, mm.shape[1], dtype=f, device=q.device)
    full[:n, :n] = dmm
    return dQ[:, :T], dK[:, :T], dV[:, :T], full


class CausalFunction(torch.autograd.Function):
    """forward: CUDA kernel; backward: the same kernel on permuted / time-reversed operands
    (``causal_backward_native``; ``MHLA_TORCH_BACKWARD=1`` selects the plain-torch ``causal_backward``)."""

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
        fn = causal_backward if _TORCH_BACKWARD else causal_backward_native
        dq, dk, dv, dmm = fn(q, k, v, mm, do, ctx.chunk_size, ctx.scale)
        return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), dmm.reshape(mm.shape).to(mm.dtype), None, None, None
