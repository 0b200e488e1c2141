"""The analytic backward of the MHLA operators (mhla_b200/autograd.py - pure torch, device-agnostic) against
torch.autograd of the oracle, on CPU in float64-free fp32 with tight tolerances."""
import pytest
import torch

import oracle
from mhla_b200.autograd import blockmix_backward, causal_backward


@pytest.mark.parametrize("normalize,rope", [(True, False), (False, False), (True, True), (False, True)])
def test_blockmix_backward_matches_autograd(normalize, rope):
    g = torch.Generator().manual_seed(0)
    G, M, w, D = 3, 5, 12, 16
    mk = lambda relu: ((torch.relu(torch.randn(G, M, w, D, generator=g)) + 0.1) if relu  # noqa: E731
                       else torch.randn(G, M, w, D, generator=g)).double().requires_grad_(True)
    q, k, v = mk(True), mk(True), mk(False)
    qr, kr = (mk(False), mk(False)) if rope else (None, None)
    W = (torch.rand(M, M, generator=g) / M + 0.3 * torch.eye(M)).double().requires_grad_(True)
    do = torch.randn(G, M, w, D, generator=g).double()
    out = oracle.blockmix_fwd(q, k, v, W, eps=1e-6, normalize=normalize, q_rope=qr, k_rope=kr, dtype=torch.float64)
    ins = [t for t in (q, k, v, W, qr, kr) if t is not None]
    ref = torch.autograd.grad(out, ins, do, allow_unused=True)
    got = blockmix_backward(q.detach().float(), k.detach().float(), v.detach().float(), W.detach().float(), do.float(),
                            q_rope=None if qr is None else qr.detach().float(),
                            k_rope=None if kr is None else kr.detach().float(), eps=1e-6, normalize=normalize)
    got = [g_ for g_, t in zip(got, (q, k, v, W, qr, kr)) if t is not None]
    for r, o in zip(ref, got):
        if r is None:      # un-roped q,k do not enter the output without the normaliser
            assert float(o.abs().max()) == 0.0
            continue
        assert oracle.err_ratio(r, o) < 2e-5


@pytest.mark.parametrize("T,K,V", [(128, 16, 24), (100, 8, 8), (40, 16, 8)])
def test_causal_backward_matches_autograd(T, K, V):
    g = torch.Generator().manual_seed(1)
    B, H, L = 2, 2, 32
    q = torch.randn(B, T, H, K, generator=g).double().requires_grad_(True)
    k = torch.randn(B, T, H, K, generator=g).double().requires_grad_(True)
    v = torch.randn(B, T, H, V, generator=g).double().requires_grad_(True)
    mm = torch.clamp(torch.rand(L, L, generator=g), 1e-5, 1).tril().double().requires_grad_(True)
    do = torch.randn(B, T, H, V, generator=g).double()
    out = oracle.causal_chunk_fwd(q, k, v, mm, dtype=torch.float64)
    ref = torch.autograd.grad(out, [q, k, v, mm], do)
    got = causal_backward(q.detach().float(), k.detach().float(), v.detach().float(), mm.detach().float(), do.float())
    for r, o in zip(ref, got):
        assert oracle.err_ratio(r, o) < 2e-5
    n = (T + 63) // 64
    assert float(got[3][n:].abs().max()) == 0.0 and float(got[3].triu(1).abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------------------------------
# The GPU backward is a composition of FORWARD kernel launches with permuted / time-reversed operands
# (autograd.blockmix_backward_native / causal_backward_native).  Here the launches are replaced by the oracle's forward
# (float64), which checks the composition itself - operand order, W^T, time reversal, slicing of the value dim, the
# summaries read back from the launches' workspaces - against torch.autograd of the oracle.
def _stub_blockmix(q, k, v, mix, *, q_rope=None, k_rope=None, eps=1e-6, normalize=True, ws_out=None, **kw):
    out = oracle.blockmix_fwd(q, k, v, mix, eps=eps, normalize=normalize, q_rope=q_rope, k_rope=k_rope, dtype=torch.float64)
    if ws_out is not None:
        kn = k if k_rope is None else k_rope
        S = oracle.blockmix_summaries(kn.double(), v.double())            # [..., M, Dk, Dv], row-major like the workspace
        M, D = S.shape[-3], S.shape[-1]
        ws_out.update(S=S.reshape(-1, M, D * D), D=D)
    return out


def _stub_causal(q, k, v, mm, chunk_size=64, scale=None, **kw):
    o = oracle.causal_chunk_fwd(q.double(), k.double(), v.double(), mm.double(), chunk_size=chunk_size, dtype=torch.float64)
    return o * (float(scale) / q.shape[-1] ** -0.5)


@pytest.mark.parametrize("normalize,rope,use_ws", [(True, False, True), (False, False, True), (True, True, True),
                                                   (False, True, False)])
def test_blockmix_backward_native_composition(monkeypatch, normalize, rope, use_ws):
    from mhla_b200 import autograd, ops
    monkeypatch.setitem(ops._DT, torch.float64, 0)
    stub = _stub_blockmix if use_ws else (lambda *a, ws_out=None, **kw: _stub_blockmix(*a, **kw))
    monkeypatch.setattr(ops, "_blockmix_fwd", stub)
    g = torch.Generator().manual_seed(2)
    G, M, w, D = 3, 5, 12, 16
    mk = lambda relu: ((torch.relu(torch.randn(G, M, w, D, generator=g)) + 0.1) if relu  # noqa: E731
                       else torch.randn(G, M, w, D, generator=g)).double().requires_grad_(True)
    q, k, v = mk(True), mk(True), mk(False)
    qr, kr = (mk(False), mk(False)) if rope else (None, None)
    W = (torch.rand(M, M, generator=g) / M + 0.3 * torch.eye(M)).double().requires_grad_(True)
    do = torch.randn(G, M, w, D, generator=g).double()
    out = oracle.blockmix_fwd(q, k, v, W, eps=1e-6, normalize=normalize, q_rope=qr, k_rope=kr, dtype=torch.float64)
    ins = [t for t in (q, k, v, W, qr, kr) if t is not None]
    ref = torch.autograd.grad(out, ins, do, allow_unused=True)
    d = lambda t: None if t is None else t.detach()   # noqa: E731
    got = autograd.blockmix_backward_native(d(q), d(k), d(v), d(W), do, out.detach(), q_rope=d(qr), k_rope=d(kr), eps=1e-6,
                                            normalize=normalize)
    got = [g_ for g_, t in zip(got, (q, k, v, W, qr, kr)) if t is not None]
    for r, o in zip(ref, got):
        if r is None:
            assert float(o.abs().max()) == 0.0
            continue
        assert oracle.err_ratio(r, o.double()) < 2e-6


@pytest.mark.parametrize("T,K,V", [(128, 16, 24), (100, 8, 8), (192, 16, 160)])
def test_causal_backward_native_composition(monkeypatch, T, K, V):
    from mhla_b200 import autograd, ops
    monkeypatch.setitem(ops._DT, torch.float64, 0)
    monkeypatch.setattr(ops, "_causal_fwd", _stub_causal)
    g = torch.Generator().manual_seed(3)
    B, H, L = 2, 2, 32
    q = torch.randn(B, T, H, K, generator=g).double().requires_grad_(True)
    k = torch.randn(B, T, H, K, generator=g).double().requires_grad_(True)
    v = torch.randn(B, T, H, V, generator=g).double().requires_grad_(True)
    mm = torch.clamp(torch.rand(L, L, generator=g), 1e-5, 1).tril().double().requires_grad_(True)
    do = torch.randn(B, T, H, V, generator=g).double()
    out = oracle.causal_chunk_fwd(q, k, v, mm, dtype=torch.float64)
    ref = torch.autograd.grad(out, [q, k, v, mm], do)
    got = autograd.causal_backward_native(q.detach(), k.detach(), v.detach(), mm.detach(), do)
    for r, o in zip(ref, got):
        assert oracle.err_ratio(r, o.double()) < 1e-6


@pytest.mark.parametrize("with_weight,with_plain", [(True, False), (False, False), (True, True)])
def test_wan_prep_backward_matches_autograd(with_weight, with_plain):
    """autograd.wan_prep_backward (the backward of WanPrepFunction: rotation^T, relu mask, RMSNorm over the full channel
    dim) against torch.autograd of the differentiable restatement of the pre-processing (mhla_utils.py:267-276, :127-156),
    float64.  The restatement itself is checked against the reference's RoPE in test_modules (rope_apply) and the CUDA
    kernel against the reference pre-processing in test_wan_prep_kernel_matches_reference_preprocessing."""
    from mhla_b200.autograd import wan_prep_backward, wan_prep_reference
    torch.manual_seed(11)
    B, N, H, D = 2, 12, 3, 8
    x = torch.randn(B, N, H * D, dtype=torch.float64, requires_grad=True)
    w = (torch.rand(H * D, dtype=torch.float64) + 0.5).requires_grad_(True) if with_weight else None
    ang = torch.rand(N, D // 2, dtype=torch.float64) * 6.28
    cos, sin = torch.cos(ang), torch.sin(ang)
    y = wan_prep_reference(x, w, cos, sin, D, eps_norm=1e-5, eps=1e-6)
    g = torch.randn_like(y)
    loss = (y * g).sum()
    gp = None
    if with_plain:   # the un-roped output relu(.) + eps gets a gradient of its own (normaliser operands)
        r = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5)
        plain = torch.relu(x * r * w) + 1e-6
        gp = torch.randn_like(plain)
        loss = loss + (plain * gp).sum()
    grads = torch.autograd.grad(loss, [x] + ([w] if with_weight else []))
    gx, gw = wan_prep_backward(x.detach(), None if w is None else w.detach(), cos, sin, D, 1e-5, g, gp)
    assert torch.allclose(gx, grads[0], rtol=1e-9, atol=1e-11)
    if with_weight:
        assert torch.allclose(gw, grads[1], rtol=1e-9, atol=1e-11)
    else:
        assert gw is None


def test_wan_prep_reference_is_the_modules_preprocessing():
    """wan_prep_reference == relu(WanRMSNorm(x)) + eps followed by the module's rope_apply (the reference-style path of
    modules/wan.py, itself checked against the reference's complex RoPE)."""
    from mhla_b200.autograd import wan_prep_reference
    from mhla_b200.modules.wan import WanRMSNorm, rope_apply, _rope_tables
    torch.manual_seed(12)
    grid, H, D = (2, 3, 4), 2, 12
    N = grid[0] * grid[1] * grid[2]
    x = torch.randn(2, N, H * D)
    norm = WanRMSNorm(H * D, eps=1e-5)
    norm.weight.data = torch.rand(H * D) + 0.5
    freqs = oracle.rope_freqs_wan(D)
    ref = rope_apply((torch.relu(norm(x)) + 1e-6).view(2, N, H, D), torch.tensor([list(grid)] * 2), freqs)
    cos, sin = _rope_tables(grid, freqs, x.device)
    got = wan_prep_reference(x, norm.weight.detach(), cos, sin, D, eps_norm=1e-5, eps=1e-6)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)
