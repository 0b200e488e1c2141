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
