"""GPU tests of the native backward (mhla_b200/autograd.py): the gradient contractions run as launches of the FORWARD
CUDA kernels with permuted / time-reversed operands.  Checked against torch.autograd of the CPU oracle (fp32, same
bf16-rounded inputs) and against the plain-torch statement of the same gradients.  Tolerance: gradients pass through two
16-bit roundings (dO / den, the kernel's own output), so the RMS bound is 2e-2 of the gradient's RMS."""
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _grads_oracle(q, k, v, W, do, qr, kr, normalize):
    ins = [t.float().clone().requires_grad_(True) if t is not None else None for t in (q, k, v, W, qr, kr)]
    out = oracle.blockmix_fwd(ins[0], ins[1], ins[2], ins[3], eps=1e-6, normalize=normalize, q_rope=ins[4], k_rope=ins[5])
    out.backward(do.float())
    return [None if t is None else t.grad for t in ins]


@pytest.mark.parametrize("B,H,M,w,D,normalize,rope", [
    (1, 2, 16, 64, 64, True, False),      # general kernel (N = 1024 per unit), packed mixing tile
    (2, 2, 128, 256, 64, True, False),    # headline block shape
    (1, 2, 128, 256, 64, False, False),
    (2, 6, 16, 16, 64, True, False),      # DiT-S/2: short-sequence kernel (no workspace: summaries for dW from cuBLAS)
    (1, 2, 20, 210, 128, False, True),    # Wan-shaped blocks, shipped config (no normaliser), roped numerator
    (1, 2, 20, 210, 128, True, True),     # ... with the normaliser
    (1, 2, 9, 64, 72, True, False),       # DiT-XL head dim (zero-padded to 128)
])
def test_blockmix_native_backward(B, H, M, w, D, normalize, rope):
    import mhla_b200
    from mhla_b200 import autograd
    assert not autograd._TORCH_BACKWARD
    g = torch.Generator().manual_seed(11)
    mk = lambda relu: ((torch.relu(torch.randn(B, H, M, w, D, generator=g)) + 1e-2) if relu  # noqa: E731
                       else torch.randn(B, H, M, w, D, generator=g)).bfloat16()
    q, k, v = mk(True), mk(True), mk(False)
    qr, kr = (mk(False), mk(False)) if rope else (None, None)
    W = torch.rand(M, M, generator=g) / M + 0.3 * torch.eye(M)
    do = torch.randn(B, H, M, w, D, generator=g).bfloat16()
    ref = _grads_oracle(q, k, v, W, do, qr, kr, normalize)
    dev = [None if t is None else t.cuda().requires_grad_(True) for t in (q, k, v, W, qr, kr)]
    before = mhla_b200.last_launch_count()
    out = mhla_b200.mhla(dev[0], dev[1], dev[2], dev[3], q_rope=dev[4], k_rope=dev[5], normalize=normalize)
    out.backward(do.cuda())
    torch.cuda.synchronize()
    assert before >= 0
    names = ("dq", "dk", "dv", "dW", "dq_rope", "dk_rope")
    for name, r, t in zip(names, ref, dev):
        if t is None:
            continue
        if r is None or float(r.abs().max()) == 0.0:     # un-roped q, k without the normaliser
            assert t.grad is None or float(t.grad.abs().max()) == 0.0, name
            continue
        err = oracle.err_ratio(r, t.grad.float().cpu())
        assert err < 2e-2, (name, err)


def test_blockmix_native_backward_equals_torch_statement():
    """Same inputs through both statements of the gradient on the GPU: native launches vs fp32 cuBLAS einsums."""
    from mhla_b200 import autograd, ops
    g = torch.Generator().manual_seed(12)
    G, M, w, D = 4, 32, 256, 64
    q = (torch.relu(torch.randn(G, M, w, D, generator=g)) + 1e-2).bfloat16().cuda()
    k = (torch.relu(torch.randn(G, M, w, D, generator=g)) + 1e-2).bfloat16().cuda()
    v = torch.randn(G, M, w, D, generator=g).bfloat16().cuda()
    W = (torch.rand(M, M, generator=g) / M + 0.3 * torch.eye(M)).cuda()
    do = torch.randn(G, M, w, D, generator=g).bfloat16().cuda()
    out = ops._blockmix_fwd(q, k, v, W, normalize=True)
    a = autograd.blockmix_backward_native(q, k, v, W, do, out, normalize=True)
    b = autograd.blockmix_backward(q, k, v, W, do, normalize=True)
    for x, y in zip(a[:4], b[:4]):
        assert oracle.err_ratio(y.float().cpu(), x.float().cpu()) < 1e-2


@pytest.mark.parametrize("B,T,H,K,V", [(2, 256, 2, 64, 64), (2, 512, 2, 128, 256), (1, 200, 2, 32, 48)])
def test_causal_native_backward(B, T, H, K, V):
    import mhla_b200
    g = torch.Generator().manual_seed(13)
    q, k, v = (torch.randn(B, T, H, d, generator=g).bfloat16() for d in (K, K, V))
    mm = torch.clamp(torch.rand(32, 32, generator=g), 1e-5, 1).tril()
    do = torch.randn(B, T, H, V, generator=g).bfloat16()
    qc, kc, vc, mc = (t.float().clone().requires_grad_(True) for t in (q, k, v, mm))
    oracle.causal_chunk_fwd(qc, kc, vc, mc).backward(do.float())
    qg, kg, vg = (t.cuda().requires_grad_(True) for t in (q, k, v))
    mg = mm.cuda().requires_grad_(True)
    out = mhla_b200.mhla_causal(qg, kg, vg, mg)
    out.backward(do.cuda())
    torch.cuda.synchronize()
    for name, ref, got in (("dq", qc.grad, qg.grad), ("dk", kc.grad, kg.grad), ("dv", vc.grad, vg.grad), ("dmm", mc.grad, mg.grad)):
        err = oracle.err_ratio(ref, got.float().cpu())
        assert err < 2e-2, (name, err)


def test_blockmix_grid_native_backward():
    """Token-major 3-D block view (Wan): autograd.BlockmixGridFunction against torch.autograd of the oracle on the
    block-major rearrangement (mhla_utils.py:317-326)."""
    import mhla_b200
    from einops import rearrange
    g = torch.Generator().manual_seed(14)
    B, nh, D, grid, layout = 1, 2, 128, (6, 12, 10), (3, 2, 2)
    N = grid[0] * grid[1] * grid[2]
    M = layout[0] * layout[1] * layout[2]
    q, k, v, do = (torch.randn(B, N, nh, D, generator=g).bfloat16() for _ in range(4))
    W = torch.rand(M, M, generator=g) / M + 0.3 * torch.eye(M)
    pat = "b (fb p1 hb p2 wb p3) h d -> b h (fb hb wb) (p1 p2 p3) d"
    kw = dict(fb=layout[0], hb=layout[1], wb=layout[2], p1=grid[0] // layout[0], p2=grid[1] // layout[1], p3=grid[2] // layout[2])
    qc, kc, vc, Wc = (t.float().clone().requires_grad_(True) for t in (q, k, v, W))
    out = oracle.blockmix_fwd(rearrange(qc, pat, **kw), rearrange(kc, pat, **kw), rearrange(vc, pat, **kw), Wc, normalize=False)
    out.backward(rearrange(do.float(), pat, **kw))
    qg, kg, vg, Wg = (t.cuda().requires_grad_(True) for t in (q, k, v, W))
    og = mhla_b200.mhla_blockmix_grid(qg, kg, vg, Wg, grid, layout, normalize=False)
    og.backward(do.cuda())
    torch.cuda.synchronize()
    for name, r, t in (("dq", qc.grad, qg.grad), ("dk", kc.grad, kg.grad), ("dv", vc.grad, vg.grad), ("dW", Wc.grad, Wg.grad)):
        err = oracle.err_ratio(r, t.float().cpu())
        assert err < 2e-2, (name, err)
