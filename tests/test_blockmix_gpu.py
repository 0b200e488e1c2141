"""GPU parity tests for the block-mixed operator (variants A/B): CUDA path (through the C ABI) vs the CPU oracle
and the committed golden fixtures.  Tolerance: kernel output (bf16) vs the fp32 oracle evaluated on
the same bf16-rounded inputs: RMS error ratio <= 5e-3 and max-abs error <= 2e-2 * max|ref| for bf16
(<= 1e-3 / 5e-3 for fp16)."""
import pytest
import torch

import oracle
from conftest import load_golden

pytestmark = pytest.mark.gpu

TOL = {torch.bfloat16: (5e-3, 2e-2), torch.float16: (1e-3, 5e-3)}


def _check(ref, out, dtype, scale=1.0):
    rms, mx = TOL[dtype]
    out = out.float().cpu()
    assert not torch.isnan(out).any()
    assert oracle.err_ratio(ref, out) <= rms * scale
    assert float((ref - out).abs().max()) <= mx * scale * float(ref.abs().max())


def _inputs(B, H, M, w, D, dtype, seed=0, rope=False):
    g = torch.Generator().manual_seed(seed)
    q = (torch.relu(torch.randn(B, H, M, w, D, generator=g)) + 1e-6).to(dtype)
    k = (torch.relu(torch.randn(B, H, M, w, D, generator=g)) + 1e-6).to(dtype)
    v = torch.randn(B, H, M, w, D, generator=g).to(dtype)
    qr = kr = None
    if rope:
        qr = torch.randn(B, H, M, w, D, generator=g).to(dtype)
        kr = torch.randn(B, H, M, w, D, generator=g).to(dtype)
    return q, k, v, qr, kr


def _run(q, k, v, W, qr=None, kr=None, normalize=True, eps=1e-6, **kw):
    import mhla_b200
    dev = "cuda"
    args = [t.to(dev) for t in (q, k, v)]
    out = mhla_b200.mhla(*args, W.to(dev), q_rope=None if qr is None else qr.to(dev),
                         k_rope=None if kr is None else kr.to(dev), eps=eps, normalize=normalize, **kw)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("B,H,M,w,D,normalize,rope,dtype", [
    (1, 1, 2, 128, 64, True, False, torch.bfloat16),
    (1, 4, 16, 64, 64, True, False, torch.bfloat16),        # BASELINE cfg1: B=1 H=4 N=1024 D=64
    (2, 6, 16, 16, 64, True, False, torch.bfloat16),        # BASELINE cfg2: DiT-S/2, N=256
    (2, 4, 32, 256, 64, True, False, torch.bfloat16),
    (2, 4, 32, 256, 64, False, False, torch.bfloat16),
    (1, 2, 6, 210, 128, False, True, torch.bfloat16),       # Wan-shaped block (w=210, D=128), shipped: no normaliser
    (1, 2, 6, 210, 128, True, True, torch.bfloat16),
    (1, 3, 150, 48, 64, True, False, torch.bfloat16),       # M=150 (Wan block count): M not a multiple of 4/32/128
    (1, 2, 5, 100, 64, True, False, torch.float16),         # ragged w, fp16
    (1, 1, 1, 256, 64, True, False, torch.bfloat16),        # single block
])
@pytest.mark.parametrize("path", ["default", "no_smalln", "three_launch"])
def test_blockmix_vs_oracle(B, H, M, w, D, normalize, rope, dtype, path):
    """Every launch structure of the C ABI against the oracle: the default (the short-sequence kernel for units of at most
    256 tokens, else the single fused kernel), the fused general kernel forced (no_smalln) and the three PDL-chained phase
    launches."""
    q, k, v, qr, kr = _inputs(B, H, M, w, D, dtype, rope=rope)
    g = torch.Generator().manual_seed(1)
    W = torch.rand(M, M, generator=g) / M + 0.5 * torch.eye(M) / M
    out = _run(q, k, v, W, qr, kr, normalize=normalize, **({} if path == "default" else {path: True}))
    ref = oracle.blockmix_fwd(q, k, v, W, normalize=normalize, q_rope=qr, k_rope=kr)
    _check(ref, out, dtype)


@pytest.mark.parametrize("name", ["a_dit_s2", "a_qknorm", "b_norm", "b_nonorm"])
def test_blockmix_vs_reference_golden(name):
    """Inputs and outputs produced by the reference's own code (tests/golden/make_golden.py).  The fixtures are fp32;
    the kernel computes in bf16, so compare against the oracle on bf16-rounded inputs AND against the reference's fp32
    output with the rounding of the inputs added to the budget."""
    g = load_golden(name)
    bf = torch.bfloat16          # (D = 32 fixtures run zero-padded to 64 channels inside the shim)
    q, k, v = g["q"].to(bf), g["k"].to(bf), g["v"].to(bf)
    qr = g["q_rope"].to(bf) if "q_rope" in g else None
    kr = g["k_rope"].to(bf) if "k_rope" in g else None
    normalize = bool(g.get("normalize_out", 1))
    out = _run(q, k, v, g["W"], qr, kr, normalize=normalize, eps=g["eps"])
    ref = oracle.blockmix_fwd(q, k, v, g["W"], eps=g["eps"], normalize=normalize, q_rope=qr, k_rope=kr)
    _check(ref, out, bf)
    _check(g["out"], out, bf, scale=3.0)   # vs the reference's fp32 result on un-rounded inputs


def test_blockmix_strided_token_major_view():
    """q,k,v as views of a token-major [B, N, H, D] tensor (blocks are contiguous token ranges): consumed in place."""
    B, H, M, w, D = 2, 3, 4, 128, 64
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, B, M * w, H, D, generator=g).to(torch.bfloat16).cuda()
    q, k, v = (x[i].view(B, M, w, H, D).permute(0, 3, 1, 2, 4) for i in range(3))
    q, k = torch.relu(q) + 1e-6, torch.relu(k) + 1e-6        # makes q,k contiguous copies; v stays a strided view
    W = torch.rand(M, M, generator=g) / M
    import mhla_b200
    out = mhla_b200.mhla(q, k, v, W.cuda())
    ref = oracle.blockmix_fwd(q.cpu(), k.cpu(), v.cpu(), W)
    _check(ref, out, torch.bfloat16)


def test_blockmix_properties_full_size():
    """BASELINE full size (B=2, H=16, N=32768, D=64, w=256) through size-independent properties:
    (i) W = I  -> block-local linear attention, checked exactly on a few sampled blocks;
    (ii) linearity in v;  (iii) (b,h)-shard equivalence: a slice of the batch gives bit-identical results."""
    import mhla_b200
    B, H, M, w, D = 2, 16, 128, 256, 64
    g = torch.Generator(device="cuda").manual_seed(0)
    q = (torch.relu(torch.randn(B, H, M, w, D, generator=g, device="cuda")) + 1e-6).bfloat16()
    k = (torch.relu(torch.randn(B, H, M, w, D, generator=g, device="cuda")) + 1e-6).bfloat16()
    v = torch.randn(B, H, M, w, D, generator=g, device="cuda").bfloat16()
    eye = torch.eye(M, device="cuda")
    out = mhla_b200.mhla(q, k, v, eye, normalize=True)
    for (b, h, j) in [(0, 0, 0), (1, 7, 63), (1, 15, 127)]:
        ref = oracle.blockmix_fwd(q[b, h, j][None].cpu(), k[b, h, j][None].cpu(), v[b, h, j][None].cpu(), torch.eye(1))
        _check(ref[0], out[b, h, j], torch.bfloat16)
    Wm = oracle.block_distance_matrix((M, 1, 1), "linear").cuda()
    o1 = mhla_b200.mhla(q, k, v, Wm, normalize=False).float()
    o2 = mhla_b200.mhla(q, k, (2 * v), Wm, normalize=False).float()
    assert oracle.err_ratio(2 * o1.cpu(), o2.cpu()) < 1e-6       # scaling by 2 is exact in bf16
    osl = mhla_b200.mhla(q[1:, 4:8], k[1:, 4:8], v[1:, 4:8], Wm, normalize=True)
    ofull = mhla_b200.mhla(q, k, v, Wm, normalize=True)
    assert torch.equal(osl, ofull[1:, 4:8])
    # spot-check a few rows of the dense-W result against the oracle on one (b,h)
    ref = oracle.blockmix_fwd(q[1, 5].cpu(), k[1, 5].cpu(), v[1, 5].cpu(), Wm.cpu(), normalize=True)
    _check(ref, ofull[1, 5], torch.bfloat16)


def test_blockmix_permutation_catches_fixed_normaliser():
    """Permuting tokens inside ONE block changes the reference's quirky normaliser (it pairs equal in-block indices
    across blocks) - a 'fixed' textbook normaliser would be invariant.  SURVEY.md section 4."""
    B, H, M, w, D = 1, 1, 4, 128, 64
    q, k, v, _, _ = _inputs(B, H, M, w, D, torch.bfloat16, seed=5)
    W = oracle.block_distance_matrix((2, 2), "linear")
    perm = torch.randperm(w, generator=torch.Generator().manual_seed(0))
    q2, k2, v2 = q.clone(), k.clone(), v.clone()
    q2[:, :, 1], k2[:, :, 1], v2[:, :, 1] = q[:, :, 1][:, :, perm], k[:, :, 1][:, :, perm], v[:, :, 1][:, :, perm]
    o1, o2 = _run(q, k, v, W).float().cpu(), _run(q2, k2, v2, W).float().cpu()
    r1, r2 = oracle.blockmix_fwd(q, k, v, W), oracle.blockmix_fwd(q2, k2, v2, W)
    _check(r1, o1, torch.bfloat16)
    _check(r2, o2, torch.bfloat16)
    assert oracle.err_ratio(r1[:, :, 0], r2[:, :, 0]) > 1e-3      # block 0's outputs change although its tokens did not


@pytest.mark.parametrize("name,B,H,M,w,D,normalize,rope", [
    ("cfg2 DiT-S/2 256x256, batch 64", 64, 6, 16, 16, 64, True, False),          # BASELINE cfg2 at a realistic batch
    ("cfg4 Wan2.1-1.3B 81x480x800, shipped (no normaliser)", 1, 12, 150, 210, 128, False, True),
    ("cfg4 Wan2.1-1.3B, normaliser on, CFG batch 2", 2, 12, 150, 210, 128, True, True),
])
def test_blockmix_full_size_configs(name, B, H, M, w, D, normalize, rope):
    """BASELINE configs at their full sizes: the whole batch runs on the GPU, a few (b,h) units are checked against the
    oracle (the oracle is per-unit independent, so the rest is covered by the bitwise shard-equivalence test)."""
    import mhla_b200
    g = torch.Generator(device="cuda").manual_seed(3)
    mk = lambda relu: (torch.relu(torch.randn(B, H, M, w, D, generator=g, device="cuda")) + 1e-6 if relu  # noqa: E731
                       else torch.randn(B, H, M, w, D, generator=g, device="cuda")).bfloat16()
    q, k, v = mk(True), mk(True), mk(False)
    qr, kr = (mk(False), mk(False)) if rope else (None, None)
    W = torch.rand(M, M, generator=torch.Generator().manual_seed(4)) / M + 0.5 * torch.eye(M) / M
    out = mhla_b200.mhla(q, k, v, W.cuda(), q_rope=qr, k_rope=kr, normalize=normalize)
    torch.cuda.synchronize()
    assert not torch.isnan(out).any()
    for (b, h) in {(0, 0), (B - 1, H - 1), (B // 2, H // 2)}:
        sl = lambda t: None if t is None else t[b, h][None].cpu()  # noqa: E731
        ref = oracle.blockmix_fwd(sl(q), sl(k), sl(v), W, normalize=normalize, q_rope=sl(qr), k_rope=sl(kr))
        _check(ref[0], out[b, h], torch.bfloat16)


@pytest.mark.parametrize("B,H,M,w,D,normalize,rope", [(1, 2, 6, 210, 128, False, True), (2, 3, 8, 256, 64, True, False)])
def test_blockmix_fused_output_rmsnorm(B, H, M, w, D, normalize, rope):
    """Fused per-(token, head) RMSNorm of the output (MHLA_Video_Uni's g_norm, mhla_utils.py:360-362 with WanRMSNorm
    wan/model.py:181-196) against the oracle followed by the same normalisation in fp32."""
    q, k, v, qr, kr = _inputs(B, H, M, w, D, torch.bfloat16, seed=21, rope=rope)
    g = torch.Generator().manual_seed(22)
    W = torch.rand(M, M, generator=g) / M
    wgt = torch.rand(D, generator=g) + 0.5
    eps = 1e-5
    out = _run(q, k, v, W, qr, kr, normalize=normalize, out_rms_weight=wgt.cuda(), out_rms_eps=eps)
    ref = oracle.blockmix_fwd(q, k, v, W, normalize=normalize, q_rope=qr, k_rope=kr)
    ref = ref * torch.rsqrt(ref.pow(2).mean(dim=-1, keepdim=True) + eps) * wgt
    _check(ref, out, torch.bfloat16)


@pytest.mark.parametrize("B,H,M,w,normalize,dtype", [
    (2, 6, 16, 16, True, torch.bfloat16),      # DiT-S/2 256x256 (BASELINE cfg2)
    (3, 2, 4, 49, True, torch.bfloat16),       # the modules' default block_size = 49 (N = 196: ragged second row tile)
    (1, 3, 3, 7, True, torch.bfloat16),        # N = 21: one partial row tile
    (2, 2, 64, 4, True, torch.bfloat16),       # M = 64 blocks of 4 tokens
    (1, 2, 16, 8, False, torch.bfloat16),      # N = 128: exactly one row tile, no normaliser
    (1, 2, 37, 1, True, torch.bfloat16),       # one token per block
    (2, 2, 1, 256, True, torch.bfloat16),      # a single block of 256 tokens
    (2, 3, 4, 64, True, torch.float16),        # fp16
    (150, 6, 16, 16, True, torch.bfloat16),    # 900 units: several units per CTA (ring reuse, barrier phases)
])
def test_smalln_kernel_vs_oracle_and_general_kernel(B, H, M, w, normalize, dtype):
    """The short-sequence kernel (csrc/smalln_kernel.cuh) against the oracle, and against the general kernel on the
    same inputs (two independent formulations of mhla.py:262-268)."""
    import mhla_b200
    D = 64
    q, k, v, _, _ = _inputs(B, H, M, w, D, dtype, seed=41)
    g = torch.Generator().manual_seed(42)
    W = torch.rand(M, M, generator=g) / M + 0.5 * torch.eye(M) / M
    out = _run(q, k, v, W, normalize=normalize)
    assert mhla_b200.last_launch_count() == 1
    sel = slice(0, min(B, 3))
    _check(oracle.blockmix_fwd(q[sel], k[sel], v[sel], W, normalize=normalize), out[sel], dtype)
    _check(oracle.blockmix_fwd(q[-1:], k[-1:], v[-1:], W, normalize=normalize), out[-1:], dtype)
    gen = _run(q, k, v, W, normalize=normalize, no_smalln=True)
    assert oracle.err_ratio(gen.float().cpu(), out.float().cpu()) < 6e-3
    again = _run(q, k, v, W, normalize=normalize)
    assert torch.equal(out, again)                       # deterministic
    if B > 1:
        assert torch.equal(_run(q[1:], k[1:], v[1:], W, normalize=normalize), out[1:])   # unit-shard equivalence


def test_blockmix_padded_head_dim():
    """DiT-XL heads (1152 / 16 = 72 channels, mhla_dit/models.py:478-549) run zero-padded to 128 in the shim."""
    B, H, M, w, D = 1, 2, 16, 16, 72
    q, k, v, _, _ = _inputs(B, H, M, w, D, torch.bfloat16, seed=31)
    W = oracle.block_distance_matrix((4, 4), "linear")
    out = _run(q, k, v, W)
    assert out.shape[-1] == D
    _check(oracle.blockmix_fwd(q, k, v, W), out, torch.bfloat16)


def test_host_pipeline_matches_device_call():
    """mhla_host (pinned host tensors, copies and kernels pipelined over ranges of (b,h) units) is bit-identical to one
    device call on the whole batch - the units are independent."""
    import mhla_b200
    B, H, M, w, D = 2, 6, 16, 64, 64
    q, k, v, _, _ = _inputs(B, H, M, w, D, torch.bfloat16, seed=11)
    W = oracle.block_distance_matrix((4, 4), "linear")
    ref = _run(q, k, v, W)
    for chunks in (1, 4, 5):
        out = mhla_b200.mhla_host(q.pin_memory(), k.pin_memory(), v.pin_memory(), W.cuda(), chunks=chunks)
        torch.cuda.synchronize()
        assert torch.equal(out, ref.cpu())


def test_errors_are_loud():
    import mhla_b200
    q = torch.zeros(1, 1, 2, 16, 160, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(Exception):
        mhla_b200.mhla(q, q, q, torch.eye(2, device="cuda"))          # D = 160 outside the envelope (D <= 128)
    q = torch.zeros(1, 1, 2, 16, 32, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):
        mhla_b200.mhla(q.cpu(), q.cpu(), q.cpu(), torch.eye(2))       # CPU tensors: no fallback


def test_cuda_graph_capture_of_a_dit_stack():
    """28 DiT-S/2 layers x batch 2 (BASELINE cfg2) captured into ONE CUDA graph: the operator only enqueues on the current
    stream (no allocation of its own after warm-up, no sync), so a whole sampler step can be replayed without the Python /
    ctypes enqueue cost.  The replayed result must equal the eagerly computed one bit for bit."""
    import mhla_b200
    B, H, M, w, D, layers = 2, 6, 16, 16, 64, 28
    q, k, v, _, _ = _inputs(B, H, M, w, D, torch.bfloat16, seed=51)
    q, k, v = q.cuda(), k.cuda(), v.cuda()
    Ws = [(torch.rand(M, M, generator=torch.Generator().manual_seed(60 + i)) / M).cuda() for i in range(layers)]
    outs = [torch.empty_like(q) for _ in range(layers)]
    eager = [mhla_b200.mhla(q, k, v, Ws[i]).clone() for i in range(layers)]     # also warms the descriptor caches
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(layers):
            mhla_b200.mhla(q, k, v, Ws[i], out=outs[i])
    for o in outs:
        o.zero_()
    g.replay()
    torch.cuda.synchronize()
    for i in range(layers):
        assert torch.equal(outs[i], eager[i])
    # the general kernel (persistent self-cleaning workspace) is capturable as well
    qg, kg, vg, _, _ = _inputs(1, 2, 8, 128, 64, torch.bfloat16, seed=52)
    qg, kg, vg = qg.cuda(), kg.cuda(), vg.cuda()
    Wg = (torch.rand(8, 8) / 8).cuda()
    ref = mhla_b200.mhla(qg, kg, vg, Wg).clone()
    og = torch.empty_like(qg)
    torch.cuda.synchronize()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        for _ in range(3):
            mhla_b200.mhla(qg, kg, vg, Wg, out=og)
    g2.replay()
    torch.cuda.synchronize()
    assert torch.equal(og, ref)


@pytest.mark.parametrize("B,nh,D,grid,layout,normalize,rope", [
    (2, 2, 64, (2, 4, 8), (1, 2, 2), False, True),       # the b_nonorm fixture's geometry: 4 blocks of 16 tokens
    (1, 3, 128, (6, 10, 20), (3, 5, 10), False, True),   # Wan-like: 150 blocks of 2*2*2 = 8 tokens
    (1, 2, 128, (7, 12, 10), (1, 2, 2), True, True),     # Wan's block shape: 7*6*5 = 210 tokens (two sub-tiles, ragged tail)
    (2, 2, 128, (14, 6, 10), (2, 1, 2), True, True),     # batch 2 (the (b, f) axis of the tensor map), 7*6*5 tokens
    (1, 2, 64, (8, 4, 4), (2, 2, 2), True, False),       # 4*2*2 = 16 tokens, no rope, D = 64
    (1, 1, 64, (3, 8, 16), (1, 1, 1), True, False),      # one block of 3*8*16 = 384 tokens? no: exceeds 256 -> see below
])
@pytest.mark.parametrize("three_launch", [False, True])
def test_blockmix_3d_block_view(B, nh, D, grid, layout, normalize, rope, three_launch):
    """Token-major [B, N, heads, D] tensors consumed in place through the 3-D block TMA view against the oracle on the
    rearranged (block-major) tensors - the reference's own layout transformation, mhla_utils.py:317-326 / :345-354."""
    import mhla_b200
    from einops import rearrange
    F_, H_, W_ = grid
    fb, hb, wb = layout
    p1, p2, p3 = F_ // fb, H_ // hb, W_ // wb
    N, M, w = F_ * H_ * W_, fb * hb * wb, p1 * p2 * p3
    g = torch.Generator().manual_seed(7)
    mk = lambda relu: ((torch.relu(torch.randn(B, N, nh, D, generator=g)) + 1e-6) if relu  # noqa: E731
                       else torch.randn(B, N, nh, D, generator=g)).bfloat16()
    q, k, v = mk(True), mk(True), mk(False)
    qr, kr = (mk(False), mk(False)) if rope else (None, None)
    W = torch.rand(M, M, generator=g) / M + 0.5 * torch.eye(M) / M
    cu = lambda t: None if t is None else t.cuda()  # noqa: E731
    if w > 256:
        with pytest.raises(Exception):
            mhla_b200.mhla_blockmix_grid(cu(q), cu(k), cu(v), W.cuda(), grid, layout, normalize=normalize)
        return
    out = mhla_b200.mhla_blockmix_grid(cu(q), cu(k), cu(v), W.cuda(), grid, layout, q_rope=cu(qr), k_rope=cu(kr),
                                       normalize=normalize, three_launch=three_launch)
    torch.cuda.synchronize()
    pat = "b (fb p1 hb p2 wb p3) h d -> b h (fb hb wb) (p1 p2 p3) d"
    kw = dict(fb=fb, hb=hb, wb=wb, p1=p1, p2=p2, p3=p3)
    blk = lambda t: None if t is None else rearrange(t, pat, **kw)  # noqa: E731
    ref = oracle.blockmix_fwd(blk(q), blk(k), blk(v), W, normalize=normalize, q_rope=blk(qr), k_rope=blk(kr))
    ref = rearrange(ref, "b h (fb hb wb) (p1 p2 p3) d -> b (fb p1 hb p2 wb p3) h d", **kw)
    _check(ref, out, torch.bfloat16)
    # strided views of one fused [B, N, 3, heads, D] projection output are consumed in place as well
    if not rope:
        qkv = torch.stack([q, k, v], dim=2).cuda()
        out2 = mhla_b200.mhla_blockmix_grid(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], W.cuda(), grid, layout,
                                            normalize=normalize, three_launch=three_launch)
        assert torch.equal(out2, out)


def _post(ref, gate, add, wgt=None, eps=1e-5):
    """out = (o [rms-normed per row, * wgt]) * silu(gate) + add in fp32 (mhla_utils.py:357-364)."""
    if wgt is not None:
        ref = ref * torch.rsqrt(ref.pow(2).mean(dim=-1, keepdim=True) + eps) * wgt
    if gate is not None:
        ref = ref * torch.nn.functional.silu(gate.float())
    if add is not None:
        ref = ref + add.float()
    return ref


@pytest.mark.parametrize("B,H,M,w,D,normalize,rms,gate,add", [
    (2, 2, 8, 256, 64, True, False, True, True),      # general kernel, D = 64
    (1, 2, 20, 210, 128, False, True, True, True),    # Wan block shape, per-head norm + gate + lepe
    (1, 2, 20, 210, 128, True, False, False, True),   # additive term only
    (2, 6, 16, 16, 64, True, False, False, True),     # DiT-S/2: "+ lepe" inside the short-sequence kernel
    (3, 2, 4, 49, 64, True, False, False, True),      # ... ragged second row tile
    (2, 6, 16, 16, 64, True, False, True, True),      # a gate at a short-sequence shape takes the general kernel
])
def test_blockmix_fused_gate_and_add(B, H, M, w, D, normalize, rms, gate, add):
    """ABI v4 post-ops of the readout epilogue: out = o * silu(out_gate) + out_add (after the optional per-head RMS norm),
    the gate / add tensors being permuted views of [B, M, w, H*D] tensors (as the modules pass them)."""
    q, k, v, _, _ = _inputs(B, H, M, w, D, torch.bfloat16, seed=21)
    g = torch.Generator().manual_seed(22)
    W = torch.rand(M, M, generator=g) / M
    wgt = (torch.rand(D, generator=g) + 0.5) if rms else None
    gt = torch.randn(B, M, w, H * D, generator=g).bfloat16() if gate else None
    ad = torch.randn(B, M, w, H * D, generator=g).bfloat16() if add else None
    view = lambda t: None if t is None else t.cuda().view(B, M, w, H, D).permute(0, 3, 1, 2, 4)   # noqa: E731
    kw = dict(out_rms_weight=wgt.cuda(), out_rms_eps=1e-5) if rms else {}
    out = _run(q, k, v, W, normalize=normalize, out_gate=view(gt), out_add=view(ad), **kw)
    ref = oracle.blockmix_fwd(q, k, v, W, normalize=normalize)
    cpuview = lambda t: None if t is None else t.view(B, M, w, H, D).permute(0, 3, 1, 2, 4)   # noqa: E731
    ref = _post(ref, cpuview(gt), cpuview(ad), wgt)
    _check(ref, out, torch.bfloat16)


@pytest.mark.parametrize("B,nh,D,grid,layout,normalize", [
    (1, 2, 128, (7, 12, 10), (1, 2, 2), False),     # Wan's block shape (two sub-tiles, ragged tail)
    (2, 2, 128, (14, 6, 10), (2, 1, 2), True),      # batch 2, with the normaliser
    (1, 2, 64, (8, 4, 4), (2, 2, 2), True),         # D = 64
])
def test_blockmix_3d_block_view_fused_gate_and_add(B, nh, D, grid, layout, normalize):
    """Token-major 3-D block view with the fused per-head norm, SiLU gate and additive term: gate / add are plain
    [B, N, heads*D] projection-shaped tensors viewed as [B, N, heads, D] (what MHLA_Video_Uni._forward_fused passes)."""
    import mhla_b200
    from einops import rearrange
    F_, H_, W_ = grid
    fb, hb, wb = layout
    p1, p2, p3 = F_ // fb, H_ // hb, W_ // wb
    N, M = F_ * H_ * W_, fb * hb * wb
    g = torch.Generator().manual_seed(23)
    mk = lambda relu: ((torch.relu(torch.randn(B, N, nh, D, generator=g)) + 1e-6) if relu  # noqa: E731
                       else torch.randn(B, N, nh, D, generator=g)).bfloat16()
    q, k, v, gt, ad = mk(True), mk(True), mk(False), mk(False), mk(False)
    W = torch.rand(M, M, generator=g) / M + 0.5 * torch.eye(M) / M
    wgt = torch.rand(D, generator=g) + 0.5
    out = mhla_b200.mhla_blockmix_grid(q.cuda(), k.cuda(), v.cuda(), W.cuda(), grid, layout, normalize=normalize,
                                       out_rms_weight=wgt.cuda(), out_rms_eps=1e-5, out_gate=gt.cuda(), out_add=ad.cuda())
    torch.cuda.synchronize()
    pat = "b (fb p1 hb p2 wb p3) h d -> b h (fb hb wb) (p1 p2 p3) d"
    kw = dict(fb=fb, hb=hb, wb=wb, p1=p1, p2=p2, p3=p3)
    blk = lambda t: rearrange(t, pat, **kw)  # noqa: E731
    ref = oracle.blockmix_fwd(blk(q), blk(k), blk(v), W, normalize=normalize)
    ref = rearrange(ref, "b h (fb hb wb) (p1 p2 p3) d -> b (fb p1 hb p2 wb p3) h d", **kw)
    ref = _post(ref, gt, ad, wgt)
    _check(ref, out, torch.bfloat16)
