"""Drop-in modules: constructor / state_dict compatibility with the reference (CPU) and forward parity against the
reference modules' recorded outputs (GPU; fixtures from tests/golden/make_golden.py)."""
import pytest
import torch

import oracle
from conftest import load_golden
from mhla_b200.modules import MHLA, MHLA4DiT, MHLA_Normed_Torch, MHLA_Video_Uni, WAN_SELFATTENTION_CLASSES, rope_apply


def _sd(g):
    return {k[3:]: v for k, v in g.items() if k.startswith("sd.")}


def _build(name, g):
    if name in ("a_dit_s2", "a_qknorm"):
        dim = g["x"].shape[-1]
        return MHLA4DiT(dim, heads=int(g["heads"]), dropout=0.0, qk_norm=bool(g["qk_norm"]),
                        block_size=int(g["block_size"]), embed_len=int(g["embed_len"]), qkv_bias=True)
    if name == "a_vit_twin":
        return MHLA_Normed_Torch(g["x"].shape[-1], heads=int(g["heads"]), dropout=0.0, qk_norm=True,
                                 window_size=int(g["window_size"]), embed_len=int(g["embed_len"]))
    if name in ("b_norm", "b_nonorm"):
        return MHLA_Video_Uni(g["x"].shape[-1], int(g["heads"]), None, 0.0, None, True,
                              tuple(int(v) for v in g["layout"]), normalize_out=bool(g["normalize_out"]),
                              is_gated=bool(g["gated"]))
    if name.startswith("bp_"):      # the positional call of WanAttentionBlock (wan/model.py:1644-1646)
        return WAN_SELFATTENTION_CLASSES[name[3:]](
            g["x"].shape[-1], int(g["heads"]), (-1, -1), True, 1e-6, rope_after=False, without_rope=False, power=1.0,
            out_rmsnorm=bool(g["out_rmsnorm"]), normalize_out=bool(g["normalize_out"]), is_gated=False, is_lepe=False,
            block_layout=tuple(int(v) for v in g["layout"]))
    raise KeyError(name)


BPRIME = ["bp_mhla", "bp_gated_mhla", "bp_mhla_nope", "bp_mhla_lepe", "bp_gated_mhla_lepe"]


def test_wan_registry_keys():
    assert set(WAN_SELFATTENTION_CLASSES) == {"mhla", "gated_mhla", "mhla_nope", "mhla_lepe", "gated_mhla_lepe", "mhla_uni"}


@pytest.mark.parametrize("name", ["a_dit_s2", "a_qknorm", "a_vit_twin", "b_norm", "b_nonorm"] + BPRIME)
def test_state_dict_keys_match_reference(name):
    g = load_golden(name)
    m = _build(name, g)
    sd = _sd(g)
    assert set(m.state_dict().keys()) == set(sd.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd, strict=True)


def test_nlp_layer_state_dict_and_init():
    m = MHLA(mode="chunk", hidden_size=256, expand_k=0.5, expand_v=1.0, num_heads=2, feature_map="relu")
    keys = set(m.state_dict().keys())
    assert keys == {"q_proj.weight", "k_proj.weight", "v_proj.weight", "g_proj.weight", "o_proj.weight",
                    "mixing_matrix", "g_norm_swish_gate.weight"}
    mm = m.mixing_matrix.view(32, 32)
    assert tuple(m.mixing_matrix.shape) == (32, 32, 1, 1, 1, 1)
    torch.testing.assert_close(mm[5, :6], torch.full((6,), 1 / 6))
    assert float(mm.triu(1).abs().max()) == 0.0
    m2 = MHLA(hidden_size=256, num_heads=2, feature_map="relu", use_short_conv=True, use_output_gate=False)
    assert {"q_conv1d.weight", "k_conv1d.weight", "v_conv1d.weight", "g_norm.weight"} <= set(m2.state_dict().keys())
    with pytest.raises(NotImplementedError):
        MHLA(feature_map="nope")


def test_mixing_modules_match_reference_weights():
    import mhla_b200
    g2, g3 = load_golden("blockdist2d"), load_golden("blockdist3d")
    bd = mhla_b200.BlockDistanceConv(num_patches_per_side=16, patch_group_size=16, transform="linear")
    torch.testing.assert_close(bd.get_weight_matrix(), g2["W_side16_group16_linear"], rtol=1e-6, atol=1e-7)
    assert tuple(bd.conv.weight.shape) == (16, 16, 1, 1)
    bd3 = mhla_b200.BlockDistanceConv3D(blocks_layout=(3, 5, 10), transform="linear")
    torch.testing.assert_close(bd3.get_weight_matrix(), g3["W_3x5x10_linear"], rtol=1e-6, atol=1e-7)
    with pytest.raises(ValueError):
        mhla_b200.BlockDistanceConv3D(transform="bogus")


@pytest.mark.parametrize("name", ["b_norm", "b_nonorm"])
def test_rope_real_arithmetic_matches_reference(name):
    g = load_golden(name)
    d = g["q_tok"].shape[-1]
    freqs = oracle.rope_freqs_wan(d)
    grid = torch.tensor([[int(v) for v in g["grid"]]] * g["q_tok"].shape[0])
    torch.testing.assert_close(rope_apply(g["q_tok"], grid, freqs), g["q_rope_tok"], rtol=1e-5, atol=1e-5)


def test_ops_refuse_cpu_tensors():
    import mhla_b200
    q = torch.zeros(1, 1, 2, 16, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        mhla_b200.mhla(q, q, q, torch.eye(2))


# ------------------------------------------------------------------------------------------------ GPU forward parity
def _fwd_err(name, atol_scale=1.0):
    g = load_golden(name)
    m = _build(name, g).eval()
    m.load_state_dict(_sd(g), strict=True)
    m = m.cuda()
    x = g["x"].cuda()
    with torch.no_grad():
        if name.startswith("b_") or name.startswith("bp_"):
            B = x.shape[0]
            grid = torch.tensor([[int(v) for v in g["grid"]]] * B, dtype=torch.long)
            d = x.shape[-1] // int(g["heads"])
            y = m(x, torch.tensor([x.shape[1]] * B), grid, oracle.rope_freqs_wan(d))
        else:
            y = m(x)
    torch.cuda.synchronize()
    return oracle.err_ratio(g["y"], y.float().cpu())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a_dit_s2", "a_qknorm", "a_vit_twin", "b_norm", "b_nonorm"] + BPRIME)
def test_module_forward_matches_reference_module(name):
    """Whole-module forward (fp32 weights, operator in bf16) vs the reference module's recorded fp32 output.
    Budget: the operator's 5e-3 plus bf16 rounding of q,k,v (2^-9 each).  The D = 32 fixtures run zero-padded."""
    assert _fwd_err(name) < 1.5e-2


@pytest.mark.gpu
def test_dit_module_padded_head_dim_and_fp16_autocast():
    """DiT-XL heads (1152 / 16 = 72 channels) through the MODULE (the shim zero-pads; no out= view), and fp16 autocast
    (LayerNorm returns fp32, the projections fp16: the output buffer follows q's dtype)."""
    torch.manual_seed(0)
    m = MHLA4DiT(144, heads=2, dropout=0.0, block_size=16, embed_len=64, qkv_bias=True).eval().cuda()
    x = torch.randn(2, 4, 16, 144, device="cuda")
    with torch.no_grad():
        y = m(x)
        xn = m.norm(x)
        q, k, v, lepe = m._mlp_lepe(xn)
        q, k = torch.relu(q) + m.eps, torch.relu(k) + m.eps
        sp = lambda t: t.reshape(2, 4, 16, 2, 72).permute(0, 3, 1, 2, 4).float().cpu()   # noqa: E731
        o = oracle.blockmix_fwd(sp(q), sp(k), sp(v), m.piece_attn.get_weight_matrix().cpu(), eps=m.eps)
        ref = m.to_out(o.permute(0, 2, 3, 1, 4).reshape(2, 4, 16, 144).cuda() + lepe)
        assert oracle.err_ratio(ref.float().cpu(), y.float().cpu()) < 1.5e-2
        m64 = MHLA4DiT(128, heads=2, dropout=0.0, block_size=16, embed_len=64).eval().cuda()
        with torch.autocast("cuda", dtype=torch.float16):
            y16 = m64(torch.randn(2, 4, 16, 128, device="cuda"))
        assert torch.isfinite(y16.float()).all()


@pytest.mark.gpu
def test_dit_module_lepe_paths_agree_above_the_short_sequence_size():
    """MHLA4DiT at N = 1024 tokens (16 blocks of 64; the general kernel) under bf16 autocast: "+ lepe" as one streaming
    launch behind the operator (default) against the plain torch add, and at N = 256 (short-sequence kernel: lepe added in
    the readout epilogue) against the same."""
    torch.manual_seed(1)
    for embed_len, bs in ((1024, 64), (256, 16)):
        m = MHLA4DiT(128, heads=2, dropout=0.0, block_size=bs, embed_len=embed_len, qkv_bias=True).eval().cuda()
        x = torch.randn(2, embed_len // bs, bs, 128, device="cuda")
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            m.fuse_lepe = True
            y1 = m(x)
            m.fuse_lepe = False
            y2 = m(x)
        assert oracle.err_ratio(y2.float().cpu(), y1.float().cpu()) < 8e-3


@pytest.mark.gpu
def test_modules_train_through_the_operator():
    """loss.backward() reaches the projections AND the trainable mixing matrix (the reference trainers clamp
    piece_attn.conv.weight after every step, mhla_dit/train.py:308-310); gradients match the oracle's autograd."""
    torch.manual_seed(0)
    m = MHLA4DiT(128, heads=2, dropout=0.0, block_size=16, embed_len=64, qkv_bias=True).cuda()
    x = torch.randn(2, 4, 16, 128, device="cuda")
    y = m(x)
    y.square().mean().backward()
    gW = m.piece_attn.conv.weight.grad
    assert gW is not None and float(gW.abs().sum()) > 0 and m.to_qkv.weight.grad is not None

    # the same forward with the oracle in the operator's place, on CPU in fp32
    mc = MHLA4DiT(128, heads=2, dropout=0.0, block_size=16, embed_len=64, qkv_bias=True)
    mc.load_state_dict({k_: v_.cpu() for k_, v_ in m.state_dict().items()})
    xc = x.cpu()
    xn = mc.norm(xc)
    q, k, v, lepe = mc._mlp_lepe(xn)
    q, k = torch.relu(q) + mc.eps, torch.relu(k) + mc.eps
    sp = lambda t: t.reshape(2, 4, 16, 2, 64).permute(0, 3, 1, 2, 4)   # noqa: E731
    o = oracle.blockmix_fwd(sp(q), sp(k), sp(v), mc.piece_attn.conv.weight.view(4, 4), eps=mc.eps)
    yc = mc.to_out(o.permute(0, 2, 3, 1, 4).reshape(2, 4, 16, 128) + lepe)
    yc.square().mean().backward()
    assert oracle.err_ratio(mc.piece_attn.conv.weight.grad, gW.cpu()) < 3e-2
    assert oracle.err_ratio(mc.to_qkv.weight.grad, m.to_qkv.weight.grad.cpu()) < 3e-2


@pytest.mark.gpu
def test_dit_module_accepts_3d_tokens():
    g = load_golden("a_dit_s2")
    m = _build("a_dit_s2", g).eval()
    m.load_state_dict(_sd(g))
    m = m.cuda()
    x = g["x"].cuda()
    with torch.no_grad():
        y4 = m(x)
        y3 = m(x.flatten(1, 2))
    assert y3.shape == (x.shape[0], x.shape[1] * x.shape[2], x.shape[3])
    torch.testing.assert_close(y3.view_as(y4), y4)


@pytest.mark.gpu
@pytest.mark.parametrize("in_dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("plain", [False, True])
def test_wan_prep_kernel_matches_reference_preprocessing(in_dtype, plain):
    """csrc/wan_prep_kernel.cuh against the reference sequence: WanRMSNorm over the full channel dim (wan/model.py:181-196),
    relu + eps (mhla_utils.py:267-276), complex RoPE on interleaved pairs (:127-156; oracle.rope_apply_wan)."""
    import mhla_b200
    from mhla_b200.modules.wan import _rope_tables
    torch.manual_seed(0)
    B, grid, nh, D = 2, (3, 4, 6), 3, 64
    N, Cc = grid[0] * grid[1] * grid[2], nh * D
    xq, xk = torch.randn(B, N, Cc), torch.randn(B, N, Cc)
    wq, wk = 1 + 0.1 * torch.randn(Cc), 1 + 0.1 * torch.randn(Cc)
    freqs = oracle.rope_freqs_wan(D)
    xq_, xk_ = xq.to(in_dtype), xk.to(in_dtype)

    def ref(x, w):
        xf = x.float()
        y = torch.relu(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * w) + 1e-6
        y = y.view(B, N, nh, D)
        return y, oracle.rope_apply_wan(y, grid, freqs)

    (qp_ref, qr_ref), (kp_ref, kr_ref) = ref(xq_, wq), ref(xk_, wk)
    cos, sin = _rope_tables(grid, freqs, torch.device("cuda"))
    qr, kr, qp, kp = mhla_b200.wan_prep(xq_.cuda(), xk_.cuda(), wq.cuda(), wk.cuda(), cos, sin, D, want_plain=plain)
    torch.cuda.synchronize()
    assert qr.dtype == torch.bfloat16 and tuple(qr.shape) == (B, N, nh, D)
    for got, want in ((qr, qr_ref), (kr, kr_ref)) + (((qp, qp_ref), (kp, kp_ref)) if plain else ()):
        assert oracle.err_ratio(want, got.float().cpu()) < 3e-3           # bf16 output rounding (2^-9)
    if not plain:
        assert qp is None and kp is None


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,gated,added", [(torch.bfloat16, True, True), (torch.float16, True, False),
                                               (torch.bfloat16, False, True)])
def test_gate_add_kernel(dtype, gated, added):
    """csrc/gated_norm_kernel.cuh gate_add_kernel (mhla_gate_add): out = x * silu(g) + add, fp32 math, one rounding -
    the post-ops of mhla_utils.py:360-366 / wan/model.py:1001-1003 as one streaming launch; also in place."""
    import mhla_b200
    g_ = torch.Generator().manual_seed(5)
    x = torch.randn(2, 333, 12, 128, generator=g_).to(dtype)
    gt = torch.randn(2, 333, 12, 128, generator=g_).to(dtype) if gated else None
    ad = torch.randn(2, 333, 12, 128, generator=g_).to(dtype) if added else None
    ref = x.float()
    if gated:
        ref = ref * torch.nn.functional.silu(gt.float())
    if added:
        ref = ref + ad.float()
    xd = x.cuda()
    y = mhla_b200.gate_add(xd, None if gt is None else gt.cuda(), None if ad is None else ad.cuda())
    tol = 3e-3 if dtype == torch.bfloat16 else 5e-4
    assert oracle.err_ratio(ref, y.float().cpu()) < tol
    y2 = mhla_b200.gate_add(xd, None if gt is None else gt.cuda(), None if ad is None else ad.cuda(), out=xd)   # in place
    assert y2.data_ptr() == xd.data_ptr() and torch.equal(y2, y)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,bias,grid", [(torch.bfloat16, True, (3, 4, 5)), (torch.float16, False, (1, 6, 7)),
                                             (torch.bfloat16, True, (5, 2, 1))])
def test_dwconv3d_tokens_matches_conv3d(dtype, bias, grid):
    """csrc/gated_norm_kernel.cuh dwconv3d_kernel (mhla_dwconv3d) against the reference's LePE path: nn.Conv3d(C, C, 3,
    padding=1, groups=C) on the NCDHW rearrangement of v (mhla_utils.py:289-296), fp32 on CPU as the yardstick."""
    import mhla_b200
    torch.manual_seed(9)
    B, C_ = 2, 64
    F_, H_, W_ = grid
    conv = torch.nn.Conv3d(C_, C_, 3, padding=1, groups=C_, bias=bias)
    x = torch.randn(B, F_ * H_ * W_, C_).to(dtype)
    with torch.no_grad():
        ref = conv(x.float().view(B, F_, H_, W_, C_).permute(0, 4, 1, 2, 3)).permute(0, 2, 3, 4, 1).reshape(B, -1, C_)
        y = mhla_b200.dwconv3d_tokens(x.cuda(), conv.weight.cuda(), None if conv.bias is None else conv.bias.cuda(), grid)
    tol = 4e-3 if dtype == torch.bfloat16 else 6e-4
    assert oracle.err_ratio(ref, y.float().cpu()) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("normalize_out,gated,lepe,post", [(False, False, False, True), (True, True, False, True),
                                                           (True, False, True, True), (False, True, True, True),
                                                           (False, True, True, "epilogue"), (True, True, False, "epilogue"),
                                                           (False, True, True, False)])
def test_wan_module_fused_path_equals_reference_style_path(normalize_out, gated, lepe, post):
    """MHLA_Video_Uni at Wan's block shape (7*6*5 = 210 tokens per block, D = 128): the fused inference path (one
    pre-processing launch + the 3-D block view) against the module's own reference-style path (torch pre-processing +
    block-major rearrange copies), bf16 autocast as in the reference's sampler (inference.py:284)."""
    torch.manual_seed(3)
    dim, heads, layout, grid = 256, 2, (1, 2, 2), (7, 12, 10)
    N = grid[0] * grid[1] * grid[2]
    m = MHLA_Video_Uni(dim, heads, None, 0.0, None, True, layout, normalize_out=normalize_out, is_gated=gated,
                       is_lepe=lepe, fuse_post=post).cuda().eval()   # post-ops: streaming launch / readout epilogue / torch
    m.block_attn.conv.weight.data.mul_(1.0 + 0.1 * torch.rand_like(m.block_attn.conv.weight))
    x = torch.randn(2, N, dim, device="cuda")
    gs = torch.tensor([list(grid)] * 2, dtype=torch.long)
    freqs = oracle.rope_freqs_wan(dim // heads)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y_fast = m(x, torch.tensor([N, N]), gs, freqs)
        m.fast_path = False
        y_slow = m(x, torch.tensor([N, N]), gs, freqs)
    assert oracle.err_ratio(y_slow.float().cpu(), y_fast.float().cpu()) < 1.5e-2


@pytest.mark.gpu
def test_wan_module_training_step_through_the_3d_block_view():
    """Training in the shipped Wan configuration (norm_output: false): forward and the three gradient launches of
    autograd.BlockmixGridFunction gather / scatter the blocks by TMA.  Gradients of the projections and of the mixing
    matrix against the module's block-major path (rearrange copies + BlockmixFunction) on the same weights."""
    torch.manual_seed(4)
    dim, heads, layout, grid = 256, 2, (1, 2, 2), (7, 12, 10)
    N = grid[0] * grid[1] * grid[2]
    m = MHLA_Video_Uni(dim, heads, None, 0.0, None, True, layout, normalize_out=False).cuda().train()
    x = torch.randn(2, N, dim, device="cuda")
    gs = torch.tensor([list(grid)] * 2, dtype=torch.long)
    freqs = oracle.rope_freqs_wan(dim // heads)
    m.norm_q.weight.data.add_(0.2 * torch.rand_like(m.norm_q.weight))
    m.norm_k.weight.data.add_(0.2 * torch.rand_like(m.norm_k.weight))
    grads = []
    # (fast, fused pre-processing): 3-D view + autograd.WanPrepFunction (default) | 3-D view + torch pre-processing |
    # block-major copies + torch pre-processing (reference-style)
    for fast, prep in ((True, True), (True, False), (False, False)):
        m.fast_path, m.train_fused_prep = fast, prep
        m.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x, torch.tensor([N, N]), gs, freqs)
        y.float().square().mean().backward()
        grads.append([p.grad.detach().float().cpu().clone() for p in (m.q.weight, m.k.weight, m.v.weight,
                                                                       m.block_attn.conv.weight, m.norm_q.weight,
                                                                       m.norm_k.weight)])
    for g_ in grads[:2]:
        for a, b in zip(g_, grads[2]):
            assert float(b.abs().max()) > 0
            assert oracle.err_ratio(b, a) < 2e-2
