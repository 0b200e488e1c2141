"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE'S OWN CODE.

Run in the build container only (the reference checkout is mounted read-only at /root/reference and does
not exist on the GPU box):

    TORCHDYNAMO_DISABLE=1 python tests/golden/make_golden.py

Workarounds: ``torch.compile`` decorators are disabled (Inductor's CPU build is broken in
the image); ``mhla_utils.py`` is loaded by path with stub ``diffusion.model.wan.model`` modules exposing the
reference's ``WanRMSNorm``; ``naive.py`` is loaded by file path because the in-repo ``fla`` package cannot be
imported.  The inline operator cores (variants A/B) are re-executed here statement by statement around the
reference's own ``BlockDistanceConv(3D)`` instances and asserted equal to the full reference module forward
before anything is written, so every stored tensor is a value the reference itself produced.
"""
import importlib.util
import os
import sys
import types

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

import numpy as np
import torch
from einops import rearrange

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_grad_enabled(False)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


ONLY = set(sys.argv[1:])      # optional fixture names: regenerate just those


def save(name, **arrays):
    if ONLY and name not in ONLY:
        return
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


# ---------------------------------------------------------------------------------------------------
# variant A : mhla_dit/mhla/mhla.py (MHLA4DiT) and its image-classification twin
# ---------------------------------------------------------------------------------------------------
def golden_variant_a():
    sys.path.insert(0, os.path.join(REF, "mhla_dit"))
    from mhla.mhla import MHLA4DiT, BlockDistanceConv  # reference code
    sys.path.pop(0)

    # mixing matrices for several layouts / transforms
    ws = {}
    for side, group, tr in [(16, 16, "linear"), (16, 16, "cos"), (8, 4, "exp"), (8, 16, "gaussian"), (8, 4, "local"),
                            (14, 49, "linear")]:
        bd = BlockDistanceConv(num_patches_per_side=side, patch_group_size=group, transform=tr)
        ws[f"W_side{side}_group{group}_{tr}"] = bd.get_weight_matrix()
    save("blockdist2d", **ws)

    for tag, dim, heads, block_size, embed_len, qk_norm, B in [
        ("a_dit_s2", 128, 2, 16, 256, False, 2),       # DiT-like: M=16, w=16, D=64
        ("a_qknorm", 96, 3, 4, 64, True, 1),           # M=16, w=4, D=32, RMSNorm on q,k
    ]:
        torch.manual_seed(0)
        m = MHLA4DiT(dim, heads=heads, dropout=0.0, qk_norm=qk_norm, block_size=block_size, embed_len=embed_len,
                     qkv_bias=True).eval()
        # perturb W so that it is not the pristine init (trainable parameter)
        m.piece_attn.conv.weight.data.mul_(1.0 + 0.1 * torch.rand_like(m.piece_attn.conv.weight))
        M, w = embed_len // block_size, block_size
        x = torch.randn(B, M, w, dim)
        y = m(x)                                                         # reference module forward
        # re-run the reference's forward statement by statement (mhla.py:252-275) to expose the core tensors
        xn = m.norm(x)
        q, k, v, lepe = m._mlp_lepe(xn)
        q, kt, v = m._process_qkv_impl(q, k, v, B, M, heads, m.head_dim)   # kt is k transposed [(BH),M,D,w]
        kv = torch.matmul(kt, v)
        kv_mixed = m.piece_attn(kv)
        k_sum = kt.sum(dim=-1, keepdim=True)
        normalizer = m.piece_attn(torch.matmul(q, k_sum)) + m.eps
        out = torch.matmul(q, kv_mixed) / normalizer
        y2 = m.to_out(rearrange(out, "(b h) n w d -> b n w (h d)", b=B, h=heads) + lepe)
        assert torch.equal(y, y2), "restated forward must be bit-identical to the reference module forward"
        sd = {f"sd.{k_}": v_ for k_, v_ in m.state_dict().items()}
        big = {} if m.head_dim > 32 else dict(kv=kv, kv_mixed=kv_mixed)   # keep the fixtures small
        save(tag, x=x, y=y, q=q, k=kt.transpose(-2, -1).contiguous(), v=v, W=m.piece_attn.get_weight_matrix(),
             normalizer=normalizer, out=out, eps=np.float32(m.eps), **big,
             heads=np.int32(heads), block_size=np.int32(block_size), embed_len=np.int32(embed_len),
             qk_norm=np.int32(qk_norm), **sd)

    # image-classification twin: same operator, different defaults (5x5 LePE, transform="cos", window_size kwarg)
    sys.path.insert(0, os.path.join(REF, "mhla_image_classification", "models", "modules", "attention"))
    twin = _load(os.path.join(REF, "mhla_image_classification/models/modules/attention/mhla.py"), "ref_vit_mhla")
    sys.path.pop(0)
    torch.manual_seed(1)
    m = twin.MHLA_Normed_Torch(64, heads=2, dropout=0.0, qk_norm=True, window_size=16, embed_len=64).eval()
    x = torch.randn(2, 4, 16, 64)
    y = m(x)
    sd = {f"sd.{k_}": v_ for k_, v_ in m.state_dict().items()}
    save("a_vit_twin", x=x, y=y, heads=np.int32(2), window_size=np.int32(16), embed_len=np.int32(64), **sd)


# ---------------------------------------------------------------------------------------------------
# variant B : mhla_videogen/diffusion/model/wan/mhla_utils.py (MHLA_Video_Uni)
# ---------------------------------------------------------------------------------------------------
def golden_variant_b_stubs():
    class WanRMSNorm(torch.nn.Module):   # verbatim behaviour of wan/model.py:181-196 (stub for the heavy import)
        def __init__(self, dim, eps=1e-5):
            super().__init__()
            self.dim, self.eps = dim, eps
            self.weight = torch.nn.Parameter(torch.ones(dim))

        def forward(self, x):
            return self._norm(x.float()).type_as(x) * self.weight

        def _norm(self, x):
            return x * torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + self.eps)

    for name in ["diffusion", "diffusion.model", "diffusion.model.wan", "diffusion.model.wan.model"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["diffusion.model.wan.model"].WanRMSNorm = WanRMSNorm


def golden_variant_b():
    golden_variant_b_stubs()
    mu = _load(os.path.join(REF, "mhla_videogen/diffusion/model/wan/mhla_utils.py"), "ref_mhla_utils")

    ws = {}
    for layout, tr in [((3, 5, 10), "linear"), ((2, 2, 3), "cos"), ((4, 1, 1), "linear"), ((2, 3, 2), "exp"),
                       ((2, 2, 2), "local"), ((2, 2, 2), "gaussian")]:
        bd = mu.BlockDistanceConv3D(blocks_layout=layout, transform=tr)
        ws["W_" + "x".join(map(str, layout)) + "_" + tr] = bd.get_weight_matrix()
    save("blockdist3d", **ws)

    def rope_params(max_seq_len, dim, theta=10000):          # wan/model.py:139-146
        freqs = torch.outer(torch.arange(max_seq_len),
                            1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
        return torch.polar(torch.ones_like(freqs), freqs)

    for tag, dim, heads, layout, grid, normalize_out, gated in [
        ("b_norm", 64, 2, (2, 2, 3), (4, 4, 6), True, False),     # D=32, M=12, w=2*2*2=8
        ("b_nonorm", 128, 2, (1, 2, 2), (2, 4, 8), False, True),  # D=64, M=4, w=2*2*4=16 (shipped: norm_output false)
    ]:
        torch.manual_seed(2)
        m = mu.MHLA_Video_Uni(dim, heads, None, 0.0, None, True, layout, normalize_out=normalize_out,
                              is_gated=gated).eval()
        for p in (m.norm_q.weight, m.norm_k.weight, m.g_norm.weight):
            p.data.add_(0.1 * torch.randn_like(p))
        m.block_attn.conv.weight.data.mul_(1.0 + 0.1 * torch.rand_like(m.block_attn.conv.weight))
        d = dim // heads
        freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                           rope_params(1024, 2 * (d // 6))], dim=1)                     # wan/model.py:1933-1936
        B = 2
        N = grid[0] * grid[1] * grid[2]
        x = torch.randn(B, N, dim)
        grid_sizes = torch.tensor([list(grid)] * B, dtype=torch.long)
        seq_lens = torch.tensor([N] * B)
        y = m(x, seq_lens, grid_sizes, freqs)                                          # reference forward
        # statement-by-statement re-execution of mhla_utils.py:292-366 to expose the operator's tensors
        F_, H_, W_ = grid
        bs = (F_ // layout[0], H_ // layout[1], W_ // layout[2])
        q, k, v, lepe = m._qkv_fn(x, F_, H_, W_)
        q, k, v = m._process_qkv_impl(q.float(), k.float(), v.float(), B, N, heads, d)
        q, k, v = (rearrange(t, "b n (h d) -> b n h d", h=heads) for t in (q, k, v))
        q_rope, k_rope = mu.rope_apply(q, grid_sizes, freqs), mu.rope_apply(k, grid_sizes, freqs)
        pat = "b (fb p1 hb p2 wb p3) h c -> (b h) (fb hb wb) (p1 p2 p3) c"
        kw = dict(fb=layout[0], hb=layout[1], wb=layout[2], p1=bs[0], p2=bs[1], p3=bs[2])
        qb, kb, vb, qrb, krb = (rearrange(t, pat, **kw).contiguous() for t in (q, k, v, q_rope, k_rope))
        kv = m.block_attn(torch.matmul(krb.transpose(-2, -1), vb))
        if normalize_out:
            k_sum = kb.transpose(-2, -1).sum(dim=-1, keepdim=True)
            normalizer = m.block_attn(torch.matmul(qb, k_sum)) + m.eps
            out = torch.matmul(qrb, kv) / normalizer
        else:
            out = torch.matmul(qrb, kv)
        o = rearrange(out, "(b h) n w d -> b n w (h d)", b=B, h=heads)
        o = rearrange(o, "b (fb hb wb) (p1 p2 p3) c -> b (fb p1 hb p2 wb p3) c", **kw)
        o = rearrange(m.g_norm(rearrange(o, "b n (h d) -> b n h d", h=heads)), "b n h d -> b n (h d)")
        if gated:
            o = o * m.g_fn(m.g(x))
        y2 = m.o(o)
        assert torch.equal(y, y2), "restated forward must be bit-identical to the reference module forward"
        sd = {f"sd.{k_}": v_ for k_, v_ in m.state_dict().items()}
        save(tag, x=x, y=y, q=qb, k=kb, v=vb, q_rope=qrb, k_rope=krb, W=m.block_attn.get_weight_matrix(), out=out,
             q_tok=q, q_rope_tok=q_rope, eps=np.float32(m.eps), heads=np.int32(heads),
             layout=np.array(layout), grid=np.array(grid), normalize_out=np.int32(normalize_out),
             gated=np.int32(gated), **sd)


# ---------------------------------------------------------------------------------------------------
# variant B' : the five further MHLA self-attention classes of wan/model.py:428-1390 (registry :1592-1605)
# ---------------------------------------------------------------------------------------------------
def golden_variant_b_prime():
    """wan/model.py cannot be imported here (diffusers / mmcv / ... are not installed), so the five class definitions
    are cut out of the reference file by AST and executed unmodified in a namespace that provides exactly the names
    they use: torch / nn / rearrange, the reference's own BlockDistanceConv3D and rope_apply (mhla_utils.py, loaded as
    in golden_variant_b) and WanRMSNorm (also cut from model.py:181-196)."""
    import ast
    golden_variant_b_stubs()
    mu = _load(os.path.join(REF, "mhla_videogen/diffusion/model/wan/mhla_utils.py"), "ref_mhla_utils_bp")
    path = os.path.join(REF, "mhla_videogen/diffusion/model/wan/model.py")
    src = open(path).read()
    tree = ast.parse(src)
    names = ["WanRMSNorm", "Gated_MHLA_Video", "MHLA_Video_Nope", "Gated_MHLA_Video_LePE", "MHLA_Video_LePE", "MHLA_Video"]
    ns = {"torch": torch, "nn": torch.nn, "rearrange": rearrange, "BlockDistanceConv3D": mu.BlockDistanceConv3D,
          "rope_apply": mu.rope_apply, "__name__": "ref_wan_model_cut"}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)

    def rope_params(max_seq_len, dim, theta=10000):          # wan/model.py:139-146
        freqs = torch.outer(torch.arange(max_seq_len),
                            1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
        return torch.polar(torch.ones_like(freqs), freqs)

    dim, heads, layout, grid = 128, 2, (1, 2, 2), (2, 4, 8)
    d = dim // heads
    freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                       rope_params(1024, 2 * (d // 6))], dim=1)
    for key, cls, kw in [("gated_mhla", "Gated_MHLA_Video", dict(normalize_out=True)),
                         ("mhla_nope", "MHLA_Video_Nope", dict(normalize_out=False, out_rmsnorm=True)),
                         ("gated_mhla_lepe", "Gated_MHLA_Video_LePE", dict(normalize_out=False)),
                         ("mhla_lepe", "MHLA_Video_LePE", dict(normalize_out=True, out_rmsnorm=False)),
                         ("mhla", "MHLA_Video", dict(normalize_out=False, out_rmsnorm=True))]:
        torch.manual_seed(5)
        # the positional call of WanAttentionBlock (wan/model.py:1644-1646): window_size lands in dim_head, qk_norm in
        # dropout, eps in fixed_weight_value (-> every weight starts at 1e-6; re-randomised below like a checkpoint load)
        m = ns[cls](dim, heads, (-1, -1), True, 1e-6, rope_after=False, without_rope=False, power=1.0,
                    out_rmsnorm=kw.get("out_rmsnorm", False), normalize_out=kw["normalize_out"], is_gated=False,
                    is_lepe=False, block_layout=layout).eval()
        for n_, p in m.named_parameters():
            if n_ == "block_attn.conv.weight":
                p.data = mu.BlockDistanceConv3D(blocks_layout=layout).conv.weight.data * (1.0 + 0.1 * torch.rand_like(p))
            elif n_.endswith("weight") and p.dim() == 1:
                p.data = 1.0 + 0.1 * torch.randn_like(p)
            elif p.dim() >= 2:
                p.data = torch.randn_like(p) * (p.shape[1] ** -0.5 if p.dim() == 2 else 0.2)
            else:
                p.data = 0.1 * torch.randn_like(p)
        B = 2
        N = grid[0] * grid[1] * grid[2]
        x = torch.randn(B, N, dim)
        y = m(x, torch.tensor([N] * B), torch.tensor([list(grid)] * B, dtype=torch.long), freqs)
        sd = {f"sd.{k_}": v_ for k_, v_ in m.state_dict().items()}
        save("bp_" + key, x=x, y=y, heads=np.int32(heads), layout=np.array(layout), grid=np.array(grid),
             normalize_out=np.int32(kw["normalize_out"]), out_rmsnorm=np.int32(kw.get("out_rmsnorm", False)), **sd)


# ---------------------------------------------------------------------------------------------------
# variant C : mhla_nlp/fla/ops/mhla/naive.py
# ---------------------------------------------------------------------------------------------------
def golden_variant_c():
    nv = _load(os.path.join(REF, "mhla_nlp/fla/ops/mhla/naive.py"), "ref_naive")
    L = 32
    init = (torch.tril(torch.ones(L, L)) / (torch.arange(L, dtype=torch.float32).unsqueeze(1) + 1.0)).view(L, L, 1, 1, 1, 1)
    for tag, B, T, H, K, V, signed, rand_mm in [
        ("c_t256", 1, 256, 2, 32, 64, True, True),
        ("c_t200_ragged", 2, 200, 2, 64, 64, False, False),     # T % 64 != 0 -> zero padding path
        ("c_cfg1", 1, 1024, 4, 64, 64, True, True),             # BASELINE cfg1 (CPU plumbing case)
        ("c_kv_128_256", 1, 192, 1, 128, 256, True, False),     # NLP 340M head shape
    ]:
        g = torch.Generator().manual_seed(3)
        q = torch.randn(B, T, H, K, generator=g)
        k = torch.randn(B, T, H, K, generator=g)
        if not signed:
            q, k = torch.relu(q), torch.relu(k)
        v = torch.randn(B, T, H, V, generator=g)
        mm = torch.clamp(torch.rand(L, L, generator=g), 1e-5, 1).tril().view(L, L, 1, 1, 1, 1) if rand_mm else init
        o = nv.naive_chunk_simple_mhla_fixed(q, k, v, mm)
        if tag == "c_cfg1":   # 1 MiB per tensor: store the output only; the test regenerates q,k,v,mm from seed 3
            save(tag, o=o, seed=np.int32(3), shape=np.array([B, T, H, K, V]), q_sum=q.double().sum(), v_sum=v.double().sum())
        else:
            save(tag, q=q, k=k, v=v, mm=mm.view(L, L), o=o)
    # token-recurrent form, T <= 64 (the only regime the layer uses it in, layers/mhla.py:247)
    g = torch.Generator().manual_seed(4)
    q, k, v = torch.randn(2, 48, 2, 32, generator=g), torch.randn(2, 48, 2, 32, generator=g), torch.randn(2, 48, 2, 64, generator=g)
    o, S = nv.naive_recurrent_mhla(q, k, v, init)
    o_chunk = nv.naive_chunk_simple_mhla_fixed(q, k, v, init)
    save("c_recurrent_t48", q=q, k=k, v=v, mm=init.view(L, L), o=o, o_chunk=o_chunk, S=S)
    # the same with a head shape inside the kernel envelope (K=64, V=128), a ragged T < 64 and a random mixing matrix
    g = torch.Generator().manual_seed(5)
    q, k, v = torch.randn(2, 57, 2, 64, generator=g), torch.randn(2, 57, 2, 64, generator=g), torch.randn(2, 57, 2, 128, generator=g)
    mm = torch.clamp(torch.rand(L, L, generator=g), 1e-5, 1).tril().view(L, L, 1, 1, 1, 1)
    o, S = nv.naive_recurrent_mhla(q, k, v, mm)
    o_chunk = nv.naive_chunk_simple_mhla_fixed(q, k, v, mm)
    save("c_recurrent_k64", q=q, k=k, v=v, mm=mm.view(L, L), o=o, o_chunk=o_chunk)


if __name__ == "__main__":
    golden_variant_a()
    golden_variant_b()
    golden_variant_b_prime()
    golden_variant_c()
