"""Variant-B modules: ``MHLA_Video_Uni`` (mhla_videogen/diffusion/model/wan/mhla_utils.py:158-366) and the five further
MHLA self-attention classes of wan/model.py:428-1390 (``Gated_MHLA_Video``, ``MHLA_Video_Nope``,
``Gated_MHLA_Video_LePE``, ``MHLA_Video_LePE``, ``MHLA_Video``) - one operator core, different post-processing; the
registry ``WAN_SELFATTENTION_CLASSES`` below carries the reference's keys (wan/model.py:1592-1605).

Same constructor signature (including the odd positional call ``cls(dim, num_heads, window_size, qk_norm, eps, ...)``
of wan/model.py:1644-1646, whose third positional lands in ``dim_head`` and is ignored), attributes and
``state_dict`` keys (q,k,v,o,[g],norm_q,norm_k,g_norm,block_attn.conv.weight,[lepe]).  The operator core
(:328-341) is the CUDA kernel; RoPE is evaluated in fp32 real arithmetic on the GPU instead of the reference's
complex128 per-sample Python loop (:127-156).
"""
from __future__ import annotations

import torch
from einops import rearrange
from torch import nn

from ..mixing import BlockDistanceConv3D
from ..ops import dwconv3d_tokens, gate_add, mhla_blockmix, mhla_blockmix_grid, wan_prep


class WanRMSNorm(nn.Module):
    """wan/model.py:181-196."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.dim = dim
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return self._norm(x.float()).type_as(x) * self.weight

    def _norm(self, x):
        return x * torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + self.eps)


_ROPE_CACHE = {}


def _rope_tables(grid, freqs: torch.Tensor, device):
    """cos/sin [N, D/2] for a (F, H, W) token grid from the reference's complex table [1024, D/2] (wan/model.py:1933-1936)."""
    f, h, w = grid
    key = (f, h, w, freqs.data_ptr(), freqs.shape[1], str(device))
    hit = _ROPE_CACHE.get(key)
    if hit is not None:
        return hit
    c = freqs.shape[1]
    fr = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    fi = torch.cat([
        fr[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1),
        fr[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
        fr[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(f * h * w, -1)
    cos = fi.real.to(device=device, dtype=torch.float32).contiguous()
    sin = fi.imag.to(device=device, dtype=torch.float32).contiguous()
    if len(_ROPE_CACHE) > 8:
        _ROPE_CACHE.clear()
    _ROPE_CACHE[key] = (cos, sin)
    return cos, sin


def rope_apply(x: torch.Tensor, grid_sizes, freqs: torch.Tensor) -> torch.Tensor:
    """x [B, N, H, D] -> fp32 roped tensor; interleaved-pair rotation (mhla_utils.py:127-156).  Every sample uses the
    grid of sample 0, as the module does (:297)."""
    g0 = grid_sizes[0]
    grid = tuple(int(v) for v in (g0.tolist() if torch.is_tensor(g0) else g0))
    B, N, H, D = x.shape
    seq = grid[0] * grid[1] * grid[2]
    cos, sin = _rope_tables(grid, freqs, x.device)
    xf = x.float()
    xr = xf[:, :seq].reshape(B, seq, H, D // 2, 2)
    a, b = xr[..., 0], xr[..., 1]
    c, s = cos[None, :, None, :], sin[None, :, None, :]
    out = torch.stack((a * c - b * s, a * s + b * c), dim=-1).reshape(B, seq, H, D)
    if seq < N:
        out = torch.cat([out, xf[:, seq:]], dim=1)
    return out


class _MHLAVideoBase(nn.Module):
    """Shared implementation.  Class attributes select what the reference classes differ in (wan/model.py:428-1390 and
    mhla_utils.py:158-366 were diffed statement by statement: only the post-processing differs):
      _gate       'always' | 'kwarg' | 'never'   SiLU(g(x)) gate
      _gnorm      'dim' (RMSNorm over the whole channel dim) | 'head' (per head) | None
      _lepe       'always' | 'kwarg' | 'never'   depthwise 3x3x3 Conv3d on v
      _out_norm   True: ``out_rmsnorm`` (RMSNorm over dim after ``o``) when the kwarg is set
    """
    _gate, _gnorm, _lepe, _out_norm = "kwarg", "head", "kwarg", False

    def __init__(self, dim, num_heads=8, dim_head=None, dropout=0.1, fixed_weight_value=None, qk_norm=True,
                 block_layout=(3, 5, 10), transform="linear", qkv_bias=False, eps=1e-6, is_gated=False, is_lepe=False,
                 **kwargs):
        super().__init__()
        # (WanAttentionBlock calls cls(dim, num_heads, window_size, qk_norm, eps, ...): window_size lands in dim_head -
        #  ignored, as in the reference - and eps in fixed_weight_value, wan/model.py:1644-1646)
        dim_head = dim // num_heads
        self.dim = dim
        self.num_heads = num_heads
        self.head_dim = dim_head
        gated = self._gate == "always" or (self._gate == "kwarg" and is_gated)
        lepe = self._lepe == "always" or (self._lepe == "kwarg" and is_lepe)
        self.q = nn.Linear(dim, dim)
        self.k = nn.Linear(dim, dim)
        self.v = nn.Linear(dim, dim)
        if self._gate == "always":
            self.g, self.g_fn = nn.Linear(dim, dim), nn.SiLU()
        elif self._gate == "kwarg":
            self.g, self.g_fn = (nn.Linear(dim, dim), nn.SiLU()) if gated else (None, None)
        if self._gnorm is not None:
            self.g_norm = WanRMSNorm(dim if self._gnorm == "dim" else dim_head, eps=eps)
        self.is_gated = gated
        self.fuse_out_norm = kwargs.get("fuse_out_norm", True)   # extension: per-head g_norm inside the kernel epilogue
        self.fast_path = kwargs.get("fast_path", True)           # extension: fused pre-processing + 3-D block TMA view
        # extension: SiLU gate and "+ lepe" behind the operator.  True / "stream": ONE streaming launch (ops.gate_add) instead
        # of three eager elementwise passes; "epilogue": inside the operator's readout epilogue (ABI v4 out_gate / out_add;
        # measured slower than the streaming pass on a B200 - the epilogue's row-wise global loads queue behind the TMA
        # stream, profiles/r02c_notes.md); False: plain torch ops
        self.fuse_post = kwargs.get("fuse_post", True)
        # extension: training goes through the fused pre-processing launch with an analytic backward (autograd.WanPrepFunction)
        self.train_fused_prep = kwargs.get("train_fused_prep", True)
        self.is_lepe = lepe
        self.norm_q = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()
        self.norm_k = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()
        self.out_norm = kwargs.get("out_rmsnorm", False)
        if self._out_norm:
            self.out_rmsnorm = WanRMSNorm(dim, eps=eps) if self.out_norm else nn.Identity()
        self.normalize_out = kwargs.get("normalize_out", True)
        self.blocks_layout = tuple(block_layout)
        self.num_blocks = self.blocks_layout[0] * self.blocks_layout[1] * self.blocks_layout[2]
        self.block_attn = BlockDistanceConv3D(blocks_layout=self.blocks_layout, transform=transform)
        if self._lepe == "always":
            self.lepe = nn.Conv3d(dim, dim, kernel_size=(3, 3, 3), stride=1, padding=(1, 1, 1), groups=dim)
        elif self._lepe == "kwarg":
            self.lepe = nn.Conv3d(dim, dim, kernel_size=(3, 3, 3), stride=1, padding=(1, 1, 1), groups=dim) if lepe else None
        self.eps = eps
        self.o = nn.Linear(dim, dim)
        self.rope_after = kwargs.get("rope_after", False)
        self.power = kwargs.get("power", 1.0)
        self.without_rope = kwargs.get("without_rope", False)
        if fixed_weight_value is not None:
            self._init_weights_with_fixed_value(fixed_weight_value)

    def _init_weights_with_fixed_value(self, value):
        for name, param in self.named_parameters():
            if "weight" in name:
                nn.init.constant_(param, value)
            elif "bias" in name and param is not None:
                nn.init.zeros_(param)

    @staticmethod
    def init_to_value(model, value=1.0):
        for name, param in model.named_parameters():
            if "weight" in name:
                nn.init.constant_(param, value)
            elif "bias" in name and param is not None:
                nn.init.zeros_(param)
        return model

    def _forward_fused(self, x, q, k, v, lepe, grid, grid_sizes, freqs):
        """Inference path without a single layout copy: ONE pre-processing launch (RMSNorm over C, relu + eps, RoPE,
        16-bit token-major outputs: csrc/wan_prep_kernel.cuh) and ONE operator launch that gathers the 3-D blocks by
        TMA and scatters the output back (csrc/blockmix_kernel.cuh, 3-D block view) - instead of .float() copies, a
        complex128 RoPE with a host sync, a 5-tensor cat and two rearranges (mhla_utils.py:303-326, :345-354)."""
        B, N, C = x.shape
        nh, D = self.num_heads, self.head_dim
        cos, sin = _rope_tables(grid, freqs, x.device)
        wq = self.norm_q.weight if isinstance(self.norm_q, WanRMSNorm) else None
        wk = self.norm_k.weight if isinstance(self.norm_k, WanRMSNorm) else None
        eps_n = self.norm_q.eps if isinstance(self.norm_q, WanRMSNorm) else 1e-6
        q_rope, k_rope, q_n, k_n = wan_prep(q, k, wq, wk, cos, sin, D, eps_norm=eps_n, eps=self.eps,
                                            want_plain=self.normalize_out)
        cdtype = q_rope.dtype
        v4 = (v if v.dtype == cdtype else v.to(cdtype)).view(B, N, nh, D)
        fuse = (dict(out_rms_weight=self.g_norm.weight, out_rms_eps=self.g_norm.eps)
                if (self._gnorm == "head" and self.fuse_out_norm) else {})
        # the SiLU gate and "+ lepe" ride in the same epilogue (out = g_norm(o) * silu(g(x)) + lepe in fp32, one rounding:
        # mhla_utils.py:357-364, wan/model.py:1001-1003) whenever nothing un-fused sits between the operator and them
        post = {}
        if self.fuse_post == "epilogue" and ((self._gnorm == "head" and fuse) or self._gnorm is None):
            if self.is_gated:
                post["out_gate"] = self.g(x).view(B, N, nh, D)
            if self.is_lepe:
                post["out_add"] = lepe.view(B, N, nh, D)
        Wm = self.block_attn.conv.weight
        if self.normalize_out:
            out = mhla_blockmix_grid(q_n, k_n, v4, Wm, grid, self.blocks_layout, q_rope=q_rope, k_rope=k_rope, eps=self.eps,
                                     normalize=True, **fuse, **post)
        else:
            out = mhla_blockmix_grid(q_rope, k_rope, v4, Wm, grid, self.blocks_layout, eps=self.eps, normalize=False, **fuse,
                                     **post)
        out = out.to(q.dtype)
        if self._gnorm == "head" and not fuse:
            out = self.g_norm(out)
        out = out.reshape(B, N, C)
        if self._gnorm == "dim":
            out = self.g_norm(out)
        need_gate, need_add = self.is_gated and "out_gate" not in post, self.is_lepe and "out_add" not in post
        if (need_gate or need_add) and self.fuse_post and out.dtype in (torch.bfloat16, torch.float16) and C % 8 == 0:
            gate = self.g(x) if need_gate else None
            if gate is not None and gate.dtype != out.dtype:
                gate = gate.to(out.dtype)
            out = gate_add(out.contiguous(), gate, lepe if need_add else None)     # one launch: out * silu(g(x)) + lepe
        else:
            if need_gate:
                out = out * self.g_fn(self.g(x))
            if need_add:
                out = out + lepe
        out = self.o(out)
        return self.out_rmsnorm(out) if self._out_norm else out

    def forward(self, x: torch.Tensor, seq_lens, grid_sizes, freqs) -> torch.Tensor:
        B, N, C = x.shape
        g0 = grid_sizes[0]
        F_, H_, W_ = (int(v) for v in (g0.tolist() if torch.is_tensor(g0) else g0))
        fb, hb, wb = self.blocks_layout
        p1, p2, p3 = F_ // fb, H_ // hb, W_ // wb
        nh, D = self.num_heads, self.head_dim

        q, k, v = self.q(x), self.k(x), self.v(x)                                   # mhla_utils.py:279-288
        lepe = None
        if self.is_lepe:
            lepe_grad = torch.is_grad_enabled() and (v.requires_grad or self.lepe.weight.requires_grad)
            if (self.fast_path and x.is_cuda and not lepe_grad and v.dtype in (torch.bfloat16, torch.float16) and C % 8 == 0
                    and N == F_ * H_ * W_):
                # the depthwise Conv3d straight on the token-major v (no NCDHW rearrangement, no cuDNN depthwise path)
                lepe = dwconv3d_tokens(v, self.lepe.weight, self.lepe.bias, (F_, H_, W_))
            else:
                lepe = self.lepe(rearrange(v, "b (f h w) c -> b c f h w", f=F_, h=H_, w=W_))
                lepe = rearrange(lepe, "b c f h w -> b (f h w) c")
        dtype = q.dtype
        W = self.block_attn.conv.weight
        training = torch.is_grad_enabled() and (q.requires_grad or v.requires_grad or W.requires_grad)
        view3d = (self.fast_path and x.is_cuda and D in (64, 128) and N == F_ * H_ * W_ and p2 * p3 <= 128
                  and -(-p1 // max(1, min(p1, 128 // (p2 * p3)))) <= 2)   # the kernel's 3-D block view applies
        if view3d and not training:
            return self._forward_fused(x, q, k, v, lepe, (F_, H_, W_), grid_sizes, freqs)
        norms_ok = all(isinstance(n_, (WanRMSNorm, nn.Identity)) for n_ in (self.norm_q, self.norm_k))
        if view3d and training and not self.normalize_out and self.train_fused_prep and norms_ok:
            # training in the shipped configuration (norm_output: false) without a layout copy AND without autograd through
            # the fp32 / complex pre-processing: the forward is the fused pre-processing launch, its backward is analytic
            # (autograd.WanPrepFunction); the operator trains through the 3-D block view (autograd.BlockmixGridFunction)
            from ..autograd import WanPrepFunction
            cos, sin = _rope_tables((F_, H_, W_), freqs, x.device)
            wq = self.norm_q.weight if isinstance(self.norm_q, WanRMSNorm) else None
            wk = self.norm_k.weight if isinstance(self.norm_k, WanRMSNorm) else None
            eps_n = self.norm_q.eps if isinstance(self.norm_q, WanRMSNorm) else 1e-6
            q_rope, k_rope = WanPrepFunction.apply(q, k, wq, wk, cos, sin, D, eps_n, self.eps)
            v4 = (v if v.dtype == q_rope.dtype else v.to(q_rope.dtype)).view(B, N, nh, D)
            out = mhla_blockmix_grid(q_rope, k_rope, v4, W, (F_, H_, W_), self.blocks_layout, eps=self.eps,
                                     normalize=False).to(dtype)
            return self._post(x, out, lepe, B, N, C, fuse_norm=False)
        q = torch.relu(self.norm_q(q.float())) + self.eps                          # :308, 267-276
        k = torch.relu(self.norm_k(k.float())) + self.eps
        q, k, v = (t.view(B, N, nh, D) for t in (q, k, v))
        q_rope, k_rope = rope_apply(q, grid_sizes, freqs), rope_apply(k, grid_sizes, freqs)   # :314

        cdtype = torch.float16 if dtype == torch.float16 else torch.bfloat16
        pat = "b (fb p1 hb p2 wb p3) h d -> b h (fb hb wb) (p1 p2 p3) d"
        kw = dict(fb=fb, hb=hb, wb=wb, p1=p1, p2=p2, p3=p3)
        blk = lambda t: rearrange(t.to(cdtype), pat, **kw).contiguous()           # noqa: E731  (:317-326, one 16-bit copy each)
        # the per-head g_norm (:360-364) is fused into the kernel's readout epilogue (fp32, before the single rounding)
        fuse_norm = self._gnorm == "head" and self.fuse_out_norm and not training and D in (64, 128)
        fuse = dict(out_rms_weight=self.g_norm.weight, out_rms_eps=self.g_norm.eps) if fuse_norm else {}
        if self.normalize_out:
            out = mhla_blockmix(blk(q), blk(k), blk(v), W, q_rope=blk(q_rope), k_rope=blk(k_rope), eps=self.eps,
                                normalize=True, **fuse)
        elif training and view3d:
            # training in the shipped configuration (norm_output: false): forward AND the three gradient launches gather /
            # scatter the blocks by TMA (autograd.BlockmixGridFunction) - no block-major copies in either direction
            out = mhla_blockmix_grid(q_rope.to(cdtype), k_rope.to(cdtype), v.to(cdtype), W, (F_, H_, W_), self.blocks_layout,
                                     eps=self.eps, normalize=False).to(dtype)
        else:  # shipped Wan config (norm_output: false): the un-roped q/k are not needed at all
            out = mhla_blockmix(blk(q_rope), blk(k_rope), blk(v), W, eps=self.eps, normalize=False, **fuse)
        if out.dim() == 5:
            out = rearrange(out, "b h (fb hb wb) (p1 p2 p3) d -> b (fb p1 hb p2 wb p3) h d", **kw).to(dtype)   # :343-356
        return self._post(x, out, lepe, B, N, C, fuse_norm)

    def _post(self, x, out, lepe, B, N, C, fuse_norm):
        """Everything between the operator and the layer's output (mhla_utils.py:357-366), plain torch (differentiable)."""
        if self._gnorm == "head" and not fuse_norm:
            out = self.g_norm(out)                                                  # :360-364 per-head RMSNorm
        out = out.reshape(B, N, C)
        if self._gnorm == "dim":
            out = self.g_norm(out)                                                  # wan/model.py:617 (whole channel dim)
        if self.is_gated:
            out = out * self.g_fn(self.g(x))
        if self.is_lepe:
            out = out + lepe
        out = self.o(out)
        return self.out_rmsnorm(out) if self._out_norm else out


class MHLA_Video_Uni(_MHLAVideoBase):
    """mhla_utils.py:158-366: optional gate / LePE (kwargs), per-head g_norm."""
    _gate, _gnorm, _lepe, _out_norm = "kwarg", "head", "kwarg", False


class Gated_MHLA_Video(_MHLAVideoBase):
    """wan/model.py:428-619: always gated, g_norm over the whole channel dim."""
    _gate, _gnorm, _lepe, _out_norm = "always", "dim", "never", False


class MHLA_Video_Nope(_MHLAVideoBase):
    """wan/model.py:621-806: no gate, optional ``out_rmsnorm`` after ``o``."""
    _gate, _gnorm, _lepe, _out_norm = "never", None, "never", True


class Gated_MHLA_Video_LePE(_MHLAVideoBase):
    """wan/model.py:808-1008: gate, per-head g_norm, LePE added before ``o``."""
    _gate, _gnorm, _lepe, _out_norm = "always", "head", "always", False


class MHLA_Video_LePE(_MHLAVideoBase):
    """wan/model.py:1010-1203: LePE, optional ``out_rmsnorm``."""
    _gate, _gnorm, _lepe, _out_norm = "never", None, "always", True


class MHLA_Video(_MHLAVideoBase):
    """wan/model.py:1205-1390: plain (optional ``out_rmsnorm``)."""
    _gate, _gnorm, _lepe, _out_norm = "never", None, "never", True


# the MHLA entries of wan/model.py:1592-1605; a maintainer updates the reference's dict with this one
WAN_SELFATTENTION_CLASSES = {
    "mhla": MHLA_Video, "gated_mhla": Gated_MHLA_Video, "mhla_nope": MHLA_Video_Nope, "mhla_lepe": MHLA_Video_LePE,
    "gated_mhla_lepe": Gated_MHLA_Video_LePE, "mhla_uni": MHLA_Video_Uni,
}
