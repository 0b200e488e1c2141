"""Variant-A modules: ``MHLA4DiT`` (mhla_dit/mhla/mhla.py:141-275) and its image-classification twin
``MHLA_Normed_Torch`` (mhla_image_classification/models/modules/attention/mhla.py:141-289).

Same constructor arguments, attributes (``piece_attn.conv.weight`` is what the trainers clamp,
mhla_dit/train.py:308-310) and ``state_dict`` keys as the reference, so checkpoints load unchanged.  ``forward``
keeps the reference's pre/post-processing in PyTorch and replaces lines :262-268 by one ``mhla_b200.mhla`` call that
consumes q/k/v as strided [B, H, M, w, D] views (no head-split copy).  Additionally accepts the 3-D ``[B, M*w, C]``
input ``DiT_MHLA`` actually feeds (the reference raises there).
"""
from __future__ import annotations

import torch
from einops import rearrange
from torch import nn

from ..mixing import BlockDistanceConv
from ..ops import gate_add, mhla_blockmix


class _MHLAImage(nn.Module):
    _block_kw = "block_size"
    _lepe_kernel = 3

    def __init__(self, dim, heads=8, dim_head=None, dropout=0.1, fixed_weight_value=None, qk_norm=False,
                 transform="linear", **kwargs):
        super().__init__()
        if dim_head is None:
            dim_head = dim // heads
        inner_dim = dim_head * heads
        self.num_heads = heads
        self.head_dim = dim_head
        self.scale = dim_head ** -0.5

        self.norm = nn.LayerNorm(dim)
        is_bias = kwargs["qkv_bias"] if "qkv_bias" in kwargs else False
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=is_bias)
        self.q_norm = nn.RMSNorm(dim) if qk_norm else nn.Identity()
        self.k_norm = nn.RMSNorm(dim) if qk_norm else nn.Identity()
        kk = self._lepe_kernel
        self.lepe = nn.Conv2d(dim, dim, kk, 1, kk // 2, groups=dim)

        bs = kwargs[self._block_kw] if self._block_kw in kwargs else 49
        setattr(self, self._block_kw, bs)
        setattr(self, "block_len" if self._block_kw == "block_size" else "window_len", int(bs ** 0.5))
        self._bs, self._bl = bs, int(bs ** 0.5)
        self.embed_len = kwargs["embed_len"] if "embed_len" in kwargs else 196
        self.num_pieces = self.embed_len // bs
        self.pieces_len = int(self.num_pieces ** 0.5)
        self.piece_attn = BlockDistanceConv(
            num_patches_per_side=int(self.embed_len ** 0.5), patch_group_size=bs, transform=transform,
            local_thres=kwargs.get("local_thres", 1.5), exp_sigma=kwargs.get("exp_sigma", 3))
        self.eps = kwargs.get("eps", 1e-6)
        self.fuse_lepe = kwargs.get("fuse_lepe", True)   # extension: "+ lepe" inside the kernel's readout epilogue
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))
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

    def _mlp_lepe(self, x):
        q, k, v = self.to_qkv(x).chunk(3, dim=-1)
        pl, bl = self.pieces_len, self._bl
        lepe = self.lepe(rearrange(v, "b (h w) (p1 p2) d -> b d (h p1) (w p2)", h=pl, w=pl, p1=bl, p2=bl))
        lepe = rearrange(lepe, "b d (h p1) (w p2) -> b (h w) (p1 p2) d", h=pl, w=pl, p1=bl, p2=bl)
        return q, k, v, lepe

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        squeeze = x.dim() == 3
        if squeeze:                                   # [B, M*w, C] block-major tokens (DiT_MHLA.forward)
            x = x.view(x.shape[0], self.num_pieces, self._bs, x.shape[-1])
        x = self.norm(x)
        B, M, w, C = x.shape
        H, D = self.num_heads, self.head_dim
        q, k, v, lepe = self._mlp_lepe(x)
        q = torch.relu(self.q_norm(q)) + self.eps      # mhla.py:226-230
        k = torch.relu(self.k_norm(k)) + self.eps
        # "b n w (h d) -> (b h) n w d" as zero-copy views [B, H, M, w, D]
        q5, k5, v5 = (t.reshape(B, M, w, H, D).permute(0, 3, 1, 2, 4) for t in (q, k, v))
        W = self.piece_attn.conv.weight
        training = torch.is_grad_enabled() and (q.requires_grad or v.requires_grad or W.requires_grad)
        if D in (64, 128) and not training:
            # inference: the kernel writes straight into the "(b h) n w d -> b n w (h d)" layout (no head-merge copy)
            cdtype = q.dtype if q.dtype in (torch.bfloat16, torch.float16) else torch.bfloat16
            obuf = torch.empty((B, M, w, H, D), dtype=cdtype, device=x.device)
            # "+ lepe" (mhla.py:271-273): short sequences (the whole unit on chip, csrc/smalln_kernel.cuh) add it in the
            # readout epilogue, in fp32 before the single rounding; above that size the general kernel's epilogue has no
            # latency budget for a second input stream and ONE streaming launch behind the operator adds it
            small = D == 64 and M <= 64 and M * w <= 256
            add = lepe.reshape(B, M, w, H, D).permute(0, 3, 1, 2, 4) if (self.fuse_lepe and small) else None
            mhla_blockmix(q5, k5, v5, W, eps=self.eps, normalize=True, out=obuf.permute(0, 3, 1, 2, 4), out_add=add)   # mhla.py:262-268
            out = obuf.view(B, M, w, H * D)
            if add is None and self.fuse_lepe and lepe.dtype == cdtype and (H * D) % 8 == 0:
                out = gate_add(out, None, lepe, out=out).to(x.dtype)
            elif add is None:
                out = out.to(x.dtype) + lepe
            else:
                out = out.to(x.dtype)
            out = self.to_out(out)
            return out.view(B, M * w, -1) if squeeze else out
        else:
            # training (autograd through the operator) or a head dim the shim zero-pads (DiT-XL: 1152 / 16 = 72)
            o5 = mhla_blockmix(q5, k5, v5, W, eps=self.eps, normalize=True)
            out = o5.permute(0, 2, 3, 1, 4).reshape(B, M, w, H * D)
        out = out.to(x.dtype) + lepe                                              # "(b h) n w d -> b n w (h d)"
        out = self.to_out(out)
        return out.view(B, M * w, -1) if squeeze else out


class MHLA4DiT(_MHLAImage):
    """mhla_dit/mhla/mhla.py:141-275 (kwargs: qkv_bias, block_size=49, embed_len=196, local_thres, exp_sigma, eps)."""
    _block_kw = "block_size"
    _lepe_kernel = 3

    def __init__(self, dim, heads=8, dim_head=None, dropout=0.1, fixed_weight_value=None, qk_norm=False,
                 transform="linear", **kwargs):
        super().__init__(dim, heads, dim_head, dropout, fixed_weight_value, qk_norm, transform, **kwargs)


class MHLA_Normed_Torch(_MHLAImage):
    """mhla_image_classification/models/modules/attention/mhla.py:141-289 (window_size kwarg, 5x5 LePE, "cos")."""
    _block_kw = "window_size"
    _lepe_kernel = 5

    def __init__(self, dim, heads=8, dim_head=None, dropout=0.1, fixed_weight_value=None, qk_norm=False,
                 transform="cos", **kwargs):
        super().__init__(dim, heads, dim_head, dropout, fixed_weight_value, qk_norm, transform, **kwargs)
