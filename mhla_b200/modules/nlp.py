"""Variant-C layer: ``MHLA`` (mhla_nlp/fla/layers/mhla.py:29-365) with the reference's constructor signature,
``forward`` contract ``(o, None, past_key_values)`` and ``state_dict`` keys (q_proj,k_proj,v_proj,g_proj,o_proj,
mixing_matrix, g_norm_swish_gate.weight / g_norm.weight, optional {q,k,v}_conv1d.*).

The token mixer (layers/mhla.py:318-337 -> fla/ops/mhla/naive.py) is the CUDA kernel.  The neighbours the reference
takes from ``fla.modules`` (Triton) are restated here in plain PyTorch with identical parameters so the layer has
no Triton dependency: NeoX-style rotary (fla/modules/rotary.py, interleaved=False), gated RMSNorm
``rmsnorm(o) * w * g * sigmoid(g)`` (fla/modules/fused_norm_gate.py:77-99) and the depthwise causal short conv.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from einops import rearrange, repeat

from ..ops import naive_chunk_simple_mhla_fixed, naive_recurrent_mhla


class RotaryEmbedding(nn.Module):
    """Non-interleaved (rotate-half) rotary embedding over the full head dim, base 10000."""

    def __init__(self, dim: int, base: float = 10000.0):
        super().__init__()
        self.dim, self.base = dim, float(base)
        inv = 1.0 / (self.base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
        self.register_buffer("inv_freq", inv, persistent=False)

    def forward(self, q, k, seqlen_offset=0, max_seqlen=None, cu_seqlens=None):
        if cu_seqlens is not None:
            raise NotImplementedError("variable-length (cu_seqlens) rotary is a next-row item")
        T = q.shape[1]
        off = int(seqlen_offset) if not torch.is_tensor(seqlen_offset) else seqlen_offset
        t = torch.arange(T, device=q.device, dtype=torch.float32)
        if torch.is_tensor(off):
            t = t[None, :] + off.to(t)[:, None]
        else:
            t = (t + off)[None, :]
        freqs = t[..., None] * self.inv_freq.to(q.device)                     # [b|1, T, D/2]
        cos, sin = freqs.cos()[:, :, None, :], freqs.sin()[:, :, None, :]

        def rot(x):
            x1, x2 = x.float().chunk(2, dim=-1)
            return torch.cat((x1 * cos - x2 * sin, x1 * sin + x2 * cos), dim=-1).to(x.dtype)

        return rot(q), rot(k)


class RMSNorm(nn.Module):
    def __init__(self, hidden_size, elementwise_affine=True, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden_size)) if elementwise_affine else None

    def forward(self, x):
        y = x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + self.eps)
        if self.weight is not None:
            y = y * self.weight.float()
        return y.to(x.dtype)


class FusedRMSNormGated(RMSNorm):
    """y = rmsnorm(x) * weight * g * sigmoid(g)  (activation 'swish')."""

    def forward(self, x, g):
        y = x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + self.eps)
        if self.weight is not None:
            y = y * self.weight.float()
        gf = g.float()
        return (y * gf * torch.sigmoid(gf)).to(x.dtype)


class ShortConvolution(nn.Conv1d):
    """Depthwise causal conv over time followed by SiLU (fla/modules/convolution.py); weight [hidden, 1, k]."""

    def __init__(self, hidden_size, kernel_size, bias=False, activation="silu"):
        super().__init__(hidden_size, hidden_size, kernel_size, groups=hidden_size, bias=bias, padding=kernel_size - 1)
        self.hidden_size, self.activation = hidden_size, activation

    def forward(self, x, cache=None, output_final_state=False, cu_seqlens=None):
        if cache is not None or cu_seqlens is not None:
            raise NotImplementedError("short-conv state caching / varlen is a next-row item")
        T = x.shape[1]
        y = super().forward(x.transpose(1, 2))[..., :T].transpose(1, 2)
        if self.activation in ("silu", "swish"):
            y = F.silu(y)
        return y, None


_ACT = {"swish": F.silu, "silu": F.silu, "sigmoid": torch.sigmoid, "relu": F.relu, "gelu": F.gelu}


class MHLA(nn.Module):
    def __init__(self, mode: str = "chunk", hidden_size: int = 1024, expand_k: float = 0.5, expand_v: float = 1.0,
                 num_heads: int = 4, num_kv_heads: Optional[int] = None, feature_map: Optional[str] = None,
                 use_short_conv: bool = False, conv_size: int = 4, conv_bias: bool = False,
                 use_output_gate: bool = True, gate_fn: str = "swish", elementwise_affine: Optional[bool] = True,
                 norm_eps: float = 1e-5, gate_logit_normalizer: int = 16, gate_low_rank_dim: int = 16,
                 clamp_min: Optional[float] = None, fuse_norm: bool = True, layer_idx: int = None):
        super().__init__()
        self.mode = mode
        self.hidden_size = hidden_size
        self.expand_k, self.expand_v = expand_k, expand_v
        self.num_heads = num_heads
        self.num_kv_heads = num_kv_heads if num_kv_heads is not None else num_heads
        self.num_kv_groups = self.num_heads // self.num_kv_heads
        self.key_dim = int(hidden_size * expand_k)
        self.value_dim = int(hidden_size * expand_v)
        self.head_k_dim = self.key_dim // num_heads
        self.head_v_dim = self.value_dim // num_heads

        if feature_map == "relu":
            self.feature_map_q, self.feature_map_k = nn.ReLU(), nn.ReLU()
        elif feature_map == "identity":
            self.feature_map_q, self.feature_map_k = nn.Identity(), nn.Identity()
        elif feature_map == "elu":
            def elu(x):
                return F.elu(x) + 1
            self.feature_map_q = self.feature_map_k = elu
        elif feature_map in ("hedgehog", "t2r", "elementwise_product", "dpfp"):
            from fla.modules import feature_map as fm   # learnable upstream feature maps (layers/mhla.py:113-131)
            cls = {"hedgehog": fm.HedgehogFeatureMap, "t2r": fm.T2RFeatureMap,
                   "elementwise_product": fm.HadamardFeatureMap, "dpfp": fm.DPFPFeatureMap}[feature_map]
            self.feature_map_q, self.feature_map_k = cls(head_dim=self.head_k_dim), cls(head_dim=self.head_k_dim)
        else:
            raise NotImplementedError(f"Not supported feature map `{feature_map}`.")

        self.use_short_conv, self.conv_size, self.conv_bias = use_short_conv, conv_size, conv_bias
        self.use_output_gate = use_output_gate
        self.key_dim_per_group = self.key_dim // self.num_kv_groups
        self.value_dim_per_group = self.value_dim // self.num_kv_groups
        self.clamp_min = clamp_min
        self.layer_idx = layer_idx
        assert mode in ["chunk", "fused_recurrent", "fused_chunk"], f"Not supported mode `{mode}`."
        assert self.key_dim % num_heads == 0, f"key dim must be divisible by num_heads of {num_heads}"
        assert self.value_dim % num_heads == 0, f"value dim must be divisible by num_heads of {num_heads}"

        self.q_proj = nn.Linear(hidden_size, self.key_dim, bias=False)
        self.k_proj = nn.Linear(hidden_size, self.key_dim_per_group, bias=False)
        self.v_proj = nn.Linear(hidden_size, self.value_dim_per_group, bias=False)
        if self.use_output_gate:
            self.g_proj = nn.Linear(hidden_size, self.value_dim, bias=False)
        if use_short_conv:
            self.q_conv1d = ShortConvolution(self.key_dim, conv_size, bias=conv_bias, activation="silu")
            self.k_conv1d = ShortConvolution(self.key_dim_per_group, conv_size, bias=conv_bias, activation="silu")
            self.v_conv1d = ShortConvolution(self.value_dim_per_group, conv_size, bias=conv_bias, activation="silu")

        L = 32                                                         # layers/mhla.py:196-200
        lower_tri = torch.tril(torch.ones(L, L, dtype=torch.float32))
        lower_tri = lower_tri / (torch.arange(L, dtype=torch.float32).unsqueeze(1) + 1.0)
        self.mixing_matrix = nn.Parameter(lower_tri.view(L, L, 1, 1, 1, 1))
        self.o_proj = nn.Linear(self.value_dim, hidden_size, bias=False)

        if gate_fn == "swish" and fuse_norm and use_output_gate:
            self.g_norm_swish_gate = FusedRMSNormGated(self.head_v_dim, elementwise_affine, norm_eps)
            self.fuse_norm_and_gate = True
        else:
            self.fuse_norm_and_gate = False
            self.g_norm = RMSNorm(self.head_v_dim, elementwise_affine, norm_eps)
            self.gate_fn = _ACT[gate_fn]
        self.gate_logit_normalizer = gate_logit_normalizer
        assert self.head_k_dim <= 256, "head_k_dim must be less than or equal to 256"
        self.rotary = RotaryEmbedding(dim=self.head_k_dim)

    def forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                past_key_values=None, use_cache: Optional[bool] = False, output_attentions: Optional[bool] = False,
                **kwargs):
        # clamp the mixing matrix weight into [0, 1] (layers/mhla.py:237)
        self.mixing_matrix.data = torch.clamp(self.mixing_matrix.data, 1e-5, 1).tril()
        if attention_mask is not None:
            assert len(attention_mask.shape) == 2, "Expected attention_mask as a 0-1 matrix with shape [batch_size, seq_len]"
            if not bool(attention_mask.all()):
                raise NotImplementedError("padding masks (varlen unpad) are a next-row item (SURVEY.md 8f rank 4)")
        if kwargs.get("cu_seqlens") is not None:
            raise NotImplementedError("cu_seqlens packing is a next-row item (SURVEY.md 8f rank 4)")
        batch_size, q_len, _ = hidden_states.shape
        mode = "fused_recurrent" if q_len <= 64 else self.mode

        if self.use_short_conv:
            q, _ = self.q_conv1d(self.q_proj(hidden_states))
            k, _ = self.k_conv1d(self.k_proj(hidden_states))
            v, _ = self.v_conv1d(self.v_proj(hidden_states))
        else:
            q, k, v = self.q_proj(hidden_states), self.k_proj(hidden_states), self.v_proj(hidden_states)
        q = rearrange(q, "... (h d) -> ... h d", d=self.head_k_dim)
        if self.num_kv_groups > 1:
            k = repeat(k, "... (h d) -> ... (h g) d", g=self.num_kv_groups, d=self.head_k_dim)
            v = repeat(v, "... (h d) -> ... (h g) d", g=self.num_kv_groups, d=self.head_v_dim)
        else:
            k = rearrange(k, "... (h d) -> ... h d", d=self.head_k_dim)
            v = rearrange(v, "... (h d) -> ... h d", d=self.head_v_dim)
        q, k = self.feature_map_q(q), self.feature_map_k(k)

        seqlen_offset = 0
        if past_key_values is not None:
            seqlen_offset = past_key_values.get_seq_length(self.layer_idx)
        q, k = self.rotary(q, k, seqlen_offset=seqlen_offset)

        if mode == "fused_recurrent":
            o, recurrent_state = naive_recurrent_mhla(q=q, k=k, v=v, mixing_matrix=self.mixing_matrix,
                                                      initial_state=None, output_final_state=use_cache)
        elif mode == "chunk":
            o = naive_chunk_simple_mhla_fixed(q=q, k=k, v=v, mixing_matrix=self.mixing_matrix,
                                              output_final_state=use_cache)
            recurrent_state = None
        else:
            raise NotImplementedError(f"Not supported mode `{mode}`.")

        if past_key_values is not None:
            past_key_values.update(recurrent_state=recurrent_state, conv_state=None, layer_idx=self.layer_idx,
                                   offset=q_len)
        if self.use_output_gate:
            g = self.g_proj(hidden_states)
            if self.fuse_norm_and_gate:
                g = rearrange(g, "... (h d) -> ... h d", d=self.head_v_dim)
                o = rearrange(self.g_norm_swish_gate(o, g), "... h d -> ... (h d)")
            else:
                o = rearrange(self.g_norm(o), "... h d -> ... (h d)") * self.gate_fn(g)
        else:
            o = rearrange(self.g_norm(o), "... h d -> ... (h d)")
        o = self.o_proj(o)
        return o, None, past_key_values
