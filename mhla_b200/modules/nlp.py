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

from ..decode import MHLAState, causal_with_state, get_unpad_data, pad_input
from ..ops import mhla_causal, naive_chunk_simple_mhla_fixed, naive_recurrent_mhla  # noqa: F401


class Cache:
    """The slice of ``fla.models.utils.Cache`` the layer uses (layers/mhla.py:247-252,301-348): per-layer dicts with
    ``recurrent_state`` / ``conv_state``, ``update(...)`` and ``get_seq_length``.  ``recurrent_state`` holds an
    ``MHLAState`` (completed chunk summaries + the open chunk's k, v rows) - a state that actually continues the
    sequence, unlike the reference's all-zero one."""

    def __init__(self, seen_tokens: int = 0):
        self.states = []
        self._seen_tokens = seen_tokens

    def __getitem__(self, layer_idx):
        return self.states[layer_idx]

    def __len__(self):
        return len(self.states)

    def update(self, recurrent_state=None, attn_state=None, conv_state=None, ffn_state=None, layer_idx=0, offset=1,
               cache_kwargs=None):
        if layer_idx == 0:
            self._seen_tokens += offset
        if len(self.states) <= layer_idx:
            self.states.append(dict(recurrent_state=recurrent_state, attn_state=attn_state, conv_state=conv_state,
                                    ffn_state=ffn_state))
        else:
            st = self.states[layer_idx]
            if recurrent_state is not None:
                st["recurrent_state"] = recurrent_state
            if conv_state is not None:
                st["conv_state"] = conv_state
        return self.states[layer_idx]

    def get_seq_length(self, layer_idx=0):
        if len(self.states) <= layer_idx:
            return 0
        return self._seen_tokens


class RotaryEmbedding(nn.Module):
    """Non-interleaved (rotate-half) rotary embedding over the full head dim, base 10000."""

    def __init__(self, dim: int, base: float = 10000.0):
        super().__init__()
        self.dim, self.base = dim, float(base)
        inv = 1.0 / (self.base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
        self.register_buffer("inv_freq", inv, persistent=False)

    def forward(self, q, k, seqlen_offset=0, max_seqlen=None, cu_seqlens=None):
        T = q.shape[1]
        off = int(seqlen_offset) if not torch.is_tensor(seqlen_offset) else seqlen_offset
        t = torch.arange(T, device=q.device, dtype=torch.float32)
        if cu_seqlens is not None:
            # packed batch [1, total, ...]: positions restart at every sequence start (fla/modules/rotary.py, varlen path)
            cu = cu_seqlens.to(device=q.device, dtype=torch.long)
            seq_id = torch.bucketize(torch.arange(T, device=q.device), cu[1:], right=True)
            t = (t - cu[seq_id].float())
            if torch.is_tensor(off):
                t = t + off.to(t)[seq_id]
            else:
                t = t + off
            t = t[None, :]
        elif torch.is_tensor(off):
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
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or g.requires_grad or
                                                  (self.weight is not None and self.weight.requires_grad))
        if (x.is_cuda and not needs_grad and x.dtype in (torch.bfloat16, torch.float16)
                and x.shape[-1] in (64, 128, 256) and g.shape == x.shape):
            from ..ops import gated_rmsnorm      # one launch: csrc/gated_norm_kernel.cuh
            return gated_rmsnorm(x, g, self.weight, self.eps)
        y = x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + self.eps)
        if self.weight is not None:
            y = y * self.weight.float()
        gf = g.float()
        return (y * gf * torch.sigmoid(gf)).to(x.dtype)


class ShortConvolution(nn.Conv1d):
    """Depthwise causal conv over time followed by SiLU (fla/modules/convolution.py); weight [hidden, 1, k]."""

    def __init__(self, hidden_size, kernel_size, bias=False, activation="silu"):
        super().__init__(hidden_size, hidden_size, kernel_size, groups=hidden_size, bias=bias, padding=0)
        self.hidden_size, self.activation = hidden_size, activation

    def forward(self, x, cache=None, output_final_state=False, cu_seqlens=None):
        """x [B, T, C] -> (y, final_state): causal depthwise conv.  ``cache`` [B, C, k-1] holds the previous inputs
        (decode); ``cu_seqlens`` (packed batch, B = 1) keeps the window from reaching across sequence starts."""
        B, T, Cc = x.shape
        kw_ = self.kernel_size[0]
        xt = x.transpose(1, 2)                                                      # [B, C, T]
        left = cache.to(xt.dtype) if cache is not None else xt.new_zeros(B, Cc, kw_ - 1)
        xp = torch.cat([left, xt], dim=-1)                                          # [B, C, T + k - 1]
        if cu_seqlens is not None:
            # zero the taps that would read the previous sequence: tap d of token t is valid iff t - d >= its sequence start
            cu = cu_seqlens.to(device=x.device, dtype=torch.long)
            pos = torch.arange(T, device=x.device)
            start = cu[torch.bucketize(pos, cu[1:], right=True)]
            w = self.weight[:, 0, :]                                                # [C, k], tap j multiplies x[t - (k-1-j)]
            y = xt.new_zeros(B, Cc, T)
            for j in range(kw_):
                d = kw_ - 1 - j
                ok = (pos - d >= start).to(xt.dtype)
                y = y + xp[..., j:j + T] * w[:, j].view(1, -1, 1) * ok.view(1, 1, -1)
            if self.bias is not None:
                y = y + self.bias.view(1, -1, 1)
        else:
            y = F.conv1d(xp, self.weight, self.bias, groups=Cc)
        y = y.transpose(1, 2)
        if self.activation in ("silu", "swish"):
            y = F.silu(y)
        final = xp[..., -(kw_ - 1):].contiguous() if output_final_state else None
        return y, final


_ACT = {"swish": F.silu, "silu": F.silu, "sigmoid": torch.sigmoid, "relu": F.relu, "gelu": F.gelu}


class MHLA(nn.Module):
    def __init__(self, mode: str = "chunk", hidden_size: int = 1024, expand_k: float = 0.5, expand_v: float = 1.0,
                 num_heads: int = 4, num_kv_heads: Optional[int] = None, feature_map: Optional[str] = None,
                 use_short_conv: bool = False, conv_size: int = 4, conv_bias: bool = False,
                 use_output_gate: bool = True, gate_fn: str = "swish", elementwise_affine: Optional[bool] = True,
                 norm_eps: float = 1e-5, gate_logit_normalizer: int = 16, gate_low_rank_dim: int = 16,
                 clamp_min: Optional[float] = None, fuse_norm: bool = True, layer_idx: int = None,
                 max_chunks: int = 32, varlen: str = "reference"):
        """``max_chunks`` (extension): side of the mixing matrix, 32 in the reference (layers/mhla.py:196) - sequences
        longer than 32 * 64 tokens need more.  ``varlen`` (extension): "reference" evaluates a padded batch as the
        reference does - un-padded and packed into ONE sequence (layers/mhla.py:254-256; chunks run across sequence
        borders); "per_sequence" evaluates every sequence on its own."""
        super().__init__()
        self.mode = mode
        self.varlen = varlen
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

        L = int(max_chunks)                                            # 32: layers/mhla.py:196-200
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
        batch_size, q_len, _ = hidden_states.shape
        layer = self.layer_idx if self.layer_idx is not None else 0
        last_state = None
        if past_key_values is not None and len(past_key_values) > layer:
            last_state = past_key_values[layer]

        cu_seqlens = kwargs.get("cu_seqlens", None)
        indices = None
        if attention_mask is not None and not bool(attention_mask[:, -q_len:].all()):
            if last_state is not None:
                raise NotImplementedError("decoding with a padded batch is not supported (left-pad the prompts instead)")
            indices, cu_seqlens, _ = get_unpad_data(attention_mask[:, -q_len:])            # layers/mhla.py:254-256
            hidden_states = hidden_states.reshape(batch_size * q_len, -1)[indices].unsqueeze(0)

        conv_state = None
        if self.use_short_conv:
            cq = ck = cv = None
            if last_state is not None and last_state.get("conv_state") is not None:
                cq, ck, cv = last_state["conv_state"]
            q, cq = self.q_conv1d(self.q_proj(hidden_states), cache=cq, output_final_state=use_cache, cu_seqlens=cu_seqlens)
            k, ck = self.k_conv1d(self.k_proj(hidden_states), cache=ck, output_final_state=use_cache, cu_seqlens=cu_seqlens)
            v, cv = self.v_conv1d(self.v_proj(hidden_states), cache=cv, output_final_state=use_cache, cu_seqlens=cu_seqlens)
            conv_state = (cq, ck, cv)
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
            seqlen_offset = past_key_values.get_seq_length(layer)
        q, k = self.rotary(q, k, seqlen_offset=seqlen_offset, cu_seqlens=cu_seqlens)

        prev = last_state["recurrent_state"] if last_state is not None else None
        if self.mode not in ("chunk", "fused_recurrent"):
            raise NotImplementedError(f"Not supported mode `{self.mode}`.")
        recurrent_state = None
        if isinstance(prev, MHLAState) or use_cache:
            # stateful evaluation: the prompt runs through the kernel, continuations are evaluated from the chunk
            # summaries kept in the cache (mhla_b200/decode.py)
            o, recurrent_state = causal_with_state(q, k, v, self.mixing_matrix, prev, prefill_op=mhla_causal)
        elif indices is not None and self.varlen == "per_sequence":
            cu = cu_seqlens.tolist()
            o = torch.cat([mhla_causal(q[:, a:b], k[:, a:b], v[:, a:b], self.mixing_matrix) for a, b in zip(cu[:-1], cu[1:])], dim=1)
        elif q.shape[1] <= 64:        # layers/mhla.py:247: 'fused_recurrent' for short inputs; equals the chunk form there
            o, _ = naive_recurrent_mhla(q=q, k=k, v=v, mixing_matrix=self.mixing_matrix, initial_state=None,
                                        output_final_state=False)
        else:
            o = naive_chunk_simple_mhla_fixed(q=q, k=k, v=v, mixing_matrix=self.mixing_matrix, output_final_state=use_cache)

        if past_key_values is not None:
            past_key_values.update(recurrent_state=recurrent_state, conv_state=conv_state if self.use_short_conv else None,
                                   layer_idx=layer, offset=q_len)
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
        if indices is not None:
            o = pad_input(o.squeeze(0), indices, batch_size, q_len)                         # layers/mhla.py:362-363
        return o, None, past_key_values
