"""Incremental (prefill + decode) evaluation of the causal MHLA operator and variable-length helpers
(reference call sites mhla_nlp/fla/layers/mhla.py:249-256 and :301-348, op naive.py:10-142).

The reference has no usable decode path: ``naive_recurrent_mhla`` returns an all-zero "final state" and its recurrence
disagrees with the chunk form beyond the first chunk, so generation silently loses the prefix.  The
operator itself, however, has an exact incremental form.  With chunks of c = 64 tokens,
    o_t = scale * ( q_t . sum_{j < i} mm[i, j] S_j  +  mm[i, i] * sum_{s <= t, s in chunk i} (q_t . k_s) v_s ),   i = t // c,
so the state after T tokens is: the summaries S_j = k_j^T v_j of the COMPLETED chunks ([B, H, n, K, V] - they never
change again) and the k, v rows of the current partial chunk (at most c - 1 tokens).  ``MHLAState`` holds exactly that.

Prefill runs the CUDA kernel on the prompt; the per-chunk summaries of the prompt are produced by one batched matmul
(only when a cache is requested).  A decode step touches one token per sequence: n small dot products, done with torch
ops on the tensors' own device - launch-bound bookkeeping, not a streaming hot path.  Everything here is plain torch and
device-agnostic, so the CPU test-suite checks prefill + token-by-token decode against the oracle on the full sequence.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class MHLAState:
    """Decode state of one layer: completed chunk summaries and the open chunk's k, v rows."""
    S: torch.Tensor            # [B, H, n_done, K, V] fp32
    k_tail: torch.Tensor       # [B, t_open, H, K]   (0 <= t_open < chunk)
    v_tail: torch.Tensor       # [B, t_open, H, V]
    chunk: int = 64

    @property
    def seen_tokens(self) -> int:
        return self.S.shape[2] * self.chunk + self.k_tail.shape[1]


def empty_state(B, H, K, V, device, dtype, chunk=64) -> MHLAState:
    return MHLAState(torch.zeros(B, H, 0, K, V, dtype=torch.float32, device=device),
                     torch.zeros(B, 0, H, K, dtype=dtype, device=device), torch.zeros(B, 0, H, V, dtype=dtype, device=device), chunk)


def _mm2d(mixing_matrix):
    L = mixing_matrix.shape[0]
    return mixing_matrix.detach().reshape(L, mixing_matrix.shape[1]).float()


def state_from_prompt(k, v, chunk: int = 64) -> MHLAState:
    """State after a prompt of T tokens (k [B,T,H,K], v [B,T,H,V]): summaries of the T // chunk complete chunks + tail."""
    B, T, H, K = k.shape
    V = v.shape[-1]
    n = T // chunk
    kc = k[:, :n * chunk].float().reshape(B, n, chunk, H, K).permute(0, 3, 1, 2, 4)      # b h n c k
    vc = v[:, :n * chunk].float().reshape(B, n, chunk, H, V).permute(0, 3, 1, 2, 4)
    S = torch.matmul(kc.transpose(-2, -1), vc)                                           # [B,H,n,K,V]
    return MHLAState(S, k[:, n * chunk:].contiguous(), v[:, n * chunk:].contiguous(), chunk)


def causal_with_state(q, k, v, mixing_matrix, state: Optional[MHLAState], scale: Optional[float] = None, chunk: int = 64,
                      prefill_op: Optional[Callable] = None) -> Tuple[torch.Tensor, MHLAState]:
    """o for the new tokens q,k [B,T,H,K], v [B,T,H,V] continuing ``state`` (None = start of sequence), and the new state.

    A fresh sequence is handed to ``prefill_op(q, k, v, mixing_matrix)`` (the CUDA kernel) when given; continuations - the
    decode steps - are evaluated directly from the state."""
    B, T, H, K = q.shape
    V = v.shape[-1]
    sc = float(K ** -0.5 if scale is None else scale)
    mm = _mm2d(mixing_matrix).to(q.device)
    if state is None or state.seen_tokens == 0:
        n_total = (T + chunk - 1) // chunk
        if n_total > mm.shape[0]:
            raise IndexError(f"mixing matrix is {mm.shape[0]}x{mm.shape[0]} but {n_total} chunks are needed")
        if prefill_op is not None:
            o = prefill_op(q, k, v, mixing_matrix)
        else:
            o = _direct(q, k, v, mm, None, sc, chunk)
        return o, state_from_prompt(k, v, chunk)
    o = _direct(q, k, v, mm, state, sc, chunk)
    k_all, v_all = torch.cat([state.k_tail, k], dim=1), torch.cat([state.v_tail, v], dim=1)
    new = state_from_prompt(k_all, v_all, chunk)
    return o, MHLAState(torch.cat([state.S, new.S], dim=2), new.k_tail, new.v_tail, chunk)


def _direct(q, k, v, mm, state, sc, chunk):
    """Reference-exact evaluation from (state, new tokens) with fp32 torch ops; the new tokens may span chunk borders."""
    B, T, H, K = q.shape
    V = v.shape[-1]
    n_done = 0 if state is None else state.S.shape[2]
    t_open = 0 if state is None else state.k_tail.shape[1]
    k_all = k if state is None else torch.cat([state.k_tail, k], dim=1)     # tokens of the open chunk onwards
    v_all = v if state is None else torch.cat([state.v_tail, v], dim=1)
    Ta = k_all.shape[1]
    n_new = (Ta + chunk - 1) // chunk
    if n_done + n_new > mm.shape[0]:
        raise IndexError(f"mixing matrix is {mm.shape[0]}x{mm.shape[0]} but {n_done + n_new} chunks are needed")
    pad = n_new * chunk - Ta
    kf = F.pad(k_all.float(), (0, 0, 0, 0, 0, pad)).reshape(B, n_new, chunk, H, K).permute(0, 3, 1, 2, 4)   # b h n c k
    vf = F.pad(v_all.float(), (0, 0, 0, 0, 0, pad)).reshape(B, n_new, chunk, H, V).permute(0, 3, 1, 2, 4)
    qf = F.pad(torch.cat([q.new_zeros(B, t_open, H, K), q], dim=1).float(), (0, 0, 0, 0, 0, pad))
    qf = qf.reshape(B, n_new, chunk, H, K).permute(0, 3, 1, 2, 4)
    S_new = torch.matmul(kf.transpose(-2, -1), vf)                                         # [B,H,n_new,K,V] (last may be partial)
    S_all = S_new if state is None else torch.cat([state.S.to(S_new.device), S_new], dim=2)
    rows = mm[n_done:n_done + n_new, :n_done + n_new]                                      # mixing rows of the new chunks
    lower = torch.tril(rows, diagonal=n_done - 1)                                          # strictly before the chunk itself
    prefix = torch.einsum("ij,bhjkv->bhikv", lower, S_all)                                 # [B,H,n_new,K,V]
    diag = torch.diagonal(rows, offset=n_done)                                             # mm[i, i]
    tril = torch.tril(torch.ones(chunk, chunk, dtype=torch.float32, device=q.device))
    attn = torch.matmul(qf, kf.transpose(-2, -1)) * tril
    o = torch.matmul(qf, prefix) + diag.view(1, 1, -1, 1, 1) * torch.matmul(attn, vf)
    o = (o * sc).permute(0, 2, 3, 1, 4).reshape(B, n_new * chunk, H, V)[:, t_open:t_open + T]
    return o.to(q.dtype)


# ------------------------------------------------------------------------------------------------ variable length
def get_unpad_data(attention_mask: torch.Tensor):
    """indices of the real tokens in the flattened [B*T] batch, cu_seqlens [B+1] (int32) and the longest length -
    the triple ``fla.layers.utils.get_unpad_data`` returns (layers/mhla.py:254)."""
    lens = attention_mask.sum(dim=-1, dtype=torch.int32)
    indices = torch.nonzero(attention_mask.flatten(), as_tuple=False).flatten()
    cu = F.pad(torch.cumsum(lens, dim=0, dtype=torch.int32), (1, 0))
    return indices, cu, int(lens.max())


def pad_input(x: torch.Tensor, indices: torch.Tensor, batch: int, seqlen: int) -> torch.Tensor:
    """Scatter packed rows [total, ...] back to [batch, seqlen, ...] (zeros at the padding) - layers/mhla.py:363."""
    out = x.new_zeros((batch * seqlen,) + tuple(x.shape[1:]))
    out[indices] = x
    return out.view(batch, seqlen, *x.shape[1:])
