"""CPU oracle for the MHLA operator (TEST INFRASTRUCTURE, not product code).

A torch-CPU fp32/fp64 restatement of the reference's algorithm for the hot path named in
SURVEY.md section 8(a).  Every function cites the reference file:line it follows (paths relative to
the upstream checkout, DAGroup-PKU/MHLA @ ccf97b2).  The restatement is *pinned*: the fixtures in
``tests/golden/*.npz`` were produced by importing and executing the reference's own code in the build
container (``tests/golden/make_golden.py``), and ``tests/test_oracle_golden.py`` checks this file
against them.

The reference is pure PyTorch, so the port keeps PyTorch-CPU ops (``torch.matmul`` and the 1x1
block-mixing convolution written as an einsum) - this is also what ``bench.py`` times as the
``cpu_baseline`` (kind "port").
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch


# ------------------------------------------------------------------------------------------------
# Mixing matrices (rows A0 / B0 of SURVEY.md 8a)
# ------------------------------------------------------------------------------------------------
def block_distance_matrix(
    blocks_layout: Sequence[int],
    transform: str = "linear",
    local_thres: float = 1.5,
    exp_sigma: float = 3.0,
) -> torch.Tensor:
    """W[M, M] of ``BlockDistanceConv`` (2-D layout, mhla_dit/mhla/mhla.py:63-122) and
    ``BlockDistanceConv3D`` (3-D layout, mhla_videogen/diffusion/model/wan/mhla_utils.py:60-118).

    Block centres sit at (i+.5, j+.5[, k+.5]) in raster order; W is a transform of the Euclidean
    centre distance, column-normalised (``mat / mat.sum(dim=0)``) except for "gaussian".
    """
    grids = torch.meshgrid(*[torch.arange(n, dtype=torch.float32) + 0.5 for n in blocks_layout], indexing="ij")
    centres = torch.stack([g.reshape(-1) for g in grids], dim=-1)  # raster order == nested python loops
    diff = centres[:, None, :] - centres[None, :, :]
    dist = torch.linalg.vector_norm(diff, ord=2, dim=-1)
    if transform == "linear":
        mat = 1.0 - dist / dist.max()
        return mat / mat.sum(dim=0, keepdim=True)
    if transform == "cos":
        mat = torch.cos(dist / dist.max() * math.pi / 4)
        return mat / mat.sum(dim=0, keepdim=True)
    if transform == "exp":
        mat = torch.exp(-dist / exp_sigma)
        return mat / mat.sum(dim=0, keepdim=True)
    if transform == "gaussian":
        sigma = dist.max() / 3
        return torch.exp(-(dist ** 2) / (2 * sigma ** 2))
    if transform == "local":
        mat = (dist <= local_thres).float()
        return mat / mat.sum(dim=0, keepdim=True)
    raise ValueError(f"Unknown transform: {transform}")


# ------------------------------------------------------------------------------------------------
# Variant A / B: non-causal block-mixed linear attention (rows A2-A4, B3)
# ------------------------------------------------------------------------------------------------
def blockmix_summaries(k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """S_j = K_j^T V_j, [..., M, Dk, Dv]  (mhla_dit/mhla/mhla.py:262, mhla_utils.py:331)."""
    return torch.matmul(k.transpose(-2, -1), v)


def blockmix_fwd(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    W: torch.Tensor,
    eps: float = 1e-6,
    normalize: bool = True,
    q_rope: Optional[torch.Tensor] = None,
    k_rope: Optional[torch.Tensor] = None,
    dtype: torch.dtype = torch.float32,
) -> torch.Tensor:
    """out[..., M, w, Dv] for block-major q,k [..., M, w, Dk], v [..., M, w, Dv], W [M, M].

    Follows mhla_dit/mhla/mhla.py:262-268 (variant A) and
    mhla_videogen/diffusion/model/wan/mhla_utils.py:328-341 (variant B: roped q/k feed the numerator,
    un-roped q/k feed the normaliser).  The 1x1 ``Conv2d`` over the block axis is
    ``y[:, i] = sum_j W[i, j] x[:, j]`` (mhla.py:124-134).

    The normaliser reproduces the reference's quirk: ``matmul(q, k_sum)`` is the
    *local* n_loc[j, t] = q_{j,t} . ksum_j and the conv then mixes equal in-block indices t across
    blocks: den[i, t] = sum_j W[i, j] n_loc[j, t] + eps.
    """
    q, k, v, W = (t.to(dtype) for t in (q, k, v, W))
    qn = q if q_rope is None else q_rope.to(dtype)
    kn = k if k_rope is None else k_rope.to(dtype)
    kv = torch.matmul(kn.transpose(-2, -1), v)                 # [..., M, Dk, Dv]       (:262 / :331)
    kv = torch.einsum("ij,...jab->...iab", W, kv)              # piece_attn / block_attn (:263 / :332)
    out = torch.matmul(qn, kv)                                 # (:268 / :339-341)
    if normalize:
        k_sum = k.sum(dim=-2)                                  # [..., M, Dk]            (:265 / :335)
        n_loc = torch.einsum("...jtd,...jd->...jt", q, k_sum)  # matmul(q, k_sum)        (:266 / :337)
        den = torch.einsum("ij,...jt->...it", W, n_loc) + eps  # conv over blocks + eps
        out = out / den.unsqueeze(-1)
    return out


# ------------------------------------------------------------------------------------------------
# Variant C: causal chunked (row C2) and the token-recurrent form restricted to one chunk (row C3)
# ------------------------------------------------------------------------------------------------
def _mm2d(mixing_matrix: torch.Tensor) -> torch.Tensor:
    L = mixing_matrix.shape[0]
    return mixing_matrix.reshape(L, mixing_matrix.shape[1]).to(torch.float32)


def causal_chunk_fwd(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    mixing_matrix: torch.Tensor,
    chunk_size: int = 64,
    dtype: torch.dtype = torch.float32,
) -> torch.Tensor:
    """o[B, T, H, V] for q,k [B, T, H, K], v [B, T, H, V], mixing_matrix [L, L, (1,1,1,1)].

    Follows mhla_nlp/fla/ops/mhla/naive.py:10-83 step by step: fp32 (:39), zero-pad T to a multiple of
    the chunk (:46-51), q *= K^-1/2 (:42,58), S_j = k_j^T v_j (:59-64), then per chunk i
    o_i = q_i (sum_{j<i} mm[i,j] S_j) + mm[i,i] ((q_i k_i^T) * tril) v_i (:66-78); un-pad, cast (:82).
    """
    out_dtype = q.dtype
    mm = _mm2d(mixing_matrix).to(dtype)
    q, k, v = (t.transpose(1, 2).to(dtype) for t in (q, k, v))       # b h t d
    B, H, T, K = q.shape
    V = v.shape[-1]
    c = chunk_size
    pad = (c - T % c) % c
    if pad:
        q, k, v = (torch.nn.functional.pad(t, (0, 0, 0, pad)) for t in (q, k, v))
    n = (T + pad) // c
    if n > mm.shape[0]:
        raise IndexError(f"mixing matrix is {mm.shape[0]}x{mm.shape[0]} but {n} chunks are needed")
    mm = mm[:n, :n]
    q = q.reshape(B, H, n, c, K) * (K ** -0.5)
    k = k.reshape(B, H, n, c, K)
    v = v.reshape(B, H, n, c, V)
    S = torch.matmul(k.transpose(-2, -1), v)                              # [B,H,n,K,V]
    tril = torch.tril(torch.ones(c, c, dtype=dtype))
    o = torch.zeros_like(v)
    for i in range(n):
        attn = torch.matmul(q[:, :, i], k[:, :, i].transpose(-2, -1)) * tril
        prefix = torch.einsum("j,bhjkv->bhkv", mm[i, :i], S[:, :, :i])
        o[:, :, i] = torch.matmul(q[:, :, i], prefix) + mm[i, i] * torch.matmul(attn, v[:, :, i])
    o = o.reshape(B, H, n * c, V)[:, :, :T].transpose(1, 2)
    return o.to(out_dtype)


def causal_closed_form(q, k, v, mixing_matrix, chunk_size: int = 64, dtype=torch.float64) -> torch.Tensor:
    """Independent cross-check: O = K^-1/2 ((Q K^T) * Mask) V with
    Mask[t, s] = mm[t // c, s // c] * 1[s <= t].  Quadratic - small cases only."""
    mm = _mm2d(mixing_matrix).to(dtype)
    q, k, v = (t.transpose(1, 2).to(dtype) for t in (q, k, v))
    T, K = q.shape[-2], q.shape[-1]
    idx = torch.arange(T)
    mask = mm[idx[:, None] // chunk_size, idx[None, :] // chunk_size] * (idx[None, :] <= idx[:, None]).to(dtype)
    o = torch.matmul((torch.matmul(q, k.transpose(-2, -1)) * mask), v) * (K ** -0.5)
    return o.transpose(1, 2)


def recurrent_first_chunk_fwd(q, k, v, mixing_matrix, chunk_size: int = 64, dtype=torch.float32):
    """``naive_recurrent_mhla`` (naive.py:88-142) restricted to T <= chunk_size, the only regime in which
    the layer selects it (layers/mhla.py:247) and in which it agrees with the chunk form: token t reads
    mm[0,0] * sum_{s<=t} k_s v_s^T.  Returns (o, None): the reference's "final state" is all zeros /
    unusable, so no state is modelled."""
    T = q.shape[1]
    if T > chunk_size:
        raise ValueError("recurrent form is only defined for T <= chunk_size (SURVEY.md 0.4)")
    return causal_chunk_fwd(q, k, v, mixing_matrix, chunk_size, dtype), None


# ------------------------------------------------------------------------------------------------
# Wan 3-axis RoPE (row B1) - pre-op, restated so module-level parity can be checked
# ------------------------------------------------------------------------------------------------
def _rope_params(max_seq_len: int, dim: int, theta: float = 10000.0) -> torch.Tensor:
    """wan/model.py:139-146."""
    freqs = torch.outer(
        torch.arange(max_seq_len), 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim))
    )
    return torch.polar(torch.ones_like(freqs), freqs)


def rope_freqs_wan(head_dim: int, max_seq_len: int = 1024) -> torch.Tensor:
    """complex128 [max_seq_len, head_dim/2] table, wan/model.py:1933-1936."""
    d = head_dim
    return torch.cat(
        [_rope_params(max_seq_len, d - 4 * (d // 6)), _rope_params(max_seq_len, 2 * (d // 6)),
         _rope_params(max_seq_len, 2 * (d // 6))], dim=1)


def rope_apply_wan(x: torch.Tensor, grid: Tuple[int, int, int], freqs: torch.Tensor) -> torch.Tensor:
    """x [B, N, H, D] -> roped fp32, interleaved-pair rotation with per-axis (f,h,w) frequency bands
    (mhla_utils.py:127-156).  All samples share one grid here (the module uses grid_sizes[0])."""
    B, N, H, D = x.shape
    c = D // 2
    f, h, w = grid
    assert f * h * w == N
    fr = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    fi = torch.cat([
        fr[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1),
        fr[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
        fr[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(N, 1, -1)
    xc = torch.view_as_complex(x.to(torch.float64).reshape(B, N, H, c, 2))
    return torch.view_as_real(xc * fi).flatten(3).float()


# ------------------------------------------------------------------------------------------------
# Error metric used by every parity test (fla/utils.py:72-93 ``get_err_ratio``)
# ------------------------------------------------------------------------------------------------
def err_ratio(ref: torch.Tensor, out: torch.Tensor) -> float:
    ref = ref.double().flatten()
    out = out.double().flatten()
    err = (ref - out).square().mean().sqrt().item()
    base = ref.square().mean().sqrt().item()
    return err / (base + 1e-12)
