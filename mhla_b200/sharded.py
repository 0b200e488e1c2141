"""(b,h)-unit sharding of the MHLA operator across the GPUs of one node.

The operator has no term coupling different (batch, head) pairs - the mixing matrix is shared and read-only - so the
flattened B*H axis is partitioned contiguously over the ranks and every rank runs the single-GPU kernel on its
slice: zero communication inside the op.  Only when the consumer needs all heads on every rank (Wan's ``o`` Linear
over the full channel dim, mhla_utils.py:366) is ONE NCCL all-gather of the outputs issued, on the current stream,
right behind the kernel.  One process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch) for the plumbing.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def unit_range(n_units: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of the flattened (b,h) units owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n_units, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def mhla_sharded(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mix: torch.Tensor, *, group=None,
                 gather: bool = True, inputs: str = "full", total_units: Optional[int] = None,
                 compute: Optional[Callable] = None, **kw) -> torch.Tensor:
    """Block-mixed MHLA on (b,h)-sharded units.

    q, k, v : [G, M, w, D] with G = B*H flattened units (``inputs="full"``: every rank holds all G units and works on
              its own slice; ``inputs="local"``: the tensors already are this rank's slice, e.g. behind head-sharded
              projections).  Extra keyword tensors ``q_rope`` / ``k_rope`` follow the same convention.
    total_units : ``inputs="local"`` only - the global number of units G when the caller knows it (skips the count
              exchange, so the call enqueues the kernel and nothing else).
    gather  : all-gather the outputs so every rank returns the full [G, M, w, D]; otherwise return the local slice.
    compute : the single-GPU operator (defaults to ``mhla_b200.mhla_blockmix``; the CPU tests inject the oracle to
              exercise this host logic under gloo).
    """
    if compute is None:
        from .ops import mhla_blockmix as compute
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rope = {n: kw.pop(n) for n in ("q_rope", "k_rope") if kw.get(n) is not None}
    for n in ("q_rope", "k_rope"):
        kw.pop(n, None)
    if inputs == "full":
        G = q.shape[0]
        lo, hi = unit_range(G, world, rank)
        ql, kl, vl = q[lo:hi], k[lo:hi], v[lo:hi]
        rope = {n: t[lo:hi] for n, t in rope.items()}
    elif inputs == "local":
        ql, kl, vl = q, k, v
        counts = torch.tensor([q.shape[0]], device=q.device) if total_units is None else None
        if total_units is not None:
            G = int(total_units)
        elif world > 1:
            allc = [torch.zeros_like(counts) for _ in range(world)]
            dist.all_gather(allc, counts, group=group)
            G = int(sum(int(c) for c in allc))
        else:
            G = q.shape[0]
        lo, hi = unit_range(G, world, rank)
        if hi - lo != q.shape[0]:
            raise ValueError("local shards must follow unit_range(): contiguous, sizes differing by at most one")
    else:
        raise ValueError("inputs must be 'full' or 'local'")
    out_local = compute(ql, kl, vl, mix, **rope, **kw) if hi > lo else ql.new_empty((0,) + tuple(q.shape[1:]))
    if not gather or world == 1:
        return out_local
    # one all-gather of the outputs; uneven partitions are padded to the largest slice and trimmed afterwards
    per = -(-G // world)
    if G % world == 0:
        full = out_local.new_empty((G,) + tuple(out_local.shape[1:]))
        dist.all_gather_into_tensor(full, out_local.contiguous(), group=group)
        return full
    padded = out_local.new_zeros((per,) + tuple(out_local.shape[1:]))
    padded[: hi - lo] = out_local
    buf = out_local.new_empty((world * per,) + tuple(out_local.shape[1:]))
    dist.all_gather_into_tensor(buf, padded, group=group)
    pieces = []
    for r in range(world):
        a, b = unit_range(G, world, r)
        pieces.append(buf[r * per: r * per + (b - a)])
    return torch.cat(pieces, dim=0)


# ------------------------------------------------------------------------------------------------ block-range split
def _phase_call(q, k, v, mix, normalize, eps, q_rope, k_rope, flags_extra, ws, out=None):
    """One phase (or phase range) of the general kernel on a caller-owned workspace, through the C ABI's phase flags."""
    import ctypes as C
    from . import _capi
    from .ops import _DT, _t5
    B, H, M, w, D = q.shape
    d = _capi.BlockmixDesc()
    d.B, d.H, d.M, d.w, d.D = B, H, M, w, D
    d.dtype, d.eps = _DT[q.dtype], float(eps)
    d.flags = (_capi.FLAG_NORMALIZE if normalize else 0) | _capi.FLAG_UNFUSED | _capi.FLAG_NO_SMALLN | flags_extra
    o = out if out is not None else torch.empty_like(q)
    d.q, d.k, d.v, d.out = _t5(q), _t5(k), _t5(v), _t5(o)
    d.q_rope, d.k_rope = _t5(q_rope), _t5(k_rope)
    d.mix, d.mix_ld = mix.data_ptr(), mix.stride(0)
    L = _capi.lib()
    nbytes = L.mhla_blockmix_workspace_bytes(C.byref(d))
    lay = (C.c_size_t * 8)()
    _capi.check(L.mhla_blockmix_workspace_layout(C.byref(d), C.byref(lay)), "mhla_blockmix_workspace_layout")
    if ws is None:
        ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=q.device)
    base = (ws.data_ptr() + 1023) // 1024 * 1024 - ws.data_ptr()
    d.workspace, d.workspace_bytes = ws.data_ptr() + base, nbytes
    with torch.cuda.device(q.device):
        _capi.check(L.mhla_fwd_blockmix(C.byref(d), torch.cuda.current_stream().cuda_stream), "mhla_fwd_blockmix")
    ncols, wpad = int(lay[5]), int(lay[6])
    G = B * H
    views = dict(
        S=ws[base + lay[0]: base + lay[0] + G * M * ncols * 2].view(torch.int16).view(G, M, ncols),
        St=ws[base + lay[1]: base + lay[1] + G * M * D * D * 2].view(torch.int16).view(G, M, D * D),
        den=(ws[base + lay[2]: base + lay[2] + G * M * 2 * wpad * 4].view(torch.float32).view(G, M, 2 * wpad) if wpad else None))
    return o, ws, views


def mhla_block_sharded(q, k, v, mix, *, group=None, normalize: bool = True, eps: float = 1e-6, q_rope=None, k_rope=None):
    """Block-range split (second axis) for B*H smaller than / not divisible by the number of GPUs - Wan with
    guidance off: 12 heads on 8 GPUs.  Every rank holds ALL (b,h) units but only its contiguous range of blocks:
    q, k, v (and the roped copies) are [B, H, Mloc, w, D] with Mloc = unit_range(M, world, rank); ``mix`` is the full [M, M]
    matrix.  Three phases of the same kernel with ONE exchange step between them:
      1. block summaries S_j (+ ksum / n_loc) of the local blocks                       (phase 1 on the local blocks)
      2. all-gather of the summaries - M * (D^2 + 2 wpad) 16-bit values per unit - and the block mixing of the full set
         (phase 2; small and identical on every rank: replicated rather than exchanged a second time)
      3. readout of the local blocks against their rows of S~ / den                      (phase 3 on the local blocks)
    Returns the local [B, H, Mloc, w, D] slice of the output: no gather is needed when the consumer is token-sharded."""
    from . import _capi
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B, H, Mloc, w, D = q.shape
    M = mix.shape[0]
    lo, hi = unit_range(M, world, rank)
    if hi - lo != Mloc:
        raise ValueError("local block slices must follow unit_range(M, world, rank)")
    mixf = mix.detach().reshape(M, M).float().contiguous()
    dummy = mixf[:Mloc, :Mloc].contiguous()
    # 1. local summaries
    _, ws_loc, v_loc = _phase_call(q, k, v, dummy, normalize, eps, q_rope, k_rope, _capi.FLAG_STOP_AFTER_P1, None)
    # 2. exchange + mixing of the full set
    per = -(-M // world)
    G, ncols = B * H, v_loc["S"].shape[-1]
    send = v_loc["S"].new_zeros((G, per, ncols))
    send[:, :Mloc] = v_loc["S"]
    recv = send.new_empty((world, G, per, ncols))
    if world > 1:   # (NCCL has no int16: the 16-bit summaries travel as raw bytes)
        dist.all_gather_into_tensor(recv.view(torch.uint8), send.view(torch.uint8), group=group)
    else:
        recv[0] = send
    qf = q.new_empty((B, H, M, w, D))          # shape carrier for the full descriptor (phase 2 touches no q/k/v)
    ws_full = None
    # (allocate the full workspace through a dry layout query: phase 2 only)
    import ctypes as C
    from .ops import _DT
    d = _capi.BlockmixDesc()
    d.B, d.H, d.M, d.w, d.D = B, H, M, w, D
    d.dtype, d.flags, d.eps = _DT[q.dtype], (_capi.FLAG_NORMALIZE if normalize else 0) | _capi.FLAG_UNFUSED | _capi.FLAG_NO_SMALLN, float(eps)
    if q_rope is not None:
        d.q_rope.ptr = d.k_rope.ptr = 1
    L = _capi.lib()
    nbytes = L.mhla_blockmix_workspace_bytes(C.byref(d))
    lay = (C.c_size_t * 8)()
    _capi.check(L.mhla_blockmix_workspace_layout(C.byref(d), C.byref(lay)), "mhla_blockmix_workspace_layout")
    ws_full = torch.empty(nbytes + 1024, dtype=torch.uint8, device=q.device)
    base = (ws_full.data_ptr() + 1023) // 1024 * 1024 - ws_full.data_ptr()
    S_full = ws_full[base + lay[0]: base + lay[0] + G * M * ncols * 2].view(torch.int16).view(G, M, ncols)
    for r in range(world):
        a, b = unit_range(M, world, r)
        S_full[:, a:b] = recv[r, :, : b - a]
    qr_f = qf if q_rope is not None else None
    _, _, v_full = _phase_call(qf, qf, qf, mixf, normalize, eps, qr_f, qr_f, _capi.FLAG_ONLY_P2, ws_full)
    # 3. readout of the local blocks
    v_loc["St"].copy_(v_full["St"][:, lo:hi])
    if normalize:
        v_loc["den"].copy_(v_full["den"][:, lo:hi])
    out, _, _ = _phase_call(q, k, v, dummy, normalize, eps, q_rope, k_rope, _capi.FLAG_ONLY_P3, ws_loc)
    return out
