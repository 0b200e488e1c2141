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
