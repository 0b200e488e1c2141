"""world_size-2 `gloo` tests (CPU) of the (b,h)-unit sharding host logic.  The per-rank compute is injected (the CPU
oracle) - the CUDA kernel itself is covered by the `-m gpu` tests; what is checked here is partitioning, the single
all-gather and the shard-equivalence property."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from mhla_b200.sharded import mhla_sharded, unit_range


def test_unit_range_partitions_contiguously():
    for n in (1, 5, 24, 32):
        for world in (1, 2, 3, 8):
            spans = [unit_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, G, inputs_mode, q_out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        M, w, D = 4, 8, 16
        q, k, v = torch.rand(G, M, w, D, generator=g), torch.rand(G, M, w, D, generator=g), torch.randn(G, M, w, D, generator=g)
        W = oracle.block_distance_matrix((2, 2), "linear")
        compute = lambda a, b, c, mix, **kw: oracle.blockmix_fwd(a, b, c, mix, **kw)  # noqa: E731
        if inputs_mode == "full":
            out = mhla_sharded(q, k, v, W, gather=True, inputs="full", compute=compute, normalize=True)
        else:
            lo, hi = unit_range(G, world, rank)
            out = mhla_sharded(q[lo:hi], k[lo:hi], v[lo:hi], W, gather=True, inputs="local", compute=compute, normalize=True)
        ref = oracle.blockmix_fwd(q, k, v, W, normalize=True)
        local = mhla_sharded(q, k, v, W, gather=False, inputs="full", compute=compute, normalize=True)
        lo, hi = unit_range(G, world, rank)
        # (bitwise shard-equivalence of the CUDA kernel is asserted in tests/test_blockmix_gpu.py; the injected CPU
        # oracle's batched matmul may block differently on a slice, hence a tolerance here)
        ok = torch.allclose(out, ref, rtol=1e-5, atol=1e-6) and torch.allclose(local, ref[lo:hi], rtol=1e-5, atol=1e-6)
        q_out.put((rank, bool(ok), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("G,inputs_mode", [(6, "full"), (5, "full"), (5, "local"), (1, "full")])
def test_sharded_equals_single_rank_gloo(G, inputs_mode):
    world = 2
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, G, inputs_mode, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = sorted(q_out.get(timeout=10) for _ in range(world))
    assert all(ok for _, ok, _ in results), results
    assert all(shape[0] == G for _, _, shape in results)
