"""NCCL tests of the (b,h)-unit sharding on real GPUs (`-m gpu`, needs >= 2 devices; the driver's single-GPU run skips
them - `gpurun --gpus 2 -- python -m pytest tests/test_sharded_nccl.py -m gpu` is how the log in profiles/ was made).
One process per GPU, torch.distributed over NCCL; every rank runs the single-GPU kernel on its slice of the units and ONE
all_gather_into_tensor reassembles the outputs.  The result must be BITWISE equal to the single-GPU run."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, normalize, q_out):
    import torch.distributed as dist
    import mhla_b200
    from mhla_b200.sharded import mhla_sharded, unit_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        G, M, w, D = shape
        g = torch.Generator().manual_seed(0)
        q = (torch.relu(torch.randn(G, M, w, D, generator=g)) + 1e-6).bfloat16().to(dev)
        k = (torch.relu(torch.randn(G, M, w, D, generator=g)) + 1e-6).bfloat16().to(dev)
        v = torch.randn(G, M, w, D, generator=g).bfloat16().to(dev)
        W = (torch.rand(M, M, generator=g) / M).to(dev)
        full = mhla_b200.mhla(q, k, v, W, normalize=normalize)                      # single-GPU reference on this rank
        lo, hi = unit_range(G, world, rank)
        a = mhla_sharded(q, k, v, W, gather=True, inputs="full", normalize=normalize)
        b = mhla_sharded(q[lo:hi], k[lo:hi], v[lo:hi], W, gather=True, inputs="local", total_units=G, normalize=normalize)
        c = mhla_sharded(q[lo:hi], k[lo:hi], v[lo:hi], W, gather=False, inputs="local", total_units=G, normalize=normalize)
        torch.cuda.synchronize()
        if M > 64:
            ok = bool(torch.equal(a, full) and torch.equal(b, full) and torch.equal(c, full[lo:hi]))
        else:
            # M <= 64: the kernel packs floor(128 / M) units into one 128-row mixing tile when that divides the unit count,
            # so a rank's slice may be mixed with a different packing factor than the full batch - same products, another
            # fp32 summation grouping (a few 16-bit roundings flip).  Equal up to that, and bitwise among the sharded forms.
            rel = lambda x, y: float((x.float() - y.float()).norm() / y.float().norm())  # noqa: E731
            ok = bool(rel(a, full) < 1e-3 and torch.equal(a, b) and torch.equal(c, a[lo:hi]))
        q_out.put((rank, ok, tuple(a.shape)))
    finally:
        dist.destroy_process_group()


def _run(world, shape, normalize):
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, normalize, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    results = sorted(q_out.get(timeout=10) for _ in range(world))
    assert all(ok for _, ok, _ in results), results
    assert all(s[0] == shape[0] for _, _, s in results)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("shape,normalize", [((32, 32, 256, 64), True), ((12, 70, 210, 128), False), ((12, 20, 210, 128), False), ((5, 16, 16, 64), True)])
def test_sharded_nccl_bitwise_equals_single_gpu(world, shape, normalize):
    """BASELINE's 32 units (at N = 8192), Wan's 12 heads (uneven over 8 ranks: padded all-gather; with 70 blocks bitwise,
    with 20 blocks - packed mixing tiles - up to the fp32 summation grouping) and a short-sequence case with fewer units
    than ranks."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _run(world, shape, normalize)


def _worker_blocks(rank, world, port, shape, normalize, rope, q_out):
    import torch.distributed as dist
    import mhla_b200
    from mhla_b200.sharded import mhla_block_sharded, unit_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        B, H, M, w, D = shape
        g = torch.Generator().manual_seed(0)
        mk = lambda relu: ((torch.relu(torch.randn(B, H, M, w, D, generator=g)) + 1e-6) if relu  # noqa: E731
                           else torch.randn(B, H, M, w, D, generator=g)).bfloat16().to(dev)
        q, k, v = mk(True), mk(True), mk(False)
        qr, kr = (mk(False), mk(False)) if rope else (None, None)
        W = (torch.rand(M, M, generator=g) / M).to(dev)
        full = mhla_b200.mhla(q, k, v, W, q_rope=qr, k_rope=kr, normalize=normalize, three_launch=True)
        lo, hi = unit_range(M, world, rank)
        sl = lambda t: None if t is None else t[:, :, lo:hi].contiguous()  # noqa: E731
        out = mhla_block_sharded(sl(q), sl(k), sl(v), W, normalize=normalize, q_rope=sl(qr), k_rope=sl(kr))
        torch.cuda.synchronize()
        q_out.put((rank, bool(torch.equal(out, full[:, :, lo:hi])), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("shape,normalize,rope", [((1, 3, 30, 210, 128), False, True), ((1, 2, 19, 64, 64), True, False)])
def test_block_range_split_nccl_bitwise(world, shape, normalize, rope):
    """Block-range split (one all-gather of the block summaries): each rank's slice of the output is BITWISE the
    single-GPU result - Wan-shaped blocks (210 tokens, D = 128, roped numerator) and an uneven split with the normaliser."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_blocks, args=(r, world, port, shape, normalize, rope, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    results = sorted(q_out.get(timeout=10) for _ in range(world))
    assert all(ok for _, ok, _ in results), results
