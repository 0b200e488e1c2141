#!/usr/bin/env python
"""Headline benchmark: MHLA forward tokens/s at seq_len 32768 (B=2, H=16, D=64) on N x B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path (one JSON line on rank 0)
    python bench.py --impl reference [...]                          # reference CPU path (oracle port) on host cores
    python bench.py --impl reference-gpu [--sweep]                  # context arm: the reference's PyTorch core on the GPU
                                                                    # (eager + torch.compile), fla linear attention, FlashAttention
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N    # N>1: one rank per GPU, (b,h) units sharded

A "step" is one pass of the operator over one batch of synthetic block-major [B, H, M, w, D] tensors
(w = 256, M = N / 256, W = BlockDistanceConv3D((M,1,1), "linear")), normaliser ON (DiT semantics): ONE launch of the
fused persistent kernel per step and rank.  Multi-GPU is STRONG scaling of exactly that batch: the 32 independent
(b,h) units are partitioned contiguously over the ranks (32/16/8/4 units per GPU, `mhla_b200.sharded.mhla_sharded`),
no data-path collective; the same run also times the variant with ONE NCCL all-gather of the outputs
(`all_gather_into_tensor`, for consumers that need every head on every rank) and reports it as `with_gather`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, H, N, D, WBLK = 2, 16, 32768, 64, 256
UNITS = B * H
METRIC = "MHLA fwd tokens/sec at seq_len 32768 (B=2,H=16,D=64)"
L2_BYTES = 126e6


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    for name in ("r02c_traffic.json", "r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return float(json.load(f)["dram_bytes_per_launch"]), name
        except Exception:
            continue
    return None, None


def make_units(lo, hi, n=N, d=D, w=WBLK, pin=False, salt=0):
    """(b,h) units [lo, hi) of the synthetic workload, [units, M, w, D] bf16; unit u always comes from seed 1000+u, so a
    sharded run works on exactly the tensors of the single-GPU run."""
    m = n // w
    qs, ks, vs = [], [], []
    for u in range(lo, hi):
        g = torch.Generator().manual_seed(1000 + u + 100000 * salt)
        qs.append((torch.relu(torch.randn(m, w, d, generator=g)) + 1e-6).to(torch.bfloat16))
        ks.append((torch.relu(torch.randn(m, w, d, generator=g)) + 1e-6).to(torch.bfloat16))
        vs.append(torch.randn(m, w, d, generator=g).to(torch.bfloat16))
    q, k, v = torch.stack(qs), torch.stack(ks), torch.stack(vs)
    if pin:
        q, k, v = q.pin_memory(), k.pin_memory(), v.pin_memory()
    return q, k, v


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.interval = float(os.environ.get("MHLA_BENCH_SAMPLE_S", "0.0005"))
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.interval)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------- reference (CPU)
def cpu_oracle_leg(steps, warmup, threads, n=N):
    """Reference CPU path (oracle port of the reference's PyTorch-eager code) on the WHOLE workload: all 32 (b,h) units."""
    import oracle
    torch.set_num_threads(threads)
    m = n // WBLK
    q, k, v = make_units(0, UNITS, n=n)
    q, k, v = q.float(), k.float(), v.float()
    W = oracle.block_distance_matrix((m, 1, 1), "linear")
    for _ in range(warmup):
        oracle.blockmix_fwd(q, k, v, W, normalize=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.blockmix_fwd(q, k, v, W, normalize=True)
    dt = (time.perf_counter() - t0) / steps
    return B * n / dt, dt


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps, warm = max(1, args.steps), max(1, args.warmup)
    tok_s, dt = cpu_oracle_leg(steps, warm, threads)
    sample = (f"the whole workload: all {UNITS} (b,h) units per step, fp32, torch-CPU eager restatement of "
              "mhla_dit/mhla/mhla.py:262-268 (oracle port; the reference is Python and cannot travel to the GPU box)")
    line = {
        "impl": "reference", "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"blockmix A (mhla_dit core): B={B} H={H} N={N} D={D} w={WBLK} M={N // WBLK} normalize=1",
                   "sample": sample},
        "cpu_baseline": {"value": tok_s, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": tok_s, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- reference (GPU)
def _timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def run_reference_gpu(args):
    """Context arm (NOT the driver's reference arm): what the reference's own GPU path costs on this B200.
    (i)  the reference core mhla_dit/mhla/mhla.py:262-268 - two batched matmuls, a bias-free 1x1 nn.Conv2d over the
         block axis (twice), a divide - restated in torch around a real nn.Conv2d, eager and under torch.compile, in the
         two precisions the reference trains in (fp32 with TF32 as mhla_dit/train.py:12-13 sets; bf16 autocast);
    (ii) fla.ops.linear_attn.chunk_linear_attn (pip fla, Triton; plain causal linear attention, not MHLA) and
    (iii) flash_attn_func at the same B, H, N, D, as context rows.  Ours is timed beside them."""
    import mhla_b200
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    ns = [1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072] if args.sweep else [N]
    rows = []

    def core(q, kt, v, conv, eps):
        kv = conv(torch.matmul(kt, v))                                   # :262-263
        normalizer = conv(torch.matmul(q, kt.sum(dim=-1, keepdim=True))) + eps   # :265-266
        return torch.matmul(q, kv) / normalizer                          # :268

    def compiled_core():
        # a fresh compiled instance per (shape, precision): dynamo otherwise hits its recompile limit over the sweep and
        # silently falls back to eager
        try:
            import torch._dynamo
            torch._dynamo.reset()
            return torch.compile(core, dynamic=False)
        except Exception as e:   # noqa: BLE001
            print(f"torch.compile unavailable: {e}", file=sys.stderr)
            return None

    for n in ns:
        m = n // WBLK
        row = {"N": n, "M": m, "w": WBLK, "B": B, "H": H, "D": D}
        g = torch.Generator(device=dev).manual_seed(0)
        q = torch.relu(torch.randn(UNITS, m, WBLK, D, generator=g, device=dev)) + 1e-6
        k = torch.relu(torch.randn(UNITS, m, WBLK, D, generator=g, device=dev)) + 1e-6
        v = torch.randn(UNITS, m, WBLK, D, generator=g, device=dev)
        Wm = mhla_b200.block_distance_matrix((m, 1, 1), "linear").to(dev) if m > 1 else torch.ones(1, 1, device=dev)
        conv = torch.nn.Conv2d(m, m, 1, bias=False).to(dev)
        conv.weight.data = Wm.view(m, m, 1, 1).clone()
        reps = 20 if n <= 32768 else 5
        qb, kb, vb = q.bfloat16(), k.bfloat16(), v.bfloat16()
        out = torch.empty_like(qb)
        with torch.no_grad():
            row["ours_us"] = _timed(lambda: mhla_b200.mhla(qb, kb, vb, Wm, normalize=True, out=out), reps)
            kt = k.transpose(-2, -1).contiguous()        # the reference materialises k^T in _process_qkv_impl (:236)
            row["ref_eager_tf32_us"] = _timed(lambda: core(q, kt, v, conv, 1e-6), reps)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                row["ref_eager_bf16_autocast_us"] = _timed(lambda: core(q, kt, v, conv, 1e-6), reps)
            try:
                cc = compiled_core()
                if cc is not None:
                    row["ref_compiled_tf32_us"] = _timed(lambda: cc(q, kt, v, conv, 1e-6), reps)
                cc = compiled_core()
                if cc is not None:
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        row["ref_compiled_bf16_autocast_us"] = _timed(lambda: cc(q, kt, v, conv, 1e-6), reps)
            except Exception as e:   # noqa: BLE001
                row["ref_compiled_error"] = repr(e)[:200]
            del kt
            # context rows at the same B, H, N, D (token-major [B, N, H, D])
            qt = qb.view(B, H, n, D).transpose(1, 2).contiguous()
            ktok = kb.view(B, H, n, D).transpose(1, 2).contiguous()
            vt = vb.view(B, H, n, D).transpose(1, 2).contiguous()
            try:
                from fla.ops.linear_attn import chunk_linear_attn
                row["fla_chunk_linear_attn_us"] = _timed(lambda: chunk_linear_attn(qt, ktok, vt, normalize=False), reps)
            except Exception as e:   # noqa: BLE001
                row["fla_error"] = repr(e)[:200]
            try:
                from flash_attn import flash_attn_func
                row["flash_attn_us"] = _timed(lambda: flash_attn_func(qt, ktok, vt, causal=False), 3 if n > 32768 else 5, warm=1)
            except Exception as e:   # noqa: BLE001
                row["flash_attn_error"] = repr(e)[:200]
        best_ref = min(v_ for k_, v_ in row.items() if k_.startswith("ref_") and k_.endswith("_us"))
        row["speedup_vs_best_reference"] = best_ref / row["ours_us"]
        row["ours_tokens_per_s"] = B * n / (row["ours_us"] * 1e-6)
        rows.append(row)
        print(json.dumps(row), file=sys.stderr)
        del q, k, v, qb, kb, vb, out, qt, ktok, vt
        torch.cuda.empty_cache()
    head = next(r for r in rows if r["N"] == N) if any(r["N"] == N for r in rows) else rows[-1]
    line = {"impl": "reference-gpu", "metric": METRIC, "unit": "us per forward (lower is better)", "n_gpus": 1,
            "config": {"workload": f"B={B} H={H} D={D} w={WBLK}, N swept" if args.sweep else f"B={B} H={H} N={N} D={D} w={WBLK}"},
            "headline": head, "rows": rows}
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "r02_comparators.json"), "w") as f:
            json.dump(line, f, indent=1)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- ours
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--sweep", action="store_true", help="reference-gpu: sweep N = 1k .. 128k")
    ap.add_argument("--no-gather", action="store_true", help="N>1: skip the extra timed loop with the all-gather of the outputs")
    ap.add_argument("--no-normalize", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-chunks", type=int, default=8, help="ranges of (b,h) units the host entry point pipelines over")
    ap.add_argument("--three-launch", action="store_true", help="three PDL-chained phase launches instead of the fused kernel")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.impl == "reference-gpu":
        if rank == 0:
            run_reference_gpu(args)
        return

    import mhla_b200
    from mhla_b200.sharded import mhla_sharded, unit_range
    import oracle  # checker / cpu_baseline only

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    normalize = not args.no_normalize
    path_kw = {"three_launch": True} if args.three_launch else {}
    warm = max(args.warmup, 3)
    M = N // WBLK
    lo, hi = unit_range(UNITS, world, rank)
    nloc = hi - lo

    # Rotating input sets keep the per-rank working set above L2 (126 MB): one set is 4 * nloc * N * D * 2 bytes.
    set_bytes = 4 * nloc * N * D * 2
    nsets = max(1, -(-int(3 * L2_BYTES) // set_bytes)) if set_bytes < 3 * L2_BYTES else 1
    hq, hk, hv = make_units(lo, hi, pin=True)
    sets = [(hq.to(dev), hk.to(dev), hv.to(dev), torch.empty((nloc, M, WBLK, D), dtype=torch.bfloat16, device=dev))]
    for s_ in range(1, nsets):
        g = torch.Generator(device=dev).manual_seed(7 + s_)
        mk = lambda relu: ((torch.relu(torch.randn(nloc, M, WBLK, D, generator=g, device=dev)) + 1e-6) if relu  # noqa: E731
                           else torch.randn(nloc, M, WBLK, D, generator=g, device=dev)).bfloat16()
        sets.append((mk(True), mk(True), mk(False), torch.empty((nloc, M, WBLK, D), dtype=torch.bfloat16, device=dev)))
    W = mhla_b200.block_distance_matrix((M, 1, 1), "linear").to(dev)
    gathered = torch.empty((UNITS, M, WBLK, D), dtype=torch.bfloat16, device=dev) if world > 1 else None
    it = [0]

    def launch(i, gather=False):
        q, k, v, out = sets[i % nsets]
        mhla_sharded(q, k, v, W, inputs="local", total_units=UNITS, gather=False, normalize=normalize, out=out, **path_kw)
        if gather:
            dist.all_gather_into_tensor(gathered, out)

    # A step = one launch of the operator on this rank's units.  The K timed steps are captured into ONE CUDA graph (after
    # warm-up) and replayed: the launches keep their programmatic-dependent-launch edges inside the graph, so consecutive
    # steps overlap prologue and tail exactly like eager back-to-back launches, but the Python / ctypes enqueue cost
    # (~25 us per call) is off the timed path - it matters from 4 ranks on, where the per-rank kernel is that short.
    # `--no-graph` times eager calls.
    graph = [None]
    it = [0]

    def step(gather=False):
        i = it[0]
        it[0] += 1
        launch(i, gather)

    # quick self-check of one (b,h) unit against the oracle before timing anything
    step()
    torch.cuda.synchronize()
    ref = oracle.blockmix_fwd(hq[0][None].float(), hk[0][None].float(), hv[0][None].float(), W.cpu(), normalize=normalize)
    err = oracle.err_ratio(ref[0], sets[0][3][0].float().cpu())
    if not err < 5e-3:
        raise SystemExit(f"bench self-check failed: err_ratio {err}")
    it[0] = 0
    if not args.no_graph:
        try:
            for i in range(nsets):
                launch(i)                                  # warm every set's descriptors / workspace
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for i in range(args.steps):
                    launch(i)
            graph[0] = gr
        except Exception as e:   # noqa: BLE001
            print(f"CUDA graph capture failed ({e!r}); timing eager launches", file=sys.stderr)
            graph[0] = None

    def timed_loop(gather):
        for _ in range(warm):
            step(gather)
        if graph[0] is not None and not gather:
            graph[0].replay()                              # untimed: the first replay of a graph also uploads it to the device
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        if graph[0] is not None and not gather:
            graph[0].replay()                              # = exactly args.steps launches
        else:
            for _ in range(args.steps):
                step(gather)
        e1.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total = timed_loop(False)
    sampler.stop_flag = True
    launches_per_step = mhla_b200.last_launch_count()
    ms_gather = timed_loop(True) if (world > 1 and not args.no_gather) else None

    # ---- end to end through the public API with HOST buffers (pinned H2D of q,k,v and D2H of the output per step)
    hout = torch.empty((nloc, M, WBLK, D), dtype=torch.bfloat16).pin_memory()
    h5 = lambda t: t.view(nloc, 1, M, WBLK, D)   # noqa: E731

    def e2e_step():
        # the public host-tensor entry point: pinned H2D of q,k,v, the kernel and the D2H of the output, pipelined over
        # ranges of (b,h) units on three streams (mhla_b200.ops.mhla_host)
        mhla_b200.mhla_host(h5(hq), h5(hk), h5(hv), W, out=h5(hout), normalize=normalize, chunks=min(args.e2e_chunks, nloc))

    e2e_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.e2e_steps):
        e2e_step()
    f1.record()
    torch.cuda.synchronize()
    ms_e2e = f0.elapsed_time(f1)
    e2e_ok = bool(torch.equal(hout, sets[0][3].cpu())) if nsets == 1 else None

    if dist is not None:
        t = torch.tensor([ms_total, ms_e2e, ms_gather or 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e, mg = t.tolist()
        ms_gather = mg if ms_gather is not None else None
    ms_step = ms_total / args.steps
    tokens_per_step = B * N                       # strong scaling: the batch is fixed, the ranks share it
    value = tokens_per_step / (ms_step * 1e-3)
    e2e_value = tokens_per_step / (ms_e2e / args.e2e_steps * 1e-3)
    unit_bytes = N * D * 2
    in_bytes, out_bytes = 3 * nloc * unit_bytes, nloc * unit_bytes          # this rank's share
    peak, peak_src = measured_peaks()
    alg_bytes = 4 * nloc * unit_bytes
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9      # per GPU: algorithmic bytes of its launch / step time
    traffic, traffic_src = measured_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {
            "workload": f"blockmix A (mhla_dit core): B={B} H={H} N={N} D={D} w={WBLK} M={M} normalize={int(normalize)} (whole job)",
            "parallelism": (f"{UNITS} (b,h) units sharded over {world} rank(s), {nloc} per GPU, no data-path collective"
                            + ("; with_gather adds one NCCL all_gather_into_tensor of the outputs behind the kernel" if ms_gather else "")),
            "launch": (f"one CUDA graph of the {args.steps} timed launches (PDL edges kept), one untimed warm-up replay, then replayed once for the timing" if graph[0] is not None
                       else "eager launch per step"),
            "l2": (f"{nsets} rotating input/output set(s) of {set_bytes / 1e6:.0f} MB per rank: working set "
                   f"{nsets * set_bytes / 1e6:.0f} MB > 126 MB L2 (no explicit flush)"),
        },
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (traffic if (normalize and not path_kw and world == 1) else None), "peak_source": peak_src,
            "note": f"algorithmic bytes = Q,K,V read + O write = 4*units*N*D*2 = {alg_bytes / 1e6:.1f} MB per launch on this "
                    "GPU; duration = CUDA-event time per step = one launch of blockmix_kernel<64> (the only kernel of a "
                    f"step); traffic = dram read + write bytes of that launch from the committed ncu --set full capture "
                    f"(profiles/{traffic_src})",
        },
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": in_bytes * world, "d2h_bytes_per_step": out_bytes * world,
                "ms_per_step": ms_e2e / args.e2e_steps, "bit_identical_to_device_path": e2e_ok},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": sampler.summary(),
        "self_check_err_ratio": err,
    }
    if ms_gather is not None:
        mg = ms_gather / args.steps
        line["with_gather"] = {"value": tokens_per_step / (mg * 1e-3), "unit": "tokens/s", "ms_per_step": mg,
                               "all_gather_bytes": UNITS * unit_bytes}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            tok_s, dt = cpu_oracle_leg(3, 1, threads)
            line["cpu_baseline"] = {"value": tok_s, "unit": "tokens/s", "cores": threads, "kind": "port",
                                    "sample": f"the whole workload (all {UNITS} (b,h) units) per step, 3 steps, fp32 torch-CPU eager oracle"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
