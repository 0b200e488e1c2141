#!/usr/bin/env python
"""Headline benchmark: MHLA forward tokens/s at seq_len 32768 (B=2, H=16, D=64) on N x B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path (one JSON line on rank 0)
    python bench.py --impl reference [...]                          # reference CPU path (oracle port) on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N    # N>1: one rank per GPU, (b,h)-units sharded

A "step" is one pass of the operator over one batch of synthetic block-major [B, H, M, w, D] tensors
(w = 256, M = N / 256, W = BlockDistanceConv3D((M,1,1), "linear")), normaliser ON (DiT semantics): ONE launch of the fused
persistent kernel per step (`--three-launch` / `--two-launch` time the multi-launch variants of the same kernel).  Multi-GPU is
weak scaling over independent (b,h) units: every rank processes a full B=2,H=16 batch (global batch 2N), no
data-path collective (`--gather` adds one NCCL all-gather of the outputs for the consumers that need all heads).
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
METRIC = "MHLA fwd tokens/sec at seq_len 32768 (B=2,H=16,D=64)"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(path) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


def make_inputs(device, seed, b=B, h=H, n=N, d=D, w=WBLK, pin=False):
    m = n // w
    g = torch.Generator().manual_seed(seed)
    q = (torch.relu(torch.randn(b, h, m, w, d, generator=g)) + 1e-6).to(torch.bfloat16)
    k = (torch.relu(torch.randn(b, h, m, w, d, generator=g)) + 1e-6).to(torch.bfloat16)
    v = torch.randn(b, h, m, w, d, generator=g).to(torch.bfloat16)
    if pin:
        q, k, v = q.pin_memory(), k.pin_memory(), v.pin_memory()
    return q, k, v


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
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
            time.sleep(0.0005)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_oracle_leg(steps, warmup, units, threads):
    """Reference CPU path (oracle port of the reference's PyTorch-eager code) on `units` (b,h) units of the workload."""
    import oracle
    torch.set_num_threads(threads)
    m = N // WBLK
    q, k, v = make_inputs("cpu", 0, b=1, h=units)
    q, k, v = q.float(), k.float(), v.float()
    W = oracle.block_distance_matrix((m, 1, 1), "linear")
    for _ in range(warmup):
        oracle.blockmix_fwd(q, k, v, W, normalize=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.blockmix_fwd(q, k, v, W, normalize=True)
    dt = (time.perf_counter() - t0) / steps
    # tokens are counted per full-width (H=16) batch element: `units` units = units/H of a sequence of N tokens
    tok_s = (units / H) * N / dt
    return tok_s, dt


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    units = 2
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    tok_s, dt = cpu_oracle_leg(steps, warm, units, threads)
    sample = f"{units} of {B * H} (b,h) units of the headline workload per step (cost is linear in units), fp32, torch-CPU eager"
    line = {
        "impl": "reference", "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"blockmix A: B={B} H={H} N={N} D={D} w={WBLK} M={N // WBLK} normalize=1", "sample": sample},
        "cpu_baseline": {"value": tok_s, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": tok_s, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gather", action="store_true", help="add one NCCL all-gather of the outputs per step (N>1)")
    ap.add_argument("--no-normalize", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--three-launch", action="store_true", help="three PDL-chained phase launches instead of the fused kernel")
    ap.add_argument("--two-launch", action="store_true", help="summaries+mixing kernel followed by the readout kernel")
    ap.add_argument("--fused", action="store_true", help="(default path; kept for old command lines)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import mhla_b200
    import oracle  # checker / cpu_baseline only

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    normalize = not args.no_normalize
    path_kw = {"three_launch": True} if args.three_launch else ({"two_launch": True} if args.two_launch else {})
    warm = max(args.warmup, 3)
    M = N // WBLK

    hq, hk, hv = make_inputs("cpu", 100 + rank, pin=True)
    q, k, v = hq.to(dev), hk.to(dev), hv.to(dev)
    W = oracle.block_distance_matrix((M, 1, 1), "linear").to(dev)
    out = torch.empty_like(q)
    gathered = torch.empty((world,) + tuple(out.shape), dtype=out.dtype, device=dev) if (args.gather and world > 1) else None

    def step():
        mhla_b200.mhla(q, k, v, W, normalize=normalize, out=out, **path_kw)
        if gathered is not None:
            dist.all_gather_into_tensor(gathered, out)

    # quick self-check of one (b,h) unit against the oracle before timing anything
    step()
    torch.cuda.synchronize()
    ref = oracle.blockmix_fwd(hq[0, 0].float(), hk[0, 0].float(), hv[0, 0].float(), W.cpu(), normalize=normalize)
    err = oracle.err_ratio(ref, out[0, 0].float().cpu())
    if not err < 5e-3:
        raise SystemExit(f"bench self-check failed: err_ratio {err}")

    for _ in range(warm):
        step()
    sampler = ClockSampler(local_rank)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    sampler.stop_flag = True
    launches_per_step = mhla_b200.last_launch_count()

    # ---- end to end through the public API with HOST buffers (pinned H2D of q,k,v and D2H of the output per step)
    hout = torch.empty(out.shape, dtype=out.dtype).pin_memory()

    def e2e_step():
        # the public host-tensor entry point: pinned H2D of q,k,v, the kernel and the D2H of the output, pipelined over
        # ranges of (b,h) units on three streams (mhla_b200.ops.mhla_host)
        if path_kw:
            dq, dk, dv = hq.to(dev, non_blocking=True), hk.to(dev, non_blocking=True), hv.to(dev, non_blocking=True)
            hout.copy_(mhla_b200.mhla(dq, dk, dv, W, normalize=normalize, **path_kw), non_blocking=True)
        else:
            mhla_b200.mhla_host(hq, hk, hv, W, out=hout, normalize=normalize)

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

    if dist is not None:
        t = torch.tensor([ms_total, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = t.tolist()
    ms_step = ms_total / args.steps
    tokens_per_step = world * B * N
    value = tokens_per_step / (ms_step * 1e-3)
    e2e_value = tokens_per_step / (ms_e2e / args.e2e_steps * 1e-3)
    in_bytes = 3 * B * H * N * D * 2
    out_bytes = B * H * N * D * 2
    peak, peak_src = measured_peaks()
    achieved = (in_bytes + out_bytes) / (ms_step * 1e-3) / 1e9      # per GPU: algorithmic bytes / step time
    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {
            "workload": f"blockmix A (mhla_dit core): B={B} H={H} N={N} D={D} w={WBLK} M={M} normalize={int(normalize)} per GPU",
            "parallelism": f"(b,h)-sharded x{world}, no collective" + (" + all-gather(out)" if gathered is not None else ""),
            "l2": "inputs+outputs 537 MB per step > 126 MB L2 (no explicit flush)",
        },
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (measured_traffic() if (normalize and not path_kw) else None), "peak_source": peak_src,
            "note": "algorithmic bytes = Q,K,V read + O write = 4*B*H*N*D*2 = 536.9 MB per launch; duration = CUDA-event "
                    "time per step = one launch of blockmix_kernel<64> (the only kernel of a step); traffic = dram read + "
                    "write bytes of that launch from the committed ncu --set full capture (profiles/r01_traffic.json)",
        },
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                "ms_per_step": ms_e2e / args.e2e_steps},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": sampler.summary(),
        "self_check_err_ratio": err,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            tok_s, dt = cpu_oracle_leg(3, 1, 2, threads)
            line["cpu_baseline"] = {"value": tok_s, "unit": "tokens/s", "cores": threads, "kind": "port",
                                    "sample": "2 of 32 (b,h) units per step, 3 steps, fp32 torch-CPU eager oracle"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
