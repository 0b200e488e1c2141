"""Training-step timing of the block-mixed operator on one B200: forward kernel, native backward (three more launches
of the forward kernel, mhla_b200/autograd.py) and - as the yardstick - the plain-torch statement of the same gradients
(fp32 cuBLAS einsums) and torch.autograd through the reference's own formulation (bf16 autocast, nn.Conv2d mixing).
    python tools/bwd_bench.py [--wan]"""
import json
import sys

import torch

sys.path.insert(0, ".")
import mhla_b200  # noqa: E402
from mhla_b200 import autograd, ops  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def reference_step(q, k, v, conv, do, normalize):
    """mhla_dit/mhla/mhla.py:262-268 under bf16 autocast, forward + backward through torch.autograd."""
    q, k, v = (t.detach().requires_grad_(True) for t in (q, k, v))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        kv = torch.matmul(k.transpose(-2, -1), v)
        kv = conv(kv)
        out = torch.matmul(q, kv)
        if normalize:
            ksum = k.sum(dim=-2, keepdim=True).transpose(-2, -1)
            out = out / (conv(torch.matmul(q, ksum)) + 1e-6)
    out.backward(do)


def main():
    wan = "--wan" in sys.argv
    shapes = [(1, 12, 150, 210, 128, False)] if wan else [(2, 16, 128, 256, 64, True), (2, 16, 128, 256, 64, False)]
    g = torch.Generator(device="cuda").manual_seed(0)
    for B, H, M, w, D, normalize in shapes:
        mk = lambda relu: ((torch.relu(torch.randn(B * H, M, w, D, generator=g, device="cuda")) + 1e-6) if relu  # noqa: E731
                           else torch.randn(B * H, M, w, D, generator=g, device="cuda")).bfloat16()
        q, k, v, do = mk(True), mk(True), mk(False), mk(False)
        W = (torch.rand(M, M, device="cuda", generator=g) / M + 0.3 * torch.eye(M, device="cuda"))
        out = ops._blockmix_fwd(q, k, v, W, normalize=normalize)
        t_f = timeit(lambda: ops._blockmix_fwd(q, k, v, W, normalize=normalize))
        qg, kg, vg = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        Wg = W.detach().clone().requires_grad_(True)

        def step(with_w=True):   # what a trainer runs: autograd.BlockmixFunction forward + backward
            qg.grad = kg.grad = vg.grad = Wg.grad = None      # (no accumulation passes: the reference arm uses fresh leaves)
            o = mhla_b200.mhla_blockmix(qg, kg, vg, Wg if with_w else W, normalize=normalize)
            o.backward(do)
        t_step = timeit(step)
        t_step_now = timeit(lambda: step(False))
        t_bn, t_bn_now = t_step - t_f, t_step_now - t_f
        t_bt = timeit(lambda: autograd.blockmix_backward(q, k, v, W, do, normalize=normalize), iters=3, warm=1)
        conv = torch.nn.Conv2d(M, M, 1, bias=False).cuda()
        with torch.no_grad():
            conv.weight.copy_(W.view(M, M, 1, 1))
        t_ref = timeit(lambda: reference_step(q, k, v, conv, do, normalize), iters=3, warm=1)
        a = autograd.blockmix_backward_native(q, k, v, W, do, out, normalize=normalize)
        b = autograd.blockmix_backward(q, k, v, W, do, normalize=normalize)
        rel = lambda x, y: float((x.float() - y.float()).norm() / y.float().norm())  # noqa: E731
        print(json.dumps({
            "shape": dict(B=B, H=H, M=M, w=w, D=D, normalize=normalize),
            "fwd_us": round(t_f, 1), "bwd_native_us": round(t_bn, 1), "bwd_native_no_dW_us": round(t_bn_now, 1),
            "bwd_torch_fp32_us": round(t_bt, 1), "reference_autocast_fwd_bwd_us": round(t_ref, 1),
            "step_native_us": round(t_step, 1), "step_native_no_dW_us": round(t_step_now, 1),
            "rel_diff_native_vs_torch": {n: rel(x, y) for n, x, y in zip(("dq", "dk", "dv", "dW"), a, b)},
        }), flush=True)


if __name__ == "__main__":
    main()
