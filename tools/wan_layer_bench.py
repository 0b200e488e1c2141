"""Wan2.1-1.3B MHLA layer (dim 1536, 12 heads, D = 128, 21x30x50 tokens, layout (3,5,10)): the fused inference path
(one pre-processing launch + the blockmix kernel's 3-D block view) against the module's reference-style path (torch
pre-processing + block-major rearrange copies), whole-module forward under bf16 autocast.  GPU box only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhla_b200  # noqa: E402
from mhla_b200.modules import MHLA_Video_Uni  # noqa: E402
from mhla_b200.modules.wan import _rope_tables  # noqa: E402


def rope_freqs(d, n=1024):
    def rp(dim):
        fr = torch.outer(torch.arange(n), 1.0 / torch.pow(10000, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
        return torch.polar(torch.ones_like(fr), fr)
    return torch.cat([rp(d - 4 * (d // 6)), rp(2 * (d // 6)), rp(2 * (d // 6))], dim=1)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def timed_graph(fn, reps=10):
    """Device time per call with the Python / ctypes / allocator cost off the timed path: `reps` calls in ONE CUDA graph."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


dim, heads, layout, grid = 1536, 12, (3, 5, 10), (21, 30, 50)
N = grid[0] * grid[1] * grid[2]
freqs = rope_freqs(dim // heads)
for B, norm in ((1, False), (2, False), (2, True)):
    m = MHLA_Video_Uni(dim, heads, None, 0.0, None, True, layout, normalize_out=norm).cuda().eval()
    x = torch.randn(B, N, dim, device="cuda")
    gs = torch.tensor([list(grid)] * B, dtype=torch.long)
    sl = torch.tensor([N] * B)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        m.fast_path = True
        t_fast = timed(lambda: m(x, sl, gs, freqs))
        y1 = m(x, sl, gs, freqs)
        m.fast_path = False
        t_slow = timed(lambda: m(x, sl, gs, freqs))
        y2 = m(x, sl, gs, freqs)
        # the operator alone on token-major tensors (3-D block view)
        q = torch.randn(B, N, heads, dim // heads, device="cuda").bfloat16()
        W = m.block_attn.conv.weight
        t_op = timed(lambda: mhla_b200.mhla_blockmix_grid(q, q, q, W, grid, layout, normalize=False))
        cos, sin = _rope_tables(grid, freqs, x.device)
        xq = torch.randn(B, N, dim, device="cuda").bfloat16()
        xk = torch.randn(B, N, dim, device="cuda").bfloat16()
        t_prep = timed_graph(lambda: mhla_b200.wan_prep(xq, xk, m.norm_q.weight, m.norm_k.weight, cos, sin, dim // heads, want_plain=norm))
    err = float((y1.float() - y2.float()).norm() / y2.float().norm())
    print(f"B={B} normalize_out={norm}: module forward fused {t_fast:.0f} us vs reference-style {t_slow:.0f} us "
          f"({t_slow / t_fast:.2f}x), rel diff {err:.2e}; operator (3-D view, no norm) {t_op:.0f} us, prep kernel (CUDA-graph device time, distinct q / k inputs) {t_prep:.1f} us")

# gate + LePE (MHLA_Video_Uni(is_gated, is_lepe) = Gated_MHLA_Video_LePE's post-processing): SiLU gate and "+ lepe" inside
# the readout epilogue (ABI v4) against the same fused path with those two as separate elementwise passes
m = MHLA_Video_Uni(dim, heads, None, 0.0, None, True, layout, normalize_out=False, is_gated=True, is_lepe=True).cuda().eval()
x = torch.randn(1, N, dim, device="cuda")
gs, sl = torch.tensor([list(grid)], dtype=torch.long), torch.tensor([N])
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    m.fuse_post = True          # (streaming gate_add launch behind the operator)
    t_f = timed(lambda: m(x, sl, gs, freqs))
    y1 = m(x, sl, gs, freqs)
    m.fuse_post = False
    t_u = timed(lambda: m(x, sl, gs, freqs))
    y2 = m(x, sl, gs, freqs)
    m.fuse_post = True
    m.fast_path = False
    t_ref = timed(lambda: m(x, sl, gs, freqs))
    y3 = m(x, sl, gs, freqs)
    m.fast_path = True
    # the LePE convolution alone: token-major depthwise kernel vs the reference's NCDHW rearrangement + cuDNN depthwise Conv3d
    vv = torch.randn(1, N, dim, device="cuda").bfloat16()
    t_dw = timed_graph(lambda: mhla_b200.dwconv3d_tokens(vv, m.lepe.weight, m.lepe.bias, grid))
    from einops import rearrange as _re
    t_cudnn = timed(lambda: _re(m.lepe(_re(vv, "b (f h w) c -> b c f h w", f=grid[0], h=grid[1], w=grid[2])), "b c f h w -> b (f h w) c"))
print(f"B=1 gated + lepe layer: fused path {t_f:.0f} us vs reference-style path {t_ref:.0f} us ({t_ref / t_f:.2f}x), rel diff "
      f"{float((y1.float() - y3.float()).norm() / y3.float().norm()):.2e}; LePE conv alone: token-major kernel {t_dw:.0f} us vs "
      f"rearrange + cuDNN depthwise Conv3d {t_cudnn:.0f} us")
print(f"B=1 gated + lepe layer: one gate_add launch {t_f:.0f} us vs torch elementwise passes {t_u:.0f} us, rel diff "
      f"{float((y1.float() - y2.float()).norm() / y2.float().norm()):.2e}")

# training step (forward + backward of the whole layer, bf16 autocast) in the shipped configuration: 3-D block view with
# the native backward (autograd.BlockmixGridFunction) against the block-major path (rearrange copies + BlockmixFunction)
m = MHLA_Video_Uni(dim, heads, None, 0.0, None, True, layout, normalize_out=False).cuda().train()


def train_step():
    m.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x, sl, gs, freqs)
    y.float().square().mean().backward()


m.fast_path, m.train_fused_prep = True, True
t_p = timed(train_step, reps=5)
m.fast_path, m.train_fused_prep = True, False
t_g = timed(train_step, reps=5)
m.fast_path = False
t_b = timed(train_step, reps=5)
print(f"B=1 training step of the layer: 3-D block view + fused pre-processing with analytic backward {t_p:.0f} us, 3-D block "
      f"view + torch pre-processing {t_g:.0f} us, block-major copies (reference-style) {t_b:.0f} us")

# the fused gate / "+ lepe" epilogue at OPERATOR level (the layer numbers above are dominated by the depthwise Conv3d):
# one launch with out_gate / out_add against the plain launch followed by the eager elementwise passes
q = torch.randn(1, N, heads, dim // heads, device="cuda").bfloat16()
gt = torch.randn(1, N, heads, dim // heads, device="cuda").bfloat16()
ad = torch.randn(1, N, heads, dim // heads, device="cuda").bfloat16()
W = m.block_attn.conv.weight.detach()
with torch.no_grad():
    t_plain = timed(lambda: mhla_b200.mhla_blockmix_grid(q, q, q, W, grid, layout, normalize=False), reps=20)
    t_fused = timed(lambda: mhla_b200.mhla_blockmix_grid(q, q, q, W, grid, layout, normalize=False, out_gate=gt, out_add=ad), reps=20)
    t_sep = timed(lambda: mhla_b200.mhla_blockmix_grid(q, q, q, W, grid, layout, normalize=False) * torch.nn.functional.silu(gt) + ad, reps=20)

    def streamed():
        o = mhla_b200.mhla_blockmix_grid(q, q, q, W, grid, layout, normalize=False)
        return mhla_b200.gate_add(o, gt, ad, out=o)
    t_stream = timed(streamed, reps=20)
print(f"operator, Wan B=1: plain {t_plain:.0f} us, gate + add fused in the epilogue {t_fused:.0f} us, plain + eager silu/mul/add "
      f"{t_sep:.0f} us, plain + ONE streaming gate_add launch {t_stream:.0f} us")
