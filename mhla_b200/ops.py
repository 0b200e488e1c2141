"""Functional API of the B200-native MHLA forward operator.

``mhla(q, k, v, mix, ...)`` is the block-mixed operator the reference writes inline
(mhla_dit/mhla/mhla.py:262-268; mhla_videogen/diffusion/model/wan/mhla_utils.py:328-341);
``naive_chunk_simple_mhla_fixed`` / ``naive_recurrent_mhla`` keep the reference's causal entry points
(mhla_nlp/fla/ops/mhla/naive.py:10-83, :88-142) name-for-name so ``fla.layers.mhla`` can import them unchanged.

Everything here is a thin host shim: argument checking, output/workspace allocation from the torch caching
allocator, and one C-ABI call (``libmhla_b200.so``) on the current CUDA stream.  There is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _capi

__all__ = ["mhla", "mhla_blockmix", "mhla_blockmix_grid", "wan_prep", "gated_rmsnorm", "gate_add", "dwconv3d_tokens", "mhla_host", "mhla_causal", "naive_chunk_simple_mhla_fixed", "naive_recurrent_mhla",
           "last_launch_count"]

_DT = {torch.bfloat16: _capi.MHLA_BF16, torch.float16: _capi.MHLA_FP16}


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mhla_b200 operators run on CUDA (sm_100a) tensors only; there is no CPU fallback")


def _as5(t: torch.Tensor) -> torch.Tensor:
    """View as [B, H, M, w, D]; the reference's '(b h) n w d' 4-D layout becomes H = 1."""
    if t.dim() == 4:
        return t.unsqueeze(1)
    if t.dim() != 5:
        raise ValueError(f"expected a [B,H,M,w,D] or [(B H),M,w,D] tensor, got shape {tuple(t.shape)}")
    return t


def _tma_ok(t: torch.Tensor) -> bool:
    if t.stride(-1) != 1 or t.data_ptr() % 16:
        return False
    return all(s % 8 == 0 for s, n in zip(t.stride()[:-1], t.shape[:-1]) if n > 1)


def _prep(t: Optional[torch.Tensor], dtype: torch.dtype) -> Optional[torch.Tensor]:
    if t is None:
        return None
    t = _as5(t)
    if t.dtype != dtype:
        t = t.to(dtype)
    if not _tma_ok(t):
        t = t.contiguous()
    return t


# Persistent workspaces of the fused path, one per (device, stream, size): zeroed once, then self-cleaning (the kernel's
# last CTA re-zeroes the control block), so a call is a single kernel launch.  A handful of shapes are kept alive.
_WS_CACHE: "dict[tuple, torch.Tensor]" = {}
_WS_CACHE_MAX = 4


class _Local(__import__("threading").local):
    """Per-thread descriptor cache: a cached ctypes descriptor is mutated (pointers) on every call."""

    def __init__(self):
        self.desc = {}


_TLS = _Local()


def _new_ws(device: torch.device, nbytes: int, off_cnt: int) -> torch.Tensor:
    # only the control block (dependency counters, tickets, flags) has to start out zero - the kernel keeps it
    # that way; zeroing just that keeps a CUDA-graph capture of the first call from recording a memset of the
    # whole (tens of MB) workspace that every replay would repeat
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    base = (ws.data_ptr() + 1023) // 1024 * 1024 - ws.data_ptr()
    ws[base + off_cnt:].zero_()
    return ws


def _persistent_ws(device: torch.device, nbytes: int, stream_handle: int, sig: tuple, off_cnt: int) -> torch.Tensor:
    # keyed by the full call signature: two shapes of equal total size lay the control block out differently
    key = (device.index, stream_handle, nbytes) + tuple(sig)
    ws = _WS_CACHE.get(key)
    if ws is None:
        if len(_WS_CACHE) >= _WS_CACHE_MAX:
            _WS_CACHE.pop(next(iter(_WS_CACHE)))
        ws = _WS_CACHE[key] = _new_ws(device, nbytes, off_cnt)
    return ws


def _ws_views(ws: torch.Tensor, lay, G: int, M: int, D: int) -> dict:
    """Typed views of a workspace (mhla_blockmix_workspace_layout): block summaries S [G, M, ncols] (D*D summaries, then
    n_loc hi | lo), mixed summaries St [G, M, D*D] (16-bit, viewed as int16) and den [G, M, 2*wpad] fp32."""
    base = (ws.data_ptr() + 1023) // 1024 * 1024 - ws.data_ptr()
    ncols, wpad = int(lay[5]), int(lay[6])
    return dict(
        S=ws[base + lay[0]: base + lay[0] + G * M * ncols * 2].view(torch.int16).view(G, M, ncols),
        St=ws[base + lay[1]: base + lay[1] + G * M * D * D * 2].view(torch.int16).view(G, M, D * D),
        den=(ws[base + lay[2]: base + lay[2] + G * M * 2 * wpad * 4].view(torch.float32).view(G, M, 2 * wpad) if wpad else None),
        ncols=ncols, wpad=wpad)


def _t5(t: Optional[torch.Tensor]) -> _capi.Tensor5:
    if t is None:
        return _capi.Tensor5(None, 0, 0, 0, 0)
    sb, sh, sm, sw, _ = t.stride()
    return _capi.Tensor5(t.data_ptr(), sb, sh, sm, sw)


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def mhla_blockmix(q, k, v, mix, *, q_rope=None, k_rope=None, eps: float = 1e-6, normalize: bool = True,
                  out: Optional[torch.Tensor] = None, out_rms_weight: Optional[torch.Tensor] = None,
                  out_gate: Optional[torch.Tensor] = None, out_add: Optional[torch.Tensor] = None, **kw) -> torch.Tensor:
    """Non-causal block-mixed MHLA (see ``_blockmix_fwd`` for the arguments).  Differentiable: when gradients are
    enabled and an input requires them the call goes through ``autograd.BlockmixFunction`` (CUDA forward, analytic
    backward for q, k, v, q_rope, k_rope and the mixing matrix); ``out=`` and the fused output RMSNorm are
    inference-only and raise instead of silently dropping the graph."""
    if _needs_grad(q, k, v, mix, q_rope, k_rope, out_rms_weight, out_gate, out_add):
        if out is not None or out_rms_weight is not None or out_gate is not None or out_add is not None:
            raise NotImplementedError("out= and the fused epilogue (out_rms_weight / out_gate / out_add) are inference-only; "
                                      "call without them when training")
        from .autograd import BlockmixFunction
        return BlockmixFunction.apply(q, k, v, mix, q_rope, k_rope, float(eps), bool(normalize), kw)
    return _blockmix_fwd(q, k, v, mix, q_rope=q_rope, k_rope=k_rope, eps=eps, normalize=normalize, out=out,
                         out_rms_weight=out_rms_weight, out_gate=out_gate, out_add=out_add, **kw)


def _blockmix_fwd(q, k, v, mix, *, q_rope=None, k_rope=None, eps: float = 1e-6, normalize: bool = True,
                  out: Optional[torch.Tensor] = None, fused: bool = True, unfused: Optional[bool] = None,
                  three_launch: bool = False, two_launch: bool = False, force_fused: bool = False, debug_flags: int = 0,
                  out_rms_weight: Optional[torch.Tensor] = None, out_rms_eps: float = 1e-6,
                  no_smalln: bool = False, ws_out: Optional[dict] = None, out_gate: Optional[torch.Tensor] = None,
                  out_add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Non-causal block-mixed MHLA forward on block-major tensors (no autograd graph).

    q, k, v : [B, H, M, w, D] (or the reference's [(B H), M, w, D]); bf16 / fp16 (fp32 is computed in bf16
              with fp32 accumulation and returned as fp32).  Strided views are consumed in place when every
              stride is a multiple of 8 elements.
    mix     : [M, M] (or the Conv2d weight [M, M, 1, 1]); out_i = sum_j mix[i, j] (.)_j.
    q_rope, k_rope : roped copies for the numerator (variant B); q, k then only feed the normaliser.
    out_rms_weight : optional [D] weight of a per-(token, head) RMSNorm fused into the readout epilogue
              (out = o * rsqrt(mean_d(o^2) + out_rms_eps) * weight; MHLA_Video_Uni's g_norm, mhla_utils.py:360-362).
    out_gate, out_add : optional tensors shaped like the output (any TMA-compatible strides, e.g. permuted views of a
              [B, M, w, H*D] projection): out = o * silu(out_gate) + out_add, fused into the readout epilogue in fp32 before
              the single rounding (the SiLU gate and "+ lepe" of the Wan classes, the "+ lepe" of MHLA4DiT; ABI v4).
    no_smalln : units of at most 256 tokens (M*w <= 256, D = 64, M <= 64: DiT / ViT) normally take the short-sequence
              kernel (whole unit on chip, no workspace, csrc/smalln_kernel.cuh); True forces the general kernel.
    fused   : (default) one persistent kernel: items are scheduled at run time, cross-CTA dependencies go through
              per-group counters.  ``three_launch`` / ``unfused=True`` run the phases as three PDL-chained launches of
              the same kernel (an independent cross-check), ``two_launch`` as summaries+mixing followed by the readout.
    ws_out  : optional dict; the call then runs on a PRIVATE workspace and fills the dict with typed views of it
              (``_ws_views``: the block summaries S_j = k_j^T v_j the kernel produced) - the backward pass reads the
              summaries of its own launches from there (autograd.py).  Left empty when the shape takes no workspace.
    """
    _require_cuda(q, k, v, mix, q_rope, k_rope)
    q, k, v = q.detach(), k.detach(), v.detach()
    q_rope = None if q_rope is None else q_rope.detach()
    k_rope = None if k_rope is None else k_rope.detach()
    if (q_rope is None) != (k_rope is None):
        raise ValueError("q_rope and k_rope must be given together")
    in_dtype = q.dtype
    cdtype = in_dtype if in_dtype in _DT else torch.bfloat16
    squeeze = q.dim() == 4
    q5, k5, v5 = _prep(q, cdtype), _prep(k, cdtype), _prep(v, cdtype)
    qr5, kr5 = _prep(q_rope, cdtype), _prep(k_rope, cdtype)
    B, H, M, w, D = q5.shape
    for name, t in (("k", k5), ("v", v5), ("q_rope", qr5), ("k_rope", kr5)):
        if t is not None and tuple(t.shape) != (B, H, M, w, D):
            raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {(B, H, M, w, D)}")
    D_in = D
    if D not in (64, 128) and D < 128:
        # head dims between the kernel's two widths (DiT-XL: 1152 / 16 = 72) run zero-padded: padded channels add exact
        # zeros to K^T V, to the k-sums and to Q.S~, so the result is unchanged; the output is sliced back
        if out is not None:
            raise ValueError("out= is only supported for head dims 64 and 128")
        D = 64 if D < 64 else 128
        padc = lambda t: None if t is None else torch.nn.functional.pad(t, (0, D - D_in))  # noqa: E731
        q5, k5, v5, qr5, kr5 = padc(q5), padc(k5), padc(v5), padc(qr5), padc(kr5)
        if out_rms_weight is not None or out_gate is not None or out_add is not None:
            raise ValueError("the fused epilogue (out_rms_weight / out_gate / out_add) needs a head dim of 64 or 128")
    mix2 = mix.reshape(mix.shape[0], mix.shape[1]) if mix.dim() != 2 else mix
    if tuple(mix2.shape) != (M, M):
        raise ValueError(f"mix must be [{M}, {M}], got {tuple(mix.shape)}")
    if mix2.dtype != torch.float32 or mix2.stride(1) != 1 or mix2.requires_grad:
        mix2 = mix2.detach().to(torch.float32).contiguous()
    if out is None:
        o5 = torch.empty((B, H, M, w, D), dtype=cdtype, device=q.device)
    else:
        o5 = _as5(out)
        if o5.dtype != cdtype or not _tma_ok(o5) or tuple(o5.shape) != (B, H, M, w, D):
            raise ValueError("out must be a TMA-compatible tensor of the compute dtype and the shape of q")

    if unfused is not None:          # spelling used by the tests: unfused=True -> three launches
        three_launch = bool(unfused)
    if two_launch or not fused:      # (the two-launch variant of round 1 is gone: it now means phase-by-phase launches)
        three_launch = True
    flags = ((_capi.FLAG_NORMALIZE if normalize else 0) | (_capi.FLAG_UNFUSED if three_launch else 0) | int(debug_flags) |
             (_capi.FLAG_NO_SMALLN if no_smalln else 0))
    single = not (three_launch or debug_flags)
    if single:
        flags |= _capi.FLAG_WS_PERSISTENT | (_capi.FLAG_FUSED if force_fused else 0)
    L = _capi.lib()
    # the descriptor (shape, flags, workspace size) is cached per call signature; only pointers and strides change
    post = []
    for name, t in (("out_gate", out_gate), ("out_add", out_add)):
        if t is None:
            post.append(None)
            continue
        t5 = _as5(t.detach())
        if tuple(t5.shape) != (B, H, M, w, D):
            raise ValueError(f"{name} has shape {tuple(t.shape)}, expected the output's {(B, H, M, w, D)}")
        if t5.dtype != cdtype:
            t5 = t5.to(cdtype)
        post.append(t5 if _tma_ok(t5) else t5.contiguous())
    g5, a5 = post
    sig = (B, H, M, w, D, cdtype, flags, float(eps), qr5 is not None, out_rms_weight is not None, g5 is not None,
           a5 is not None)
    cache = _TLS.desc
    ent = cache.get(sig)
    if ent is None:
        d = _capi.BlockmixDesc()
        d.B, d.H, d.M, d.w, d.D = B, H, M, w, D
        d.dtype = _DT[cdtype]
        d.flags = flags
        d.eps = float(eps)
        nbytes = L.mhla_blockmix_workspace_bytes(C.byref(d))
        if nbytes == 0:
            raise _capi.MhlaError(
                f"unsupported blockmix shape B={B} H={H} M={M} w={w} D={D} (need D in {{64,128}}, w<=256)")
        # short sequences (no rope, no fused output norm) run without a workspace (csrc/smalln_kernel.cuh)
        d.q_rope.ptr, d.k_rope.ptr = (1 if qr5 is not None else None), (1 if kr5 is not None else None)
        d.out_rms_weight = 1 if out_rms_weight is not None else None
        d.out_gate.ptr, d.out_add.ptr = (1 if g5 is not None else None), (1 if a5 is not None else None)
        needs_ws = bool(L.mhla_blockmix_needs_workspace(C.byref(d))) if hasattr(L, "mhla_blockmix_needs_workspace") else True
        lay = (C.c_size_t * 8)()
        _capi.check(L.mhla_blockmix_workspace_layout(C.byref(d), C.byref(lay)), "mhla_blockmix_workspace_layout")
        if len(cache) > 64:
            cache.clear()
        ent = cache[sig] = (d, nbytes, needs_ws, tuple(int(x) for x in lay))
    d, nbytes, needs_ws, lay = ent
    off_cnt = lay[4]
    d.q, d.k, d.v, d.out = _t5(q5), _t5(k5), _t5(v5), _t5(o5)
    d.q_rope, d.k_rope = _t5(qr5), _t5(kr5)
    d.out_gate, d.out_add = _t5(g5), _t5(a5)
    d.mix, d.mix_ld = mix2.data_ptr(), mix2.stride(0)
    rms_w = None
    if out_rms_weight is not None:
        rms_w = out_rms_weight.detach().to(device=q.device, dtype=torch.float32).contiguous()
        if rms_w.numel() != D:
            raise ValueError(f"out_rms_weight must have {D} elements")
    d.out_rms_weight, d.out_rms_eps = (rms_w.data_ptr() if rms_w is not None else None), float(out_rms_eps)
    stream = torch.cuda.current_stream(q.device)
    if not needs_ws:
        d.workspace, d.workspace_bytes = None, 0
    else:
        if single and ws_out is None:
            ws = _persistent_ws(q.device, nbytes, stream.cuda_stream, sig, off_cnt)
        elif single:
            ws = _new_ws(q.device, nbytes, off_cnt)
        else:
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=q.device)
        d.workspace, d.workspace_bytes = (ws.data_ptr() + 1023) // 1024 * 1024, nbytes
        if ws_out is not None:
            ws_out.update(_ws_views(ws, lay, B * H, M, D), ws=ws, D=D)
    if torch.cuda.current_device() == q.device.index:
        _capi.check(L.mhla_fwd_blockmix(C.byref(d), stream.cuda_stream), "mhla_fwd_blockmix")
    else:
        with torch.cuda.device(q.device):
            _capi.check(L.mhla_fwd_blockmix(C.byref(d), stream.cuda_stream), "mhla_fwd_blockmix")
    # (operands and temporaries belong to the current stream, like those of any other torch op: the caching allocator's
    #  stream-ordered reuse keeps them valid for the enqueued kernel - no record_stream traffic per call)
    res = o5 if out is None else out
    if out is None:
        if D_in != D:
            res = res[..., :D_in]
        if squeeze:
            res = res.squeeze(1)
        if in_dtype != cdtype:
            res = res.to(in_dtype)
    return res


def mhla_blockmix_grid(q, k, v, mix, grid, layout, *, q_rope=None, k_rope=None, eps: float = 1e-6, normalize: bool = True,
                       out: Optional[torch.Tensor] = None, out_rms_weight: Optional[torch.Tensor] = None,
                       out_rms_eps: float = 1e-6, three_launch: bool = False, ws_out: Optional[dict] = None,
                       out_gate: Optional[torch.Tensor] = None, out_add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Block-mixed MHLA forward on TOKEN-major tensors with a 3-D block structure (variant B, Wan): q, k, v (and the
    roped copies) are [B, F*H*W, heads, D] - e.g. plain views of the q/k/v projections - and ``layout = (fb, hb, wb)``
    cuts the ``grid = (F, H, W)`` token grid into M = fb*hb*wb blocks of (F/fb)*(H/hb)*(W/wb) tokens, exactly the
    ``"b (fb p1 hb p2 wb p3) h c -> (b h) (fb hb wb) (p1 p2 p3) c"`` rearrangement of mhla_utils.py:317-326.  The kernel
    gathers the blocks with TMA boxes and scatters the output back (:345-354), so no block-major copy is ever made.
    Returns [B, F*H*W, heads, D].  Differentiable without the normaliser (``autograd.BlockmixGridFunction``: the shipped
    Wan configuration); with it, training goes through the block-major ``mhla_blockmix``."""
    if _needs_grad(q, k, v, mix, q_rope, k_rope, out_rms_weight, out_gate, out_add):
        if (normalize or q_rope is not None or out is not None or out_rms_weight is not None or out_gate is not None
                or out_add is not None):
            raise NotImplementedError("the 3-D block view is differentiable for normalize=False without out= / fused "
                                      "output norm (pass the roped q, k as q, k); use mhla_blockmix otherwise")
        from .autograd import BlockmixGridFunction
        return BlockmixGridFunction.apply(q, k, v, mix, tuple(grid), tuple(layout), float(eps))
    return _blockmix_grid_fwd(q, k, v, mix, grid, layout, q_rope=q_rope, k_rope=k_rope, eps=eps, normalize=normalize, out=out,
                              out_rms_weight=out_rms_weight, out_rms_eps=out_rms_eps, three_launch=three_launch, ws_out=ws_out,
                              out_gate=out_gate, out_add=out_add)


def _blockmix_grid_fwd(q, k, v, mix, grid, layout, *, q_rope=None, k_rope=None, eps: float = 1e-6, normalize: bool = True,
                       out: Optional[torch.Tensor] = None, out_rms_weight: Optional[torch.Tensor] = None,
                       out_rms_eps: float = 1e-6, three_launch: bool = False, ws_out: Optional[dict] = None,
                       out_gate: Optional[torch.Tensor] = None, out_add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``mhla_blockmix_grid`` without the autograd dispatch (``ws_out``: as in ``_blockmix_fwd``)."""
    _require_cuda(q, k, v, mix, q_rope, k_rope)
    if (q_rope is None) != (k_rope is None):
        raise ValueError("q_rope and k_rope must be given together")
    F_, H_, W_ = (int(x) for x in grid)
    fb, hb, wb = (int(x) for x in layout)
    B, N, nh, D = q.shape
    if N != F_ * H_ * W_:
        raise ValueError(f"q has {N} tokens, grid {grid} has {F_ * H_ * W_}")
    if D not in (64, 128):
        raise ValueError("the 3-D block view needs a head dim of 64 or 128")
    M, w = fb * hb * wb, (F_ // fb) * (H_ // hb) * (W_ // wb)
    cdtype = q.dtype if q.dtype in _DT else torch.bfloat16

    def prep(t):
        if t is None:
            return None
        t = t.detach()
        if tuple(t.shape) != (B, N, nh, D):
            raise ValueError(f"expected a [{B}, {N}, {nh}, {D}] tensor, got {tuple(t.shape)}")
        if t.dtype != cdtype:
            t = t.to(cdtype)
        ok = (t.stride(-1) == 1 and t.data_ptr() % 16 == 0 and t.stride(1) % 8 == 0 and t.stride(2) % 8 == 0 and
              (B == 1 or t.stride(0) == N * t.stride(1)))
        return t if ok else t.contiguous()

    q4, k4, v4, qr4, kr4 = prep(q), prep(k), prep(v), prep(q_rope), prep(k_rope)
    g4, a4 = prep(out_gate), prep(out_add)     # [B, N, heads, D] views of e.g. the gate projection / the LePE conv output
    o4 = torch.empty((B, N, nh, D), dtype=cdtype, device=q.device) if out is None else out
    if out is not None and (o4.dtype != cdtype or tuple(o4.shape) != (B, N, nh, D) or prep(o4) is not o4):
        raise ValueError("out must be a [B, N, heads, D] tensor of the compute dtype with TMA-compatible strides")
    mix2 = mix.detach().reshape(mix.shape[0], mix.shape[1]).to(torch.float32).contiguous()
    if tuple(mix2.shape) != (M, M):
        raise ValueError(f"mix must be [{M}, {M}], got {tuple(mix.shape)}")
    t5 = lambda t: _capi.Tensor5(None, 0, 0, 0, 0) if t is None else _capi.Tensor5(t.data_ptr(), t.stride(0), t.stride(2), 0, t.stride(1))  # noqa: E731
    flags = (_capi.FLAG_NORMALIZE if normalize else 0) | (_capi.FLAG_UNFUSED if three_launch else _capi.FLAG_WS_PERSISTENT)
    L = _capi.lib()
    d = _capi.BlockmixDesc()
    d.B, d.H, d.M, d.w, d.D = B, nh, M, w, D
    d.dtype, d.flags, d.eps = _DT[cdtype], flags, float(eps)
    d.grid[0], d.grid[1], d.grid[2] = F_, H_, W_
    d.layout[0], d.layout[1], d.layout[2] = fb, hb, wb
    nbytes = L.mhla_blockmix_workspace_bytes(C.byref(d))
    if nbytes == 0:
        raise _capi.MhlaError(f"unsupported 3-D block view: grid {tuple(grid)} layout {tuple(layout)} D={D} "
                              "(needs p2*p3 <= 128 and at most two sub-tiles per block)")
    lay = (C.c_size_t * 8)()
    _capi.check(L.mhla_blockmix_workspace_layout(C.byref(d), C.byref(lay)), "mhla_blockmix_workspace_layout")
    d.q, d.k, d.v, d.out, d.q_rope, d.k_rope = t5(q4), t5(k4), t5(v4), t5(o4), t5(qr4), t5(kr4)
    d.out_gate, d.out_add = t5(g4), t5(a4)
    d.mix, d.mix_ld = mix2.data_ptr(), mix2.stride(0)
    rms_w = None
    if out_rms_weight is not None:
        rms_w = out_rms_weight.detach().to(device=q.device, dtype=torch.float32).contiguous()
    d.out_rms_weight, d.out_rms_eps = (rms_w.data_ptr() if rms_w is not None else None), float(out_rms_eps)
    stream = torch.cuda.current_stream(q.device)
    sig = ("grid", B, nh, F_, H_, W_, fb, hb, wb, D, cdtype, flags, qr4 is not None, g4 is not None, a4 is not None)
    if three_launch:
        ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=q.device)
    elif ws_out is not None:
        ws = _new_ws(q.device, nbytes, int(lay[4]))
    else:
        ws = _persistent_ws(q.device, nbytes, stream.cuda_stream, sig, int(lay[4]))
    d.workspace, d.workspace_bytes = (ws.data_ptr() + 1023) // 1024 * 1024, nbytes
    if ws_out is not None:
        ws_out.update(_ws_views(ws, tuple(int(x) for x in lay), B * nh, M, D), ws=ws, D=D)
    with torch.cuda.device(q.device):
        _capi.check(L.mhla_fwd_blockmix(C.byref(d), stream.cuda_stream), "mhla_fwd_blockmix")
    return o4


def wan_prep(xq: torch.Tensor, xk: torch.Tensor, wq: Optional[torch.Tensor], wk: Optional[torch.Tensor], cos: Optional[torch.Tensor],
             sin: Optional[torch.Tensor], head_dim: int, *, eps_norm: float = 1e-6, eps: float = 1e-6, want_plain: bool = False,
             out_dtype: Optional[torch.dtype] = None):
    """Fused Wan pre-processing (one launch, csrc/wan_prep_kernel.cuh): xq, xk [B, N, C] projection outputs (bf16 / fp16 /
    fp32) -> relu(rmsnorm_C(x) * w) + eps, RoPE by the [N, D/2] cos / sin tables, written token-major in 16 bit.
    Returns (q_rope, k_rope, q_plain, k_plain) as [B, N, heads, D] views; the plain pair is None unless ``want_plain``.
    Replaces mhla_utils.py:267-276 + :127-156 + :303-316 (fp32 / complex128 temporaries, ~10 launches, a host sync)."""
    _require_cuda(xq, xk, wq, wk, cos, sin)
    B, N, Cc = xq.shape
    D = int(head_dim)
    in_code = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}[xq.dtype]
    odt = out_dtype or (torch.float16 if xq.dtype == torch.float16 else torch.bfloat16)
    xq2 = xq.detach().reshape(B * N, Cc)
    xk2 = xk.detach().to(xq.dtype).reshape(B * N, Cc)
    if xq2.stride(1) != 1 or xq2.stride(0) % 8:
        xq2 = xq2.contiguous()
    if xk2.stride(1) != 1 or xk2.stride(0) != xq2.stride(0):
        xk2 = xk2.contiguous()
        xq2 = xq2.contiguous()
    mk = lambda: torch.empty((B, N, Cc // D, D), dtype=odt, device=xq.device)  # noqa: E731
    qr, kr = mk(), mk()
    qp, kp = (mk(), mk()) if want_plain else (None, None)
    f32 = lambda t: None if t is None else t.detach().to(device=xq.device, dtype=torch.float32).contiguous()  # noqa: E731
    wq, wk, cos, sin = f32(wq), f32(wk), f32(cos), f32(sin)
    d = _capi.WanPrepDesc()
    d.rows, d.N, d.C, d.D = B * N, N, Cc, D
    d.in_dtype, d.out_dtype = in_code, _DT[odt]
    d.xq, d.xk, d.ld_in = xq2.data_ptr(), xk2.data_ptr(), xq2.stride(0)
    d.q_rope, d.k_rope = qr.data_ptr(), kr.data_ptr()
    d.q_plain, d.k_plain = (qp.data_ptr(), kp.data_ptr()) if want_plain else (None, None)
    d.wq, d.wk = (wq.data_ptr() if wq is not None else None), (wk.data_ptr() if wk is not None else None)
    d.cos_table, d.sin_table = (cos.data_ptr(), sin.data_ptr()) if cos is not None else (None, None)
    d.eps_norm, d.eps = float(eps_norm), float(eps)
    with torch.cuda.device(xq.device):
        _capi.check(_capi.lib().mhla_wan_prep(C.byref(d), torch.cuda.current_stream().cuda_stream), "mhla_wan_prep")
    return qr, kr, qp, kp


def gated_rmsnorm(x: torch.Tensor, g: Optional[torch.Tensor], weight: Optional[torch.Tensor], eps: float = 1e-5) -> torch.Tensor:
    """y = x * rsqrt(mean(x^2, -1) + eps) * weight * g * sigmoid(g) per row of the last dim, one launch
    (csrc/gated_norm_kernel.cuh; FusedRMSNormGated of the NLP layer, fla/modules/fused_norm_gate.py:77-99 as called at
    layers/mhla.py:350-356).  ``g=None``: plain RMSNorm.  x, g: [..., D] bf16 / fp16 CUDA tensors, D in {64, 128, 256};
    Inference only."""
    _require_cuda(x, g, weight)
    D = x.shape[-1]
    if x.dtype not in _DT or D not in (64, 128, 256):
        raise ValueError("gated_rmsnorm needs a bf16 / fp16 tensor with a last dim of 64, 128 or 256")

    def rows(t):   # [rows, D]; a ragged / padded operator output (sliced view) is made contiguous first
        t = t.detach()
        if t.dtype != x.dtype:
            t = t.to(x.dtype)
        return t.contiguous().view(-1, D)
    x2 = rows(x)
    g2 = rows(g) if g is not None else None
    out = torch.empty(tuple(x.shape), dtype=x.dtype, device=x.device)
    d = _capi.GatedNormDesc()
    d.rows, d.D, d.dtype = x2.shape[0], D, _DT[x.dtype]
    d.x, d.ld_x = x2.data_ptr(), x2.stride(0)
    d.g, d.ld_g = (g2.data_ptr(), g2.stride(0)) if g2 is not None else (None, 0)
    wf = None if weight is None else weight.detach().to(device=x.device, dtype=torch.float32).contiguous()
    d.weight, d.eps, d.out = (wf.data_ptr() if wf is not None else None), float(eps), out.data_ptr()
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib().mhla_gated_rmsnorm(C.byref(d), torch.cuda.current_stream().cuda_stream), "mhla_gated_rmsnorm")
    return out


def gate_add(x: torch.Tensor, gate: Optional[torch.Tensor] = None, add: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = x * silu(gate) + add in one streaming launch (``mhla_gate_add``, csrc/gated_norm_kernel.cuh): the SiLU gate
    and "+ lepe" that follow the operator in the Wan / DiT layers (mhla_utils.py:360-366, wan/model.py:1001-1003,
    mhla.py:268-273).  x, gate, add: [..., C] bf16 / fp16 CUDA tensors of one shape (C % 8 == 0); ``out`` may be ``x``
    (in place).  Inference only."""
    _require_cuda(x, gate, add)
    if x.dtype not in _DT or x.shape[-1] % 8:
        raise ValueError("gate_add needs a bf16 / fp16 tensor whose last dim is a multiple of 8")
    Cc = x.shape[-1]

    def rows(t):   # [rows, C] with one row pitch; anything else is made contiguous first
        t = t.detach()
        if t.dtype != x.dtype:
            t = t.to(x.dtype)
        if tuple(t.shape) != tuple(x.shape):
            raise ValueError("gate_add: gate / add must have the shape of x")
        if not t.is_contiguous():
            t = t.contiguous()
        return t.view(-1, Cc)
    x2 = rows(x)
    g2 = rows(gate) if gate is not None else None
    a2 = rows(add) if add is not None else None
    if out is None:
        out = torch.empty(tuple(x.shape), dtype=x.dtype, device=x.device)
    elif out.dtype != x.dtype or tuple(out.shape) != tuple(x.shape) or not out.is_contiguous():
        raise ValueError("gate_add: out must be a contiguous tensor of x's shape and dtype")
    d = _capi.GateAddDesc()
    d.rows, d.C, d.dtype = x2.shape[0], Cc, _DT[x.dtype]
    d.x, d.ld_x = x2.data_ptr(), x2.stride(0)
    d.g, d.ld_g = (g2.data_ptr(), g2.stride(0)) if g2 is not None else (None, 0)
    d.add, d.ld_add = (a2.data_ptr(), a2.stride(0)) if a2 is not None else (None, 0)
    d.out, d.ld_out = out.data_ptr(), Cc
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib().mhla_gate_add(C.byref(d), torch.cuda.current_stream().cuda_stream), "mhla_gate_add")
    return out


def dwconv3d_tokens(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], grid) -> torch.Tensor:
    """Depthwise 3x3x3 convolution (zero padding 1) over the (F, H, W) token grid of a TOKEN-major tensor: x [B, F*H*W, C]
    bf16 / fp16 -> [B, F*H*W, C] (``mhla_dwconv3d``, csrc/gated_norm_kernel.cuh).  ``weight`` / ``bias`` are the parameters
    of the reference's ``self.lepe = nn.Conv3d(C, C, 3, padding=1, groups=C)`` ([C, 1, 3, 3, 3] / [C]): the result equals
    ``rearrange(lepe(rearrange(x, "b (f h w) c -> b c f h w")), "b c f h w -> b (f h w) c")`` of mhla_utils.py:289-296
    without the two rearrangements and cuDNN's depthwise path.  Inference only."""
    _require_cuda(x, weight, bias)
    B, N, Cc = x.shape
    F_, H_, W_ = (int(v) for v in grid)
    if x.dtype not in _DT or Cc % 8 or N != F_ * H_ * W_ or tuple(weight.shape) != (Cc, 1, 3, 3, 3):
        raise ValueError("dwconv3d_tokens needs a bf16 / fp16 [B, F*H*W, C] tensor (C % 8 == 0) and a [C, 1, 3, 3, 3] weight")
    x = x.detach()
    if x.stride(-1) != 1 or x.stride(0) != N * x.stride(1):
        x = x.contiguous()
    wt = weight.detach().to(torch.float32).reshape(Cc, 27).t().contiguous()          # [27, C]
    bf = None if bias is None else bias.detach().to(torch.float32).contiguous()
    out = torch.empty((B, N, Cc), dtype=x.dtype, device=x.device)
    d = _capi.DwConv3dDesc()
    d.B, d.F, d.H, d.W, d.C, d.dtype = B, F_, H_, W_, Cc, _DT[x.dtype]
    d.x, d.ld_x, d.wt, d.bias, d.out = x.data_ptr(), x.stride(1), wt.data_ptr(), (bf.data_ptr() if bf is not None else None), out.data_ptr()
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib().mhla_dwconv3d(C.byref(d), torch.cuda.current_stream().cuda_stream), "mhla_dwconv3d")
    return out


def block_wsum(x: torch.Tensor, wgt: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[..., :] = sum_t wgt[..., t] * x[..., t, :] over the token axis of every block (``mhla_block_wsum``): x [..., w, D]
    16-bit, wgt [..., w] fp32 or None (plain sums) -> [..., D] fp32.  ksum and its gradient partner of the backward pass."""
    _require_cuda(x, wgt)
    w, D = x.shape[-2], x.shape[-1]
    x = x.contiguous()
    out = torch.empty(tuple(x.shape[:-2]) + (D,), dtype=torch.float32, device=x.device)
    d = _capi.BlockWsumDesc()
    d.blocks, d.w, d.D, d.dtype = x.numel() // (w * D), w, D, _DT[x.dtype]
    wf = None if wgt is None else wgt.contiguous().to(torch.float32)
    d.x, d.wgt, d.out = x.data_ptr(), (wf.data_ptr() if wf is not None else None), out.data_ptr()
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib().mhla_block_wsum(C.byref(d), torch.cuda.current_stream().cuda_stream), "mhla_block_wsum")
    return out


def bwd_prep(do: torch.Tensor, out: torch.Tensor, den: torch.Tensor):
    """dnum = do / den[..., None] (16 bit) and dden = -(do . out) / den (fp32) in one pass over the token rows
    (csrc/bwd_aux_kernel.cuh; ``mhla_bwd_prep``).  do, out: [..., D] 16-bit, den: [...] fp32; D in {64, 128}."""
    _require_cuda(do, out, den)
    D = do.shape[-1]
    do, out = do.contiguous(), out.contiguous().to(do.dtype)
    den = den.contiguous().to(torch.float32)
    dnum = torch.empty_like(do)
    dden = torch.empty_like(den)
    d = _capi.BwdPrepDesc()
    d.rows, d.D, d.dtype = do.numel() // D, D, _DT[do.dtype]
    d.dout, d.out, d.den, d.dnum, d.dden = do.data_ptr(), out.data_ptr(), den.data_ptr(), dnum.data_ptr(), dden.data_ptr()
    with torch.cuda.device(do.device):
        _capi.check(_capi.lib().mhla_bwd_prep(C.byref(d), torch.cuda.current_stream().cuda_stream), "mhla_bwd_prep")
    return dnum, dden


def bwd_post(dqn: Optional[torch.Tensor], dkn: Optional[torch.Tensor], dnl: torch.Tensor, ksum: torch.Tensor,
             dksum: torch.Tensor, dtype: torch.dtype):
    """dq = dqn + dnl[..., None] * ksum[..., None, :], dk = dkn + dksum[..., None, :] in one pass (``mhla_bwd_post``).
    dqn, dkn: [..., M, w, D] 16-bit or None; dnl [..., M, w] fp32; ksum, dksum [..., M, D] fp32."""
    _require_cuda(dnl, ksum, dksum, dqn, dkn)
    w, D = dnl.shape[-1], ksum.shape[-1]
    dnl, ksum, dksum = (t.contiguous().to(torch.float32) for t in (dnl, ksum, dksum))
    if dqn is not None:
        dqn, dkn = dqn.contiguous(), dkn.contiguous()
    dq = torch.empty(tuple(dnl.shape) + (D,), dtype=dtype, device=dnl.device)
    dk = torch.empty_like(dq)
    d = _capi.BwdPostDesc()
    d.rows, d.w, d.D, d.dtype = dnl.numel(), w, D, _DT[dtype]
    d.dqn, d.dkn = (dqn.data_ptr(), dkn.data_ptr()) if dqn is not None else (None, None)
    d.dnl, d.ksum, d.dksum, d.dq, d.dk = dnl.data_ptr(), ksum.data_ptr(), dksum.data_ptr(), dq.data_ptr(), dk.data_ptr()
    with torch.cuda.device(dnl.device):
        _capi.check(_capi.lib().mhla_bwd_post(C.byref(d), torch.cuda.current_stream().cuda_stream), "mhla_bwd_post")
    return dq, dk


_HOST_STREAMS: "dict[int, tuple]" = {}
_HOST_BUFS: "dict[tuple, tuple]" = {}


def mhla_host(q, k, v, mix, *, out: Optional[torch.Tensor] = None, device=None, eps: float = 1e-6, normalize: bool = True,
              chunks: int = 4) -> torch.Tensor:
    """Block-mixed forward for HOST tensors ([B, H, M, w, D], ideally pinned): the (b,h) units are independent, so the
    batch is cut into `chunks` contiguous ranges of units and the host->device copies of range c+1, the kernel of range
    c and the device->host copy of range c-1 run concurrently on three streams (PCIe is full duplex; the kernel time
    disappears behind the copies).  Returns a host tensor (`out`, pinned if given so, else freshly pinned).

    The device->host copies are ASYNCHRONOUS: the caller's current stream is made to wait for them, so any later work on
    that stream is ordered, but the HOST must synchronise (``torch.cuda.current_stream().synchronize()``) before reading
    `out` from Python.  `out` must be contiguous (it is written through a reshaped view)."""
    if q.is_cuda:
        raise ValueError("mhla_host takes host tensors; use mhla() for device tensors")
    if out is not None and not out.is_contiguous():
        raise ValueError("mhla_host: out must be contiguous")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    B, H, M, w, D = q.shape
    if out is None:
        out = torch.empty(q.shape, dtype=q.dtype).pin_memory()
    units = B * H
    chunks = max(1, min(chunks, units))
    qf, kf, vf, of = (t.reshape(units, 1, M, w, D) for t in (q, k, v, out))   # a unit is a (b,h) pair: [units, 1, M, w, D]
    with torch.cuda.device(dev):
        if dev.index not in _HOST_STREAMS:
            _HOST_STREAMS[dev.index] = (torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream())
        s_in, s_run, s_out = _HOST_STREAMS[dev.index]
        cur = torch.cuda.current_stream()
        for s_ in (s_in, s_run, s_out):
            s_.wait_stream(cur)
        mix_d = mix.to(dev, non_blocking=True) if not mix.is_cuda else mix
        s_run.wait_stream(cur)
        # device staging buffers are kept per shape: every call orders itself after the previous one through the
        # caller's stream (first and last statements of this block), so reuse is safe and no allocator traffic is left
        bkey = (dev.index, units, M, w, D, q.dtype)
        bufs = _HOST_BUFS.get(bkey)
        if bufs is None:
            if len(_HOST_BUFS) >= 2:
                _HOST_BUFS.pop(next(iter(_HOST_BUFS)))
            bufs = _HOST_BUFS[bkey] = tuple(torch.empty((units, 1, M, w, D), dtype=q.dtype, device=dev) for _ in range(4))
        dq, dk, dv, do = bufs
        per = (units + chunks - 1) // chunks
        for c in range(chunks):
            lo, hi = c * per, min(units, (c + 1) * per)
            if lo >= hi:
                break
            with torch.cuda.stream(s_in):
                dq[lo:hi].copy_(qf[lo:hi], non_blocking=True)
                dk[lo:hi].copy_(kf[lo:hi], non_blocking=True)
                dv[lo:hi].copy_(vf[lo:hi], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            with torch.cuda.stream(s_run):
                s_run.wait_event(ev_in)
                mhla_blockmix(dq[lo:hi], dk[lo:hi], dv[lo:hi], mix_d, eps=eps, normalize=normalize, out=do[lo:hi])
                ev_run = torch.cuda.Event()
                ev_run.record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_run)
                of[lo:hi].copy_(do[lo:hi], non_blocking=True)
        cur.wait_stream(s_out)
    return out


def _t4(t: torch.Tensor) -> _capi.Tensor4:
    sb, st, sh, _ = t.stride()
    return _capi.Tensor4(t.data_ptr(), sb, st, sh)


def mhla_causal(q, k, v, mixing_matrix, chunk_size: int = 64, scale: Optional[float] = None,
                unfused: Optional[bool] = None) -> torch.Tensor:
    """Causal chunked MHLA (see ``_causal_fwd``); differentiable through ``autograd.CausalFunction``."""
    if _needs_grad(q, k, v, mixing_matrix):
        from .autograd import CausalFunction
        return CausalFunction.apply(q, k, v, mixing_matrix, int(chunk_size), scale, unfused)
    return _causal_fwd(q, k, v, mixing_matrix, chunk_size=chunk_size, scale=scale, unfused=unfused)


def _causal_fwd(q, k, v, mixing_matrix, chunk_size: int = 64, scale: Optional[float] = None,
                unfused: Optional[bool] = None, debug_flags: int = 0) -> torch.Tensor:
    """Causal chunked MHLA forward; q,k [B,T,H,K], v [B,T,H,V] -> o [B,T,H,V] in q.dtype.

    T is zero-padded to a multiple of the chunk exactly as the reference does (naive.py:46-51); the pad is a host-side
    copy that only happens for ragged T."""
    _require_cuda(q, k, v, mixing_matrix)
    q, k, v = q.detach(), k.detach(), v.detach()
    in_dtype = q.dtype
    cdtype = in_dtype if in_dtype in _DT else torch.bfloat16
    B, T_in, H, K_in = q.shape
    V_in = v.shape[-1]
    # head dims between the kernel's widths run zero-padded (exact: padded channels add zeros to q.k and to S)
    K = 64 if K_in <= 64 else 128
    V = 64 if V_in <= 64 else (128 if V_in <= 128 else 256)
    if K_in > 128 or V_in > 256:
        raise _capi.MhlaError(f"unsupported causal head dims K={K_in} V={V_in} (K <= 128, V <= 256)")
    pad = (chunk_size - T_in % chunk_size) % chunk_size
    T = T_in + pad

    def prep(t, dpad):
        t = t.to(cdtype) if t.dtype != cdtype else t
        if pad or dpad:
            t = torch.nn.functional.pad(t, (0, dpad, 0, 0, 0, pad))
        return t if _tma_ok(t) else t.contiguous()

    q4, k4, v4 = prep(q, K - K_in), prep(k, K - K_in), prep(v, V - V_in)
    o4 = torch.empty((B, T, H, V), dtype=cdtype, device=q.device)
    Lm = mixing_matrix.shape[0]
    mm = mixing_matrix.detach().reshape(Lm, mixing_matrix.shape[1]).to(torch.float32).contiguous()
    n = T // chunk_size
    if n > Lm:
        raise IndexError(f"mixing matrix is {Lm}x{Lm} but T={T_in} needs {n} chunks of {chunk_size}")
    d = _capi.CausalDesc()
    d.B, d.T, d.H, d.K, d.V = B, T, H, K, V
    # unfused: None = let the library choose (three launches for large batches), True / False force either structure
    d.chunk, d.dtype = chunk_size, _DT[cdtype]
    d.flags = (0 if unfused is None else (_capi.FLAG_UNFUSED if unfused else _capi.FLAG_FUSED)) | int(debug_flags)
    d.scale = float(K_in ** -0.5 if scale is None else scale)
    d.q, d.k, d.v, d.out = _t4(q4), _t4(k4), _t4(v4), _t4(o4)
    d.mm, d.mm_ld, d.L = mm.data_ptr(), mm.stride(0), Lm
    L = _capi.lib()
    nbytes = L.mhla_causal_workspace_bytes(C.byref(d))
    if nbytes == 0:
        raise _capi.MhlaError(f"unsupported causal shape T={T} K={K} V={V} chunk={chunk_size}")
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=q.device)
    d.workspace, d.workspace_bytes = (ws.data_ptr() + 1023) // 1024 * 1024, nbytes
    with torch.cuda.device(q.device):
        _capi.check(L.mhla_fwd_causal(C.byref(d), torch.cuda.current_stream().cuda_stream), "mhla_fwd_causal")
    if pad or V != V_in:
        o4 = o4[:, :T_in, :, :V_in]
    return o4 if in_dtype == cdtype else o4.to(in_dtype)


def mhla(q, k, v, mix, *, q_rope=None, k_rope=None, eps: float = 1e-6, normalize: bool = True,
         causal: bool = False, chunk: Optional[int] = None, **kw) -> torch.Tensor:
    """The operator signature SURVEY.md 8b defines: one entry point for the three reference variants."""
    if causal:
        return mhla_causal(q, k, v, mix, chunk_size=chunk or 64)
    return mhla_blockmix(q, k, v, mix, q_rope=q_rope, k_rope=k_rope, eps=eps, normalize=normalize, **kw)


def naive_chunk_simple_mhla_fixed(q, k, v, mixing_matrix, output_final_state: bool = False, chunk_size: int = 64,
                                  *args, **kwargs):
    """Drop-in for mhla_nlp/fla/ops/mhla/naive.py:10-83 (returns o only; ``output_final_state`` is a no-op there too)."""
    return mhla_causal(q, k, v, mixing_matrix, chunk_size=chunk_size)


def naive_recurrent_mhla(q, k, v, mixing_matrix, chunk_size: int = 64, scale: Optional[float] = None,
                         initial_state=None, output_final_state: bool = True):
    """Drop-in for naive.py:88-142.  The layer only selects it for q_len <= 64 (layers/mhla.py:247), where it equals
    the chunk form; beyond one chunk the reference's recurrence is inconsistent and its returned state is all zeros
    (SURVEY.md 0.4), so this raises instead of reproducing the bug.  Returns (o, None)."""
    if q.shape[1] > chunk_size:
        raise ValueError("naive_recurrent_mhla is only defined for T <= chunk_size (reference defect, SURVEY.md 0.4)")
    if initial_state is not None:
        raise NotImplementedError("initial_state is not supported (the reference never produces a usable state)")
    return mhla_causal(q, k, v, mixing_matrix, chunk_size=chunk_size, scale=scale), None


def last_launch_count() -> int:
    return int(_capi.lib().mhla_last_launch_count())
