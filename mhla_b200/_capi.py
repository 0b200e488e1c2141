"""ctypes binding of libmhla_b200.so (see include/mhla_b200.h).  No CPU fallback: if the library is missing
or a call fails the caller gets an exception."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MHLA_B200_LIB") or os.path.join(_HERE, "libmhla_b200.so")   # env override: A/B builds while tuning

MHLA_BF16, MHLA_FP16 = 0, 1
FLAG_NORMALIZE = 1 << 0
FLAG_FUSED = 1 << 7
FLAG_UNFUSED = 1 << 8
FLAG_STOP_AFTER_P1 = 1 << 9
FLAG_STOP_AFTER_P2 = 1 << 10
FLAG_ONLY_P3 = 1 << 11
FLAG_ONLY_P2 = 1 << 12
FLAG_TWO_LAUNCH = 1 << 13
FLAG_WS_PERSISTENT = 1 << 14
FLAG_NO_SMALLN = 1 << 15

EXPORTS = [
    "mhla_abi_version", "mhla_strerror", "mhla_last_cuda_error", "mhla_last_launch_count",
    "mhla_blockmix_workspace_bytes", "mhla_blockmix_needs_workspace", "mhla_blockmix_workspace_layout", "mhla_fwd_blockmix", "mhla_blockmix_workspace_init", "mhla_causal_workspace_bytes", "mhla_fwd_causal", "mhla_wan_prep", "mhla_bwd_prep", "mhla_bwd_post", "mhla_block_wsum", "mhla_gated_rmsnorm", "mhla_gate_add", "mhla_dwconv3d",
]


class Tensor5(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride_b", C.c_int64), ("stride_h", C.c_int64), ("stride_m", C.c_int64),
                ("stride_w", C.c_int64)]


class BlockmixDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("M", C.c_int32), ("w", C.c_int32), ("D", C.c_int32),
        ("dtype", C.c_int32), ("flags", C.c_uint32), ("eps", C.c_float),
        ("q", Tensor5), ("k", Tensor5), ("v", Tensor5), ("q_rope", Tensor5), ("k_rope", Tensor5), ("out", Tensor5),
        ("mix", C.c_void_p), ("mix_ld", C.c_int64), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("out_rms_weight", C.c_void_p), ("out_rms_eps", C.c_float),
        ("grid", C.c_int32 * 3), ("layout", C.c_int32 * 3),
        ("out_gate", Tensor5), ("out_add", Tensor5),
    ]


class Tensor4(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride_b", C.c_int64), ("stride_t", C.c_int64), ("stride_h", C.c_int64)]


class CausalDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("T", C.c_int32), ("H", C.c_int32), ("K", C.c_int32), ("V", C.c_int32),
        ("chunk", C.c_int32), ("dtype", C.c_int32), ("flags", C.c_uint32), ("scale", C.c_float),
        ("q", Tensor4), ("k", Tensor4), ("v", Tensor4), ("out", Tensor4),
        ("mm", C.c_void_p), ("mm_ld", C.c_int64), ("L", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class WanPrepDesc(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("N", C.c_int32), ("C", C.c_int32), ("D", C.c_int32),
        ("in_dtype", C.c_int32), ("out_dtype", C.c_int32),
        ("xq", C.c_void_p), ("xk", C.c_void_p), ("ld_in", C.c_int64),
        ("q_rope", C.c_void_p), ("k_rope", C.c_void_p), ("q_plain", C.c_void_p), ("k_plain", C.c_void_p),
        ("wq", C.c_void_p), ("wk", C.c_void_p), ("cos_table", C.c_void_p), ("sin_table", C.c_void_p),
        ("eps_norm", C.c_float), ("eps", C.c_float),
    ]


class BwdPrepDesc(C.Structure):
    _fields_ = [("rows", C.c_int64), ("D", C.c_int32), ("dtype", C.c_int32), ("dout", C.c_void_p), ("out", C.c_void_p),
                ("den", C.c_void_p), ("dnum", C.c_void_p), ("dden", C.c_void_p)]


class BwdPostDesc(C.Structure):
    _fields_ = [("rows", C.c_int64), ("w", C.c_int32), ("D", C.c_int32), ("dtype", C.c_int32), ("dqn", C.c_void_p),
                ("dkn", C.c_void_p), ("dnl", C.c_void_p), ("ksum", C.c_void_p), ("dksum", C.c_void_p),
                ("dq", C.c_void_p), ("dk", C.c_void_p)]


class BlockWsumDesc(C.Structure):
    _fields_ = [("blocks", C.c_int64), ("w", C.c_int32), ("D", C.c_int32), ("dtype", C.c_int32), ("x", C.c_void_p),
                ("wgt", C.c_void_p), ("out", C.c_void_p)]


class GatedNormDesc(C.Structure):
    _fields_ = [("rows", C.c_int64), ("D", C.c_int32), ("dtype", C.c_int32), ("x", C.c_void_p), ("ld_x", C.c_int64),
                ("g", C.c_void_p), ("ld_g", C.c_int64), ("weight", C.c_void_p), ("eps", C.c_float), ("out", C.c_void_p)]


class GateAddDesc(C.Structure):
    _fields_ = [("rows", C.c_int64), ("C", C.c_int32), ("dtype", C.c_int32), ("x", C.c_void_p), ("ld_x", C.c_int64),
                ("g", C.c_void_p), ("ld_g", C.c_int64), ("add", C.c_void_p), ("ld_add", C.c_int64),
                ("out", C.c_void_p), ("ld_out", C.c_int64)]


class DwConv3dDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("F", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("dtype", C.c_int32),
                ("x", C.c_void_p), ("ld_x", C.c_int64), ("wt", C.c_void_p), ("bias", C.c_void_p), ("out", C.c_void_p)]


class MhlaError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load the extension (once).  Raises ImportError when it has not been built - there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m mhla_b200.build` (nvcc, sm_100a). "
                "mhla_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        L.mhla_abi_version.restype = C.c_int
        L.mhla_strerror.restype = C.c_char_p
        L.mhla_strerror.argtypes = [C.c_int]
        L.mhla_last_cuda_error.restype = C.c_char_p
        L.mhla_last_launch_count.restype = C.c_int
        L.mhla_blockmix_workspace_bytes.restype = C.c_size_t
        L.mhla_blockmix_workspace_bytes.argtypes = [C.POINTER(BlockmixDesc)]
        if hasattr(L, "mhla_blockmix_needs_workspace"):     # (absent in older A/B builds loaded through MHLA_B200_LIB)
            L.mhla_blockmix_needs_workspace.restype = C.c_int
            L.mhla_blockmix_needs_workspace.argtypes = [C.POINTER(BlockmixDesc)]
        L.mhla_blockmix_workspace_layout.restype = C.c_int
        L.mhla_blockmix_workspace_layout.argtypes = [C.POINTER(BlockmixDesc), C.POINTER(C.c_size_t * 8)]
        L.mhla_fwd_blockmix.restype = C.c_int
        L.mhla_fwd_blockmix.argtypes = [C.POINTER(BlockmixDesc), C.c_void_p]
        L.mhla_blockmix_workspace_init.restype = C.c_int
        L.mhla_blockmix_workspace_init.argtypes = [C.POINTER(BlockmixDesc), C.c_void_p]
        L.mhla_causal_workspace_bytes.restype = C.c_size_t
        L.mhla_causal_workspace_bytes.argtypes = [C.POINTER(CausalDesc)]
        if hasattr(L, "mhla_wan_prep"):
            L.mhla_wan_prep.restype = C.c_int
            L.mhla_wan_prep.argtypes = [C.POINTER(WanPrepDesc), C.c_void_p]
        L.mhla_fwd_causal.restype = C.c_int
        L.mhla_fwd_causal.argtypes = [C.POINTER(CausalDesc), C.c_void_p]
        if hasattr(L, "mhla_bwd_prep"):
            L.mhla_bwd_prep.restype = C.c_int
            L.mhla_bwd_prep.argtypes = [C.POINTER(BwdPrepDesc), C.c_void_p]
            L.mhla_bwd_post.restype = C.c_int
            L.mhla_bwd_post.argtypes = [C.POINTER(BwdPostDesc), C.c_void_p]
        if hasattr(L, "mhla_block_wsum"):
            L.mhla_block_wsum.restype = C.c_int
            L.mhla_block_wsum.argtypes = [C.POINTER(BlockWsumDesc), C.c_void_p]
        if hasattr(L, "mhla_gated_rmsnorm"):
            L.mhla_gated_rmsnorm.restype = C.c_int
            L.mhla_gated_rmsnorm.argtypes = [C.POINTER(GatedNormDesc), C.c_void_p]
        if hasattr(L, "mhla_gate_add"):
            L.mhla_gate_add.restype = C.c_int
            L.mhla_gate_add.argtypes = [C.POINTER(GateAddDesc), C.c_void_p]
        if hasattr(L, "mhla_dwconv3d"):
            L.mhla_dwconv3d.restype = C.c_int
            L.mhla_dwconv3d.argtypes = [C.POINTER(DwConv3dDesc), C.c_void_p]
        if L.mhla_abi_version() != 4:
            raise ImportError("libmhla_b200.so ABI version mismatch; rebuild with `python -m mhla_b200.build --force`")
        _lib = L
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        L = lib()
        msg = L.mhla_strerror(status).decode()
        detail = L.mhla_last_cuda_error().decode()
        raise MhlaError(f"{what} failed: {msg} ({status})" + (f" -- {detail}" if detail else ""))
