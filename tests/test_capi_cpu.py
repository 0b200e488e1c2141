"""The C-ABI library loads without a GPU and exports every entry point include/mhla_b200.h declares; host-only entry
points (planning, status strings) behave.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from mhla_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "mhla_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mhla_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    names = _declared_functions()
    assert {"mhla_fwd_blockmix", "mhla_fwd_causal", "mhla_blockmix_workspace_bytes", "mhla_strerror"} <= set(names)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert set(_capi.EXPORTS) <= set(names)
    assert L.mhla_abi_version() == 4


def test_status_strings():
    L = _capi.lib()
    assert L.mhla_strerror(0) == b"ok"
    for code in range(-6, 0):
        assert len(L.mhla_strerror(code)) > 0
    assert L.mhla_strerror(-99) == b"unknown status"


def _desc(B=2, H=16, M=128, w=256, D=64, flags=1, dtype=0):
    d = _capi.BlockmixDesc()
    d.B, d.H, d.M, d.w, d.D, d.dtype, d.flags = B, H, M, w, D, dtype, flags
    return d


def test_blockmix_workspace_planning():
    L = _capi.lib()
    d = _desc()
    n = L.mhla_blockmix_workspace_bytes(C.byref(d))
    lay = (C.c_size_t * 8)()
    assert L.mhla_blockmix_workspace_layout(C.byref(d), C.byref(lay)) == 0
    offS, offSt, offDen, offW, offC, ncols, wpad, Mp = [int(x) for x in lay]
    G = 32
    assert ncols == 64 * 64 + 2 * 256 and wpad == 256 and Mp == 128
    assert offS == 0 and offSt >= G * 128 * ncols * 2 and offDen >= offSt + G * 128 * 4096 * 2
    assert n > offC and n % 1024 == 0
    # without the normaliser no n_loc / den columns are planned
    assert int(L.mhla_blockmix_workspace_bytes(C.byref(_desc(flags=0)))) < n
    # envelope: D in {64,128}, w <= 256
    assert L.mhla_blockmix_workspace_bytes(C.byref(_desc(D=32))) == 0
    assert L.mhla_blockmix_workspace_bytes(C.byref(_desc(w=300))) == 0
    assert L.mhla_blockmix_workspace_bytes(C.byref(_desc(dtype=7))) == 0


def test_small_block_counts_are_packed_into_one_mixing_tile():
    """Planner: with M <= 64 blocks per group, `pack` consecutive (b,h) groups (a divisor of B*H, pack * M <= 128) are
    scheduled as one group with a block-diagonal mixing matrix; layout[7] is the padded size of that matrix."""
    L = _capi.lib()
    lay = (C.c_size_t * 8)()
    for (B, H, M, want) in [(64, 6, 16, 128), (2, 16, 128, 128), (1, 5, 16, 80), (1, 1, 16, 16), (2, 12, 150, 152), (8, 4, 32, 128)]:
        assert L.mhla_blockmix_workspace_layout(C.byref(_desc(B=B, H=H, M=M)), C.byref(lay)) == 0
        assert int(lay[7]) == want, (B, H, M, int(lay[7]))


def test_invalid_arguments_are_rejected_before_any_cuda_call():
    L = _capi.lib()
    d = _desc()
    assert L.mhla_fwd_blockmix(C.byref(d), None) == -1          # NULL tensors -> MHLA_ERR_INVALID_ARGUMENT
    c = _capi.CausalDesc()
    c.B, c.T, c.H, c.K, c.V, c.chunk, c.dtype, c.L = 1, 128, 2, 64, 64, 64, 0, 32
    assert L.mhla_causal_workspace_bytes(C.byref(c)) > 0
    c.K = 96
    assert L.mhla_causal_workspace_bytes(C.byref(c)) == 0
    c.K, c.T = 64, 100                                            # ragged T is padded by the host shim, not the kernel
    assert L.mhla_causal_workspace_bytes(C.byref(c)) == 0
    c.T, c.L = 64 * 40, 32                                        # needs 40 mixing rows
    assert L.mhla_causal_workspace_bytes(C.byref(c)) == 0


def test_streaming_entry_points_validate_their_descriptors():
    """mhla_gate_add / mhla_dwconv3d / mhla_gated_rmsnorm reject bad descriptors before touching CUDA (status codes of
    include/mhla_b200.h: -1 invalid argument, -2 unsupported shape, -3 alignment)."""
    L = _capi.lib()
    g = _capi.GateAddDesc()
    g.rows, g.C, g.dtype = 16, 64, 0
    assert L.mhla_gate_add(C.byref(g), None) == -1                # x / out NULL
    g.x, g.out, g.ld_x, g.ld_out = 0x1000, 0x2000, 64, 64
    g.C = 60
    assert L.mhla_gate_add(C.byref(g), None) == -2                # C % 8
    g.C, g.x = 64, 0x1004
    assert L.mhla_gate_add(C.byref(g), None) == -3                # 16-byte alignment
    g.x, g.ld_x = 0x1000, 32
    assert L.mhla_gate_add(C.byref(g), None) == -3                # pitch shorter than a row
    g.ld_x, g.dtype = 64, 7
    assert L.mhla_gate_add(C.byref(g), None) == -1                # dtype

    d = _capi.DwConv3dDesc()
    d.B, d.F, d.H, d.W, d.C, d.dtype = 1, 3, 4, 5, 64, 0
    assert L.mhla_dwconv3d(C.byref(d), None) == -1                # NULL tensors
    d.x, d.wt, d.out, d.ld_x = 0x1000, 0x2000, 0x3000, 64
    d.C = 12
    assert L.mhla_dwconv3d(C.byref(d), None) == -2                # C % 8
    d.C, d.F = 64, 0
    assert L.mhla_dwconv3d(C.byref(d), None) == -2                # empty grid
    d.F, d.ld_x = 3, 60
    assert L.mhla_dwconv3d(C.byref(d), None) == -3                # pitch

    n = _capi.GatedNormDesc()
    n.rows, n.D, n.dtype = 8, 256, 0
    assert L.mhla_gated_rmsnorm(C.byref(n), None) == -1
    n.x, n.out, n.ld_x = 0x1000, 0x2000, 256
    n.D = 96
    assert L.mhla_gated_rmsnorm(C.byref(n), None) == -2           # D not in {64, 128, 256}


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _capi.lib()
