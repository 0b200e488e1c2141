"""Build libmhla_b200.so (the C-ABI extension) in-tree with nvcc for sm_100a.

    python -m mhla_b200.build            # rebuild if sources are newer than the library
    python -m mhla_b200.build --force
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmhla_b200.so")
SOURCES = ["mhla_capi.cu"]
HEADERS = ["ptx.cuh", "blockmix_kernel.cuh", "causal_kernel.cuh", "smalln_kernel.cuh", "wan_prep_kernel.cuh", "bwd_aux_kernel.cuh", "gated_norm_kernel.cuh", os.path.join("..", "..", "include", "mhla_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "--use_fast_math",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libmhla_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


LIB_DIAG = os.path.join(HERE, "libmhla_b200_diag.so")


def build_diag() -> str:
    """The diagnostics build (-DMHLA_DIAG: time-bounded waits that record a stall in host-mapped memory before trapping);
    tools/stress.py loads it through MHLA_B200_LIB.  Not used by the product path."""
    cmd = [_nvcc(), *NVCC_FLAGS, "-DMHLA_DIAG", "-o", LIB_DIAG, *[os.path.join(CSRC, s) for s in SOURCES]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{proc.stdout}\n{proc.stderr}")
    return LIB_DIAG


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    if "--diag" in sys.argv:
        print(build_diag())
