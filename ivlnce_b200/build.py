"""Builds the CUDA library in-tree: nvcc -> ivlnce_b200/csrc/libivlnmap.so (sm_100a only).

-fmad=false: no multiply-add contraction anywhere (the fused steps are written
as explicit __fmaf_rn in ivm_core.h); no -use_fast_math; IEEE division.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(CSRC, "libivlnmap.so")
SOURCES = [os.path.join(CSRC, "ivm_kernels.cu")]
HEADERS = [os.path.join(CSRC, "ivm_core.h"), os.path.join(os.path.dirname(_HERE), "include", "ivln_map.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; the CUDA library cannot be built")
    return nvcc


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if force or needs_build():
        cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
        subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB
