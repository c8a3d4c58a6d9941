"""The C-ABI library loads without a GPU and exports every symbol include/ivln_map.h declares
(no compute calls here)."""
import ctypes
import os
import re

from ivlnce_b200 import _lib
from ivlnce_b200.build import build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ivln_map.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ivm_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    build_library()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ivln_map.h but not exported"


def test_binding_loads_and_reports_version():
    L = _lib.load()
    assert b"sm_100a" in L.ivm_version()


def test_workspace_size_and_argument_checks():
    L = _lib.load()
    cfg = _lib.IvmConfig(max_envs=16, height=256, width=256, map_rows=128, map_cols=128, res=0.05, half_res=0.025,
                         half_h=3.2, half_w=3.2, store_rows=2048, store_cols=2048, mode=0, known_capacity=0)
    n = L.ivm_workspace_bytes(ctypes.byref(cfg))
    # 16-byte record + 8-byte frame candidate word per half-cell, plus small bookkeeping
    assert 24 * 2048 * 2048 * 16 < n < 25 * 2048 * 2048 * 16
    bad = _lib.IvmConfig(max_envs=0)
    assert L.ivm_workspace_bytes(ctypes.byref(bad)) == 0
    ctx = ctypes.c_void_p()
    assert L.ivm_create(ctypes.byref(cfg), None, n, ctypes.byref(ctx)) == 1       # IVM_E_INVALID: no workspace
    assert L.ivm_create(ctypes.byref(cfg), 4096 + 8, n, ctypes.byref(ctx)) == 2   # IVM_E_WORKSPACE: misaligned
    assert L.ivm_step_iterative(None, 1, None, None, None, 0, None, None, None, None, None, 0, None, None, None, None) == 1


def test_config_struct_matches_header_layout():
    # 5 i32, 4 f32, 3 i32, i64, 2 i32, 4 i32 with natural alignment
    assert ctypes.sizeof(_lib.IvmConfig) == 80
    assert _lib.IvmConfig.known_capacity.offset == 48
    assert ctypes.sizeof(_lib.IvmStatus) == 72
