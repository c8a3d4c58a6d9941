"""TEST INFRASTRUCTURE ONLY -- ctypes front end of tests/emu/emulator.cpp, the serial
host emulator that runs the kernels' per-thread logic (ivm_core.h) on the CPU."""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np
import torch

from ivlnce_b200.geometry import camera_scale_tables, camera_to_world_rows, ego_rotation

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SRC = os.path.join(_HERE, "emu", "emulator.cpp")
_CORE = os.path.join(_ROOT, "ivlnce_b200", "csrc", "ivm_core.h")
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "libemu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        os.makedirs(_BUILD, exist_ok=True)
        newest = max(os.path.getmtime(_SRC), os.path.getmtime(_CORE))
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                            "-Wall", "-o", _SO, _SRC, "-lm"], check=True)
        L = ctypes.CDLL(_SO)
        f32p, u8p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint8)
        i64p, u64p, u32p = ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32)
        L.emu_create.restype = ctypes.c_void_p
        L.emu_create.argtypes = [ctypes.c_int, ctypes.c_int, f32p, f32p] + [ctypes.c_float] * 4 + [ctypes.c_int] * 8 + [ctypes.c_longlong]
        L.emu_destroy.argtypes = [ctypes.c_void_p]
        L.emu_set_order.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.emu_set_fix_cap.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.emu_set_direct.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.emu_step_iterative.argtypes = [ctypes.c_void_p, ctypes.c_int, f32p, u8p, f32p, f32p, f32p, ctypes.c_void_p, ctypes.c_int, u8p, u8p, u8p]
        L.emu_known_load.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, f32p, u8p, ctypes.c_int, ctypes.c_int]
        L.emu_known_clear.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.emu_step_known.argtypes = [ctypes.c_void_p, ctypes.c_int, f32p, f32p, u8p, u8p]
        L.emu_export_world.restype = ctypes.c_longlong
        L.emu_export_world.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, i64p, f32p, u8p, u64p]
        L.emu_status.argtypes = [ctypes.c_void_p, u32p, u64p]
        L.emu_set_stamp_period.argtypes = [ctypes.c_void_p, ctypes.c_uint]
        L.emu_cand_current.restype = ctypes.c_longlong
        L.emu_cand_current.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


class EmuMapper:
    def __init__(self, height, width, vfov, map_m, resolution, max_envs, mode="iterative", store=1024,
                 tile=32, known_clouds=None, known_capacity=1 << 16, order=0, kernel_trig=False, fix_cap=None, direct=False):
        self.kernel_trig = kernel_trig
        self.H, self.W = height, width
        self.R = math.ceil(map_m / resolution)
        self.C = math.ceil(map_m / resolution)
        self.mode = mode
        self.known = known_clouds or {}
        self.store = store
        self.half_res = np.float32(resolution / 2)
        xs, ys = camera_scale_tables(height, width, vfov)
        self.xs, self.ys = xs.numpy(), ys.numpy()
        self.maxB = max_envs
        self._h = lib().emu_create(height, width, _p(self.xs, ctypes.c_float), _p(self.ys, ctypes.c_float),
                                   ctypes.c_float(np.float32(resolution)), ctypes.c_float(self.half_res),
                                   ctypes.c_float(np.float32(map_m / 2)), ctypes.c_float(np.float32(map_m / 2)),
                                   self.R, self.C, store, store, max_envs, tile, tile, 0 if mode == "iterative" else 1,
                                   known_capacity)
        lib().emu_set_order(self._h, order)
        if fix_cap is not None:
            lib().emu_set_fix_cap(self._h, fix_cap)
        lib().emu_set_direct(self._h, 1 if direct else 0)
        self._B = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().emu_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def step(self, masks, pose, orientation, depth=None, labels=None, env_names=None):
        L = lib()
        B = int(masks.shape[0])
        self._B = B
        masks = np.ascontiguousarray(masks, dtype=np.uint8).reshape(B)
        pose = np.ascontiguousarray(pose, dtype=np.float32)
        ori = torch.from_numpy(np.ascontiguousarray(orientation))
        cs = np.ascontiguousarray(ego_rotation(ori[:, 1]).numpy())
        occ = np.zeros((B, self.R, self.C), dtype=np.uint8)
        sem = np.zeros((B, self.R, self.C), dtype=np.uint8)
        if self.mode == "iterative":
            T12 = np.ascontiguousarray(camera_to_world_rows(torch.from_numpy(pose), ori[:, 0], ori[:, 1]).numpy())
            depth = np.ascontiguousarray(depth, dtype=np.float32)
            labels = np.ascontiguousarray(labels, dtype=np.uint8)
            o = np.ascontiguousarray(orientation)
            rc = L.emu_step_iterative(self._h, B, _p(depth, ctypes.c_float), _p(labels, ctypes.c_uint8),
                                      _p(T12, ctypes.c_float), _p(pose, ctypes.c_float), _p(cs, ctypes.c_float),
                                      o.ctypes.data_as(ctypes.c_void_p) if self.kernel_trig else None,
                                      1 if o.dtype == np.float64 else 0,
                                      _p(masks, ctypes.c_uint8), _p(occ, ctypes.c_uint8), _p(sem, ctypes.c_uint8))
            assert rc == 0
        else:
            for b in range(self._known_hi if hasattr(self, "_known_hi") else 0):
                if b >= B:
                    L.emu_known_clear(self._h, b)
            self._known_hi = B
            for b in range(B):
                if masks[b] == 0:
                    xyz, s = self.known[env_names[b]]
                    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
                    s8 = np.ascontiguousarray(np.asarray(s).astype(np.int64).astype(np.uint8))
                    o_r = int(np.floor(xyz[:, 2].min() / self.half_res)) - 2
                    o_c = int(np.floor(xyz[:, 0].min() / self.half_res)) - 2
                    rc = L.emu_known_load(self._h, b, xyz.shape[0], _p(xyz, ctypes.c_float), _p(s8, ctypes.c_uint8), o_r, o_c)
                    assert rc == 0
            L.emu_step_known(self._h, B, _p(pose, ctypes.c_float), _p(cs, ctypes.c_float), _p(occ, ctypes.c_uint8),
                             _p(sem, ctypes.c_uint8))
        return occ, sem

    def status(self):
        err = ctypes.c_uint32(0)
        stats = np.zeros(8, dtype=np.uint64)
        lib().emu_status(self._h, ctypes.byref(err), _p(stats, ctypes.c_uint64))
        return int(err.value), stats

    def set_step(self, step):
        lib().emu_set_step(self._h, ctypes.c_uint(step))

    def set_stamp_period(self, period):
        lib().emu_set_stamp_period(self._h, ctypes.c_uint(period))

    def cand_current(self):
        return int(lib().emu_cand_current(self._h))

    def world(self):
        """Live records in the reference's list order (sorted by the reference key)."""
        cap = 1 << 22
        env = np.zeros(cap, dtype=np.int64)
        xyz = np.zeros((cap, 3), dtype=np.float32)
        lab = np.zeros(cap, dtype=np.uint8)
        key = np.zeros(cap, dtype=np.uint64)
        n = int(lib().emu_export_world(self._h, self._B, cap, _p(env, ctypes.c_int64), _p(xyz, ctypes.c_float),
                                       _p(lab, ctypes.c_uint8), _p(key, ctypes.c_uint64)))
        assert n <= cap
        order = np.argsort(key[:n], kind="stable")
        return env[:n][order], xyz[:n][order], lab[:n][order]
