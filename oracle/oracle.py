"""TEST INFRASTRUCTURE ONLY -- python front end of the C oracle.

`OracleMapper` drives oracle/mapper_oracle.c (compiled on demand with gcc into
oracle/_build/liboracle.so) through the same per-step sequence as the
reference's `MappingModule.forward` (mapper.py:921-944).  Parity status: the
reference ships no golden vectors for this path; the oracle is pinned against
outputs of the unmodified reference code generated in the build container
(tests/golden/).  Only tests/, `__graft_entry__.smoke()` and bench.py's
cpu_baseline / `--impl reference` legs may import this module.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "liboracle.so")
_SRC = os.path.join(_HERE, "mapper_oracle.c")


def build_oracle(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off (no fast-math): IEEE fp32, fma only where written."""
    os.makedirs(_BUILD, exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
               "-Wall", "-o", _SO, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
            build_oracle()
        L = ctypes.CDLL(_SO)
        f32p = ctypes.POINTER(ctypes.c_float)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.c_int, ctypes.c_int, f32p, f32p, ctypes.c_float, ctypes.c_float,
                                 ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_clear.argtypes = [ctypes.c_void_p, ctypes.c_int, u8p]
        L.orc_ingest.argtypes = [ctypes.c_void_p, ctypes.c_int, f32p, u8p, f32p, f32p]
        L.orc_append_cloud.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, f32p, u8p]
        L.orc_raster.argtypes = [ctypes.c_void_p, ctypes.c_int, f32p, f32p]
        L.orc_argmax_labels.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, u8p]
        L.orc_occupancy.restype = u8p
        L.orc_occupancy.argtypes = [ctypes.c_void_p]
        L.orc_semantic.restype = u8p
        L.orc_semantic.argtypes = [ctypes.c_void_p]
        L.orc_world_size.restype = ctypes.c_int64
        L.orc_world_size.argtypes = [ctypes.c_void_p]
        L.orc_world_export.argtypes = [ctypes.c_void_p, i64p, f32p, u8p]
        L.orc_counters.argtypes = [ctypes.c_void_p, i64p]
        _lib = L
    return _lib


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


# --------------------------------------------------------------- host geometry
def camera_tables(height: int, width: int, vfov: float) -> Tuple[np.ndarray, np.ndarray]:
    """x_scale[u], y_scale[v] of projector/core.py:70-107: fx, fy, cx, cy are python
    doubles stored into an f32 tensor, then (u + 0.5 - cx) / fx in f32."""
    hfov = width / height * vfov
    fx = np.float32(width / (2.0 * math.tan(hfov / 2.0)))
    fy = np.float32(height / (2.0 * math.tan(vfov / 2.0)))
    cx = np.float32(width / 2.0)
    cy = np.float32(height / 2.0)
    xs = ((np.arange(width, dtype=np.float32) + np.float32(0.5)) - cx) / fx
    ys = ((np.arange(height, dtype=np.float32) + np.float32(0.5)) - cy) / fy
    return np.ascontiguousarray(xs, dtype=np.float32), np.ascontiguousarray(ys, dtype=np.float32)


def camera_to_world(pose: torch.Tensor, elevation: torch.Tensor, heading: torch.Tensor) -> np.ndarray:
    """f32 [B,16]: projector/core.py:6-37 called with elevation + pi
    (mapper.py:132-138).  Trig and products are evaluated in the angles' dtype and
    rounded to f32 when stored."""
    ex = elevation + torch.pi
    cx, sx = torch.cos(ex), torch.sin(ex)
    cy, sy = torch.cos(heading), torch.sin(heading)
    T = torch.zeros(pose.shape[0], 4, 4, dtype=torch.float32)
    T[:, 0, 0] = cy
    T[:, 0, 1] = sx * sy
    T[:, 0, 2] = cx * sy
    T[:, 0, 3] = pose[:, 0]
    T[:, 1, 1] = cx
    T[:, 1, 2] = -sx
    T[:, 1, 3] = pose[:, 1]
    T[:, 2, 0] = -sy
    T[:, 2, 1] = cy * sx
    T[:, 2, 2] = cy * cx
    T[:, 2, 3] = pose[:, 2]
    T[:, 3, 3] = 1
    return np.ascontiguousarray(T.reshape(-1, 16).numpy())


def ego_rotation(heading: torch.Tensor) -> np.ndarray:
    """f32 [B,2] = (cos(-heading), sin(-heading)) as stored by rotate_around_y_matrix
    (mapper.py:38-48) for shift_origin (mapper.py:264-266)."""
    a = -heading
    cs = torch.zeros(heading.shape[0], 2, dtype=torch.float32)
    cs[:, 0] = torch.cos(a)
    cs[:, 1] = torch.sin(a)
    return np.ascontiguousarray(cs.numpy())


class OracleMapper:
    """CPU oracle with the call sequence of `MappingModule.forward`.

    mode = "iterative" (depth + labels ingested every step) or "known" (scene
    clouds appended on reset, `known_clouds[env_name] = (xyz f32 [N,3], sem [N])`).
    """

    def __init__(self, height: int, width: int, vfov: float, map_height_m: float, map_width_m: float,
                 resolution: float, mode: str = "iterative",
                 known_clouds: Optional[Dict[str, Tuple[np.ndarray, np.ndarray]]] = None):
        self.H, self.W = int(height), int(width)
        self.R = math.ceil(map_height_m / resolution)  # mapper.py:97-99
        self.C = math.ceil(map_width_m / resolution)
        self.mode = mode
        self.known = known_clouds or {}
        xs, ys = camera_tables(self.H, self.W, vfov)
        self.xs, self.ys = xs, ys
        self._h = lib().orc_create(
            self.H, self.W, _p(xs, ctypes.c_float), _p(ys, ctypes.c_float),
            ctypes.c_float(np.float32(resolution)), ctypes.c_float(np.float32(resolution / 2)),
            ctypes.c_float(np.float32(map_height_m / 2)), ctypes.c_float(np.float32(map_width_m / 2)),
            self.R, self.C)
        self.counters = {}

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def step(self, masks: np.ndarray, pose: np.ndarray, orientation: np.ndarray,
             depth: Optional[np.ndarray] = None, labels: Optional[np.ndarray] = None,
             env_names: Optional[Sequence[str]] = None):
        """masks u8 [B]; pose f32 [B,3]; orientation [B,2] (elevation, heading) f32|f64;
        depth f32 [B,H,W]; labels u8 [B,H,W].  Returns (occupancy, semantic) u8 [B,R,C]."""
        L = lib()
        B = int(masks.shape[0])
        masks = np.ascontiguousarray(masks, dtype=np.uint8).reshape(B)
        pose = np.ascontiguousarray(pose, dtype=np.float32)
        ori = torch.from_numpy(np.ascontiguousarray(orientation))
        pose_t = torch.from_numpy(pose)
        L.orc_clear(self._h, B, _p(masks, ctypes.c_uint8))
        if self.mode == "iterative":
            T = camera_to_world(pose_t, ori[:, 0], ori[:, 1])
            depth = np.ascontiguousarray(depth, dtype=np.float32).reshape(B, self.H, self.W)
            labels = np.ascontiguousarray(labels, dtype=np.uint8).reshape(B, self.H, self.W)
            L.orc_ingest(self._h, B, _p(depth, ctypes.c_float), _p(labels, ctypes.c_uint8),
                         _p(T, ctypes.c_float), _p(pose, ctypes.c_float))
        else:
            for b in range(B):
                if masks[b] == 0:
                    xyz, sem = self.known[env_names[b]]
                    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
                    sem8 = np.ascontiguousarray(np.asarray(sem).astype(np.int64).astype(np.uint8))
                    L.orc_append_cloud(self._h, b, xyz.shape[0], _p(xyz, ctypes.c_float), _p(sem8, ctypes.c_uint8))
        cs = ego_rotation(ori[:, 1])
        L.orc_raster(self._h, B, _p(pose, ctypes.c_float), _p(cs, ctypes.c_float))
        n = B * self.R * self.C
        occ = np.ctypeslib.as_array(L.orc_occupancy(self._h), shape=(n,)).reshape(B, self.R, self.C).copy()
        sem = np.ctypeslib.as_array(L.orc_semantic(self._h), shape=(n,)).reshape(B, self.R, self.C).copy()
        c = np.zeros(6, dtype=np.int64)
        L.orc_counters(self._h, _p(c, ctypes.c_int64))
        self.counters = dict(zip(("n_valid", "n_local", "n_world", "n_band", "n_in", "n_ties"), c.tolist()))
        return occ, sem

    def world(self):
        """(batch_indices i64 [P], xyz f32 [P,3], semantics u8 [P]) in world-cloud order."""
        L = lib()
        n = int(L.orc_world_size(self._h))
        b = np.zeros(max(n, 1), dtype=np.int64)
        xyz = np.zeros((max(n, 1), 3), dtype=np.float32)
        sem = np.zeros(max(n, 1), dtype=np.uint8)
        if n:
            L.orc_world_export(self._h, _p(b, ctypes.c_int64), _p(xyz, ctypes.c_float), _p(sem, ctypes.c_uint8))
        return b[:n], xyz[:n], sem[:n]


def argmax_labels(scores: np.ndarray) -> np.ndarray:
    """scores f32 [B,Cls,H,W] -> u8 [B,H,W] (mapper.py:795-798)."""
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    B, K, H, W = scores.shape
    out = np.zeros((B, H, W), dtype=np.uint8)
    lib().orc_argmax_labels(_p(scores, ctypes.c_float), B, K, H * W, _p(out, ctypes.c_uint8))
    return out


def map_features_oracle(occupancy: np.ndarray, semantic: np.ndarray, num_classes: int = 13) -> np.ndarray:
    """CPU restatement of SemanticMapEncoder.generate_map_features
    (ivlnce_baselines/models/encoders/map_encoder.py:85-90): cat(occupancy, one_hot(semantic)) as float32
    [B, 1 + K, R, C].  Raises like F.one_hot if a class value is out of range.  Test infrastructure only."""
    occupancy = np.asarray(occupancy)
    semantic = np.asarray(semantic).astype(np.int64)
    if semantic.size and semantic.max() >= num_classes:
        raise RuntimeError("Class values must be smaller than num_classes.")
    B, R, C = occupancy.shape
    out = np.zeros((B, 1 + num_classes, R, C), dtype=np.float32)
    out[:, 0] = occupancy.astype(np.float32)
    for k in range(num_classes):
        out[:, 1 + k] = (semantic == k).astype(np.float32)
    return out
