"""Seeded synthetic inputs for the map-update path.

Everything is generated on the host with numpy so that the same bytes can be
fed to the reference (in the build container), to the CPU oracle and to the
CUDA path.  Layouts are the ones the habitat sensors deliver to the reference
(`habitat_extensions/sensors.py:161-367`): depth f32 [B,H,W,1] in [0,1],
semantic12 u8 [B,H,W,1], world_robot_pose f32 [B,3], world_robot_orientation
(elevation, heading) f64 [B,2], not_done_masks u8 [B,1], env_name list[str].

Two depth modes:
  * "iid"   - U(0.05, 0.95) per pixel (no height ties, ~23 % of pixels survive
              the depth/height filters),
  * "scene" - an analytic box room with box obstacles, ray-cast from the pose
              (coherent surfaces, exact height ties, shared cells).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

DEPTH_SCALE = 10.0  # reference mapper.py:381-384


@dataclass
class ScenarioConfig:
    name: str = "scenario"
    num_envs: int = 2
    height: int = 64
    width: int = 64
    steps: int = 8
    vfov_radians: float = math.pi / 2.0
    map_meters: float = 6.4
    resolution: float = 0.1
    num_labels: int = 13
    depth_mode: str = "iid"  # "iid" | "scene"
    angle_dtype: str = "float64"  # dtype of world_robot_orientation
    env_spacing: float = 40.0  # start offset between envs (0 => overlapping)
    reset_steps: Dict[int, List[int]] = field(default_factory=dict)  # step -> env list
    reset_every: Optional[int] = None  # episodic: all envs reset every k steps
    forward_step: float = 0.25
    turn_degrees: float = 15.0
    sensor_height: float = 1.25
    elevation: float = 0.0
    roam_radius: Optional[float] = None  # keep the walk within this distance of its start (a house-sized area)
    seed: int = 1000


def camera_tables(height: int, width: int, vfov: float):
    """f32 per-column / per-row scale tables, same arithmetic as
    projector/core.py:70-107 (fx, cx are doubles cast to f32; (u + 0.5 - cx) / fx
    evaluated in f32)."""
    hfov = width / height * vfov
    fx = np.float32(width / (2.0 * math.tan(hfov / 2.0)))
    fy = np.float32(height / (2.0 * math.tan(vfov / 2.0)))
    cx = np.float32(width / 2.0)
    cy = np.float32(height / 2.0)
    half = np.float32(0.5)
    xs = ((np.arange(width, dtype=np.float32) + half) - cx) / fx
    ys = ((np.arange(height, dtype=np.float32) + half) - cy) / fy
    return xs.astype(np.float32), ys.astype(np.float32)


def _rotation(elevation: float, heading: float) -> np.ndarray:
    """3x3 part of projector/core.py:6-37 with elevation + pi (mapper.py:132-138),
    in float64 (used only to build synthetic scenes, not for parity)."""
    ex = elevation + math.pi
    cx, sx = math.cos(ex), math.sin(ex)
    cy, sy = math.cos(heading), math.sin(heading)
    return np.array(
        [[cy, sx * sy, cx * sy], [0.0, cx, -sx], [-sy, cy * sx, cy * cx]]
    )


class BoxRoom:
    """Axis-aligned room with box obstacles; labelled surfaces."""

    def __init__(self, rng: np.random.Generator, num_labels: int, half=(5.0, 6.0), n_boxes=6):
        self.half = half
        self.ceiling = 2.6
        k = max(num_labels - 1, 1)
        # boxes: (xmin, ymin, zmin, xmax, ymax, zmax, label)
        boxes = []
        for i in range(n_boxes):
            cx = rng.uniform(-half[0] + 1.0, half[0] - 1.0)
            cz = rng.uniform(-half[1] + 1.0, half[1] - 1.0)
            sx = rng.uniform(0.3, 1.2)
            sz = rng.uniform(0.3, 1.2)
            hy = rng.choice([0.4, 0.8, 1.2, 1.6, 2.2])
            boxes.append((cx - sx, 0.0, cz - sz, cx + sx, hy, cz + sz, 1 + (i % k)))
        self.boxes = np.array(boxes, dtype=np.float64)
        self.wall_label = 1 + (n_boxes % k)
        self.floor_label = 0  # label 0 is dropped from the semantic map
        self.ceiling_label = 1 + ((n_boxes + 1) % k)

    def inside_free_space(self, x, z, margin=0.35):
        if abs(x) > self.half[0] - margin or abs(z) > self.half[1] - margin:
            return False
        b = self.boxes
        hit = (
            (x > b[:, 0] - margin) & (x < b[:, 3] + margin)
            & (z > b[:, 2] - margin) & (z < b[:, 5] + margin)
        )
        return not bool(hit.any())

    def raycast(self, origin: np.ndarray, dirs: np.ndarray):
        """origin [3], dirs [N,3] (per unit camera-z).  Returns (t, label):
        smallest positive parameter t with origin + t*dir on a surface."""
        n = dirs.shape[0]
        best = np.full(n, np.inf)
        label = np.zeros(n, dtype=np.int64)
        with np.errstate(divide="ignore", invalid="ignore"):
            # room shell: 6 planes, inside-out
            planes = [
                (0, -self.half[0], self.wall_label), (0, self.half[0], self.wall_label),
                (2, -self.half[1], self.wall_label), (2, self.half[1], self.wall_label),
                (1, 0.0, self.floor_label), (1, self.ceiling, self.ceiling_label),
            ]
            for axis, val, lab in planes:
                t = (val - origin[axis]) / dirs[:, axis]
                ok = (t > 1e-6) & (t < best)
                best = np.where(ok, t, best)
                label = np.where(ok, lab, label)
            for bx in self.boxes:
                lo = (bx[0:3] - origin) / dirs
                hi = (bx[3:6] - origin) / dirs
                tmin = np.minimum(lo, hi).max(axis=1)
                tmax = np.maximum(lo, hi).min(axis=1)
                ok = (tmax >= tmin) & (tmin > 1e-6) & (tmin < best)
                best = np.where(ok, tmin, best)
                label = np.where(ok, int(bx[6]), label)
        return best, label


def random_walk(cfg: ScenarioConfig, rng: np.random.Generator, rooms=None):
    """pose f32 [T,B,3], orientation [T,B,2] (elevation, heading) in cfg.angle_dtype."""
    T, B = cfg.steps, cfg.num_envs
    pose = np.zeros((T, B, 3), dtype=np.float64)
    heading = np.zeros((T, B), dtype=np.float64)
    turn = math.radians(cfg.turn_degrees)
    for b in range(B):
        ox = cfg.env_spacing * b
        oz = -cfg.env_spacing * 0.5 * b
        x, z = rng.uniform(-1.0, 1.0), rng.uniform(-1.0, 1.0)
        h = rng.uniform(-math.pi, math.pi)
        room = rooms[b] if rooms is not None else None
        if room is not None:
            for _ in range(200):
                if room.inside_free_space(x, z):
                    break
                x = rng.uniform(-room.half[0] + 0.5, room.half[0] - 0.5)
                z = rng.uniform(-room.half[1] + 0.5, room.half[1] - 0.5)
        x0, z0 = x, z
        for t in range(T):
            pose[t, b] = (x + ox, cfg.sensor_height, z + oz)
            heading[t, b] = h
            a = rng.integers(0, 4)
            if a <= 1:
                nx = x - math.sin(h) * cfg.forward_step
                nz = z - math.cos(h) * cfg.forward_step
                inside = True
                if cfg.roam_radius is not None:
                    inside = (nx - x0) ** 2 + (nz - z0) ** 2 <= cfg.roam_radius ** 2
                if inside and (room is None or room.inside_free_space(nx, nz)):
                    x, z = nx, nz
                else:
                    h += 2 * turn
            elif a == 2:
                h += turn
            else:
                h -= turn
            h = (h + math.pi) % (2 * math.pi) - math.pi
    orient = np.zeros((T, B, 2), dtype=np.float64)
    orient[..., 0] = cfg.elevation
    orient[..., 1] = heading
    return pose.astype(np.float32), orient.astype(np.dtype(cfg.angle_dtype))


def reset_masks(cfg: ScenarioConfig) -> np.ndarray:
    """not_done_masks u8 [T,B]: 0 = episode/tour finished => wipe that env's
    world state before ingesting this frame (mapper.py:320-326)."""
    m = np.ones((cfg.steps, cfg.num_envs), dtype=np.uint8)
    m[0, :] = 0
    if cfg.reset_every:
        m[:: cfg.reset_every, :] = 0
    for t, envs in cfg.reset_steps.items():
        for b in envs:
            m[int(t), int(b)] = 0
    return m


def make_scenario(cfg: ScenarioConfig) -> Dict[str, np.ndarray]:
    """All inputs of a run: depth [T,B,H,W] f32, labels [T,B,H,W] u8,
    pose [T,B,3] f32, orientation [T,B,2], masks [T,B] u8."""
    rng = np.random.default_rng(cfg.seed)
    T, B, H, W = cfg.steps, cfg.num_envs, cfg.height, cfg.width
    rooms = None
    if cfg.depth_mode == "scene":
        shared = cfg.env_spacing == 0.0
        first = BoxRoom(rng, cfg.num_labels)
        rooms = [first if shared else BoxRoom(rng, cfg.num_labels) for _ in range(B)]
        if not shared:
            rooms[0] = first
    pose, orient = random_walk(cfg, rng, rooms)
    masks = reset_masks(cfg)
    if cfg.depth_mode == "iid":
        depth = rng.uniform(0.05, 0.95, size=(T, B, H, W)).astype(np.float32)
        labels = rng.integers(0, cfg.num_labels, size=(T, B, H, W), dtype=np.uint8)
    elif cfg.depth_mode == "scene":
        xs, ys = camera_tables(H, W, cfg.vfov_radians)
        dcam = np.stack(
            [np.broadcast_to(xs[None, :], (H, W)), np.broadcast_to(ys[:, None], (H, W)),
             np.ones((H, W), dtype=np.float32)], axis=-1,
        ).reshape(-1, 3).astype(np.float64)
        depth = np.zeros((T, B, H, W), dtype=np.float32)
        labels = np.zeros((T, B, H, W), dtype=np.uint8)
        for b in range(B):
            ox = cfg.env_spacing * b
            oz = -cfg.env_spacing * 0.5 * b
            for t in range(T):
                R = _rotation(float(orient[t, b, 0]), float(orient[t, b, 1]))
                origin = pose[t, b].astype(np.float64) - np.array([ox, 0.0, oz])
                dw = dcam @ R.T
                tz, lab = rooms[b].raycast(origin, dw)
                d = np.clip(tz / DEPTH_SCALE, 0.0, 1.0)
                depth[t, b] = d.reshape(H, W).astype(np.float32)
                labels[t, b] = lab.reshape(H, W).astype(np.uint8)
    else:
        raise ValueError(cfg.depth_mode)
    return {"depth": depth, "labels": labels, "pose": pose, "orientation": orient, "masks": masks}


def make_logits(labels_shape, num_classes: int, seed: int) -> np.ndarray:
    """f32 N(0,1) class scores [.., Cls, H, W] (NCHW per frame), config 2."""
    rng = np.random.default_rng(seed)
    lead, (H, W) = labels_shape[:-2], labels_shape[-2:]
    return rng.standard_normal(size=(*lead, num_classes, H, W), dtype=np.float32)


def make_known_cloud(num_points: int, extent_m: float, num_labels: int, seed: int, center=(0.0, 0.0)):
    """npz-style scene cloud for known-map mode (mapper.py:283-294): xyz f32 [N,3]
    uniform over extent x 3 m x extent, semantics int64 [N]."""
    rng = np.random.default_rng(seed)
    xyz = np.empty((num_points, 3), dtype=np.float32)
    xyz[:, 0] = rng.uniform(center[0] - extent_m / 2, center[0] + extent_m / 2, num_points)
    xyz[:, 1] = rng.uniform(0.0, 3.0, num_points)
    xyz[:, 2] = rng.uniform(center[1] - extent_m / 2, center[1] + extent_m / 2, num_points)
    sem = rng.integers(0, num_labels, num_points).astype(np.int64)
    return xyz, sem


def obs_dict_for_step(scn: Dict[str, np.ndarray], t: int, num_envs: Optional[int] = None,
                      env_names: Optional[List[str]] = None):
    """One step in the obs-dict layout `Mapper.forward` receives
    (setup_mapping_module.py:56-89), as torch CPU tensors."""
    import torch

    B = scn["depth"].shape[1] if num_envs is None else num_envs
    names = env_names if env_names is not None else [f"scene{b}" for b in range(B)]
    return {
        "depth": torch.from_numpy(scn["depth"][t, :B]).unsqueeze(-1).contiguous(),
        "semantic12": torch.from_numpy(scn["labels"][t, :B]).unsqueeze(-1).contiguous(),
        "world_robot_pose": torch.from_numpy(scn["pose"][t, :B]).contiguous(),
        "world_robot_orientation": torch.from_numpy(scn["orientation"][t, :B]).contiguous(),
        "not_done_masks": torch.from_numpy(scn["masks"][t, :B]).unsqueeze(-1).contiguous(),
        "env_name": list(names[:B]),
    }
