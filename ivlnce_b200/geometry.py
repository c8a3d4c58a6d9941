"""Host-side geometry of the map update: the tiny per-env matrices the kernels
consume.  All trigonometry stays here, in torch, in the dtype the angles arrive
in (float64 from the habitat sensors, `habitat_extensions/sensors.py:240-250`),
and is rounded to float32 exactly where the reference stores it into float32
tensors -- the kernels never recompute trig ("Trig").
"""
from __future__ import annotations

import math
from typing import Tuple

import torch


def camera_scale_tables(height: int, width: int, vertical_fov_radians: float,
                        device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """x_scale[u] = (u + 0.5 - cx) / fx and y_scale[v] = (v + 0.5 - cy) / fy as f32.

    Follows reference projector/core.py:70-107: the intrinsics are Python doubles
    stored in an f32 tensor (hfov = W/H * vfov, fx = W / (2 tan(hfov/2))), the
    pixel index is converted to f32 and the three operations are f32.
    """
    hfov = width / height * vertical_fov_radians
    fx = torch.tensor(width / (2.0 * math.tan(hfov / 2.0)), dtype=torch.float32)
    fy = torch.tensor(height / (2.0 * math.tan(vertical_fov_radians / 2.0)), dtype=torch.float32)
    cx = torch.tensor(width / 2.0, dtype=torch.float32)
    cy = torch.tensor(height / 2.0, dtype=torch.float32)
    xs = (torch.arange(width, dtype=torch.float32) + 0.5 - cx) / fx
    ys = (torch.arange(height, dtype=torch.float32) + 0.5 - cy) / fy
    return xs.to(device).contiguous(), ys.to(device).contiguous()


def camera_to_world_rows(pose: torch.Tensor, elevation: torch.Tensor, heading: torch.Tensor) -> torch.Tensor:
    """f32 [B,12]: rows 0..2 of the camera->world matrix Rx(elevation + pi) * Ry(heading) | pose.

    Same element formulas as reference projector/core.py:6-37, called as in
    mapper.py:132-138 (elevation + pi).  Products are formed in the angles' dtype
    and rounded to f32 on assignment, as the reference's `T[:, i, j] = ...` does.
    """
    ex = elevation + torch.pi
    cx, sx = torch.cos(ex), torch.sin(ex)
    cy, sy = torch.cos(heading), torch.sin(heading)
    T = torch.zeros(pose.shape[0], 12, dtype=torch.float32, device=pose.device)
    T[:, 0] = cy
    T[:, 1] = sx * sy
    T[:, 2] = cx * sy
    T[:, 3] = pose[:, 0]
    # T[:, 4] = 0
    T[:, 5] = cx
    T[:, 6] = -sx
    T[:, 7] = pose[:, 1]
    T[:, 8] = -sy
    T[:, 9] = cy * sx
    T[:, 10] = cy * cx
    T[:, 11] = pose[:, 2]
    return T


def ego_rotation(heading: torch.Tensor) -> torch.Tensor:
    """f32 [B,2] = (cos(-heading), sin(-heading)): the two distinct entries of the
    rotate-about-y matrix of mapper.py:38-48 as used by shift_origin (mapper.py:264-266)."""
    a = -heading
    cs = torch.zeros(heading.shape[0], 2, dtype=torch.float32, device=heading.device)
    cs[:, 0] = torch.cos(a)
    cs[:, 1] = torch.sin(a)
    return cs
