"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's PyTorch CPU path, restated.

`TorchReferencePath` performs, with eager torch CPU ops, the same sequence of tensor
operations as the reference `MappingModule.forward` (mapper.py:904-947): every stage
materialises the same temporaries, the world state is the same growing point list, the
occupancy and semantic rasters repeat the same transform twice.  It exists so that
bench.py can time "the reference's PyTorch CPU path" on the GPU box's host cores
(`cpu_baseline`, `--impl reference`; /root/reference itself does not exist there) and as
a second, independent check of the C oracle (tests/test_torch_path.py).  It is a port:
`scatter_max` is the torch-only stand-in described in oracle/ref_loader.py, because
torch-scatter cannot be installed offline.  Never imported by the product package.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch


def _scatter_max_arg(src: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """argmax part of torch_scatter.scatter_max (first index among the maxima, N for empty groups)."""
    n = src.shape[0]
    groups = int(index.max()) + 1
    out = torch.full((groups,), torch.finfo(src.dtype).min, dtype=src.dtype)
    out = out.scatter_reduce(0, index, src, "amax", include_self=True)
    pos = torch.arange(n)
    cand = torch.where(src == out[index], pos, torch.full_like(pos, n))
    return torch.full((groups,), n, dtype=torch.long).scatter_reduce(0, index, cand, "amin", include_self=True)


class TorchReferencePath:
    def __init__(self, height: int, width: int, vfov: float, map_height_m: float, map_width_m: float,
                 resolution: float, known_clouds: Optional[Dict[str, Tuple]] = None):
        self.H, self.W = height, width
        self.res = resolution
        self.mh, self.mw = map_height_m, map_width_m
        self.R = math.ceil(map_height_m / resolution)
        self.C = math.ceil(map_width_m / resolution)
        self.known = known_clouds
        # projector/core.py:70-107
        hfov = width / height * vfov
        K = torch.Tensor([[width / (2.0 * math.tan(hfov / 2.0)), 0, width / 2.0],
                          [0, height / (2.0 * math.tan(vfov / 2.0)), height / 2.0], [0, 0, 1.0]])
        u = torch.arange(width).unsqueeze(0).repeat(height, 1).float()
        v = torch.arange(height).unsqueeze(1).repeat(1, width).float()
        self.x_scale = ((u + 0.5 - K[0, 2]) / K[0, 0]).unsqueeze(0)
        self.y_scale = ((v + 0.5 - K[1, 2]) / K[1, 1]).unsqueeze(0)
        self.w_b = self.w_xyz = self.w_sem = None  # world point list

    # ---- stages
    @staticmethod
    def _camera_matrix(pose, elevation, heading):
        ex = elevation + torch.pi
        cx, sx, cy, sy = torch.cos(ex), torch.sin(ex), torch.cos(heading), torch.sin(heading)
        T = torch.zeros(pose.shape[0], 4, 4)
        T[:, 0, 0] = cy; T[:, 0, 1] = sx * sy; T[:, 0, 2] = cx * sy; T[:, 0, 3] = pose[:, 0]
        T[:, 1, 1] = cx; T[:, 1, 2] = -sx; T[:, 1, 3] = pose[:, 1]
        T[:, 2, 0] = -sy; T[:, 2, 1] = cy * sx; T[:, 2, 2] = cy * cx; T[:, 2, 3] = pose[:, 2]
        T[:, 3, 3] = 1
        return T

    def _frame_cloud(self, depth, labels, pose, T):
        B, H, W = depth.shape[0], self.H, self.W
        z = (depth[:, 0] * 10) / 1.0                                  # mapper.py:381-384, core.py:142
        x, y = z * self.x_scale, z * self.y_scale
        xyz1 = torch.cat((x.unsqueeze(3), y.unsqueeze(3), z.unsqueeze(3), torch.ones(B, H, W, 1)), dim=3)
        xyz1 = xyz1.reshape(B, H * W, 4).transpose(1, 2)
        world = torch.bmm(T, xyz1).transpose(1, 2)[:, :, :3]          # core.py:171
        world = world - torch.zeros(3)
        xyz = world.reshape(B, H, W, 3).permute(0, 3, 1, 2)           # point_cloud.py:82
        bidx = torch.arange(B).view(-1, 1, 1, 1).expand(B, 1, H, W)
        flat = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).squeeze()   # mapper.py:32-35
        b, p, s, d = flat(bidx), flat(xyz), flat(labels), flat(depth)
        keep = torch.logical_and(d > 0.01, d < 0.99)                 # mapper.py:236-243
        b, p, s = b[keep], p[keep], s[keep]
        h = pose[:, 1][b]
        keep = torch.logical_and(p[:, 1] > (h - 1.0), p[:, 1] < (h + 0.5))   # mapper.py:245-253
        return b[keep], p[keep], s[keep]

    def _keep_highest(self, b, p, s):
        if p.shape[0] == 0:
            return b, p, s
        half = self.res / 2
        rows = (p[:, 2] / half).round().long(); rows = rows - rows.min()   # mapper.py:461-466
        cols = (p[:, 0] / half).round().long(); cols = cols - cols.min()
        flat = b * (rows.max() * cols.max()) + rows * cols.max() + cols    # mapper.py:468-469
        arg = _scatter_max_arg(p[:, 1], flat)
        arg = arg[arg != flat.shape[0]]
        return b[arg], p[arg], s[arg]

    def _raster(self, B, b, p, s, pose, heading, semantic: bool):
        p = p.clone(); b = b.clone(); s = s.clone()                    # .copy() = deepcopy, mapper.py:562
        p = p + (-pose)[b]                                             # translate
        a = -heading
        M = torch.zeros(a.shape[0], 3, 3)
        M[:, 0, 0] = torch.cos(a); M[:, 0, 2] = torch.sin(a); M[:, 1, 1] = 1
        M[:, 2, 0] = -torch.sin(a); M[:, 2, 2] = torch.cos(a)
        p = torch.bmm(M[b], p.unsqueeze(-1)).squeeze(-1)               # mapper.py:258-262
        rows = ((p[:, 2] + self.mh / 2) / self.res).round().long()     # mapper.py:101-114
        cols = ((p[:, 0] + self.mw / 2) / self.res).round().long()
        ok = torch.logical_and(torch.logical_and(rows >= 0, rows < self.R), torch.logical_and(cols >= 0, cols < self.C))
        b, rows, cols, s = b[ok], rows[ok], cols[ok], s[ok]
        n_in = int(b.shape[0])
        data = torch.zeros(B, self.R, self.C, dtype=torch.uint8)
        if semantic:
            nz = s != 0                                                # mapper.py:611
            b, rows, cols, s = b[nz], rows[nz], cols[nz], s[nz]
            data[b, rows, cols] = s
        else:
            data[b, rows, cols] = 1
        return data, n_in

    # ---- one MappingModule.forward
    @torch.no_grad()
    def step(self, masks, pose, elevation, heading, depth=None, labels=None, scores=None,
             env_names: Optional[Sequence[str]] = None):
        """masks u8 [B]; pose f32 [B,3]; depth f32 [B,1,H,W]; labels u8 [B,1,H,W] or scores f32 [B,Cls,H,W]."""
        B = masks.shape[0]
        if scores is not None:
            labels = scores.argmax(1, keepdims=True).to(torch.uint8)   # mapper.py:795-798
        self.labels = labels
        if self.w_xyz is not None:                                     # mapper.py:310-326
            keep = self.w_b < B
            for fi in torch.arange(B)[masks == 0]:
                keep = torch.logical_and(keep, self.w_b != fi)
            self.w_b, self.w_xyz, self.w_sem = self.w_b[keep], self.w_xyz[keep], self.w_sem[keep]
        if self.known is None:
            T = self._camera_matrix(pose, elevation, heading)
            lb, lp, ls = self._keep_highest(*self._frame_cloud(depth, labels, pose, T))
            if self.w_xyz is None:
                self.w_b, self.w_xyz, self.w_sem = lb, lp, ls
            else:
                self.w_b, self.w_xyz, self.w_sem = (torch.cat((self.w_b, lb)), torch.cat((self.w_xyz, lp)),
                                                    torch.cat((self.w_sem, ls)))
            self.w_b, self.w_xyz, self.w_sem = self._keep_highest(self.w_b, self.w_xyz, self.w_sem)
        else:
            for bi in torch.arange(B)[masks == 0].tolist():
                xyz, sem = self.known[env_names[bi]]
                xyz = torch.as_tensor(xyz, dtype=torch.float32)
                sem = torch.as_tensor(sem).long().to(torch.uint8)
                nb = torch.full((xyz.shape[0],), bi, dtype=torch.long)
                if self.w_xyz is None:
                    self.w_b, self.w_xyz, self.w_sem = nb, xyz, sem
                else:
                    self.w_b, self.w_xyz, self.w_sem = (torch.cat((self.w_b, nb)), torch.cat((self.w_xyz, xyz)),
                                                        torch.cat((self.w_sem, sem)))
        if self.w_xyz is None:
            z = torch.zeros(B, self.R, self.C, dtype=torch.uint8)
            return z, z.clone()
        b, p, s = self.w_b.clone(), self.w_xyz.clone(), self.w_sem.clone()   # mapper.py:895
        h = pose[:, 1][b]
        band = torch.logical_and(p[:, 1] > (h - 1.25), p[:, 1] < (h + 0.75))
        b, p, s = b[band], p[band], s[band]
        occ, self.n_in = self._raster(B, b, p, s, pose, heading, semantic=False)
        sem, _ = self._raster(B, b, p, s, pose, heading, semantic=True)
        return occ, sem
