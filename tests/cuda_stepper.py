"""Drives the CUDA mapping module with the stepper signature used by tests/scenarios.run_mapper."""
from __future__ import annotations

import os
import tempfile

import numpy as np
import torch

from ivlnce_b200.mapper import (CameraParameters, EpisodesInfo, MapDimensions, Observations, PrecomputedScores,
                                RobotCurrentState, create_gt_semantics_iterative_mapper, create_iterative_mapper,
                                create_known_mapper)


class CudaStepper:
    def __init__(self, cfg, known_clouds=None, device="cuda:0", host_trig=True, pred=False, store_cells=None,
                 max_envs=None, raster_tile=0, trig=None, scatter_variant=0, stamp_period=0):
        self.dev = torch.device(device)
        self.cfg = cfg
        self.pred = pred
        md = MapDimensions(cfg["map_m"], cfg["map_m"], cfg["resolution"])
        if store_cells is None:
            store_cells = 2048 if cfg["resolution"] < 0.1 else 1024
        kw = dict(store_cells=store_cells, max_envs=max_envs, raster_tile=raster_tile, scatter_variant=scatter_variant, stamp_period=stamp_period,
                  trig=trig if trig is not None else ("host" if host_trig else "torch"))
        if cfg["mode"] == "iterative":
            cam = CameraParameters(cfg["vfov"], (cfg["height"], cfg["width"]), 0.1)
            if pred:
                self.mm = create_iterative_mapper(self.dev, cam, md, PrecomputedScores(), **kw)
            else:
                self.mm = create_gt_semantics_iterative_mapper(self.dev, cam, md, **kw)
        else:
            self._tmp = tempfile.TemporaryDirectory()
            for name, (xyz, sem) in (known_clouds or {}).items():
                np.savez(os.path.join(self._tmp.name, f"{name}.npz"), xyz=xyz, semantics=sem)
            self.mm = create_known_mapper(self.dev, md, self._tmp.name, known_capacity=1 << 16, **kw)
        self.last_obs = None

    def step(self, masks, pose, orientation, depth=None, labels=None, env_names=None, logits=None):
        B = masks.shape[0]
        dev = self.dev
        names = list(env_names) if env_names is not None else [f"scene{b}" for b in range(B)]
        ei = EpisodesInfo(torch.from_numpy(np.ascontiguousarray(masks)).reshape(B, 1).to(dev), names)
        ori = torch.from_numpy(np.ascontiguousarray(orientation)).to(dev)
        st = RobotCurrentState(torch.from_numpy(np.ascontiguousarray(pose)).to(dev), ori[:, 0], ori[:, 1])
        if self.cfg["mode"] == "iterative":
            d = torch.from_numpy(np.ascontiguousarray(depth)).to(dev).unsqueeze(-1).permute(0, 3, 1, 2)  # NHWC view
            if self.pred:
                obs = Observations(None, d, torch.from_numpy(np.ascontiguousarray(logits)).to(dev))
            else:
                lab = torch.from_numpy(np.ascontiguousarray(labels)).to(dev).unsqueeze(-1).permute(0, 3, 1, 2)
                obs = Observations(lab, d, None)
        else:
            obs = Observations(None, None, None)
        out = self.mm(ei, obs, st)
        self.last_obs = obs
        return out.occupancy.cpu().numpy(), out.semantic.cpu().numpy()

    def world(self):
        w = self.mm.get_world_semantic_pointcloud()
        if w.xyz is None:
            return np.zeros(0, np.int64), np.zeros((0, 3), np.float32), np.zeros(0, np.uint8)
        return w.batch_indices.cpu().numpy(), w.xyz.cpu().numpy(), w.semantics.cpu().numpy()
