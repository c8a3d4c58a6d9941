"""Drives the REAL reference mapping module (build container only) step by step
on a synthetic scenario.  Used by tests/golden/make_golden.py and by the
`not gpu` tests that pin the oracle against the reference when /root/reference
is present."""
from __future__ import annotations

import os
import tempfile
from typing import Dict, List, Optional

import numpy as np
import torch

from oracle.ref_loader import load_reference_mapper, reference_available  # noqa: F401


class ReferenceRunner:
    def __init__(self, height, width, vfov, map_m, resolution, mode="iterative",
                 known_clouds: Optional[Dict[str, tuple]] = None):
        torch.set_num_threads(1)  # the reference's semantic write is racy multi-threaded
        self.ref = load_reference_mapper()
        ref = self.ref
        self.md = ref.MapDimensions(map_m, map_m, resolution)
        self.mode = mode
        if mode == "iterative":
            cam = ref.CameraParameters(vfov, (height, width), 0.1)
            self.mm = ref.create_gt_semantics_iterative_mapper(torch.device("cpu"), cam, self.md)
        else:
            self._tmp = tempfile.TemporaryDirectory()
            for name, (xyz, sem) in (known_clouds or {}).items():
                np.savez(os.path.join(self._tmp.name, f"{name}.npz"), xyz=xyz, semantics=sem)
            self.mm = ref.create_known_mapper(torch.device("cpu"), self.md, self._tmp.name)

    def step(self, masks, pose, orientation, depth=None, labels=None, env_names: Optional[List[str]] = None):
        ref = self.ref
        B = masks.shape[0]
        names = list(env_names) if env_names is not None else [f"scene{b}" for b in range(B)]
        ei = ref.EpisodesInfo(torch.from_numpy(np.ascontiguousarray(masks)).reshape(B, 1), names)
        ori = torch.from_numpy(np.ascontiguousarray(orientation))
        if self.mode == "iterative":
            obs = ref.Observations(
                torch.from_numpy(np.ascontiguousarray(labels)).unsqueeze(1),
                torch.from_numpy(np.ascontiguousarray(depth)).unsqueeze(1), None)
        else:
            obs = ref.Observations(None, None, None)
        st = ref.RobotCurrentState(torch.from_numpy(np.ascontiguousarray(pose)), ori[:, 0], ori[:, 1])
        out = self.mm(ei, obs, st)
        return out.occupancy.numpy().copy(), out.semantic.numpy().copy()

    def world(self):
        w = self.mm.get_world_semantic_pointcloud()
        if w.xyz is None:
            return (np.zeros(0, np.int64), np.zeros((0, 3), np.float32), np.zeros(0, np.uint8))
        return (w.batch_indices.numpy().astype(np.int64).copy(), w.xyz.numpy().copy(),
                w.semantics.numpy().astype(np.uint8).copy())
