"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's two host-side staging functions
(SURVEY.md section 8, rows f-2 / f-3), used to check ivlnce_b200/staging.py.

  * `batch_obs_oracle`               ivlnce_baselines/common/utils.py:57-92
  * `add_map_to_observations_oracle` ivlnce_baselines/trainers/iterative_collection_dagger_trainer.py:28-58

Pinned against the unmodified reference functions in tests/test_staging.py (build container only, where
/root/reference exists; oracle/ref_loader.py loads them by path with stub `gym` / `habitat_baselines` modules).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Optional, Set

import numpy as np
import torch


def batch_obs_oracle(observations: List[Dict], device: Optional[torch.device] = None,
                     ignore_keys: Optional[Set[str]] = None) -> Dict:
    if ignore_keys is None:
        ignore_keys = {"env_name"}                              # utils.py:70-71
    batch = defaultdict(list)
    for obs in observations:                                    # utils.py:75-81
        for sensor in obs:
            value = obs[sensor]
            if isinstance(value, np.ndarray) and value.dtype == np.uint32:
                value = np.int32(value)
            if sensor not in ignore_keys:
                value = torch.as_tensor(value)
            obs[sensor] = value
            batch[sensor].append(value)
    out: Dict = {}
    for sensor in batch:                                        # utils.py:85-90
        if sensor not in ignore_keys:
            out[sensor] = torch.stack(batch[sensor], dim=0).to(device)
        else:
            out[sensor] = batch[sensor]
    return out


def add_map_to_observations_oracle(observations: List[Dict], batch: Dict, num_envs: int) -> List[Dict]:
    k_sum = int("occupancy_map" in batch) + int("semantic_map" in batch)
    if k_sum == 1:
        raise RuntimeError("either both map keys should exist in the batch or neither")
    elif k_sum != 2:
        return observations
    for i in range(num_envs):
        for k in ["occupancy_map", "semantic_map"]:
            observations[i][k] = batch[k][i].cpu().numpy()
        for k in ["semantic", "semantic12", "world_robot_pose", "world_robot_orientation", "env_name"]:
            if k in observations[i]:
                del observations[i][k]
    return observations
