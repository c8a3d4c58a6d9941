"""The obs-transform plugin shim: `Mapper` and its four registered variants.

Mirror of reference `ivlnce_baselines/common/obs_transforms.py:30-176`: same class names
(the registry key is the class name), `from_config`, `transform_observation_space`,
`forward(dict) -> dict`; adds `occupancy_map` / `semantic_map`, deletes the consumed keys.
When habitat-lab is importable the classes register themselves with its
`baseline_registry` and derive from its `ObservationTransformer`; otherwise a tiny local
registry with the same lookup call is used, so the module works standalone.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch.nn as nn
from torch import Tensor

from .mapper import (CameraParameters, MapDimensions, create_gt_semantics_iterative_mapper,
                     create_gt_semantics_known_mapper, create_predicted_semantics_iterative_mapper,
                     create_predicted_semantics_known_mapper)
from .setup_mapping_module import (extract_camera_parameters, extract_egocentric_map_parameters,
                                   setup_inputs_from_obs_dict)

try:  # real plugin API (habitat-lab/habitat_baselines/common/obs_transformers.py:45-64)
    from habitat_baselines.common.baseline_registry import baseline_registry
    from habitat_baselines.common.obs_transformers import ObservationTransformer
except Exception:  # standalone
    class ObservationTransformer(nn.Module):
        def transform_observation_space(self, observation_space, **kwargs):
            return observation_space

        @classmethod
        def from_config(cls, config):
            raise NotImplementedError

        def forward(self, observations):
            return observations

    class _Registry:
        def __init__(self):
            self._obs_transformers = {}

        def register_obs_transformer(self, to_register=None, *, name=None):
            def wrap(cls):
                self._obs_transformers[name or cls.__name__] = cls
                return cls
            return wrap if to_register is None else wrap(to_register)

        def get_obs_transformer(self, name):
            return self._obs_transformers.get(name)

    baseline_registry = _Registry()

try:
    from gym import spaces
    _Box = spaces.Box
except Exception:
    class _Box:  # minimal stand-in for gym.spaces.Box
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


@baseline_registry.register_obs_transformer()
class Mapper(ObservationTransformer):
    def __init__(self, camera_parameters: CameraParameters, map_dimensions: MapDimensions, visualize=False,
                 **mapper_kwargs):
        super().__init__()
        self.camera_parameters = camera_parameters
        self.map_dimensions = map_dimensions
        self.visualize = visualize  # rendering (visualize_semantic_map.py) is out of scope here
        self.mapping_module = None
        self.mapper_kwargs = mapper_kwargs
        self.keys_to_delete = ["world_robot_orientation", "world_robot_pose", "semantic", "semantic12", "env_name"]

    def transform_observation_space(self, observation_space, **kwargs):
        md = self.map_dimensions
        nrows = math.ceil(md.height_meters / md.resolution_meters)
        ncols = math.ceil(md.width_meters / md.resolution_meters)
        for new_key in ("occupancy_map", "semantic_map"):
            observation_space.spaces[new_key] = _Box(low=np.iinfo(np.uint8).min, high=np.iinfo(np.uint8).max,
                                                     shape=(nrows, ncols), dtype=np.uint8)
        for key in self.keys_to_delete:
            if key in observation_space.spaces:
                del observation_space.spaces[key]
        return observation_space

    def forward(self, observations: Dict[str, Tensor]) -> Dict[str, Tensor]:
        self.setup_mapping_module(observations)
        observations = self.update_maps_from_observations(observations)
        return self.delete_extra_information(observations)

    def setup_mapping_module(self, observations: Dict[str, Tensor]):
        raise NotImplementedError

    def update_maps_from_observations(self, observations):
        episodes_info, input_observations, robot_current_state = setup_inputs_from_obs_dict(observations)
        maps = self.mapping_module(episodes_info, input_observations, robot_current_state)
        observations["occupancy_map"] = maps.occupancy
        observations["semantic_map"] = maps.semantic
        return observations

    def delete_extra_information(self, observations):
        for key in self.keys_to_delete:
            if key in observations:
                del observations[key]
        return observations

    @classmethod
    def from_config(cls, config, visualize=False):
        mapper_cfg = config.RL.POLICY.OBS_TRANSFORMS.EGOCENTRIC_MAPPER
        # optional keys beside the reference's (yacs nodes are created with new_allowed=True): B200_MAPPER = dict of
        # keyword arguments for the mapping module (store_cells, maps_location, max_envs, ...)
        extra = getattr(mapper_cfg, "B200_MAPPER", None)
        kwargs = dict(extra) if extra is not None else {}
        return cls(
            camera_parameters=extract_camera_parameters(
                depth_sensor_params=config.TASK_CONFIG.SIMULATOR.DEPTH_SENSOR, map_sensor_params=mapper_cfg),
            map_dimensions=extract_egocentric_map_parameters(map_sensor_params=mapper_cfg),
            visualize=(len(getattr(config, "VIDEO_OPTION", [])) > 0) or visualize,
            **kwargs,
        )


@baseline_registry.register_obs_transformer()
class GTSemanticsIterativeMapper(Mapper):
    def setup_mapping_module(self, observations):
        if self.mapping_module is None:
            self.mapping_module = create_gt_semantics_iterative_mapper(
                device=observations["depth"].device, camera_parameters=self.camera_parameters,
                map_dimensions=self.map_dimensions, **self.mapper_kwargs)


@baseline_registry.register_obs_transformer()
class PredictedSemanticsIterativeMapper(Mapper):
    def setup_mapping_module(self, observations):
        if self.mapping_module is None:
            self.mapping_module = create_predicted_semantics_iterative_mapper(
                device=observations["depth"].device, camera_parameters=self.camera_parameters,
                map_dimensions=self.map_dimensions, **self.mapper_kwargs)


@baseline_registry.register_obs_transformer()
class GTSemanticsKnownMapper(Mapper):
    def setup_mapping_module(self, observations):
        if self.mapping_module is None:
            self.mapping_module = create_gt_semantics_known_mapper(
                device=observations["depth"].device, map_dimensions=self.map_dimensions, **self.mapper_kwargs)


@baseline_registry.register_obs_transformer()
class PredictedSemanticsKnownMapper(Mapper):
    def setup_mapping_module(self, observations):
        if self.mapping_module is None:
            self.mapping_module = create_predicted_semantics_known_mapper(
                device=observations["depth"].device, map_dimensions=self.map_dimensions, **self.mapper_kwargs)


def get_active_obs_transforms(config):
    """habitat-lab obs_transformers.py:1194-1206."""
    out = []
    if hasattr(config.RL.POLICY, "OBS_TRANSFORMS"):
        for name in config.RL.POLICY.OBS_TRANSFORMS.ENABLED_TRANSFORMS:
            out.append(baseline_registry.get_obs_transformer(name).from_config(config))
    return out


def apply_obs_transforms_batch(batch, obs_transforms):
    """habitat-lab obs_transformers.py:1209-1215."""
    for t in obs_transforms:
        batch = t(batch)
    return batch
