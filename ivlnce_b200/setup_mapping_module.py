"""Input adapter: obs dict (NHWC sensor tensors) -> the mapping module's dataclasses.

Mirror of reference `ivlnce_baselines/common/mapping_module/setup_mapping_module.py:13-89`
(same function names and semantics).  Config objects only need attribute access
(`HFOV`, `HEIGHT`, `WIDTH`, `height_clip`, `height_meters`, ...), so yacs nodes, habitat
`Config`s and plain namespaces all work.
"""
from __future__ import annotations

import math

from .mapper import CameraParameters, EpisodesInfo, MapDimensions, Observations, RobotCurrentState


def calculate_vertical_fov_in_degrees(depth_sensor_params) -> float:
    return depth_sensor_params.HFOV * (depth_sensor_params.HEIGHT / depth_sensor_params.WIDTH)


def calculate_verticial_fov_in_radians(depth_sensor_params) -> float:  # (sic) reference spelling
    return math.radians(calculate_vertical_fov_in_degrees(depth_sensor_params))


def extract_camera_parameters(depth_sensor_params, map_sensor_params) -> CameraParameters:
    return CameraParameters(
        vertical_fov_radians=calculate_verticial_fov_in_radians(depth_sensor_params),
        features_spatial_dimensions=(depth_sensor_params.HEIGHT, depth_sensor_params.WIDTH),
        height_clip=map_sensor_params.height_clip,
    )


def extract_egocentric_map_parameters(map_sensor_params) -> MapDimensions:
    return MapDimensions(
        height_meters=map_sensor_params.height_meters,
        width_meters=map_sensor_params.width_meters,
        resolution_meters=map_sensor_params.resolution_meters,
    )


def channel_first_representation(x):
    return None if x is None else x.permute(0, 3, 1, 2)


def setup_observations(observations_dict: dict) -> Observations:
    return Observations(
        semantics=channel_first_representation(observations_dict.get("semantic12", None)),
        depth_normalized=channel_first_representation(observations_dict.get("depth", None)),
        rgb=channel_first_representation(observations_dict.get("rgb", None)),
    )


def setup_inputs_from_obs_dict(observations_dict: dict):
    observations = setup_observations(observations_dict)
    episodes_info = EpisodesInfo(observations_dict["not_done_masks"], observations_dict["env_name"])
    orientation = observations_dict["world_robot_orientation"]
    robot_current_state = RobotCurrentState(
        pose=observations_dict["world_robot_pose"], elevation=orientation[:, 0], heading=orientation[:, 1])
    return episodes_info, observations, robot_current_state
