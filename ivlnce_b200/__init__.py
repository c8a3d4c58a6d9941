"""B200-native semantic-map update for IVLN-CE's MapCMA agents (hot path only).

Public surface mirrors the reference's mapping module:
  ivlnce_b200.mapper                 MappingModule, factories, dataclasses
  ivlnce_b200.setup_mapping_module   obs-dict adapter
  ivlnce_b200.obs_transforms         Mapper plugin shim (+ 4 registered variants)
  ivlnce_b200.sharding               env/tour partitioning across GPUs + gathers
The compute lives in csrc/ (sm_100a CUDA behind the C ABI of include/ivln_map.h).
"""
from .mapper import (CameraParameters, EpisodesInfo, MapDimensions, MappingModule, Observations,  # noqa: F401
                     RobotCurrentState, create_gt_semantics_iterative_mapper, create_gt_semantics_known_mapper,
                     create_iterative_mapper, create_known_mapper, create_predicted_semantics_iterative_mapper,
                     create_predicted_semantics_known_mapper)

__version__ = "0.1.0"
