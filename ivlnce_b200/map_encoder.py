"""Host-side mirror of the map-feature front end of the reference's `SemanticMapEncoder`
(ivlnce_baselines/models/encoders/map_encoder.py:85-90): the tensor the CMA policy's map CNN consumes.

    occupancy = observations["occupancy_map"].unsqueeze(1)
    semantic  = F.one_hot(observations["semantic_map"].long(), num_classes).permute(0, 3, 1, 2)
    features  = torch.cat((occupancy, semantic), 1).to(torch.float)          # [B, 1 + K, R, C]

Here it is ONE kernel (`ivm_map_features`, a pure write stream) on the uint8 maps a step produces; the CNN itself is
outside this repository's scope.  CUDA tensors only -- there is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib

NUM_SEMANTIC_CLASSES = 13  # reference default (map_encoder.py:29)


class MapFeatures:
    """`generate_map_features` of the reference encoder.  The output buffer is reused between calls (like the maps of
    `MappingModule`, it is overwritten by the next call) unless `out` is given."""

    def __init__(self, num_semantic_classes: int = NUM_SEMANTIC_CLASSES):
        self.num_semantic_classes = int(num_semantic_classes)
        self._out: Optional[torch.Tensor] = None
        self._err: Optional[torch.Tensor] = None

    def __call__(self, observations: Dict[str, torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.generate_map_features(observations, out)

    def generate_map_features(self, observations: Dict[str, torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        for k in ("occupancy_map", "semantic_map"):
            if k not in observations:
                raise ValueError(f"Observation `{k}` is missing.")  # map_encoder.py:93-95
        occ, sem = observations["occupancy_map"], observations["semantic_map"]
        if occ.device.type != "cuda" or sem.device != occ.device:
            raise _lib.MapLibraryError("map features run on CUDA tensors only; there is no CPU fallback")
        if occ.dtype is not torch.uint8 or not occ.is_contiguous():
            occ = occ.to(torch.uint8).contiguous()
        if sem.dtype is not torch.uint8 or not sem.is_contiguous():
            sem = sem.to(torch.uint8).contiguous()
        B, R, C = occ.shape
        assert tuple(sem.shape) == (B, R, C)
        K = self.num_semantic_classes
        if out is None:
            if self._out is None or tuple(self._out.shape) != (B, 1 + K, R, C) or self._out.device != occ.device:
                self._out = torch.empty((B, 1 + K, R, C), dtype=torch.float32, device=occ.device)
            out = self._out
        assert out.dtype is torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, 1 + K, R, C)
        if self._err is None or self._err.device != occ.device:
            self._err = torch.zeros(1, dtype=torch.int32, device=occ.device)
        lib = _lib.load()
        dev_index = occ.device.index if occ.device.index is not None else torch.cuda.current_device()
        stream = torch._C._cuda_getCurrentRawStream(dev_index)
        with torch.cuda.device(occ.device):
            _lib.check(lib.ivm_map_features(occ.data_ptr(), sem.data_ptr(), B, R, C, K, out.data_ptr(), self._err.data_ptr(),
                                            stream), None, "ivm_map_features")
        return out

    def check_errors(self) -> None:
        """Synchronises; raises like F.one_hot does if a semantic value was >= num_semantic_classes."""
        if self._err is not None and int(self._err.item()) != 0:
            self._err.zero_()
            raise RuntimeError("Class values must be smaller than num_classes.")
