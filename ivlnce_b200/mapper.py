"""Host-side mirror of the reference mapping module's interface.

Same names, argument meaning, tensor layouts and error behaviour as
`ivlnce_baselines/common/mapping_module/mapper.py` (reference), so that the
IVLN-CE trainers and the obs-transform plugin (`obs_transforms.py:89-103`)
can call it unchanged:

    mapper = create_gt_semantics_iterative_mapper(device, camera_parameters, map_dimensions)
    maps = mapper(episodes_info, observations, robot_current_state)
    maps.occupancy, maps.semantic        # uint8 [B, num_rows, num_cols], views of internal buffers

Everything below the dataclasses is different: the world state is a dense
per-env half-cell store in HBM and every step is five CUDA kernel launches
through the C ABI of libivlnmap.so (include/ivln_map.h).  There is no CPU
implementation in this package.
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .geometry import camera_scale_tables, camera_to_world_rows, ego_rotation

DEFAULT_STORE_CELLS = 2048       # half-cells per side of an env's world store
DEFAULT_KNOWN_CAPACITY = 1 << 22  # points per env in known-map mode


# ----------------------------------------------------------------------------
# Interface dataclasses (reference mapper.py:61-138, 195-200, 336-340)
class EpisodesInfo:
    """reference mapper.py:61-86: `not_done_masks` [B,1]; 0 = episode (or tour) finished.

    Same constructor, attributes and methods as the reference's dataclass.  The two conversions of its
    `__post_init__` (`np.asarray(env_names)`, `not_done_masks.squeeze(1)`) are made on first use: the map update reads
    the masks straight from the caller's tensor and compares the names as the sequence they came in."""

    EPISODE_FINISHED = 0
    EPISODE_UNFINISHED = 1
    __slots__ = ("_masks_raw", "_masks", "_names_raw", "_names")

    def __init__(self, not_done_masks: torch.Tensor, env_names: Sequence[str]):
        self._masks_raw, self._masks = not_done_masks, None
        self._names_raw, self._names = env_names, None

    @property
    def not_done_masks(self) -> torch.Tensor:
        if self._masks is None:
            self._masks = self._masks_raw.squeeze(1)
        return self._masks

    @not_done_masks.setter
    def not_done_masks(self, value: torch.Tensor):
        self._masks_raw = self._masks = value

    @property
    def env_names(self) -> np.ndarray:
        if self._names is None:
            self._names = np.asarray(self._names_raw)
        return self._names

    @env_names.setter
    def env_names(self, value):
        self._names_raw, self._names = value, None

    def __repr__(self):
        return f"EpisodesInfo(not_done_masks={self.not_done_masks!r}, env_names={self.env_names!r})"

    def __eq__(self, other):
        return (isinstance(other, EpisodesInfo) and torch.equal(self.not_done_masks, other.not_done_masks)
                and np.array_equal(self.env_names, other.env_names))

    @property
    def indices(self) -> torch.Tensor:  # built on demand: the hot path never needs it
        return torch.arange(self.not_done_masks.shape[0], device=self.not_done_masks.device)

    def finished(self) -> torch.Tensor:
        return self.not_done_masks == self.EPISODE_FINISHED

    def unfinished(self) -> torch.Tensor:
        return self.not_done_masks == self.EPISODE_UNFINISHED

    def finished_indices(self) -> torch.Tensor:
        return self.indices[self.finished()]

    @property
    def num_envs(self) -> int:
        return self._masks_raw.shape[0]

    def masks_flat(self) -> torch.Tensor:
        """The masks as the kernels read them, [B] elements in memory: the caller's [B,1] tensor itself when it is
        contiguous (no torch op), else the squeezed view."""
        raw = self._masks_raw
        if raw.dim() == 2 and raw.shape[1] == 1 and raw.stride(0) == 1:
            return raw
        return self.not_done_masks


@dataclass
class MapDimensions:
    """reference mapper.py:89-114."""
    height_meters: float
    width_meters: float
    resolution_meters: float
    num_rows: int = field(init=False)
    num_cols: int = field(init=False)

    def __post_init__(self):
        self.num_rows = math.ceil(self.height_meters / self.resolution_meters)
        self.num_cols = math.ceil(self.width_meters / self.resolution_meters)


@dataclass
class CameraParameters:
    """reference mapper.py:336-340 (`height_clip` is plumbed but unused on this path)."""
    vertical_fov_radians: float
    features_spatial_dimensions: tuple
    height_clip: float = 0.0


@dataclass
class Observations:
    """reference mapper.py:195-200; NCHW views of the sensor tensors."""
    semantics: Optional[torch.Tensor]
    depth_normalized: Optional[torch.Tensor]
    rgb: Optional[torch.Tensor]


@dataclass
class State:
    pose: Optional[torch.Tensor] = None
    elevation: Optional[torch.Tensor] = None
    heading: Optional[torch.Tensor] = None


class RobotCurrentState(State):
    """reference mapper.py:127-138."""

    @property
    def height(self):
        return self.pose[:, 1]

    def get_camera_matrix(self) -> torch.Tensor:
        """f32 [B,4,4] camera->world matrix (projector/core.py:6-37 with elevation + pi)."""
        rows = camera_to_world_rows(self.pose, self.elevation, self.heading)
        T = torch.zeros(rows.shape[0], 4, 4, dtype=torch.float32, device=rows.device)
        T[:, :3, :] = rows.view(-1, 3, 4)
        T[:, 3, 3] = 1
        return T


class RobotStartState(State):
    """reference mapper.py:141-178: pose of each env at its last reset (kept for API parity; nothing reads it)."""

    def __init__(self):
        super().__init__()
        self.batch_size = None

    def update(self, episodes_info: EpisodesInfo, current_state: RobotCurrentState):
        B, dev = episodes_info.num_envs, current_state.pose.device
        if self.batch_size != B:
            if self.batch_size is None or B > self.batch_size:
                self.pose = torch.zeros((B, 3), device=dev)
                self.elevation = torch.zeros((B,), device=dev)
                self.heading = torch.zeros((B,), device=dev)
            else:
                self.pose, self.elevation, self.heading = self.pose[:B], self.elevation[:B], self.heading[:B]
            self.batch_size = B
        done = episodes_info.finished()
        self.pose = torch.where(done.unsqueeze(1), current_state.pose.to(self.pose.dtype), self.pose)


class LocalizeRobot:
    """reference mapper.py:180-192."""

    def __init__(self):
        self.current_state = RobotCurrentState()
        self.start_state = RobotStartState()

    def __call__(self, episodes_info: EpisodesInfo, robot_current_state: RobotCurrentState):
        self.current_state = robot_current_state
        self.start_state.update(episodes_info, robot_current_state)
        return self


@dataclass
class SemanticPointcloud:
    """reference mapper.py:203-281 (read-only view handed out by get_world_semantic_pointcloud)."""
    batch_indices: Optional[torch.Tensor] = None
    xyz: Optional[torch.Tensor] = None
    semantics: Optional[torch.Tensor] = None

    def __len__(self):
        return 0 if self.xyz is None else self.xyz.shape[0]


class OccupancySemanticMapMemory:
    """reference mapper.py:620-648: `.occupancy` / `.semantic` are the live uint8 [B,R,C]
    buffers; they are overwritten by the next forward call."""

    def __init__(self):
        self._occ: Optional[torch.Tensor] = None
        self._sem: Optional[torch.Tensor] = None

    @property
    def occupancy(self) -> torch.Tensor:
        return self._occ

    @property
    def semantic(self) -> torch.Tensor:
        return self._sem

    @property
    def data(self) -> torch.Tensor:
        # the reference property (mapper.py:646-648) references undefined attributes; this is what it means
        return torch.stack((self._occ, self._sem), 1)


# ----------------------------------------------------------------------------
# Semantics front ends (reference mapper.py:651-800)
class ComputeSemantics(nn.Module):
    pass


class GTSemantics(ComputeSemantics):
    def forward(self, observations: Observations) -> Observations:
        if observations.semantics is None:
            raise Exception("Semantic Sensor not in use")
        return observations


class PredictSemantics(ComputeSemantics):
    """reference mapper.py:703-800.  The segmentation network itself (RedNet, dense convolutions on cuDNN, trained
    weights) is outside this repository's scope; everything round it is here:

      * the front end -- rgb / 255 -> bilinear resize to the depth size -> per-channel normalisation, and the depth
        normalisation (mapper.py:715-736, 788-793) -- is ONE kernel (`ivm_rednet_preprocess`);
      * the tail `scores.argmax(1, keepdims=True).to(uint8)` (mapper.py:795-798) is NOT done here: the scores go
        straight into the step kernel, which does the argmax while streaming the planes.

    `model`: callable (rgb_normalized f32 [B,3,H,W], depth_normalized f32 [B,1,H,W]) -> class scores f32 [B,Cls,H,W].
    Like the reference, the model is built lazily on the first forward (`setup_finetuned_rednet`, mapper.py:738-752):
    from `model_factory(device)` if given, else from the factory named by the environment variable
    IVLN_SEMANTICS_MODEL_FACTORY ("package.module:callable"), else from the reference's own RedNet class and weight
    file (`data/rednet_mp3d_best_model.pkl`) when `ivlnce_baselines` is importable.  If none of these exists the
    first forward raises, naming the three options."""

    REDNET_WEIGHTS = "data/rednet_mp3d_best_model.pkl"

    def __init__(self, model: Optional[Callable] = None, model_factory: Optional[Callable] = None):
        super().__init__()
        self.model = model
        self.model_factory = model_factory
        self._rgb_mean = self._rgb_std = None

    # -- the reference's lazy model construction
    def setup_finetuned_rednet(self, device: torch.device):
        if self.model is not None:
            return
        factory = self.model_factory
        spec = os.environ.get("IVLN_SEMANTICS_MODEL_FACTORY", "")
        if factory is None and spec:
            import importlib

            mod, _, fn = spec.partition(":")
            factory = getattr(importlib.import_module(mod), fn)
        if factory is not None:
            self.model = factory(device)
        else:
            try:
                from ivlnce_baselines.common.mapping_module.rednet import RedNet  # the reference's own network
            except Exception as exc:
                raise Exception(
                    "PredictSemantics needs a segmentation model: pass model= / model_factory= (callable returning class "
                    "scores [B,Cls,H,W]), or set IVLN_SEMANTICS_MODEL_FACTORY=package.module:callable, or make the "
                    "reference's ivlnce_baselines package (RedNet + data/rednet_mp3d_best_model.pkl) importable") from exc
            cfg_rednet = {"arch": "rednet", "resnet_pretrained": False, "finetune": True, "SUNRGBD_pretrained_weights": "",
                          "n_classes": 13, "upsample_prediction": True, "load_model": self.REDNET_WEIGHTS}
            model = RedNet(cfg_rednet).to(device)
            state = torch.load(cfg_rednet["load_model"])["model_state"]
            if next(iter(state)).split(".")[0] == "module":          # convert_weights_cuda_cpu(..., "cpu")
                state = {".".join(k.split(".")[1:]): v for k, v in state.items()}
            model.load_state_dict(state)
            self.model = model
        if isinstance(self.model, nn.Module):
            self.model.eval()
            for p in self.model.parameters():
                p.requires_grad = False

    def preprocess_torch(self, observations: Observations):
        """mapper.py:715-736, 788-793 as the reference's torch ops (the comparison path of the tests)."""
        depth = observations.depth_normalized
        if self._rgb_mean is None:
            dev = observations.rgb.device
            self._rgb_mean = torch.as_tensor([0.485, 0.456, 0.406], device=dev)[None, :, None, None]
            self._rgb_std = torch.as_tensor([0.229, 0.224, 0.225], device=dev)[None, :, None, None]
        rgb = torch.nn.functional.interpolate(observations.rgb.float() / 255.0, size=tuple(depth.shape[2:]), mode="bilinear")
        rgb = (rgb - self._rgb_mean) / self._rgb_std
        d = (depth - 0.213) / 0.285
        return rgb, d

    def preprocess(self, observations: Observations):
        """The same through the fused kernel: rgb u8 [B,3,h,w] (any strides: the sensor's NHWC memory seen as NCHW is
        read in place) + depth f32 [B,1,H,W] -> normalised f32 NCHW tensors."""
        rgb, depth = observations.rgb, observations.depth_normalized
        if rgb.device.type != "cuda":
            raise _lib.MapLibraryError("the segmentation front end runs on CUDA devices only; there is no CPU fallback")
        if rgb.dtype != torch.uint8:
            rgb = rgb.to(torch.uint8)
        depth = _as_f32(depth, rgb.device)
        B, _, H, W = depth.shape
        assert rgb.shape[0] == B and rgb.shape[1] == 3
        rgb_out = torch.empty((B, 3, H, W), dtype=torch.float32, device=rgb.device)
        depth_out = torch.empty((B, 1, H, W), dtype=torch.float32, device=rgb.device)
        strides = (ctypes.c_int64 * 4)(*rgb.stride())
        lib = _lib.load()
        with torch.cuda.device(rgb.device):
            _lib.check(lib.ivm_rednet_preprocess(rgb.data_ptr(), strides, B, int(rgb.shape[2]), int(rgb.shape[3]), depth.data_ptr(),
                                                 H, W, rgb_out.data_ptr(), depth_out.data_ptr(),
                                                 torch.cuda.current_stream(rgb.device).cuda_stream), None, "ivm_rednet_preprocess")
        return rgb_out, depth_out

    def scores(self, observations: Observations) -> torch.Tensor:
        if observations.rgb is None:
            raise Exception("RGB Sensor not in use")
        self.setup_finetuned_rednet(observations.depth_normalized.device)
        with torch.no_grad():
            rgb, d = self.preprocess(observations)
            return self.model(rgb, d)

    def forward(self, observations: Observations) -> Observations:
        observations.semantics = self.scores(observations).argmax(1, keepdims=True).to(torch.uint8)
        return observations


class PrecomputedScores(PredictSemantics):
    """Front end for callers that already hold the class scores (benchmarks, tests, an
    external segmentation service): `observations.rgb` carries the f32 [B,Cls,H,W] scores."""

    def scores(self, observations: Observations) -> torch.Tensor:
        if observations.rgb is None:
            raise Exception("RGB Sensor not in use")
        return observations.rgb


# ----------------------------------------------------------------------------
class _MapEngine:
    """Owns one libivlnmap context + its device workspace and output buffers."""

    def __init__(self, device: torch.device, map_dimensions: MapDimensions, camera: Optional[CameraParameters],
                 mode: str, max_envs: int, store_cells: int, known_capacity: int, tile: int = 0,
                 scatter_variant: int = 0, stamp_period: int = 0):
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.MapLibraryError(
                f"the B200 mapping module runs on CUDA devices only (got {device}); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = device
        self._dev_index = device.index if device.index is not None else torch.cuda.current_device()
        self.md = map_dimensions
        self.camera = camera
        self.mode = mode
        self.store_cells = int(store_cells)
        self.known_capacity = int(known_capacity)
        self.tile = int(tile)
        self.scatter_variant = int(scatter_variant)  # 0 = auto (fused persistent step kernel when it applies), 1 / 2 = four kernels (register-staged / bulk-async score loads)
        self.stamp_period = int(stamp_period)        # test switch: period of the candidate-plane stamp (0 = longest)
        self.ctx = None
        self.workspace = None
        self.max_envs = 0
        self._create(max_envs)

    # -- context lifetime
    def _config(self, max_envs: int) -> _lib.IvmConfig:
        md = self.md
        H, W = (self.camera.features_spatial_dimensions if self.camera is not None else (0, 0))
        reserved = (ctypes.c_int32 * 4)(self.scatter_variant, int(os.environ.get("IVM_DEBUG_FLAGS", "0")), 0,
                                        int(self.stamp_period))
        return _lib.IvmConfig(
            reserved=reserved, max_envs=max_envs, height=int(H), width=int(W), map_rows=md.num_rows, map_cols=md.num_cols,
            res=np.float32(md.resolution_meters), half_res=np.float32(md.resolution_meters / 2),
            half_h=np.float32(md.height_meters / 2), half_w=np.float32(md.width_meters / 2),
            store_rows=self.store_cells, store_cols=self.store_cells, mode=0 if self.mode == "iterative" else 1,
            known_capacity=self.known_capacity if self.mode == "known" else 0,
            tile_rows=self.tile, tile_cols=self.tile)

    def _create(self, max_envs: int):
        cfg = self._config(max_envs)
        nbytes = self.lib.ivm_workspace_bytes(ctypes.byref(cfg))
        if nbytes == 0:
            raise _lib.MapLibraryError("invalid map configuration")
        with torch.cuda.device(self.device):
            workspace = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = workspace.data_ptr()
            aligned = (base + 255) & ~255
            ctx = ctypes.c_void_p()
            _lib.check(self.lib.ivm_create(ctypes.byref(cfg), aligned, nbytes, ctypes.byref(ctx)), None, "ivm_create")
            stream = torch.cuda.current_stream(self.device).cuda_stream
            old = self.ctx
            if old is not None:
                _lib.check(self.lib.ivm_copy_state(ctx, old, stream), ctx, "ivm_copy_state")
                torch.cuda.current_stream(self.device).synchronize()
                self.lib.ivm_destroy(old)
            elif self.camera is not None:
                H, W = self.camera.features_spatial_dimensions
                xs, ys = camera_scale_tables(int(H), int(W), float(self.camera.vertical_fov_radians), self.device)
                _lib.check(self.lib.ivm_set_camera(ctx, xs.data_ptr(), ys.data_ptr(), stream), ctx, "ivm_set_camera")
                self._tables = (xs, ys)
            self.ctx, self.workspace, self.max_envs = ctx, workspace, max_envs
            R, C = self.md.num_rows, self.md.num_cols
            # both output maps in one block, so that a full batch leaves for the host in a single copy (staging.MapEgress)
            self.maps = torch.zeros((2, max_envs, R, C), dtype=torch.uint8, device=self.device)
            self.occ, self.sem = self.maps[0], self.maps[1]
            self.occ_ptr, self.sem_ptr, self.labels_out_ptr = self.occ.data_ptr(), self.sem.data_ptr(), None
            if self.mode == "iterative":
                H, W = self.camera.features_spatial_dimensions
                self.labels_out = torch.zeros((max_envs, int(H), int(W)), dtype=torch.uint8, device=self.device)
                self.labels_out_ptr = self.labels_out.data_ptr()

    def ensure_capacity(self, num_envs: int) -> bool:
        """True if a new context was created (its per-context switches have to be set again)."""
        if num_envs > self.max_envs:
            self._create(max(num_envs, 2 * self.max_envs))
            return True
        return False

    def close(self):
        if self.ctx is not None:
            self.lib.ivm_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stream(self) -> int:
        # raw handle of torch's current stream on this device (no Stream object per call)
        return torch._C._cuda_getCurrentRawStream(self._dev_index)

    def status(self) -> Tuple[int, List[int]]:
        st = _lib.IvmStatus()
        _lib.check(self.lib.ivm_read_status(self.ctx, ctypes.byref(st), self.stream()), self.ctx, "ivm_read_status")
        return int(st.error_flags), [int(v) for v in st.stats]


def _as_u8_masks(masks: torch.Tensor, device) -> torch.Tensor:
    if masks.dtype is torch.uint8 and masks.device == device and masks.is_contiguous():
        return masks
    m = masks.to(device=device, non_blocking=True)
    if m.dtype != torch.uint8:
        m = (m != 0).to(torch.uint8)
    return m.contiguous()


def _as_f32(t: torch.Tensor, device) -> torch.Tensor:
    """`t` as a contiguous float32 tensor on `device` (no torch call at all when it already is one)."""
    if t.dtype is torch.float32 and t.device == device and t.is_contiguous():
        return t
    return t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


class _Hold:
    __slots__ = ("keepalive", "num_envs", "views_B", "occ_base", "occ_view", "sem_view", "err_host", "err_event", "err_calls",
                 "err_stream")

    def __init__(self):
        self.keepalive, self.num_envs, self.views_B = None, 0, -1
        self.occ_base = self.occ_view = self.sem_view = None
        self.err_host, self.err_event, self.err_calls, self.err_stream = None, None, 0, None   # error polling (_poll_errors)


class MappingModule(nn.Module):
    """Drop-in for reference `MappingModule` (mapper.py:904-947): same forward signature, returns an
    object with `.occupancy` / `.semantic` uint8 [B, num_rows, num_cols] on the input device.

    mode "iterative": world state built from depth + semantics every step (episodic or iterative
    depending on which masks the caller passes, base_il_trainer.py:746-753);
    mode "known": scene clouds `{maps_location}/{env_name}.npz` loaded on reset (mapper.py:851-881).
    """

    def __init__(self, device, map_dimensions: MapDimensions, camera_parameters: Optional[CameraParameters] = None,
                 compute_semantics: Optional[ComputeSemantics] = None, mode: str = "iterative",
                 maps_location: Optional[str] = None, max_envs: Optional[int] = None,
                 store_cells: int = DEFAULT_STORE_CELLS, known_capacity: int = DEFAULT_KNOWN_CAPACITY,
                 trig: str = "kernel", host_trig: bool = False, raster_tile: int = 0,
                 track_start_state: bool = False, scatter_variant: int = 0, stamp_period: int = 0,
                 pipelined: bool = False, error_poll_interval: int = 32, on_overflow: str = "raise"):
        super().__init__()
        assert mode in ("iterative", "known")
        self.device = torch.device(device)
        self.map_dimensions = map_dimensions
        self.camera_parameters = camera_parameters
        self.compute_semantics = compute_semantics
        self.mode = mode
        self.maps_location = maps_location
        self.localize_robot = LocalizeRobot()
        self.map_memory = OccupancySemanticMapMemory()
        # where sin/cos of the pose angles are evaluated:
        #   "kernel": inside the prep kernel, in the angles' dtype (no torch launches per step);
        #   "torch" : torch ops on the tensors' device (what the reference does on a GPU);
        #   "host"  : torch ops on the host CPU (bit-identical to a CPU run of the reference; syncs).
        self.trig = "host" if host_trig else trig
        assert self.trig in ("kernel", "torch", "host")
        self.track_start_state = track_start_state
        self._engine_args = dict(store_cells=store_cells, known_capacity=known_capacity, tile=raster_tile,
                                 scatter_variant=scatter_variant, stamp_period=stamp_period)
        self._engine: Optional[_MapEngine] = None
        self._pipelined = bool(pipelined)
        # Errors reach a caller that never asks for them: every `error_poll_interval` forwards the device's error flags
        # are copied into pinned host memory (asynchronously, no synchronisation); a later forward that finds the copy
        # complete and a flag set raises MapLibraryError (on_overflow="warn": window overflows only warn, once).
        assert on_overflow in ("raise", "warn")
        self.error_poll_interval = int(error_poll_interval)
        self.on_overflow = on_overflow
        self._warned = False
        self._initial_max_envs = max_envs
        self._hold = _Hold()   # per-call state kept off the nn.Module attribute machinery
        self._known_cache: Dict[str, Tuple[torch.Tensor, torch.Tensor, int, int]] = {}
        self._known_order: List[int] = []   # env indices in load order (world-cloud order in known mode)
        self._known_loaded: Dict[int, str] = {}
        self._known_names = None    # env_names of the last call if every env holds the scene it is in
        self._known_names_raw = None   # ... as the list the caller passed (a copy), for the cheap comparison
        self._known_B = -1

    @property
    def _num_envs(self) -> int:
        return self._hold.num_envs

    # -- engine
    def engine(self, num_envs: int) -> _MapEngine:
        if self._engine is None:
            n = max(num_envs, self._initial_max_envs or 0)
            self._engine = _MapEngine(self.device, self.map_dimensions, self.camera_parameters, self.mode, n,
                                      **self._engine_args)
            if self._pipelined and self.mode == "iterative":
                _lib.check(self._engine.lib.ivm_set_pipelined(self._engine.ctx, 1))
        if self._engine.ensure_capacity(num_envs) and self.mode == "iterative":
            _lib.check(self._engine.lib.ivm_set_pipelined(self._engine.ctx, 1 if self._pipelined else 0))
        return self._engine

    def set_pipelined(self, enabled: bool = True):
        """Pipelined stepping (include/ivln_map.h, ivm_set_pipelined): the caller promises that the tensors passed to
        a forward() call are complete on the stream before the PREVIOUS forward() call was issued (resident or replayed
        inputs; inputs staged by copies rather than kernels).  Back-to-back calls then overlap on the GPU: the next
        step's score stream and depth filter run beside the last ego tiles of the current step.  Same results."""
        self._pipelined = bool(enabled)
        if self._engine is not None and self.mode == "iterative":
            _lib.check(self._engine.lib.ivm_set_pipelined(self._engine.ctx, 1 if self._pipelined else 0))
        return self

    def _matrices(self, state: RobotCurrentState):
        pose, elev, head = state.pose, state.elevation, state.heading
        if self.trig == "kernel":
            pose32 = _as_f32(pose, self.device)
            if elev.dtype != head.dtype or head.dtype not in (torch.float32, torch.float64):
                elev, head = elev.to(torch.float64), head.to(torch.float64)
            n = elev.shape[0]
            interleaved = (elev.device == self.device and head.device == self.device and elev.dim() == 1
                           and head.data_ptr() - elev.data_ptr() == elev.element_size()
                           and (n == 1 or (elev.stride(0) == 2 and head.stride(0) == 2)))
            if interleaved:
                orient = elev  # the two views of the [B,2] world_robot_orientation tensor: use it in place
            else:
                orient = torch.stack((elev, head), 1).to(self.device, non_blocking=True).contiguous()
            return None, None, pose32, orient
        if self.trig == "host":
            p, e, h = pose.detach().cpu(), elev.detach().cpu(), head.detach().cpu()
            T12 = camera_to_world_rows(p, e, h).to(self.device, non_blocking=True)
            cs = ego_rotation(h).to(self.device, non_blocking=True)
        else:
            pose = pose.to(self.device, non_blocking=True)
            elev = elev.to(self.device, non_blocking=True)
            head = head.to(self.device, non_blocking=True)
            T12 = camera_to_world_rows(pose, elev, head)
            cs = ego_rotation(head)
        pose32 = pose.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        return T12.contiguous(), cs.contiguous(), pose32, None

    # -- forward
    def forward(self, episodes_info: EpisodesInfo, observations: Observations,
                robot_current_state: RobotCurrentState) -> OccupancySemanticMapMemory:
        # The step is ONE kernel of a few tens of microseconds, so the host path is kept as short: no autograd
        # bookkeeping (nothing here builds a graph: the library is called with raw pointers), no nn.Module
        # attribute traffic, no device context switch unless the current device really differs.
        if self.track_start_state:
            self.localize_robot(episodes_info, robot_current_state)
        else:
            self.localize_robot.current_state = robot_current_state
        B = episodes_info.num_envs
        eng = self._engine
        if eng is None or B > eng.max_envs:
            eng = self.engine(B)
        hold = self._hold
        if torch.cuda.current_device() != eng._dev_index:
            with torch.cuda.device(self.device):
                return self._forward_on_device(eng, B, episodes_info, observations, robot_current_state, hold)
        return self._forward_on_device(eng, B, episodes_info, observations, robot_current_state, hold)

    def _poll_errors(self, eng, hold):
        # (the counters live in `hold`, a plain object: an attribute store on an nn.Module costs microseconds)
        ev = hold.err_event
        if ev is not None and ev.query():           # the copy issued a while ago has landed
            hold.err_event = None
            flags = int(hold.err_host[0])
            if flags:
                self._raise_flags(eng, flags)
        hold.err_calls += 1
        if hold.err_event is None and hold.err_calls >= self.error_poll_interval:
            hold.err_calls = 0
            if hold.err_host is None:
                hold.err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
                # The flags are sticky, so WHEN they are read does not matter: the copy runs on a side stream and nothing
                # is put between two steps on the caller's stream (an event or a copy there would end the overlap of
                # back-to-back steps).  The side stream is ordered once behind the work that created the context.
                hold.err_stream = torch.cuda.Stream(self.device)
                hold.err_stream.wait_stream(torch.cuda.current_stream(self.device))
            _lib.check(eng.lib.ivm_copy_error_flags_async(eng.ctx, hold.err_host.data_ptr(), hold.err_stream.cuda_stream), eng.ctx,
                       "ivm_copy_error_flags_async")
            hold.err_event = torch.cuda.Event()
            hold.err_event.record(hold.err_stream)

    def _raise_flags(self, eng, flags: int):
        overflow = flags & (_lib.ERR_STORE_OVERFLOW | _lib.ERR_KNOWN_OVERFLOW)
        fatal = flags & ~overflow
        if fatal & _lib.ERR_GRID_BARRIER:
            self._engine = None                       # the context's barrier counters are out of step: start afresh
            raise _lib.MapLibraryError("grid barrier time-out in the fused step kernel: the maps are invalid and the world "
                                       "state has been dropped (a new context is created on the next call)")
        if fatal & _lib.ERR_EDGE_OVERFLOW:
            raise _lib.MapLibraryError("edge list capacity exceeded: bounding-box edge collisions may be unresolved")
        if fatal:
            raise _lib.MapLibraryError(f"map library error flags {flags}")
        msg = ("points fell outside an env's world store window and were DROPPED (the reference's unbounded cloud keeps "
               f"them): raise store_cells (now {eng.store_cells} half-cells = "
               f"{eng.store_cells * self.map_dimensions.resolution_meters / 2:.1f} m per side)")
        _lib.check(eng.lib.ivm_clear_error_flags(eng.ctx, eng.stream()), eng.ctx, "ivm_clear_error_flags")
        if self.on_overflow == "raise":
            raise _lib.MapLibraryError(msg)
        if not self._warned:
            import warnings

            warnings.warn(msg)
            self._warned = True

    def _forward_on_device(self, eng, B, episodes_info, observations, robot_current_state, hold):
        if self.error_poll_interval > 0:
            self._poll_errors(eng, hold)
        T12, cs, pose, orient = self._matrices(robot_current_state)
        hold.keepalive = (T12, cs, pose, orient)
        if self.mode == "iterative":
            self._forward_iterative(eng, B, episodes_info, observations, T12, cs, pose, orient)
        else:
            self._forward_known(eng, B, episodes_info, cs, pose, orient)
        hold.num_envs = B
        mem = self.map_memory
        if hold.views_B != B or hold.occ_base is not eng.occ:
            hold.views_B, hold.occ_base, hold.occ_view, hold.sem_view = B, eng.occ, eng.occ[:B], eng.sem[:B]
        mem._occ = hold.occ_view
        mem._sem = hold.sem_view
        return mem

    def _forward_iterative(self, eng, B, episodes_info, observations, T12, cs, pose, orient):
        H, W = self.camera_parameters.features_spatial_dimensions
        H, W = int(H), int(W)
        # (an nn.Module attribute that is itself a module is found only after a failed normal lookup: ask the dicts)
        sem_mod = self.__dict__.get("compute_semantics")
        if sem_mod is None:
            sem_mod = self._modules.get("compute_semantics")
        scores = None
        if isinstance(sem_mod, PredictSemantics):
            scores = sem_mod.scores(observations)
        elif type(sem_mod) is GTSemantics:      # (its forward, without the nn.Module call machinery)
            if observations.semantics is None:
                raise Exception("Semantic Sensor not in use")
        elif sem_mod is not None:
            observations = sem_mod(observations)
        depth = observations.depth_normalized
        assert depth.shape[2] == H  # projector/point_cloud.py:68-69
        assert depth.shape[3] == W
        depth = _as_f32(depth, self.device)
        masks = _as_u8_masks(episodes_info.masks_flat(), self.device)
        labels_ptr, logits_ptr, ncls = None, None, 0
        if scores is not None:
            scores = _as_f32(scores, self.device)
            assert scores.shape[0] == B and scores.shape[2] == H and scores.shape[3] == W
            logits_ptr, ncls = scores.data_ptr(), int(scores.shape[1])
            # side effect of PredictSemantics.forward (mapper.py:796-798)
            observations.semantics = eng.labels_out[:B].view(B, 1, H, W)
        else:
            labels = observations.semantics
            if not (labels.dtype is torch.uint8 and labels.device == self.device and labels.is_contiguous()):
                labels = labels.to(device=self.device, non_blocking=True)
                if labels.dtype != torch.uint8:
                    labels = labels.to(torch.uint8)
                labels = labels.contiguous()
            assert labels.numel() == B * H * W
            labels_ptr = labels.data_ptr()
        args = (eng.ctx, B, depth.data_ptr(), labels_ptr, logits_ptr, ncls, eng.labels_out_ptr,
                None if T12 is None else T12.data_ptr(), pose.data_ptr(), None if cs is None else cs.data_ptr(),
                None if orient is None else orient.data_ptr(),
                1 if (orient is not None and orient.dtype == torch.float64) else 0,
                masks.data_ptr(), eng.occ_ptr, eng.sem_ptr, eng.stream())
        rc = eng.lib.ivm_step_iterative(*args)
        if rc == 4:  # 2^24 - 1 steps: rebase the stamps and retry once
            _lib.check(eng.lib.ivm_rebase_stamps(eng.ctx, eng.stream()), eng.ctx, "ivm_rebase_stamps")
            rc = eng.lib.ivm_step_iterative(*args)
        _lib.check(rc, eng.ctx, "ivm_step_iterative")

    # -- known-map mode
    def get_map_file(self, env_name: str) -> str:
        return os.path.join(self.maps_location, f"{env_name}.npz")

    KNOWN_CACHE_SCENES = 8   # scene clouds kept resident on the device (least recently used beyond that are re-read from
                             # disk on their next reset, which is what the reference does on EVERY reset)

    def _known_cloud(self, env_name: str):
        if env_name in self._known_cache:
            self._known_cache[env_name] = self._known_cache.pop(env_name)   # most recently used last
        else:
            with np.load(self.get_map_file(env_name)) as f:  # SemanticPointcloud.from_npz_file, mapper.py:283-294
                xyz = np.ascontiguousarray(f["xyz"], dtype=np.float32)
                sem = np.ascontiguousarray(np.asarray(f["semantics"]).astype(np.int64).astype(np.uint8))
            hr = np.float32(self.map_dimensions.resolution_meters / 2)
            if xyz.shape[0]:
                o_r = int(np.floor(float(xyz[:, 2].min()) / float(hr))) - 2
                o_c = int(np.floor(float(xyz[:, 0].min()) / float(hr))) - 2
            else:
                o_r = o_c = 0
            cap = self._engine_args["known_capacity"]
            if xyz.shape[0] > cap:
                raise _lib.MapLibraryError(f"scene cloud {self.get_map_file(env_name)} holds {xyz.shape[0]} points, more than "
                                           f"known_capacity = {cap} per env: construct the mapper with a larger known_capacity")
            self._known_cache[env_name] = (torch.from_numpy(xyz).to(self.device), torch.from_numpy(sem).to(self.device),
                                           o_r, o_c)
            while len(self._known_cache) > max(self.KNOWN_CACHE_SCENES, 1):
                self._known_cache.pop(next(iter(self._known_cache)))
        return self._known_cache[env_name]

    def _forward_known(self, eng, B, episodes_info, cs, pose, orient):
        lib, st = eng.lib, eng.stream()
        for b in [b for b in self._known_order if b >= B]:      # clear_paused_episodes, mapper.py:315-318
            _lib.check(lib.ivm_known_clear(eng.ctx, b, st), eng.ctx, "ivm_known_clear")
            self._known_order.remove(b)
            self._known_loaded.pop(b, None)
        # Which envs reload their scene cloud (mask == 0, mapper.py:873-878)?  A reload of the scene an env already
        # holds changes nothing, so the masks matter only where the scene differs from the loaded one (or nothing is
        # loaded yet): only then are they read -- a device sync if they live on the device; never in steady state.
        raw = episodes_info._names_raw
        last = self._known_names
        if last is not None and self._known_B == B and (type(raw) is list and raw == self._known_names_raw):
            finished = []          # every env still holds the scene it is in (the same list, or an equal one)
        elif last is not None and last.shape == episodes_info.env_names.shape and bool((last == episodes_info.env_names).all()):
            finished = []          # ... (one vectorised comparison)
            self._known_names_raw, self._known_B = (list(raw) if type(raw) is list else None), B
        else:
            env_names = episodes_info.env_names
            names = [str(n) for n in env_names[:B]]
            m = episodes_info.not_done_masks
            finished = [b for b in (m == episodes_info.EPISODE_FINISHED).nonzero().flatten().tolist()
                        if self._known_loaded.get(b) != names[b]]
        for b in finished:
            name = str(episodes_info.env_names[b])
            xyz, sem, o_r, o_c = self._known_cloud(name)
            _lib.check(lib.ivm_known_load(eng.ctx, b, xyz.shape[0], xyz.data_ptr(), sem.data_ptr(), o_r, o_c, st),
                       eng.ctx, "ivm_known_load")
            if b in self._known_order:
                self._known_order.remove(b)
            self._known_order.append(b)
            self._known_loaded[b] = name
        if finished or self._known_names is None:
            env_names = episodes_info.env_names
            ok = all(self._known_loaded.get(b) == str(env_names[b]) for b in range(B))
            self._known_names = np.array(env_names[:B], copy=True) if ok else None
            self._known_names_raw, self._known_B = (list(raw) if (ok and type(raw) is list) else None), B
        _lib.check(lib.ivm_step_known(eng.ctx, B, pose.data_ptr(), None if cs is None else cs.data_ptr(),
                                      None if orient is None else orient.data_ptr(),
                                      1 if (orient is not None and orient.dtype == torch.float64) else 0,
                                      eng.occ_ptr, eng.sem_ptr, st), eng.ctx, "ivm_step_known")

    # -- inspection
    def get_world_semantic_pointcloud(self) -> SemanticPointcloud:
        """reference mapper.py:946-947: the world cloud (batch_indices i64 [P], xyz f32 [P,3],
        semantics u8 [P]) in the reference's list order."""
        if self._engine is None or self._num_envs == 0:
            return SemanticPointcloud()
        eng, B = self._engine, self._num_envs
        if self.mode == "known":
            bs, xs, ss = [], [], []
            for b in self._known_order:
                xyz, sem, _, _ = self._known_cloud(self._known_loaded[b])
                bs.append(torch.full((xyz.shape[0],), b, dtype=torch.long, device=self.device))
                xs.append(xyz)
                ss.append(sem)
            if not xs:
                return SemanticPointcloud()
            return SemanticPointcloud(torch.cat(bs), torch.cat(xs), torch.cat(ss))
        with torch.cuda.device(self.device):
            _, stats = eng.status()
            cap = max(int(stats[2]), 1)
            env = torch.zeros(cap, dtype=torch.int64, device=self.device)
            xyz = torch.zeros((cap, 3), dtype=torch.float32, device=self.device)
            lab = torch.zeros(cap, dtype=torch.uint8, device=self.device)
            key = torch.zeros(cap, dtype=torch.int64, device=self.device)
            cnt = torch.zeros(1, dtype=torch.int64, device=self.device)
            _lib.check(eng.lib.ivm_export_world(eng.ctx, B, cap, env.data_ptr(), xyz.data_ptr(), lab.data_ptr(),
                                                key.data_ptr(), cnt.data_ptr(), eng.stream()), eng.ctx, "ivm_export_world")
            n = int(cnt.item())
            assert n <= cap, (n, cap)
            order = torch.sort(key[:n], stable=True).indices
            return SemanticPointcloud(env[:n][order], xyz[:n][order], lab[:n][order])

    def status(self):
        """(error_flags, stats) of the last step; synchronises the stream."""
        if self._engine is None:
            return 0, [0] * 8
        return self._engine.status()

    def check_errors(self):
        flags, _ = self.status()
        if flags & _lib.ERR_STORE_OVERFLOW:
            raise _lib.MapLibraryError("a point fell outside an env's world store window (raise store_cells)")
        if flags & _lib.ERR_EDGE_OVERFLOW:
            raise _lib.MapLibraryError("edge list capacity exceeded")
        if flags & _lib.ERR_KNOWN_OVERFLOW:
            raise _lib.MapLibraryError("known-map cloud outside the store window / over capacity")
        if flags & _lib.ERR_CAND_OVERFLOW:
            raise _lib.MapLibraryError("frame candidate table full")
        if flags & _lib.ERR_GRID_BARRIER:
            raise _lib.MapLibraryError("grid barrier time-out in the fused step kernel (results invalid)")

    def phase_ns(self) -> List[int]:
        """%globaltimer stamps (ns) of the last fused step: start, ingest done, resolve done, fix-up done,
        raster released, end.  Synchronises the stream."""
        out = (ctypes.c_uint64 * 24)()
        eng = self._engine
        _lib.check(eng.lib.ivm_read_phase_ns(eng.ctx, out, eng.stream()), eng.ctx, "ivm_read_phase_ns")
        return [int(v) for v in out][:8]

    def cta_trace_ns(self, num_ctas: int = 296):
        """Per-CTA %globaltimer stamps of the last fused step as an int64 array [num_ctas, 16] (see ivln_map.h)."""
        out = (ctypes.c_uint64 * (16 * num_ctas))()
        eng = self._engine
        _lib.check(eng.lib.ivm_read_cta_trace(eng.ctx, out, num_ctas, eng.stream()), eng.ctx, "ivm_read_cta_trace")
        return np.frombuffer(out, dtype=np.uint64).astype(np.int64).reshape(num_ctas, 16)

    def fixup_trace_ns(self) -> List[int]:
        """%globaltimer stamps (ns) of the milestones inside the last edge fix-up (see ivln_map.h)."""
        out = (ctypes.c_uint64 * 24)()
        eng = self._engine
        _lib.check(eng.lib.ivm_read_phase_ns(eng.ctx, out, eng.stream()), eng.ctx, "ivm_read_phase_ns")
        return [int(v) for v in out][8:]

    def kernel_launches(self) -> int:
        return 0 if self._engine is None else int(self._engine.lib.ivm_kernel_launches(self._engine.ctx))

    def set_timing(self, enabled: bool):
        _lib.check(self._engine.lib.ivm_set_timing(self._engine.ctx, 1 if enabled else 0))

    def stage_times(self, reset: bool = True):
        ms = (ctypes.c_float * 5)()
        n = (ctypes.c_int32 * 5)()
        _lib.check(self._engine.lib.ivm_stage_times(self._engine.ctx, ms, n, 1 if reset else 0))
        return list(ms), list(n)


# ----------------------------------------------------------------------------
# Factories (reference mapper.py:950-1028); extra keyword arguments are optional.
def create_iterative_mapper(device, camera_parameters: CameraParameters, map_dimensions: MapDimensions,
                            semantics_module: ComputeSemantics, **kw) -> MappingModule:
    return MappingModule(device, map_dimensions, camera_parameters, semantics_module, mode="iterative", **kw)


def create_known_mapper(device, map_dimensions: MapDimensions, maps_location, **kw) -> MappingModule:
    return MappingModule(device, map_dimensions, None, None, mode="known", maps_location=maps_location, **kw)


def create_gt_semantics_iterative_mapper(device, camera_parameters: CameraParameters,
                                         map_dimensions: MapDimensions, **kw) -> MappingModule:
    return create_iterative_mapper(device, camera_parameters, map_dimensions, GTSemantics(), **kw)


def create_predicted_semantics_iterative_mapper(device, camera_parameters: CameraParameters,
                                                map_dimensions: MapDimensions, semantics_module=None,
                                                **kw) -> MappingModule:
    return create_iterative_mapper(device, camera_parameters, map_dimensions,
                                   semantics_module if semantics_module is not None else PredictSemantics(), **kw)


def create_gt_semantics_known_mapper(device, map_dimensions: MapDimensions, **kw) -> MappingModule:
    return create_known_mapper(device, map_dimensions, kw.pop("maps_location", "data/known_maps/gt_semantics"), **kw)


def create_predicted_semantics_known_mapper(device, map_dimensions: MapDimensions, **kw) -> MappingModule:
    return create_known_mapper(device, map_dimensions,
                               kw.pop("maps_location", "data/known_maps/predicted_semantics"), **kw)
