"""Host <-> device staging either side of the map update.

f-2  `batch_obs` -- reference `ivlnce_baselines/common/utils.py:57-92`: a list of per-env observation dicts
     (numpy arrays from the simulator workers) becomes a dict of batched tensors on the device.  The reference
     builds a python list per sensor, `torch.stack`s it (a pageable host tensor) and calls a blocking
     `.to(device)` per sensor, every step.  Here each sensor has preallocated PINNED slabs `[B, ...]`; every env's
     array is copied straight into its row, and ONE asynchronous H2D copy per sensor runs on a copy stream while
     the previous step's kernels are still busy; the compute stream waits for an event, the host never blocks.
     Same signature, same keys, dtypes, shapes and values as the reference function.

f-3  map egress -- reference `ivlnce_baselines/trainers/iterative_collection_dagger_trainer.py:28-58`
     (`add_map_to_observations`; also `dagger_trainer.py:436-447`): per env and per map a `.cpu().numpy()`, i.e.
     2 B tiny blocking D2H copies per step (the reference even performs each of them twice).  Here both maps leave
     in ONE asynchronous D2H copy of a `[2, B, R, C]` pinned buffer and the per-env arrays are views of it (copied
     on request, as the reference's are fresh arrays).  Same dict layout afterwards.

Neither function contains map arithmetic; the device side is plain copies (torch owns memory and streams).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Set, Tuple

import numpy as np
import torch

MAP_KEYS = ("occupancy_map", "semantic_map")
# keys the mapper consumed and the trainers drop before storing an episode
# (iterative_collection_dagger_trainer.py:47-56, dagger_trainer.py:449-458)
CONSUMED_KEYS = ("semantic", "semantic12", "world_robot_pose", "world_robot_orientation", "env_name")


def _as_array(value) -> np.ndarray:
    """What `torch.as_tensor(value)` would see (utils.py:76-80): numpy uint32 becomes int32 first."""
    if isinstance(value, torch.Tensor):
        return value.detach().cpu().numpy()
    arr = np.asarray(value)
    if arr.dtype == np.uint32:
        arr = arr.astype(np.int32)  # utils.py:50-54, 77-78
    return arr


_TORCH_DTYPE = {
    np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.float16): torch.float16,
    np.dtype(np.uint8): torch.uint8, np.dtype(np.int8): torch.int8, np.dtype(np.int16): torch.int16,
    np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64, np.dtype(np.bool_): torch.bool,
}


class _Slab:
    """One pinned staging buffer of a sensor + the event of the last copy that read it."""

    __slots__ = ("host", "view", "event")

    def __init__(self, shape, dtype: torch.dtype, pin: bool):
        self.host = torch.empty(shape, dtype=dtype, pin_memory=pin)
        self.view = self.host.numpy()
        self.event: Optional[torch.cuda.Event] = None


class ObservationStager:
    """Drop-in for `batch_obs` (utils.py:57-92) that stages through pinned ring buffers.

        stager = ObservationStager()
        batch = stager.batch_obs(observations, device)      # same result as the reference function

    `depth` = number of pinned slabs per sensor (a slab is reused once the copy that read it has completed).
    The returned device tensors are freshly allocated every call, as in the reference."""

    def __init__(self, depth: int = 3, mutate_inputs: bool = True):
        self.depth = max(int(depth), 1)
        self.mutate_inputs = mutate_inputs   # replace the numpy arrays in the caller's dicts by tensors, as the reference does
        self._slabs: Dict[Tuple, List[_Slab]] = {}
        self._turn = 0
        self._copy_stream: Dict[torch.device, torch.cuda.Stream] = {}
        self.h2d_bytes = 0   # bytes of the last call (what bench.py reports)
        self._threads = None

    PARALLEL_SENSOR_BYTES = int(__import__("os").environ.get("IVM_STAGE_PAR", 1 << 20))

    def _pool(self):
        if self._threads is None:
            import os
            from concurrent.futures import ThreadPoolExecutor

            n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            self._threads = ThreadPoolExecutor(max_workers=max(1, min(8, n)))
        return self._threads

    def _slab(self, key, shape, dtype, pin) -> _Slab:
        k = (key, tuple(shape), dtype, pin)
        ring = self._slabs.get(k)
        if ring is None:
            ring = [_Slab(shape, dtype, pin) for _ in range(self.depth)]
            self._slabs[k] = ring
        slab = ring[self._turn % self.depth]
        if slab.event is not None:
            slab.event.synchronize()   # the H2D copy issued `depth` calls ago (long finished)
            slab.event = None
        return slab

    def batch_obs(self, observations: Sequence[Dict], device: Optional[torch.device] = None,
                  ignore_keys: Optional[Set[str]] = None, prebatched: Optional[Dict[str, torch.Tensor]] = None) -> Dict:
        """`prebatched` (an extension): already batched host tensors (pinned, for an asynchronous copy) that ride along on
        the same copy stream, e.g. the output of a host-side model."""
        if ignore_keys is None:
            ignore_keys = {"env_name"}
        device = torch.device(device) if device is not None else None
        on_gpu = device is not None and device.type == "cuda"
        B = len(observations)
        out: Dict = {}
        self._turn += 1
        self.h2d_bytes = 0
        if B == 0:
            return out
        stream = main = None
        if on_gpu:
            stream = self._copy_stream.get(device)
            if stream is None:
                stream = self._copy_stream[device] = torch.cuda.Stream(device)
            main = torch.cuda.current_stream(device)
        sensors = list(observations[0].keys())
        for obs in observations[1:]:           # the reference keeps every key that appears, in first-seen order
            for s in obs:
                if s not in sensors:
                    sensors.append(s)
        pending = []
        for sensor in sensors:
            rows = [obs[sensor] for obs in observations if sensor in obs]
            if sensor in ignore_keys:
                out[sensor] = rows
                continue
            first = _as_array(rows[0])
            tdtype = _TORCH_DTYPE.get(first.dtype)
            if tdtype is None or len(rows) != B:
                # exotic dtypes / ragged presence: the reference's own route (torch.stack raises the same errors)
                t = torch.stack([torch.as_tensor(_as_array(r)) for r in rows], dim=0)
                out[sensor] = t.to(device) if device is not None else t
                continue
            slab = self._slab(sensor, (B,) + tuple(first.shape), tdtype, on_gpu)
            arrays = [first] + [_as_array(r) for r in rows[1:]]
            for i, a in enumerate(arrays):
                if a.shape != first.shape or a.dtype != first.dtype:
                    raise RuntimeError(f"stack expects each tensor to be equal size, but got {tuple(first.shape)} at entry 0 "
                                       f"and {tuple(a.shape)} at entry {i}")
            if first.nbytes * B >= self.PARALLEL_SENSOR_BYTES and B > 1:
                # a sensor of a megabyte or more per step (depth, labels, RGB, score planes): the per-env copies run on a
                # few threads, a contiguous group of envs each (numpy releases the GIL while it copies)
                pool, view = self._pool(), slab.view
                groups = min(pool._max_workers, B)
                bounds = [(g * B // groups, (g + 1) * B // groups) for g in range(groups)]

                def copy_rows(lo_hi, view=view, arrays=arrays):
                    for i in range(lo_hi[0], lo_hi[1]):
                        view[i, ...] = arrays[i]

                list(pool.map(copy_rows, bounds))
            else:
                for i, a in enumerate(arrays):
                    slab.view[i, ...] = a
            if self.mutate_inputs:
                for i, (r, a) in enumerate(zip(rows, arrays)):
                    if not isinstance(r, torch.Tensor):
                        observations[i][sensor] = torch.as_tensor(a)   # the reference rewrites the caller's dicts (utils.py:79-80)
            if not on_gpu:
                out[sensor] = slab.host.clone()      # a fresh tensor, as torch.stack returns
                continue
            pending.append((sensor, slab, tdtype))
        if prebatched and on_gpu:
            for k, host_t in prebatched.items():
                sensors.append(k)
                pending.append((k, None, host_t))
        elif prebatched:
            for k, host_t in prebatched.items():
                sensors.append(k)
                out[k] = host_t
        if pending:
            # The device tensors are allocated and filled on the COPY stream (so the copies need not wait for the
            # kernels still queued on the compute stream) and handed to the compute stream with an event.
            with torch.cuda.stream(stream):
                for sensor, slab, tdtype in pending:
                    src = slab.host if slab is not None else tdtype      # (a prebatched host tensor travels in the third field)
                    dst = torch.empty(src.shape, dtype=src.dtype, device=device)
                    dst.copy_(src, non_blocking=True)
                    dst.record_stream(main)
                    out[sensor] = dst
                    self.h2d_bytes += dst.numel() * dst.element_size()
                ev = torch.cuda.Event()
                ev.record(stream)
            for _, slab, _ in pending:
                if slab is not None:
                    slab.event = ev
            main.wait_event(ev)                      # stream-ordered: later work on the compute stream sees the batch
            out = {k: out[k] for k in sensors if k in out}   # the reference's key order
        return out


_default_stager: Optional[ObservationStager] = None


def batch_obs(observations: Sequence[Dict], device: Optional[torch.device] = None,
              ignore_keys: Optional[Set[str]] = None) -> Dict:
    """Module-level drop-in with the reference's signature (utils.py:57-61); uses one shared stager."""
    global _default_stager
    if _default_stager is None:
        _default_stager = ObservationStager()
    return _default_stager.batch_obs(observations, device, ignore_keys)


class MapEgress:
    """Both ego maps of a step to the host in ONE asynchronous copy (f-3).

        egress = MapEgress()
        ticket = egress.start(batch)                       # enqueue the D2H copy (no host wait)
        ...                                                # e.g. policy.act on the same batch
        observations = egress.add_map_to_observations(observations, batch, num_envs, ticket)

    `add_map_to_observations` has the reference's name, arguments and effect
    (iterative_collection_dagger_trainer.py:28-58): observations[i]["occupancy_map" / "semantic_map"] become numpy
    uint8 [R, C] arrays and the keys the mapper consumed are deleted.  Without a ticket it starts the copy itself."""

    def __init__(self, depth: int = 2, copy: bool = True):
        self.depth = max(int(depth), 1)
        self.copy = copy          # hand out copies (the reference's arrays are fresh) or views of the pinned buffer
        self._bufs: Dict[Tuple, List] = {}
        self._turn = 0
        self.d2h_bytes = 0

    def start(self, batch: Dict):
        present = [k in batch for k in MAP_KEYS]
        if any(present) and not all(present):
            raise RuntimeError("either both map keys should exist in the batch or neither")
        if not all(present):
            return None
        occ, sem = batch[MAP_KEYS[0]], batch[MAP_KEYS[1]]
        if occ.device.type != "cuda":
            return (occ.detach().numpy(), sem.detach().numpy(), None)
        key = (tuple(occ.shape), occ.dtype, occ.device)
        ring = self._bufs.get(key)
        if ring is None:
            ring = [[torch.empty((2,) + tuple(occ.shape), dtype=occ.dtype, pin_memory=True), None] for _ in range(self.depth)]
            self._bufs[key] = ring
        self._turn += 1
        slot = ring[self._turn % self.depth]
        if slot[1] is not None:
            slot[1].synchronize()
        host = slot[0]
        # the mapping module keeps both maps in one [2, max_envs, R, C] block: a full batch is ONE contiguous copy,
        # a partial batch two (still asynchronous, still no per-env copies)
        n = occ.numel() * occ.element_size()
        if (occ.is_contiguous() and sem.is_contiguous() and sem.data_ptr() == occ.data_ptr() + n
                and occ.untyped_storage().data_ptr() == sem.untyped_storage().data_ptr()):
            both = torch.as_strided(occ, (2,) + tuple(occ.shape), (occ.numel(),) + tuple(occ.stride()))
            host.copy_(both, non_blocking=True)
        else:
            host[0].copy_(occ, non_blocking=True)
            host[1].copy_(sem, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(occ.device))
        slot[1] = ev
        self.d2h_bytes = 2 * occ.numel() * occ.element_size()
        return (host[0].numpy(), host[1].numpy(), ev)

    def add_map_to_observations(self, observations: List[Dict], batch: Dict, num_envs: int, ticket=None):
        if ticket is None:
            ticket = self.start(batch)
            if ticket is None:
                return observations
        occ, sem, ev = ticket
        if ev is not None:
            ev.synchronize()
        for i in range(num_envs):
            observations[i][MAP_KEYS[0]] = occ[i].copy() if self.copy else occ[i]
            observations[i][MAP_KEYS[1]] = sem[i].copy() if self.copy else sem[i]
            for k in CONSUMED_KEYS:
                if k in observations[i]:
                    del observations[i][k]
        return observations


# ------------------------------------------------------------------------------------------------------------------
# Episode records of the collection trainers ("keep the msgpack layout").
# reference `iterative_collection_dagger_trainer.py:60-80` (`save_episode_to_disk`; `dagger_trainer.py:258-270` reads it
# back): an episode = [ {sensor: array [T, ...]}, prev_actions i64 [T], oracle_actions i64 [T] ], written as
# `msgpack_numpy.packb(..., use_bin_type=True)`.  msgpack-numpy (0.4.x, not installed here) encodes an ndarray as the map
# {b"nd": True, b"type": dtype.str, b"kind": b"", b"shape": shape, b"data": raw bytes}; the two functions below write
# and read exactly that with plain msgpack, so records are interchangeable with the reference's LMDB files.
def _encode_ndarray(obj):
    if isinstance(obj, np.ndarray):
        if obj.dtype.kind in "VO":
            raise TypeError(f"cannot pack arrays of dtype {obj.dtype}")
        return {b"nd": True, b"type": obj.dtype.str, b"kind": b"", b"shape": obj.shape, b"data": obj.tobytes()}
    if isinstance(obj, (np.bool_, np.number)):
        return {b"nd": False, b"type": obj.dtype.str, b"data": obj.tobytes()}
    raise TypeError(f"cannot pack {type(obj)}")


def _decode_ndarray(obj):
    if b"nd" in obj:
        if obj[b"nd"] is True:
            return np.frombuffer(obj[b"data"], dtype=np.dtype(obj[b"type"])).reshape(obj[b"shape"])
        return np.frombuffer(obj[b"data"], dtype=np.dtype(obj[b"type"]))[0]
    return obj


def pack_episode(episode: Sequence[Tuple[Dict, int, int]], expert_uuid: Optional[str] = None, lmdb_fp16: bool = False) -> bytes:
    """`save_episode_to_disk` without the LMDB transaction: `episode` = [(observation dict, prev_action, oracle_action)]
    per step (observations as `MapEgress.add_map_to_observations` leaves them); returns the value the reference stores
    under `str(lmdb_idx).encode()`."""
    import msgpack

    traj_obs = batch_obs([step[0] for step in episode], device=torch.device("cpu"))
    if expert_uuid is not None:
        del traj_obs[expert_uuid]
    for k, v in traj_obs.items():
        traj_obs[k] = v.numpy()
        if lmdb_fp16:
            traj_obs[k] = traj_obs[k].astype(np.float16)
    transposed_ep = [traj_obs, np.array([step[1] for step in episode], dtype=np.int64),
                     np.array([step[2] for step in episode], dtype=np.int64)]
    return msgpack.packb(transposed_ep, default=_encode_ndarray, use_bin_type=True)


def unpack_episode(record: bytes):
    """`msgpack_numpy.unpackb(record, raw=False)` (dagger_trainer.py:258-270): [obs dict of arrays, prev_actions, oracle_actions]."""
    import msgpack

    return msgpack.unpackb(record, object_hook=_decode_ndarray, raw=False, strict_map_key=False)
