"""Multi-GPU partitioning of the map update: one process per GPU, envs (= tours) split into
static contiguous blocks, no data-path collective.

Each env's world state depends only on its own frames (reference mapper.py:406-413 keeps
clouds apart by batch index; resets are per env, mapper.py:320-326), so the path shards with
no exchange step.  NCCL (or gloo on CPU, for tests) is used only to gather per-rank metric
vectors and, on request, the ego maps; never inside a step.

Note: the reference couples the envs of one batch through the
batch-global bounding box of its de-dup key, so bit-exact parity is defined per shard --
against a reference run on the same env subset.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(total_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of envs owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(total_envs, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(total_envs: int, world_size: int) -> List[int]:
    return [shard_range(total_envs, world_size, r)[1] - shard_range(total_envs, world_size, r)[0]
            for r in range(world_size)]


def owner_of(env: int, total_envs: int, world_size: int) -> int:
    for r in range(world_size):
        s, e = shard_range(total_envs, world_size, r)
        if s <= env < e:
            return r
    raise IndexError(env)


def slice_obs_dict(obs: dict, start: int, stop: int) -> dict:
    """The rank-local part of a global obs dict (tensors sliced on dim 0, env_name list sliced)."""
    out = {}
    for k, v in obs.items():
        if torch.is_tensor(v):
            out[k] = v[start:stop]
        elif isinstance(v, (list, tuple)):
            out[k] = list(v[start:stop])
        else:
            out[k] = v
    return out


def _initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def gather_metrics(vec: torch.Tensor) -> torch.Tensor:
    """[world, n] matrix of every rank's metric vector (frames, seconds, bytes, checksums ...)."""
    if not _initialized():
        return vec.unsqueeze(0)
    world = dist.get_world_size()
    flat = vec.contiguous().reshape(-1)
    out = torch.empty(world * flat.numel(), dtype=vec.dtype, device=vec.device)
    dist.all_gather_into_tensor(out, flat)
    return out.reshape((world,) + tuple(vec.shape))


def gather_maps(local_maps: torch.Tensor, total_envs: int) -> torch.Tensor:
    """All ranks' ego maps u8 [B_r, R, C] concatenated in env order -> [total_envs, R, C].
    Shards may differ in size by one env; they are padded to the largest for the collective."""
    if not _initialized():
        return local_maps
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(total_envs, world)
    assert local_maps.shape[0] == sizes[rank], (local_maps.shape, sizes, rank)
    mx = max(sizes)
    pad = local_maps
    if local_maps.shape[0] < mx:
        pad = torch.zeros((mx,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
        pad[: local_maps.shape[0]] = local_maps
    out = torch.empty((world * mx,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
    dist.all_gather_into_tensor(out, pad.contiguous())
    parts = [out[r * mx: r * mx + sizes[r]] for r in range(world)]
    return torch.cat(parts, 0)


def map_checksum(maps: torch.Tensor) -> torch.Tensor:
    """Order-sensitive 64-bit checksum of a u8 map tensor (for cross-rank / cross-run comparison)."""
    flat = maps.reshape(-1).to(torch.int64)
    idx = torch.arange(flat.numel(), device=flat.device, dtype=torch.int64)
    return ((flat + 1) * ((idx % 1000003) + 1)).sum()


class ShardedMapper:
    """Runs one rank's block of envs through a mapping-module plugin.  `forward` takes the GLOBAL
    obs dict (as produced for all envs) and returns the rank-local result dict."""

    def __init__(self, plugin, total_envs: int, rank: int = None, world_size: int = None):
        self.plugin = plugin
        self.total_envs = total_envs
        self.rank = dist.get_rank() if rank is None and _initialized() else (rank or 0)
        self.world_size = dist.get_world_size() if world_size is None and _initialized() else (world_size or 1)
        self.start, self.stop = shard_range(total_envs, self.world_size, self.rank)

    def forward(self, global_obs: dict) -> dict:
        return self.plugin(slice_obs_dict(global_obs, self.start, self.stop))

    __call__ = forward
