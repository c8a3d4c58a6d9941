"""Env/tour partitioning and the metric/map gathers, world_size 2 over gloo on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ivlnce_b200.sharding import (ShardedMapper, gather_maps, gather_metrics, map_checksum, owner_of, shard_range,
                                  shard_sizes, slice_obs_dict)


def test_shard_ranges_cover_everything():
    for total in (1, 7, 16, 256, 257):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(total, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert max(shard_sizes(total, world)) - min(shard_sizes(total, world)) <= 1
    assert shard_range(256, 8, 3) == (96, 128)
    assert owner_of(100, 256, 8) == 3


def test_slice_obs_dict():
    obs = {"depth": torch.arange(12.).reshape(6, 2), "env_name": list("abcdef"), "flag": 3}
    s = slice_obs_dict(obs, 2, 5)
    assert s["depth"].shape == (3, 2) and s["env_name"] == ["c", "d", "e"] and s["flag"] == 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_envs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, stop = shard_range(total_envs, world, rank)

        class FakePlugin:  # stands in for the CUDA plugin: map cell value = global env index + 1
            def __call__(self, obs):
                ids = obs["ids"]
                return {"occupancy_map": (ids + 1).to(torch.uint8).view(-1, 1, 1).expand(-1, 4, 4).contiguous()}

        sm = ShardedMapper(FakePlugin(), total_envs)
        assert (sm.start, sm.stop) == (start, stop)
        out = sm({"ids": torch.arange(total_envs), "env_name": [str(i) for i in range(total_envs)]})
        local = out["occupancy_map"]
        assert local.shape[0] == stop - start
        full = gather_maps(local, total_envs)
        expect = (torch.arange(total_envs) + 1).to(torch.uint8).view(-1, 1, 1).expand(-1, 4, 4)
        assert torch.equal(full, expect)
        m = gather_metrics(torch.tensor([float(stop - start), float(rank)], dtype=torch.float64))
        assert m.shape == (world, 2) and m[:, 0].sum().item() == total_envs
        q.put((rank, int(map_checksum(full))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total_envs", [8, 5])
def test_gather_two_ranks_gloo(total_envs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total_envs, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    sums = dict(q.get(timeout=10) for _ in range(2))
    assert sums[0] == sums[1]
