"""f-2 / f-3 staging (ivlnce_b200/staging.py) against the restated reference functions (oracle/staging_oracle.py)
and -- in the build container -- the restatement against the unmodified reference source."""
import copy

import numpy as np
import pytest
import torch

from ivlnce_b200.staging import MapEgress, ObservationStager, batch_obs
from oracle.ref_loader import load_reference_function, reference_available
from oracle.staging_oracle import add_map_to_observations_oracle, batch_obs_oracle


def make_observations(B, H=32, W=32, seed=0, with_rgb=True):
    rng = np.random.default_rng(seed)
    obs = []
    for b in range(B):
        o = {
            "depth": rng.random((H, W, 1), dtype=np.float32),
            "semantic12": rng.integers(0, 13, (H, W, 1), dtype=np.uint8),
            "semantic": rng.integers(0, 2 ** 31, (H, W), dtype=np.uint32),      # uint32 -> int32 (utils.py:50-54)
            "world_robot_pose": rng.standard_normal(3).astype(np.float32),
            "world_robot_orientation": rng.standard_normal(2),                   # float64
            "instruction": rng.integers(0, 2000, 200, dtype=np.int64),
            "progress": np.float32(rng.random()),                                # 0-d
            "env_name": f"scene{b}",
        }
        if with_rgb:
            o["rgb"] = rng.integers(0, 255, (24, 24, 3), dtype=np.uint8)
        obs.append(o)
    return obs


def assert_same_batch(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in a:
        if isinstance(a[k], list):
            assert a[k] == b[k]
        else:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
            assert torch.equal(a[k].cpu(), b[k].cpu()), k


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container)")
def test_restated_batch_obs_matches_reference_source():
    ref = load_reference_function("ivlnce_baselines/common/utils.py", "batch_obs")
    for B in (1, 3):
        o1, o2 = make_observations(B, seed=B), make_observations(B, seed=B)
        assert_same_batch(ref(o1, torch.device("cpu")), batch_obs_oracle(o2, torch.device("cpu")))
        for x, y in zip(o1, o2):   # the in-place rewrite of the caller's dicts
            assert all(type(x[k]) is type(y[k]) for k in x)


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container)")
def test_restated_add_map_matches_reference_source():
    ref = load_reference_function("ivlnce_baselines/trainers/iterative_collection_dagger_trainer.py",
                                  "add_map_to_observations", class_name="IterativeCollectionDaggerTrainer")
    B = 3
    batch = {"occupancy_map": torch.randint(0, 2, (B, 8, 8), dtype=torch.uint8),
             "semantic_map": torch.randint(0, 13, (B, 8, 8), dtype=torch.uint8)}
    o1, o2 = make_observations(B), make_observations(B)
    r1 = ref(None, o1, batch, B)
    r2 = add_map_to_observations_oracle(o2, batch, B)
    assert [sorted(x.keys()) for x in r1] == [sorted(x.keys()) for x in r2]
    for x, y in zip(r1, r2):
        assert np.array_equal(x["occupancy_map"], y["occupancy_map"]) and np.array_equal(x["semantic_map"], y["semantic_map"])
    with pytest.raises(RuntimeError):
        ref(None, make_observations(B), {"occupancy_map": batch["occupancy_map"]}, B)


@pytest.mark.parametrize("B", [1, 4])
def test_batch_obs_host_matches_oracle(B):
    stager = ObservationStager()
    for rep in range(4):   # slab reuse
        o1, o2 = make_observations(B, seed=10 * B + rep), make_observations(B, seed=10 * B + rep)
        got = stager.batch_obs(o1, torch.device("cpu"))
        assert_same_batch(got, batch_obs_oracle(o2, torch.device("cpu")))
        assert all(isinstance(o1[0][k], torch.Tensor) for k in o1[0] if k != "env_name")
    got = batch_obs(make_observations(2), None)
    assert got["depth"].shape == (2, 32, 32, 1) and got["env_name"] == ["scene0", "scene1"]
    assert batch_obs([], torch.device("cpu")) == {}


def test_batch_obs_errors_like_stack():
    o = make_observations(2)
    o[1]["depth"] = np.zeros((16, 16, 1), np.float32)
    with pytest.raises(RuntimeError):
        ObservationStager().batch_obs(o, torch.device("cpu"))
    with pytest.raises(RuntimeError):
        batch_obs_oracle(copy.deepcopy(o), torch.device("cpu"))


def test_map_egress_host_matches_oracle():
    B = 3
    batch = {"occupancy_map": torch.randint(0, 2, (B, 8, 8), dtype=torch.uint8),
             "semantic_map": torch.randint(0, 13, (B, 8, 8), dtype=torch.uint8)}
    r1 = MapEgress().add_map_to_observations(make_observations(B), batch, B)
    r2 = add_map_to_observations_oracle(make_observations(B), batch, B)
    assert [sorted(x.keys()) for x in r1] == [sorted(x.keys()) for x in r2]
    for x, y in zip(r1, r2):
        assert np.array_equal(x["occupancy_map"], y["occupancy_map"]) and np.array_equal(x["semantic_map"], y["semantic_map"])
    with pytest.raises(RuntimeError):
        MapEgress().add_map_to_observations(make_observations(B), {"semantic_map": batch["semantic_map"]}, B)
    obs = make_observations(B)
    assert MapEgress().add_map_to_observations(obs, {}, B) is obs


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 5])
def test_batch_obs_device_matches_oracle(B):
    dev = torch.device("cuda:0")
    stager = ObservationStager(depth=2)
    outs = []
    for rep in range(6):   # ring reuse with copies in flight
        o1, o2 = make_observations(B, 64, 64, seed=rep), make_observations(B, 64, 64, seed=rep)
        got = stager.batch_obs(o1, dev)
        outs.append((got, batch_obs_oracle(o2, torch.device("cpu"))))
    torch.cuda.synchronize()
    for got, want in outs:     # earlier batches stay intact while later ones are staged
        assert all(v.device.type == "cuda" for k, v in got.items() if k != "env_name")
        assert_same_batch(got, want)
    assert stager.h2d_bytes == sum(v.numel() * v.element_size() for k, v in outs[-1][0].items() if k != "env_name")


@pytest.mark.gpu
def test_map_egress_device_through_plugin():
    """The plugin's maps leave through MapEgress exactly as the reference's per-env .cpu().numpy() would."""
    from ivlnce_b200.mapper import CameraParameters, MapDimensions
    from ivlnce_b200.obs_transforms import GTSemanticsIterativeMapper
    from ivlnce_b200.synthetic import ScenarioConfig, make_scenario, obs_dict_for_step

    dev = torch.device("cuda:0")
    cfg = ScenarioConfig(num_envs=3, height=64, width=64, steps=3, resolution=0.1, seed=5)
    scn = make_scenario(cfg)
    plugin = GTSemanticsIterativeMapper(CameraParameters(cfg.vfov_radians, (64, 64), 0.1), MapDimensions(6.4, 6.4, 0.1),
                                        store_cells=1024, max_envs=3)
    egress = MapEgress()
    for t in range(cfg.steps):
        obs = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in obs_dict_for_step(scn, t).items()}
        batch = plugin(obs)
        want = add_map_to_observations_oracle([{"env_name": "x"} for _ in range(3)], batch, 3)
        ticket = egress.start(batch)
        got = egress.add_map_to_observations([{"env_name": "x"} for _ in range(3)], batch, 3, ticket)
        for x, y in zip(got, want):
            assert np.array_equal(x["occupancy_map"], y["occupancy_map"]) and np.array_equal(x["semantic_map"], y["semantic_map"])
            assert "env_name" not in x
    assert egress.d2h_bytes == 2 * 3 * 64 * 64


def test_episode_record_keeps_the_msgpack_layout():
    """f-3: an episode packed here is what the reference's `save_episode_to_disk` stores
    (iterative_collection_dagger_trainer.py:60-80): msgpack-numpy's ndarray maps inside [obs, prev_actions, oracle_actions]."""
    import msgpack

    from ivlnce_b200.staging import pack_episode, unpack_episode

    rng = np.random.default_rng(3)
    T = 5
    episode = []
    for t in range(T):
        obs = {"rgb": rng.integers(0, 255, (4, 4, 3), dtype=np.uint8), "depth": rng.random((4, 4, 1), dtype=np.float32),
               "occupancy_map": rng.integers(0, 2, (8, 8), dtype=np.uint8), "semantic_map": rng.integers(0, 13, (8, 8), dtype=np.uint8),
               "expert": np.array([t % 4], dtype=np.int64)}
        episode.append((obs, t % 4, (t + 1) % 4))
    rec = pack_episode(episode, expert_uuid="expert")
    raw = msgpack.unpackb(rec, raw=True, strict_map_key=False)            # the wire layout, without the numpy hook
    assert isinstance(raw, list) and len(raw) == 3
    assert set(raw[0].keys()) == {b"rgb", b"depth", b"occupancy_map", b"semantic_map"}
    m = raw[0][b"occupancy_map"]
    assert m[b"nd"] is True and m[b"type"] == b"|u1" and m[b"kind"] == b"" and list(m[b"shape"]) == [T, 8, 8] and len(m[b"data"]) == T * 64
    assert raw[1][b"type"] == b"<i8" and list(raw[1][b"shape"]) == [T]
    obs, prev, oracle = unpack_episode(rec)
    for k in ("rgb", "depth", "occupancy_map", "semantic_map"):
        want = np.stack([np.asarray(step[0][k].cpu()) if hasattr(step[0][k], "cpu") else step[0][k] for step in episode])
        assert obs[k].dtype == want.dtype and np.array_equal(obs[k], want), k
    assert np.array_equal(prev, np.arange(T) % 4) and np.array_equal(oracle, (np.arange(T) + 1) % 4)
    half = unpack_episode(pack_episode(episode, expert_uuid="expert", lmdb_fp16=True))
    assert half[0]["depth"].dtype == np.float16 and half[0]["occupancy_map"].dtype == np.float16   # (the reference casts every key)
    try:
        import msgpack_numpy
    except ImportError:
        return
    ref = msgpack_numpy.unpackb(rec, raw=False)
    assert np.array_equal(ref[0]["semantic_map"], obs["semantic_map"])
