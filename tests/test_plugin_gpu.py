"""The obs-transform plugin under all four registered names, the segmentation front end and the
start-state bookkeeping (row a4), on the GPU through the public interfaces."""
import math
import os
import tempfile
import types

import numpy as np
import pytest
import torch

from ivlnce_b200.mapper import (CameraParameters, EpisodesInfo, LocalizeRobot, MapDimensions, Observations,
                                PredictSemantics, RobotCurrentState)
from ivlnce_b200.obs_transforms import baseline_registry
from ivlnce_b200.synthetic import ScenarioConfig, make_known_cloud, make_scenario, obs_dict_for_step
from oracle.oracle import OracleMapper, argmax_labels
from oracle.ref_loader import load_reference_mapper, reference_available


class StubSegmenter(torch.nn.Module):
    """Deterministic stand-in for RedNet: 13 class-score planes from the normalised rgb and depth."""

    def __init__(self, classes=13):
        super().__init__()
        self.classes = classes
        self.last_scores = None
        self.calls = 0

    def forward(self, rgb, depth):
        k = torch.arange(self.classes, device=rgb.device, dtype=torch.float32).view(1, -1, 1, 1)
        s = torch.sin(rgb.mean(1, keepdim=True) * (k + 1.0)) + torch.cos(depth * (0.5 * k + 0.3))
        self.last_scores = s
        self.calls += 1
        return s


def _rgb_for(scn, t, B, h=48, w=40):
    rng = np.random.default_rng(100 + t)
    return torch.from_numpy(rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [((224, 224), (256, 256)), ((48, 40), (64, 64)), ((300, 280), (96, 128)), ((64, 64), (64, 64))])
def test_preprocess_kernel_matches_torch_ops(shape):
    """rgb / 255 -> bilinear resize -> normalise, depth normalise: one kernel vs the reference's torch ops
    (mapper.py:715-736, 788-793), up- and down-scaling, NHWC-as-NCHW view and contiguous NCHW.  Tolerance 1e-5."""
    (h, w), (H, W) = shape
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(h * 1000 + W)
    rgb_nhwc = torch.randint(0, 256, (3, h, w, 3), generator=g, dtype=torch.uint8).to(dev)
    depth = torch.rand((3, 1, H, W), generator=g).to(dev)
    ps = PredictSemantics(model=StubSegmenter())
    for rgb in (rgb_nhwc.permute(0, 3, 1, 2), rgb_nhwc.permute(0, 3, 1, 2).contiguous()):
        obs = Observations(None, depth, rgb)
        r1, d1 = ps.preprocess(obs)
        r0, d0 = ps.preprocess_torch(obs)
        assert r1.shape == r0.shape and d1.shape == d0.shape and r1.is_contiguous()
        assert float((r1 - r0).abs().max()) <= 1e-5, float((r1 - r0).abs().max())
        assert float((d1 - d0).abs().max()) <= 1e-5


def _check_against_oracle(outs, ref):
    for t, ((o, s), (o_ref, s_ref)) in enumerate(zip(outs, ref)):
        assert np.array_equal(o, o_ref), f"occupancy differs at step {t}"
        assert np.array_equal(s, s_ref), f"semantic map differs at step {t}"


@pytest.mark.gpu
def test_gt_and_predicted_iterative_plugins_by_registered_name():
    dev = torch.device("cuda:0")
    cfg = ScenarioConfig(num_envs=3, height=64, width=64, steps=5, resolution=0.1, reset_steps={3: [1]}, seed=11)
    scn = make_scenario(cfg)
    cam, md = CameraParameters(cfg.vfov_radians, (64, 64), 0.1), MapDimensions(6.4, 6.4, 0.1)
    # ---- GT
    cls = baseline_registry.get_obs_transformer("GTSemanticsIterativeMapper")
    plugin = cls(cam, md, store_cells=1024, host_trig=True)
    orc = OracleMapper(64, 64, cfg.vfov_radians, 6.4, 6.4, 0.1)
    outs, ref = [], []
    for t in range(cfg.steps):
        obs = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in obs_dict_for_step(scn, t).items()}
        out = plugin(obs)
        assert "semantic12" not in out and "world_robot_pose" not in out and "env_name" not in out and "depth" in out
        outs.append((out["occupancy_map"].cpu().numpy(), out["semantic_map"].cpu().numpy()))
        ref.append(orc.step(scn["masks"][t], scn["pose"][t], scn["orientation"][t], depth=scn["depth"][t], labels=scn["labels"][t]))
    _check_against_oracle(outs, ref)
    # ---- predicted: rgb -> front-end kernel -> stub network -> scores -> argmax inside the step kernel
    cls = baseline_registry.get_obs_transformer("PredictedSemanticsIterativeMapper")
    stub = StubSegmenter()
    plugin = cls(cam, md, store_cells=1024, host_trig=True, semantics_module=PredictSemantics(model=stub))
    orc = OracleMapper(64, 64, cfg.vfov_radians, 6.4, 6.4, 0.1)
    outs, ref = [], []
    for t in range(cfg.steps):
        obs = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in obs_dict_for_step(scn, t).items()}
        del obs["semantic12"]
        obs["rgb"] = _rgb_for(scn, t, 3).to(dev)
        out = plugin(obs)
        labels = argmax_labels(stub.last_scores.cpu().numpy())            # of the scores the module really used
        outs.append((out["occupancy_map"].cpu().numpy(), out["semantic_map"].cpu().numpy()))
        ref.append(orc.step(scn["masks"][t], scn["pose"][t], scn["orientation"][t], depth=scn["depth"][t], labels=labels))
    assert stub.calls == cfg.steps
    _check_against_oracle(outs, ref)
    plugin.mapping_module.check_errors()


@pytest.mark.gpu
def test_predicted_plugin_model_factory_hooks(monkeypatch):
    """`from_config` route: the registered PredictedSemanticsIterativeMapper builds its network lazily from a factory
    (explicit, or named by IVLN_SEMANTICS_MODEL_FACTORY); without any it raises on the first forward with a clear text."""
    dev = torch.device("cuda:0")
    cfg = ScenarioConfig(num_envs=2, height=32, width=32, steps=2, resolution=0.1, seed=12)
    scn = make_scenario(cfg)
    mapper_cfg = types.SimpleNamespace(resolution_meters=0.1, height_clip=0.1, height_meters=6.4, width_meters=6.4,
                                       B200_MAPPER={"store_cells": 1024})
    config = types.SimpleNamespace(
        RL=types.SimpleNamespace(POLICY=types.SimpleNamespace(OBS_TRANSFORMS=types.SimpleNamespace(EGOCENTRIC_MAPPER=mapper_cfg))),
        TASK_CONFIG=types.SimpleNamespace(SIMULATOR=types.SimpleNamespace(DEPTH_SENSOR=types.SimpleNamespace(HFOV=90, HEIGHT=32, WIDTH=32))),
        VIDEO_OPTION=[])
    cls = baseline_registry.get_obs_transformer("PredictedSemanticsIterativeMapper")

    def obs_at(t):
        obs = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in obs_dict_for_step(scn, t).items()}
        del obs["semantic12"]
        obs["rgb"] = _rgb_for(scn, t, 2, 24, 24).to(dev)
        return obs

    monkeypatch.delenv("IVLN_SEMANTICS_MODEL_FACTORY", raising=False)
    plugin = cls.from_config(config)
    with pytest.raises(Exception, match="needs a segmentation model"):
        plugin(obs_at(0))
    monkeypatch.setenv("IVLN_SEMANTICS_MODEL_FACTORY", "test_plugin_gpu:make_stub")
    plugin = cls.from_config(config)
    out = plugin(obs_at(0))
    assert out["semantic_map"].shape == (2, 64, 64) and out["semantic_map"].dtype == torch.uint8
    assert isinstance(plugin.mapping_module.compute_semantics.model, StubSegmenter)


def make_stub(device):
    return StubSegmenter().to(device)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["GTSemanticsKnownMapper", "PredictedSemanticsKnownMapper"])
def test_known_map_plugins_by_registered_name(name):
    dev = torch.device("cuda:0")
    T, B = 6, 2
    cfg = ScenarioConfig(num_envs=B, height=8, width=8, steps=T, resolution=0.1, env_spacing=0.0, seed=21)
    scn = make_scenario(cfg)
    known = {f"scene{i}": make_known_cloud(5000, 14.0, 13, seed=300 + i) for i in range(3)}
    names = [["scene0", "scene1"]] * 3 + [["scene0", "scene2"]] * 3
    scn["masks"][3, 1] = 0
    with tempfile.TemporaryDirectory() as tmp:
        for k, (xyz, sem) in known.items():
            np.savez(os.path.join(tmp, f"{k}.npz"), xyz=xyz, semantics=sem)
        cls = baseline_registry.get_obs_transformer(name)
        plugin = cls(None, MapDimensions(6.4, 6.4, 0.1), store_cells=1024, host_trig=True, maps_location=tmp,
                     known_capacity=1 << 14)
        orc = OracleMapper(8, 8, cfg.vfov_radians, 6.4, 6.4, 0.1, mode="known", known_clouds=known)
        outs, ref = [], []
        for t in range(T):
            obs = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in obs_dict_for_step(scn, t, env_names=names[t]).items()}
            out = plugin(obs)
            outs.append((out["occupancy_map"].cpu().numpy(), out["semantic_map"].cpu().numpy()))
            ref.append(orc.step(scn["masks"][t], scn["pose"][t], scn["orientation"][t], env_names=names[t]))
        _check_against_oracle(outs, ref)
        plugin.mapping_module.check_errors()


def _start_state_sequence():
    rng = np.random.default_rng(3)
    seq = []
    for B, zeros in [(3, [0, 1, 2]), (3, []), (3, [1]), (2, []), (2, [0]), (4, [3]), (4, [])]:
        m = np.ones((B, 1), dtype=np.uint8)
        m[zeros] = 0
        seq.append((torch.from_numpy(m), torch.from_numpy(rng.standard_normal((B, 3)).astype(np.float32)),
                    torch.from_numpy(rng.standard_normal(B)), torch.from_numpy(rng.standard_normal(B))))
    return seq


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container)")
def test_start_state_matches_reference():
    """Row a4: LocalizeRobot / RobotStartState (mapper.py:141-192) against the unmodified reference classes on a
    sequence with resets, batch shrink and batch grow (the grown state starts from zeros, as in the reference)."""
    ref = load_reference_mapper()
    ours, theirs = LocalizeRobot(), ref.LocalizeRobot()
    for m, pose, elev, head in _start_state_sequence():
        names = [f"s{i}" for i in range(m.shape[0])]
        a = ours(EpisodesInfo(m.clone(), names), RobotCurrentState(pose, elev, head))
        b = theirs(ref.EpisodesInfo(m.clone(), names), ref.RobotCurrentState(pose, elev, head))
        assert torch.equal(a.start_state.pose, b.start_state.pose)
        assert torch.equal(a.start_state.elevation, b.start_state.elevation)
        assert torch.equal(a.start_state.heading, b.start_state.heading)
        assert a.start_state.batch_size == b.start_state.batch_size
        assert a.current_state.pose is pose


@pytest.mark.gpu
def test_start_state_tracked_on_the_device():
    from ivlnce_b200.mapper import create_gt_semantics_iterative_mapper

    dev = torch.device("cuda:0")
    cfg = ScenarioConfig(num_envs=2, height=32, width=32, steps=4, resolution=0.1, reset_steps={2: [1]}, seed=31)
    scn = make_scenario(cfg)
    mm = create_gt_semantics_iterative_mapper(dev, CameraParameters(cfg.vfov_radians, (32, 32), 0.1), MapDimensions(6.4, 6.4, 0.1),
                                              store_cells=1024, track_start_state=True)
    want = np.zeros((2, 3), np.float32)
    for t in range(cfg.steps):
        o = obs_dict_for_step(scn, t)
        ori = o["world_robot_orientation"].to(dev)
        mm(EpisodesInfo(o["not_done_masks"].to(dev), o["env_name"]),
           Observations(o["semantic12"].to(dev).permute(0, 3, 1, 2), o["depth"].to(dev).permute(0, 3, 1, 2), None),
           RobotCurrentState(o["world_robot_pose"].to(dev), ori[:, 0], ori[:, 1]))
        for b in range(2):
            if scn["masks"][t, b] == 0:
                want[b] = scn["pose"][t, b]
        assert np.array_equal(mm.localize_robot.start_state.pose.cpu().numpy(), want)
