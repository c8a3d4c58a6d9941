"""Host-side logic of the product package on CPU: geometry, dataclasses, the input adapter
and the plugin shim (with a stand-in mapping module; no CUDA here)."""
import math
import types

import numpy as np
import pytest
import torch

from ivlnce_b200 import _lib
from ivlnce_b200.geometry import camera_scale_tables, camera_to_world_rows, ego_rotation
from ivlnce_b200.mapper import (CameraParameters, EpisodesInfo, GTSemantics, MapDimensions, Observations,
                                PredictSemantics, RobotCurrentState, create_gt_semantics_iterative_mapper)
from ivlnce_b200.obs_transforms import (GTSemanticsIterativeMapper, Mapper, baseline_registry,
                                        get_active_obs_transforms)
from ivlnce_b200.setup_mapping_module import (extract_camera_parameters, extract_egocentric_map_parameters,
                                              setup_inputs_from_obs_dict)
from oracle import oracle as orc
from oracle.ref_loader import reference_available


def test_map_dimensions_ceil():
    md = MapDimensions(6.4, 6.4, 0.05)
    assert (md.num_rows, md.num_cols) == (128, 128)
    md = MapDimensions(6.4, 3.3, 0.1)
    assert (md.num_rows, md.num_cols) == (64, math.ceil(3.3 / 0.1))


def test_geometry_matches_oracle_restatement():
    xs, ys = camera_scale_tables(48, 64, 1.1)
    xo, yo = orc.camera_tables(48, 64, 1.1)
    assert np.array_equal(xs.numpy().view(np.uint32), xo.view(np.uint32))
    assert np.array_equal(ys.numpy().view(np.uint32), yo.view(np.uint32))
    g = torch.Generator().manual_seed(0)
    for dt in (torch.float64, torch.float32):
        pose = torch.randn(7, 3, generator=g)
        elev = (torch.randn(7, generator=g) * 0.1).to(dt)
        head = (torch.rand(7, generator=g) * 6.28 - 3.14).to(dt)
        T12 = camera_to_world_rows(pose, elev, head).numpy()
        T16 = orc.camera_to_world(pose, elev, head).reshape(-1, 16)[:, :12]
        assert np.array_equal(T12.view(np.uint32), T16.view(np.uint32))
        assert np.array_equal(ego_rotation(head).numpy().view(np.uint32), orc.ego_rotation(head).view(np.uint32))


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container only)")
def test_geometry_matches_reference_code():
    from oracle.ref_loader import load_reference_mapper

    ref = load_reference_mapper()
    g = torch.Generator().manual_seed(1)
    pose = torch.randn(5, 3, generator=g)
    for dt in (torch.float64, torch.float32):
        elev = torch.zeros(5, dtype=dt)
        head = (torch.rand(5, generator=g) * 6.28 - 3.14).to(dt)
        st = ref.RobotCurrentState(pose, elev, head)
        T_ref = st.get_camera_matrix()
        T12 = camera_to_world_rows(pose, elev, head)
        assert torch.equal(T_ref[:, :3, :].reshape(5, 12), T12)
        mine = RobotCurrentState(pose, elev, head).get_camera_matrix()
        assert torch.equal(mine, T_ref)
        R = ref.rotate_around_y_matrix(-head)
        cs = ego_rotation(head)
        assert torch.equal(R[:, 0, 0], cs[:, 0]) and torch.equal(R[:, 0, 2], cs[:, 1])
        assert torch.equal(R[:, 2, 0], -cs[:, 1]) and torch.equal(R[:, 2, 2], cs[:, 0])
    cam = ref.CameraParameters(1.3, (32, 48), 0.1)
    pc = ref.PointCloud(cam, batch_size=1, world_shift_origin=torch.zeros(3), device="cpu")
    xs, ys = camera_scale_tables(32, 48, 1.3)
    assert torch.equal(pc.x_scale[0, 0, :], xs) and torch.equal(pc.y_scale[0, :, 0], ys)


def test_episodes_info_and_adapter():
    B, H, W = 3, 4, 6
    obs = {
        "depth": torch.rand(B, H, W, 1),
        "semantic12": torch.randint(0, 13, (B, H, W, 1), dtype=torch.uint8),
        "world_robot_pose": torch.rand(B, 3),
        "world_robot_orientation": torch.rand(B, 2, dtype=torch.float64),
        "not_done_masks": torch.tensor([[0], [1], [0]], dtype=torch.uint8),
        "env_name": ["a", "b", "c"],
    }
    ei, o, st = setup_inputs_from_obs_dict(obs)
    assert ei.num_envs == 3 and ei.finished().tolist() == [True, False, True]
    assert ei.finished_indices().tolist() == [0, 2]
    assert o.depth_normalized.shape == (B, 1, H, W) and o.depth_normalized.is_contiguous()
    assert o.semantics.shape == (B, 1, H, W) and o.rgb is None
    assert st.elevation.dtype == torch.float64 and torch.equal(st.height, obs["world_robot_pose"][:, 1])


def test_config_extraction():
    depth = types.SimpleNamespace(HFOV=90, HEIGHT=256, WIDTH=256)
    emap = types.SimpleNamespace(height_clip=0.1, height_meters=6.4, width_meters=6.4, resolution_meters=0.1)
    cam = extract_camera_parameters(depth, emap)
    assert cam.vertical_fov_radians == pytest.approx(math.pi / 2) and cam.features_spatial_dimensions == (256, 256)
    md = extract_egocentric_map_parameters(emap)
    assert (md.num_rows, md.num_cols) == (64, 64)


def test_semantics_front_end_errors():
    with pytest.raises(Exception, match="Semantic Sensor not in use"):
        GTSemantics()(Observations(None, torch.zeros(1, 1, 2, 2), None))
    with pytest.raises(Exception, match="RGB Sensor not in use"):
        PredictSemantics().scores(Observations(None, torch.zeros(1, 1, 2, 2), None))


def test_no_cpu_fallback():
    mm = create_gt_semantics_iterative_mapper("cpu", CameraParameters(1.57, (8, 8), 0.1), MapDimensions(6.4, 6.4, 0.1))
    ei = EpisodesInfo(torch.zeros(1, 1, dtype=torch.uint8), ["a"])
    obs = Observations(torch.zeros(1, 1, 8, 8, dtype=torch.uint8), torch.zeros(1, 1, 8, 8), None)
    st = RobotCurrentState(torch.zeros(1, 3), torch.zeros(1), torch.zeros(1))
    with pytest.raises(_lib.MapLibraryError, match="no CPU fallback"):
        mm(ei, obs, st)


def test_plugin_shim_contract():
    for n in ("Mapper", "GTSemanticsIterativeMapper", "PredictedSemanticsIterativeMapper",
              "GTSemanticsKnownMapper", "PredictedSemanticsKnownMapper"):
        assert baseline_registry.get_obs_transformer(n) is not None
    cfg = types.SimpleNamespace(
        TASK_CONFIG=types.SimpleNamespace(SIMULATOR=types.SimpleNamespace(
            DEPTH_SENSOR=types.SimpleNamespace(HFOV=90, HEIGHT=8, WIDTH=8))),
        RL=types.SimpleNamespace(POLICY=types.SimpleNamespace(OBS_TRANSFORMS=types.SimpleNamespace(
            ENABLED_TRANSFORMS=["GTSemanticsIterativeMapper"],
            EGOCENTRIC_MAPPER=types.SimpleNamespace(height_clip=0.1, height_meters=6.4, width_meters=6.4,
                                                    resolution_meters=0.1)))),
        VIDEO_OPTION=[])
    (plugin,) = get_active_obs_transforms(cfg)
    assert isinstance(plugin, GTSemanticsIterativeMapper)
    space = types.SimpleNamespace(spaces={"depth": 1, "semantic12": 2, "world_robot_pose": 3, "env_name": 4})
    space = plugin.transform_observation_space(space)
    assert set(space.spaces) == {"depth", "occupancy_map", "semantic_map"}
    assert space.spaces["occupancy_map"].shape == (64, 64)
    with pytest.raises(NotImplementedError):
        Mapper(plugin.camera_parameters, plugin.map_dimensions).forward({})

    class FakeModule:  # stands in for the CUDA module: the shim only moves keys around
        def __call__(self, ei, obs, st):
            return types.SimpleNamespace(occupancy=torch.ones(ei.num_envs, 64, 64, dtype=torch.uint8),
                                         semantic=torch.zeros(ei.num_envs, 64, 64, dtype=torch.uint8))

    plugin.mapping_module = FakeModule()
    obs = {"depth": torch.rand(2, 8, 8, 1), "semantic12": torch.zeros(2, 8, 8, 1, dtype=torch.uint8),
           "world_robot_pose": torch.zeros(2, 3), "world_robot_orientation": torch.zeros(2, 2, dtype=torch.float64),
           "not_done_masks": torch.zeros(2, 1, dtype=torch.uint8), "env_name": ["a", "b"], "rgb": torch.zeros(2, 4, 4, 3)}
    out = plugin(obs)
    assert set(out) == {"depth", "rgb", "not_done_masks", "occupancy_map", "semantic_map"}
    assert out["occupancy_map"].shape == (2, 64, 64) and out["occupancy_map"].dtype == torch.uint8


@pytest.mark.parametrize("d", [0.05, 0.025, 0.1, 0.2, 0.07, 1.0 / 3.0])
def test_rint_of_quotient_without_division_is_exact(d):
    """ivm_rint_div (ivm_core.h) = rint(a / d) of the true fp32 division for EVERY float a with 2^-20 <= |a| <= 700 d
    (both signs; ~0.4 G values per divisor): the raster and the scatter use it instead of an IEEE division per point."""
    import ctypes

    from emu_wrapper import lib

    L = lib()
    L.emu_check_rint_div.restype = ctypes.c_longlong
    L.emu_check_rint_div.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.POINTER(ctypes.c_longlong)]
    amb = ctypes.c_longlong(0)
    bad = L.emu_check_rint_div(np.float32(d), np.float32(2.0 ** -20), np.float32(700.0 * d), ctypes.byref(amb))
    assert bad == 0
    assert 0 < amb.value < 4_000_000   # the fallback exists, and is rare
    # huge, tiny and non-finite numerators take the fallback or agree trivially
    for a in (0.0, 1e-38, 1e-45, 3.0e38, float("inf"), float("nan"), 1.0e9, 123456.789):
        amb2 = ctypes.c_longlong(0)
        assert L.emu_check_rint_div(np.float32(d), np.float32(a), np.float32(a), ctypes.byref(amb2)) == 0 or a != a
