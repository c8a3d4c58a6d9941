"""The dense-store + edge-fix-up algorithm (the kernels' per-thread logic from
ivlnce_b200/csrc/ivm_core.h, run serially by tests/emu/emulator.cpp) against the golden
fixtures of the unmodified reference and against the oracle on random scenarios.
This is how the CUDA algorithm is validated on a box without a GPU; the emulator is test
infrastructure and is never used by the product."""
import numpy as np
import pytest

from emu_wrapper import EmuMapper
from golden_io import golden_names, load_golden
from ivlnce_b200.synthetic import ScenarioConfig, make_scenario
from oracle.oracle import OracleMapper
from scenarios import run_mapper


def _emu(scn, order=0, tile=32, **kw):
    c = scn["cfg"]
    return EmuMapper(c["height"], c["width"], c["vfov"], c["map_m"], c["resolution"],
                     max_envs=scn["masks"].shape[1], mode=c["mode"], known_clouds=scn.get("known"),
                     order=order, tile=tile, store=2048 if c["resolution"] < 0.1 else 1024, **kw)


@pytest.mark.parametrize("order", [0, 2])
@pytest.mark.parametrize("name", golden_names())
def test_emulator_matches_golden(name, order):
    scn = load_golden(name)
    emu = _emu(scn, order=order)
    iterative = scn["cfg"]["mode"] == "iterative"
    outs, sizes = run_mapper(emu.step, scn, world_fn=emu.world if iterative else None)
    for t, (o, s) in enumerate(outs):
        B = o.shape[0]
        assert np.array_equal(o, scn["ref_occupancy"][t, :B]), f"occupancy differs at step {t}"
        assert np.array_equal(s, scn["ref_semantic"][t, :B]), f"semantic differs at step {t}"
    err, _ = emu.status()
    assert err == 0
    if iterative:
        assert sizes == scn["ref_world_sizes"].tolist()
        b, xyz, sem = emu.world()
        assert np.array_equal(b, scn["ref_world_b"])
        assert np.array_equal(xyz.view(np.uint32), scn["ref_world_xyz"].view(np.uint32))  # same bits, same order
        assert np.array_equal(sem, scn["ref_world_sem"])


@pytest.mark.parametrize("name", ["iid_f64", "scene_overlap", "batch_shrink_grow"])
def test_candidate_table_stamp_wrap(name):
    """The frame candidate plane is never cleared between steps: its words carry a step stamp that
    wraps (65 535 steps at 256x256; the period is shortened here).  Start just below the wrap so that it
    happens in mid-run."""
    scn = load_golden(name)
    emu = _emu(scn)
    emu.set_stamp_period(255)
    emu.set_step(255 * 3 - 4)
    outs, sizes = run_mapper(emu.step, scn, world_fn=emu.world)
    for t, (o, s) in enumerate(outs):
        B = o.shape[0]
        assert np.array_equal(o, scn["ref_occupancy"][t, :B]) and np.array_equal(s, scn["ref_semantic"][t, :B]), t
    assert sizes == scn["ref_world_sizes"].tolist()
    assert emu.cand_current() > 0
    assert emu.status()[0] == 0


@pytest.mark.parametrize("name", ["degenerate", "identical_envs", "scene_f32", "scene_overlap"])
def test_emulator_hash_path_and_kernel_trig(name):
    """Same goldens through the hash-based class resolution (the small-class fast path disabled) and
    with the pose matrices derived by the K1 code path (ivm_pose_matrices) where angles are float64."""
    scn = load_golden(name)
    f64 = scn["orientation"].dtype == np.float64
    emu = _emu(scn, order=1, fix_cap=1, kernel_trig=f64)
    outs, sizes = run_mapper(emu.step, scn, world_fn=emu.world)
    for t, (o, s) in enumerate(outs):
        B = o.shape[0]
        assert np.array_equal(o, scn["ref_occupancy"][t, :B]) and np.array_equal(s, scn["ref_semantic"][t, :B]), t
    assert sizes == scn["ref_world_sizes"].tolist()
    b, xyz, sem = emu.world()
    assert np.array_equal(b, scn["ref_world_b"]) and np.array_equal(xyz.view(np.uint32), scn["ref_world_xyz"].view(np.uint32))


def test_edge_collisions_are_exercised():
    merged = 0
    for name in ("degenerate", "identical_envs", "scene_f32"):
        scn = load_golden(name)
        emu = _emu(scn)
        run_mapper(emu.step, scn)
        merged += int(emu.status()[1][6])
    assert merged >= 50  # records deleted by the reference's key-collision quirk


def _fuzz_case(seed):
    rng = np.random.default_rng(seed)
    B = int(rng.integers(1, 5)); H = int(rng.choice([8, 12, 16, 24])); W = int(rng.choice([8, 16, 20, 32]))
    T = int(rng.integers(4, 12)); res = float(rng.choice([0.1, 0.05, 0.2]))
    mode = str(rng.choice(["iid", "scene", "narrow", "quant"]))
    cfg = ScenarioConfig(num_envs=B, height=H, width=W, steps=T, resolution=res,
                         depth_mode="scene" if mode == "scene" else "iid",
                         env_spacing=float(rng.choice([0.0, 0.0, 0.5, 3.0])),
                         angle_dtype=str(rng.choice(["float64", "float32"])), seed=seed,
                         forward_step=float(rng.choice([0.25, 0.05, 0.0])),
                         turn_degrees=float(rng.choice([15.0, 0.0, 90.0])))
    cfg.vfov_radians = (np.pi / 2) * min(1.0, H / W)
    scn = make_scenario(cfg)
    if mode == "narrow":
        scn["depth"] = (0.1 + 0.05 * scn["depth"]).astype(np.float32)
    if mode == "quant":
        scn["depth"] = (np.round(scn["depth"] * 8) / 8).astype(np.float32)
    if rng.random() < 0.3:  # axis-aligned headings: exact height ties, straight walls
        q = np.pi / 2
        scn["orientation"][..., 1] = (np.round(scn["orientation"][..., 1] / q) * q).astype(scn["orientation"].dtype)
    scn["masks"][rng.random(scn["masks"].shape) < 0.1] = 0
    if rng.random() < 0.5:
        scn["depth"][rng.random(scn["depth"].shape) < rng.choice([0.3, 0.9, 0.99])] = 1.0
    return cfg, scn, int(rng.integers(0, 3)), int(rng.choice([8, 16, 32, 64]))


@pytest.mark.parametrize("seed", range(40))
def test_emulator_matches_oracle_fuzz(seed):
    cfg, scn, order, tile = _fuzz_case(10_000 + seed)
    orc = OracleMapper(cfg.height, cfg.width, cfg.vfov_radians, cfg.map_meters, cfg.map_meters, cfg.resolution)
    emu = EmuMapper(cfg.height, cfg.width, cfg.vfov_radians, cfg.map_meters, cfg.resolution, max_envs=cfg.num_envs,
                    store=2048 if cfg.resolution < 0.1 else 1024, order=order, tile=tile)
    for t in range(cfg.steps):
        a = (scn["masks"][t], scn["pose"][t], scn["orientation"][t])
        kw = dict(depth=scn["depth"][t], labels=scn["labels"][t])
        o1, s1 = orc.step(*a, **kw)
        o2, s2 = emu.step(*a, **kw)
        assert emu.status()[0] == 0
        assert np.array_equal(o1, o2) and np.array_equal(s1, s2), f"maps differ at step {t}"
        b1, x1, m1 = orc.world()
        b2, x2, m2 = emu.world()
        assert np.array_equal(b1, b2) and np.array_equal(x1.view(np.uint32), x2.view(np.uint32)) and np.array_equal(m1, m2)
        _, stats = emu.status()
        assert int(stats[0]) == orc.counters["n_valid"] and int(stats[1]) == orc.counters["n_local"]
        assert int(stats[2]) == orc.counters["n_world"] and int(stats[3]) == orc.counters["n_in"]


@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("name", [n for n in golden_names() if n != "known_map"])
def test_emulator_direct_edge_resolution_matches_golden(name, order):
    """Direct resolution of the edge collisions (every edge cell looks up the <= 5 cells that share its key instead
    of grouping entries by key; ivm_partners / ivm_frame_edge_loses / ivm_world_edge_loses) against the goldens."""
    scn = load_golden(name)
    emu = _emu(scn, order=order, direct=True)
    outs, sizes = run_mapper(emu.step, scn, world_fn=emu.world)
    for t, (o, s) in enumerate(outs):
        B = o.shape[0]
        assert np.array_equal(o, scn["ref_occupancy"][t, :B]) and np.array_equal(s, scn["ref_semantic"][t, :B]), t
    assert emu.status()[0] == 0
    assert sizes == scn["ref_world_sizes"].tolist()
    b, xyz, sem = emu.world()
    assert np.array_equal(b, scn["ref_world_b"]) and np.array_equal(xyz.view(np.uint32), scn["ref_world_xyz"].view(np.uint32))
    assert np.array_equal(sem, scn["ref_world_sem"])


@pytest.mark.parametrize("seed", range(60))
def test_emulator_direct_edge_resolution_fuzz(seed):
    cfg, scn, order, tile = _fuzz_case(20_000 + seed)
    orc = OracleMapper(cfg.height, cfg.width, cfg.vfov_radians, cfg.map_meters, cfg.map_meters, cfg.resolution)
    emu = EmuMapper(cfg.height, cfg.width, cfg.vfov_radians, cfg.map_meters, cfg.resolution, max_envs=cfg.num_envs,
                    store=2048 if cfg.resolution < 0.1 else 1024, order=order, tile=tile, direct=True)
    merged = 0
    for t in range(cfg.steps):
        a = (scn["masks"][t], scn["pose"][t], scn["orientation"][t])
        kw = dict(depth=scn["depth"][t], labels=scn["labels"][t])
        o1, s1 = orc.step(*a, **kw)
        o2, s2 = emu.step(*a, **kw)
        assert emu.status()[0] == 0
        assert np.array_equal(o1, o2) and np.array_equal(s1, s2), f"maps differ at step {t}"
        b1, x1, m1 = orc.world()
        b2, x2, m2 = emu.world()
        assert np.array_equal(b1, b2) and np.array_equal(x1.view(np.uint32), x2.view(np.uint32)) and np.array_equal(m1, m2)
        _, stats = emu.status()
        assert int(stats[0]) == orc.counters["n_valid"] and int(stats[1]) == orc.counters["n_local"]
        assert int(stats[2]) == orc.counters["n_world"] and int(stats[3]) == orc.counters["n_in"]
