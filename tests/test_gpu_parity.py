"""GPU parity tests: the CUDA path (through the public MappingModule -> C ABI) against the
golden fixtures of the unmodified reference and against the CPU oracle.  Integer state
(cell indices, labels, occupancy) must be bit-exact; float map values (the world records)
are compared bitwise as well, which is stricter than the 1e-5 the north star asks for."""
import numpy as np
import pytest
import torch

from golden_io import golden_names, load_golden
from scenarios import run_mapper

pytestmark = pytest.mark.gpu


def _run_cuda(scn, **kw):
    from cuda_stepper import CudaStepper

    pred = "logits" in scn
    cs = CudaStepper(scn["cfg"], known_clouds=scn.get("known"), pred=pred, max_envs=int(scn["num_envs"].max()), **kw)
    T = scn["masks"].shape[0]
    outs, sizes = [], []
    for t in range(T):
        B = int(scn["num_envs"][t])
        k = {}
        if scn["cfg"]["mode"] == "iterative":
            k["depth"] = scn["depth"][t, :B]
            if pred:
                k["logits"] = scn["logits"][t, :B]
            else:
                k["labels"] = scn["labels"][t, :B]
        else:
            k["env_names"] = scn["env_names"][t][:B]
        o, s = cs.step(scn["masks"][t, :B], scn["pose"][t, :B], scn["orientation"][t, :B], **k)
        outs.append((o, s))
        if pred:
            lab = cs.last_obs.semantics.cpu().numpy().reshape(B, *scn["labels_for_map"].shape[2:])
            assert np.array_equal(lab, scn["labels_for_map"][t, :B]), f"argmax labels differ at step {t}"
        if scn["cfg"]["mode"] == "iterative":
            sizes.append(len(cs.world()[0]))
    return cs, outs, sizes


@pytest.mark.parametrize("name", golden_names())
def test_cuda_matches_reference_golden(name):
    scn = load_golden(name)
    cs, outs, sizes = _run_cuda(scn)
    for t, (o, s) in enumerate(outs):
        B = o.shape[0]
        assert np.array_equal(o, scn["ref_occupancy"][t, :B]), f"occupancy differs at step {t}"
        assert np.array_equal(s, scn["ref_semantic"][t, :B]), f"semantic differs at step {t}"
    cs.mm.check_errors()
    if scn["cfg"]["mode"] == "iterative":
        assert sizes == scn["ref_world_sizes"].tolist()
    b, xyz, sem = cs.world()
    assert np.array_equal(b, scn["ref_world_b"])
    assert np.array_equal(xyz.view(np.uint32), scn["ref_world_xyz"].view(np.uint32))
    assert np.array_equal(sem, scn["ref_world_sem"])


@pytest.mark.parametrize("tile", [4, 8, 16, 64])
@pytest.mark.parametrize("name", ["iid_f32_res005", "degenerate", "identical_envs", "scene_overlap"])
def test_raster_tile_sizes(name, tile):
    """Ego tile sizes (0 = the library chooses).  `degenerate` deletes records in stage 2 of the edge fix-up while
    other CTAs raster beside it: with many small tiles this caught a race on the env boxes under reconstruction."""
    scn = load_golden(name)
    cs, outs, _ = _run_cuda(scn, raster_tile=tile)
    for t, (o, s) in enumerate(outs):
        B = o.shape[0]
        assert np.array_equal(o, scn["ref_occupancy"][t, :B]) and np.array_equal(s, scn["ref_semantic"][t, :B]), t
    cs.mm.check_errors()


@pytest.mark.parametrize("name", [n for n in golden_names() if n != "known_map"])
def test_cuda_four_kernel_path_matches_reference_golden(name):
    """The multi-kernel step (taken when the image does not tile evenly, or on request) against the goldens."""
    scn = load_golden(name)
    cs, outs, sizes = _run_cuda(scn, scatter_variant=2)
    for t, (o, s) in enumerate(outs):
        B = o.shape[0]
        assert np.array_equal(o, scn["ref_occupancy"][t, :B]), f"occupancy differs at step {t}"
        assert np.array_equal(s, scn["ref_semantic"][t, :B]), f"semantic differs at step {t}"
    cs.mm.check_errors()
    assert sizes == scn["ref_world_sizes"].tolist()
    assert all(v == 0 for v in cs.mm.phase_ns())  # the fused kernel never ran


def test_fused_kernel_is_the_default_path():
    """64x64 frames tile evenly: the step must run as the single fused persistent kernel."""
    scn = load_golden("iid_f64")
    cs, outs, _ = _run_cuda(scn)
    ns = cs.mm.phase_ns()
    assert ns[0] > 0 and ns[0] <= ns[1] <= ns[2] <= ns[3] <= ns[4] <= ns[5], ns
    T = scn["masks"].shape[0]
    assert cs.mm.kernel_launches() <= T + 1 + 5 * T  # one step kernel per call (+ init, + the world exports of the test)


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_predicted_ingest_variants(variant):
    """Fused persistent kernel / four kernels with register-staged loads / four kernels with the bulk-async ring."""
    scn = load_golden("predicted")
    cs, outs, _ = _run_cuda(scn, scatter_variant=variant)
    for t, (o, s) in enumerate(outs):
        assert np.array_equal(o, scn["ref_occupancy"][t]) and np.array_equal(s, scn["ref_semantic"][t]), t


def test_predicted_full_size_against_oracle():
    """BASELINE config 2 shape on 4 envs: 40-class score planes, 256x256, ties and NaNs included."""
    from ivlnce_b200.synthetic import ScenarioConfig, make_scenario
    from oracle.oracle import OracleMapper, argmax_labels
    from scenarios import _wrap

    c = ScenarioConfig(num_envs=4, height=256, width=256, steps=4, resolution=0.05, num_labels=40, seed=78)
    scn = _wrap(c, make_scenario(c))
    rng = np.random.default_rng(5)
    lg = np.round(rng.standard_normal((c.steps, c.num_envs, 40, 256, 256)).astype(np.float32) * 4) / 4
    lg[rng.random(lg.shape) < 1e-4] = np.nan
    scn["logits"] = lg
    scn["labels_for_map"] = np.stack([argmax_labels(lg[t]) for t in range(c.steps)])
    orc = OracleMapper(c.height, c.width, c.vfov_radians, c.map_meters, c.map_meters, c.resolution)
    ref_outs, _ = run_mapper(orc.step, scn)
    for variant in (0, 1, 2):
        cs, outs, _ = _run_cuda(scn, scatter_variant=variant)
        for t in range(c.steps):
            assert np.array_equal(outs[t][0], ref_outs[t][0]) and np.array_equal(outs[t][1], ref_outs[t][1]), (variant, t)
        cs.mm.check_errors()


@pytest.mark.parametrize("variant", [0, 2])
def test_candidate_table_stamp_wrap(variant):
    """300 steps of a small scenario with the stamp period of the frame candidate plane shortened to 255 / 7 steps:
    the stamp wraps (and the plane is cleared by the host) in mid-run; results must stay bit-exact against the oracle."""
    from ivlnce_b200.synthetic import ScenarioConfig, make_scenario
    from oracle.oracle import OracleMapper
    from scenarios import _wrap

    c = ScenarioConfig(num_envs=2, height=32, width=32, steps=300, resolution=0.1, num_labels=13,
                       reset_steps={100: [0], 257: [1]}, seed=91, roam_radius=3.0)
    scn = _wrap(c, make_scenario(c))
    orc = OracleMapper(c.height, c.width, c.vfov_radians, c.map_meters, c.map_meters, c.resolution)
    ref_outs, _ = run_mapper(orc.step, scn)
    from cuda_stepper import CudaStepper

    for period in (255, 7):
        cs = CudaStepper(scn["cfg"], max_envs=2, scatter_variant=variant, stamp_period=period)
        for t in range(c.steps):
            o, s = cs.step(scn["masks"][t], scn["pose"][t], scn["orientation"][t], depth=scn["depth"][t], labels=scn["labels"][t])
            assert np.array_equal(o, ref_outs[t][0]) and np.array_equal(s, ref_outs[t][1]), (period, t)
        cs.mm.check_errors()


def test_device_trig_f64_matches():
    """sin/cos evaluated by torch on the GPU in float64 round to the same float32 matrices."""
    scn = load_golden("iid_f64")
    cs, outs, _ = _run_cuda(scn, host_trig=False)
    for t, (o, s) in enumerate(outs):
        assert np.array_equal(o, scn["ref_occupancy"][t]) and np.array_equal(s, scn["ref_semantic"][t])


@pytest.mark.parametrize("name", ["iid_f64", "scene_overlap", "single_long", "identical_envs", "degenerate", "known_map"])
def test_kernel_trig_f64_matches(name):
    """Default mode: pose matrices derived inside the prep kernel from the float64 angles."""
    scn = load_golden(name)
    assert scn["orientation"].dtype == np.float64
    cs, outs, _ = _run_cuda(scn, trig="kernel")
    for t, (o, s) in enumerate(outs):
        B = o.shape[0]
        assert np.array_equal(o, scn["ref_occupancy"][t, :B]) and np.array_equal(s, scn["ref_semantic"][t, :B]), t


def test_full_size_against_oracle():
    """BASELINE config 1 shape (256x256 depth, 27 labels, 0.05 m cells) on 3 envs, 8 steps."""
    from ivlnce_b200.synthetic import ScenarioConfig, make_scenario
    from oracle.oracle import OracleMapper
    from scenarios import _wrap

    c = ScenarioConfig(num_envs=3, height=256, width=256, steps=8, resolution=0.05, num_labels=27,
                       reset_steps={5: [2]}, seed=77)
    scn = _wrap(c, make_scenario(c))
    orc = OracleMapper(c.height, c.width, c.vfov_radians, c.map_meters, c.map_meters, c.resolution)
    ref_outs, _ = run_mapper(orc.step, scn)
    cs, outs, _ = _run_cuda(scn)
    for t in range(c.steps):
        assert np.array_equal(outs[t][0], ref_outs[t][0]), t
        assert np.array_equal(outs[t][1], ref_outs[t][1]), t
    b1, x1, s1 = orc.world()
    b2, x2, s2 = cs.world()
    assert np.array_equal(b1, b2) and np.array_equal(x1.view(np.uint32), x2.view(np.uint32)) and np.array_equal(s1, s2)
    cs.mm.check_errors()


def test_batch_grow_keeps_world():
    """Growing the batch beyond the context capacity re-creates the context and copies the stores."""
    scn = load_golden("batch_shrink_grow")
    from cuda_stepper import CudaStepper

    cs = CudaStepper(scn["cfg"], max_envs=1)  # forces two grows (1 -> 2 -> 4)
    T = scn["masks"].shape[0]
    for t in range(T):
        B = int(scn["num_envs"][t])
        o, s = cs.step(scn["masks"][t, :B], scn["pose"][t, :B], scn["orientation"][t, :B],
                       depth=scn["depth"][t, :B], labels=scn["labels"][t, :B])
        assert np.array_equal(o, scn["ref_occupancy"][t, :B]) and np.array_equal(s, scn["ref_semantic"][t, :B]), t


def test_store_overflow_is_reported():
    from ivlnce_b200._lib import MapLibraryError

    scn = load_golden("iid_f32_res005")
    cs, outs, _ = _run_cuda(scn, store_cells=256)  # 6.4 m window: far points fall outside
    with pytest.raises(MapLibraryError):
        cs.mm.check_errors()


def test_smoke_entry():
    import __graft_entry__ as g

    g.smoke()


@pytest.mark.parametrize("name", ["iid_f64", "scene_overlap", "identical_envs", "degenerate", "batch_shrink_grow", "predicted"])
def test_step_kernel_multi_chunk_path(name, monkeypatch):
    """Debug bit 128 makes the persistent step kernel hold only 2 tiles per CTA at a time, so that CTAs with more
    tiles take the multi-chunk path (queues rebuilt in the resolve phase, label slots recycled between the argmax
    and the resolve warps) -- the path a GPU takes with more than ~18 (scores) / ~37 (GT labels) 256x256 envs."""
    monkeypatch.setenv("IVM_DEBUG_FLAGS", "128")
    scn = load_golden(name)
    cs, outs, sizes = _run_cuda(scn)
    for t, (o, s) in enumerate(outs):
        B = o.shape[0]
        assert np.array_equal(o, scn["ref_occupancy"][t, :B]), f"occupancy differs at step {t}"
        assert np.array_equal(s, scn["ref_semantic"][t, :B]), f"semantic differs at step {t}"
    cs.mm.check_errors()
    assert sizes == scn["ref_world_sizes"].tolist()
    b, xyz, sem = cs.world()
    assert np.array_equal(b, scn["ref_world_b"])
    assert np.array_equal(xyz.view(np.uint32), scn["ref_world_xyz"].view(np.uint32))
    assert np.array_equal(sem, scn["ref_world_sem"])


def test_many_envs_against_oracle():
    """48 envs of 128x128 depth (GT labels) and 24 envs with class scores: more tiles than co-resident CTAs, every
    CTA holds several tiles of different envs; envs reset at different steps."""
    from ivlnce_b200.synthetic import ScenarioConfig, make_scenario
    from oracle.oracle import OracleMapper, argmax_labels
    from scenarios import _wrap

    for pred, nenv in ((False, 48), (True, 24)):
        c = ScenarioConfig(num_envs=nenv, height=128, width=128, steps=5, resolution=0.1, num_labels=13,
                           reset_steps={2: [1, 7], 3: [0]}, seed=91 + nenv)
        scn = _wrap(c, make_scenario(c))
        if pred:
            rng = np.random.default_rng(6)
            lg = np.round(rng.standard_normal((c.steps, c.num_envs, 13, 128, 128)).astype(np.float32) * 4) / 4
            scn["logits"] = lg
            scn["labels_for_map"] = np.stack([argmax_labels(lg[t]) for t in range(c.steps)])
            scn["labels"] = scn["labels_for_map"]
        orc = OracleMapper(c.height, c.width, c.vfov_radians, c.map_meters, c.map_meters, c.resolution)
        ref_outs, _ = run_mapper(orc.step, scn)
        b1, x1, s1 = orc.world()
        # "host": pose matrices from the host's libm; "kernel": derived in the kernel's prep phase (sincos in f64) --
        # the world records (their x / z bits come straight from the matrix entries) must agree bitwise either way
        for trig in ("host", "kernel"):
            cs, outs, _ = _run_cuda(scn, store_cells=1024, trig=trig)
            for t in range(c.steps):
                assert np.array_equal(outs[t][0], ref_outs[t][0]), (pred, trig, t)
                assert np.array_equal(outs[t][1], ref_outs[t][1]), (pred, trig, t)
            b2, x2, s2 = cs.world()
            assert np.array_equal(b1, b2) and np.array_equal(x1.view(np.uint32), x2.view(np.uint32)) and np.array_equal(s1, s2), trig
            cs.mm.check_errors()


def test_tour_accumulation_against_oracle():
    """BASELINE config 3 shape, shortened: ONE env walks a coherent scene for 3 episodes x 40 steps with a single
    reset at the tour start (iterative map), 128x128 depth, 0.05 m cells, into a 2048 x 2048 half-cell store."""
    from ivlnce_b200.synthetic import ScenarioConfig, make_scenario
    from oracle.oracle import OracleMapper
    from scenarios import _wrap

    c = ScenarioConfig(num_envs=1, height=128, width=128, steps=120, resolution=0.05, num_labels=27, depth_mode="scene",
                       roam_radius=6.0, seed=301)
    scn = _wrap(c, make_scenario(c))
    assert int((scn["masks"] == 0).sum()) == 1  # the tour is reset once, at its start
    orc = OracleMapper(c.height, c.width, c.vfov_radians, c.map_meters, c.map_meters, c.resolution)
    ref_outs, ref_sizes = run_mapper(orc.step, scn, world_fn=orc.world)
    cs, outs, sizes = _run_cuda(scn, store_cells=2048)
    for t in range(c.steps):
        assert np.array_equal(outs[t][0], ref_outs[t][0]), t
        assert np.array_equal(outs[t][1], ref_outs[t][1]), t
    assert sizes == ref_sizes and sizes[-1] > sizes[10]  # the world cloud keeps growing over the tour
    b1, x1, s1 = orc.world()
    b2, x2, s2 = cs.world()
    assert np.array_equal(b1, b2) and np.array_equal(x1.view(np.uint32), x2.view(np.uint32)) and np.array_equal(s1, s2)
    cs.mm.check_errors()


def test_known_map_64_envs_against_oracle():
    """BASELINE config 5 shape: known-map registration + rotated ego crop at 64 envs, 0.05 m cells (1024 x 1024
    half-cell stores), 60 k-point scene clouds, scenes swapped mid-run."""
    from ivlnce_b200.synthetic import ScenarioConfig, make_known_cloud, make_scenario
    from oracle.oracle import OracleMapper
    from scenarios import _wrap

    T, B = 4, 64
    c = ScenarioConfig(name="known64", num_envs=B, height=8, width=8, steps=T, resolution=0.05, env_spacing=0.0, seed=411)
    s = make_scenario(c)
    known = {f"scene{i}": make_known_cloud(60000, 16.0, 27, seed=5000 + i) for i in range(6)}
    names = [[f"scene{(b + (t >= 2)) % 6}" for b in range(B)] for t in range(T)]
    s["masks"][2, :] = 0  # every env moves to another scene at t = 2
    scn = _wrap(c, s, mode="known")
    scn["env_names"] = names
    scn["known"] = known
    del scn["depth"], scn["labels"]
    orc = OracleMapper(c.height, c.width, c.vfov_radians, c.map_meters, c.map_meters, c.resolution, mode="known",
                       known_clouds=known)
    ref_outs, _ = run_mapper(orc.step, scn)
    cs, outs, _ = _run_cuda(scn, store_cells=1024)
    for t in range(T):
        assert np.array_equal(outs[t][0], ref_outs[t][0]), t
        assert np.array_equal(outs[t][1], ref_outs[t][1]), t
        assert outs[t][0].any()
    cs.mm.check_errors()


@pytest.mark.parametrize("variant", [0, 2])
def test_step_statistics_match_oracle(variant):
    """The counters the roofline is computed from -- valid pixels, frame de-dup survivors, live world records and
    world records rastered into the ego windows -- equal the oracle's after every step (persistent kernel, where the
    raster runs beside the fix-up, and four-kernel path)."""
    from ivlnce_b200.synthetic import ScenarioConfig, make_scenario
    from oracle.oracle import OracleMapper
    from scenarios import _wrap
    from cuda_stepper import CudaStepper

    c = ScenarioConfig(num_envs=6, height=128, width=128, steps=6, resolution=0.05, num_labels=27, depth_mode="scene",
                       reset_steps={3: [2]}, seed=512)
    scn = _wrap(c, make_scenario(c))
    orc = OracleMapper(c.height, c.width, c.vfov_radians, c.map_meters, c.map_meters, c.resolution)
    cs = CudaStepper(scn["cfg"], max_envs=c.num_envs, scatter_variant=variant)
    for t in range(c.steps):
        a = (scn["masks"][t], scn["pose"][t], scn["orientation"][t])
        kw = dict(depth=scn["depth"][t], labels=scn["labels"][t])
        o1, s1 = orc.step(*a, **kw)
        o2, s2 = cs.step(*a, **kw)
        assert np.array_equal(o1, o2) and np.array_equal(s1, s2), t
        flags, stats = cs.mm.status()
        assert flags == 0
        want = [orc.counters[k] for k in ("n_valid", "n_local", "n_world", "n_in")]
        assert [int(v) for v in stats[:4]] == want, (t, stats[:4], want)


def _run_back_to_back(scn, pipelined, repeat_from=None, **kw):
    """All inputs resident on the device first, then every step enqueued back to back with nothing in between
    (consecutive step kernels are programmatically serialised and -- pipelined -- overlap); only the last step's
    maps and the final world cloud are read back."""
    from cuda_stepper import CudaStepper
    from ivlnce_b200.mapper import EpisodesInfo, Observations, RobotCurrentState

    pred = "logits" in scn
    cs = CudaStepper(scn["cfg"], pred=pred, max_envs=int(scn["num_envs"].max()), trig="kernel", **kw)
    cs.mm.set_pipelined(pipelined)
    dev = cs.dev
    T = scn["masks"].shape[0]
    B = int(scn["num_envs"][0])
    assert all(int(b) == B for b in scn["num_envs"])
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    masks, pose, ori = up(scn["masks"]), up(scn["pose"]), up(scn["orientation"])
    depth = up(scn["depth"]).unsqueeze(2)                       # [T,B,1,H,W]
    sem = up(scn["logits"]) if pred else up(scn["labels"]).unsqueeze(2)
    names = [f"scene{b}" for b in range(B)]
    torch.cuda.synchronize()
    out = None
    for t in range(T):
        ei = EpisodesInfo(masks[t].reshape(B, 1), names)
        obs = Observations(None, depth[t], sem[t]) if pred else Observations(sem[t], depth[t], None)
        out = cs.mm(ei, obs, RobotCurrentState(pose[t], ori[t, :, 0], ori[t, :, 1]))
    torch.cuda.synchronize()
    return cs, out.occupancy.cpu().numpy(), out.semantic.cpu().numpy()


@pytest.mark.parametrize("pipelined", [False, True])
@pytest.mark.parametrize("name", ["iid_f64", "scene_overlap", "single_long", "identical_envs", "degenerate", "thresholds",
                                  "predicted"])
def test_back_to_back_steps_match_reference_golden(name, pipelined):
    """Steps enqueued back to back overlap on the GPU (the next step's CTAs move in while the current step still
    rasters).  The last maps and the final world cloud -- which depends on every step -- must still be the
    reference's, bit for bit."""
    scn = load_golden(name)
    assert scn["orientation"].dtype == np.float64
    cs, occ, sem = _run_back_to_back(scn, pipelined)
    T = scn["masks"].shape[0]
    assert np.array_equal(occ, scn["ref_occupancy"][T - 1]) and np.array_equal(sem, scn["ref_semantic"][T - 1])
    cs.mm.check_errors()
    b, xyz, s = cs.world()
    assert np.array_equal(b, scn["ref_world_b"])
    assert np.array_equal(xyz.view(np.uint32), scn["ref_world_xyz"].view(np.uint32))
    assert np.array_equal(s, scn["ref_world_sem"])


@pytest.mark.parametrize("flags", [128, 128 + 64])
def test_back_to_back_multi_chunk(flags, monkeypatch):
    """Same with two tiles per chunk (the queues are rebuilt in the resolve), pipelined."""
    monkeypatch.setenv("IVM_DEBUG_FLAGS", str(flags))
    scn = load_golden("iid_f32_res005") if False else load_golden("single_long")
    cs, occ, sem = _run_back_to_back(scn, True)
    T = scn["masks"].shape[0]
    assert np.array_equal(occ, scn["ref_occupancy"][T - 1]) and np.array_equal(sem, scn["ref_semantic"][T - 1])
    b, xyz, s = cs.world()
    assert np.array_equal(xyz.view(np.uint32), scn["ref_world_xyz"].view(np.uint32)) and np.array_equal(s, scn["ref_world_sem"])


@pytest.mark.parametrize("pred", [False, True])
def test_back_to_back_full_size_against_oracle(pred):
    """BASELINE shapes (256x256 depth, 0.05 m cells; 12 envs with 27 GT labels / 6 envs with 40 score planes), 10 steps
    enqueued back to back in pipelined mode with a mid-run reset: every step kernel runs with its full grid and
    overlaps its predecessor; final maps + world cloud against the oracle."""
    from ivlnce_b200.synthetic import ScenarioConfig, make_scenario
    from oracle.oracle import OracleMapper, argmax_labels
    from scenarios import _wrap

    B = 6 if pred else 12
    c = ScenarioConfig(num_envs=B, height=256, width=256, steps=10, resolution=0.05, num_labels=40 if pred else 27,
                       reset_steps={6: [1, B - 1]}, seed=177, env_spacing=0.0, roam_radius=4.0)
    scn = _wrap(c, make_scenario(c))
    if pred:
        rng = np.random.default_rng(6)
        scn["logits"] = np.round(rng.standard_normal((c.steps, B, 40, 256, 256)).astype(np.float32) * 4) / 4
        scn["labels_for_map"] = np.stack([argmax_labels(scn["logits"][t]) for t in range(c.steps)])
    orc = OracleMapper(c.height, c.width, c.vfov_radians, c.map_meters, c.map_meters, c.resolution)
    ref_outs, _ = run_mapper(orc.step, scn)
    cs, occ, sem = _run_back_to_back(scn, True)
    assert np.array_equal(occ, ref_outs[-1][0]) and np.array_equal(sem, ref_outs[-1][1])
    cs.mm.check_errors()
    b1, x1, s1 = orc.world()
    b2, x2, s2 = cs.world()
    assert np.array_equal(b1, b2) and np.array_equal(x1.view(np.uint32), x2.view(np.uint32)) and np.array_equal(s1, s2)


def test_store_overflow_reaches_a_caller_that_never_asks():
    """A trainer calls forward() and nothing else: a world-store overflow (points dropped, unlike the reference's
    unbounded cloud) must surface by itself -- raised by a later forward(), without any synchronising status call --
    or, with on_overflow="warn", warned once while the run continues."""
    import warnings

    from cuda_stepper import CudaStepper
    from ivlnce_b200._lib import MapLibraryError

    scn = load_golden("iid_f32_res005")
    T = scn["masks"].shape[0]

    def run(cs, steps):
        for i in range(steps):
            t = i % T
            cs.step(scn["masks"][t], scn["pose"][t], scn["orientation"][t], depth=scn["depth"][t], labels=scn["labels"][t])

    cs = CudaStepper(scn["cfg"], store_cells=256, max_envs=2)          # 6.4 m window: far points fall outside
    cs.mm.error_poll_interval = 2
    with pytest.raises(MapLibraryError, match="store window"):
        run(cs, 12)
    cs = CudaStepper(scn["cfg"], store_cells=256, max_envs=2)
    cs.mm.error_poll_interval, cs.mm.on_overflow = 2, "warn"
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        run(cs, 12)
    assert len(w) == 1 and "DROPPED" in str(w[0].message)
    cs = CudaStepper(scn["cfg"], max_envs=2)                           # the default window holds everything: silent
    cs.mm.error_poll_interval = 2
    run(cs, 12)


def test_step_counter_rebase_keeps_parity():
    """the 24-bit step stamps are rebased after 2^24 - 1 steps (ivm_rebase_stamps).  Start a context just below
    the limit so that the rebase happens in mid-run; maps and world cloud must stay the reference's."""
    import ctypes

    from cuda_stepper import CudaStepper
    from ivlnce_b200 import _lib

    scn = load_golden("iid_f64")
    cs = CudaStepper(scn["cfg"], max_envs=3, trig="kernel")
    eng = cs.mm.engine(3)
    _lib.check(eng.lib.ivm_debug_set_step(eng.ctx, ctypes.c_uint32(0xFFFFFF - 5)))
    T = scn["masks"].shape[0]
    for t in range(T):
        o, s = cs.step(scn["masks"][t], scn["pose"][t], scn["orientation"][t], depth=scn["depth"][t], labels=scn["labels"][t])
        assert np.array_equal(o, scn["ref_occupancy"][t]) and np.array_equal(s, scn["ref_semantic"][t]), t
    cs.mm.check_errors()
    b, xyz, sem = cs.world()
    assert np.array_equal(b, scn["ref_world_b"]) and np.array_equal(xyz.view(np.uint32), scn["ref_world_xyz"].view(np.uint32))
    assert np.array_equal(sem, scn["ref_world_sem"])


@pytest.mark.parametrize("nenv,mode", [(1, "iid"), (3, "iid"), (2, "scene")])
def test_edge_scan_skipping_keeps_parity(nenv, mode):
    """The step kernel scans the edge lines of the world bounding box (where the reference's key quirk merges distinct
    cells, mapper.py:461-474) only when a collision can have appeared since the last scan.  Long runs in a small
    area: the box settles, later frames keep landing on its edge lines, envs are reset in mid-run.  Maps and the
    size of the world cloud after EVERY step, and the final cloud bitwise, against the oracle; and both kinds of
    steps (scanned / skipped) must have occurred."""
    from ivlnce_b200.synthetic import ScenarioConfig, make_scenario
    from oracle.oracle import OracleMapper
    from scenarios import _wrap

    c = ScenarioConfig(num_envs=nenv, height=64, width=64, steps=90, resolution=0.1, num_labels=13, depth_mode=mode,
                       roam_radius=1.0, reset_steps={40: [0], 63: [nenv - 1]}, seed=700 + nenv)
    scn = _wrap(c, make_scenario(c))
    orc = OracleMapper(c.height, c.width, c.vfov_radians, c.map_meters, c.map_meters, c.resolution)
    ref_outs, ref_sizes = run_mapper(orc.step, scn, world_fn=orc.world)
    from cuda_stepper import CudaStepper

    cs = CudaStepper(scn["cfg"], max_envs=nenv, store_cells=1024)
    scanned = skipped = 0
    for t in range(c.steps):
        o, s = cs.step(scn["masks"][t], scn["pose"][t], scn["orientation"][t], depth=scn["depth"][t], labels=scn["labels"][t])
        assert np.array_equal(o, ref_outs[t][0]) and np.array_equal(s, ref_outs[t][1]), t
        assert len(cs.world()[0]) == ref_sizes[t], t
        _, stats = cs.mm.status()
        if int(stats[7]) >> 32:
            scanned += 1
        else:
            skipped += 1
    b1, x1, s1 = orc.world()
    b2, x2, s2 = cs.world()
    assert np.array_equal(b1, b2) and np.array_equal(x1.view(np.uint32), x2.view(np.uint32)) and np.array_equal(s1, s2)
    assert scanned > 0 and skipped > 0, (scanned, skipped)
    cs.mm.check_errors()


def test_two_contexts_with_different_tiles_alternate():
    """the dynamic shared memory attribute of the step kernel belongs to the function, not to a context.
    Two live mappers whose raster tiles (hence shared memory sizes) differ step alternately."""
    scn = load_golden("scene_overlap")
    from cuda_stepper import CudaStepper

    a = CudaStepper(scn["cfg"], max_envs=int(scn["num_envs"].max()), raster_tile=64)
    b = CudaStepper(scn["cfg"], max_envs=int(scn["num_envs"].max()), raster_tile=4)
    T = scn["masks"].shape[0]
    for t in range(T):
        B = int(scn["num_envs"][t])
        args = (scn["masks"][t, :B], scn["pose"][t, :B], scn["orientation"][t, :B])
        kw = dict(depth=scn["depth"][t, :B], labels=scn["labels"][t, :B])
        oa, sa = a.step(*args, **kw)
        ob, sb = b.step(*args, **kw)
        assert np.array_equal(oa, ob) and np.array_equal(sa, sb), t
        assert np.array_equal(oa, scn["ref_occupancy"][t, :B]) and np.array_equal(sa, scn["ref_semantic"][t, :B]), t
    a.mm.check_errors(); b.mm.check_errors()


def test_known_cloud_cache_is_bounded_and_capacity_error_is_clear(tmp_path):
    """scene clouds are kept on the device for a bounded number of scenes (re-read from disk beyond that), and a
    cloud larger than known_capacity raises a message that says so."""
    from ivlnce_b200 import _lib
    from ivlnce_b200.mapper import EpisodesInfo, MapDimensions, Observations, RobotCurrentState, create_known_mapper

    rng = np.random.default_rng(5)
    dev = torch.device("cuda:0")
    names = [f"scene{i}" for i in range(4)]
    for i, n in enumerate(names):
        k = 2500 + 400 * i
        xyz = (rng.random((k, 3), dtype=np.float32) * np.array([8, 1.5, 8], np.float32)).astype(np.float32)
        np.savez(tmp_path / f"{n}.npz", xyz=xyz, semantics=rng.integers(0, 13, k))
    np.savez(tmp_path / "huge.npz", xyz=rng.random((5000, 3), dtype=np.float32), semantics=np.zeros(5000, np.int64))
    md = MapDimensions(6.4, 6.4, 0.1)
    mm = create_known_mapper(dev, md, str(tmp_path), known_capacity=4096, store_cells=1024)
    mm.KNOWN_CACHE_SCENES = 2
    pose = torch.tensor([[4.0, 0.5, 4.0]], device=dev)
    ori = torch.zeros(1, 2, dtype=torch.float64, device=dev)
    outs = {}
    for rnd in range(2):                                   # second round: every scene but the last two is re-read from disk
        for n in names:
            ei = EpisodesInfo(torch.zeros(1, 1, dtype=torch.uint8, device=dev), [n])
            o = mm(ei, Observations(None, None, None), RobotCurrentState(pose, ori[:, 0], ori[:, 1]))
            occ = o.occupancy.cpu().numpy().copy()
            assert len(mm._known_cache) <= 2
            if rnd == 0:
                outs[n] = occ
                assert occ.any()
            else:
                assert np.array_equal(occ, outs[n]), n
    with pytest.raises(_lib.MapLibraryError, match="known_capacity"):
        ei = EpisodesInfo(torch.zeros(1, 1, dtype=torch.uint8, device=dev), ["huge"])
        mm(ei, Observations(None, None, None), RobotCurrentState(pose, ori[:, 0], ori[:, 1]))


def test_long_tour_against_oracle():
    """BASELINE config 3 at length: ONE env, one reset at the start, 800 steps of a walk within 5 m (128x128 depth,
    0.05 m cells, 2048^2 half-cell store, ~0.6 M world records at the end).  Maps against the oracle every 25th step (and
    the last 10), the size of the world cloud at those steps, the final cloud bitwise.  Exercises what only long runs
    reach: a settled world box with the edge-line scan skipped for hundreds of steps, re-scans when the box moves, a
    store that keeps filling."""
    from ivlnce_b200.synthetic import ScenarioConfig, make_scenario
    from oracle.oracle import OracleMapper
    from scenarios import _wrap
    from cuda_stepper import CudaStepper

    c = ScenarioConfig(num_envs=1, height=128, width=128, steps=800, resolution=0.05, num_labels=27, depth_mode="iid",
                       roam_radius=5.0, seed=311)
    scn = _wrap(c, make_scenario(c))
    assert int((scn["masks"] == 0).sum()) == 1
    orc = OracleMapper(c.height, c.width, c.vfov_radians, c.map_meters, c.map_meters, c.resolution)
    cs = CudaStepper(scn["cfg"], max_envs=1, store_cells=2048)
    checked = 0
    for t in range(c.steps):
        args = (scn["masks"][t], scn["pose"][t], scn["orientation"][t])
        kw = dict(depth=scn["depth"][t], labels=scn["labels"][t])
        o_ref, s_ref = orc.step(*args, **kw)
        o, s = cs.step(*args, **kw)
        if t % 25 == 0 or t >= c.steps - 10:
            assert np.array_equal(o, o_ref) and np.array_equal(s, s_ref), t
            assert len(cs.world()[0]) == len(orc.world()[0]), t
            checked += 1
    b1, x1, s1 = orc.world()
    b2, x2, s2 = cs.world()
    assert np.array_equal(b1, b2) and np.array_equal(x1.view(np.uint32), x2.view(np.uint32)) and np.array_equal(s1, s2)
    assert checked >= 40 and len(b1) > 50_000
    cs.mm.check_errors()


@pytest.mark.parametrize("name", ["scene_overlap", "identical_envs", "iid_f64"])
def test_fused_and_four_kernel_steps_alternate_in_one_context(name):
    """A depth tensor that is not 16-byte aligned sends a step down the four-kernel path; the next aligned one goes back
    to the persistent kernel.  Both work on the same world state (store, env boxes, edge-scan record, per-step
    counters), so a run that alternates between them must still match the reference step for step."""
    from cuda_stepper import CudaStepper
    from ivlnce_b200.mapper import EpisodesInfo, Observations, RobotCurrentState

    scn = load_golden(name)
    cs = CudaStepper(scn["cfg"], max_envs=int(scn["num_envs"].max()))
    dev, mm = cs.dev, cs.mm
    T = scn["masks"].shape[0]
    launches = []
    for t in range(T):
        B = int(scn["num_envs"][t])
        d = np.ascontiguousarray(scn["depth"][t, :B])
        if t % 2 == 1:                       # element offset 1: the same values at a 4-byte-aligned address
            buf = torch.empty(d.size + 1, dtype=torch.float32, device=dev)
            buf[1:] = torch.from_numpy(d.reshape(-1)).to(dev)
            dt = buf[1:].view(B, 1, *d.shape[1:])
            assert dt.data_ptr() % 16 != 0
        else:
            dt = torch.from_numpy(d).to(dev).unsqueeze(1)
        lab = torch.from_numpy(np.ascontiguousarray(scn["labels"][t, :B])).to(dev).unsqueeze(1)
        ori = torch.from_numpy(np.ascontiguousarray(scn["orientation"][t, :B])).to(dev)
        ei = EpisodesInfo(torch.from_numpy(np.ascontiguousarray(scn["masks"][t, :B])).reshape(B, 1).to(dev), [f"scene{b}" for b in range(B)])
        n0 = mm.kernel_launches()
        out = mm(ei, Observations(lab, dt, None), RobotCurrentState(torch.from_numpy(np.ascontiguousarray(scn["pose"][t, :B])).to(dev), ori[:, 0], ori[:, 1]))
        launches.append(mm.kernel_launches() - n0)
        assert np.array_equal(out.occupancy.cpu().numpy(), scn["ref_occupancy"][t, :B]), t
        assert np.array_equal(out.semantic.cpu().numpy(), scn["ref_semantic"][t, :B]), t
        assert len(cs.world()[0]) == int(scn["ref_world_sizes"][t]), t
    H, W = scn["depth"].shape[2], scn["depth"].shape[3]
    assert min(launches[1::2]) >= 4, launches                                # the misaligned steps took the four-kernel path
    if (H * W) % 512 == 0 and W % 4 == 0:
        assert min(launches[2::2]) == 1, launches                            # ... and the others the persistent kernel
    b, xyz, sem = cs.world()
    assert np.array_equal(b, scn["ref_world_b"]) and np.array_equal(xyz.view(np.uint32), scn["ref_world_xyz"].view(np.uint32))
    assert np.array_equal(sem, scn["ref_world_sem"])
    mm.check_errors()
