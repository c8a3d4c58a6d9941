"""Full-size reference goldens (tests/golden/full/*.npz: BASELINE shapes, outputs of the UNMODIFIED reference run by
tests/golden/make_golden_full.py; inputs rebuilt from seeds) against the C oracle (CPU) and the CUDA path (GPU)."""
import numpy as np
import pytest

from full_scenarios import build_full, full_names, load_full, world_digest
from oracle.oracle import OracleMapper
from scenarios import run_mapper

_CACHE = {}


def _scn(name):
    if name not in _CACHE:
        _CACHE.clear()   # one full-size scenario in memory at a time
        scn = build_full(name)
        if "logits" in scn:
            from oracle.oracle import argmax_labels

            scn["labels_for_map"] = np.stack([argmax_labels(scn["logits"][t]) for t in range(scn["logits"].shape[0])])
        _CACHE[name] = scn
    return _CACHE[name]


def test_full_fixtures_exist():
    assert set(full_names()) >= {"full_gt16", "full_pred16", "full_scene8"}


@pytest.mark.parametrize("name", full_names())
def test_oracle_matches_full_size_reference(name):
    scn, ref = _scn(name), load_full(name)
    c = scn["cfg"]
    orc = OracleMapper(c["height"], c["width"], c["vfov"], c["map_m"], c["map_m"], c["resolution"])
    outs, sizes = run_mapper(orc.step, scn, world_fn=orc.world)
    for t, (o, s) in enumerate(outs):
        assert np.array_equal(o, ref["ref_occupancy"][t]), f"occupancy differs at step {t}"
        assert np.array_equal(s, ref["ref_semantic"][t]), f"semantic differs at step {t}"
    assert sizes == ref["ref_world_sizes"].tolist()
    assert world_digest(*orc.world()) == ref["ref_world_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", full_names())
def test_cuda_matches_full_size_reference(name):
    """The CUDA path at BASELINE size against the reference itself (not via the oracle): both maps of every step, the
    world-cloud size of every step and the final world cloud (values, labels, env ids, order) by digest."""
    from test_gpu_parity import _run_cuda

    scn, ref = _scn(name), load_full(name)
    cs, outs, sizes = _run_cuda(scn, trig="kernel")
    for t, (o, s) in enumerate(outs):
        assert np.array_equal(o, ref["ref_occupancy"][t]), f"occupancy differs at step {t}"
        assert np.array_equal(s, ref["ref_semantic"][t]), f"semantic differs at step {t}"
    cs.mm.check_errors()
    assert sizes == ref["ref_world_sizes"].tolist()
    b, xyz, sem = cs.world()
    assert np.array_equal(xyz[:256].view(np.uint32), ref["ref_world_head_xyz"].view(np.uint32))
    assert world_digest(b, xyz, sem) == ref["ref_world_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", full_names())
def test_cuda_pipelined_matches_full_size_reference(name):
    """Same inputs, steps enqueued back to back in pipelined mode (consecutive step kernels overlap): last maps and
    final world cloud."""
    from test_gpu_parity import _run_back_to_back

    scn, ref = _scn(name), load_full(name)
    cs, occ, sem = _run_back_to_back(scn, True)
    assert np.array_equal(occ, ref["ref_occupancy"][-1]) and np.array_equal(sem, ref["ref_semantic"][-1])
    cs.mm.check_errors()
    assert world_digest(*cs.world()) == ref["ref_world_sha256"]
