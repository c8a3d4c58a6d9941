"""the map-feature epilogue (one-hot + occupancy concat) against the reference expression."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.oracle import map_features_oracle


def _reference(occ: torch.Tensor, sem: torch.Tensor, K: int) -> torch.Tensor:
    # ivlnce_baselines/models/encoders/map_encoder.py:85-90, verbatim semantics
    occupancy = occ.unsqueeze(1)
    semantic = F.one_hot(sem.long(), K).permute(0, 3, 1, 2)
    return torch.cat((occupancy, semantic), 1).to(dtype=torch.float)


@pytest.mark.parametrize("shape,K", [((2, 64, 64), 13), ((3, 7, 5), 13), ((1, 128, 128), 27)])
def test_oracle_matches_reference_expression(shape, K):
    rng = np.random.default_rng(3)
    occ = rng.integers(0, 2, shape, dtype=np.uint8)
    sem = rng.integers(0, K, shape, dtype=np.uint8)
    want = _reference(torch.from_numpy(occ), torch.from_numpy(sem), K).numpy()
    assert np.array_equal(map_features_oracle(occ, sem, K), want)
    with pytest.raises(RuntimeError):
        map_features_oracle(occ, np.full(shape, K, np.uint8), K)


def test_cpu_tensors_are_refused():
    from ivlnce_b200._lib import MapLibraryError
    from ivlnce_b200.map_encoder import MapFeatures

    with pytest.raises(MapLibraryError):
        MapFeatures()({"occupancy_map": torch.zeros(1, 4, 4, dtype=torch.uint8), "semantic_map": torch.zeros(1, 4, 4, dtype=torch.uint8)})
    with pytest.raises(ValueError):
        MapFeatures()({"occupancy_map": torch.zeros(1, 4, 4)})


@pytest.mark.gpu
@pytest.mark.parametrize("shape,K", [((16, 128, 128), 13), ((3, 7, 5), 13), ((2, 64, 64), 27), ((5, 6, 6), 4)])
def test_cuda_matches_oracle(shape, K):
    from ivlnce_b200.map_encoder import MapFeatures

    rng = np.random.default_rng(11)
    occ = rng.integers(0, 2, shape, dtype=np.uint8)
    sem = rng.integers(0, K, shape, dtype=np.uint8)
    mf = MapFeatures(K)
    dev = torch.device("cuda:0")
    out = mf({"occupancy_map": torch.from_numpy(occ).to(dev), "semantic_map": torch.from_numpy(sem).to(dev)})
    mf.check_errors()
    assert out.dtype is torch.float32 and tuple(out.shape) == (shape[0], 1 + K, shape[1], shape[2])
    assert np.array_equal(out.cpu().numpy(), map_features_oracle(occ, sem, K))
    # out-of-range class values are reported the way F.one_hot reports them
    sem[0, 0, 0] = K
    mf({"occupancy_map": torch.from_numpy(occ).to(dev), "semantic_map": torch.from_numpy(sem).to(dev)})
    with pytest.raises(RuntimeError):
        mf.check_errors()


@pytest.mark.gpu
def test_features_of_a_mapping_step():
    """End of the path: MappingModule.forward -> maps -> features, against the oracle on the same maps."""
    from golden_io import load_golden
    from test_gpu_parity import _run_cuda
    from ivlnce_b200.map_encoder import MapFeatures

    scn = load_golden("scene_overlap")
    cs, outs, _ = _run_cuda(scn)
    mem = cs.mm.map_memory
    feats = MapFeatures(13)({"occupancy_map": mem.occupancy, "semantic_map": mem.semantic})
    want = map_features_oracle(outs[-1][0], outs[-1][1], 13)
    assert np.array_equal(feats.cpu().numpy(), want)
