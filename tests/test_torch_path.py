"""oracle/torch_path.py (the torch-op restatement used as the timed CPU baseline) must agree with
the golden fixtures of the unmodified reference (single-threaded: the reference's semantic write is
racy with more threads)."""
import numpy as np
import pytest
import torch

from golden_io import load_golden
from oracle.torch_path import TorchReferencePath


@pytest.mark.parametrize("name", ["iid_f64", "scene_overlap", "degenerate", "predicted", "known_map", "batch_shrink_grow"])
def test_torch_path_matches_golden(name):
    torch.set_num_threads(1)
    scn = load_golden(name)
    c = scn["cfg"]
    tp = TorchReferencePath(c["height"], c["width"], c["vfov"], c["map_m"], c["map_m"], c["resolution"],
                            known_clouds=scn.get("known"))
    for t in range(scn["masks"].shape[0]):
        B = int(scn["num_envs"][t])
        ori = torch.from_numpy(scn["orientation"][t, :B])
        kw = {}
        if c["mode"] == "iterative":
            kw["depth"] = torch.from_numpy(scn["depth"][t, :B]).unsqueeze(1)
            if "logits" in scn:
                kw["scores"] = torch.from_numpy(scn["logits"][t, :B])
            else:
                kw["labels"] = torch.from_numpy(scn["labels"][t, :B]).unsqueeze(1)
        else:
            kw["env_names"] = scn["env_names"][t][:B]
        occ, sem = tp.step(torch.from_numpy(scn["masks"][t, :B]), torch.from_numpy(scn["pose"][t, :B]),
                           ori[:, 0], ori[:, 1], **kw)
        assert np.array_equal(occ.numpy(), scn["ref_occupancy"][t, :B]), t
        assert np.array_equal(sem.numpy(), scn["ref_semantic"][t, :B]), t
