"""Full-size parity scenarios (BASELINE shapes), rebuilt from seeds: shared by tests/golden/make_golden_full.py (which
runs the unmodified reference on them in the build container) and by the tests that compare the oracle and the CUDA
path with the stored reference outputs (tests/golden/full/*.npz)."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

from ivlnce_b200.synthetic import ScenarioConfig, make_scenario
from scenarios import _wrap

FULL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full")

FULL_SCENARIOS = {
    # BASELINE config 1 / 4 shape: GT labels (27 classes), 16 envs whose coordinates overlap, a mid-run reset
    "full_gt16": dict(num_envs=16, steps=5, num_labels=27, depth_mode="iid", seed=5101, reset_steps={"3": [2, 9]}, roam=6.0),
    # BASELINE config 2 shape: 40 class-score planes per env (argmax labels), 16 envs
    "full_pred16": dict(num_envs=16, steps=3, num_labels=40, depth_mode="iid", seed=5102, reset_steps={}, roam=6.0, logits=40),
    # coherent depth: 8 envs walking through ONE box room -- walls on the world bounding box, exact height ties,
    # thousands of key collisions per step at full size
    "full_scene8": dict(num_envs=8, steps=5, num_labels=13, depth_mode="scene", seed=5103, reset_steps={"3": [1]}, roam=None),
}


def build_full(name):
    spec = FULL_SCENARIOS[name]
    c = ScenarioConfig(name=name, num_envs=spec["num_envs"], height=256, width=256, steps=spec["steps"], resolution=0.05,
                       num_labels=spec["num_labels"], depth_mode=spec["depth_mode"], env_spacing=0.0, seed=spec["seed"],
                       reset_steps={int(k): v for k, v in spec["reset_steps"].items()}, roam_radius=spec["roam"])
    scn = _wrap(c, make_scenario(c))
    if spec.get("logits"):
        rng = np.random.default_rng(spec["seed"] + 1)
        scn["logits"] = rng.standard_normal(size=(c.steps, c.num_envs, spec["logits"], 256, 256), dtype=np.float32)
    return scn


def world_digest(b, xyz, sem) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(b, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(xyz, dtype=np.float32).tobytes())
    h.update(np.ascontiguousarray(sem, dtype=np.uint8).tobytes())
    return h.hexdigest()


def full_names():
    return [n for n in FULL_SCENARIOS if os.path.exists(os.path.join(FULL_DIR, f"{n}.npz"))]


def load_full(name):
    z = np.load(os.path.join(FULL_DIR, f"{name}.npz"), allow_pickle=False)
    assert json.loads(str(z["spec"])) == FULL_SCENARIOS[name], "fixture was generated from a different scenario spec"
    R = z["ref_semantic"].shape[-1]
    return dict(ref_occupancy=np.unpackbits(z["ref_occupancy"], axis=-1)[..., :R], ref_semantic=z["ref_semantic"],
                ref_world_sizes=z["ref_world_sizes"], ref_world_sha256=str(z["ref_world_sha256"]),
                ref_world_head_xyz=z["ref_world_head_xyz"], ref_world_head_b=z["ref_world_head_b"],
                ref_world_head_sem=z["ref_world_head_sem"])
