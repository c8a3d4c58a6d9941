"""Parity scenarios shared by the golden generator, the oracle tests, the host
emulator tests and the GPU parity tests.  Each builder returns a dict:

  cfg      : dict(height, width, vfov, map_m, resolution, mode)
  depth    : f32 [T,Bmax,H,W]   (iterative)
  labels   : u8  [T,Bmax,H,W]   (iterative)
  logits   : f32 [T,Bmax,Cls,H,W] (optional: predicted-semantics front end)
  pose     : f32 [T,Bmax,3]
  orientation : f32|f64 [T,Bmax,2]  (elevation, heading)
  masks    : u8  [T,Bmax]
  num_envs : i64 [T]            (batch size of each call)
  env_names: list[T] of list[str]      (known mode)
  known    : dict name -> (xyz f32 [N,3], sem i64 [N])   (known mode)
"""
from __future__ import annotations

import math
from typing import Callable, Dict

import numpy as np

from ivlnce_b200.synthetic import ScenarioConfig, make_known_cloud, make_scenario


def _wrap(cfg: ScenarioConfig, scn: Dict[str, np.ndarray], mode="iterative"):
    T = scn["masks"].shape[0]
    out = dict(scn)
    out["cfg"] = dict(height=cfg.height, width=cfg.width, vfov=cfg.vfov_radians, map_m=cfg.map_meters,
                      resolution=cfg.resolution, mode=mode)
    out["num_envs"] = np.full(T, scn["masks"].shape[1], dtype=np.int64)
    return out


def iid_f64():
    c = ScenarioConfig(name="iid_f64", num_envs=3, height=64, width=64, steps=12, resolution=0.1,
                       num_labels=13, reset_steps={5: [1]}, seed=1001)
    return _wrap(c, make_scenario(c))


def iid_f32_res005():
    c = ScenarioConfig(name="iid_f32_res005", num_envs=2, height=64, width=64, steps=10, resolution=0.05,
                       num_labels=27, angle_dtype="float32", reset_steps={6: [0]}, seed=1002)
    return _wrap(c, make_scenario(c))


def scene_overlap():
    """Coherent scene, both envs in the same room (cross-env key collisions, height ties)."""
    c = ScenarioConfig(name="scene_overlap", num_envs=2, height=96, width=96, steps=12, resolution=0.05,
                       depth_mode="scene", env_spacing=0.0, reset_every=8, seed=1003)
    return _wrap(c, make_scenario(c))


def scene_f32():
    c = ScenarioConfig(name="scene_f32", num_envs=3, height=64, width=64, steps=12, resolution=0.1,
                       depth_mode="scene", env_spacing=0.0, angle_dtype="float32", seed=1004)
    return _wrap(c, make_scenario(c))


def single_long():
    """One env, no reset after t=0: within-env bbox-edge collisions every step."""
    c = ScenarioConfig(name="single_long", num_envs=1, height=64, width=64, steps=30, resolution=0.05,
                       seed=1005)
    return _wrap(c, make_scenario(c))


def identical_envs():
    """Four envs fed byte-identical frames and poses: every cross-env collision is an exact tie."""
    c = ScenarioConfig(name="identical_envs", num_envs=1, height=48, width=48, steps=10, resolution=0.1,
                       depth_mode="scene", env_spacing=0.0, seed=1006)
    s = make_scenario(c)
    s = {k: np.repeat(v, 4, axis=1) for k, v in s.items()}
    s["masks"][6, 2] = 0
    c.num_envs = 4
    return _wrap(c, s)


def thresholds():
    """Depth values exactly on / next to the strict 0.01 / 0.99 cut-offs (mapper.py:416-418)."""
    c = ScenarioConfig(name="thresholds", num_envs=2, height=32, width=32, steps=4, resolution=0.1, seed=1007)
    s = make_scenario(c)
    lo, hi = np.float32(0.01), np.float32(0.99)
    special = np.array([lo, np.nextafter(lo, np.float32(1)), np.nextafter(lo, np.float32(0)),
                        hi, np.nextafter(hi, np.float32(1)), np.nextafter(hi, np.float32(0)),
                        0.0, 1.0, 0.5], dtype=np.float32)
    rng = np.random.default_rng(5)
    pick = rng.integers(0, len(special), size=s["depth"].shape)
    use = rng.random(size=s["depth"].shape) < 0.5
    s["depth"] = np.where(use, special[pick], s["depth"]).astype(np.float32)
    return _wrap(c, s)


def degenerate():
    """Single-row / single-pixel / single-column / empty frames: degenerate bboxes of the
    de-dup key (rows.max()==0 or cols.max()==0 collapse whole rows / envs)."""
    H = W = 32
    T, B = 8, 2
    c = ScenarioConfig(name="degenerate", num_envs=B, height=H, width=W, steps=T, resolution=0.1, seed=1008)
    rng = np.random.default_rng(c.seed)
    depth = np.ones((T, B, H, W), dtype=np.float32)
    labels = rng.integers(0, 13, size=(T, B, H, W), dtype=np.uint8)
    pose = np.zeros((T, B, 3), dtype=np.float32)
    pose[:, :, 1] = 1.25
    pose[:, 1, 0] = 0.3
    ori = np.zeros((T, B, 2), dtype=np.float64)
    masks = np.ones((T, B), dtype=np.uint8)
    masks[0] = 0
    depth[0, :, 14, :] = 0.3            # one image row at constant depth: rows.max() == 0
    depth[1, 0, 15, 7] = 0.41           # a single pixel
    depth[2, :, 8:24, 16] = np.linspace(0.2, 0.5, 16, dtype=np.float32)[None, :]  # one image column
    masks[3, 1] = 0                     # empty frame + reset of env 1
    depth[4, :, 15, :] = 0.3            # same row again after the reset
    depth[4, 0, 15, :] = 0.31
    depth[5] = rng.uniform(0.05, 0.95, size=(B, H, W)).astype(np.float32)
    masks[6, :] = 0                     # both reset, frame is a single pixel per env
    depth[6, :, 16, 16] = 0.25
    depth[7, :, 16, 3:29] = 0.25
    scn = dict(depth=depth, labels=labels, pose=pose, orientation=ori, masks=masks)
    return _wrap(c, scn)


def batch_shrink_grow():
    """Batch 3 -> 2 -> 3 (paused env dropped, mapper.py:315-318, 533-553); env 2 comes
    back without a reset at t=7 and is reset at t=9."""
    c = ScenarioConfig(name="batch_shrink_grow", num_envs=3, height=48, width=48, steps=11, resolution=0.1,
                       env_spacing=3.0, seed=1009)
    s = make_scenario(c)
    out = _wrap(c, s)
    ne = np.full(c.steps, 3, dtype=np.int64)
    ne[4:7] = 2
    out["num_envs"] = ne
    out["masks"][9, 2] = 0
    return out


def predicted():
    """Predicted-semantics front end: 40-class scores -> argmax -> map update.  Scores are
    quantised so that ties occur, and sprinkled with NaNs."""
    c = ScenarioConfig(name="predicted", num_envs=2, height=32, width=32, steps=5, resolution=0.05,
                       num_labels=40, seed=1010)
    s = make_scenario(c)
    rng = np.random.default_rng(c.seed + 1)
    logits = rng.standard_normal(size=(c.steps, c.num_envs, 40, c.height, c.width)).astype(np.float32)
    logits = np.round(logits * 2.0) / 2.0
    nan = rng.random(size=logits.shape) < 0.002
    logits[nan] = np.nan
    s["logits"] = logits.astype(np.float32)
    return _wrap(c, s)


def known_map():
    """Known-map mode (mapper.py:851-881): scene clouds appended on reset, no de-dup."""
    T, B = 8, 2
    c = ScenarioConfig(name="known_map", num_envs=B, height=8, width=8, steps=T, resolution=0.1,
                       env_spacing=0.0, seed=1011)
    s = make_scenario(c)
    known = {}
    for i, name in enumerate(["sceneA", "sceneB", "sceneC"]):
        xyz, sem = make_known_cloud(6000, 14.0, 13, seed=2000 + i)
        known[name] = (xyz, sem)
    names = [["sceneA", "sceneB"] for _ in range(T)]
    for t in range(4, T):
        names[t] = ["sceneA", "sceneC"]
    s["masks"][4, 1] = 0
    out = _wrap(c, s, mode="known")
    out["env_names"] = names
    out["known"] = known
    del out["depth"], out["labels"]
    return out


SCENARIOS: Dict[str, Callable[[], dict]] = {
    f.__name__: f
    for f in (iid_f64, iid_f32_res005, scene_overlap, scene_f32, single_long, identical_envs, thresholds,
              degenerate, batch_shrink_grow, predicted, known_map)
}


def run_mapper(stepper, scn, world_fn=None):
    """Drive any stepper with signature step(masks, pose, orientation, depth=, labels=, env_names=)
    over a scenario; returns list of (occ, sem) per step (+ world sizes when world_fn is given)."""
    T = scn["masks"].shape[0]
    outs, sizes = [], []
    for t in range(T):
        B = int(scn["num_envs"][t])
        kw = {}
        if scn["cfg"]["mode"] == "iterative":
            kw["depth"] = scn["depth"][t, :B]
            kw["labels"] = scn["labels_for_map"][t, :B] if "labels_for_map" in scn else scn["labels"][t, :B]
        else:
            kw["env_names"] = scn["env_names"][t][:B]
        o, s = stepper(scn["masks"][t, :B], scn["pose"][t, :B], scn["orientation"][t, :B], **kw)
        outs.append((np.asarray(o).copy(), np.asarray(s).copy()))
        if world_fn is not None:
            sizes.append(len(world_fn()[0]))
    return outs, sizes
