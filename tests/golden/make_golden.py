"""Generate tests/golden/*.npz by running the UNMODIFIED reference mapping module
(/root/reference, loaded through oracle/ref_loader.py) single-threaded on the
scenarios of tests/scenarios.py.  Build container only:

    python tests/golden/make_golden.py

Each fixture holds the scenario inputs and, per step, the reference's
occupancy / semantic maps, world-cloud size, and the full world cloud after
the last step (in the reference's own order).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ref_harness import ReferenceRunner  # noqa: E402
from scenarios import SCENARIOS, run_mapper  # noqa: E402


def main():
    torch.set_num_threads(1)
    for name, build in SCENARIOS.items():
        scn = build()
        cfg = scn["cfg"]
        if "logits" in scn:
            # PredictSemantics tail, mapper.py:795-798, with torch's own argmax
            lg = torch.from_numpy(scn["logits"])
            T, B = lg.shape[:2]
            lab = lg.reshape(T * B, *lg.shape[2:]).argmax(1, keepdims=True).to(torch.uint8)
            scn["labels_for_map"] = lab.reshape(T, B, *lg.shape[3:]).numpy()
        ref = ReferenceRunner(cfg["height"], cfg["width"], cfg["vfov"], cfg["map_m"], cfg["resolution"],
                              mode=cfg["mode"], known_clouds=scn.get("known"))
        outs, sizes = run_mapper(ref.step, scn, world_fn=ref.world)
        Bmax = scn["masks"].shape[1]
        T = len(outs)
        R, C = outs[0][0].shape[1:]
        occ = np.zeros((T, Bmax, R, C), dtype=np.uint8)
        sem = np.zeros((T, Bmax, R, C), dtype=np.uint8)
        for t, (o, s) in enumerate(outs):
            occ[t, : o.shape[0]] = o
            sem[t, : s.shape[0]] = s
        wb, wxyz, wsem = ref.world()
        save = dict(cfg=json.dumps(cfg), num_envs=scn["num_envs"], pose=scn["pose"],
                    orientation=scn["orientation"], masks=scn["masks"],
                    ref_occupancy=occ, ref_semantic=sem, ref_world_sizes=np.asarray(sizes, dtype=np.int64),
                    ref_world_b=wb, ref_world_xyz=wxyz, ref_world_sem=wsem)
        for k in ("depth", "labels", "logits", "labels_for_map"):
            if k in scn:
                save[k] = scn[k]
        if "known" in scn:
            save["env_names"] = np.asarray(scn["env_names"])
            for kn, (xyz, s) in scn["known"].items():
                save[f"known_xyz_{kn}"] = xyz
                save[f"known_sem_{kn}"] = s
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **save)
        print(f"{name}: T={T} B={Bmax} map={R}x{C} world={sizes[-1]} occ_cells={int(occ.sum())} "
              f"sem_cells={int((sem > 0).sum())} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
