"""Full-size golden fixtures from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_full.py

BASELINE shapes -- 256x256 depth, 0.05 m cells, 128x128 ego maps -- at 16 envs (GT labels; 40-class scores) and a
coherent-scene case (walls on the world bounding box: the reference's key-collision quirk at full size).  The inputs
are NOT stored: they are regenerated from the seeds by tests/full_scenarios.py (same numpy generator streams); the
fixtures hold only the reference's outputs -- both maps of every step, the world-cloud size of every step and a
SHA-256 of the final world cloud (env ids, xyz bits, labels, in the reference's list order).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from full_scenarios import FULL_SCENARIOS, build_full, world_digest  # noqa: E402
from ref_harness import ReferenceRunner  # noqa: E402
from scenarios import run_mapper  # noqa: E402


def main():
    torch.set_num_threads(1)
    os.makedirs(os.path.join(HERE, "full"), exist_ok=True)
    for name in FULL_SCENARIOS:
        t0 = time.time()
        scn = build_full(name)
        cfg = scn["cfg"]
        if "logits" in scn:  # PredictSemantics tail, mapper.py:795-798, with torch's own argmax
            lg = torch.from_numpy(scn["logits"])
            T, B = lg.shape[:2]
            lab = lg.reshape(T * B, *lg.shape[2:]).argmax(1, keepdims=True).to(torch.uint8)
            scn["labels_for_map"] = lab.reshape(T, B, *lg.shape[3:]).numpy()
        ref = ReferenceRunner(cfg["height"], cfg["width"], cfg["vfov"], cfg["map_m"], cfg["resolution"])
        outs, sizes = run_mapper(ref.step, scn, world_fn=ref.world)
        occ = np.stack([o for o, _ in outs])
        sem = np.stack([s for _, s in outs])
        wb, wxyz, wsem = ref.world()
        path = os.path.join(HERE, "full", f"{name}.npz")
        np.savez_compressed(path, spec=json.dumps(FULL_SCENARIOS[name]), ref_occupancy=np.packbits(occ, axis=-1), ref_semantic=sem,
                            ref_world_sizes=np.asarray(sizes, dtype=np.int64), ref_world_sha256=world_digest(wb, wxyz, wsem),
                            ref_world_head_xyz=wxyz[:256], ref_world_head_b=wb[:256], ref_world_head_sem=wsem[:256])
        print(f"{name}: T={len(outs)} B={occ.shape[1]} world={sizes[-1]} occ_cells={int(occ.sum())} "
              f"sem_cells={int((sem > 0).sum())} -> {os.path.getsize(path) / 1024:.0f} KiB in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
