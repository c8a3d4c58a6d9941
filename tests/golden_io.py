"""Load a tests/golden/<name>.npz fixture back into the scenario dict layout of
tests/scenarios.py (+ the reference outputs)."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False)
    scn = {k: z[k] for k in z.files if not k.startswith("known_") and k not in ("cfg", "env_names")}
    scn["cfg"] = json.loads(str(z["cfg"]))
    if "env_names" in z.files:
        scn["env_names"] = [list(map(str, row)) for row in z["env_names"]]
        known = {}
        for k in z.files:
            if k.startswith("known_xyz_"):
                n = k[len("known_xyz_"):]
                known[n] = (z[k], z[f"known_sem_{n}"])
        scn["known"] = known
    return scn
