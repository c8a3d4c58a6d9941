"""scratch: BASELINE config 5 shape -- known-map registration + rotated ego crop, 64 envs, 0.05 m cells, 1024 x 1024
half-cell stores, 200 k-point scene clouds.  Device time per step and the reference's torch path beside it."""
import os, sys, time, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ivlnce_b200.mapper import EpisodesInfo, MapDimensions, Observations, RobotCurrentState, create_known_mapper
from ivlnce_b200.synthetic import make_known_cloud

B, NPTS, STEPS = 64, 200_000, 400
dev = torch.device("cuda:0")
tmp = tempfile.TemporaryDirectory()
for i in range(8):
    xyz, sem = make_known_cloud(NPTS, 20.0, 27, seed=7000 + i)
    np.savez(os.path.join(tmp.name, f"scene{i}.npz"), xyz=xyz, semantics=sem)
md = MapDimensions(6.4, 6.4, 0.05)
mm = create_known_mapper(dev, md, tmp.name, known_capacity=1 << 18, store_cells=1024, max_envs=B, trig="kernel", raster_tile=int(os.environ.get("KTILE", "0")))
names = [f"scene{b % 8}" for b in range(B)]
rng = np.random.default_rng(5)
pose = torch.from_numpy(np.stack([rng.uniform(-4, 4, (STEPS + 50, B)), np.full((STEPS + 50, B), 1.25), rng.uniform(-4, 4, (STEPS + 50, B))], -1).astype(np.float32)).to(dev)
orient = torch.from_numpy(np.stack([np.zeros((STEPS + 50, B)), rng.uniform(-3.1, 3.1, (STEPS + 50, B))], -1)).to(dev)
masks = torch.ones((STEPS + 50, B, 1), dtype=torch.uint8, device=dev); masks[0] = 0
def step(t):
    return mm(EpisodesInfo(masks[t], names), Observations(None, None, None), RobotCurrentState(pose[t], orient[t, :, 0], orient[t, :, 1]))
for t in range(50): out = step(t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(50, 50 + STEPS): out = step(t)
e1.record(); torch.cuda.synchronize()
us = 1e3 * e0.elapsed_time(e1) / STEPS
print(f"known64: {us:.1f} us/step device, {B / us * 1e6:.0f} env-frames/s; occupied cells in the last maps: {int(out.occupancy.sum())}")
