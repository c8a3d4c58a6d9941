"""scratch: known64 -- host submit time vs device time per step."""
import os, sys, time, math, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ivlnce_b200.mapper import EpisodesInfo, MapDimensions, Observations, RobotCurrentState, create_known_mapper
from ivlnce_b200.synthetic import make_known_cloud
cfg = dict(bench.WORKLOADS["known64"]); dev = torch.device("cuda:0"); B = cfg["envs"]
tmp = tempfile.TemporaryDirectory()
for s in range(cfg["scenes"]):
    xyz, sem = make_known_cloud(cfg["points"], cfg["store"] * cfg["res"] / 2 - 0.2, cfg["classes"], seed=7 + 31 * s)
    np.savez(os.path.join(tmp.name, f"scene{s}.npz"), xyz=xyz, semantics=sem)
names = [f"scene{b % cfg['scenes']}" for b in range(B)]
pose, orient, masks = bench.make_poses(cfg, 1200, 9)
pose_d, orient_d = torch.from_numpy(pose).to(dev), torch.from_numpy(orient).to(dev)
masks_h = torch.from_numpy(masks)
mm = create_known_mapper(dev, MapDimensions(cfg["map_m"], cfg["map_m"], cfg["res"]), tmp.name, store_cells=cfg["store"],
                         known_capacity=cfg["points"] + 1024, max_envs=B)
def step(t):
    return mm(EpisodesInfo(masks_h[t].view(-1, 1), names), Observations(None, None, None),
              RobotCurrentState(pose_d[t], orient_d[t, :, 0], orient_d[t, :, 1]))
for t in range(50): step(t)
torch.cuda.synchronize()
K = 500
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for t in range(50, 50 + K): step(t)
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"submit {1e6*(t1-t0)/K:.1f} us/step (CPU), device window {1e3*e0.elapsed_time(e1)/K:.1f} us/step")
mm.set_timing(True); mm.stage_times(reset=True)
for t in range(600, 700): step(t)
torch.cuda.synchronize()
ms, n = mm.stage_times(reset=True)
print("per-kernel (events, serialised): pose %.1f us, raster %.1f us" % (1e3 * ms[0] / max(n[0], 1), 1e3 * ms[4] / max(n[4], 1)))
mm.set_timing(False)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for t in range(700, 1000): step(t)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(10)
