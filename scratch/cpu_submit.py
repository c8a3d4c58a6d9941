"""scratch: is the bench loop CPU-bound?  wall time of K submissions (no sync) vs device time."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "pred16"
cfg = dict(bench.WORKLOADS[wl]); dev = torch.device("cuda:0"); B = cfg["envs"]
pose, orient, masks = bench.make_poses(cfg, 8000, 1002)
depth, sem = bench.make_frames(cfg, dev, 1002)
pose_d, orient_d, masks_d = (torch.from_numpy(x).to(dev) for x in (pose, orient, masks))
mm = bench.build_module(cfg, dev, B, 0, os.environ.get("IVM_PIPELINED", "1") != "0")
names = [f"s{b}" for b in range(B)]
def step(t): bench.call_module(mm, cfg, names, masks_d[t], pose_d[t], orient_d[t], depth[t % 4], sem[t % 4])
for t in range(200): step(t)
torch.cuda.synchronize()
K = 2000
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for t in range(200, 200 + K): step(t)
e1.record(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"{wl}: submit {1e6*(t1-t0)/K:.1f} us/step (CPU), device {1e3*e0.elapsed_time(e1)/K:.1f} us/step, wall incl. drain {1e6*(t2-t0)/K:.1f} us/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for t in range(200, 700): step(t)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
