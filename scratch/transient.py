"""scratch: where do the extra ~13 us/step of a 20-step timed region (vs 300 steps) come from?
Per-step host submit stamps and per-step GPU completion stamps (events between the steps) after a sync."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
if os.environ.get("TILE"):   # raster tile override (cells per side)
    import ivlnce_b200.mapper as _M
    _orig = _M.MappingModule.__init__
    def _init(self, *a, **k):
        k["raster_tile"] = int(os.environ["TILE"]); _orig(self, *a, **k)
    _M.MappingModule.__init__ = _init
wl = sys.argv[1] if len(sys.argv) > 1 else "pred16"
cfg = dict(bench.WORKLOADS[wl]); dev = torch.device("cuda:0")
if os.environ.get("ENVS"): cfg["envs"] = int(os.environ["ENVS"])
B = cfg["envs"]
pose, orient, masks = bench.make_poses(cfg, 8000, 1002)
depth, sem = bench.make_frames(cfg, dev, 1002)
pose_d, orient_d, masks_d = (torch.from_numpy(x).to(dev) for x in (pose, orient, masks))
mm = bench.build_module(cfg, dev, B, 0, os.environ.get("IVM_PIPELINED", "1") != "0")
names = [f"s{b}" for b in range(B)]
def step(t): bench.call_module(mm, cfg, names, masks_d[t], pose_d[t], orient_d[t], depth[t % 4], sem[t % 4])
for t in range(4500): step(t)
torch.cuda.synchronize()
K = 24
for rep in range(3):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    host = np.zeros(K + 1)
    torch.cuda.synchronize()
    if rep == 2:
        time.sleep(0.5)   # idle GPU before the region
    t0 = time.perf_counter(); evs[0].record()
    for i in range(K):
        step(4500 + i); evs[i + 1].record(); host[i + 1] = time.perf_counter() - t0
    torch.cuda.synchronize()
    gpu = np.array([evs[0].elapsed_time(e) * 1e3 for e in evs])
    print(f"rep {rep} {wl}: total {gpu[-1]:.0f} us = {gpu[-1]/K:.1f} us/step")
    print("  host submit (us since start):", " ".join(f"{x*1e6:.0f}" for x in host[1:]))
    print("  gpu done    (us since start):", " ".join(f"{x:.0f}" for x in gpu[1:]))
    print("  gpu deltas:", " ".join(f"{x:.0f}" for x in np.diff(gpu)))
# without events between the steps
for K in (20, 100, 300):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): step(4600 + i)
    e1.record(); torch.cuda.synchronize()
    print(f"K={K}: {1e3*e0.elapsed_time(e1)/K:.1f} us/step")
