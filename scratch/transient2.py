"""scratch: replicate bench.py's timed-region preamble; which part causes the one-off cost?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = "pred16"
cfg = dict(bench.WORKLOADS[wl]); dev = torch.device("cuda:0"); B = cfg["envs"]
pose, orient, masks = bench.make_poses(cfg, 3000, 1002)
depth, sem = bench.make_frames(cfg, dev, 1002)
pose_d, orient_d, masks_d = (torch.from_numpy(x).to(dev) for x in (pose, orient, masks))
mm = bench.build_module(cfg, dev, B, 0, os.environ.get("IVM_PIPELINED", "1") != "0")
names = [f"s{b}" for b in range(B)]
def step(t): bench.call_module(mm, cfg, names, masks_d[t], pose_d[t], orient_d[t], depth[t % 4], sem[t % 4])
def region(K, label, pre):
    for t in range(30): step(t)
    torch.cuda.synchronize()
    pre()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    h0 = time.perf_counter()
    e0.record()
    hs = []
    for i in range(K):
        step(100 + i); hs.append(time.perf_counter() - h0)
    e1.record(); torch.cuda.synchronize()
    print(f"{label}: {1e3*e0.elapsed_time(e1)/K:.1f} us/step; host first 3 submits at", [round(x*1e6) for x in hs[:3]], "last", round(hs[-1]*1e6))
for rep in range(2):
    region(20, "plain", lambda: None)
    region(20, "check_errors", lambda: mm.check_errors())
    region(20, "status+launches", lambda: (mm.status(), mm.kernel_launches()))
    s = bench.ClockSampler(0); s.start(); time.sleep(0.3)
    region(20, "sampler running", lambda: None)
    region(20, "sampler running", lambda: None)
    region(20, "sampler+check", lambda: mm.check_errors())
    print(s.stop())
    region(20, "sleep 50ms", lambda: time.sleep(0.05))
    region(20, "sleep 5ms", lambda: time.sleep(0.005))
