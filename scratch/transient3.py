"""scratch: bench.py's exact warm-up / timed-region sequence with per-step GPU deltas."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = "pred16"
cfg = dict(bench.WORKLOADS[wl]); dev = torch.device("cuda:0"); B = cfg["envs"]
Wm, K = 5, 20
total = 2 * (Wm + K) + 64
pose, orient, masks = bench.make_poses(cfg, total, 1002)
depth, sem = bench.make_frames(cfg, dev, 1002)
pose_d, orient_d, masks_d = (torch.from_numpy(x).to(dev) for x in (pose, orient, masks))
mm = bench.build_module(cfg, dev, B, 0, os.environ.get("IVM_PIPELINED", "1") != "0")
names = [f"s{b}" for b in range(B)]
def step(t): bench.call_module(mm, cfg, names, masks_d[t], pose_d[t], orient_d[t], depth[t % 4], sem[t % 4])
for rep in range(3):
    t = 0
    t_w = time.perf_counter()
    while True:
        for _ in range(max(Wm, 3)):
            step(t % (Wm + 8)); t += 1
        torch.cuda.synchronize(dev)
        if time.perf_counter() - t_w > 0.4:
            break
    print("warm-up steps", t, "last index", (t - 1) % (Wm + 8))
    t = Wm + 8
    mm.check_errors()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    torch.cuda.synchronize()
    evs[0].record()
    for i in range(K):
        step(t); t += 1
        if rep > 0: evs[i + 1].record()
    if rep == 0: evs[K].record()
    torch.cuda.synchronize()
    if rep == 0:
        print(f"no inner events: {1e3*evs[0].elapsed_time(evs[K])/K:.1f} us/step")
    else:
        gpu = np.array([evs[0].elapsed_time(e) * 1e3 for e in evs])
        print(f"total {gpu[-1]/K:.1f} us/step; deltas:", " ".join(f"{x:.0f}" for x in np.diff(gpu)))
    print("status", mm.status())
