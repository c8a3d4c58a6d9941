"""scratch: fix-up milestones of CTA 0 (ttrace) relative to grid barrier 2, mean over single steps of a burst."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "gt1"
cfg = dict(bench.WORKLOADS[wl]); dev = torch.device("cuda:0")
if os.environ.get("ENVS"): cfg["envs"] = int(os.environ["ENVS"])
B = cfg["envs"]
pose, orient, masks = bench.make_poses(cfg, 6000, 1002)
depth, sem = bench.make_frames(cfg, dev, 1002)
pose_d, orient_d, masks_d = (torch.from_numpy(x).to(dev) for x in (pose, orient, masks))
mm = bench.build_module(cfg, dev, B, 0, True)
names = [f"s{b}" for b in range(B)]
def step(t): bench.call_module(mm, cfg, names, masks_d[t], pose_d[t], orient_d[t], depth[t % 4], sem[t % 4])
for t in range(3000): step(t)
torch.cuda.synchronize()
acc = []; t = 3000
for rep in range(20):
    for _ in range(7 + rep % 5):
        step(t); t += 1
    ph = np.array(mm.phase_ns(), dtype=np.float64); ft = np.array(mm.fixup_trace_ns(), dtype=np.float64)
    b2 = ph[2]
    acc.append(np.concatenate([(ph - b2) / 1e3, (ft - b2) / 1e3]))
acc = np.array(acc)
np.set_printoptions(linewidth=220, precision=2, suppress=True)
print(wl, "tstamp[0..7] (start, bar1, bar2, fix3, fix4, maxDend, depwait, -) rel. bar2, us:"); print(acc[:, :8].mean(0))
print("ttrace[0..15] rel. bar2, us (median):"); print(np.median(acc[:, 8:], 0))
flags, stats = mm.status(); print("stats", stats)
print("per sample: seg, scan_end, del_end, fix_end, maxDend (us after bar2)")
print(np.stack([acc[:, 8], acc[:, 9], acc[:, 10], acc[:, 3], acc[:, 5]], 1).T)
