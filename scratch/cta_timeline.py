"""scratch: per-CTA timeline of the fused step kernel (run on the GPU box)."""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
cfg = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "pred16"])
dev = torch.device("cuda:0")
B = cfg["envs"]
pose, orient, masks = bench.make_poses(cfg, 400, 1002)
depth, sem = bench.make_frames(cfg, dev, 1002)
pose_d, orient_d, masks_d = (torch.from_numpy(x).to(dev) for x in (pose, orient, masks))
mm = bench.build_module(cfg, dev, B, 0)
names = [f"s{b}" for b in range(B)]
for t in range(300):
    bench.call_module(mm, cfg, names, masks_d[t], pose_d[t], orient_d[t], depth[t % 4], sem[t % 4])
acc = None
N = 16
for t in range(300, 300 + N):
    bench.call_module(mm, cfg, names, masks_d[t], pose_d[t], orient_d[t], depth[t % 4], sem[t % 4])
    tr = mm.cta_trace_ns(296).astype(np.float64)
    t0 = tr[:, 7].min()
    rel = (tr[:, :11] - t0) / 1e3
    acc = rel if acc is None else acc + rel
acc /= N
names_s = ["B.start", "B.slots", "B.filter", "B.drain", "B.flush", "B.pastbar", "A.done", "start", "D.start", "D.end", "A0.done"]
order = [7, 10, 6, 0, 1, 2, 3, 4, 5, 8, 9]
print("stamp      min    mean     max   (us since first CTA start, mean over %d steps)" % N)
for k in order:
    print(f"{names_s[k]:10s} {acc[:, k].min():7.2f} {acc[:, k].mean():7.2f} {acc[:, k].max():7.2f}")
d = lambda a, b: acc[:, a] - acc[:, b]
for nm, a, b in [("A0 work", 10, 7), ("A work", 6, 7), ("barrier1 wait", 0, 6), ("B setup", 1, 0), ("B filter", 2, 1), ("B drain", 3, 2), ("B flush", 4, 3), ("barrier2 wait", 5, 4), ("D work", 9, 8)]:
    x = d(a, b)
    print(f"{nm:14s} min {x.min():6.2f} mean {x.mean():6.2f} max {x.max():6.2f}")
