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
mm = bench.build_module(cfg, dev, B, 0, os.environ.get("IVM_PIPELINED", "1") != "0")
names = [f"s{b}" for b in range(B)]
for t in range(300):
    bench.call_module(mm, cfg, names, masks_d[t], pose_d[t], orient_d[t], depth[t % 4], sem[t % 4])
acc = None
N = 16
for t in range(300, 300 + N):
    bench.call_module(mm, cfg, names, masks_d[t], pose_d[t], orient_d[t], depth[t % 4], sem[t % 4])
    tr = mm.cta_trace_ns(444 if not cfg['pred'] else 296).astype(np.float64)
    t0 = tr[:, 7].min()
    rel = (tr[:, :16] - t0) / 1e3
    rel[:, 13:16] = tr[:, 13:16] / 1e3
    acc = rel if acc is None else acc + rel
acc /= N
# stamps of k_step_overlap: 7 start, 10 G1 (scatter) done, 0 past barrier 1, 3 G3 (resolve) done, 6 argmax warps done,
# 5 past barrier 2, 8 raster start, 9 end
names = {7: "start", 12: "A1.done", 4: "dep.passed", 1: "G1.prep", 2: "G1.pixels", 10: "G1.done", 0: "bar1.passed", 3: "G3.done", 6: "argmax.done", 5: "bar2.passed", 11: "D1.end", 8: "D.start", 9: "D.end"}
print("stamp          min    mean     max   (us since first CTA start, mean over %d steps)" % N)
for k in [7, 12, 4, 1, 2, 10, 0, 3, 6, 5, 11, 8, 9]:
    print(f"{names[k]:12s} {acc[:, k].min():7.2f} {acc[:, k].mean():7.2f} {acc[:, k].max():7.2f}")
d = lambda a, b: acc[:, a] - acc[:, b]
for nm, a, b in [("prep+A1 filter", 12, 7), ("state prep", 1, 4), ("A2+B", 2, 1), ("G1 flush", 10, 2), ("G1 work", 10, 7), ("barrier1 wait", 0, 10), ("G3 work", 3, 0), ("argmax total", 6, 7), ("G3 tail after argmax", 3, 6), ("barrier2 wait", 5, 3), ("safe raster / fix-up", 11, 5), ("release wait", 8, 11), ("D2 work", 9, 8)]:
    x = d(a, b)
    print(f"{nm:22s} min {x.min():6.2f} mean {x.mean():6.2f} max {x.max():6.2f}")

for nm, k in [("raster: zero+geom+spans (sum over the group's tiles)", 13), ("raster: record loop", 14), ("raster: write-out", 15)]:
    x = acc[1:, k]
    print(f"{nm:55s} min {x.min():6.2f} mean {x.mean():6.2f} max {x.max():6.2f}")

order_end = np.argsort(-acc[:, 9])[:5]
print("slowest CTAs (D.end):", [(int(i), round(float(acc[i, 9]), 2)) for i in order_end])
