"""scratch: per-step log of CTA 0 over a burst of 250 steps: period and its parts."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "gt32"
cfg = dict(bench.WORKLOADS[wl]); dev = torch.device("cuda:0")
if os.environ.get("ENVS"): cfg["envs"] = int(os.environ["ENVS"])
B = cfg["envs"]
pose, orient, masks = bench.make_poses(cfg, 8000, 1002)
depth, sem = bench.make_frames(cfg, dev, 1002)
pose_d, orient_d, masks_d = (torch.from_numpy(x).to(dev) for x in (pose, orient, masks))
mm = bench.build_module(cfg, dev, B, 0, os.environ.get("IVM_PIPELINED", "1") != "0")
names = [f"s{b}" for b in range(B)]
def step(t): bench.call_module(mm, cfg, names, masks_d[t], pose_d[t], orient_d[t], depth[t % 4], sem[t % 4])
W0 = int(os.environ.get("WARM", "4500"))
for t in range(W0): step(t)
torch.cuda.synchronize()
K = 250
for t in range(W0, W0 + K): step(t)
tr = mm.cta_trace_ns(1024)[512:768].astype(np.float64) / 1e3
order = np.argsort(tr[:, 0]); tr = tr[order][-K + 10:]       # the burst's steps in time order, minus the transient
t0, b1, b2, fx, de, res = (tr[:, i] for i in range(6))
per = np.diff(t0)
np.set_printoptions(linewidth=220, precision=1, suppress=True)
print(f"{wl} envs={B}: period mean {per.mean():.1f} (min {per.min():.1f} max {per.max():.1f})")
print("  t0->bar1 %.1f  bar1->bar2 %.1f  bar2->fix end %.1f  bar2->max D.end %.1f  max D.end->next t0 %.1f  resident->t0 %.1f" % (
    (b1 - t0).mean(), (b2 - b1).mean(), (fx - b2)[fx > 0].mean(), (de - b2).mean(), (t0[1:] - de[:-1]).mean(), (t0 - res).mean()))
print("  fix end - bar2 histogram (us):", np.histogram((fx - b2)[fx > 0], bins=[0, 4, 8, 12, 16, 24, 32, 64])[0])
print("  period sample:", per[100:130])
