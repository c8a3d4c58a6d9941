"""scratch: step time as a function of the map's age (steps since the reset)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "pred16"
cfg = dict(bench.WORKLOADS[wl]); dev = torch.device("cuda:0"); B = cfg["envs"]
N = 330
pose, orient, masks = bench.make_poses(cfg, N + 10, 1002)
depth, sem = bench.make_frames(cfg, dev, 1002)
pose_d, orient_d, masks_d = (torch.from_numpy(x).to(dev) for x in (pose, orient, masks))
mm = bench.build_module(cfg, dev, B, 0, os.environ.get("IVM_PIPELINED", "1") != "0")
names = [f"s{b}" for b in range(B)]
def step(t): bench.call_module(mm, cfg, names, masks_d[t], pose_d[t], orient_d[t], depth[t % 4], sem[t % 4])
for rep in range(2):
    for t in range(40): step(t)   # warm (ends mid-walk; t=0 below resets)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
    evs[0].record()
    for t in range(N):
        step(t); evs[t + 1].record()
    torch.cuda.synchronize()
    d = np.array([evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(N)])
    print(f"rep {rep} per-step GPU deltas by age:", " ".join(f"{a}-{a+10}:{d[a:a+10].mean():.1f}" for a in range(0, N, 10)))
# in-kernel phase sums by age (sync per step)
ph = []
st = []
for t in range(N):
    step(t)
    ns = mm.phase_ns()
    ph.append((ns[5] - ns[0]) / 1e3)
    if t % 10 == 9: st.append(mm.status()[1])
ph = np.array(ph)
print("in-kernel (start..last CTA end) by age:", " ".join(f"{a}:{ph[a:a+10].mean():.1f}" for a in range(0, N, 10)))
print("stats valid/local/world/in/e1/e2 by age:", [(s[0], s[1], s[2], s[3], s[4], s[5]) for s in st[::3]])
