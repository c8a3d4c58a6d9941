"""scratch: per-CTA timeline of the LAST of a run of back-to-back steps (programmatic dependent launches stay
overlapped), relative to the moment the previous kernel completed (= first CTA past the dependency wait)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "pred16"
cfg = dict(bench.WORKLOADS[wl]); dev = torch.device("cuda:0"); B = cfg["envs"]
pose, orient, masks = bench.make_poses(cfg, int(os.environ.get("NWALK", "8000")), 1002)
depth, sem = bench.make_frames(cfg, dev, 1002)
pose_d, orient_d, masks_d = (torch.from_numpy(x).to(dev) for x in (pose, orient, masks))
mm = bench.build_module(cfg, dev, B, 0, os.environ.get("IVM_PIPELINED", "1") != "0")
names = [f"s{b}" for b in range(B)]
def step(t): bench.call_module(mm, cfg, names, masks_d[t], pose_d[t], orient_d[t], depth[t % 4], sem[t % 4])
W0 = int(os.environ.get("WARM", "4500"))
for t in range(W0): step(t)
torch.cuda.synchronize()
nct = 296 if cfg["pred"] else 444
acc = None; N = int(os.environ.get('REPS', '12')); t = W0; c0 = []
BURST = 32 if (int(os.environ.get('IVM_DEBUG_FLAGS', '0')) & 4096) else 30
for rep in range(N):
    torch.cuda.synchronize()
    for _ in range(BURST):
        step(t); t += 1
    tr = mm.cta_trace_ns(nct).astype(np.float64)
    tr = tr[tr[:, 4] > 0]
    t0 = tr[:, 4].min()          # first CTA past the dependency wait
    rel = (tr - t0) / 1e3
    rel[:, 13:16] = tr[:, 13:16] / 1e3
    acc = rel if acc is None else acc + rel
    c0.append(round(float(rel[0, 9] - rel[0, 5]), 1))
acc /= N
nm = {7: "resident", 12: "A1.done", 4: "dep.wait passed", 1: "G1.prep", 2: "G1.pixels", 10: "G1.done", 0: "bar1.passed", 3: "G3.done", 6: "argmax.done", 5: "bar2.passed", 11: "D1.end", 8: "D.start", 9: "D.end"}
print(f"{wl}: stamp            min    mean     max   (us since the previous kernel completed, mean over {N} runs of 30 steps)")
for k in [7, 12, 4, 1, 2, 10, 0, 3, 6, 5, 11, 8, 9]:
    if k == 6 and not cfg["pred"]: continue
    print(f"{nm[k]:16s} {acc[:, k].min():7.2f} {acc[:, k].mean():7.2f} {acc[:, k].max():7.2f}")
order_end = np.argsort(-acc[:, 9])[:5]
print("CTA 0: D.end - bar2 per rep:", c0)
print("slowest CTAs (D.end):", [(int(i), round(float(acc[i, 9]), 2)) for i in order_end])
print("period estimate = last D.end =", round(float(acc[:, 9].max()), 2))
if os.environ.get("IVM_TL_DUMP"):
    np.set_printoptions(linewidth=250, precision=1, suppress=True)
    for k in (7, 12, 4, 10, 3, 11, 9):
        print(nm[k], "by CTA index (every 8th):")
        print(acc[::8, k])
    print("my tiles = total/grid; resident-sorted D.end of the same kernel:")
    print(np.sort(acc[:, 9])[::8])
    print(np.sort(acc[:, 7])[::8])
