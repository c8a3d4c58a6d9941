"""scratch: why is a rank's pred16 step ~10 us slower under torchrun (N >= 2) than alone?
torchrun --nproc-per-node N scratch/scale_diag.py [--no-pin] [--smi] [--no-dist-barrier]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
flags = set(sys.argv[1:])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)
cores = sorted(os.sched_getaffinity(0))
if "--no-pin" not in flags and world > 1:
    per = len(cores) // world
    os.sched_setaffinity(0, cores[lr * per:(lr + 1) * per])
cfg = dict(bench.WORKLOADS["pred16"]); B = cfg["envs"]
pose, orient, masks = bench.make_poses(cfg, 8000, 1002 + (rank if "--seeds" in flags else 0))
depth, sem = bench.make_frames(cfg, dev, 1002 + (rank if "--seeds" in flags else 0))
pose_d, orient_d, masks_d = (torch.from_numpy(x).to(dev) for x in (pose, orient, masks))
mm = bench.build_module(cfg, dev, B, 0, True)
names = [f"s{b}" for b in range(B)]
def step(t): bench.call_module(mm, cfg, names, masks_d[t], pose_d[t], orient_d[t], depth[t % 4], sem[t % 4])
for t in range(4500): step(t)
torch.cuda.synchronize()
smi = bench.ClockSampler(list(range(world))).start() if ("--smi" in flags and rank == 0) else None
if "--smi0" in flags and rank == 0:
    smi = bench.ClockSampler([0]).start()
nv = None
if "--nvml" in flags:
    import threading, pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(lr)
    nv = {"stop": False, "n": 0, "sm": [], "reasons": 0, "dt": []}
    def loop():
        while not nv["stop"]:
            a = time.perf_counter()
            nv["sm"].append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            nv["reasons"] |= pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            nv["dt"].append(time.perf_counter() - a)
            time.sleep(0.02)
    th = threading.Thread(target=loop, daemon=True); th.start()
def bar():
    if world > 1 and "--no-dist-barrier" not in flags:
        torch.distributed.barrier()
    torch.cuda.synchronize()
t = 4500
for K in (20, 300, 20, 300, 20, 20, 20, 100, 20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bar()
    h0 = time.perf_counter(); e0.record()
    for i in range(K): step(t); t += 1
    e1.record(); h1 = time.perf_counter()
    bar()
    ms = e0.elapsed_time(e1)
    out = torch.tensor([ms * 1e3 / K, (h1 - h0) * 1e6 / K], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.zeros_like(out) for _ in range(world)]
        torch.distributed.all_gather(allv, out)
    else:
        allv = [out]
    if rank == 0:
        print(f"K={K:4d} dev us/step per rank: " + " ".join(f"{v[0].item():6.1f}" for v in allv) +
              "   host submit us/step: " + " ".join(f"{v[1].item():5.1f}" for v in allv), flush=True)
if smi is not None:
    print(smi.stop())
if nv is not None:
    nv["stop"] = True; th.join()
    print(rank, "nvml samples", len(nv["sm"]), "median sm", np.median(nv["sm"]), "reasons", hex(nv["reasons"]), "call ms median/max", 1e3*np.median(nv["dt"]), 1e3*max(nv["dt"]))
if rank == 0:
    print("cores total", len(cores), "flags", sorted(flags))
if world > 1:
    torch.distributed.destroy_process_group()
