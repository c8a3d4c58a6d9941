#!/usr/bin/env python
"""Benchmark of the semantic-map update (BASELINE.json metric: env-frames/s, HBM GB/s vs roofline,
CPU reference beside it).

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference's PyTorch CPU path (port)
    torchrun --nproc-per-node N bench.py --gpus N ...         # N ranks, envs sharded by tour

A "step" is one `MappingModule.forward` over one batch of synthetic frames.  The headline line is
BASELINE.json configs[1]: predicted-semantics map update, 16 envs per GPU (weak scaling), 256x256 depth,
40-class f32 score planes (argmax fused into the step kernel), 0.05 m cells, 128x128 ego map.  The other
BASELINE configs ride in the same JSON line under "configs", each with its step time, roofline fraction, CPU
baseline and an in-bench parity check of the CUDA path against the C oracle on the first steps of the same inputs:

    N = 1:  gt1 (config 1), gt32 (config 4, per-GPU shape at 8 GPUs), tour (config 3: 100 episodes x 60 steps into one
            2048^2 store), known64 (config 5: 64 envs, 1024^2 stores, 2 M-point clouds), pred16_scene (coherent depth)
    N > 1:  gt256 (config 4: 256 envs split by tour over the N ranks)

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BASE = dict(H=256, W=256, res=0.05, map_m=6.4, store=2048, depth="iid", roam=8.0)
WORKLOADS = {
    # BASELINE.json configs[1] -- the configuration the metric is quoted on
    "pred16": dict(BASE, envs=16, pred=True, classes=40,
                   desc="predicted-semantics (40-class f32 scores) map update, 16 envs/GPU, 256x256 depth, "
                        "0.05 m cells, 128x128 ego map"),
    "pred16_scene": dict(BASE, envs=16, pred=True, classes=40, depth="scene",
                         desc="pred16 with coherent depth: a box room with obstacles ray-cast from the poses (walls, floor "
                              "rows: many pixels per half-cell, exact height ties)"),
    "gt1": dict(BASE, envs=1, pred=False, classes=27,
                desc="GT-semantics map update, 1 env, 256x256 depth + 27 labels, 0.05 m cells (BASELINE config 1)"),
    "gt32": dict(BASE, envs=32, pred=False, classes=27,
                 desc="GT-semantics map update, 32 envs/GPU (the per-GPU shape of 256 envs over 8 GPUs), 256x256 depth, 0.05 m cells"),
    "gt256": dict(BASE, envs=256, pred=False, classes=27,
                  desc="GT-semantics map update, 256 envs partitioned by tour over the ranks (BASELINE config 4)"),
    "tour": dict(BASE, envs=1, pred=False, classes=27, roam=10.0, tour_steps=6000,
                 desc="iterative-map accumulation over a 100-episode tour (100 x 60 steps, one reset at the start) into one "
                      "persistent 2048x2048 half-cell store, 1 env (BASELINE config 3)"),
    "known64": dict(H=0, W=0, res=0.1, map_m=6.4, store=1024, envs=64, pred=False, classes=13, points=2_000_000, scenes=8,
                    roam=18.0, depth="none",
                    desc="known-map registration + ego raster, 64 envs, 1024x1024 half-cell stores (51.2 m at 0.05 m), scene "
                         "clouds of 2 M points (BASELINE config 5)"),
}
RING = 4  # distinct resident input frames per env (4 x 168 MB for pred16: larger than the 126 MB L2)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pred16", choices=sorted(WORKLOADS), help="headline workload")
    ap.add_argument("--envs-per-gpu", type=int, default=0)
    ap.add_argument("--cpu-steps", type=int, default=12, help="steps of the CPU baseline sample (ours arm)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="wall budget of the reference arm")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-configs", action="store_true", help="headline workload only")
    ap.add_argument("--only", default="", help="comma list of extra configs to run (default: all that apply)")
    ap.add_argument("--seed", type=int, default=1002)
    ap.add_argument("--store", type=int, default=0, help="override the world-store extent (half-cells per side)")
    ap.add_argument("--no-pipeline", action="store_true", help="do not overlap consecutive steps (ivm_set_pipelined off)")
    ap.add_argument("--variant", type=int, default=0,
                    help="0 = fused persistent step kernel (default), 1/2 = four-kernel step (register / bulk-async loads)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ inputs
def make_poses(cfg, steps, seed):
    """Random walk of 0.25 m forward steps / 15 degree turns per env, kept within `roam` metres
    of its start (a house-sized area).  Every env lives in its own scene, and scene coordinates all lie near the
    origin (as in MP3D), so the envs' world coordinates overlap -- this keeps the reference's de-dup key space
    (batch-global bbox x envs, mapper.py:468-469) at a realistic size.
    pose f32 [S,B,3], orientation f64 [S,B,2], masks u8 [S,B] (0 only at t = 0)."""
    from ivlnce_b200.synthetic import ScenarioConfig, random_walk, reset_masks

    sc = ScenarioConfig(num_envs=cfg["envs"], height=max(cfg["H"], 1), width=max(cfg["W"], 1), steps=steps,
                        resolution=cfg["res"], map_meters=cfg["map_m"], num_labels=cfg["classes"], seed=seed,
                        env_spacing=0.0, roam_radius=cfg["roam"])
    pose, orient = random_walk(sc, np.random.default_rng(seed))
    return pose, orient, reset_masks(sc)


def make_frames(cfg, device, seed):
    """RING distinct frames: depth f32 [RING,B,1,H,W] ~ U(0.05,0.95); scores f32 [RING,B,Cls,H,W] ~ N(0,1)
    (pred) or labels u8 [RING,B,1,H,W]."""
    g = torch.Generator(device=device).manual_seed(seed)
    B, H, W = cfg["envs"], cfg["H"], cfg["W"]
    depth = torch.rand((RING, B, 1, H, W), generator=g, device=device) * 0.9 + 0.05
    if cfg["pred"]:
        sem = torch.randn((RING, B, cfg["classes"], H, W), generator=g, device=device)
    else:
        sem = torch.randint(0, cfg["classes"], (RING, B, 1, H, W), generator=g, device=device, dtype=torch.uint8)
    return depth, sem


def make_scene_depth(cfg, device, seed, steps):
    """Coherent depth: one box room with obstacles (ivlnce_b200.synthetic.BoxRoom), every env walks in it, depth
    ray-cast on the device from the walk's poses.  Returns pose f32 [S,B,3], orient f64 [S,B,2] (numpy) and depth
    f32 [S,B,1,H,W] (device); replayed cyclically by the caller."""
    from ivlnce_b200.synthetic import BoxRoom, ScenarioConfig, camera_tables, random_walk

    B, H, W = cfg["envs"], cfg["H"], cfg["W"]
    rng = np.random.default_rng(seed)
    room = BoxRoom(rng, 13)
    sc = ScenarioConfig(num_envs=B, height=H, width=W, steps=steps, resolution=cfg["res"], map_meters=cfg["map_m"],
                        seed=seed, env_spacing=0.0)
    pose, orient = random_walk(sc, rng, [room] * B)
    xs, ys = camera_tables(H, W, math.pi / 2)
    f64 = torch.float64
    dcam = torch.stack([torch.from_numpy(xs).to(device, f64)[None, :].expand(H, W),
                        torch.from_numpy(ys).to(device, f64)[:, None].expand(H, W),
                        torch.ones(H, W, dtype=f64, device=device)], -1).reshape(-1, 3)          # [HW,3]
    ex = torch.from_numpy(orient[..., 0].astype(np.float64)).to(device) + math.pi
    hd = torch.from_numpy(orient[..., 1].astype(np.float64)).to(device)
    cx, sx, cy, sy = torch.cos(ex), torch.sin(ex), torch.cos(hd), torch.sin(hd)
    zero = torch.zeros_like(cx)
    R = torch.stack([torch.stack([cy, sx * sy, cx * sy], -1), torch.stack([zero, cx, -sx], -1),
                     torch.stack([-sy, cy * sx, cy * cx], -1)], -2)                              # [S,B,3,3]
    planes = [(0, -room.half[0]), (0, room.half[0]), (2, -room.half[1]), (2, room.half[1]), (1, 0.0), (1, room.ceiling)]
    boxes = torch.from_numpy(room.boxes).to(device)
    depth = torch.empty((steps, B, 1, H, W), dtype=torch.float32, device=device)
    for t in range(steps):
        origin = torch.from_numpy(pose[t].astype(np.float64)).to(device)                         # [B,3]
        dirs = torch.einsum("nk,bjk->bnj", dcam, R[t])                                           # [B,HW,3]
        best = torch.full((B, H * W), float("inf"), dtype=f64, device=device)
        for axis, val in planes:
            tt = (val - origin[:, None, axis]) / dirs[..., axis]
            best = torch.where((tt > 1e-6) & (tt < best), tt, best)
        for bx in boxes:
            lo = (bx[0:3] - origin[:, None, :]) / dirs
            hi = (bx[3:6] - origin[:, None, :]) / dirs
            tmin = torch.minimum(lo, hi).amax(-1)
            tmax = torch.maximum(lo, hi).amin(-1)
            best = torch.where((tmax >= tmin) & (tmin > 1e-6) & (tmin < best), tmin, best)
        depth[t, :, 0] = (best / 10.0).clamp(0.0, 1.0).reshape(B, H, W).to(torch.float32)
    return pose, orient, depth


def build_module(cfg, device, max_envs, variant=0, pipelined=True):
    from ivlnce_b200.mapper import (CameraParameters, MapDimensions, PrecomputedScores,
                                    create_gt_semantics_iterative_mapper, create_iterative_mapper)

    cam = CameraParameters(math.pi / 2, (cfg["H"], cfg["W"]), 0.1)
    md = MapDimensions(cfg["map_m"], cfg["map_m"], cfg["res"])
    kw = dict(store_cells=cfg["store"], max_envs=max_envs, trig="kernel", scatter_variant=variant, pipelined=pipelined)
    if cfg["pred"]:
        return create_iterative_mapper(device, cam, md, PrecomputedScores(), **kw)
    return create_gt_semantics_iterative_mapper(device, cam, md, **kw)


def call_module(mm, cfg, names, masks_t, pose_t, orient_t, depth_t, sem_t):
    """One public-API call: MappingModule.forward(EpisodesInfo, Observations, RobotCurrentState)."""
    from ivlnce_b200.mapper import EpisodesInfo, Observations, RobotCurrentState

    ei = EpisodesInfo(masks_t.view(-1, 1), names)
    obs = Observations(None, depth_t, sem_t) if cfg["pred"] else Observations(sem_t, depth_t, None)
    st = RobotCurrentState(pose_t, orient_t[:, 0], orient_t[:, 1])
    return mm(ei, obs, st)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampling DURING the timed regions (B200_PROFILING.md clocks line): ONE process for all the GPUs of
    the job, started by rank 0 only."""
    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, indices):
        self.indices, self.proc, self.path = list(indices), None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "--id=" + ",".join(str(i) for i in self.indices),
                                          f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm), "gpus": self.indices}
        return out


# ------------------------------------------------------------------------------------------ CPU arm
def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_reference_run(cfg, frames_host, pose, orient, masks, warmup, steps, budget_s):
    """The reference's PyTorch CPU path (oracle/torch_path.py port) on the host cores, all threads."""
    from oracle.torch_path import TorchReferencePath

    threads = host_threads()
    torch.set_num_threads(threads)
    depth_h, sem_h = frames_host
    tp = TorchReferencePath(cfg["H"], cfg["W"], math.pi / 2, cfg["map_m"], cfg["map_m"], cfg["res"])
    B = cfg["envs"]

    def one(t):
        o = torch.from_numpy(orient[t])
        kw = dict(depth=depth_h[t % depth_h.shape[0]])
        if cfg["pred"]:
            kw["scores"] = sem_h[t % sem_h.shape[0]]
        else:
            kw["labels"] = sem_h[t % sem_h.shape[0]]
        return tp.step(torch.from_numpy(masks[t]), torch.from_numpy(pose[t]), o[:, 0], o[:, 1], **kw)

    t = 0
    for _ in range(warmup):
        one(t); t += 1
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        one(t); t += 1; done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=B * done / dt, seconds=dt, steps=done, threads=threads, ms_per_step=1e3 * dt / max(done, 1),
                warmup=warmup)


def peak_hbm():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "hbm_gbs" in peaks:
            return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def traffic_for(workload, kernel):
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        w = tj.get(workload, {})
        v = w.get(kernel)
        return v, (w.get("_source") if v is not None else None)   # the capture the figure comes from (per workload)
    except Exception:
        return None, None


class Env:
    """Process-wide context of a bench run."""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = None

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(self.dev)

    def _reduce(self, x, op):
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            torch.distributed.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, x):
        return self._reduce(x, torch.distributed.ReduceOp.MAX)

    def min_over_ranks(self, x):
        return self._reduce(x, torch.distributed.ReduceOp.MIN)

    def sum_over_ranks(self, x):
        return self._reduce(x, torch.distributed.ReduceOp.SUM)


def pin_cores(env):
    """Each rank keeps to its own share of the host cores (all ranks of a node otherwise share one affinity mask and
    their Python launch loops disturb each other)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", env.world))
        if local_world > 1 and len(cores) >= 2 * local_world:
            per = len(cores) // local_world
            mine = cores[env.local_rank * per:(env.local_rank + 1) * per]
            os.sched_setaffinity(0, mine)
            torch.set_num_threads(max(1, min(per, 8)))
            return mine
        return cores
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ iterative workloads
def run_iterative(env, name, cfg, K, Wm, headline=False, e2e=False, cpu_steps=0, parity_steps=3, seed=0):
    """Builds the inputs and the module of one workload on this rank, checks the first steps against the C oracle,
    times K back-to-back steps (device-resident inputs), optionally the end-to-end path and the CPU port."""
    args, dev, rank, world = env.args, env.dev, env.rank, env.world
    B = cfg["envs"]
    HW = cfg["H"] * cfg["W"]
    R = math.ceil(cfg["map_m"] / cfg["res"])
    pipelined = not args.no_pipeline
    tour = "tour_steps" in cfg
    Kt = cfg["tour_steps"] if tour else K
    n_walk = (Kt + 8) if tour else 4096 + 3 * K + 256
    res = {"workload": f"{name}: {cfg['desc']}", "envs_per_gpu": B}

    # ---- inputs (device-resident)
    if cfg["depth"] == "scene":
        n_scene = 16
        pose_s, orient_s, depth = make_scene_depth(cfg, dev, seed, n_scene)
        reps = (n_walk + n_scene - 1) // n_scene
        pose, orient = np.tile(pose_s, (reps, 1, 1))[:n_walk], np.tile(orient_s, (reps, 1, 1))[:n_walk]
        masks = np.ones((n_walk, B), dtype=np.uint8); masks[0] = 0
        _, sem = make_frames(cfg, dev, seed)
        nd = n_scene
        res["scene"] = f"{n_scene} consistent steps of a walk through one box room, replayed cyclically (the map saturates)"
    else:
        pose, orient, masks = make_poses(cfg, n_walk, seed)
        depth, sem = make_frames(cfg, dev, seed)
        nd = RING
    pose_d, orient_d, masks_d = (torch.from_numpy(x).to(dev) for x in (pose, orient, masks))
    names = [f"scene{rank}_{b}" for b in range(B)]
    mm = build_module(cfg, dev, B, args.variant, pipelined)

    # per-step views of the resident inputs, made once (the timed call is the module call with its three argument objects)
    from ivlnce_b200.mapper import EpisodesInfo, Observations, RobotCurrentState

    masks_v = list(masks_d.unsqueeze(-1).unbind(0))
    pose_v, elev_v, head_v = list(pose_d.unbind(0)), list(orient_d[:, :, 0].unbind(0)), list(orient_d[:, :, 1].unbind(0))
    depth_v, sem_v = list(depth.unbind(0)), list(sem.unbind(0))
    is_pred = bool(cfg["pred"])

    def step(t):
        obs = Observations(None, depth_v[t % nd], sem_v[t % RING]) if is_pred else Observations(sem_v[t % RING], depth_v[t % nd], None)
        return mm(EpisodesInfo(masks_v[t], names), obs, RobotCurrentState(pose_v[t], elev_v[t], head_v[t]))

    # ---- parity: the CUDA path against the C oracle on the first steps of these very inputs (t = 0 resets every env)
    if parity_steps > 0:
        from oracle.oracle import OracleMapper, argmax_labels

        orc = OracleMapper(cfg["H"], cfg["W"], math.pi / 2, cfg["map_m"], cfg["map_m"], cfg["res"])
        occ_eq = sem_eq = True
        tp0 = time.perf_counter()
        for t in range(parity_steps):
            out = step(t)
            occ, smp = out.occupancy.cpu().numpy(), out.semantic.cpu().numpy()
            d_h = depth[t % nd][:, 0].cpu().numpy()
            lab_h = argmax_labels(sem[t % RING].cpu().numpy()) if cfg["pred"] else sem[t % RING][:, 0].cpu().numpy()
            o_ref, s_ref = orc.step(masks[t], pose[t], orient[t], depth=d_h, labels=lab_h)
            occ_eq = occ_eq and bool(np.array_equal(occ, o_ref))
            sem_eq = sem_eq and bool(np.array_equal(smp, s_ref))
        res["parity"] = {"steps_checked": parity_steps, "occupancy_equal": occ_eq, "semantic_equal": sem_eq,
                         "envs_checked": B, "seconds": round(time.perf_counter() - tp0, 2),
                         "against": "oracle/mapper_oracle.c (C restatement of mapper.py:825-947) on the first steps of the "
                                    "timed inputs, per rank"}
        del orc
        if world > 1:
            ok = env.min_over_ranks(1.0 if (occ_eq and sem_eq) else 0.0)
            res["parity"]["all_ranks_equal"] = bool(ok == 1.0)

    # ---- warm-up: at least Wm steps and ~0.3 s of load along the walk (the map matures), then the timed region
    t = 0
    t_w = time.perf_counter()
    while not tour:
        for _ in range(max(Wm, 3)):
            step(t); t += 1
        torch.cuda.synchronize(dev)
        if time.perf_counter() - t_w > 0.3 or t + max(Wm, 3) > 4096:
            break
    warm_steps = t
    if tour:       # a tour starts from an empty map: warm-up on the first steps, then the timed tour resets at t = 0
        for tt in range(max(Wm, 3)):
            step(tt)
        torch.cuda.synchronize(dev)
        warm_steps, t = max(Wm, 3), 0
    mm.check_errors()
    launches0 = mm.kernel_launches()
    marks = [0, Kt // 10, Kt - Kt // 10, Kt] if tour else [0, Kt]
    evs = [torch.cuda.Event(enable_timing=True) for _ in marks]
    env.barrier()
    for i in range(Kt + 1):
        if i in marks:
            evs[marks.index(i)].record()
        if i < Kt:
            step(t); t += 1
    env.barrier()
    ms = evs[0].elapsed_time(evs[-1])
    launches = mm.kernel_launches() - launches0
    flags, stats = mm.status()
    assert flags == 0, f"{name}: map error flags {flags}"
    ms_max = env.max_over_ranks(ms)
    total_envs = env.sum_over_ranks(float(B))
    res.update({"steps": Kt, "warmup_steps": warm_steps, "ms_per_step": ms_max / Kt,
                "value": total_envs * Kt / (ms_max * 1e-3), "gpu_launches": int(launches)})
    if tour:
        first, last = evs[0].elapsed_time(evs[1]) / marks[1], evs[2].elapsed_time(evs[3]) / (marks[3] - marks[2])
        res["tour"] = {"episodes": 100, "steps_per_episode": Kt // 100, "first_10pct_ms_per_step": first,
                       "last_10pct_ms_per_step": last, "world_records_at_end": int(stats[2])}

    # ---- roofline: one kernel per step; its launch duration = timed region / launches (back-to-back launches
    #      overlap when pipelined, so events round each kernel would not time the production schedule)
    p_local, p_in = stats[1] / B, stats[3] / B
    bytes_in = HW * 4 + (HW * 4 * cfg["classes"] if cfg["pred"] else HW)
    bytes_frame = bytes_in + 2 * R * R + 16 * (2 * p_local + p_in)
    peak, peak_src = peak_hbm()
    fused = launches == Kt
    gbs = B * bytes_frame / ((ms / Kt) * 1e-3) / 1e9
    traffic, tsrc = traffic_for(name, "step_overlap")
    res["roofline"] = {"bound": "hbm",
                       "kernel": ("k_step_overlap<pred>" if cfg["pred"] else "k_step_overlap<gt>") if fused else "four-kernel step",
                       "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": traffic,
                       "traffic_source": tsrc, "peak_source": peak_src,
                       "alg_bytes_per_env_frame": bytes_frame, "alg_bytes_per_launch": B * bytes_frame,
                       "launch_us": 1e3 * ms / Kt, "p_local": p_local, "p_in": p_in,
                       "duration": "CUDA events round the timed region / launches in it (this rank)"}

    # ---- phase split inside the persistent kernel (%globaltimer stamps; steps run one at a time here, so the
    #      phases of consecutive steps do not overlap in this breakdown)
    if headline and fused:
        acc = np.zeros(4)
        facc = np.zeros(7)
        for _ in range(16):
            step(t); t += 1
            ns = np.asarray(mm.phase_ns()[:6], dtype=np.float64)
            acc += np.asarray([ns[1] - ns[0], ns[2] - ns[1], ns[5] - ns[2], ns[3] - ns[2]])
            tr = mm.fixup_trace_ns()
            facc += np.diff(np.asarray([ns[2]] + tr[:7], dtype=np.float64))
        ph = dict(zip(["start_to_barrier1", "barrier1_to_barrier2", "raster_beside_fixup", "fixup_cta0"], (acc / 16e3).tolist()))
        ph["fixup_split_cta0"] = dict(zip(["enter", "stage1", "bbox_segments", "stage1_flag", "team_scan", "stage2", "publish"],
                                          (facc / 16e3).tolist()))
        _, st2 = mm.status()
        ph["edge_entries"] = {"e1": int(st2[4]), "e2": int(st2[5]), "merged_total": int(st2[6]),
                              "scan_segments": int(st2[7]) >> 32, "scan_cells": int(st2[7]) & 0xFFFFFFFF}
        ph["note"] = "single steps (synchronised one by one): not the overlapped schedule of the timed region"
        res["roofline"]["phase_us_serial"] = ph

    # ---- end to end: per-env numpy observations -> batch_obs (pinned slabs, async H2D on a copy stream) ->
    #      MappingModule.forward -> both maps back to the host (one async D2H), every step inside the timed region
    if e2e:
        from ivlnce_b200.mapper import EpisodesInfo, Observations, RobotCurrentState
        from ivlnce_b200.staging import MapEgress, ObservationStager

        Ke = min(K, 48)
        depth_h = depth.cpu().numpy()
        skey = "semantic_scores" if cfg["pred"] else "semantic12"
        if cfg["pred"]:   # the class scores stand for RedNet's output: a whole-batch tensor in pinned host memory
            scores_pin = sem.cpu().pin_memory()
        else:
            sem_h = sem.cpu().numpy()

        def env_obs(tt):  # what the simulator workers hand over: one dict of numpy arrays per env (views, no copies)
            dd = depth_h[tt % nd]
            obs = [{"depth": dd[b].reshape(cfg["H"], cfg["W"], 1), "world_robot_pose": pose[tt, b],
                    "world_robot_orientation": orient[tt, b], "not_done_masks": masks[tt, b:b + 1], "env_name": names[b]}
                   for b in range(B)]
            if not cfg["pred"]:
                for b in range(B):
                    obs[b][skey] = sem_h[tt % RING, b].reshape(cfg["H"], cfg["W"], 1)
            return obs

        stager, egress = ObservationStager(depth=3, mutate_inputs=False), MapEgress(depth=2, copy=False)

        def run_e2e(n, t_start):
            ticket, prev = None, None
            for i in range(n):
                batch = stager.batch_obs(env_obs(t_start + i), dev,
                                         prebatched={skey: scores_pin[(t_start + i) % RING]} if cfg["pred"] else None)
                if ticket is not None:  # the previous step's maps, while this step's copies are in flight
                    egress.add_map_to_observations(prev, None, B, ticket)
                d = batch["depth"].permute(0, 3, 1, 2)
                if cfg["pred"]:
                    obs = Observations(None, d, batch[skey])
                else:
                    obs = Observations(batch[skey].permute(0, 3, 1, 2), d, None)
                o = batch["world_robot_orientation"]
                out = mm(EpisodesInfo(batch["not_done_masks"], batch["env_name"]), obs,
                         RobotCurrentState(batch["world_robot_pose"], o[:, 0], o[:, 1]))
                ticket = egress.start({"occupancy_map": out.occupancy, "semantic_map": out.semantic})
                prev = [dict() for _ in range(B)]
            egress.add_map_to_observations(prev, None, B, ticket)
            return t_start + n

        t = run_e2e(3, t)
        env.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        t = run_e2e(Ke, t)
        e1.record()
        env.barrier()
        wall_ms = 1e3 * (time.perf_counter() - w0)
        ems = env.max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
        res["e2e"] = {"value": total_envs * Ke / (ems * 1e-3), "unit": "env-frames/s",
                      "h2d_bytes_per_step": int(stager.h2d_bytes), "d2h_bytes_per_step": int(egress.d2h_bytes), "steps": Ke,
                      "ms_per_step": ems / Ke,
                      "note": "per-env numpy observations (depth, labels, pose, angles, masks) -> staging.batch_obs (pinned slabs, "
                              "one async H2D per sensor on a copy stream"
                              + ("; the 40 class-score planes, which stand for RedNet's device-side output, ride along as one "
                                 "pinned whole-batch tensor" if cfg["pred"] else "")
                              + ") -> MappingModule.forward -> staging.MapEgress (one async D2H of both maps into pinned "
                              "memory, per-env numpy views), every step inside the timed region; time = max(CUDA events, host wall)"}

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    if cpu_steps > 0 and rank == 0 and world == 1:
        r = cpu_reference_run(cfg, (depth.cpu(), sem.cpu()), pose, orient, masks, 1, cpu_steps, 40.0)
        res["cpu_baseline"] = {"value": r["value"], "unit": "env-frames/s", "cores": r["threads"], "kind": "port",
                               "sample": f"first {r['steps']} steps (after 1 warm-up step) of the same workload (same frames and poses, "
                                         f"{B} envs/step, world cloud growing from the reset at t=0) with oracle/torch_path.py (the "
                                         f"reference's eager torch op sequence, torch-only scatter_max stand-in), {r['ms_per_step']:.1f} ms/step"}
    del mm
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------ known-map workload
def run_known(env, name, cfg, K, Wm, cpu=True, seed=0):
    """BASELINE config 5: scene clouds loaded on reset (mapper.py:851-881), then every step = band filter + ego
    transform + raster of the cloud under the window."""
    from ivlnce_b200.mapper import EpisodesInfo, MapDimensions, Observations, RobotCurrentState, create_known_mapper
    from ivlnce_b200.synthetic import make_known_cloud

    dev, rank = env.dev, env.rank
    B, R = cfg["envs"], math.ceil(cfg["map_m"] / cfg["res"])
    extent = cfg["store"] * cfg["res"] / 2 - 0.2          # the cloud must fit the store window (1024 half-cells of 0.05 m)
    res = {"workload": f"{name}: {cfg['desc']}", "envs_per_gpu": B}
    tmp = tempfile.TemporaryDirectory()
    clouds = {}
    for s in range(cfg["scenes"]):
        xyz, sem = make_known_cloud(cfg["points"], extent, cfg["classes"], seed=seed + 31 * s)
        clouds[f"scene{s}"] = (xyz, sem)
        np.savez(os.path.join(tmp.name, f"scene{s}.npz"), xyz=xyz, semantics=sem)
    names = [f"scene{b % cfg['scenes']}" for b in range(B)]
    n_walk = 2 * K + Wm + 64
    pose, orient, masks = make_poses(cfg, n_walk, seed)
    pose_d, orient_d = torch.from_numpy(pose).to(dev), torch.from_numpy(orient).to(dev)
    masks_h = torch.from_numpy(masks)                       # host masks: the reset decision needs no device sync
    md = MapDimensions(cfg["map_m"], cfg["map_m"], cfg["res"])
    mm = create_known_mapper(dev, md, tmp.name, store_cells=cfg["store"], known_capacity=cfg["points"] + 1024, max_envs=B)

    # per-step views of the resident inputs, made once (the timed call is the module call with its three argument objects)
    masks_v = [masks_h[t].view(-1, 1) for t in range(n_walk)]
    pose_v, elev_v, head_v = list(pose_d.unbind(0)), list(orient_d[:, :, 0].unbind(0)), list(orient_d[:, :, 1].unbind(0))
    no_obs = Observations(None, None, None)

    def step(t):
        return mm(EpisodesInfo(masks_v[t], names), no_obs, RobotCurrentState(pose_v[t], elev_v[t], head_v[t]))

    t0 = time.perf_counter()
    out = step(0)
    torch.cuda.synchronize(dev)
    res["load_s"] = round(time.perf_counter() - t0, 2)
    # ---- parity (first envs; envs are independent in known mode: no de-dup, mapper.py:862-881)
    from oracle.oracle import OracleMapper

    nb, ns = 8, 2
    orc = OracleMapper(8, 8, math.pi / 2, cfg["map_m"], cfg["map_m"], cfg["res"], mode="known", known_clouds=clouds)
    occ_eq = sem_eq = True
    for t in range(ns):
        out = step(t)
        occ, smp = out.occupancy[:nb].cpu().numpy(), out.semantic[:nb].cpu().numpy()
        o_ref, s_ref = orc.step(masks[t, :nb], pose[t, :nb], orient[t, :nb], env_names=names[:nb])
        occ_eq = occ_eq and bool(np.array_equal(occ, o_ref))
        sem_eq = sem_eq and bool(np.array_equal(smp, s_ref))
    res["parity"] = {"steps_checked": ns, "envs_checked": nb, "occupancy_equal": occ_eq, "semantic_equal": sem_eq,
                     "against": "oracle/mapper_oracle.c in known mode on the first envs (envs are independent in known mode)"}
    del orc
    t = 1
    for _ in range(max(Wm, 3) + 20):
        step(t); t += 1
    torch.cuda.synchronize(dev)
    launches0 = mm.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.barrier()
    ev0.record()
    for _ in range(K):
        step(t); t += 1
    ev1.record()
    env.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = mm.kernel_launches() - launches0
    flags, stats = mm.status()
    assert flags == 0, f"map error flags {flags}"
    p_in = stats[3] / B
    bytes_frame = 2 * R * R + 16 * p_in
    peak, peak_src = peak_hbm()
    gbs = B * bytes_frame / ((ms / K) * 1e-3) / 1e9
    ktraffic, ktsrc = traffic_for(name, "raster_known")
    res.update({"steps": K, "ms_per_step": ms / K, "value": B * K / (ms * 1e-3), "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "k_raster_known", "achieved": gbs, "peak": peak, "unit": "GB/s",
                             "frac": gbs / peak, "traffic": ktraffic, "traffic_source": ktsrc, "peak_source": peak_src,
                             "alg_bytes_per_env_frame": bytes_frame,
                             "alg_bytes_per_launch": B * bytes_frame, "p_in": p_in,
                             "duration": "CUDA events round the timed region / steps"}})
    if cpu and rank == 0 and env.world == 1:
        from oracle.torch_path import TorchReferencePath

        nbc = 4
        threads = host_threads()
        torch.set_num_threads(threads)
        tp = TorchReferencePath(8, 8, math.pi / 2, cfg["map_m"], cfg["map_m"], cfg["res"], known_clouds=clouds)
        o = torch.from_numpy(orient)
        tp.step(torch.from_numpy(masks[0, :nbc]), torch.from_numpy(pose[0, :nbc]), o[0, :nbc, 0], o[0, :nbc, 1], env_names=names[:nbc])
        c0 = time.perf_counter()
        nsteps = 3
        for tt in range(1, 1 + nsteps):
            tp.step(torch.from_numpy(masks[tt, :nbc]), torch.from_numpy(pose[tt, :nbc]), o[tt, :nbc, 0], o[tt, :nbc, 1],
                    env_names=names[:nbc])
        dt = time.perf_counter() - c0
        res["cpu_baseline"] = {"value": nbc * nsteps / dt, "unit": "env-frames/s", "cores": threads, "kind": "port",
                               "sample": f"{nsteps} steady-state steps of {nbc} of the 64 envs (2 M-point clouds each) with "
                                         f"oracle/torch_path.py in known mode, {1e3 * dt / nsteps:.0f} ms/step"}
    del mm
    torch.cuda.empty_cache()
    tmp.cleanup()
    return res


# ------------------------------------------------------------------------------------------ main
def main():
    args = parse_args()
    env = Env(args)
    rank, world, local_rank = env.rank, env.world, env.local_rank
    cfg = dict(WORKLOADS[args.workload])
    if args.envs_per_gpu:
        cfg["envs"] = args.envs_per_gpu
    if args.store:
        cfg["store"] = args.store
    B = cfg["envs"]
    K, Wm = args.steps, max(args.warmup, 0)
    config = {"workload": f"{args.workload}: {cfg['desc']}", "envs_per_gpu": B, "depth": [cfg["H"], cfg["W"]],
              "classes": cfg["classes"], "cell_m": cfg["res"], "ego_map": [math.ceil(cfg["map_m"] / cfg["res"])] * 2,
              "partition": f"envs sharded by tour, {B}/GPU x {world}", "l2": f"{RING} rotating resident input frames"
              f" ({RING * B * cfg['H'] * cfg['W'] * (4 * cfg['classes'] if cfg['pred'] else 1) / 2**20:.0f} MiB/GPU > L2)"}

    # ---------------- reference arm: CPU only, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return 0
        wu = min(Wm, 3)
        total = wu + K + 1
        pose, orient, masks = make_poses(cfg, total, args.seed)
        depth, sem = make_frames(cfg, torch.device("cpu"), args.seed)
        r = cpu_reference_run(cfg, (depth, sem), pose, orient, masks, wu, K, args.cpu_budget_s)
        sample = (f"{r['steps']} timed steps of the {args.workload} workload ({B} envs/step, world cloud growing from a "
                  f"reset at t=0, {wu} untimed warm-up steps -- the CPU path needs no more, and each costs ~0.25 s) with "
                  f"oracle/torch_path.py = the reference's eager torch op sequence; scatter_max is the torch-only stand-in "
                  f"(torch-scatter not installable offline)")
        line = {"impl": "reference", "metric": "env_frames_per_sec", "value": r["value"], "unit": "env-frames/s",
                "n_gpus": args.gpus, "steps": r["steps"], "warmup": max(Wm, 3), "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "notes": {"warmup_steps_run": wu},
                "cpu_baseline": {"value": r["value"], "unit": "env-frames/s", "cores": r["threads"], "kind": "port",
                                 "sample": sample},
                "e2e": {"value": r["value"], "unit": "env-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ---------------- our arm
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    import __graft_entry__ as entry
    from ivlnce_b200.build import needs_build

    if needs_build():
        if local_rank == 0:
            entry.build()
    cores = pin_cores(env)
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist.barrier()
    env.dev = dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)

    sampler = None
    if rank == 0:
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        sampler = ClockSampler(range(local_world)).start()

    head = run_iterative(env, args.workload, cfg, K, Wm, headline=True, e2e=not args.skip_e2e,
                         cpu_steps=0 if args.skip_cpu else args.cpu_steps, seed=args.seed + 17 * rank)

    # ---- the other BASELINE configs
    configs = {}
    only = [x for x in args.only.split(",") if x]
    Ks = max(K, 60)

    def wanted(n):
        return (not args.skip_configs) and (not only or n in only)

    if world == 1:
        for n, kw in (("gt1", dict(cpu_steps=12, parity_steps=4)), ("gt32", dict(cpu_steps=3, parity_steps=3, e2e=True)),
                      ("pred16_scene", dict(cpu_steps=0, parity_steps=2)), ("tour", dict(cpu_steps=40, parity_steps=24))):
            if wanted(n):
                c = dict(WORKLOADS[n])
                if args.skip_cpu:
                    kw["cpu_steps"] = 0
                if args.skip_e2e:
                    kw["e2e"] = False
                try:
                    configs[n] = run_iterative(env, n, c, Ks, Wm, seed=args.seed + 101, **kw)
                except Exception as ex:  # a failing extra config must not take the headline line with it
                    configs[n] = {"workload": n, "error": f"{type(ex).__name__}: {ex}"}
        if wanted("known64"):
            try:
                configs["known64"] = run_known(env, "known64", dict(WORKLOADS["known64"]), Ks, Wm, cpu=not args.skip_cpu,
                                               seed=args.seed + 7)
            except Exception as ex:
                configs["known64"] = {"workload": "known64", "error": f"{type(ex).__name__}: {ex}"}
    if wanted("gt256"):
        # BASELINE config 4 at every N (N = 1 included: the 1-GPU point of the 256-env scaling curve)
        from ivlnce_b200.sharding import shard_range

        c = dict(WORKLOADS["gt256"])
        s0, s1 = shard_range(c["envs"], world, rank)
        total_envs = c["envs"]
        c["envs"] = s1 - s0
        try:
            r = run_iterative(env, "gt256", c, Ks, Wm, parity_steps=1 if world == 1 else 2, e2e=(world > 1 and not args.skip_e2e),
                              seed=args.seed + 211 + 17 * rank)
            r["partition"] = (f"{total_envs} envs in contiguous blocks by tour: {s1 - s0} per GPU x {world} "
                              f"(the total is fixed: strong scaling)")
            r["scaling"] = "strong"
        except Exception as ex:
            if world > 1:
                raise
            r = {"workload": "gt256", "error": f"{type(ex).__name__}: {ex}"}
        configs["gt256"] = r

    # ---- NCCL gather of metrics + maps (outside the timed regions; the only collectives on this path)
    if world > 1:
        from ivlnce_b200.sharding import gather_maps, gather_metrics, map_checksum

        mm = build_module(cfg, dev, B, args.variant, True)
        pose, orient, masks = make_poses(cfg, 4, args.seed + 17 * rank)
        depth, sem = make_frames(cfg, dev, args.seed + 17 * rank)
        out = call_module(mm, cfg, [f"s{b}" for b in range(B)], torch.from_numpy(masks[0]).to(dev),
                          torch.from_numpy(pose[0]).to(dev), torch.from_numpy(orient[0]).to(dev), depth[0], sem[0])
        allm = gather_maps(out.occupancy, world * B)
        met = gather_metrics(torch.tensor([float(B * K), head["ms_per_step"], float(map_checksum(out.semantic))],
                                          dtype=torch.float64, device=dev))
        assert allm.shape[0] == world * B and met.shape[0] == world
        del mm

    clocks = sampler.stop() if sampler is not None else None
    if rank == 0:
        line = {"metric": "env_frames_per_sec", "value": head["value"], "unit": "env-frames/s", "n_gpus": world, "steps": K,
                "warmup": max(Wm, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
                "gpu_launches": head["gpu_launches"], "roofline": head["roofline"], "parity": head.get("parity"),
                "notes": {"step": ("one persistent kernel per step (k_step_overlap), consecutive steps pipelined (ivm_set_pipelined)"
                                   if not args.no_pipeline else "one persistent kernel per step (k_step_overlap), steps not overlapped"),
                          "variant": args.variant, "warmup_steps_run": head["warmup_steps"],
                          "host_cores_of_rank0": len(cores) if cores else None}}
        if "e2e" in head:
            line["e2e"] = head["e2e"]
        if "cpu_baseline" in head:
            line["cpu_baseline"] = head["cpu_baseline"]
        if configs:
            line["configs"] = configs
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
