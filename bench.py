#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (SGBM frames/s and Gdisp-evals/s) on BASELINE.json's config 2:
StereoSGBM MODE_SGBM (5 paths), reference configs/sgbm.yml wiring with numDisp 64, 752x480, synthetic random-dot
stereograms with a disparity ramp.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A "step" is one pass of the hot path (mvsv_compute: prefilter -> BT cost -> box sums -> 5 path scans -> WTA/LR ->
median -> speckle) over one batch of B stereo pairs per GPU.  `value` times K steps with inputs resident in HBM
(CUDA events on the engine's stream); `e2e` times the same K steps through the host-facing C-ABI call with pinned
HOST buffers, host->device and device->host copies inside the timed region.  For N > 1 the script runs under
torchrun, one rank per GPU; frames are round-robined over ranks (no data-path collective) and the time is the
max over ranks.  `--impl reference` times the reference's own CPU implementation of the path (OpenCV's
StereoSGBM through cv2 4.13.0 -- the library call reference src/disparity.cpp:8 makes) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mvstereovision3_b200 import shard, synth  # noqa: E402

# BASELINE.json configs[1]; parameter wiring of reference src/disparity.cpp:83-95 on configs/sgbm.yml (numDisp -> 64)
CFG = dict(name="cfg2: StereoSGBM MODE_SGBM 5-path, configs/sgbm.yml (numDisp=64), 752x480",
           H=480, W=752,
           params=dict(minDisp=1, numDisp=64, blockSize=13, disp12MaxDiff=0, preFilterCap=0, uniquenessRatio=0,
                       speckleWindowSize=150, speckleRange=2, disparityMode=0, P1=0, P2=0))

# algorithmic HBM bytes per evaluated cell (x, y, d) of each aggregation-stage kernel in the current pipeline
# (int16 volumes; images, maps and per-pixel outputs are < 1 % and ignored) -- see DESIGN.md "Kernels"
KERNEL_BYTES_PER_CELL = {"sgbm_vsum": 2.0, "sgbm_h1": 6.0, "sgbm_td": 6.0, "sgbm_vdir": 6.0, "sgbm_h2_wta": 4.0}
PATH_BYTES_PER_CELL = 8.0      # MODE_SGBM aggregation roofline model (C w+r, S_h w+r)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic_per_cell():
    """dram__bytes_read.sum + dram__bytes_write.sum per evaluated cell of each kernel, from the committed
    `ncu --set full` capture of this workload (profiles/r01_ncu_traffic.json; null when absent)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as f:
            return json.load(f)["dram_bytes_per_cell"]
    except Exception:
        return {}


def bind_to_gpu_numa_node(local_rank):
    """Best effort: run this rank (and allocate its pinned host buffers) on the CPUs of the NUMA node its GPU hangs
    off, so that host<->device copies of different ranks do not cross sockets.  Returns a short description."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return "numa node unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return "numa node %d has no allowed cpu" % node
        os.sched_setaffinity(0, allowed)
        return "numa node %d (%d cpus)" % (node, len(allowed))
    except Exception as e:          # no sysfs, container restrictions, ...
        return "not bound (%s)" % type(e).__name__


def w1_of(W, p):
    maxD = p["minDisp"] + p["numDisp"]
    return (W + min(p["minDisp"], 0)) - max(maxD, 0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU reference: OpenCV StereoSGBM via cv2 (frame-parallel, one matcher per worker, cv2.setNumThreads(1))
# ---------------------------------------------------------------------------------------------------------------
def _cv2_or_none():
    try:
        import cv2
        return cv2
    except Exception:
        return None


def _make_cv_matcher(cv2, p):
    return cv2.StereoSGBM_create(minDisparity=p["minDisp"], numDisparities=p["numDisp"], blockSize=p["blockSize"],
                                 P1=p["P1"], P2=p["P2"], disp12MaxDiff=p["disp12MaxDiff"], preFilterCap=p["preFilterCap"],
                                 uniquenessRatio=p["uniquenessRatio"], speckleWindowSize=p["speckleWindowSize"],
                                 speckleRange=p["speckleRange"],
                                 mode=cv2.STEREO_SGBM_MODE_HH if p["disparityMode"] == 1 else cv2.STEREO_SGBM_MODE_SGBM)


class CpuReference:
    """Times the reference's CPU path.  kind = "reference": cv2's StereoSGBM (the OpenCV call the reference makes);
    kind = "port": the C oracle (only if cv2 cannot be imported)."""

    def __init__(self, cfg, frames):
        self.cfg, self.frames = cfg, frames
        self.cv2 = _cv2_or_none()
        self.cores = max(1, os.cpu_count() or 1)
        if self.cv2 is not None:
            self.cv2.setNumThreads(1)
            self.kind = "reference"
            from concurrent.futures import ThreadPoolExecutor
            self.pool = ThreadPoolExecutor(self.cores)
            self.matchers = [_make_cv_matcher(self.cv2, cfg["params"]) for _ in range(self.cores)]
        else:
            import oracle
            self.oracle = oracle
            self.kind = "port"
            self.cores = 1

    def run(self, frames_per_worker):
        """Every worker processes frames_per_worker frames; returns (seconds, frames, outputs of worker 0)."""
        if self.kind == "reference":
            def work(wi):
                outs = []
                for j in range(frames_per_worker):
                    l, r = self.frames[(wi + j * self.cores) % len(self.frames)]
                    outs.append(self.matchers[wi].compute(l, r))
                return outs
            t0 = time.perf_counter()
            res = list(self.pool.map(work, range(self.cores)))
            dt = time.perf_counter() - t0
            return dt, frames_per_worker * self.cores, res[0]
        p = dict(self.cfg["params"])
        p["mode"] = p.pop("disparityMode")
        t0 = time.perf_counter()
        outs = [self.oracle.sgbm(*self.frames[j % len(self.frames)], p) for j in range(frames_per_worker)]
        return time.perf_counter() - t0, frames_per_worker, outs

    def describe(self):
        if self.kind == "reference":
            return "cv2 %s StereoSGBM.compute, %d frame-parallel workers (cv2.setNumThreads(1) each)" % (
                self.cv2.__version__, self.cores)
        return "C oracle port (oracle/mvsv_oracle.c), single thread"


def run_reference_arm(args, rank):
    if rank != 0:
        return 0
    cfg = CFG
    p = cfg["params"]
    frames = [synth.stereogram(cfg["H"], cfg["W"], p["minDisp"], p["numDisp"], seed=s)[:2] for s in range(16)]
    ref = CpuReference(cfg, frames)
    fpw = 2 if ref.kind == "reference" else 1
    for _ in range(max(args.warmup, 1)):
        ref.run(1)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        dt, n, _ = ref.run(fpw)
        t_total += dt; n_total += n
    fps = n_total / t_total
    gd = cfg["W"] * cfg["H"] * p["numDisp"] * fps / 1e9
    sample = "%d steps x %d frames of the cfg-2 workload (%s)" % (args.steps, fpw * ref.cores, ref.describe())
    line = {"impl": "reference", "metric": "sgbm_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "gdisp_evals_per_s": gd,
            "config": {"workload": cfg["name"], "frames_per_step": fpw * ref.cores, "host_cores": ref.cores},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": ref.cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=148,
                    help="stereo pairs per GPU per step (148 = one per SM: every kernel's grid is a whole number of waves)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference_arm(args, rank)

    import torch
    from mvstereovision3_b200 import api
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single process, not bound"

    cfg = CFG
    p, H, W, B = cfg["params"], cfg["H"], cfg["W"], args.batch
    W1 = w1_of(W, p)
    cells_per_frame = W1 * H * p["numDisp"]

    # frames: global frame i -> rank i mod world (north_star); every rank holds B of them per step
    ids = shard.frames_for_rank(B * world, rank, world)
    hl, hr, hd = api.pinned((B, H, W), np.uint8), api.pinned((B, H, W), np.uint8), api.pinned((B, H, W), np.int16)
    uniq = min(B, 16)                 # distinct seeded pairs; the batch repeats them (content does not change the work)
    gen = [synth.stereogram(H, W, p["minDisp"], p["numDisp"], seed=ids[i])[:2] for i in range(uniq)]
    for i in range(B):
        hl.array[i], hr.array[i] = gen[i % uniq]
    dl = torch.from_numpy(hl.array).to(dev)
    dr = torch.from_numpy(hr.array).to(dev)

    eng = api.Engine(W, H, max_batch=B, device=local_rank)
    eng.set_sgbm_params(**p)
    stages = api.STAGE_SGBM

    def step_device():
        eng.compute_device(dl.data_ptr(), W, dr.data_ptr(), W, H * W, B, stages)

    def step_e2e():
        eng.compute(hl.array, hr.array, stages)
        eng.download(B, out={"disp": hd.array})

    def fence():
        eng.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- HBM-resident figure -------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    fence()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.profile_enable(True)
    l0 = eng.launch_count
    eng.timer_start()
    for _ in range(args.steps):
        step_device()
    ms = eng.timer_stop()
    launches = eng.launch_count - l0
    prof = eng.profile_read()
    eng.profile_enable(False)
    fence()
    clocks = sampler.stop() if rank == 0 else None
    ms_max, frames_total = shard.reduce_max_and_sum(dist, dev, ms, B * args.steps)
    fps = frames_total / (ms_max * 1e-3)

    # ---- end to end through the host-facing call -------------------------------------------------
    # Two engines (two streams, two sets of pinned host buffers) are software-pipelined: while one engine's
    # result is copied back and awaited, the other engine's H2D copies and kernels run.  mvsv_order_after keeps
    # the two engines' kernels in submission order (otherwise they share the GPU, finish together, and the copies
    # of both run with no kernel to overlap).  Every step still does
    # its own host->device copy of 2*B*H*W bytes and device->host copy of the B disparity maps.
    eng2 = api.Engine(W, H, max_batch=B, device=local_rank)
    eng2.set_sgbm_params(**p)
    hl2, hr2, hd2 = api.pinned((B, H, W), np.uint8), api.pinned((B, H, W), np.uint8), api.pinned((B, H, W), np.int16)
    hl2.array[:] = hl.array
    hr2.array[:] = hr.array
    lanes = [(eng, hl, hr, hd), (eng2, hl2, hr2, hd2)]

    def submit(k):
        e, a_, b_, _ = lanes[k & 1]
        if k > 0:
            e.order_after(lanes[(k - 1) & 1][0])        # kernels in submission order; copies overlap them
        e.compute(a_.array, b_.array, stages)           # async: pinned H2D (copy stream) + kernels (engine stream)

    def collect(k):
        e, _, _, d_ = lanes[k & 1]
        e.download(B, out={"disp": d_.array})           # D2H on the same stream, then stream sync

    def e2e_loop(n):
        submit(0)
        for k in range(1, n):
            submit(k)
            collect(k - 1)
        collect(n - 1)

    e2e_loop(4)
    eng2.sync()
    fence()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    eng2.sync()
    eng.sync()
    ms_e2e = (time.perf_counter() - t0) * 1e3          # host clock between two device synchronisations
    fence()
    ms_e2e_max, frames_e2e = shard.reduce_max_and_sum(dist, dev, ms_e2e, B * args.steps)
    fps_e2e = frames_e2e / (ms_e2e_max * 1e-3)
    gpu_disp = hd.array[:uniq].copy()
    assert np.array_equal(hd.array, hd2.array), "the two pipelined engines disagree"

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (live CUDA-event durations from the timed region) ----------------------
    peak, peak_src = hbm_peak()
    cells_per_launch = cells_per_frame * B
    kern = {k: {"ms_total": v[0], "launches": v[1], "ms_per_launch": v[0] / v[1]} for k, v in prof.items()}
    total_kernel_ms = sum(v["ms_total"] for v in kern.values())
    for k, v in kern.items():
        v["share"] = v["ms_total"] / total_kernel_ms if total_kernel_ms else 0.0
        if k in KERNEL_BYTES_PER_CELL:
            v["algorithmic_GBps"] = KERNEL_BYTES_PER_CELL[k] * cells_per_launch / (v["ms_per_launch"] * 1e-3) / 1e9
    dom = max(kern, key=lambda k: kern[k]["ms_total"])
    if dom in KERNEL_BYTES_PER_CELL:
        ach = kern[dom]["algorithmic_GBps"]
        tpc = measured_traffic_per_cell().get(dom)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": (tpc * cells_per_launch if tpc is not None else None), "peak_source": peak_src,
                    "bytes_per_cell": KERNEL_BYTES_PER_CELL[dom], "cells_per_launch": cells_per_launch,
                    "ms_per_launch": kern[dom]["ms_per_launch"], "share_of_step": kern[dom]["share"]}
    else:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                    "traffic": None, "peak_source": peak_src}
    # whole-path figure against the 8 B/cell aggregation model of SURVEY.md 8(d)
    path_ach = PATH_BYTES_PER_CELL * cells_per_frame * fps / world / 1e9
    roofline_path = {"model_bytes_per_cell": PATH_BYTES_PER_CELL, "achieved": path_ach, "peak": peak, "unit": "GB/s",
                     "frac": path_ach / peak, "note": "per GPU; whole step (all kernels) against the 8 B/cell model"}

    # ---- CPU baseline on this box's host cores (bounded sample) + parity gate on the frames it computed ---------
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(cfg, gen)
        ref.run(1)
        fpw = 4 if ref.kind == "reference" else 2
        dt, n, outs = ref.run(fpw)
        cpu = {"value": n / dt, "unit": "frames/s", "cores": ref.cores, "kind": ref.kind,
               "sample": "%d frames of the same workload, %s, %.1f s wall" % (n, ref.describe(), dt),
               "gdisp_evals_per_s": W * H * p["numDisp"] * (n / dt) / 1e9}
        ok, checked = True, 0
        for j, o in enumerate(outs):
            fi = (0 + j * ref.cores) % uniq
            ok &= bool(np.array_equal(o, gpu_disp[fi]))
            checked += 1
        parity = {"frames_checked": checked, "bit_exact_vs_cpu_reference": ok}

    line = {"metric": "sgbm_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "gdisp_evals_per_s": W * H * p["numDisp"] * fps / 1e9,
            "evaluated_gcells_per_s": cells_per_frame * fps / 1e9,
            "config": {"workload": cfg["name"], "batch_per_gpu": B, "frames_per_step": B * world,
                       "sharding": "frame i -> rank i mod N, no collective", "host_binding_rank0": numa,
                       "l2": "working set per step (3 int16 volumes x %d frames = %.1f GB) exceeds the 126 MB L2"
                             % (B, 3 * 2 * cells_per_frame * B / 1e9)},
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": 2 * B * H * W,
                    "d2h_bytes_per_step": 2 * B * H * W, "ms_per_step": ms_e2e_max / args.steps,
                    "pipeline": "2 engines, kernels in submission order (mvsv_order_after), copies on separate streams, pinned host buffers; host clock between device syncs",
                    "gdisp_evals_per_s": W * H * p["numDisp"] * fps_e2e / 1e9},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_path": roofline_path,
            "kernels": kern, "cpu_baseline": cpu, "parity": parity}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
