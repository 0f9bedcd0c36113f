#!/usr/bin/env python
"""bench.py -- frames/s of the TSDF fusion hot path (integrate + raycast per frame) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own TSDFGrid (oracle/_ref, CUDA;
                                                             # it has no CPU path), else the CPU oracle port

A step = one frame of the workload: TSDFGrid::Integrate of a 640x480 RGB + depth + ht/lt frame
followed by TSDFGrid::RayCast from the same pose at max_depth*2 (TSDFSystem::Render,
modules/tsdf_module.cc:45-49).  Workload = BASELINE.json configs[1]: TUM fr1-shaped synthetic
sequence (configs/TUM_RGBD_rgbd_1.yaml intrinsics, depth through uint16/5000, hand-held walk,
voxel 0.01 m, truncation 0.06 m, max_depth 4 m, seed 1).  W warm-up frames build the local map,
then exactly K further, distinct frames are timed.

value  : frames/s with the frames already resident in HBM (rsf_integrate_device +
         rsf_raycast_device), CUDA events around each step on the launching stream, L2 flushed
         (256 MiB memset) between steps outside the events.
e2e    : frames/s through the blocking host entry points rsf_integrate + rsf_raycast with pinned
         HOST buffers (H2D of the frame and D2H of rgba+normal+depth inside the timed region),
         wall clock bracketed by synchronisation.
roofline : the integrate kernel (HBM-bound voxel read-modify-write), algorithmic bytes
         / CUDA-event duration measured by the library's per-kernel events in the same timed region.
cpu_baseline : the CPU oracle (multithreaded C++ restatement of the reference algorithm; the
         reference has no CPU path of its own) on a bounded sample of the same sequence.

N > 1 (torchrun): independent sequences, one grid per GPU (BASELINE.json configs[3] pattern), no
data-path collective; weak scaling; value = all ranks' frames / max-over-ranks time.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL's own log: leave whatever the caller set; otherwise INFO / INIT (the communicator's "nranks" lines) to stderr, so
# stdout stays the one JSON line
if "NCCL_DEBUG" not in os.environ:
    os.environ["NCCL_DEBUG"] = "INFO"
    os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

METRIC = "frames/s integrate+raycast 640x480 semantic (TUM fr1-shaped)"
METRICS = {"tum": METRIC, "scannet": "frames/s integrate+raycast 640x480 semantic (ScanNet-shaped)",
           "scannet-hires": "frames/s integrate+raycast 1296x968 semantic (ScanNet-shaped)",
           "l515": "frames/s integrate+raycast 640x480 semantic (L515-shaped)"}
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--workload", choices=["tum", "scannet", "scannet-hires", "l515"], default="tum",
                    help="tum = BASELINE.json configs[1] (the headline line); the others are configs[2] / configs[3] shapes")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short 1296x968 ScanNet-shaped measurement")
    ap.add_argument("--no-sharded", action="store_true", help="skip the spatially sharded configs[4] measurement")
    ap.add_argument("--no-replicas", action="store_true", help="skip the configs[3] (L515 1024x768, one stream per GPU) measurement")
    ap.add_argument("--sharded-height", type=int, default=968)
    ap.add_argument("--sharded-width", type=int, default=1296)
    ap.add_argument("--cpu-sample-frames", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


WORKLOADS = {
    "tum": "configs[1] TUM fr1-shaped 640x480 integrate+raycast per frame",
    "scannet": "configs[2] ScanNet scene0-shaped 640x480 integrate+raycast per frame",
    "scannet-hires": "configs[2] ScanNet scene0-shaped 1296x968 integrate+raycast per frame",
    "l515": "configs[3] L515-shaped 640x480 integrate+raycast per frame",
}


def make_sequence(args, rank, n_frames):
    synth = importlib.import_module("ra-slam_b200.synth")
    wl = getattr(args, "workload", "tum")
    if wl == "scannet":
        return synth.config3(n_frames=max(n_frames, 200))
    if wl == "scannet-hires":
        return synth.config3(n_frames=max(n_frames, 200), hires=True)
    if wl == "l515":
        return synth.config4(seed=10 + rank, n_frames=max(n_frames, 200))
    # Headline line: every replica runs the SAME sequence (its own grid, buffers and stream; no collective).  With a
    # different scene per rank the max-over-ranks time is the heaviest scene's -- round 1's "flat 0.87 from N = 2" was
    # that, not launch latency -- and says nothing about scaling.  configs[3]'s eight different seeds are measured
    # separately (`replicas_config4`).
    return synth.config2(n_frames=max(n_frames, 300), H=args.height, W=args.width)


def render_frames(seq, n):
    """Pre-render n frames (CPU, numpy) -- outside every timed region."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
        return list(ex.map(seq.frame, range(n)))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.idx, self.proc = [], gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                mx = float(p[2])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(p[1]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "integrate_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


def cpu_oracle_fps(seq, frames, warm, sample, threads=0):
    """integrate+raycast frames/s of the CPU oracle on `sample` frames after `warm` map-building frames."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import OracleGrid
    rsf = importlib.import_module("ra-slam_b200")
    g = OracleGrid(seq.voxel_size, seq.truncation, threads)
    cam = rsf.CameraParams(seq.K, seq.H, seq.W)
    for i in range(warm):
        g.Integrate(*frames[i], seq.max_depth, seq.K, seq.poses[i])
    t0 = time.perf_counter()
    ti = tr = 0.0
    n = 0
    for i in range(warm, warm + sample):
        a = time.perf_counter()
        g.Integrate(*frames[i], seq.max_depth, seq.K, seq.poses[i])
        b = time.perf_counter()
        g.RayCast(seq.max_depth * 2, cam, seq.poses[i])
        c = time.perf_counter()
        ti += b - a
        tr += c - b
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, {"integrate_ms": 1e3 * ti / n, "raycast_ms": 1e3 * tr / n}


def reference_gpu_fps(seq, frames, W, K, pinned):
    """The reference's OWN TSDFGrid (oracle/_ref/libref_tsdf.so: unmodified utils/tsdf/*.cu built against
    oracle/ref_shim) through its public blocking API: Integrate(host images) + RayCast per frame, as
    TSDFSystem does (modules/tsdf_module.cc:45-49, 88-115).  Wall clock; both calls end with a stream sync."""
    import torch
    import ref_lib
    g = ref_lib.RefGrid(seq.voxel_size, seq.truncation)
    Kx = seq.K.as_array()
    hold = []
    ptrs = []
    for fr in frames:
        row = []
        for a in fr:
            t = torch.from_numpy(np.ascontiguousarray(a))
            if pinned:
                t = t.pin_memory()
            hold.append(t)
            row.append(t.data_ptr())
        ptrs.append(row)
    poses = [p.as_array() for p in seq.poses]
    rmax = seq.max_depth * 2
    for i in range(W):
        g.integrate_raw(*ptrs[i], seq.H, seq.W, seq.max_depth, Kx, poses[i])
        g.raycast_raw(rmax, seq.H, seq.W, Kx, poses[i])
    torch.cuda.synchronize()
    ti = tr = 0.0
    t0 = time.perf_counter()
    for i in range(W, W + K):
        a = time.perf_counter()
        g.integrate_raw(*ptrs[i], seq.H, seq.W, seq.max_depth, Kx, poses[i])
        b = time.perf_counter()
        g.raycast_raw(rmax, seq.H, seq.W, Kx, poses[i])
        c = time.perf_counter()
        ti += b - a
        tr += c - b
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n_active = g.NumActiveBlock()
    del g
    return K / dt, {"integrate_ms": 1e3 * ti / K, "raycast_ms": 1e3 * tr / K, "active_blocks": n_active}


def run_reference(args):
    """--impl reference.  The reference has NO CPU implementation of this path (it is CUDA-only), so its "own
    implementation" is its CUDA TSDFGrid: when oracle/_ref/libref_tsdf.so exists (built in the development
    container from the reference's unmodified sources, see oracle/Makefile) and a GPU is visible, that is what
    is timed, through the reference's blocking public API with host images.  Otherwise the multithreaded CPU
    oracle port is timed on the host cores."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_lib
    use_ref = ref_lib.available()
    if use_ref:
        try:
            import torch
            use_ref = torch.cuda.is_available()
        except Exception:
            use_ref = False
    cfg = {"workload": WORKLOADS[args.workload]}
    secondary = None
    steady = None
    if use_ref:
        seq = make_sequence(args, 0, W + K)
        frames = render_frames(seq, W + K)
        fps_pg, br_pg = reference_gpu_fps(seq, frames, W, K, pinned=False)
        fps_pin, br_pin = reference_gpu_fps(seq, frames, W, K, pinned=True)
        fps, br = (fps_pin, br_pin) if fps_pin >= fps_pg else (fps_pg, br_pg)
        P = seq.H * seq.W
        cpu = {"value": fps, "unit": UNIT, "cores": 1, "kind": "reference",
               "device": "B200: the reference's own CUDA kernels (it has no CPU path); 1 host thread drives them",
               "sample": f"{K} frames after {W} map-building frames; unmodified utils/tsdf/*.cu compiled for sm_100a "
                         "against oracle/ref_shim (Eigen/OpenCV/GL stand-ins), TSDFGrid::Integrate + RayCast, blocking; "
                         "value = the faster of pageable (stock cv::Mat) and pinned host images",
               "pageable_inputs_value": fps_pg, "pinned_inputs_value": fps_pin, **br}
        e2e = {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 15 * P, "d2h_bytes_per_step": 0}
        cfg["parallelism"] = "1 grid, reference CUDA build"
        if args.workload == "tum" and not args.no_secondary:
            synth = importlib.import_module("ra-slam_b200.synth")
            s2 = synth.config3(n_frames=200, hires=True)
            f2 = render_frames(s2, SEC_W + SEC_K)
            v_pg, _ = reference_gpu_fps(s2, f2, SEC_W, SEC_K, pinned=False)
            v_pin, b2 = reference_gpu_fps(s2, f2, SEC_W, SEC_K, pinned=True)
            secondary = {"workload": WORKLOADS["scannet-hires"], "steps": SEC_K, "warmup": SEC_W, "unit": UNIT,
                         "value": max(v_pg, v_pin), "pageable_inputs_value": v_pg, "pinned_inputs_value": v_pin, **b2}
        if args.workload == "tum" and not args.no_secondary and W < SS_W:
            # the same workload on a grown map (the driver's short runs time the first few frames of a sequence)
            sseq = make_sequence(args, 0, SS_W + SS_K)
            sfr = render_frames(sseq, SS_W + SS_K)
            s_pg, _ = reference_gpu_fps(sseq, sfr, SS_W, SS_K, pinned=False)
            s_pin, sb = reference_gpu_fps(sseq, sfr, SS_W, SS_K, pinned=True)
            steady = {"steps": SS_K, "warmup": SS_W, "unit": UNIT, "value": max(s_pg, s_pin), "pageable_inputs_value": s_pg,
                      "pinned_inputs_value": s_pin, **sb}
    else:
        # bounded sample: each step is one frame; cap the run to a few minutes of CPU work
        K = min(K, 60)
        W = min(W, 20)
        seq = make_sequence(args, 0, W + K)
        frames = render_frames(seq, W + K)
        cores = os.cpu_count() or 1
        fps, br = cpu_oracle_fps(seq, frames, W, K, cores)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{K} frames after {W} map-building frames, oracle/rsf_oracle.cpp, {cores} threads", **br}
        e2e = {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        cfg["parallelism"] = "host threads"
    cfg.update({"height": seq.H, "width": seq.W, "voxel_size": seq.voxel_size, "truncation": seq.truncation,
                "max_depth": seq.max_depth, "raycast_max_depth": seq.max_depth * 2})
    line = {
        "impl": "reference", "metric": METRICS[args.workload], "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": 0,
    }
    if secondary:
        line["scannet_1296x968"] = secondary
    if steady:
        line["steady_state_map"] = steady
    print(json.dumps(line), flush=True)


DEPTH = 3                 # frames in flight on the pipelined end-to-end path (the library has 4 staging slots)
SEC_K, SEC_W = 60, 12     # the short secondary measurement (BASELINE.json's metric also names 1296x968)
SS_K, SS_W = 50, 200      # steady-state map: the headline workload again, timed after 200 map-building frames
REP_K, REP_W = 40, 10     # the configs[3] replica measurement (L515 1024x768, one stream per GPU)
# configs/zed_native_l515.yaml:22-27 "extrinsics from l515 to zed"
L515_EXTRINSICS = [[0.9995, -0.0097, 0.0289, -37.172e-3], [0.009, 0.9997, 0.0247, -22.6717e-3],
                   [-0.0291, -0.0245, 0.9993, 2.4699e-3], [0, 0, 0, 1]]


def quick_native(rsf, torch, dev, local_rank, seq, frames, W, K, barrier=None, with_classes=True):
    """Short form of the two arms of run_native for a secondary workload: device-resident frames/s (CUDA
    events, L2 flushed between steps) and pipelined end-to-end frames/s with pinned host buffers."""
    H, Wd = seq.H, seq.W
    cam = rsf.CameraParams(seq.K, H, Wd)
    rmax = seq.max_depth * 2
    stream = torch.cuda.current_stream()
    grid = rsf.TSDFGrid(seq.voxel_size, seq.truncation, device=local_rank, stream=stream.cuda_stream)
    d_frames = [tuple(torch.from_numpy(a).to(dev) for a in fr) for fr in frames]
    out = {"rgba": torch.empty((H, Wd, 4), dtype=torch.uint8, device=dev),
           "normal": torch.empty((H, Wd, 4), dtype=torch.uint8, device=dev),
           "depth": torch.empty((H, Wd), dtype=torch.float32, device=dev)}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(W):
        grid.Integrate(*d_frames[i], seq.max_depth, seq.K, seq.poses[i])
        grid.RayCast(rmax, cam, seq.poses[i], out=out)
    grid.synchronize()
    tot0 = grid.totals()
    grid.profile_enable(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    if barrier:
        barrier()
    for k in range(K):
        flush.zero_()
        ev[k][0].record()
        grid.Integrate(*d_frames[W + k], seq.max_depth, seq.K, seq.poses[W + k])
        grid.RayCast(rmax, cam, seq.poses[W + k], out=out)
        ev[k][1].record()
    torch.cuda.synchronize()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    prof = grid.profile_read()
    tot1 = grid.totals()
    grid.close()
    del d_frames
    grid_e = rsf.TSDFGrid(seq.voxel_size, seq.truncation, device=local_rank)
    pins, h_frames = [], []
    for fr in frames:
        row = []
        for a in fr:
            pa = rsf.PinnedArray(a.shape, a.dtype)
            pa.array[...] = a
            pins.append(pa)
            row.append(pa.array)
        h_frames.append(tuple(row))
    h_outs = []
    for _ in range(DEPTH + 1):
        d = {}
        for kname, shape, dt in (("rgba", (H, Wd, 4), np.uint8), ("normal", (H, Wd, 4), np.uint8), ("depth", (H, Wd), np.float32)):
            pa = rsf.PinnedArray(shape, dt)
            pins.append(pa)
            d[kname] = pa.array
        h_outs.append(d)
    for i in range(W):
        grid_e.Integrate(*h_frames[i], seq.max_depth, seq.K, seq.poses[i])
        grid_e.RayCast(rmax, cam, seq.poses[i], outputs=(), out=h_outs[0])
    torch.cuda.synchronize()
    if barrier:
        barrier()
    t0 = time.perf_counter()
    tickets = []
    for k in range(K):
        i = W + k
        grid_e.IntegrateAsync(*h_frames[i], seq.max_depth, seq.K, seq.poses[i])
        tickets.append(grid_e.RayCastAsync(rmax, cam, seq.poses[i], h_outs[k % (DEPTH + 1)]))
        if len(tickets) > DEPTH:
            grid_e.wait(tickets.pop(0))
    for t in tickets:
        grid_e.wait(t)
    grid_e.synchronize()
    e2e_s = time.perf_counter() - t0
    grid_e.close()
    # (c) the same pipeline fed with the raw 16-bit images the reference's callers convert on the CPU every frame
    #     (rsf_integrate_u16_async: 9 instead of 15 input bytes per pixel, conversion in the staging kernel)
    e2e_u16 = None
    if seq.depth_factor:
        grid_u = rsf.TSDFGrid(seq.voxel_size, seq.truncation, device=local_rank)
        u_frames = []
        for rgb, depth, ht, lt in frames:
            row = []
            for a in (rgb, np.clip(np.rint(depth.astype(np.float64) * seq.depth_factor), 0, 65535).astype(np.uint16),
                      np.rint(ht.astype(np.float64) * 65535).astype(np.uint16), np.rint(lt.astype(np.float64) * 65535).astype(np.uint16)):
                pa = rsf.PinnedArray(a.shape, a.dtype)
                pa.array[...] = a
                pins.append(pa)
                row.append(pa.array)
            u_frames.append(tuple(row))
        for i in range(W):
            grid_u.IntegrateU16(*u_frames[i], seq.depth_factor, seq.max_depth, seq.K, seq.poses[i])
            grid_u.RayCast(rmax, cam, seq.poses[i], outputs=(), out=h_outs[0])
        torch.cuda.synchronize()
        if barrier:
            barrier()
        t0 = time.perf_counter()
        tickets = []
        for k in range(K):
            i = W + k
            grid_u.IntegrateU16(*u_frames[i], seq.depth_factor, seq.max_depth, seq.K, seq.poses[i], asynchronous=True)
            tickets.append(grid_u.RayCastAsync(rmax, cam, seq.poses[i], h_outs[k % (DEPTH + 1)]))
            if len(tickets) > DEPTH:
                grid_u.wait(tickets.pop(0))
        for t in tickets:
            grid_u.wait(t)
        grid_u.synchronize()
        e2e_u16 = {"value": K / (time.perf_counter() - t0), "unit": UNIT, "h2d_bytes_per_step": 9 * H * Wd,
                   "d2h_bytes_per_step": 12 * H * Wd, "api": "rsf_integrate_u16_async + rsf_raycast_async + rsf_wait"}
        grid_u.close()
    # (d) configs[2]: 20 per-class probability planes fused per frame (rsf_integrate_classes_device) + raycast, frames
    #     resident in HBM; then the export for evaluation (GatherValidSemantic / GatherValidClasses), timed separately
    classes20 = None
    try:
        if not with_classes:
            raise StopIteration
        synth = importlib.import_module("ra-slam_b200.synth")
        Kc, Kn, Wn = 20, min(K, 24), min(W, 8)
        grid_c = rsf.TSDFGrid(seq.voxel_size, seq.truncation, device=local_rank, stream=stream.cuda_stream, num_classes=Kc)
        c_frames = []
        for rgb, depth, ht, lt in frames[:Wn + Kn]:
            c_frames.append((torch.from_numpy(rgb).to(dev), torch.from_numpy(depth).to(dev),
                             torch.from_numpy(synth.class_planes(ht, rgb, depth, Kc)).to(dev)))
        for i in range(Wn):
            grid_c.IntegrateClasses(*c_frames[i], seq.max_depth, seq.K, seq.poses[i])
            grid_c.RayCast(rmax, cam, seq.poses[i], out=out)
        torch.cuda.synchronize()
        evc = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Kn)]
        for k in range(Kn):
            flush.zero_()
            evc[k][0].record()
            grid_c.IntegrateClasses(*c_frames[Wn + k], seq.max_depth, seq.K, seq.poses[Wn + k])
            grid_c.RayCast(rmax, cam, seq.poses[Wn + k], out=out)
            evc[k][1].record()
        torch.cuda.synchronize()
        c_ms = sum(a.elapsed_time(b) for a, b in evc)
        def timed_export(fn, reps=3):
            """[ms] of `reps` back-to-back calls (size query + chunked download through the persistent ring), and the rows"""
            ms, rows_ = [], None
            for _ in range(reps):
                t0_ = time.perf_counter()
                rows_ = fn()
                ms.append(1e3 * (time.perf_counter() - t0_))
            return ms, rows_

        def export_entry(fn):
            ms_page, rows_ = timed_export(lambda: fn(pinned=False))
            ms_pin, _ = timed_export(lambda: fn(pinned=True))
            nb = int(rows_.nbytes)
            return {"rows": int(len(rows_)), "bytes": nb, "ms": min(ms_page), "ms_calls": [round(x, 3) for x in ms_page],
                    "gbs": nb / (min(ms_page) * 1e-3) / 1e9,
                    "pinned_destination": {"ms": min(ms_pin), "ms_calls": [round(x, 3) for x in ms_pin],
                                           "gbs": nb / (min(ms_pin) * 1e-3) / 1e9},
                    "note": "pageable = fresh numpy array per call (host copy out of the pinned ring + first-touch page "
                            "faults); pinned = DMA straight into a reused pinned buffer; first call of each includes the "
                            "one-time ring / buffer allocation"}
        classes20 = {"classes": Kc, "steps": Kn, "value": Kn / (c_ms * 1e-3), "unit": UNIT, "ms_per_step": c_ms / Kn,
                     "class_input_bytes_per_step": 4 * Kc * H * Wd,
                     "export_gather_valid_semantic": export_entry(grid_c.GatherValidSemantic),
                     "export_gather_valid_classes": export_entry(grid_c.GatherValidClasses)}
        grid_c.close()
        del c_frames
    except StopIteration:
        classes20 = None
    except Exception as e:      # the extension must never take the headline measurement down with it
        classes20 = {"error": repr(e)}
    P = H * Wd
    upd = tot1["updated_voxels"] - tot0["updated_voxels"]
    vis = tot1["visible_blocks"] - tot0["visible_blocks"]
    int_ms, int_n = prof["integrate"]
    alg = 24.0 * upd + 12.0 * vis + 15.0 * P * int_n
    return {"classes20": classes20,"steps": K, "warmup": W, "unit": UNIT, "value": K / (dev_ms * 1e-3), "ms_per_step": dev_ms / K,
            "e2e": {"value": K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 15 * P, "d2h_bytes_per_step": 12 * P},
            "e2e_u16_inputs": e2e_u16,
            "kernels_ms_per_launch": {k: (v[0] / v[1] if v[1] else None) for k, v in prof.items()},
            "integrate_achieved_gbs": alg / (int_ms * 1e-3) / 1e9 if int_ms else None,
            "updated_voxels_per_launch": upd / max(int_n, 1), "visible_blocks_per_launch": vis / max(int_n, 1)}


def run_native(args):
    import torch
    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist = None
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    rsf = importlib.import_module("ra-slam_b200")
    K, W = args.steps, args.warmup
    n = W + K
    seq = make_sequence(args, rank, n)
    frames = render_frames(seq, n)
    H, Wd = seq.H, seq.W
    P = H * Wd
    cam = rsf.CameraParams(seq.K, H, Wd)
    rmax = seq.max_depth * 2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm (value) ------------------------------------------------
    stream = torch.cuda.current_stream()
    grid = rsf.TSDFGrid(seq.voxel_size, seq.truncation, device=local_rank, stream=stream.cuda_stream)
    d_frames = [tuple(torch.from_numpy(a).to(dev) for a in fr) for fr in frames]
    out = {"rgba": torch.empty((H, Wd, 4), dtype=torch.uint8, device=dev),
           "normal": torch.empty((H, Wd, 4), dtype=torch.uint8, device=dev),
           "depth": torch.empty((H, Wd), dtype=torch.float32, device=dev)}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_dev(i):
        grid.Integrate(*d_frames[i], seq.max_depth, seq.K, seq.poses[i])
        grid.RayCast(rmax, cam, seq.poses[i], out=out)

    for i in range(W):
        step_dev(i)
    grid.synchronize()
    tot0 = grid.totals()
    grid.profile_enable(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    # clocks are sampled for rank 0's GPU only: the JSON line carries one `clocks` object, and eight nvidia-smi pollers (each
    # query takes driver-wide locks) perturb the launch latency of every rank
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    t_wall0 = time.time()
    for k in range(K):
        flush.zero_()                     # L2 flush, outside the timed events
        ev[k][0].record()
        step_dev(W + k)
        ev[k][1].record()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    prof = grid.profile_read()
    grid.profile_enable(False)
    tot1 = grid.totals()
    stats_end = grid.stats()

    # the same K frames again, L2 left warm (informational)
    grid_w = rsf.TSDFGrid(seq.voxel_size, seq.truncation, device=local_rank, stream=stream.cuda_stream)
    for i in range(W):
        grid_w.Integrate(*d_frames[i], seq.max_depth, seq.K, seq.poses[i])
        grid_w.RayCast(rmax, cam, seq.poses[i], out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        grid_w.Integrate(*d_frames[W + k], seq.max_depth, seq.K, seq.poses[W + k])
        grid_w.RayCast(rmax, cam, seq.poses[W + k], out=out)
    e1.record()
    torch.cuda.synchronize()
    warm_ms = e0.elapsed_time(e1)
    grid_w.close()
    del d_frames

    # ---------------- end-to-end arm (e2e): blocking host API, pinned host buffers ----------------
    grid_e = rsf.TSDFGrid(seq.voxel_size, seq.truncation, device=local_rank)
    pins = []
    h_frames = []
    for fr in frames:
        row = []
        for a in fr:
            pa = rsf.PinnedArray(a.shape, a.dtype)
            pa.array[...] = a
            pins.append(pa)
            row.append(pa.array)
        h_frames.append(tuple(row))
    h_outs = []
    for _ in range(DEPTH + 1):      # rotating read-back targets
        d = {}
        for kname, shape, dt in (("rgba", (H, Wd, 4), np.uint8), ("normal", (H, Wd, 4), np.uint8), ("depth", (H, Wd), np.float32)):
            pa = rsf.PinnedArray(shape, dt)
            pins.append(pa)
            d[kname] = pa.array
        h_outs.append(d)

    # (a) pipelined: rsf_integrate_async + rsf_raycast_async, at most DEPTH frames in flight; every step's
    #     H2D and D2H happen inside the timed region and every step's result is waited for
    for i in range(W):
        grid_e.Integrate(*h_frames[i], seq.max_depth, seq.K, seq.poses[i])
        grid_e.RayCast(rmax, cam, seq.poses[i], outputs=(), out=h_outs[0])
    barrier()
    checksum = 0.0
    t0 = time.perf_counter()
    tickets = []
    for k in range(K):
        i = W + k
        grid_e.IntegrateAsync(*h_frames[i], seq.max_depth, seq.K, seq.poses[i])
        tickets.append(grid_e.RayCastAsync(rmax, cam, seq.poses[i], h_outs[k % (DEPTH + 1)]))
        if len(tickets) > DEPTH:               # every step's result is waited for, DEPTH frames later
            grid_e.wait(tickets.pop(0))
    for t in tickets:
        grid_e.wait(t)
    grid_e.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    checksum = float(h_outs[(K - 1) % (DEPTH + 1)]["depth"].sum())
    grid_e.close()

    # (b) blocking, reference-style calls (rsf_integrate + rsf_raycast return only when done)
    grid_b = rsf.TSDFGrid(seq.voxel_size, seq.truncation, device=local_rank)
    Kb = min(K, 100)
    for i in range(W):
        grid_b.Integrate(*h_frames[i], seq.max_depth, seq.K, seq.poses[i])
    grid_b.synchronize()
    t0 = time.perf_counter()
    for k in range(Kb):
        i = W + k
        grid_b.Integrate(*h_frames[i], seq.max_depth, seq.K, seq.poses[i])
        grid_b.RayCast(rmax, cam, seq.poses[i], outputs=(), out=h_outs[0])
    e2e_blocking_s = (time.perf_counter() - t0) / Kb
    grid_b.close()
    # (c) the same two blocking calls exactly as the reference's callers make them -- and as the reference arm of this
    #     benchmark times them: Integrate(host images), then RayCast into the grid's own device images for the GL bridge
    #     (voxel_tsdf.cu:885-902: no read-back to the host), then a stream synchronisation.  This is what
    #     rsfusion::TSDFGrid::Integrate + ::RayCast(GLImage8UC4*, GLImage8UC4*) of include/rsfusion/tsdf_grid.hpp execute.
    grid_r = rsf.TSDFGrid(seq.voxel_size, seq.truncation, device=local_rank)
    for i in range(W):
        grid_r.Integrate(*h_frames[i], seq.max_depth, seq.K, seq.poses[i])
    grid_r.synchronize()
    t0 = time.perf_counter()
    for k in range(Kb):
        i = W + k
        grid_r.Integrate(*h_frames[i], seq.max_depth, seq.K, seq.poses[i])
        grid_r.RayCastInternal(rmax, cam, seq.poses[i])
        grid_r.synchronize()
    e2e_dropin_s = (time.perf_counter() - t0) / Kb
    grid_r.close()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    # ---------------- max over ranks ---------------------------------------------------------------
    if dist is not None:
        t = torch.tensor([dev_ms, warm_ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, warm_ms, e2e_s = (float(x) for x in t.tolist())
    # ---------------- configs[4]: one large scene spatially sharded over the N GPUs (strong scaling) ----------------
    # every rank takes part; rank 0 gets the result.  N = 1 is the unsharded baseline of the same frames.
    sharded = None
    if not args.no_sharded:
        import bench_sharded
        try:
            sharded = bench_sharded.run(rank, local_rank, world, steps=min(max(K, 8), 24), warmup=12, H=args.sharded_height,
                                        W=args.sharded_width)
        except Exception as e:      # never take the headline line down
            sharded = {"error": repr(e)}
    # ---------------- configs[3] as written: one independent L515 stream per GPU (weak scaling) ----------------
    # 1024x768, the L515 intrinsics, depthmap_factor 4000, every pose composed with the camera extrinsics on the host
    # (configs/zed_native_l515.yaml:22-34, modules/tsdf_module.cc:29), seeds 10 + rank; all ranks run at the same time.
    replicas4 = None
    if not args.no_replicas:
        try:
            synth = importlib.import_module("ra-slam_b200.synth")
            s4 = synth.config4(seed=10 + rank, n_frames=REP_W + REP_K, hires=True)
            E = rsf.SE3.from_matrix(np.array(L515_EXTRINSICS, np.float32))
            s4.poses = [rsf.SE3(E.q, E.t) * p for p in s4.poses]
            f4 = [tuple(np.ascontiguousarray(t.cpu().numpy()) for t in
                        synth.render_torch(s4.scene, s4.K, s4.H, s4.W, s4.poses[i], dev, depth_factor=s4.depth_factor))
                  for i in range(REP_W + REP_K)]
            r4 = quick_native(rsf, torch, dev, local_rank, s4, f4, REP_W, REP_K, barrier=barrier, with_classes=False)
            P4 = s4.H * s4.W
            mine = torch.tensor([r4["value"], r4["e2e"]["value"], (r4["e2e_u16_inputs"] or {}).get("value", 0.0)],
                                dtype=torch.float64, device=dev)
            allr = [mine]
            if dist is not None:
                allr = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(allr, mine)
            rows = [[float(x) for x in t.tolist()] for t in allr]
            slow = [min(r[j] for r in rows) for j in range(3)]
            total = [sum(r[j] for r in rows) for j in range(3)]
            replicas4 = {"workload": "configs[3] L515-shaped 1024x768, extrinsics composed per pose, one independent sequence per GPU "
                                     "(seeds 10..), integrate+raycast per frame", "steps": REP_K, "warmup": REP_W, "unit": UNIT,
                         "scaling": "weak", "n_gpus": world,
                         "value": world * slow[0], "e2e": {"value": world * slow[1], "unit": UNIT, "h2d_bytes_per_step": 15 * P4,
                                                          "d2h_bytes_per_step": 12 * P4,
                                                          "pcie_gbs_per_rank": {"h2d": slow[1] * 15 * P4 / 1e9, "d2h": slow[1] * 12 * P4 / 1e9}},
                         "e2e_u16_inputs": {"value": world * slow[2], "unit": UNIT, "h2d_bytes_per_step": 9 * P4,
                                            "d2h_bytes_per_step": 12 * P4,
                                            "pcie_gbs_per_rank": {"h2d": slow[2] * 9 * P4 / 1e9, "d2h": slow[2] * 12 * P4 / 1e9}},
                         "per_rank_fps": {"columns": ["value", "e2e", "e2e_u16_inputs"], "rows": rows},
                         "sum_of_ranks": {"value": total[0], "e2e": total[1], "e2e_u16_inputs": total[2]},
                         "note": "aggregate = n_gpus x the slowest rank's frames/s (ranks start together behind a barrier; every rank "
                                 "has its own scene, so the slowest rank is also the heaviest scene); sum_of_ranks = what the box "
                                 "delivers, the streams being independent"}
            del f4
        except Exception as e:      # never take the headline line down
            replicas4 = {"error": repr(e)}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the integrate kernel ---------------------------------------------
    upd = tot1["updated_voxels"] - tot0["updated_voxels"]
    vis = tot1["visible_blocks"] - tot0["visible_blocks"]
    int_ms, int_n = prof["integrate"]
    # 24 B per updated voxel (12 read + 12 written: tsdf, rgbw, prob) + 12 B per visible
    # block descriptor + compulsory image reads (7 + 4K) B per pixel, K = 2 -> 15 B
    alg_bytes = 24.0 * upd + 12.0 * vis + 15.0 * P * int_n
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / (int_ms * 1e-3) / 1e9 if int_ms > 0 else 0.0
    kernels = {k: {"ms_per_launch": (v[0] / v[1] if v[1] else None), "launches": v[1],
                   "share_of_step": (v[0] / dev_ms if dev_ms else None)} for k, v in prof.items()}
    line = {
        "metric": METRICS[args.workload], "value": world * K / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "height": H, "width": Wd,
                   "voxel_size": seq.voxel_size, "truncation": seq.truncation, "max_depth": seq.max_depth,
                   "raycast_max_depth": rmax, "semantic_classes": 2,
                   "l2": "flushed: 256 MiB memset between steps, outside the timed events",
                   "parallelism": f"replicas x{world} (one grid per GPU, the same sequence on every rank, no collective)" if world > 1 else "1 grid"},
        "value_warm_l2": world * K / (warm_ms * 1e-3),
        "e2e": {"value": world * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 15 * P, "d2h_bytes_per_step": 12 * P,
                "ms_per_step": 1e3 * e2e_s / K,
                "api": "rsf_integrate_async + rsf_raycast_async + rsf_wait (3-stream pipeline, <= 3 frames in flight, "
                       "pinned host buffers; every frame's read-back is waited for)",
                "l2": "not flushed: a continuous stream, the map (~60 MB) stays L2-resident between frames while every "
                      "frame's inputs arrive from the host; `value` is the flushed (cold) figure, `value_warm_l2` the warm one",
                "blocking_value": world / e2e_blocking_s,
                "blocking_api": "rsf_integrate + rsf_raycast (three images read back to pinned host memory: 12 B per pixel D2H)",
                "blocking_like_reference_value": world / e2e_dropin_s,
                "blocking_like_reference_api": "rsf_integrate + rsf_raycast_internal_device + rsf_synchronize = TSDFGrid::Integrate + "
                                               "TSDFGrid::RayCast(GLImage8UC4*, GLImage8UC4*): host images in, render into the grid's device "
                                               "images, no read-back -- byte for byte what `--impl reference` times",
                "checksum": checksum},
        # per frame: begin_frame, allocate, depth_tiles, select_visible, integrate, tile_range, raycast (7 kernels of this library;
        # the per-kernel event pairs count allocate / select / integrate / raycast once each)
        "gpu_launches": int(sum(v[1] for v in prof.values()) + int_n + prof["select"][1] + prof["raycast"][1]),
        "roofline": {"kernel": "integrate_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                     "traffic_source": "profiles/integrate_traffic.json: DRAM bytes of ONE launch captured with ncu --set full in a separate run "
                                       "(2326 visible blocks, 22.09 MB algorithmic in that launch), not measured during this run",
                     # the kernel is bound by instruction issue before it is bound by HBM (profiles/r2_ncu_summary.md E):
                     # ~223 warp instructions per voxel (IEEE operations in the reference's order) against 24 algorithmic bytes
                     "issue_bound": {"warp_instructions_per_voxel": 223, "warp_instructions_per_visible_block": 894 * 4,
                                     "issue_peak_ginstr_s": 148 * 4 * 1.965,
                                     "floor_us_per_launch": 894 * 4 * (vis / max(int_n, 1)) / (148 * 4 * 1.965e3),
                                     "source": "ncu source view of the captured launch (8.36 M warp instructions / 2336 blocks)"},
                     "algorithmic_bytes_per_launch": alg_bytes / max(int_n, 1),
                     "updated_voxels_per_launch": upd / max(int_n, 1), "visible_blocks_per_launch": vis / max(int_n, 1),
                     "voxel_updates_per_s": upd / (int_ms * 1e-3) if int_ms else None},
        "kernels": kernels,
        "kernels_note": "allocate and select_visible run concurrently on two streams, so their two shares overlap",
        "clocks": clocks,
        "map": {"active_blocks": stats_end["active_blocks"], "high_water": stats_end["high_water"]},
    }
    line["sharded_config5"] = sharded
    line["replicas_config4"] = replicas4
    if world == 1 and args.workload == "tum" and not args.no_secondary:
        synth = importlib.import_module("ra-slam_b200.synth")
        s2 = synth.config3(n_frames=200, hires=True)
        sec = quick_native(rsf, torch, dev, local_rank, s2, render_frames(s2, SEC_W + SEC_K), SEC_W, SEC_K)
        sec["workload"] = WORKLOADS["scannet-hires"]
        peak_s, _ = measured_peak_gbs()
        sec["integrate_roofline_frac"] = (sec["integrate_achieved_gbs"] or 0.0) / peak_s
        line["scannet_1296x968"] = sec
        if W < SS_W:
            # the headline workload again on a grown map (200 map-building frames first): what a long session runs at
            sseq = make_sequence(args, rank, SS_W + SS_K)
            ss = quick_native(rsf, torch, dev, local_rank, sseq, render_frames(sseq, SS_W + SS_K), SS_W, SS_K, with_classes=False)
            ss.pop("classes20", None)
            ss["integrate_roofline_frac"] = (ss["integrate_achieved_gbs"] or 0.0) / peak_s
            line["steady_state_map"] = ss
    if not args.no_cpu_baseline:
        s = min(args.cpu_sample_frames, K)
        cores = os.cpu_count() or 1
        fps, br = cpu_oracle_fps(seq, frames, W, s, cores)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{s} frames after {W} map-building frames, oracle/rsf_oracle.cpp "
                                          f"(multithreaded restatement; the reference has no CPU path)", **br}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
