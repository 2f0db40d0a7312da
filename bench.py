#!/usr/bin/env python
"""bench.py -- KITTI-shaped depth-estimation throughput on B200 (BASELINE.json metric: frames/s and
feature-depths/s; fraction of the HBM roofline; the host-CPU reference beside it).

    python bench.py --gpus N --steps K --warmup W            # CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

A step is one pass of the hot path over this rank's block of the synthetic frame sequence
(BASELINE.json configs[1]: 10k KITTI-shaped frames, 120 000 points, 1241x376, 2000 features, yaml
parameters with the ground plane disabled), inputs resident in HBM. Frames are independent, so N
ranks each own a contiguous block (weak scaling, no data-path collective); the per-frame results are
gathered on rank 0 over NCCL once per step. `e2e` is the same metric through the C ABI's host-buffer entry
point (mld_process_frames_host): pinned host memory in, H2D + kernels + D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

N_POINTS, IMG_W, IMG_H, N_FEATURES = 120000, 1241, 376, 2000
ALGO_BYTES_PER_FRAME = 16 * N_POINTS + 4 * IMG_W * IMG_H + 28 * N_FEATURES  # 3 842 464 B
SEED = 20261017
WORKLOAD = "kitti"

# BASELINE.json configs: "kitti" = configs[1] (the headline, default), "road" = configs[2] (RANSAC ground
# plane per frame + road path), "dense" = configs[3] (128-beam sweep, 2048x1024 image, 20000 features)
WORKLOADS = {
    "kitti": dict(dense=False, road=False, frames=10000, features=2000, desc="BASELINE.json configs[1]: non-road features"),
    "road": dict(dense=False, road=True, frames=10000, features=2000,
                 desc="BASELINE.json configs[2]: RANSAC ground plane fitted per frame on the GPU + road-depth path"),
    "dense": dict(dense=True, road=False, frames=2000, features=20000,
                  desc="BASELINE.json configs[3]: 128-beam sweep (260096 pts), 2048x1024 image, 20000 features"),
}


def set_workload(name):
    global N_POINTS, IMG_W, IMG_H, N_FEATURES, ALGO_BYTES_PER_FRAME
    w = WORKLOADS[name]
    if w["dense"]:
        N_POINTS, IMG_W, IMG_H = 260096, 2048, 1024
    N_FEATURES = w["features"]
    ALGO_BYTES_PER_FRAME = 16 * N_POINTS + 4 * IMG_W * IMG_H + 28 * N_FEATURES
    return w


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()

    def summary(self, t0=None, t1=None):
        sm, smax, reasons = [], [], set()
        for t, line in self.rows:
            if t0 is not None and not (t0 - 0.03 <= t <= t1 + 0.03):
                continue
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference's DepthEstimator (the reference itself cannot be built
# here: Eigen/PCL/Ceres/OpenCV/catkin are absent), timed on the box's host cores.
# ------------------------------------------------------------------------------------------------
def cpu_reference_throughput(budget_s: float, frames_host=None, uv_host=None):
    """frames/s of the oracle on a bounded sample of the KITTI-shaped workload.

    Two ways of using the host cores are timed and the better one is reported:
      native : the reference's own parallelism -- one frame at a time, serial setInputCloud, OpenMP over
               features in CalculateDepth (DepthEstimator.cpp:455) on all cores;
      frames : one single-threaded estimator per core, each on its own frames (frames are independent)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import numpy as np
    import oracle_lib as O
    from mono_lidar_depth_b200 import synth

    cores = os.cpu_count() or 1
    wl = WORKLOADS[WORKLOAD]
    p = O.yaml_params()
    p.do_use_ransac_plane = 1 if wl["road"] else 0
    cam = synth.dense_camera() if wl["dense"] else synth.kitti_camera()
    cfg = synth.default_config(wl["dense"], road=bool(wl["road"]))

    def make():
        o = O.Oracle(p)
        o.initialize(IMG_W, IMG_H, cam.focal_length_, cam.principal_point_x_, cam.principal_point_y_, synth.KITTI_T_LIDAR_TO_CAM)
        return o

    nsample = 32 if not wl["dense"] else 8
    if frames_host is None:
        frames_host = [synth.points_host(cfg, SEED, f) for f in range(nsample)]
        uv_host = [synth.features_host(cfg, SEED, f, N_FEATURES) for f in range(nsample)]
    nsample = len(frames_host)

    def run_frame(o, i):
        # the reference's per-frame sequence: setInputCloud (+ RANSAC when the plane is not segmented) + CalculateDepth
        o.set_cloud(frames_host[i])
        plane = None
        if wl["road"]:
            rc, coeffs, inl, _ = O.ransac_plane(p, frames_host[i], SEED + i)
            plane = (coeffs, inl) if rc == 0 else None
        o.calculate_depth(uv_host[i], plane)

    # native: OpenMP inside the frame
    O.lib().orc_set_num_threads(cores)
    o = make()
    run_frame(o, 0)  # warm-up
    t0 = time.perf_counter()
    done = 0
    t_set = t_calc = 0.0
    while time.perf_counter() - t0 < budget_s * 0.4 or done < 8:
        i = done % nsample
        a = time.perf_counter()
        if wl["road"]:
            run_frame(o, i)
            b = c = time.perf_counter()
            t_set += b - a
        else:
            o.set_cloud(frames_host[i])
            b = time.perf_counter()
            o.calculate_depth(uv_host[i])
            c = time.perf_counter()
            t_set += b - a
            t_calc += c - b
        done += 1
    native = done / (time.perf_counter() - t0)
    native_detail = {"frames": done, "ms_set_input_cloud": 1e3 * t_set / done, "ms_calculate_depth": 1e3 * t_calc / done}

    # frames: one single-threaded estimator per core
    O.lib().orc_set_num_threads(1)
    counts = [0] * cores
    stop = time.perf_counter() + budget_s * 0.6

    def worker(w):
        ow = make()
        k = w
        while time.perf_counter() < stop:
            run_frame(ow, k % nsample)
            counts[w] += 1
            k += cores

    ts = [threading.Thread(target=worker, args=(w,)) for w in range(cores)]
    t1 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    frame_parallel = sum(counts) / (time.perf_counter() - t1)
    O.lib().orc_set_num_threads(cores)
    best = max(native, frame_parallel)
    # the reference's OWN sources (oracle/_ref: compiled against stand-in Eigen/PCL headers) on a short
    # sample, reported for information only: the stand-in linear algebra is not Eigen, so the (faster) port stays the baseline
    ref_sources = None
    try:
        import ref_lib as R

        if R.available():
            r = R.Reference(p)
            r.initialize(IMG_W, IMG_H, cam.focal_length_, cam.principal_point_x_, cam.principal_point_y_, synth.KITTI_T_LIDAR_TO_CAM)
            t2 = time.perf_counter()
            nref = 0
            while time.perf_counter() - t2 < min(2.0, 0.1 * budget_s) or nref < 2:
                r.set_cloud(frames_host[nref % nsample], None)  # with do_use_ransac_plane the reference fits its own RansacPlane
                r.has_plane = bool(wl["road"])
                r.calculate_depth(uv_host[nref % nsample])
                nref += 1
            ref_sources = nref / (time.perf_counter() - t2)
    except Exception as ex:  # the checker build is optional on the GPU box
        ref_sources = f"unavailable: {ex}"
    return {
        "value": best,
        "unit": "frames/s",
        "cores": cores,
        "kind": "port",
        "sample": (f"{done} frames native (OpenMP over features, {cores} threads: {native:.1f} frames/s, "
                   f"setInputCloud {native_detail['ms_set_input_cloud']:.2f} ms + CalculateDepth {native_detail['ms_calculate_depth']:.2f} ms) and "
                   f"{sum(counts)} frames frame-parallel ({cores} single-thread estimators: {frame_parallel:.1f} frames/s) of the same "
                   f"{WORKLOAD} workload, {nsample} distinct frames cycled; the better figure is reported"),
        "native_frames_per_s": native,
        "frame_parallel_frames_per_s": frame_parallel,
        "reference_sources_frames_per_s": ref_sources,
    }


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    budget = float(os.environ.get("MLD_BENCH_CPU_SECONDS", "20"))
    per = max(2.0, budget / (steps + warm))
    vals = []
    for s in range(steps + warm):
        r = cpu_reference_throughput(per)
        if s >= warm:
            vals.append(r)
    v = statistics.mean(x["value"] for x in vals)
    last = vals[-1]
    ngpu = max(1, args.gpus)
    frames_per_step = v * per
    line = {
        "impl": "reference",
        "metric": "frames_per_sec", "value": v, "unit": "frames/s", "n_gpus": ngpu, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 * per, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{WORKLOAD}: {WORKLOADS[WORKLOAD]['desc']} ({N_POINTS} pts, {IMG_W}x{IMG_H}, {N_FEATURES} features), yaml parameters; "
                               f"each step is a bounded {per:.1f} s sample (~{frames_per_step:.0f} frames) of the sequence on the host cores",
                   "points_per_frame": N_POINTS, "features_per_frame": N_FEATURES, "image": [IMG_W, IMG_H]},
        "feature_depths_per_sec": v * N_FEATURES,
        "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample", "reference_sources_frames_per_s")},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["cpu_baseline"]["value"] = v
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_gpu(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist

    from mono_lidar_depth_b200 import DepthEstimator, DepthEstimatorParameters, sharding, synth

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = WORKLOADS[WORKLOAD]
    frames_total = env_int("MLD_BENCH_FRAMES", wl["frames"]) * world  # weak scaling: the same block per GPU
    f0, nframes = sharding.frame_block(frames_total, world, rank)
    cfg = synth.default_config(wl["dense"], road=bool(wl["road"]))
    n = synth.points_per_frame(cfg)
    F = N_FEATURES
    assert n == N_POINTS
    use_road = bool(wl["road"])
    cam = synth.dense_camera() if wl["dense"] else synth.kitti_camera()

    est = DepthEstimator(device=local_rank)
    est.InitConfig(DepthEstimatorParameters.reference_yaml(do_use_ransac_plane=1 if use_road else 0))
    est.Initialize(cam, synth.KITTI_T_LIDAR_TO_CAM)

    pts = torch.empty((nframes, n, 4), dtype=torch.float32, device=dev)
    uv = torch.empty((nframes, F, 2), dtype=torch.float64, device=dev)
    # two result sets: with N > 1 the NCCL gather of step i overlaps the kernels of step i+1
    depths = [torch.empty((nframes, F), dtype=torch.float64, device=dev) for _ in range(2 if world > 1 else 1)]
    statuses = [torch.empty((nframes, F), dtype=torch.int32, device=dev) for _ in range(2 if world > 1 else 1)]
    depth, status = depths[0], statuses[0]
    coeffs = torch.zeros((nframes, 4), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    synth.points_device(est, cfg, SEED, f0, nframes, pts.data_ptr(), stream=stream)
    synth.features_device(est, cfg, SEED, f0, nframes, F, uv.data_ptr(), stream=stream)
    torch.cuda.synchronize()

    if world > 1:
        # the per-frame results are gathered on rank 0 (one ncclGather of 12 F bytes per frame); an
        # all-gather would move N times the bytes into every GPU's HBM for nothing
        per = -(-frames_total // world)
        g_depth = torch.empty((world * per, F), dtype=torch.float64, device=dev) if rank == 0 else None
        g_status = torch.empty((world * per, F), dtype=torch.int32, device=dev) if rank == 0 else None
        gl_depth = list(g_depth.split(per)) if rank == 0 else None
        gl_status = list(g_status.split(per)) if rank == 0 else None

    pending = [[], []]
    step_no = [0]

    def step():
        b = step_no[0] % len(depths)
        step_no[0] += 1
        for w in pending[b]:  # the gather that last read this result set must be done before it is overwritten
            w.wait()
        pending[b] = []
        est.processFramesDevice(pts.data_ptr(), n, n, 16, uv.data_ptr(), F, depths[b].data_ptr(), statuses[b].data_ptr(), nframes,
                                road=use_road, seed=SEED + f0, d_plane_coeffs_out=coeffs.data_ptr() if use_road else 0, stream=stream)
        if world > 1:  # gather the per-frame results (the only inter-GPU traffic of the path), asynchronously
            pending[b] = [dist.gather(depths[b], gl_depth, dst=0, async_op=True),
                          dist.gather(statuses[b], gl_status, dst=0, async_op=True)]

    def drain():
        for b in range(len(pending)):
            for w in pending[b]:
                w.wait()
            pending[b] = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warm = max(1, args.steps), max(3, args.warmup)
    for _ in range(warm):
        step()
    drain()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.3)
    launches0 = est.kernelLaunchCount()
    est.profileEnable(not os.environ.get("MLD_BENCH_NO_PROF"))  # event brackets per launch group (a few % of the step)
    est.profileRead()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tw0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step()
    drain()  # every gather has completed inside the timed region
    e1.record()
    barrier()
    tw1 = time.perf_counter()
    depth, status = depths[(step_no[0] - 1) % len(depths)], statuses[(step_no[0] - 1) % len(depths)]
    est.profileEnable(False)
    prof, prof_frames = est.profileRead()
    launches = est.kernelLaunchCount() - launches0
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    clocks.stop()
    clk = clocks.summary(tw0, tw1)
    ms_per_step = ms_total / steps
    value = frames_total / (ms_per_step * 1e-3)

    # ---- in-run parity spot check against the oracle (rank 0) ----
    parity = None
    cpu_base = None
    e2e = None
    if rank == 0 and not os.environ.get("MLD_BENCH_NO_PARITY"):  # (diagnostic kernel builds only)
        sys.path.insert(0, str(ROOT / "tests"))
        import oracle_lib as O
        import parity_util as PU

        p = O.yaml_params()
        p.do_use_ransac_plane = 1 if use_road else 0
        orc = O.Oracle(p)
        orc.initialize(IMG_W, IMG_H, cam.focal_length_, cam.principal_point_x_, cam.principal_point_y_, synth.KITTI_T_LIDAR_TO_CAM)
        checked = 0
        for i in sorted({0, nframes // 2, nframes - 1}):
            cloud_i = pts[i].cpu().numpy()
            orc.set_cloud(cloud_i)
            plane_i = None
            if use_road:  # the oracle's RANSAC with the same per-frame seed gives the inlier set; coefficients from the GPU
                rc_i, c_ref, inl_i, _ = O.ransac_plane(p, cloud_i, SEED + f0 + i)
                c_gpu = coeffs[i].cpu().numpy()
                assert rc_i == 0 and np.allclose(c_gpu, c_ref, rtol=1e-5, atol=1e-6), (c_gpu, c_ref)
                plane_i = (c_gpu, inl_i)
            d_ref, s_ref = orc.calculate_depth(uv[i].cpu().numpy(), plane_i)
            PU.assert_depth_status_equal(depth[i].cpu().numpy(), status[i].cpu().numpy(), d_ref, s_ref, f"bench frame {i}")
            checked += 1
        s_all = status.cpu().numpy()
        parity = {"frames_checked_vs_oracle": checked, "status_exact": True, "depth_rtol": PU.DEPTH_RTOL,
                  "success_fraction": float((s_all == 1).mean()), "success_road_fraction": float((s_all == 16).mean())}

    # ---- e2e: host buffers through mld_process_frames_host ----
    ne = min(nframes, env_int("MLD_BENCH_E2E_FRAMES", 256 if wl["dense"] else 1024))
    h_pts = torch.empty((ne, n, 4), dtype=torch.float32).pin_memory()
    h_uv = torch.empty((ne, F, 2), dtype=torch.float64).pin_memory()
    h_depth = torch.empty((ne, F), dtype=torch.float64).pin_memory()
    h_status = torch.empty((ne, F), dtype=torch.int32).pin_memory()
    h_pts.copy_(pts[:ne])
    h_uv.copy_(uv[:ne])
    torch.cuda.synchronize()

    def e2e_step():
        est.processFramesHostPtr(h_pts.data_ptr(), n, n, 16, h_uv.data_ptr(), F, h_depth.data_ptr(), h_status.data_ptr(), ne,
                                 road=use_road, seed=SEED + f0)

    for _ in range(2):
        e2e_step()
    barrier()
    te0 = time.perf_counter()
    e2e_steps = max(2, min(steps, 5))
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    te = time.perf_counter() - te0
    tt = torch.tensor([te], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    te = float(tt.item())
    if not os.environ.get("MLD_BENCH_NO_PARITY"):
        assert torch.equal(h_status, status[:ne].cpu()) and torch.equal(h_depth, depth[:ne].cpu()), "host pipeline != device path"
    e2e_value = world * ne * e2e_steps / te
    scale = nframes / ne  # bytes per full step of this rank's block
    e2e = {"value": e2e_value, "unit": "frames/s",
           "h2d_bytes_per_step": int((h_pts.numel() * 4 + h_uv.numel() * 8) * scale) * world,
           "d2h_bytes_per_step": int((h_depth.numel() * 8 + h_status.numel() * 4) * scale) * world,
           "frames_timed_per_rank": ne * e2e_steps, "api": "mld_process_frames_host (pinned host buffers, 3-slot H2D/compute/D2H pipeline)"}

    if rank == 0 and world == 1:
        budget = float(os.environ.get("MLD_BENCH_CPU_SECONDS", "20"))
        hp = [pts[i].cpu().numpy() for i in range(min(32, nframes))]
        hu = [uv[i].cpu().numpy() for i in range(min(32, nframes))]
        cpu_base = cpu_reference_throughput(budget, hp, hu)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        # device-resident sequences run K1 of chunk j and the gather of chunk j-1 as ONE launch
        pipelined = est.pipelineFrames()
        fused = 0 if pipelined else est.fusedChunkFrames()
        chunk = nframes if pipelined else (fused or est.chunkFrames())
        per_class = {}
        for name, (ms, ln) in prof.items():
            per_class[name] = {"ms_total": ms, "launches": ln, "avg_launch_ms": (ms / ln) if ln else None}
        # dominant kernel of the step and its algorithmic bytes per launch (DESIGN.md "roofline")
        # single kernels only: feature_depth is the sum of feature_gather + feature_solve + feature_rest (road kernels and the
        # overflow pass), listed for the share of the step but not a kernel of its own
        if pipelined:  # one persistent launch per sequence does all of it (mld_pipeline.cu)
            per_class = {"depth_pipeline": per_class["project_scatter"], "ransac": per_class.get("ransac")}
        elif fused:
            per_class["fused_project_gather"] = per_class.pop("project_scatter")
            per_class.pop("feature_gather", None)
        kernels = {k: v for k, v in per_class.items()
                   if k in ("depth_pipeline", "project_scatter", "fused_project_gather", "feature_gather", "feature_solve") and v and v["launches"] and v["ms_total"] > 0}
        per_class["note"] = ("durations are bracketed by CUDA events on the launching streams inside the timed region; launches of different "
                             "chunks overlap (front stream: fused K1 + gather launches; slot streams: solve + overflow pass), so a kernel's "
                             "duration includes time shared with other kernels; in the fused pipeline every 4th launch group is sampled")
        # Which single kernel dominates the step: the event brackets of kernels that run concurrently overlap (the solve of one
        # chunk is stretched by the fused launch of the next and vice versa), so the ranking comes from the serialised ncu launch
        # list of this same command when it is committed (profiles/traffic.json, headline workload), else from the brackets.
        traffic_file = ROOT / "profiles" / "traffic.json"
        tr = {}
        if traffic_file.exists() and WORKLOAD == "kitti" and not pipelined:
            try:
                tr = json.loads(traffic_file.read_text())
            except Exception:
                tr = {}
        shares = {k: v for k, v in tr.get("share_of_step_ncu", {}).items() if k in kernels}
        if shares:
            dom = max(shares, key=shares.get)
        else:
            dom = max(kernels, key=lambda k: kernels[k]["ms_total"]) if kernels else None
        roof = None
        if dom:
            frames_per_launch = prof_frames / per_class[dom]["launches"]
            # K1 owns the point stream and the pixel map (written once per frame in the reference's accounting; the
            # epoch-tagged map makes the actual clear traffic ~0), K2 the feature reads and the result writes
            # bytes the kernel has to move given the algorithm as built: the point stream once, the feature reads, the result
            # writes. SURVEY.md 8(d)'s 4 W H map term is NOT charged to a kernel: the epoch-tagged map is never rewritten
            # as a whole (only the cells of visible points are touched), so charging it would report more than the DRAM moved.
            per_frame_bytes = {"depth_pipeline": 16 * N_POINTS + 28 * N_FEATURES, "project_scatter": 16 * N_POINTS, "feature_gather": 16 * N_FEATURES,
                               "fused_project_gather": 16 * N_POINTS + 16 * N_FEATURES, "feature_solve": 12 * N_FEATURES}[dom]
            avg_s = per_class[dom]["avg_launch_ms"] * 1e-3
            achieved = per_frame_bytes * frames_per_launch / avg_s / 1e9
            sampled_ms = sum(v["ms_total"] for k, v in per_class.items()
                             if isinstance(v, dict) and k in ("map_clear", "depth_pipeline", "project_scatter", "fused_project_gather", "ransac", "feature_depth"))
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "kernel": dom, "peak_source": peak_src, "algorithmic_bytes_per_launch": per_frame_bytes * frames_per_launch,
                    "avg_launch_ms": per_class[dom]["avg_launch_ms"], "frames_per_launch": frames_per_launch,
                    "share_of_step": shares.get(dom) if shares else (per_class[dom]["ms_total"] / sampled_ms if sampled_ms else None),
                    "share_source": "ncu launch list (serialised), profiles/traffic.json" if shares else "event brackets (overlapping)",
                    "per_kernel": per_class,
                    "algorithmic_bytes_split": "per frame: project_scatter 16 N (point stream), feature_gather 16 F (feature reads), feature_solve "
                                               "12 F (result writes); fused_project_gather = project_scatter of one chunk + feature_gather of the "
                                               "previous one in one launch = 16 N + 16 F. SURVEY.md 8(d)'s 4 W H map term is not charged to a kernel "
                                               "(the epoch-tagged map is never rewritten as a whole); `path` reports both accountings",
                    "path": {"algorithmic_bytes_per_frame": ALGO_BYTES_PER_FRAME,
                             "achieved": ALGO_BYTES_PER_FRAME * (value / world) / 1e9, "frac": ALGO_BYTES_PER_FRAME * (value / world) / 1e9 / peak,
                             "frac_without_map_term": (16 * N_POINTS + 28 * N_FEATURES) * (value / world) / 1e9 / peak,
                             "note": "whole hot path per GPU: B * frames/s against the same peak, B = 16 N + 4 W H + 28 F (SURVEY.md 8(d)); "
                                     "frac_without_map_term leaves out the 4 W H map write that the epoch-tagged map performs without moving the bytes"}}
        if roof and tr:
            roof["traffic"] = tr.get(dom, {}).get("dram_bytes_per_launch")
            roof["traffic_source"] = tr.get("source")
            fpl = tr.get("frames_per_launch")
            if roof["traffic"] and fpl and abs(fpl - frames_per_launch) > 1:  # the capture used another launch size: scale per frame
                roof["traffic"] = int(roof["traffic"] / fpl * frames_per_launch)
        line = {
            "metric": "frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{WORKLOAD}: sequence of {frames_total // world} synthetic frames per GPU, batched; {wl['desc']}; "
                                   f"{N_POINTS} pts, {IMG_W}x{IMG_H}, {N_FEATURES} features, monolidar_fusion/parameters.yaml "
                                   "with do_use_depth_segmentation 0",
                       "frames_per_gpu": frames_total // world, "points_per_frame": N_POINTS, "features_per_frame": N_FEATURES,
                       "image": [IMG_W, IMG_H], "chunk_frames_per_launch": chunk,
                       "l2": f"inputs of one step ({nframes * n * 16 / 1e9:.1f} GB of points per GPU) are far larger than the 126 MB L2; no flush needed",
                       "parallelism": f"frames sharded in contiguous blocks over {world} GPU(s); NCCL gather of the results on rank 0 per step" if world > 1 else "single GPU"},
            "feature_depths_per_sec": value * N_FEATURES,
            "clocks": clk,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu_base,
            "parity": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = args.workload
    set_workload(WORKLOAD)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
