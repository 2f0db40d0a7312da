#!/usr/bin/env python
"""bench.py -- KITTI-shaped depth-estimation throughput on B200 (BASELINE.json metric: frames/s and
feature-depths/s; fraction of the HBM roofline; the host-CPU reference beside it).

    python bench.py --gpus N --steps K --warmup W            # CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

A step is one pass of the hot path over this rank's block of the synthetic frame sequence
(BASELINE.json configs[1]: 10k KITTI-shaped frames, 120 000 points, 1241x376, 2000 features, yaml
parameters with the ground plane disabled), inputs resident in HBM. Frames are independent, so N
ranks each own a contiguous block (weak scaling, no data-path collective); the per-frame results are
gathered on rank 0 over NCCL once, at the end of the run, inside the timed region. `--workload seq100k` is
BASELINE.json configs[4]: ONE 100k-frame sequence cut into contiguous blocks over the ranks (strong scaling).
`e2e` is the same metric through the C ABI's host-buffer entry point (mld_process_frames_host): pinned host
memory holding 32-byte pcl::PointXYZI records in, H2D + kernels + D2H inside the timed region. At N = 1 the
line also carries short passes of the other BASELINE configs (`other_workloads`: road, dense, SemanticPlane),
an in-run parity check against the oracle and the host-CPU baseline.
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
    "seq100k": dict(dense=False, road=False, frames=10000, features=2000,
                    desc="BASELINE.json configs[4]: 100k-frame sequence sharded across the GPUs (strong scaling)"),
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


def bind_to_gpu_numa_node(local_rank):
    """Pins this process (and the library's pack threads, which inherit the mask) to the CPUs of the NUMA node the rank's GPU hangs
    off, so that the pinned host buffers of the end-to-end pipeline are allocated next to the GPU's root complex. Best effort: a
    flat or hidden topology leaves the mask alone. Returns what was done (reported in e2e.host)."""
    info = {"cores_visible": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()}
    try:
        import torch

        prop = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text().strip())
        info["gpu_pci"], info["numa_node"] = bus, node
        nodes = [p for p in Path("/sys/devices/system/node").glob("node[0-9]*")]
        info["numa_nodes"] = len(nodes)
        if node < 0 or len(nodes) < 2 or os.environ.get("MLD_BENCH_NO_NUMA"):
            return info
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["bound_cpus"] = len(cpus)
    except Exception as e:  # no sysfs, no permission, older torch: run unbound
        info["numa_note"] = f"{type(e).__name__}: {e}"[:120]
    return info


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
        # the CPU arm generates its inputs with libmld_synth.so and never maps the product library
        "product_library_mapped": any("libmld_cuda" in ln for ln in open("/proc/self/maps")),
    }
    line["cpu_baseline"]["value"] = v
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
SEMANTIC_ROAD_ROW = 230  # rows >= this carry the road label in the synthetic label image of the "semantic" pass


class Sequence:
    """A device-resident synthetic sequence of one workload on one rank: estimator, inputs, result buffers."""

    def __init__(self, name, frames, f0, local_rank, results=1, semantic=False):
        import torch

        from mono_lidar_depth_b200 import DepthEstimator, DepthEstimatorParameters, synth

        self.torch, self.synth = torch, synth
        self.name, self.wl, self.f0, self.nframes = name, WORKLOADS[name], f0, frames
        self.dense, self.road, self.semantic = bool(self.wl["dense"]), bool(self.wl["road"]), semantic
        self.dev = torch.device("cuda", local_rank)
        self.cfg = synth.default_config(self.dense, road=self.road)
        self.n = synth.points_per_frame(self.cfg)
        self.F = self.wl["features"]
        self.W, self.H = (2048, 1024) if self.dense else (1241, 376)
        self.cam = synth.dense_camera() if self.dense else synth.kitti_camera()
        self.est = DepthEstimator(device=local_rank)
        self.est.InitConfig(DepthEstimatorParameters.reference_yaml(do_use_ransac_plane=1 if self.road else 0))
        self.est.Initialize(self.cam, synth.KITTI_T_LIDAR_TO_CAM)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.pts = torch.empty((frames, self.n, 4), dtype=torch.float32, device=self.dev)
        self.uv = torch.empty((frames, self.F, 2), dtype=torch.float64, device=self.dev)
        self.depths = [torch.empty((frames, self.F), dtype=torch.float64, device=self.dev) for _ in range(results)]
        self.statuses = [torch.empty((frames, self.F), dtype=torch.int32, device=self.dev) for _ in range(results)]
        self.coeffs = torch.zeros((frames, 4), dtype=torch.float32, device=self.dev)
        self.labels = None
        if semantic:
            lab = torch.zeros((self.H, self.W), dtype=torch.uint8, device=self.dev)
            lab[SEMANTIC_ROAD_ROW:, :] = 7
            self.labels = lab.unsqueeze(0).repeat(frames, 1, 1).contiguous()
        self.generate(f0)

    def generate(self, f0, count=None):
        """(Re)generates frames [f0, f0 + count) of the sequence into the resident buffers (seed = frame index)."""
        count = self.nframes if count is None else count
        self.synth.points_device(self.est, self.cfg, SEED, f0, count, self.pts.data_ptr(), stream=self.stream)
        self.synth.features_device(self.est, self.cfg, SEED, f0, count, self.F, self.uv.data_ptr(), stream=self.stream)
        self.f0 = f0
        self.torch.cuda.synchronize()

    def step(self, b=0, count=None):
        count = self.nframes if count is None else count
        if self.semantic:
            from mono_lidar_depth_b200 import SemanticPlane

            cam = SemanticPlane.Camera(self.cam.focal_length_, self.cam.principal_point_x_, self.cam.principal_point_y_, self.synth.KITTI_T_LIDAR_TO_CAM)
            self.est.processFramesDeviceSemantic(self.pts.data_ptr(), self.n, self.n, 16, self.labels.data_ptr(), self.W, self.H, cam, [6, 7, 8, 9],
                                                 0.2, self.uv.data_ptr(), self.F, self.depths[b].data_ptr(), self.statuses[b].data_ptr(), count,
                                                 self.coeffs.data_ptr(), 0, self.stream)
        else:
            self.est.processFramesDevice(self.pts.data_ptr(), self.n, self.n, 16, self.uv.data_ptr(), self.F, self.depths[b].data_ptr(),
                                         self.statuses[b].data_ptr(), count, road=self.road, seed=SEED + self.f0,
                                         d_plane_coeffs_out=self.coeffs.data_ptr() if self.road else 0, stream=self.stream)

    def algorithmic_bytes_per_frame(self, with_map=True):
        return 16 * self.n + (4 * self.W * self.H if with_map else 0) + 28 * self.F

    def parity(self, frames, b=0):
        """Spot check against the oracle (status exact, depth within DEPTH_RTOL) + the status mix of the whole block."""
        import numpy as np

        sys.path.insert(0, str(ROOT / "tests"))
        import oracle_lib as O
        import parity_util as PU

        p = O.yaml_params()
        p.do_use_ransac_plane = 1 if self.road else 0
        orc = O.Oracle(p)
        orc.initialize(self.W, self.H, self.cam.focal_length_, self.cam.principal_point_x_, self.cam.principal_point_y_, self.synth.KITTI_T_LIDAR_TO_CAM)
        depth, status = self.depths[b], self.statuses[b]
        checked = 0
        for i in frames:
            cloud_i = self.pts[i].cpu().numpy()
            orc.set_cloud(cloud_i)
            plane_i = None
            if self.semantic:  # the plane the GPU fitted from the label image (single-frame API), handed to the oracle
                from mono_lidar_depth_b200 import SemanticPlane

                cam = SemanticPlane.Camera(self.cam.focal_length_, self.cam.principal_point_x_, self.cam.principal_point_y_, self.synth.KITTI_T_LIDAR_TO_CAM)
                sp = SemanticPlane(self.labels[i].cpu().numpy(), cam, [6, 7, 8, 9], 0.2, estimator=self.est)
                sp.CalculateInliersPlane(cloud_i)
                plane_i = (sp.getModelCoeffs(), sp.getInlinersIndex())
            elif self.road:  # the oracle's RANSAC with the same per-frame seed gives the inlier set; coefficients from the GPU
                rc_i, c_ref, inl_i, _ = O.ransac_plane(p, cloud_i, SEED + self.f0 + i)
                c_gpu = self.coeffs[i].cpu().numpy()
                assert rc_i == 0 and np.allclose(c_gpu, c_ref, rtol=1e-5, atol=1e-6), (c_gpu, c_ref)
                plane_i = (c_gpu, inl_i)
            d_ref, s_ref = orc.calculate_depth(self.uv[i].cpu().numpy(), plane_i)
            PU.assert_depth_status_equal(depth[i].cpu().numpy(), status[i].cpu().numpy(), d_ref, s_ref, f"bench {self.name} frame {i}")
            checked += 1
        s_all = status.cpu().numpy()
        hist = np.bincount(s_all.ravel(), minlength=21)
        return {"frames_checked_vs_oracle": checked, "status_exact": True, "depth_rtol": PU.DEPTH_RTOL,
                "success_fraction": float(hist[1] / s_all.size), "success_road_fraction": float(hist[16] / s_all.size),
                "insufficient_points_fraction": float(hist[2] / s_all.size), "no_local_max_fraction": float(hist[3] / s_all.size)}


def timed_steps(seq, steps, warm, barrier, after=None):
    """W untimed + K timed steps of one sequence; device time by CUDA events on the launching stream. `after` (optional) runs
    once after the last step INSIDE the timed region (the end-of-run gather of the results)."""
    torch = seq.torch
    for i in range(warm):
        seq.step(i % len(seq.depths))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        seq.step(i % len(seq.depths))
    if after is not None:
        after((steps - 1) % len(seq.depths))
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def kernel_profile(est, pipelined, fused):
    """Event-bracket classes of the timed region (mld_profile_read) -> per-kernel dict."""
    prof, prof_frames = est.profileRead()
    per = {name: {"ms_total": ms, "launches": ln, "avg_launch_ms": (ms / ln) if ln else None} for name, (ms, ln) in prof.items()}
    if pipelined:
        per = {"depth_pipeline": per["project_scatter"], "ransac": per.get("ransac")}
    elif fused:
        per["fused_project_gather"] = per.pop("project_scatter")
        per.pop("feature_gather", None)
    return per, prof_frames


def run_gpu(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist

    from mono_lidar_depth_b200 import sharding
    from mono_lidar_depth_b200.buildinfo import source_hash

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_info = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cores = os.cpu_count() or 1
    # host threads that pack PointXYZI records in the end-to-end pipeline: the box's cores are shared by the ranks
    os.environ.setdefault("MLD_PACK_THREADS", str(max(1, min(14, cores // world - (1 if world > 1 else 2)))))
    host_info["cores"], host_info["pack_threads"] = cores, int(os.environ["MLD_PACK_THREADS"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wl = WORKLOADS[WORKLOAD]
    seq100k = WORKLOAD == "seq100k"
    frames_per_gpu = env_int("MLD_BENCH_FRAMES", wl["frames"])
    if seq100k:  # strong scaling: a fixed 100k-frame sequence cut into contiguous blocks; each block streams through a resident window
        frames_total = env_int("MLD_BENCH_SEQ_FRAMES", 100000)
        f0, block = sharding.frame_block(frames_total, world, rank)
        window = min(block, frames_per_gpu)
        seq = Sequence(WORKLOAD, window, f0, local_rank)
    else:  # weak scaling: the same block per GPU
        frames_total = frames_per_gpu * world
        f0, block = sharding.frame_block(frames_total, world, rank)
        seq = Sequence(WORKLOAD, block, f0, local_rank)
    est, n, F, nframes = seq.est, seq.n, seq.F, seq.nframes

    # the per-frame results are gathered on rank 0 ONCE, at the end of the run (ncclSend/Recv under dist.gather),
    # inside the timed region. (Round 1 gathered every step: rank 0 ingested 1.7 GB per step while it computed and became the
    # straggler of the max-over-ranks time.)
    gather = None
    if world > 1 and not seq100k:
        per = -(-frames_total // world)
        # 9 bytes per feature on the wire: the depth stays f64, the DepthResultType (0..20) travels as one byte
        g_depth = torch.empty((world * per, F), dtype=torch.float64, device=dev) if rank == 0 else None
        g_status = torch.empty((world * per, F), dtype=torch.uint8, device=dev) if rank == 0 else None
        s8 = torch.empty((per, F), dtype=torch.uint8, device=dev)

        def gather(b):
            s8.copy_(seq.statuses[b])
            dist.gather(seq.depths[b], list(g_depth.split(per)) if rank == 0 else None, dst=0)
            dist.gather(s8, list(g_status.split(per)) if rank == 0 else None, dst=0)

    steps, warm = max(1, args.steps), max(3, args.warmup)
    clocks = ClockSampler(local_rank)
    if seq100k:
        # untimed: regenerate the window; timed: the hot path over it. The times of the windows add up to the block's time.
        for _ in range(warm):
            seq.step()
        barrier()
        clocks.start()
        time.sleep(0.3)
        launches0 = est.kernelLaunchCount()
        est.profileEnable(not os.environ.get("MLD_BENCH_NO_PROF"))
        est.profileRead()
        tw0 = time.perf_counter()
        ms_block = 0.0
        windows = 0
        for _ in range(steps):
            done = 0
            while done < block:
                cnt = min(nframes, block - done)
                if not (done == 0 and block <= nframes):
                    seq.generate(f0 + done, cnt)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                seq.step(0, cnt)
                e1.record()
                torch.cuda.synchronize()
                ms_block += e0.elapsed_time(e1)
                done += cnt
                windows += 1
        barrier()
        tw1 = time.perf_counter()
        ms_total = max_over_ranks(ms_block)
    else:
        for i in range(warm):
            seq.step()
        if gather is not None:  # warm-up of the collective too: NCCL sets its connections up on the first call
            gather(0)
        barrier()
        clocks.start()
        time.sleep(0.3)
        launches0 = est.kernelLaunchCount()
        est.profileEnable(not os.environ.get("MLD_BENCH_NO_PROF"))  # event brackets per launch group (a few % of the step)
        est.profileRead()
        tw0 = time.perf_counter()
        ms_rank = timed_steps(seq, steps, 0, barrier, gather)
        tw1 = time.perf_counter()
        ms_total = max_over_ranks(ms_rank)
    est.profileEnable(False)
    pipelined = False
    fused = est.fusedChunkFrames()
    per_class, prof_frames = kernel_profile(est, pipelined, fused)
    launches = est.kernelLaunchCount() - launches0
    clocks.stop()
    clk = clocks.summary(tw0, tw1)
    ms_per_step = ms_total / steps
    value = frames_total / (ms_per_step * 1e-3)
    per_rank_ms = None
    if world > 1 and not seq100k:
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = ms_rank / steps
        dist.all_reduce(t)
        per_rank_ms = [round(float(x), 4) for x in t.cpu()]

    # ---- in-run parity spot check against the oracle (rank 0) ----
    parity = None
    if rank == 0 and not os.environ.get("MLD_BENCH_NO_PARITY"):  # (diagnostic kernel builds only)
        last = min(nframes, block) - 1
        parity = seq.parity(sorted({0, last // 2, last}), b=(steps - 1) % len(seq.depths) if not seq100k else 0)

    # ---- e2e: host buffers through mld_process_frames_host, pcl::PointXYZI records (the drop-in caller's layout) ----
    ne = min(nframes, env_int("MLD_BENCH_E2E_FRAMES", 256 if wl["dense"] else 1024))
    h_pts = torch.zeros((ne, n, 8), dtype=torch.float32).pin_memory()  # x y z 1 | intensity pad pad pad
    h_pts[:, :, :3].copy_(seq.pts[:ne, :, :3])
    h_pts[:, :, 3] = 1.0
    h_pts[:, :, 4].copy_(seq.pts[:ne, :, 3])
    h_uv = torch.empty((ne, F, 2), dtype=torch.float64).pin_memory()
    h_depth = torch.empty((ne, F), dtype=torch.float64).pin_memory()
    h_status = torch.empty((ne, F), dtype=torch.int32).pin_memory()
    h_uv.copy_(seq.uv[:ne])
    torch.cuda.synchronize()

    def e2e_run(ptr, stride, reps):
        for _ in range(2):
            est.processFramesHostPtr(ptr, n, n, stride, h_uv.data_ptr(), F, h_depth.data_ptr(), h_status.data_ptr(), ne, road=seq.road, seed=SEED + seq.f0)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            est.processFramesHostPtr(ptr, n, n, stride, h_uv.data_ptr(), F, h_depth.data_ptr(), h_status.data_ptr(), ne, road=seq.road, seed=SEED + seq.f0)
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0)

    e2e_steps = max(2, min(steps, 5))
    hs0 = est.hostPipelineStats()
    te = e2e_run(h_pts.data_ptr(), 32, e2e_steps)
    hs1 = est.hostPipelineStats()
    hs = {k: hs1[k] - hs0[k] for k in hs0}
    frames_moved = max(1, hs["frames_packed"] + hs["frames_direct"])  # warm-up passes included
    if not os.environ.get("MLD_BENCH_NO_PARITY") and not seq100k:
        b = (steps - 1) % len(seq.depths)
        assert torch.equal(h_status, seq.statuses[b][:ne].cpu()) and torch.equal(h_depth, seq.depths[b][:ne].cpu()), "host pipeline != device path"
    step_frames = min(nframes, block) * world  # frames of one full step of the job
    e2e = {"value": world * ne * e2e_steps / te, "unit": "frames/s",
           # counted by the library from the copies it issued (mld_host_pipeline_stats), scaled to one step of the job
           "h2d_bytes_per_step": int(hs["h2d_bytes"] / frames_moved * step_frames),
           "d2h_bytes_per_step": int(hs["d2h_bytes"] / frames_moved * step_frames),
           "frames_timed_per_rank": ne * e2e_steps,
           "frames_packed_fraction": hs["frames_packed"] / frames_moved,
           "host": host_info,
           "input": "pinned host memory, pcl::PointXYZI records (32 bytes per point, the drop-in caller's cloud layout)",
           "api": (f"mld_process_frames_host: 3-slot H2D / kernels / D2H pipeline; {os.environ['MLD_PACK_THREADS']} host threads strip the records to "
                   "12-byte xyz in pinned staging buffers while the copy engine has work queued, chunks go out as whole records when it would idle")}
    if rank == 0 or world > 1:
        # the round-1 figure for comparison: 16-byte float4 points copied as they are
        h4 = torch.empty((ne, n, 4), dtype=torch.float32).pin_memory()
        h4.copy_(seq.pts[:ne])
        t4 = e2e_run(h4.data_ptr(), 16, 2)
        e2e["float4_input_frames_per_s"] = world * ne * 2 / t4
        del h4
    del h_pts

    cpu_base = None
    if rank == 0 and world == 1 and not seq100k:
        budget = float(os.environ.get("MLD_BENCH_CPU_SECONDS", "20"))
        hp = [seq.pts[i].cpu().numpy() for i in range(min(32, nframes))]
        hu = [seq.uv[i].cpu().numpy() for i in range(min(32, nframes))]
        cpu_base = cpu_reference_throughput(budget, hp, hu)

    # ---- the other BASELINE configs, short passes on rank 0 at N = 1 (configs[2] road, configs[3] dense, production SemanticPlane) ----
    others = None
    if rank == 0 and world == 1 and WORKLOAD == "kitti" and not os.environ.get("MLD_BENCH_NO_OTHERS"):
        peak, _ = measured_peak_gbs()
        others = {}
        for name, frames_o, semantic in (("road", env_int("MLD_BENCH_OTHER_FRAMES", 2048), False), ("dense", env_int("MLD_BENCH_OTHER_FRAMES", 2048) // 4, False),
                                         ("road", env_int("MLD_BENCH_OTHER_FRAMES", 2048), True)):
            del seq.pts, seq.uv  # free the headline's 20 GB first time round (harmless afterwards)
            seq.pts = seq.uv = None
            torch.cuda.empty_cache()
            so = Sequence(name, frames_o, 0, local_rank, semantic=semantic)
            ms = timed_steps(so, 3, 3, barrier) / 3
            v = frames_o / (ms * 1e-3)
            key = "semantic" if semantic else name
            others[key] = {"config": ("production caller: SemanticPlane fitted per frame from a label image on the GPU + road path (tracklet_depth_module.cpp:269-330), "
                                      "road / non-road feature mix" if semantic else WORKLOADS[name]["desc"]),
                           "frames": frames_o, "points_per_frame": so.n, "features_per_frame": so.F, "image": [so.W, so.H],
                           "value": v, "unit": "frames/s", "ms_per_step": ms, "feature_depths_per_sec": v * so.F,
                           "roofline_path": {"algorithmic_bytes_per_frame": so.algorithmic_bytes_per_frame(),
                                             "frac": so.algorithmic_bytes_per_frame() * v / 1e9 / peak,
                                             "frac_without_map_term": so.algorithmic_bytes_per_frame(False) * v / 1e9 / peak},
                           "parity": so.parity([0, frames_o - 1])}
            del so
            torch.cuda.empty_cache()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        chunk = nframes if pipelined else (fused or est.chunkFrames())
        kernels = {k: v for k, v in per_class.items()
                   if k in ("depth_pipeline", "project_scatter", "fused_project_gather", "feature_gather", "feature_solve") and v and v["launches"] and v["ms_total"] > 0}
        per_class["note"] = ("durations are bracketed by CUDA events on the launching streams inside the timed region; launches of different "
                             "chunks overlap (front stream: fused K1 + gather launches; slot streams: solve + overflow pass), so a kernel's "
                             "duration includes time shared with other kernels; in the fused pipeline every 4th launch group is sampled")
        # Which single kernel dominates the step: the event brackets of concurrent kernels overlap, so the ranking and the DRAM traffic
        # come from the serialised ncu launch list of this same command (profiles/traffic.json) -- only when that file was taken from
        # the sources that are running now (source_hash), else from the brackets and traffic stays null.
        tr = {}
        traffic_file = ROOT / "profiles" / "traffic.json"
        if traffic_file.exists() and WORKLOAD == "kitti":
            try:
                tr = json.loads(traffic_file.read_text())
                if tr.get("source_hash") != source_hash():
                    # a list taken from other sources still ranks the kernels better than overlapping brackets do; its byte counts are dropped
                    tr = {"stale": f"profiles/traffic.json was taken from sources {tr.get('source_hash')}, running {source_hash()}: "
                                   "kernel ranking taken from it, traffic not reported",
                          "share_of_step_ncu": tr.get("share_of_step_ncu", {})}
            except Exception:
                tr = {}
        shares = {k: v for k, v in tr.get("share_of_step_ncu", {}).items() if k in kernels}
        dom = max(shares, key=shares.get) if shares else (max(kernels, key=lambda k: kernels[k]["ms_total"]) if kernels else None)
        roof = None
        if dom:
            frames_per_launch = prof_frames / per_class[dom]["launches"]
            # bytes the kernel has to move given the algorithm as built: the point stream once, the feature reads, the result writes.
            # SURVEY.md 8(d)'s 4 W H map term is NOT charged to a kernel: the epoch-tagged map is never rewritten as a whole.
            per_frame_bytes = {"depth_pipeline": 16 * n + 28 * F, "project_scatter": 16 * n, "feature_gather": 16 * F,
                               "fused_project_gather": 16 * n + 16 * F, "feature_solve": 12 * F}[dom]
            avg_s = per_class[dom]["avg_launch_ms"] * 1e-3
            achieved = per_frame_bytes * frames_per_launch / avg_s / 1e9
            sampled_ms = sum(v["ms_total"] for k, v in per_class.items()
                             if isinstance(v, dict) and k in ("map_clear", "depth_pipeline", "project_scatter", "fused_project_gather", "ransac", "feature_depth"))
            per_gpu = value / world
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "kernel": dom, "peak_source": peak_src, "algorithmic_bytes_per_launch": per_frame_bytes * frames_per_launch,
                    "avg_launch_ms": per_class[dom]["avg_launch_ms"], "frames_per_launch": frames_per_launch,
                    "share_of_step": shares.get(dom) if shares else (per_class[dom]["ms_total"] / sampled_ms if sampled_ms else None),
                    "share_source": ("ncu launch list (serialised), profiles/traffic.json" + (" (of an older build)" if tr.get("stale") else "")) if shares
                                    else "event brackets (overlapping)",
                    "per_kernel": per_class,
                    "algorithmic_bytes_split": "per frame: project_scatter 16 N (point stream), feature_gather 16 F (feature reads), feature_solve "
                                               "12 F (result writes; its point re-reads are not algorithmic bytes); fused_project_gather = project_scatter of one chunk + feature_gather of the "
                                               "previous one in one launch = 16 N + 16 F. SURVEY.md 8(d)'s 4 W H map term is not charged to a kernel "
                                               "(the epoch-tagged map is never rewritten as a whole); `path` reports both accountings",
                    "path": {"algorithmic_bytes_per_frame": seq.algorithmic_bytes_per_frame(),
                             "achieved": seq.algorithmic_bytes_per_frame() * per_gpu / 1e9, "frac": seq.algorithmic_bytes_per_frame() * per_gpu / 1e9 / peak,
                             "frac_without_map_term": seq.algorithmic_bytes_per_frame(False) * per_gpu / 1e9 / peak,
                             "note": "whole hot path per GPU: B * frames/s against the same peak, B = 16 N + 4 W H + 28 F (SURVEY.md 8(d)); "
                                     "frac_without_map_term leaves out the 4 W H map write that the epoch-tagged map performs without moving the bytes"}}
            if tr.get("stale"):
                roof["traffic_note"] = tr["stale"]
            elif tr:
                roof["traffic"] = tr.get(dom, {}).get("dram_bytes_per_launch")
                roof["traffic_source"] = tr.get("source")
                roof["all_kernels_ncu"] = {k: tr[k] for k in tr.get("share_of_step_ncu", {}) if k in tr}
                fpl = tr.get("frames_per_launch")
                if roof["traffic"] and fpl and abs(fpl - frames_per_launch) > 1:  # the capture used another launch size: scale per frame
                    roof["traffic"] = int(roof["traffic"] / fpl * frames_per_launch)
        if seq100k:
            workload = (f"seq100k: BASELINE.json configs[4], ONE sequence of {frames_total} synthetic KITTI-shaped frames cut into contiguous blocks over "
                        f"{world} GPU(s) ({block} frames on rank 0); a block streams through a resident window of {nframes} frames that is regenerated on "
                        f"the device between passes (100k frames = 192 GB of points do not fit one GPU); the timed region is the hot path over every "
                        f"window ({windows // steps} per step on rank 0), the regeneration is not timed")
        else:
            workload = (f"{WORKLOAD}: sequence of {frames_total // world} synthetic frames per GPU, batched; {wl['desc']}; "
                        f"{n} pts, {seq.W}x{seq.H}, {F} features, monolidar_fusion/parameters.yaml with do_use_depth_segmentation 0; feature mix "
                        "calibrated to the reference's status log (include/mld_synth.h)")
        line = {
            "metric": "frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if seq100k else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload,
                       "frames_per_gpu": block, "points_per_frame": n, "features_per_frame": F,
                       "image": [seq.W, seq.H], "chunk_frames_per_launch": chunk,
                       "l2": f"inputs of one step ({nframes * n * 16 / 1e9:.1f} GB of points per GPU) are far larger than the 126 MB L2; no flush needed",
                       "parallelism": (f"frames sharded in contiguous blocks over {world} GPU(s), no collective on the data path; one gather of the per-frame "
                                       "results (f64 depth + one status byte per feature) on rank 0 at the end of the run (ncclSend/Recv under dist.gather), inside the timed region"
                                       if world > 1 and not seq100k else ("contiguous blocks, no collective" if world > 1 else "single GPU"))},
            "feature_depths_per_sec": value * F,
            "clocks": clk,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu_base,
            "parity": parity,
            "per_rank_ms_per_step": per_rank_ms,
            "other_workloads": others,
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
