"""Shared helpers of the GPU parity tests: run the same inputs through the CUDA path (via the
reference-shaped DepthEstimator mirror -> C ABI) and through the CPU oracle, and compare.

Parity contract (BASELINE.json north_star): pixel indices, neighbour lists and status codes
bit-exact; depths within 1e-4 relative."""
import ctypes as C

import numpy as np

import oracle_lib as O
from mono_lidar_depth_b200 import CameraPinhole, DepthEstimator, DepthEstimatorParameters, GroundPlane
from mono_lidar_depth_b200._capi import MldParams

DEPTH_RTOL = 1e-4


def params_from_c(c: MldParams) -> DepthEstimatorParameters:
    p = DepthEstimatorParameters()
    C.memmove(C.byref(p.c_struct), C.byref(c), C.sizeof(MldParams))
    return p


def make_pair(c_params: MldParams, cam: CameraPinhole, T):
    """(GPU estimator, oracle) configured identically."""
    est = DepthEstimator()
    est.InitConfig(params_from_c(c_params))
    est.Initialize(cam, T)
    orc = O.Oracle(c_params)
    W, H = cam.getImageSize()
    orc.initialize(W, H, cam.focal_length_, cam.principal_point_x_, cam.principal_point_y_, T)
    return est, orc


def assert_depth_status_equal(d_gpu, s_gpu, d_ref, s_ref, what=""):
    s_gpu = np.asarray(s_gpu)
    s_ref = np.asarray(s_ref)
    bad = np.nonzero(s_gpu != s_ref)[0]
    assert len(bad) == 0, f"{what}: {len(bad)} status mismatches, first {bad[:5]} gpu {s_gpu[bad[:5]]} ref {s_ref[bad[:5]]}"
    both_nan = np.isnan(d_gpu) & np.isnan(d_ref)
    ok = both_nan | (np.abs(d_gpu - d_ref) <= DEPTH_RTOL * np.abs(d_ref))
    bad = np.nonzero(~ok)[0]
    assert len(bad) == 0, f"{what}: {len(bad)} depth mismatches, first {bad[:5]} gpu {d_gpu[bad[:5]]} ref {d_ref[bad[:5]]}"


def compare_frame(est, orc, cloud, uv, plane=None, what="", check_map=True, neighbor_samples=64):
    """plane: None or (coeffs[4], inlier raw indices)."""
    gp = GroundPlane(plane[0], plane[1]) if plane is not None else None
    est.setInputCloud(cloud, gp)
    orc.set_cloud(cloud)
    if check_map:
        m_gpu = est.getPixelMap()
        m_ref = orc.pixel_map_raw()
        assert np.array_equal(m_gpu, m_ref), f"{what}: pixel map differs in {(m_gpu != m_ref).sum()} cells"
    step = max(1, len(uv) // max(neighbor_samples, 1))
    for u, v in uv[::step][:neighbor_samples]:
        for sw, sh in ((1.0, 1.0), (2.0, 1.5)):
            a = est.getNeighbors(float(u), float(v), sw, sh)
            b = orc.neighbors(float(u), float(v), sw, sh)
            assert np.array_equal(a, b), f"{what}: neighbours of ({u},{v}) scale ({sw},{sh}): {a} vs {b}"
    d_gpu, s_gpu = est.CalculateDepth(uv, gp)
    d_ref, s_ref = orc.calculate_depth(uv, plane)
    assert_depth_status_equal(d_gpu, s_gpu, d_ref, s_ref, what)
    return d_gpu, s_gpu


def random_scene_cloud(rng, n, W, H, f, cx, cy, T, zmin=2.0, zmax=60.0, dense_patches=12):
    """A cloud that fills the image densely enough to exercise every branch: random points plus a few
    locally planar patches (several points per search window), expressed in the lidar frame."""
    Tm = np.vstack([np.asarray(T, np.float64)[:3], [0, 0, 0, 1]])
    Ti = np.linalg.inv(Tm)
    pts = []
    # planar patches: plane z = z0 + a*(x) + b*(y) in camera frame sampled on a pixel grid
    for _ in range(dense_patches):
        u0, v0 = rng.uniform(0, W - 40), rng.uniform(0, H - 30)
        z0 = rng.uniform(zmin, zmax)
        a, b = rng.uniform(-0.5, 0.5, 2)
        us, vs = np.meshgrid(u0 + np.arange(0, 40, rng.choice([1.3, 2.1, 3.7])), v0 + np.arange(0, 30, rng.choice([1.7, 2.9, 4.3])))
        us, vs = us.ravel(), vs.ravel()
        xn, yn = (us - cx) / f, (vs - cy) / f
        z = z0 / np.maximum(1e-3, (1 - a * xn - b * yn))
        z = z + rng.normal(0, 0.01, z.shape)
        pts.append(np.stack([xn * z, yn * z, z], 1))
    m = n - sum(len(p) for p in pts)
    if m > 0:
        us, vs = rng.uniform(-20, W + 20, m), rng.uniform(-20, H + 20, m)
        z = rng.uniform(-5.0, zmax, m)  # some behind the camera
        pts.append(np.stack([(us - cx) / f * z, (vs - cy) / f * z, z], 1))
    cam = np.concatenate(pts, 0)
    rng.shuffle(cam)
    lid = (Ti[:3, :3] @ cam.T).T + Ti[:3, 3]
    out = np.zeros((len(lid), 4), np.float32)
    out[:, :3] = lid.astype(np.float32)
    return out


VARIANTS = ["defaults", "no_hist", "no_trimax", "no_planar_check", "no_ortho", "adjust_mode", "absolute_local", "no_thresholds", "count_min3",
            "pca", "big_window", "hist_min0", "no_cut_behind", "cut_behind_only"]


def variant_params(variant) -> MldParams:
    """Parameter sets that switch every optional module of DepthEstimator::Initialize (DepthEstimator.cpp:46-124) on and off."""
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    if variant == "defaults":
        p = O.default_params()
        p.do_use_ransac_plane = 0
        p.viewray_plane_orthoganality_treshold = 0.05
    elif variant == "no_hist":
        p.do_use_histogram_segmentation = 0
    elif variant == "no_trimax":
        p.do_use_triangle_size_maximation = 0
    elif variant == "no_planar_check":
        p.do_check_triangleplanar_condition = 0
    elif variant == "no_ortho":
        p.viewray_plane_orthoganality_treshold = 0.0
    elif variant == "adjust_mode":
        p.treshold_depth_mode = 1
        p.treshold_depth_local_mode = 1
        p.treshold_depth_max = 20
        p.treshold_depth_min = 5
    elif variant == "absolute_local":
        p.treshold_depth_local_valuetype = 0
        p.treshold_depth_local_value = 0.05
    elif variant == "no_thresholds":
        p.treshold_depth_enabled = 0
        p.treshold_depth_local_enabled = 0
    elif variant == "count_min3":
        p.radiusSearch_count_min = 3
    elif variant == "pca":
        p.do_use_PCA = 1
        p.pca_treshold_2_1_rel_min = 0.5
    elif variant == "big_window":
        p.pixelarea_search_witdh = 14
        p.pixelarea_search_height = 17
    elif variant == "hist_min0":
        p.histogram_segmentation_min_pointcount = 0
    elif variant == "no_cut_behind":
        p.do_use_cut_behind_camera = 0
        p.treshold_depth_enabled = 0
        p.treshold_depth_local_enabled = 0
    elif variant == "cut_behind_only":
        p.treshold_depth_enabled = 0
        p.treshold_depth_local_enabled = 0
        p.do_check_triangleplanar_condition = 0
        p.viewray_plane_orthoganality_treshold = 0.0
    return p


def random_configuration(seed):
    """(params, (W, H, f, cx, cy), T[3x4], cloud, features, plane or None): a random camera, extrinsic, window size, set of
    thresholds and module switches on a random scene -- shared by the oracle-vs-reference and the GPU-vs-oracle tests."""
    from mono_lidar_depth_b200 import synth

    KITTI_T = synth.KITTI_T_LIDAR_TO_CAM
    rng = np.random.RandomState(1000 + seed)
    W, H = int(rng.randint(48, 420)), int(rng.randint(40, 300))
    f = float(rng.uniform(0.6, 2.5) * W)
    cx, cy = float(W * rng.uniform(0.3, 0.7)), float(H * rng.uniform(0.3, 0.7))
    # a random rigid transform: the KITTI axes permutation times a random rotation of up to ~20 degrees, random offset
    a = rng.normal(0, 0.2, 3)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    Rr = np.eye(3) + K + K @ K / 2
    U, _, Vt = np.linalg.svd(Rr)
    Rr = U @ Vt
    T = np.zeros((3, 4))
    T[:, :3] = Rr @ np.asarray(KITTI_T, np.float64)[:, :3]
    T[:, 3] = rng.uniform(-0.5, 0.5, 3)
    p = O.yaml_params()
    p.do_use_ransac_plane = int(rng.rand() < 0.5)
    p.pixelarea_search_witdh = int(rng.randint(1, 21))
    p.pixelarea_search_height = int(rng.randint(1, 21))
    p.radiusSearch_count_min = int(rng.randint(0, 5))
    p.do_use_histogram_segmentation = int(rng.rand() < 0.8)
    p.histogram_segmentation_bin_witdh = float(rng.choice([0.05, 0.1, 0.3, 0.7, 1.5]))
    p.histogram_segmentation_min_pointcount = int(rng.randint(0, 5))
    p.treshold_depth_enabled = int(rng.rand() < 0.8)
    p.treshold_depth_mode = int(rng.randint(0, 2))
    p.treshold_depth_min = int(rng.randint(0, 6))
    p.treshold_depth_max = int(rng.randint(10, 120))
    p.treshold_depth_local_enabled = int(rng.rand() < 0.8)
    p.treshold_depth_local_mode = int(rng.randint(0, 2))
    p.treshold_depth_local_valuetype = int(rng.randint(0, 2))
    p.treshold_depth_local_value = float(rng.choice([0.0, 0.05, 0.5, 2.0]))
    p.do_use_triangle_size_maximation = int(rng.rand() < 0.8)
    p.do_check_triangleplanar_condition = int(rng.rand() < 0.8)
    p.triangleplanar_crossnorm_treshold = float(rng.choice([0.0, 0.05, 0.1, 0.4]))
    p.viewray_plane_orthoganality_treshold = float(rng.choice([0.0, 0.03, 0.2, 0.6]))
    p.do_use_cut_behind_camera = int(rng.rand() < 0.7)
    road_mode = rng.randint(0, 2)
    p.plane_estimator_use_mestimator = int(road_mode == 0)
    p.plane_estimator_use_triangle_maximation = int(road_mode == 1)
    p.plane_estimator_z_x_min_relation = float(rng.choice([0.0, 0.2, 1.0]))
    p.ransac_plane_point_distance_treshold = float(rng.choice([0.05, 0.2, 1.0]))
    cloud = random_scene_cloud(rng, 4000, W, H, f, cx, cy, T, zmin=1.0, zmax=40.0, dense_patches=int(rng.randint(5, 40)))
    cloud[rng.randint(0, len(cloud), 40), rng.randint(0, 3, 40)] = np.nan
    uv = np.stack([rng.uniform(-8, W + 8, 700), rng.uniform(-8, H + 8, 700)], 1)
    uv[:200] = np.floor(uv[:200])
    plane = None
    if p.do_use_ransac_plane:
        # a plane through part of the cloud, in the lidar frame, with a random subset of the near points as inliers
        n3 = rng.normal(0, 1, 3)
        n3 /= np.linalg.norm(n3)
        fin = np.nonzero(np.isfinite(cloud[:, :3]).all(axis=1))[0]
        d0 = -float(np.median(cloud[fin, :3] @ n3))
        dist = np.abs(cloud[fin, :3].astype(np.float64) @ n3 + d0)
        near = fin[dist < np.percentile(dist, 40)]
        plane = (np.array([n3[0], n3[1], n3[2], d0], np.float32), np.sort(rng.choice(near, max(3, len(near) // 2), replace=False)).astype(np.int32))
    return p, (W, H, f, cx, cy), T, cloud, uv, plane
