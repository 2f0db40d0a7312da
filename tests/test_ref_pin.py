"""Pins the CPU oracle (oracle/mld_oracle.cpp) against the REFERENCE ITSELF: oracle/_ref/libmld_ref.so is built from
the reference's own monolidar_fusion sources (compiled where they lie under /root/reference) on stand-in Eigen/PCL
headers (oracle/ref_standin). Every comparison here is oracle vs reference on the same inputs; the CUDA path is then
compared with the oracle (tests/test_parity_gpu.py) and with reference-generated fixtures (tests/test_ref_golden.py).

Contract: visible indices, pixel map, neighbour lists and status codes bit-exact; depths bit-exact on the main path
(both sides evaluate the same FP64 expression trees) and within 1e-9 relative on the M-estimator road path (the two
one-sided Jacobi SVDs rotate in a different order).

These tests need the compiled reference and skip where it is absent (a box without /root/reference and without
the prebuilt oracle/_ref)."""
import collections

import numpy as np
import pytest

import kat_data
import oracle_lib as O
import parity_util as PU
import ref_lib as R
from mono_lidar_depth_b200 import synth

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libmld_ref.so not built (no /root/reference here)")

KT = synth.KITTI_T_LIDAR_TO_CAM
KCAM = (1241, 376, 718.856, 607.1928, 185.2157)


def pair(p, cam=KCAM, T=KT):
    o, r = O.Oracle(p), R.Reference(p)
    o.initialize(*cam, T)
    r.initialize(*cam, T)
    return o, r


def compare(o, r, cloud, uv, plane=None, what="", depth_rtol=0.0, neighbor_samples=48):
    o.set_cloud(cloud)
    r.set_cloud(cloud, plane)
    assert np.array_equal(o.point_index(), r.point_index()), f"{what}: _pointIndex differs"
    assert np.array_equal(o.image_points_visible(), r.image_points_visible()), f"{what}: visible image points differ"
    assert np.array_equal(o.points_camera(), r.points_camera(), equal_nan=True), f"{what}: camera-frame points differ"
    assert np.array_equal(o.pixel_map_visible(), r.pixel_map_visible()), f"{what}: pixel map differs"
    assert np.array_equal(o.pixel_map_raw(), r.pixel_map_raw())
    step = max(1, len(uv) // max(neighbor_samples, 1))
    for u, v in uv[::step][:neighbor_samples]:
        for sw, sh in ((1.0, 1.0), (2.0, 1.5)):
            assert np.array_equal(o.neighbors(float(u), float(v), sw, sh), r.neighbors(float(u), float(v), sw, sh)), (what, u, v)
    d_o, s_o = o.calculate_depth(uv, plane)
    d_r, s_r = r.calculate_depth(uv)
    bad = np.nonzero(s_o != s_r)[0]
    assert len(bad) == 0, f"{what}: {len(bad)} status mismatches, first {bad[:5]} oracle {s_o[bad[:5]]} reference {s_r[bad[:5]]}"
    if depth_rtol == 0.0:
        assert np.array_equal(d_o, d_r), f"{what}: depths not bit-identical, max rel {np.max(np.abs(d_o - d_r) / np.abs(d_r))}"
    else:
        assert np.all(np.abs(d_o - d_r) <= depth_rtol * np.abs(d_r)), f"{what}: max rel {np.max(np.abs(d_o - d_r) / np.abs(d_r))}"
    return d_r, s_r


@pytest.mark.skipif(not R.REF_YAML.exists(), reason="needs /root/reference/monolidar_fusion/parameters.yaml")
def test_yaml_loader_and_struct_defaults():
    """The reference's own yaml loader on its own parameters.yaml vs orc_yaml_params: identical except the three
    documented fields (shipped do_use_depth_segmentation 1 throws; ransac_plane_min_z/max_z are absent keys -> 0)."""
    a, b = O.yaml_params(), R.yaml_params()
    diff = {n: (getattr(a, n), getattr(b, n)) for n, _ in a._fields_ if getattr(a, n) != getattr(b, n)}
    assert diff == {"do_use_depth_segmentation": (0, 1), "ransac_plane_min_z": (-10000.0, 0.0), "ransac_plane_max_z": (10000.0, 0.0)}
    a, b = O.default_params(), R.default_params()
    assert {n for n, _ in a._fields_ if getattr(a, n) != getattr(b, n)} == set()
    assert b.viewray_plane_orthoganality_treshold == 1.0  # the octal literal {01} (DepthEstimatorParameters.h:155)


def test_kitti_shape_frames():
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    o, r = pair(p)
    cfg = synth.default_config()
    hist = collections.Counter()
    for frame in range(3):
        d, s = compare(o, r, synth.points_host(cfg, 11, frame), synth.features_host(cfg, 11, frame, 2000), what=f"frame {frame}")
        hist.update(s.tolist())
    assert hist[1] > 0 and hist[2] > 0 and hist[3] > 0


@pytest.mark.parametrize("seed", [0, 1])
def test_dense_random_scenes(seed):
    rng = np.random.RandomState(seed)
    W, H, f, cx, cy = 320, 240, 300.0, 160.3, 119.6
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    o, r = pair(p, (W, H, f, cx, cy))
    cloud = PU.random_scene_cloud(rng, 14000, W, H, f, cx, cy, KT, dense_patches=60)
    uv = np.stack([rng.uniform(-10, W + 10, 3000), rng.uniform(-10, H + 10, 3000)], 1)
    d, s = compare(o, r, cloud, uv, what=f"dense {seed}")
    hist = collections.Counter(s.tolist())
    assert hist[1] > 50 and len(hist) >= 5, hist


@pytest.mark.parametrize("variant", PU.VARIANTS)
def test_parameter_variants(variant):
    rng = np.random.RandomState(42)
    W, H, f, cx, cy = 256, 192, 250.0, 128.0, 96.0
    p = PU.variant_params(variant)
    o, r = pair(p, (W, H, f, cx, cy))
    cloud = PU.random_scene_cloud(rng, 10000, W, H, f, cx, cy, KT, dense_patches=45)
    uv = np.stack([rng.uniform(0, W, 2000), rng.uniform(0, H, 2000)], 1)
    # PCA: eigenvalues come from two different Jacobi sweeps; its float ratio tests can flip on a threshold
    if variant == "pca":
        o.set_cloud(cloud)
        r.set_cloud(cloud)
        d_o, s_o = o.calculate_depth(uv)
        d_r, s_r = r.calculate_depth(uv)
        same = s_o == s_r
        assert same.mean() > 0.999
        assert np.all(np.abs(d_o[same] - d_r[same]) <= 1e-9 * np.abs(d_r[same]))
    else:
        compare(o, r, cloud, uv, what=variant)


@pytest.mark.parametrize("mode", ["mestimator", "triangle"])
def test_road_path_with_injected_plane(mode):
    p = O.yaml_params()
    p.plane_estimator_use_mestimator = 1 if mode == "mestimator" else 0
    p.plane_estimator_use_triangle_maximation = 1 if mode == "triangle" else 0
    o, r = pair(p)
    cfg = synth.default_config()
    cloud = synth.points_host(cfg, 5, 1)
    rng = np.random.RandomState(1)
    coeffs = np.array([0.0, 0.0, 1.0, 1.73], np.float32)
    dist = np.abs(cloud[:, 2] + 1.73)
    inl = np.nonzero(np.isfinite(dist) & (dist < 0.1))[0].astype(np.int32)
    uv = np.stack([rng.randint(0, 1241, 3000), rng.randint(180, 376, 3000)], 1).astype(np.float64)
    d, s = compare(o, r, cloud, uv, plane=(coeffs, inl), what=mode, depth_rtol=1e-9 if mode == "mestimator" else 0.0)
    assert collections.Counter(s.tolist())[16] > 100


def test_road_path_sparse_inliers_tilted_plane():
    p = O.yaml_params()
    o, r = pair(p)
    cfg = synth.default_config()
    cloud = synth.points_host(cfg, 6, 2)
    coeffs = np.array([0.01, -0.02, 0.999, 1.70], np.float32)
    inl = np.nonzero(np.isfinite(cloud[:, 2]))[0][::7].astype(np.int32)
    compare(o, r, cloud, synth.features_host(cfg, 6, 2, 2000), plane=(coeffs, inl), what="sparse inliers", depth_rtol=1e-9)


def test_leastsquares_road_variant_is_not_part_of_the_compiled_reference():
    """R4: the Ceres fit is undefined behaviour upstream; oracle/_ref links a throwing stub in its place."""
    p = O.yaml_params()
    p.plane_estimator_use_mestimator = 0
    p.plane_estimator_use_leastsquares = 1
    r = R.Reference(p)
    r.initialize(*KCAM, KT)
    cloud = synth.points_host(synth.default_config(), 5, 1)
    dist = np.abs(cloud[:, 2] + 1.73)
    inl = np.nonzero(np.isfinite(dist) & (dist < 0.1))[0].astype(np.int32)
    r.set_cloud(cloud, (np.array([0, 0, 1, 1.73], np.float32), inl))
    with pytest.raises(RuntimeError, match="PlaneEstimationLeastSquares"):
        for u, v in np.stack([np.arange(300.0, 900.0, 10.0), np.full(60, 300.0)], 1):  # some road feature reaches the estimator
            r.calculate_depth_single(u, v)


def test_border_and_offimage_features():
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    o, r = pair(p)
    cloud = synth.points_host(synth.default_config(), 9, 0)
    feats = np.array([[-2.5, 200.0], [1243.0, 300.0], [620.0, -4.2], [620.0, 379.9], [0.0, 375.0], [1240.0, 170.0], [300.3, 250.7],
                      [-50.0, -50.0], [5000.0, 100.0], [620.0, -0.5], [1240.9, 375.9]])
    compare(o, r, cloud, feats, what="border features", neighbor_samples=len(feats))
    nan_cloud = np.full((1000, 4), np.nan, np.float32)
    compare(o, r, nan_cloud, feats[:7], what="nan cloud", neighbor_samples=7)


def test_status_precedence_region_growing_and_set_all_depths():
    cloud = synth.points_host(synth.default_config(), 1, 0)
    uv = synth.features_host(synth.default_config(), 1, 0, 64)
    p = O.yaml_params()
    p.do_use_depth_segmentation = 1  # the shipped yaml value
    p.do_use_ransac_plane = 0
    r = R.Reference(p)
    r.initialize(*KCAM, KT)
    r.set_cloud(cloud)
    with pytest.raises(RuntimeError, match="Region growing not supported"):
        for u, v in uv:  # the first feature with a non-empty window throws (DepthEstimator.cpp:605-608)
            r.calculate_depth_single(u, v)
    q = O.yaml_params()
    q.set_all_depths_to_zero = 1
    q.do_use_ransac_plane = 0
    o, r = pair(q)
    d, s = compare(o, r, cloud, uv, what="set_all_depths_to_zero")
    assert np.all(s == 1) and np.all(d == -1)


def test_reference_throw_sites():
    p = O.yaml_params()
    r = R.Reference(p)
    with pytest.raises(RuntimeError, match="without 'initialize'"):
        r.set_cloud(np.zeros((3, 4), np.float32))
    r.initialize(*KCAM, KT)
    with pytest.raises(RuntimeError, match="without 'SetInputCloud'"):
        r.calculate_depth(np.zeros((1, 2)))
    with pytest.raises(RuntimeError, match="Input pointcloud is invalid"):
        r.set_cloud(np.zeros((2, 4), np.float32))  # < 3 points with RANSAC requested (RansacPlane.cpp:44-50)
    q = O.yaml_params()
    q.neighbor_search_mode = 2
    r2 = R.Reference(q)
    with pytest.raises(RuntimeError, match="neighbor_search_mode has the invalid value"):
        r2.initialize(*KCAM, KT)


def test_histogram_golden_vector_and_random_vectors():
    """Histogram.FilterPointsMinDistBlob (test_monolidar_fusion.cpp:306-374) executed by the reference's own code."""
    ok, pos, lo, hi = R.histogram_filter(kat_data.HIST_DEPTHS, kat_data.HIST_BIN_WIDTH, kat_data.HIST_MIN_COUNT)
    assert ok and [kat_data.HIST_DEPTHS[i] for i in pos] == kat_data.HIST_EXPECTED
    rng = np.random.RandomState(3)
    for trial in range(300):
        n = rng.randint(1, 40)
        d = np.concatenate([rng.normal(rng.uniform(1, 60), rng.uniform(0.01, 1.0), n), rng.uniform(0.5, 80, rng.randint(0, 6))])
        if trial % 7 == 0:
            d[rng.randint(len(d))] = 999.0
        bw, mc = rng.choice([0.1, 0.3, 0.5, 1.0]), int(rng.choice([0, 1, 3, 5]))
        a, b = O.histogram_filter(d, bw, mc), R.histogram_filter(d, bw, mc)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]), (trial, d, bw, mc)
        if a[0]:
            assert a[2] == b[2] and a[3] == b[3]


def test_neighbor_finder_and_camera_units():
    rng = np.random.RandomState(5)
    W, H = 200, 120
    n = 4000
    img = np.stack([rng.uniform(0.01, W - 0.01, n), rng.uniform(0.01, H - 0.01, n)], 1)
    cam = np.stack([rng.normal(0, 5, n), rng.normal(0, 2, n), rng.uniform(-2, 40, n)], 1)
    for _ in range(50):
        u, v = rng.uniform(-5, W + 5), rng.uniform(-5, H + 5)
        a = np.empty(4096, np.int32)
        k = O.lib().orc_neighbor_finder(W, H, 9, 7, img.ctypes.data, cam.ctypes.data, n, u, v, a.ctypes.data, 4096)
        assert np.array_equal(a[:k], R.neighbor_finder(W, H, 9, 7, img, cam, u, v))
    for _ in range(200):
        u, v = rng.uniform(-100, 1400), rng.uniform(-100, 500)
        a = np.empty(3)
        O.lib().orc_viewing_ray(*KCAM, u, v, a.ctypes.data)
        assert np.array_equal(a, R.viewing_ray(*KCAM, u, v))
        p3 = np.array([rng.normal(0, 10), rng.normal(0, 3), rng.uniform(-5, 60)])
        uv = np.empty(2)
        ok = O.lib().orc_image_point(*KCAM, p3.ctypes.data, uv.ctypes.data)
        ok_r, uv_r = R.image_point(*KCAM, p3)
        assert bool(ok) == ok_r and np.array_equal(uv, uv_r)


def test_ransac_plane_reference_kat_on_the_reference_code():
    """RansacPlane.CalculateInlersPlane (test_monolidar_fusion.cpp:376-441): the reference's own RansacPlane.cpp on the
    regenerated KAT cloud recovers the plane within +-0.2, and agrees with the oracle's RANSAC statistically."""
    cloud = kat_data.ransac_kat_cloud()
    p = O.default_params()
    p.ransac_plane_distance_treshold = 0.2
    p.ransac_plane_max_iterations = 600
    p.ransac_plane_use_refinement = 1
    p.ransac_plane_refinement_treshold = 0.05
    p.ransac_plane_probability = 0.99
    for seed in (1, 2, 3):
        rc, c, inl = R.ransac_plane(p, cloud, seed)
        assert rc == 0
        c = c * np.sign(c[2])
        assert np.all(np.abs(c - np.array([0.0, 0.0, 1.0, 1.6])) < 0.2), c
        rc2, c2, inl2, it = O.ransac_plane(p, cloud, seed)
        c2 = c2 * np.sign(c2[2])
        assert np.all(np.abs(c2 - np.array([0.0, 0.0, 1.0, 1.6])) < 0.2), c2  # the KAT's own tolerance; its cloud is noisy
        assert np.all(np.abs(c[:3] - c2[:3]) < 0.02)


def test_ransac_on_a_kitti_sweep_reference_vs_oracle():
    """6000-point subsample, perpendicular-plane model, refinement with the un-refined coefficients (RansacPlane.cpp:120):
    same plane to 0.5 degrees / 3 cm from two different random streams; inliers are a subset of the subsample."""
    p = O.yaml_params()
    cloud = synth.points_host(synth.default_config(), 7, 0)
    rc, c, inl = R.ransac_plane(p, cloud, 9)
    rc2, c2, inl2, it = O.ransac_plane(p, cloud, 9)
    assert rc == 0 and rc2 == 0
    c, c2 = c * np.sign(c[2]), c2 * np.sign(c2[2])
    ang = np.degrees(np.arccos(np.clip(np.dot(c[:3], c2[:3]) / np.linalg.norm(c[:3]) / np.linalg.norm(c2[:3]), -1, 1)))
    assert ang < 0.5 and abs(c[3] - c2[3]) < 0.03, (c, c2)
    assert len(inl) <= 6000 and len(inl2) <= 6000 and abs(len(inl) - len(inl2)) < 0.05 * 6000
    assert np.all(np.diff(inl) > 0)  # order preserving subsample


def test_full_pipeline_with_the_references_own_ransac_plane():
    """setInputCloud with a null GroundPlane: the reference fits its RansacPlane; the oracle is handed that plane and
    must reproduce every status and depth of the road path."""
    p = O.yaml_params()
    o, r = pair(p)
    cfg = synth.default_config()
    cloud = synth.points_host(cfg, 2026, 0)
    rng = np.random.RandomState(8)
    uv = np.stack([rng.randint(0, 1241, 12000), rng.randint(230, 376, 12000)], 1).astype(np.float64)  # rows of ground returns
    R.lib().ref_set_seed(4)
    r.set_cloud(cloud, None)
    r.has_plane = True
    coeffs, inl = r.plane()
    d_r, s_r = r.calculate_depth(uv)
    o.set_cloud(cloud)
    d_o, s_o = o.calculate_depth(uv, (coeffs, inl))
    assert np.array_equal(s_o, s_r)
    assert np.all(np.abs(d_o - d_r) <= 1e-9 * np.abs(d_r))
    assert (s_r == 16).sum() >= 10  # the inlier set is a 6000-point subsample, so few windows hold 3 inliers


@pytest.mark.parametrize("seed", range(24))
def test_randomised_configurations(seed):
    """Random cameras, extrinsics, window sizes, thresholds and module switches on random scenes: the oracle must follow
    the reference's code bit for bit through whatever combination of branches a configuration opens."""
    p, cam, T, cloud, uv, plane = PU.random_configuration(seed)
    o, r = pair(p, cam, T)
    d, s = compare(o, r, cloud, uv, plane=plane, what=f"random config {seed}", depth_rtol=1e-9 if plane is not None else 0.0, neighbor_samples=24)
    assert len(set(s.tolist())) >= 2
