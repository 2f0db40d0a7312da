"""GPU parity: CUDA path (through the C ABI) against the CPU oracle on identical inputs."""
import collections
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import parity_util as PU
from mono_lidar_depth_b200 import CameraPinhole, DepthEstimator, DepthEstimatorParameters, ExceptionPclInvalid, GroundPlane, MldError, synth

pytestmark = pytest.mark.gpu

KT = synth.KITTI_T_LIDAR_TO_CAM


def kitti_pair(c_params):
    return PU.make_pair(c_params, synth.kitti_camera(), KT)


def test_kitti_shape_frames_non_road():
    """config[1] shape: HDL-64 ~120k points, 1241x376, 2000 integer features, yaml parameters, plane nullptr."""
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    est, orc = kitti_pair(p)
    cfg = synth.default_config()
    hist = collections.Counter()
    for frame in range(4):
        cloud = synth.points_host(cfg, 11, frame)
        uv = synth.features_host(cfg, 11, frame, 2000)
        d, s = PU.compare_frame(est, orc, cloud, uv, what=f"frame {frame}")
        hist.update(s.tolist())
    assert hist[1] > 0 and hist[2] > 0 and hist[3] > 0  # the mix exercises success and failures


def test_visible_and_camera_points_views():
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    est, orc = kitti_pair(p)
    cloud = synth.points_host(synth.default_config(), 3, 0)
    est.setInputCloud(cloud)
    orc.set_cloud(cloud)
    vis = est.getVisible()
    assert np.array_equal(np.nonzero(vis)[0].astype(np.int32), orc.point_index())
    cam_gpu = est.getPointsCloudCameraCs()
    cam_ref = orc.points_camera()
    assert np.array_equal(np.isnan(cam_gpu), np.isnan(cam_ref))
    assert np.array_equal(cam_gpu[~np.isnan(cam_gpu)], cam_ref[~np.isnan(cam_ref)])  # bit-exact FP64 transform
    img = est.getPointsCloudImageCs()
    assert np.array_equal(img, orc.image_points_visible())
    # visible-order compaction on the device: _pointIndex, image points, getPointDepthCamVisible
    idx, img2, dep = est.getVisiblePoints()
    assert np.array_equal(idx, orc.point_index()) and np.array_equal(img2, orc.image_points_visible())
    assert np.array_equal(dep, cam_ref[orc.point_index(), 2])
    assert est.getPointDepthCamVisible(17) == cam_ref[orc.point_index()[17], 2]
    # capacity smaller than the visible count: the count is still reported, only `capacity` entries are written
    import ctypes as C
    small = np.full(8, -1, np.int32)
    nv = C.c_int64(0)
    est._check(est._lib.mld_get_visible_points(est._h, small.ctypes.data, None, None, 5, C.byref(nv)))
    assert nv.value == len(idx) and np.array_equal(small[:5], idx[:5]) and np.all(small[5:] == -1)
    # empty and all-invisible clouds
    est.setInputCloud(np.zeros((0, 4), np.float32))
    assert len(est.getPointIndex()) == 0
    est.setInputCloud(np.full((3000, 4), np.nan, np.float32))
    assert len(est.getPointIndex()) == 0


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_dense_random_scenes_all_branches(seed):
    """Dense clouds with fractional feature coordinates: many neighbours per window, histogram blobs,
    planarity / orthogonality / threshold failures."""
    rng = np.random.RandomState(seed)
    W, H, f, cx, cy = 320, 240, 300.0, 160.3, 119.6
    cam = CameraPinhole(W, H, f, cx, cy)
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    est, orc = PU.make_pair(p, cam, KT)
    cloud = PU.random_scene_cloud(rng, 14000, W, H, f, cx, cy, KT, dense_patches=60)
    uv = np.stack([rng.uniform(-10, W + 10, 3000), rng.uniform(-10, H + 10, 3000)], 1)
    d, s = PU.compare_frame(est, orc, cloud, uv, what=f"dense {seed}")
    hist = collections.Counter(s.tolist())
    assert hist[1] > 50
    assert len(hist) >= 5, hist


@pytest.mark.parametrize("variant", PU.VARIANTS)
def test_parameter_variants(variant):
    rng = np.random.RandomState(42)
    W, H, f, cx, cy = 256, 192, 250.0, 128.0, 96.0
    cam = CameraPinhole(W, H, f, cx, cy)
    p = PU.variant_params(variant)
    est, orc = PU.make_pair(p, cam, KT)
    cloud = PU.random_scene_cloud(rng, 10000, W, H, f, cx, cy, KT, dense_patches=45)
    uv = np.stack([rng.uniform(0, W, 2000), rng.uniform(0, H, 2000)], 1)
    if variant == "pca":
        # eigenvalue ratios are compared in float against thresholds (PCA.cpp:27-37): mean, scatter and the Jacobi solver run in
        # the oracle's operation order on the GPU (mld_common.cuh eig3_sym_regs), so every status is exact, not 99.9 % of them
        est.setInputCloud(cloud)
        orc.set_cloud(cloud)
        d_gpu, s_gpu = est.CalculateDepth(uv)
        d_ref, s_ref = orc.calculate_depth(uv)
        assert np.array_equal(s_gpu, s_ref)
        assert len(set(s_ref.tolist()) & {12, 13, 14}) >= 2  # the PCA verdicts occur
        PU.assert_depth_status_equal(d_gpu, s_gpu, d_ref, s_ref, "pca")
    else:
        PU.compare_frame(est, orc, cloud, uv, what=variant)


@pytest.mark.parametrize("mode", ["mestimator", "leastsquares", "triangle"])
def test_road_path_with_injected_plane(mode):
    """Road features with the same plane + inlier set injected into both sides."""
    p = O.yaml_params()
    p.do_use_ransac_plane = 1
    p.plane_estimator_use_mestimator = 1 if mode == "mestimator" else 0
    p.plane_estimator_use_leastsquares = 1 if mode == "leastsquares" else 0
    p.plane_estimator_use_triangle_maximation = 1 if mode == "triangle" else 0
    est, orc = kitti_pair(p)
    cfg = synth.default_config()
    cloud = synth.points_host(cfg, 5, 1)
    rng = np.random.RandomState(1)
    # plane: the synthetic ground z = -1.73 in the lidar frame; inliers: every finite point near it
    coeffs = np.array([0.0, 0.0, 1.0, 1.73], np.float32)
    dist = np.abs(cloud[:, 2] + 1.73)
    inl = np.nonzero(np.isfinite(dist) & (dist < 0.1))[0].astype(np.int32)
    # features mostly in the lower image half, on the road
    uv = np.stack([rng.randint(0, 1241, 3000), rng.randint(180, 376, 3000)], 1).astype(np.float64)
    d, s = PU.compare_frame(est, orc, cloud, uv, plane=(coeffs, inl), what=mode)
    hist = collections.Counter(s.tolist())
    assert hist[16] > 100, hist  # SuccessRoad is exercised


def test_road_path_sparse_inliers_and_far_gate():
    p = O.yaml_params()
    est, orc = kitti_pair(p)
    cloud = synth.points_host(synth.default_config(), 6, 2)
    coeffs = np.array([0.01, -0.02, 0.999, 1.70], np.float32)
    fin = np.nonzero(np.isfinite(cloud[:, 2]))[0]
    inl = fin[::7].astype(np.int32)
    uv = synth.features_host(synth.default_config(), 6, 2, 2000)
    PU.compare_frame(est, orc, cloud, uv, plane=(coeffs, inl), what="sparse inliers")


def test_edge_cases_empty_nan_offimage():
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    est, orc = kitti_pair(p)
    # empty cloud
    empty = np.zeros((0, 4), np.float32)
    uv = np.array([[10.0, 10.0], [600.0, 200.0]])  # two features as Fx2: a 2x2 array needs its layout spelled out
    est.setInputCloud(empty)
    d, s = est.CalculateDepth(uv, layout="Fx2")
    assert list(s) == [2, 2] and list(d) == [-1, -1]
    # all-NaN cloud, off-image / NaN / huge features, zero features
    cloud = np.full((1000, 4), np.nan, np.float32)
    feats = np.array([[-50.0, -50.0], [5000.0, 100.0], [np.nan, 3.0], [1e12, 5.0], [0.0, 0.0], [1240.9, 375.9]])
    PU.compare_frame(est, orc, cloud, feats, what="nan cloud", neighbor_samples=6)
    d, s = est.CalculateDepth(np.zeros((0, 2)))
    assert len(d) == 0 and len(s) == 0
    # a real frame with the same odd features (window truncation at the borders, (-1,0) -> row 0)
    cloud = synth.points_host(synth.default_config(), 9, 0)
    feats = np.array([[-2.5, 200.0], [1243.0, 300.0], [620.0, -4.2], [620.0, 379.9], [0.0, 375.0], [1240.0, 170.0], [300.3, 250.7]])
    PU.compare_frame(est, orc, cloud, feats, what="border features", neighbor_samples=7)


def test_error_behaviour_mirrors_reference():
    est = DepthEstimator()
    with pytest.raises(RuntimeError, match="InitConfig"):
        est.Initialize(synth.kitti_camera(), KT)
    p = DepthEstimatorParameters.reference_yaml(0)
    est.InitConfig(p)
    with pytest.raises(RuntimeError, match="without 'initialize'"):
        est.setInputCloud(np.zeros((3, 4), np.float32))
    est.Initialize(synth.kitti_camera(), KT)
    with pytest.raises(RuntimeError, match="without 'SetInputCloud'"):
        est.CalculateDepth(np.zeros((1, 2)))
    q = DepthEstimatorParameters.reference_yaml(0)
    q.neighbor_search_mode = 2
    e2 = DepthEstimator()
    e2.InitConfig(q)
    with pytest.raises(MldError, match="neighbor_search_mode has the invalid value"):
        e2.Initialize(synth.kitti_camera(), KT)
    r = DepthEstimatorParameters.reference_yaml(0)
    r.do_use_depth_segmentation = 1  # shipped yaml value: region growing throws (DepthEstimator.cpp:608)
    e3 = DepthEstimator()
    e3.InitConfig(r)
    e3.Initialize(synth.kitti_camera(), KT)
    e3.setInputCloud(np.zeros((3, 4), np.float32))
    with pytest.raises(MldError, match="Region growing not supported"):
        e3.CalculateDepth(np.zeros((1, 2)))
    s = DepthEstimatorParameters.reference_yaml(1)
    e4 = DepthEstimator()
    e4.InitConfig(s)
    e4.Initialize(synth.kitti_camera(), KT)
    with pytest.raises(ExceptionPclInvalid):
        e4.setInputCloud(np.zeros((2, 4), np.float32))  # < 3 points with RANSAC requested (RansacPlane.cpp:44-50)


def test_set_all_depths_to_zero():
    p = O.yaml_params()
    p.set_all_depths_to_zero = 1
    est, orc = kitti_pair(p)
    cloud = synth.points_host(synth.default_config(), 1, 0)
    uv = synth.features_host(synth.default_config(), 1, 0, 100)
    est.setInputCloud(cloud, GroundPlane([0, 0, 1, 1.7], np.zeros(0, np.int32)))
    d, s = est.CalculateDepth(uv)
    assert np.all(s == 1) and np.all(d == -1)


def test_pointxyzi_stride_32():
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    est, orc = kitti_pair(p)
    cfg = synth.default_config()
    c4 = synth.points_host(cfg, 2, 0)
    c8 = np.zeros((len(c4), 8), np.float32)  # pcl::PointXYZI: xyz_ at 0..11, intensity at byte 16
    c8[:, :3] = c4[:, :3]
    c8[:, 4] = c4[:, 3]
    uv = synth.features_host(cfg, 2, 0, 500)
    est.setInputCloud(c8)
    d8, s8 = est.CalculateDepth(uv)
    est.setInputCloud(c4)
    d4, s4 = est.CalculateDepth(uv)
    assert np.array_equal(s8, s4) and np.array_equal(d8, d4)


def test_dense_128_beam_shape():
    """config[3] shape: 128 x 2032 points, 2048x1024 image, 20000 features."""
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    cam = synth.dense_camera()
    est, orc = PU.make_pair(p, cam, KT)
    cfg = synth.default_config(dense=True)
    cloud = synth.points_host(cfg, 4, 0)
    assert len(cloud) == 260096
    uv = synth.features_host(cfg, 4, 0, 20000)
    PU.compare_frame(est, orc, cloud, uv, what="dense 128")


def test_synth_host_equals_device():
    import torch

    est = DepthEstimator()
    est.InitConfig(DepthEstimatorParameters.reference_yaml(0))
    for dense in (False, True):
        cfg = synth.default_config(dense)
        n = synth.points_per_frame(cfg)
        F = 777
        pts = torch.empty((3, n, 4), dtype=torch.float32, device="cuda")
        uv = torch.empty((3, F, 2), dtype=torch.float64, device="cuda")
        synth.points_device(est, cfg, 99, 5, 3, pts.data_ptr())
        synth.features_device(est, cfg, 99, 5, 3, F, uv.data_ptr())
        torch.cuda.synchronize()
        for i in range(3):
            h = synth.points_host(cfg, 99, 5 + i)
            assert np.array_equal(h.view(np.uint32), pts[i].cpu().numpy().view(np.uint32)), f"dense={dense} frame {i}"
            assert np.array_equal(synth.features_host(cfg, 99, 5 + i, F), uv[i].cpu().numpy())


def test_batched_device_and_host_paths_match_per_frame():
    """mld_process_frames_device / _host over a sequence == per-frame setInputCloud + CalculateDepth == oracle."""
    import torch

    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    est, orc = kitti_pair(p)
    cfg = synth.default_config()
    n = synth.points_per_frame(cfg)
    F, nframes = 2000, 37  # not a multiple of the chunk size
    pts = torch.empty((nframes, n, 4), dtype=torch.float32, device="cuda")
    uv = torch.empty((nframes, F, 2), dtype=torch.float64, device="cuda")
    depth = torch.empty((nframes, F), dtype=torch.float64, device="cuda")
    status = torch.empty((nframes, F), dtype=torch.int32, device="cuda")
    synth.points_device(est, cfg, 21, 0, nframes, pts.data_ptr())
    synth.features_device(est, cfg, 21, 0, nframes, F, uv.data_ptr())
    torch.cuda.synchronize()
    est.processFramesDevice(pts.data_ptr(), n, n, 16, uv.data_ptr(), F, depth.data_ptr(), status.data_ptr(), nframes,
                            stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    d_dev, s_dev = depth.cpu().numpy(), status.cpu().numpy()
    # host pipeline
    pts_h, uv_h = pts.cpu().numpy(), uv.cpu().numpy()
    d_host = np.empty((nframes, F), np.float64)
    s_host = np.empty((nframes, F), np.int32)
    est.processFramesHost(pts_h, uv_h, d_host, s_host)
    assert np.array_equal(s_dev, s_host) and np.array_equal(d_dev, d_host)
    for i in (0, 15, 16, 36):
        orc.set_cloud(pts_h[i])
        d_ref, s_ref = orc.calculate_depth(uv_h[i])
        PU.assert_depth_status_equal(d_dev[i], s_dev[i], d_ref, s_ref, f"batched frame {i}")
        est.setInputCloud(pts_h[i])
        d1, s1 = est.CalculateDepth(uv_h[i])
        assert np.array_equal(s1, s_dev[i]) and np.array_equal(d1, d_dev[i])
    assert est.kernelLaunchCount() > 0


@pytest.mark.parametrize("floats_per_record", [4, 8])
@pytest.mark.parametrize("force", ["1", "0"])
def test_host_pipeline_packed_records(floats_per_record, force, monkeypatch):
    """mld_process_frames_host with every chunk packed to 12-byte xyz on the host (MLD_HOST_PACK=1: the AVX-512 / scalar squeeze of
    csrc/mld_host_pack.cpp, H2D of 12 bytes per point, device-side expansion) and with whole records copied (=0) gives the device
    path's results bit for bit, for float4 and for 32-byte pcl::PointXYZI records; ragged point count and chunking."""
    import torch

    monkeypatch.setenv("MLD_HOST_PACK", force)
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    est, _ = kitti_pair(p)
    cfg = synth.default_config()
    n_full = synth.points_per_frame(cfg)
    n = n_full - 13  # not a multiple of 16: pieces start off the 64-byte lines of the staging buffer
    F, nframes = 700, 41
    clouds = np.stack([synth.points_host(cfg, 77, i)[:n] for i in range(nframes)])
    uv = np.stack([synth.features_host(cfg, 77, i, F) for i in range(nframes)])
    rec = np.full((nframes, n, floats_per_record), 1e30, np.float32)  # padding / intensity must never be read as coordinates
    rec[:, :, :3] = clouds[:, :, :3]
    d_host = np.empty((nframes, F), np.float64)
    s_host = np.empty((nframes, F), np.int32)
    est.processFramesHost(rec, uv, d_host, s_host)
    stats = est.hostPipelineStats()
    pts = torch.from_numpy(clouds).cuda()
    uvd = torch.from_numpy(uv).cuda()
    depth = torch.empty((nframes, F), dtype=torch.float64, device="cuda")
    status = torch.empty((nframes, F), dtype=torch.int32, device="cuda")
    est.processFramesDevice(pts.data_ptr(), n, n, 16, uvd.data_ptr(), F, depth.data_ptr(), status.data_ptr(), nframes,
                            stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(status.cpu().numpy(), s_host) and np.array_equal(depth.cpu().numpy(), d_host)
    assert (s_host == 1).sum() > 100
    if force == "1":
        assert stats["frames_packed"] == nframes, stats
    else:
        assert stats["frames_packed"] == 0, stats


@pytest.mark.parametrize("mode", ["warp", "untagged"])
def test_alternative_kernel_modes(mode, monkeypatch):
    """The warp-per-feature kernel alone (MLD_FEATURE_MODE=warp) and cleared (un-tagged) pixel maps (MLD_TAGGED_MAPS=0) give the
    same results as the default path (split gather/solve/road kernels + epoch tags)."""
    if mode == "warp":
        monkeypatch.setenv("MLD_FEATURE_MODE", mode)
    else:
        monkeypatch.setenv("MLD_TAGGED_MAPS", "0")
    p = O.yaml_params()
    est, orc = kitti_pair(p)
    cfg = synth.default_config()
    cloud = synth.points_host(cfg, 13, 0)
    uv = synth.features_host(cfg, 13, 0, 2000)
    dist = np.abs(cloud[:, 2] + 1.73)
    inl = np.nonzero(np.isfinite(dist) & (dist < 0.1))[0].astype(np.int32)
    PU.compare_frame(est, orc, cloud, uv, plane=(np.array([0, 0, 1, 1.73], np.float32), inl), what=mode)
    rng = np.random.RandomState(5)
    W, H, f, cx, cy = 320, 240, 300.0, 160.3, 119.6
    q = O.yaml_params()
    q.do_use_ransac_plane = 0
    est2, orc2 = PU.make_pair(q, CameraPinhole(W, H, f, cx, cy), KT)
    cloud2 = PU.random_scene_cloud(rng, 14000, W, H, f, cx, cy, KT, dense_patches=60)
    uv2 = np.stack([rng.uniform(-10, W + 10, 3000), rng.uniform(-10, H + 10, 3000)], 1)
    PU.compare_frame(est2, orc2, cloud2, uv2, what=mode + " dense")


def test_overflow_features_finish_on_the_warp_kernel():
    """Windows with more neighbours than the thread-per-feature kernel holds (16) fall through to the
    warp-per-feature pass; a dense wall makes most windows overflow."""
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    W, H, f = 160, 120, 200.0
    cam = CameraPinhole(W, H, f, 80.0, 60.0)
    est, orc = PU.make_pair(p, cam, np.eye(4)[:3])
    us, vs = np.meshgrid(np.arange(1, W - 1) + 0.5, np.arange(1, H - 1) + 0.5)
    rng = np.random.RandomState(3)
    z = 9.0 + 0.002 * us.ravel() + rng.normal(0, 0.003, us.size)
    cloud = np.stack([(us.ravel() - 80.0) / f * z, (vs.ravel() - 60.0) / f * z, z, np.zeros_like(z)], 1).astype(np.float32)
    uv = np.stack([rng.uniform(0, W, 1500), rng.uniform(0, H, 1500)], 1)
    d, s = PU.compare_frame(est, orc, cloud, uv, what="overflow")
    ks = [len(orc.neighbors(u, v)) for u, v in uv[:50]]
    assert max(ks) > 16
    assert (s == 1).mean() > 0.5


def test_epoch_tag_wraparound():
    """The pixel map is cleared only when the 14-bit epoch tag is exhausted (16383 uses): run past the
    wrap on a tiny image and check the map after every few hundred clouds and around the wrap."""
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    W, H, f = 48, 32, 60.0
    cam = CameraPinhole(W, H, f, 24.0, 16.0)
    est, orc = PU.make_pair(p, cam, np.eye(4)[:3])
    rng = np.random.RandomState(0)
    clouds = []
    for i in range(7):
        n = 40 + 10 * i
        us, vs = rng.uniform(0, W, n), rng.uniform(0, H, n)
        z = rng.uniform(2, 20, n)
        clouds.append(np.stack([(us - 24.0) / f * z, (vs - 16.0) / f * z, z, np.zeros(n)], 1).astype(np.float32))
    refs = []
    for c in clouds:
        orc.set_cloud(c)
        refs.append(orc.pixel_map_raw())
    for it in range(16500):
        j = it % 7
        est.setInputCloud(clouds[j])
        if it % 997 == 0 or 16370 <= it <= 16400:
            assert np.array_equal(est.getPixelMap(), refs[j]), it


def test_pair_adaptor_matches_two_single_calls():
    """mld_calculate_depth_pair (tracklets_depth's previous + current cloud per frame) == two single calls == oracle."""
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    est, orc = kitti_pair(p)
    cfg = synth.default_config()
    c0, c1 = synth.points_host(cfg, 41, 0), synth.points_host(cfg, 41, 1)
    f0, f1 = synth.features_host(cfg, 41, 0, 1500), synth.features_host(cfg, 41, 1, 2100)
    d0, d1, pl0, pl1 = est.CalculateDepthPair(c0, f0, None, c1, f1, None)
    assert pl0 is None and pl1 is None
    for cloud, feats, d in ((c0, f0, d0), (c1, f1, d1)):
        orc.set_cloud(cloud)
        d_ref, s_ref = orc.calculate_depth(feats)
        assert np.array_equal(np.isnan(d), np.isnan(d_ref))
        np.testing.assert_allclose(d, d_ref, rtol=PU.DEPTH_RTOL)
    # the current cloud stays the estimator's cloud
    d1b, _ = est.CalculateDepth(f1)
    assert np.array_equal(d1b, d1)
    # first frame: no previous cloud -> -1 (tracklet_depth_module.cpp:97-100)
    dl, dc, _, _ = est.CalculateDepthPair(None, f0, None, c1, f1, None)
    assert np.all(dl == -1) and np.array_equal(dc, d1)
    # with RANSAC planes fitted on the GPU for both clouds
    q = O.yaml_params()
    est2, orc2 = kitti_pair(q)
    est2.ransac_seed = 7
    d0, d1, pl0, pl1 = est2.CalculateDepthPair(c0, f0, None, c1, f1, None)
    assert pl0.isSegmented() and pl1.isSegmented()
    for cloud, feats, d, pl in ((c0, f0, d0, pl0), (c1, f1, d1, pl1)):
        orc2.set_cloud(cloud)
        d_ref, s_ref = orc2.calculate_depth(feats, (pl.getModelCoeffs(), pl.getInlinersIndex()))
        np.testing.assert_allclose(d, d_ref, rtol=PU.DEPTH_RTOL)


def test_status_statistics_and_feature_point_packing():
    import torch

    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    est, orc = kitti_pair(p)
    cfg = synth.default_config()
    cloud, uv = synth.points_host(cfg, 43, 0), synth.features_host(cfg, 43, 0, 2000)
    est.setInputCloud(cloud)
    d, s = est.CalculateDepth(uv)
    stats = est.getDepthCalcStats(s)
    counts = collections.Counter(s.tolist())
    assert stats["Success"] == counts[1] and stats["RadiusSearchInsufficientPoints"] == counts[2]
    assert sum(stats.values()) == len(s)
    t_uv = torch.from_numpy(uv).cuda()
    t_d = torch.from_numpy(d).cuda()
    t_s = torch.from_numpy(s).cuda()
    out = torch.empty((len(d), 3), dtype=torch.float32, device="cuda")
    est.packFeaturePointsDevice(t_uv.data_ptr(), t_d.data_ptr(), len(d), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    hist = est.statusHistogramDevice(t_s.data_ptr(), len(s), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert np.array_equal(o[:, 0], uv[:, 0].astype(np.float32)) and np.array_equal(o[:, 2], d.astype(np.float32))
    assert hist[1] == counts[1] and hist.sum() == len(s)


def test_long_sequence_order_and_chunk_independence():
    """Size-independent properties on a sequence long enough to span many chunks, slots and epochs (1500 KITTI-shaped
    frames = 12 chunks of 128 over 3 slots): (1) reversing the frame order reverses the results bit for bit -- a frame's result
    depends neither on its position in a chunk nor on what the slot's epoch-tagged map held before; (2) a second pass over the
    same buffers reproduces the first (the first-point-wins scatter is deterministic under any schedule); (3) sampled frames
    equal the oracle; (4) the status histogram kernel agrees with a host count of the whole block."""
    import torch

    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    est, orc = kitti_pair(p)
    cfg = synth.default_config()
    n = synth.points_per_frame(cfg)
    F, nframes = 500, 1500
    st = torch.cuda.current_stream().cuda_stream
    pts = torch.empty((nframes, n, 4), dtype=torch.float32, device="cuda")
    uv = torch.empty((nframes, F, 2), dtype=torch.float64, device="cuda")
    synth.points_device(est, cfg, 314, 0, nframes, pts.data_ptr(), stream=st)
    synth.features_device(est, cfg, 314, 0, nframes, F, uv.data_ptr(), stream=st)
    out = []
    for order in ("forward", "forward", "reversed"):
        P, U = (pts, uv) if order == "forward" else (pts.flip(0).contiguous(), uv.flip(0).contiguous())
        depth = torch.empty((nframes, F), dtype=torch.float64, device="cuda")
        status = torch.empty((nframes, F), dtype=torch.int32, device="cuda")
        est.processFramesDevice(P.data_ptr(), n, n, 16, U.data_ptr(), F, depth.data_ptr(), status.data_ptr(), nframes, stream=st)
        torch.cuda.synchronize()
        out.append((depth.flip(0) if order == "reversed" else depth, status.flip(0) if order == "reversed" else status))
        del P, U
    (d0, s0), (d1, s1), (d2, s2) = out
    assert torch.equal(s0, s1) and torch.equal(d0, d1), "second pass differs from the first"
    assert torch.equal(s0, s2) and torch.equal(d0, d2), "reversed frame order changes a frame's result"
    for i in (0, 127, 128, 777, 1499):
        orc.set_cloud(synth.points_host(cfg, 314, i))
        d_ref, s_ref = orc.calculate_depth(synth.features_host(cfg, 314, i, F))
        PU.assert_depth_status_equal(d0[i].cpu().numpy(), s0[i].cpu().numpy(), d_ref, s_ref, f"long sequence frame {i}")
    counts = np.bincount(s0.cpu().numpy().ravel(), minlength=21)
    assert np.array_equal(est.statusHistogramDevice(s0.data_ptr(), nframes * F, st), counts)
    assert int(counts.sum()) == nframes * F and counts[1] > 0 and counts[2] > 0


def test_fused_and_separate_launch_pipelines_agree(monkeypatch):
    """Device-resident non-road sequences run K1 of chunk j and the gather of chunk j-1 as one heterogeneous launch (default);
    MLD_FUSE=0 keeps separate launches on per-chunk streams. Same results bit for bit, for chunk counts that exercise the
    first (K1 only) and last (gather only) launch, slot reuse, a ragged last chunk and a features-heavy block ratio."""
    import torch

    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    cfg = synth.default_config()
    n = synth.points_per_frame(cfg)
    st = torch.cuda.current_stream().cuda_stream
    for F, nframes, chunk in ((700, 50, 8), (2000, 29, 512), (9000, 31, 8)):
        outs = []
        for fuse in ("1", "0"):
            monkeypatch.setenv("MLD_FUSE", fuse)
            monkeypatch.setenv("MLD_FUSE_CHUNK", str(chunk))
            monkeypatch.setenv("MLD_CHUNK_FRAMES", str(chunk))
            est, _ = kitti_pair(p)
            assert (est.fusedChunkFrames() > 0) == (fuse == "1")
            pts = torch.empty((nframes, n, 4), dtype=torch.float32, device="cuda")
            uv = torch.empty((nframes, F, 2), dtype=torch.float64, device="cuda")
            depth = torch.full((nframes, F), -7.0, dtype=torch.float64, device="cuda")
            status = torch.full((nframes, F), -7, dtype=torch.int32, device="cuda")
            synth.points_device(est, cfg, 77, 0, nframes, pts.data_ptr(), stream=st)
            synth.features_device(est, cfg, 77, 0, nframes, F, uv.data_ptr(), stream=st)
            for _ in range(2):  # twice: the second pass reuses slots whose maps carry the first pass's epochs
                est.processFramesDevice(pts.data_ptr(), n, n, 16, uv.data_ptr(), F, depth.data_ptr(), status.data_ptr(), nframes, stream=st)
            torch.cuda.synchronize()
            outs.append((depth.cpu().numpy(), status.cpu().numpy()))
        assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][0], outs[1][0]), (F, nframes, chunk)
        assert outs[0][1].min() >= 1  # every feature got a status


def _run_sequence(est, cfg, seed, nframes, F, road=False, passes=1, stride=16, n_override=None):
    """Device-resident synthetic sequence through mld_process_frames_device; returns (depth, status, coeffs) as numpy."""
    import torch

    n = synth.points_per_frame(cfg)
    st = torch.cuda.current_stream().cuda_stream
    pts = torch.empty((nframes, n, 4), dtype=torch.float32, device="cuda")
    uv = torch.empty((nframes, F, 2), dtype=torch.float64, device="cuda")
    synth.points_device(est, cfg, seed, 0, nframes, pts.data_ptr(), stream=st)
    synth.features_device(est, cfg, seed, 0, nframes, F, uv.data_ptr(), stream=st)
    src = pts
    if stride == 32:  # pcl::PointXYZI records
        src = torch.zeros((nframes, n, 8), dtype=torch.float32, device="cuda")
        src[:, :, :3] = pts[:, :, :3]
        src[:, :, 4] = pts[:, :, 3]
    nn = n_override or n
    depth = torch.full((nframes, F), -7.0, dtype=torch.float64, device="cuda")
    status = torch.full((nframes, F), -7, dtype=torch.int32, device="cuda")
    coeffs = torch.zeros((nframes, 4), dtype=torch.float32, device="cuda")
    for _ in range(passes):
        est.processFramesDevice(src.data_ptr(), nn, n, stride, uv.data_ptr(), F, depth.data_ptr(), status.data_ptr(), nframes, road=road,
                                seed=seed, d_plane_coeffs_out=coeffs.data_ptr() if road else 0, stream=st)
    torch.cuda.synchronize()
    return depth.cpu().numpy(), status.cpu().numpy(), coeffs.cpu().numpy(), pts, uv


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["short", "slot_reuse", "ragged_features", "xyzi32", "ragged_cloud", "road"])
def test_fused_pipeline_matches_separate_launches(monkeypatch, case):
    """The fused pipeline (K1 of chunk j + gather of chunk j-1 in one launch, per-class survivor lists, solve and overflow pass on
    the slot streams, five slots) against separate launches per chunk (MLD_FUSE=0), bit for bit, with 16-frame chunks so that
    short sequences still span many launches: a sequence shorter than one chunk, slot reuse inside one call and across calls,
    feature counts that are not a multiple of the block size, 32-byte PointXYZI records, a cloud whose size is not a multiple
    of the tile, and the road path with a RANSAC plane per frame."""
    p = O.yaml_params()
    road = case == "road"
    p.do_use_ransac_plane = 1 if road else 0
    cfg = synth.default_config(road=road)
    nframes, F, passes, stride, nov = {"short": (5, 300, 1, 16, None), "slot_reuse": (150, 512, 2, 16, None),
                                       "ragged_features": (40, 2000 + 77, 1, 16, None), "xyzi32": (40, 700, 1, 32, None),
                                       "ragged_cloud": (40, 700, 1, 16, 120000 - 517), "road": (70, 1500, 2, 16, None)}[case]
    monkeypatch.setenv("MLD_FUSE_CHUNK", "16")
    monkeypatch.setenv("MLD_CHUNK_FRAMES", "16")
    outs = []
    for fuse in ("1", "0"):
        monkeypatch.setenv("MLD_FUSE", fuse)
        est, _ = kitti_pair(p)
        outs.append(_run_sequence(est, cfg, 91, nframes, F, road=road, passes=passes, stride=stride, n_override=nov))
    assert np.array_equal(outs[0][1], outs[1][1]), case
    assert np.array_equal(outs[0][0], outs[1][0]), case
    assert np.array_equal(outs[0][2], outs[1][2]), case
    assert outs[0][1].min() >= 1  # every feature got a status
    if road:
        assert (outs[0][1] == 16).sum() > 0


@pytest.mark.gpu
def test_fused_pipeline_against_oracle(monkeypatch):
    """Frames of a fused-pipeline sequence checked against the oracle directly (status exact, depth 1e-4), including frames at
    both ends of the sequence and on both sides of chunk boundaries and of the first slot reuse."""
    monkeypatch.setenv("MLD_FUSE_CHUNK", "8")
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    est, orc = kitti_pair(p)
    assert est.fusedChunkFrames() == 8
    cfg = synth.default_config()
    d, s, _, pts, uv = _run_sequence(est, cfg, 123, 70, 2000)
    for i in (0, 1, 7, 8, 39, 40, 41, 68, 69):
        orc.set_cloud(pts[i].cpu().numpy())
        d_ref, s_ref = orc.calculate_depth(uv[i].cpu().numpy())
        PU.assert_depth_status_equal(d[i], s[i], d_ref, s_ref, f"fused pipeline frame {i}")
    assert (s == 1).mean() > 0.2  # the calibrated workload (~30 % Success): the geometry tail is exercised


@pytest.mark.gpu
def test_full_windows_take_the_bigger_slabs_and_the_warp_path(monkeypatch):
    """Windows with more points than the main solve kernel's slab (9) go to its 16-entry instantiation, fuller ones to the
    warp-per-feature overflow pass: a 20 x 20 pixel window over the dense 128-beam sweep overflows nearly every feature. Fused
    pipeline against separate launches, and sampled frames against the oracle."""
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    p.pixelarea_search_witdh = 20
    p.pixelarea_search_height = 20
    cfg = synth.default_config(dense=True)
    outs = []
    for fuse in ("1", "0"):
        monkeypatch.setenv("MLD_FUSE", fuse)
        est, orc = PU.make_pair(p, synth.dense_camera(), synth.KITTI_T_LIDAR_TO_CAM)
        outs.append(_run_sequence(est, cfg, 5, 6, 3000))
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][0], outs[1][0])
    assert (outs[0][1] == 1).sum() > 100
    d, s, _, pts, uv = outs[0]
    for i in (0, 5):
        orc.set_cloud(pts[i].cpu().numpy())
        d_ref, s_ref = orc.calculate_depth(uv[i].cpu().numpy())
        PU.assert_depth_status_equal(d[i], s[i], d_ref, s_ref, f"full windows frame {i}")


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(24))
def test_randomised_configurations(seed):
    """The random configurations of tests/test_ref_pin.py (where the oracle is checked against the reference's code) on the
    GPU: pixel map, neighbour lists and statuses bit-exact, depths within 1e-4 relative."""
    p, cam, T, cloud, uv, plane = PU.random_configuration(seed)
    est, orc = PU.make_pair(p, CameraPinhole(*cam), T)
    PU.compare_frame(est, orc, cloud, uv, plane=plane, what=f"random config {seed}", neighbor_samples=24)


def test_pair_adaptor_keeps_the_previous_cloud_on_the_device():
    """Walking a sequence the way TrackletDepthModule does (previous + current cloud per frame, tracklet_depth_module.cpp:318-354):
    with resident=True frame t's cloud and pixel map stay on the device as frame t+1's previous cloud -- same depths as uploading
    both clouds every frame, with and without the road path."""
    cfg = synth.default_config(road=True)
    for road in (0, 1):
        p = O.yaml_params()
        p.do_use_ransac_plane = road
        est_a, _ = kitti_pair(p)
        est_b, _ = kitti_pair(p)
        clouds = [synth.points_host(cfg, 31, f) for f in range(4)]
        feats = [synth.features_host(cfg, 31, f, 800) for f in range(4)]
        prev_cloud, plane_a, plane_b = None, None, None
        for t in range(4):
            f_last = feats[t][:300] + 1.0  # features of the previous frame (where the new tracklets were one frame ago)
            dl_a, dc_a, pl_a, pc_a = est_a.CalculateDepthPair(prev_cloud, f_last, plane_a, clouds[t], feats[t], None)
            dl_b, dc_b, pl_b, pc_b = est_b.CalculateDepthPair(None, f_last, plane_b, clouds[t], feats[t], None, resident=True)
            assert np.array_equal(dc_a, dc_b) and np.array_equal(dl_a, dl_b), (road, t)
            if t == 0:
                assert np.all(dl_b == -1)
            else:
                assert (dl_b >= 0).sum() > 20
            prev_cloud, plane_a, plane_b = clouds[t], pc_a, pc_b  # the current plane becomes the previous one (groundPlaneLast_)
            if road:
                assert np.array_equal(pc_a.getInlinersIndex(), pc_b.getInlinersIndex())
        # a batched call in between drops the resident cloud: the next resident call has no previous cloud
        est_b.processFramesHost(np.stack(clouds[:2]), np.stack(feats[:2]), np.empty((2, 800)), np.empty((2, 800), np.int32))
        dl_b, _, _, _ = est_b.CalculateDepthPair(None, feats[0][:10], None, clouds[0], feats[0], None, resident=True)
        assert np.all(dl_b == -1)


@pytest.mark.parametrize("floats_per_record", [4, 8])
def test_per_call_upload_paths_agree(floats_per_record, monkeypatch):
    """setInputCloud uploads a cloud in pageable memory through the host-packed path (12-byte xyz in pinned staging, expanded
    on the device; the call returns before the projection has finished), a pinned or small cloud as it is. Same results bit for
    bit for pageable / pinned / MLD_HOST_PACK=0 sources, for back-to-back setInputCloud calls (staging reuse, the earlier cloud
    is dropped), for a batched call issued right behind an un-waited setInputCloud, and against the oracle."""
    import torch

    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    cfg = synth.default_config()
    n, F = synth.points_per_frame(cfg), 1500
    make = synth.points_host if floats_per_record == 4 else synth.points_host_xyzi32
    clouds = [make(cfg, 31, f) for f in range(3)]
    feats = [synth.features_host(cfg, 31, f, F) for f in range(3)]
    est, orc = kitti_pair(p)
    ref = []
    for c, uv in zip(clouds, feats):  # pageable numpy arrays: packed uploads
        est.setInputCloud(c)
        ref.append(est.CalculateDepth(uv))
    orc.set_cloud(np.ascontiguousarray(clouds[1][:, [0, 1, 2, 4 if floats_per_record == 8 else 3]]))
    d_ref, s_ref = orc.calculate_depth(feats[1])
    PU.assert_depth_status_equal(ref[1][0], ref[1][1], d_ref, s_ref, "packed per-call upload")
    # back to back: the second cloud replaces the first while the first upload may still be in flight
    est.setInputCloud(clouds[0])
    est.setInputCloud(clouds[2])
    d, s = est.CalculateDepth(feats[2])
    assert np.array_equal(s, ref[2][1]) and np.array_equal(d, ref[2][0])
    # pinned source: copied as it is
    pinned = torch.from_numpy(clouds[1]).pin_memory()
    est.setInputCloud(pinned.numpy())
    d, s = est.CalculateDepth(feats[1])
    assert np.array_equal(s, ref[1][1]) and np.array_equal(d, ref[1][0])
    # a batched call right behind an un-waited setInputCloud (both use slot 0's maps)
    pts4 = torch.from_numpy(np.ascontiguousarray(np.stack([c[:, :4] if floats_per_record == 4 else c[:, [0, 1, 2, 4]] for c in clouds]))).cuda()
    uvd = torch.from_numpy(np.stack(feats)).cuda()
    depth = torch.empty((3, F), dtype=torch.float64, device="cuda")
    status = torch.empty((3, F), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    est.setInputCloud(clouds[0])
    est.processFramesDevice(pts4.data_ptr(), n, n, 16, uvd.data_ptr(), F, depth.data_ptr(), status.data_ptr(), 3, stream=st)
    torch.cuda.synchronize()
    for i in range(3):
        assert np.array_equal(status[i].cpu().numpy(), ref[i][1]) and np.array_equal(depth[i].cpu().numpy(), ref[i][0]), i
    # the un-packed path
    monkeypatch.setenv("MLD_HOST_PACK", "0")
    est0, _ = kitti_pair(p)
    est0.setInputCloud(clouds[1])
    d, s = est0.CalculateDepth(feats[1])
    assert np.array_equal(s, ref[1][1]) and np.array_equal(d, ref[1][0])


def test_parameters_changed_between_initconfig_and_initialize_take_effect():
    """The reference keeps the shared parameter block and builds its modules from it in Initialize (DepthEstimator.cpp:46-127): a
    value changed after InitConfig but before Initialize counts. The mirror (and the C++ shim, shim_selftest) hand a changed
    block to the device again in Initialize."""
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    params = PU.params_from_c(p)
    params.radiusSearch_count_min = 5  # not the value the oracle runs with
    est = DepthEstimator()
    est.InitConfig(params)
    params.radiusSearch_count_min = p.radiusSearch_count_min
    est.Initialize(synth.kitti_camera(), KT)
    orc = O.Oracle(p)
    cam = synth.kitti_camera()
    orc.initialize(1241, 376, cam.focal_length_, cam.principal_point_x_, cam.principal_point_y_, KT)
    cfg = synth.default_config()
    PU.compare_frame(est, orc, synth.points_host(cfg, 41, 0), synth.features_host(cfg, 41, 0, 1200), what="late parameter change")
