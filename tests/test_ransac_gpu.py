"""K4 ground-plane RANSAC on the GPU against the oracle's restatement and the reference's +-0.2 KAT."""
import collections

import numpy as np
import pytest

import kat_data
import oracle_lib as O
import parity_util as PU
from mono_lidar_depth_b200 import DepthEstimator, GroundPlane, synth

pytestmark = pytest.mark.gpu
KT = synth.KITTI_T_LIDAR_TO_CAM


def _est(c_params):
    est = DepthEstimator()
    est.InitConfig(PU.params_from_c(c_params))
    est.Initialize(synth.kitti_camera(), KT)
    return est


def _ulp_close(a, b, ulps=2):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32)))


def test_reference_kat_within_0p2():
    """RansacPlane.CalculateInlersPlane (test_monolidar_fusion.cpp:376-441)."""
    cloud = kat_data.ransac_kat_cloud()
    p = O.default_params()
    p.ransac_plane_distance_treshold = 0.2
    p.ransac_plane_max_iterations = 600
    p.ransac_plane_use_refinement = 1
    p.ransac_plane_refinement_treshold = 0.05
    p.ransac_plane_probability = 0.99
    est = _est(p)
    for seed in (0, 7, 1234):
        pl = est.estimateGroundPlane(cloud, seed)
        c = pl.getModelCoeffs()
        sign = 1.0 if c[2] > 0 else -1.0
        assert abs(c[0]) < 0.2 and abs(c[1]) < 0.2 and abs(c[2] - sign) < 0.2 and abs(c[3] - sign * 1.6) < 0.2
        rc, c_ref, inl_ref, it_ref = O.ransac_plane(p, cloud, seed)
        assert rc == 0
        assert pl.iterations == it_ref
        assert np.array_equal(pl.getInlinersIndex(), inl_ref)
        assert _ulp_close(c, c_ref), (c, c_ref)


@pytest.mark.parametrize("variant", ["yaml", "passthrough", "passthrough_nan_xy", "no_refine", "few_iterations"])
def test_gpu_ransac_equals_oracle_on_synthetic_sweeps(variant):
    p = O.yaml_params()
    if variant.startswith("passthrough"):
        p.ransac_plane_min_z = -3.0
        p.ransac_plane_max_z = -0.5
    elif variant == "no_refine":
        p.ransac_plane_use_refinement = 0
    elif variant == "few_iterations":
        p.ransac_plane_max_iterations = 5
    est = _est(p)
    cfg = synth.default_config()
    for frame in range(3):
        cloud = synth.points_host(cfg, 31, frame)
        if variant == "passthrough_nan_xy":
            # pcl::PassThrough drops a point whose x or y is not finite even when its z passes the limits: such points must not
            # reach the 6000-point subsample (a NaN hypothesis / poisoned refit otherwise)
            rng = np.random.RandomState(frame)
            bad = rng.choice(len(cloud), 4000, replace=False)
            cloud[bad[:2000], 0] = np.nan
            cloud[bad[2000:], 1] = np.inf
        seed = 1000 + frame
        pl = est.estimateGroundPlane(cloud, seed)
        rc, c_ref, inl_ref, it_ref = O.ransac_plane(p, cloud, seed)
        assert rc == 0
        assert pl.iterations == it_ref, (variant, frame)
        assert np.array_equal(pl.getInlinersIndex(), inl_ref), (variant, frame)
        assert _ulp_close(pl.getModelCoeffs(), c_ref), (pl.getModelCoeffs(), c_ref)
        if variant == "passthrough_nan_xy":
            assert np.isfinite(pl.getModelCoeffs()).all() and not np.isin(bad, pl.getInlinersIndex()).any()
        # the fitted plane is the synthetic ground (z = -1.73 in the lidar frame)
        c = pl.getModelCoeffs()
        fin = np.isfinite(cloud[:, 2])
        ground_frac = (np.abs(cloud[fin, 2] + 1.73) < 0.3).mean()
        if variant != "few_iterations" and ground_frac > 0.45:
            assert abs(abs(c[2]) - 1.0) < 0.01 and abs(abs(c[3]) - 1.73) < 0.1


def test_small_clouds():
    p = O.yaml_params()
    est = _est(p)
    from mono_lidar_depth_b200 import ExceptionPclInvalid

    with pytest.raises(ExceptionPclInvalid):
        est.estimateGroundPlane(np.zeros((2, 4), np.float32), 0)
    rng = np.random.RandomState(0)
    cloud = np.zeros((50, 4), np.float32)
    cloud[:, :2] = rng.uniform(-5, 5, (50, 2))
    cloud[:, 2] = -1.5 + rng.normal(0, 0.01, 50)
    pl = est.estimateGroundPlane(cloud, 3)
    rc, c_ref, inl_ref, it_ref = O.ransac_plane(p, cloud, 3)
    assert rc == 0 and pl.iterations == it_ref and np.array_equal(pl.getInlinersIndex(), inl_ref)
    assert _ulp_close(pl.getModelCoeffs(), c_ref)


def test_set_cloud_fits_plane_and_road_depths_match_oracle():
    """setInputCloud with a null plane creates and fits a RansacPlane (DepthEstimator.cpp:274-283); the
    depths computed with it equal the oracle's when the oracle is handed the same plane."""
    p = O.yaml_params()
    est, orc = PU.make_pair(p, synth.kitti_camera(), KT)
    cfg = synth.default_config()
    cloud = synth.points_host(cfg, 8, 0)
    uv = synth.features_host(cfg, 8, 0, 2000)
    est.ransac_seed = 55
    d, s, plane = est.CalculateDepth(cloud, uv, None)
    assert plane is not None and plane.isSegmented()
    orc.set_cloud(cloud)
    d_ref, s_ref = orc.calculate_depth(uv, (plane.getModelCoeffs(), plane.getInlinersIndex()))
    PU.assert_depth_status_equal(d, s, d_ref, s_ref, "ransac road")
    # a segmented plane is reused, not re-fitted (:281-283)
    c0 = plane.getModelCoeffs().copy()
    est.setInputCloud(cloud, plane)
    assert np.array_equal(plane.getModelCoeffs(), c0)


@pytest.mark.parametrize("nframes,F", [(19, 1500), (330, 400)])
def test_batched_road_sequence_matches_oracle(nframes, F):
    """19 frames: launches of a few frames each, one 8-CTA cluster per frame (ransac_cluster_kernel); 330 frames: launches of 66
    frames, one CTA per frame with the whole subsample in its shared memory (ransac_frame_kernel). Same body, same results."""
    import torch

    p = O.yaml_params()
    est, orc = PU.make_pair(p, synth.kitti_camera(), KT)
    cfg = synth.default_config()
    n = synth.points_per_frame(cfg)
    seed = 4242
    pts = torch.empty((nframes, n, 4), dtype=torch.float32, device="cuda")
    uv = torch.empty((nframes, F, 2), dtype=torch.float64, device="cuda")
    depth = torch.empty((nframes, F), dtype=torch.float64, device="cuda")
    status = torch.empty((nframes, F), dtype=torch.int32, device="cuda")
    coeffs = torch.empty((nframes, 4), dtype=torch.float32, device="cuda")
    synth.points_device(est, cfg, 77, 0, nframes, pts.data_ptr())
    synth.features_device(est, cfg, 77, 0, nframes, F, uv.data_ptr())
    est.processFramesDevice(pts.data_ptr(), n, n, 16, uv.data_ptr(), F, depth.data_ptr(), status.data_ptr(), nframes, road=True,
                            seed=seed, d_plane_coeffs_out=coeffs.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    pts_h, uv_h = pts.cpu().numpy(), uv.cpu().numpy()
    hist = collections.Counter()
    for i in (0, 7, 16, nframes - 1):
        rc, c_ref, inl_ref, _ = O.ransac_plane(p, pts_h[i], seed + i)
        assert rc == 0
        c_gpu = coeffs[i].cpu().numpy()
        assert _ulp_close(c_gpu, c_ref), (i, c_gpu, c_ref)
        orc.set_cloud(pts_h[i])
        d_ref, s_ref = orc.calculate_depth(uv_h[i], (c_gpu, inl_ref))
        PU.assert_depth_status_equal(depth[i].cpu().numpy(), status[i].cpu().numpy(), d_ref, s_ref, f"road frame {i}")
        hist.update(s_ref.tolist())
    assert hist[1] > 0
