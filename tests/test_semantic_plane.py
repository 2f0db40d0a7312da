"""SemanticPlane (RansacPlane.cpp:159-274).

CPU: the numpy restatement (oracle/semantic_plane_np.py) against the reference's own code (oracle/_ref).
GPU: mld_semantic_ground_plane against the restatement and against reference outputs frozen in
tests/golden/ref_golden.npz (sem_* entries).

Contract: the ground-labelled set is bit-exact. The GPU accumulates the plane moments in double, PCL in sequential
float -- over ~10^4 points of a sweep that loses 4-5 digits of the second moments, so the REFERENCE's own plane carries
an error of ~1e-3 in the normal. Coefficients must therefore agree to 2e-3, and the inlier sets may differ only for
points whose distance to the plane is within 2e-3 * (1 + |p|) of the threshold (normal error times lever arm), at most
1 % of the set."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

import ref_lib as R
from mono_lidar_depth_b200 import DepthEstimator, DepthEstimatorParameters, ExceptionPclInvalid, SemanticPlane, synth

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import semantic_plane_np as SP  # noqa: E402

_spec = importlib.util.spec_from_file_location("make_ref_golden", ROOT / "tests" / "golden" / "make_ref_golden.py")
MK = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MK)
G = np.load(ROOT / "tests" / "golden" / "ref_golden.npz")

KT = synth.KITTI_T_LIDAR_TO_CAM
F_, CU, CV = 718.856, 607.1928, 185.2157
COEFF_TOL, MARGIN = 2e-3, 2e-3


def _check_against(coeffs, inl, ref_coeffs, ref_inl, first_model, cloud, thr, what):
    """first_model: the first-pass plane, the one selectWithinDistance is evaluated with (RansacPlane.cpp:248)."""
    assert np.all(np.abs(np.asarray(coeffs) - ref_coeffs) < COEFF_TOL), (what, coeffs, ref_coeffs)
    diff = np.setxor1d(inl, ref_inl)
    xyz = cloud[diff, :3].astype(np.float64)
    dist = np.abs(xyz @ first_model[:3].astype(np.float64) + float(first_model[3]))
    assert np.all(np.abs(dist - thr) < MARGIN * (1.0 + np.linalg.norm(xyz, axis=1))), (what, len(diff), dist[:5])
    assert len(diff) <= 0.01 * max(len(ref_inl), 1), (what, len(diff), len(ref_inl))


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libmld_ref.so not built (no /root/reference here)")
@pytest.mark.parametrize("case", [0, 1, 2])
def test_numpy_restatement_matches_the_reference_code(case):
    cloud, labels, gl, thr = MK.semantic_case(case)
    rc, c_ref, inl_ref = R.semantic_plane(labels, F_, CU, CV, KT, gl, thr, cloud)
    assert rc == 0
    c, inl, kept, first = SP.semantic_plane(cloud, labels, F_, CU, CV, KT, gl, thr)
    assert np.array_equal(inl, inl_ref)  # same float accumulation order -> same first model -> identical selection
    assert np.allclose(c, c_ref, rtol=0, atol=2e-6)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libmld_ref.so not built (no /root/reference here)")
def test_reference_throws_pcl_invalid_without_ground_pixels():
    cloud, labels, gl, thr = MK.semantic_case(0)
    rc, _, _ = R.semantic_plane(np.zeros_like(labels), F_, CU, CV, KT, gl, thr, cloud)
    assert rc == -4
    with pytest.raises(SP.PclInvalid):
        SP.semantic_plane(cloud, np.zeros_like(labels), F_, CU, CV, KT, gl, thr)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libmld_ref.so not built (no /root/reference here)")
def test_restatement_label_test_matches_the_reference_on_boundary_points():
    """The reference does not expose the set kept by its label test, so it is probed one point at a time: a cloud of three
    fixed ground points plus the probe is fitted only when the probe is kept (optimizeModelCoefficients needs more than 3
    points, else the dummy model (0,0,1,0) comes back, RansacPlane.cpp:236-242)."""
    cloud, labels = MK.semantic_boundary_cloud(3000)
    want = np.zeros(len(cloud), bool)
    want[SP.ground_labelled(cloud, labels, F_, CU, CV, KT, [7])] = True
    # three ground-labelled points in front of the camera, well inside the image, |z_lidar| far above the threshold
    Tm = np.vstack([KT[:3], [0, 0, 0, 1]])
    Ti = np.linalg.inv(Tm)
    anchors_cam = np.array([[(100.5 - CU) / F_ * 9, (301.5 - CV) / F_ * 9, 9], [(900.5 - CU) / F_ * 14, (333.5 - CV) / F_ * 14, 14],
                            [(500.5 - CU) / F_ * 6, (251.5 - CV) / F_ * 6, 6]])
    anchors = ((Ti[:3, :3] @ anchors_cam.T).T + Ti[:3, 3]).astype(np.float32)
    assert len(SP.ground_labelled(np.c_[anchors, np.zeros(3)], labels, F_, CU, CV, KT, [7])) == 3
    probe4 = np.zeros((4, 4), np.float32)
    probe4[:3, :3] = anchors
    dummy = np.array([0, 0, 1, 0], np.float32)
    idx = np.r_[np.arange(0, 3000, 2), np.arange(3000, 3008)]
    got = np.zeros(len(cloud), bool)
    for i in idx:
        probe4[3, :3] = cloud[i, :3]
        rc, c, inl = R.semantic_plane(labels, F_, CU, CV, KT, [7], 1e-6, probe4)
        assert rc == 0
        got[i] = not np.array_equal(c, dummy)
    assert np.array_equal(got[idx], want[idx]), np.nonzero(got[idx] != want[idx])[0][:10]
    assert 100 < want[idx].sum() < len(idx) - 100


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libmld_ref.so not built (no /root/reference here)")
@pytest.mark.parametrize("seed", range(8))
def test_restatement_matches_the_reference_on_random_cameras_and_label_images(seed):
    """Random image size, intrinsics, extrinsic, label layout, ground-label set and threshold: inlier set identical, coefficients
    to float rounding; PclInvalid when the reference throws."""
    rng = np.random.RandomState(500 + seed)
    W, H = int(rng.randint(60, 500)), int(rng.randint(40, 300))
    f, cu, cv = float(rng.uniform(0.5, 2.0) * W), float(W * rng.uniform(0.3, 0.7)), float(H * rng.uniform(0.3, 0.7))
    a = rng.normal(0, 0.15, 3)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    U, _, Vt = np.linalg.svd(np.eye(3) + K)
    T = np.zeros((3, 4))
    T[:, :3] = (U @ Vt) @ KT[:, :3]
    T[:, 3] = rng.uniform(-0.4, 0.4, 3)
    labels = rng.randint(0, 12, (H, W)).astype(np.uint8) if seed % 2 else np.zeros((H, W), np.uint8)
    labels[int(H * rng.uniform(0.4, 0.7)):, :] = rng.choice([6, 7, 8, 9])
    gl = sorted(set(rng.choice([3, 6, 7, 8, 9, 11, 200, 300, -4], rng.randint(1, 5)).tolist()))
    thr = float(rng.choice([0.02, 0.1, 0.5, 5.0]))
    cfg = synth.default_config()
    cfg.azimuth_steps = 300
    cloud = synth.points_host(cfg, 900 + seed, seed)
    rc, c_ref, inl_ref = R.semantic_plane(labels, f, cu, cv, T, gl, thr, cloud)
    if rc != 0:
        assert rc == -4
        with pytest.raises(SP.PclInvalid):
            SP.semantic_plane(cloud, labels, f, cu, cv, T, gl, thr)
        return
    c, inl, kept, first = SP.semantic_plane(cloud, labels, f, cu, cv, T, gl, thr)
    assert np.array_equal(inl, inl_ref), (len(inl), len(inl_ref))
    assert np.allclose(c, c_ref, rtol=0, atol=5e-6)


def test_restatement_reproduces_the_frozen_reference_outputs():
    for case in (0, 1, 2):
        cloud, labels, gl, thr = MK.semantic_case(case)
        c, inl, kept, first = SP.semantic_plane(cloud, labels, F_, CU, CV, KT, gl, thr)
        assert np.array_equal(inl, G[f"sem{case}_inliers"])
        assert np.allclose(c, G[f"sem{case}_coeffs"], rtol=0, atol=2e-6)


def _debug_projection(pts):
    """u, v, tz of a few points the way the restatement computes them (for assertion messages)."""
    p = pts[:, :3].astype(np.float64)
    T = KT[:3, :4]
    t = np.stack([((T[i, 0] * p[:, 0] + T[i, 1] * p[:, 1]) + T[i, 2] * p[:, 2]) + T[i, 3] for i in range(3)], 1).astype(np.float32).astype(np.float64)
    with np.errstate(all="ignore"):
        return np.stack([(F_ * t[:, 0] + CU * t[:, 2]) / t[:, 2], (F_ * t[:, 1] + CV * t[:, 2]) / t[:, 2], t[:, 2]], 1)


def _estimator():
    est = DepthEstimator()
    est.InitConfig(DepthEstimatorParameters.reference_yaml(0))
    est.Initialize(synth.kitti_camera(), KT)
    return est


@pytest.mark.gpu
@pytest.mark.parametrize("case", [0, 1, 2])
def test_gpu_semantic_plane_matches_reference_outputs(case):
    cloud, labels, gl, thr = MK.semantic_case(case)
    est = _estimator()
    plane = SemanticPlane(labels, SemanticPlane.Camera(F_, CU, CV, KT), gl, thr, est)
    plane.CalculateInliersPlane(cloud)
    assert plane.isSegmented()
    c, inl, kept, first = SP.semantic_plane(cloud, labels, F_, CU, CV, KT, gl, thr)
    _check_against(plane.getModelCoeffs(), plane.getInlinersIndex(), G[f"sem{case}_coeffs"], G[f"sem{case}_inliers"], first, cloud, thr, f"ref case {case}")
    _check_against(plane.getModelCoeffs(), plane.getInlinersIndex(), c, inl, first, cloud, thr, f"restatement case {case}")
    # the plane plugs into the road path like any GroundPlane
    uv = synth.features_host(synth.default_config(), 3, case, 500)
    d, s = est.CalculateDepth(cloud, uv, plane)[:2]
    assert len(d) == 500 and set(np.unique(s)).issubset(set(range(0, 17)))


@pytest.mark.gpu
def test_gpu_semantic_plane_32_byte_stride_and_pcl_invalid():
    cloud, labels, gl, thr = MK.semantic_case(0)
    est = _estimator()
    c8 = np.zeros((len(cloud), 8), np.float32)
    c8[:, :3] = cloud[:, :3]
    c8[:, 4] = cloud[:, 3]
    a = SemanticPlane(labels, SemanticPlane.Camera(F_, CU, CV, KT), gl, thr, est)
    a.CalculateInliersPlane(cloud)
    b = SemanticPlane(labels, SemanticPlane.Camera(F_, CU, CV, KT), gl, thr, est)
    b.CalculateInliersPlane(c8)
    assert np.array_equal(a.getInlinersIndex(), b.getInlinersIndex()) and np.array_equal(a.getModelCoeffs(), b.getModelCoeffs())
    with pytest.raises(ExceptionPclInvalid):
        SemanticPlane(np.zeros_like(labels), SemanticPlane.Camera(F_, CU, CV, KT), gl, thr, est).CalculateInliersPlane(cloud)
    with pytest.raises(ExceptionPclInvalid):
        SemanticPlane(labels, SemanticPlane.Camera(F_, CU, CV, KT), gl, thr, est).CalculateInliersPlane(np.zeros((0, 4), np.float32))


@pytest.mark.gpu
def test_gpu_semantic_plane_ground_labelled_set_is_bit_exact():
    """The set of points kept by the label test (RansacPlane.cpp:201-222) through the kernel's two float pre-filters and the
    guarded quotient, against the numpy restatement (itself identical to the reference's code on these inputs): sweeps,
    plus a cloud built to sit on every decision boundary -- pixel borders, the image frame, the camera plane, points behind
    the camera, huge / non-finite coordinates."""
    est = _estimator()
    cam = SemanticPlane.Camera(F_, CU, CV, KT)
    for case in (0, 1, 2):
        cloud, labels, gl, thr = MK.semantic_case(case)
        sp = SemanticPlane(labels, cam, gl, thr, est)
        got = est.semanticGroundLabelled(sp, cloud)
        want = np.zeros(len(cloud), bool)
        want[SP.ground_labelled(cloud, labels, F_, CU, CV, KT, gl)] = True
        assert np.array_equal(got, want), (case, int((got != want).sum()))
        assert want.sum() > 1000
    # boundary cloud: projections on / next to integer pixel coordinates, the frame, the camera plane (see make_ref_golden)
    cloud, labels = MK.semantic_boundary_cloud()
    m = len(cloud) - 8
    sp = SemanticPlane(labels, cam, [7], 0.1, est)
    got = est.semanticGroundLabelled(sp, cloud)
    want = np.zeros(len(cloud), bool)
    want[SP.ground_labelled(cloud, labels, F_, CU, CV, KT, [7])] = True
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, (len(bad), bad[:5], cloud[bad[:5]], got[bad[:5]], _debug_projection(cloud[bad[:5]]))
    assert 5000 < want.sum() < m


@pytest.mark.gpu
def test_gpu_semantic_plane_batched_device_api_matches_single_frame_calls():
    """mld_semantic_ground_plane_device: three sweeps and three label images resident on the device, one call; every frame
    must reproduce the stand-alone host call bit for bit (same kernels), and the bitmask must plug into the batched road path."""
    import ctypes as C

    import torch

    est = _estimator()
    cases = [MK.semantic_case(c) for c in (0, 1, 2)]
    thr = 0.1
    gl = [6, 7, 8, 9]
    n = len(cases[0][0])
    pitch = n + 64  # frames further apart than n points: the pitch is honoured
    pts = torch.zeros((3, pitch, 4), dtype=torch.float32, device="cuda")
    labs = torch.zeros((3, 376, 1241), dtype=torch.uint8, device="cuda")
    for i, (cloud, lab, _, _) in enumerate(cases):
        pts[i, :n] = torch.from_numpy(cloud).cuda()
        labs[i] = torch.from_numpy(lab).cuda()
    words = (n + 31) // 32
    coeffs = torch.zeros((3, 4), dtype=torch.float32, device="cuda")
    bits = torch.zeros((3, words), dtype=torch.int32, device="cuda")
    ninl = torch.zeros(3, dtype=torch.int32, device="cuda")
    rc = torch.zeros(3, dtype=torch.int32, device="cuda")
    T = np.ascontiguousarray(KT[:3, :4], np.float64)
    g = np.ascontiguousarray(gl, np.int32)
    est._check(est._lib.mld_semantic_ground_plane_device(est._h, pts.data_ptr(), n, pitch, 16, labs.data_ptr(), 1241, 376, F_, CU, CV,
                                                         T.ctypes.data, g.ctypes.data, len(g), thr, 3, coeffs.data_ptr(), bits.data_ptr(),
                                                         ninl.data_ptr(), rc.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rc.cpu().tolist() == [0, 0, 0]
    hb = bits.cpu().numpy().view(np.uint32)
    for i, (cloud, lab, _, _) in enumerate(cases):
        single = SemanticPlane(lab, SemanticPlane.Camera(F_, CU, CV, KT), gl, thr, est)
        single.CalculateInliersPlane(cloud)
        idx = np.nonzero(np.unpackbits(hb[i].view(np.uint8), bitorder="little")[:n])[0].astype(np.int32)
        assert np.array_equal(idx, single.getInlinersIndex()), i
        assert int(ninl[i]) == len(idx)
        assert np.allclose(coeffs[i].cpu().numpy(), single.getModelCoeffs(), rtol=0, atol=1e-6)  # double atomics: summation order varies
    # a frame without ground pixels reports ExceptionPclInvalid through its return code only
    labs[1].zero_()
    est._check(est._lib.mld_semantic_ground_plane_device(est._h, pts.data_ptr(), n, pitch, 16, labs.data_ptr(), 1241, 376, F_, CU, CV,
                                                         T.ctypes.data, g.ctypes.data, len(g), thr, 3, coeffs.data_ptr(), bits.data_ptr(),
                                                         ninl.data_ptr(), rc.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rc.cpu().tolist() == [0, -7, 0]  # MLD_ERR_PCL_INVALID
    assert int(ninl[1]) == 0 and not bits[1].any()


@pytest.mark.gpu
def test_gpu_batched_semantic_road_path_matches_per_frame_calls():
    """mld_process_frames_device_semantic (SemanticPlane fit + depth estimation per frame, one call) against
    (a) the same sequence with the planes fitted first and handed in (mld_process_frames_device_planes) and
    (b) the reference's per-frame call sequence through the single-frame API with the same plane, and the oracle."""
    import torch

    import oracle_lib as O

    p = DepthEstimatorParameters.reference_yaml(1)
    est = DepthEstimator()
    est.InitConfig(p)
    est.Initialize(synth.kitti_camera(), KT)
    cases = [MK.semantic_case(c) for c in (0, 1, 2)] * 3  # 9 frames -> more than one chunk of the 3-slot pipeline
    thr, gl = 0.1, [6, 7, 8, 9]
    nf, n, F = len(cases), len(cases[0][0]), 700
    cam = SemanticPlane.Camera(F_, CU, CV, KT)
    rng = np.random.RandomState(4)
    uv_h = np.stack([np.stack([rng.randint(0, 1241, F), rng.randint(150, 376, F)], 1).astype(np.float64) for _ in range(nf)])
    pts = torch.from_numpy(np.stack([c[0] for c in cases])).cuda()
    labs = torch.from_numpy(np.stack([c[1] for c in cases])).cuda()
    uv = torch.from_numpy(uv_h).cuda()
    dep = torch.empty((nf, F), dtype=torch.float64, device="cuda")
    sta = torch.empty((nf, F), dtype=torch.int32, device="cuda")
    coeffs = torch.zeros((nf, 4), dtype=torch.float32, device="cuda")
    prc = torch.full((nf,), 99, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    est.processFramesDeviceSemantic(pts.data_ptr(), n, n, 16, labs.data_ptr(), 1241, 376, cam, gl, thr, uv.data_ptr(), F, dep.data_ptr(),
                                    sta.data_ptr(), nf, coeffs.data_ptr(), prc.data_ptr(), st)
    torch.cuda.synchronize()
    assert prc.cpu().tolist() == [0] * nf
    d_sem, s_sem = dep.cpu().numpy().copy(), sta.cpu().numpy().copy()
    assert (s_sem == 16).sum() > 50  # SuccessRoad is exercised
    # (a) planes fitted up front on the device, then handed in
    words = (n + 31) // 32
    c2 = torch.zeros((nf, 4), dtype=torch.float32, device="cuda")
    bits = torch.zeros((nf, words), dtype=torch.int32, device="cuda")
    ninl = torch.zeros(nf, dtype=torch.int32, device="cuda")
    rc = torch.zeros(nf, dtype=torch.int32, device="cuda")
    T = np.ascontiguousarray(KT[:3, :4], np.float64)
    g = np.ascontiguousarray(gl, np.int32)
    est._check(est._lib.mld_semantic_ground_plane_device(est._h, pts.data_ptr(), n, n, 16, labs.data_ptr(), 1241, 376, F_, CU, CV, T.ctypes.data,
                                                         g.ctypes.data, len(g), thr, nf, c2.data_ptr(), bits.data_ptr(), ninl.data_ptr(),
                                                         rc.data_ptr(), st))
    est.processFramesDevicePlanes(pts.data_ptr(), n, n, 16, c2.data_ptr(), bits.data_ptr(), uv.data_ptr(), F, dep.data_ptr(), sta.data_ptr(), nf, st)
    torch.cuda.synchronize()
    assert np.array_equal(coeffs.cpu().numpy(), c2.cpu().numpy())
    assert np.array_equal(sta.cpu().numpy(), s_sem) and np.array_equal(dep.cpu().numpy(), d_sem)
    # (b) per frame through the reference-shaped API with the same plane, and the oracle
    hb = bits.cpu().numpy().view(np.uint32)
    orc = O.Oracle(O.yaml_params())
    orc.initialize(1241, 376, 718.856, 607.1928, 185.2157, KT)
    from mono_lidar_depth_b200 import GroundPlane
    import parity_util as PU
    for i in (0, 4, 8):
        idx = np.nonzero(np.unpackbits(hb[i].view(np.uint8), bitorder="little")[:n])[0].astype(np.int32)
        plane = GroundPlane(c2[i].cpu().numpy(), idx)
        d1, s1 = est.CalculateDepth(cases[i][0], uv_h[i], plane)[:2]
        assert np.array_equal(s1, s_sem[i]) and np.array_equal(d1, d_sem[i])
        orc.set_cloud(cases[i][0])
        d_o, s_o = orc.calculate_depth(uv_h[i], (c2[i].cpu().numpy(), idx))
        PU.assert_depth_status_equal(d_sem[i], s_sem[i], d_o, s_o, f"semantic batch frame {i}")
    # without a road estimator the call is refused like the reference's Initialize would leave _roadDepthEstimator null
    e2 = DepthEstimator()
    e2.InitConfig(DepthEstimatorParameters.reference_yaml(0))
    e2.Initialize(synth.kitti_camera(), KT)
    from mono_lidar_depth_b200 import MldError
    with pytest.raises(MldError):
        e2.processFramesDeviceSemantic(pts.data_ptr(), n, n, 16, labs.data_ptr(), 1241, 376, cam, gl, thr, uv.data_ptr(), F, dep.data_ptr(),
                                       sta.data_ptr(), nf, 0, 0, st)


@pytest.mark.gpu
def test_unsegmented_semantic_plane_through_setinputcloud_segments_itself():
    """DepthEstimator::setInputCloud dispatches virtually to groundPlane->CalculateInliersPlane (DepthEstimator.cpp:281-283): an
    un-segmented SemanticPlane handed to CalculateDepth(cloud, features, plane) must be fitted from its label image, not by the GPU
    RANSAC (what the Python mirror did before round 2) -- same result as segmenting it explicitly first."""
    p = DepthEstimatorParameters.reference_yaml(1)
    est = DepthEstimator()
    est.InitConfig(p)
    est.Initialize(synth.kitti_camera(), KT)
    cloud, labels, gl, thr = MK.semantic_case(1)
    cam = SemanticPlane.Camera(F_, CU, CV, KT)
    uv = synth.features_host(synth.default_config(road=True), 5, 0, 1500)
    explicit = SemanticPlane(labels, cam, gl, thr, estimator=est)
    explicit.CalculateInliersPlane(cloud)
    d_ref, s_ref, _ = est.CalculateDepth(cloud, uv, explicit)
    lazy = SemanticPlane(labels, cam, gl, thr, estimator=est)
    assert not lazy.isSegmented()
    d, s, plane = est.CalculateDepth(cloud, uv, lazy)
    assert plane is lazy and lazy.isSegmented()
    assert np.array_equal(lazy.getModelCoeffs(), explicit.getModelCoeffs())
    assert np.array_equal(lazy.getInlinersIndex(), explicit.getInlinersIndex())
    assert np.array_equal(s, s_ref) and np.array_equal(d, d_ref)
    assert (s == 16).sum() > 20
    # the pair adaptor follows the same rule
    lazy2 = SemanticPlane(labels, cam, gl, thr, estimator=est)
    dl, dc, pl_last, pl_cur = est.CalculateDepthPair(None, np.empty((0, 2)), None, cloud, uv, lazy2)
    assert pl_cur is lazy2 and lazy2.isSegmented() and np.array_equal(dc, d_ref)


def test_feature_layout_is_explicit_for_2x2():
    f = DepthEstimator._features
    a = np.arange(10, dtype=np.float64).reshape(2, 5)
    assert np.array_equal(f(a), a.T) and np.array_equal(f(a.T), a.T)
    b = np.array([[1.0, 2.0], [3.0, 4.0]])
    with pytest.raises(ValueError):
        f(b)
    assert np.array_equal(f(b, "2xF"), b.T) and np.array_equal(f(b, "Fx2"), b)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [0, 1, 2])
def test_gpu_semantic_plane_exact_mode_is_bit_identical_to_the_reference(case):
    """Exact mode (mld_set_semantic_exact): PCL's nine sequential float accumulators in index order, float covariance, the same Jacobi
    solver -- coefficients and inlier set equal the outputs of the reference's own code (tests/golden/ref_golden.npz, made by
    oracle/_ref) bit for bit, where the default double-precision moments only agree to 2e-3."""
    cloud, labels, gl, thr = MK.semantic_case(case)
    est = _estimator()
    est.setSemanticExact(True)
    plane = SemanticPlane(labels, SemanticPlane.Camera(F_, CU, CV, KT), gl, thr, est)
    plane.CalculateInliersPlane(cloud)
    assert np.array_equal(plane.getInlinersIndex(), G[f"sem{case}_inliers"])
    assert np.array_equal(np.asarray(plane.getModelCoeffs(), np.float32).view(np.uint32), G[f"sem{case}_coeffs"].astype(np.float32).view(np.uint32)), (
        plane.getModelCoeffs(), G[f"sem{case}_coeffs"])
    # 32-byte records give the same
    c8 = np.zeros((len(cloud), 8), np.float32)
    c8[:, :3] = cloud[:, :3]
    p8 = SemanticPlane(labels, SemanticPlane.Camera(F_, CU, CV, KT), gl, thr, est)
    p8.CalculateInliersPlane(c8)
    assert np.array_equal(p8.getInlinersIndex(), plane.getInlinersIndex()) and np.array_equal(p8.getModelCoeffs(), plane.getModelCoeffs())


@pytest.mark.gpu
def test_road_depths_under_the_default_fit_stay_within_tolerance_of_the_exact_fit():
    """End to end: the default (double-precision) SemanticPlane fit differs from the reference's float fit by ~2e-3 in the
    coefficients and by a few boundary points in the inlier set. What that does to the road depths: features whose status agrees
    under both planes (all but a small fraction) get depths within the 1e-4 relative contract."""
    import torch

    p = DepthEstimatorParameters.reference_yaml(1)
    cases = [MK.semantic_case(c) for c in (0, 1, 2)]
    nf, n, F = len(cases), len(cases[0][0]), 2000
    thr, gl = 0.1, [6, 7, 8, 9]
    cam = SemanticPlane.Camera(F_, CU, CV, KT)
    cfg = synth.default_config(road=True)
    uv_h = np.stack([synth.features_host(cfg, 9, f, F) for f in range(nf)])
    pts = torch.from_numpy(np.stack([c[0] for c in cases])).cuda()
    labs = torch.from_numpy(np.stack([c[1] for c in cases])).cuda()
    uv = torch.from_numpy(uv_h).cuda()
    st = torch.cuda.current_stream().cuda_stream
    res = []
    for exact in (False, True):
        est = DepthEstimator()
        est.InitConfig(p)
        est.Initialize(synth.kitti_camera(), KT)
        est.setSemanticExact(exact)
        dep = torch.empty((nf, F), dtype=torch.float64, device="cuda")
        sta = torch.empty((nf, F), dtype=torch.int32, device="cuda")
        est.processFramesDeviceSemantic(pts.data_ptr(), n, n, 16, labs.data_ptr(), 1241, 376, cam, gl, thr, uv.data_ptr(), F, dep.data_ptr(),
                                        sta.data_ptr(), nf, 0, 0, st)
        torch.cuda.synchronize()
        res.append((dep.cpu().numpy(), sta.cpu().numpy()))
    (d0, s0), (d1, s1) = res
    assert (s1 == 16).sum() > 100  # the road path is exercised
    same = s0 == s1
    assert same.mean() > 0.995, same.mean()
    road = same & (s1 == 16)
    rel = np.abs(d0[road] - d1[road]) / np.abs(d1[road])
    assert np.quantile(rel, 0.99) <= 1e-4, (np.quantile(rel, 0.99), rel.max())
    assert rel.max() <= 1e-3, rel.max()
    assert np.array_equal(d0[same & (s1 != 16)], d1[same & (s1 != 16)])  # everything off the road path is untouched by the plane
