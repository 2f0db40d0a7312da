"""The source-compatible C++ DepthEstimator shim (shim/) driven the way tracklets_depth drives the
reference, diffed against the oracle."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
import parity_util as PU
from mono_lidar_depth_b200 import synth

ROOT = Path(__file__).resolve().parent.parent


def test_shim_builds_against_stub_headers():
    subprocess.check_call(["make", "-C", str(ROOT / "shim")])
    assert (ROOT / "shim" / "libmonolidar_fusion_b200.so").exists()
    assert (ROOT / "shim" / "shim_selftest").exists()
    out = subprocess.check_output(["nm", "-DC", str(ROOT / "shim" / "libmonolidar_fusion_b200.so")], text=True)
    for sym in ("Mono_Lidar::DepthEstimator::Initialize", "Mono_Lidar::DepthEstimator::InitConfig",
                "Mono_Lidar::DepthEstimator::setInputCloud", "Mono_Lidar::DepthEstimator::CalculateDepth",
                "Mono_Lidar::RansacPlane::CalculateInliersPlane", "Mono_Lidar::SemanticPlane::CalculateInliersPlane",
                "Mono_Lidar::DepthEstimator::getPointDepthCamVisible", "Mono_Lidar::DepthEstimator::getCloudInterpolated",
                "Mono_Lidar::DepthEstimator::getCloudInterpolatedPlane", "Mono_Lidar::DepthEstimator::getCloudNeighbors",
                "Mono_Lidar::DepthEstimator::getCloudTriangleCorners", "Mono_Lidar::DepthEstimator::getCloudRansacPlane",
                "Mono_Lidar::DepthEstimator::getPointsCloudImageCs", "Mono_Lidar::DepthEstimator::getCloudCameraCs"):
        assert sym in out, sym


REF_CALLER = Path("/root/reference/tracklets_depth/src/tracklet_depth_module.cpp")


@pytest.mark.skipif(not REF_CALLER.exists(), reason="/root/reference is not present (GPU box): the prebuilt binary is used")
def test_the_references_own_caller_compiles_unmodified_against_the_shim():
    """tracklets_depth/src/tracklet_depth_module.cpp and its header (which calls getDepthCalcStats / getCloudCameraCs /
    getCloudInterpolated / getPointsCloudImageCs inline, tracklet_depth_module.h:109-123) are compiled from /root/reference, byte
    for byte as they are, against shim/include; only ROS / OpenCV / feature_tracking come from stand-ins (tests/stubs_ros)."""
    subprocess.check_call(["make", "-C", str(ROOT / "shim"), "-B", "caller_dropin"])
    assert (ROOT / "shim" / "_ref_caller" / "tracklet_depth_module.o").exists()
    out = subprocess.check_output(["nm", "-C", str(ROOT / "shim" / "_ref_caller" / "tracklet_depth_module.o")], text=True)
    # the caller's translation unit references the shim's DepthEstimator entry points (undefined here, resolved by the shim library)
    for sym in ("U Mono_Lidar::DepthEstimator::CalculateDepth", "U Mono_Lidar::DepthEstimator::Initialize",
                "U Mono_Lidar::DepthEstimator::InitConfig", "U Mono_Lidar::SemanticPlane::SemanticPlane"):
        assert sym in out, sym


@pytest.mark.gpu
def test_the_references_own_caller_runs_on_the_gpu():
    """TrackletDepthModule::process (the reference's code) drives the shim for three frames: SemanticPlane per frame, previous and
    current cloud per frame; the depths it stores in its tracklets equal direct calls into the shim."""
    exe = ROOT / "shim" / "_ref_caller" / "caller_dropin"
    if not exe.exists():
        pytest.skip("shim/_ref_caller/caller_dropin was not built (needs /root/reference at build time)")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    assert "caller drop-in ok" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("use_plane", [0, 1])
def test_shim_matches_oracle(tmp_path, use_plane):
    subprocess.check_call(["make", "-C", str(ROOT / "shim")])
    cfg = synth.default_config()
    cloud = synth.points_host(cfg, 77, 0)
    uv = synth.features_host(cfg, 77, 0, 2000)
    cloud.tofile(tmp_path / "pts.f32")
    uv.tofile(tmp_path / "uv.f64")
    r = subprocess.run([str(ROOT / "shim" / "shim_selftest"), str(tmp_path / "pts.f32"), str(tmp_path / "uv.f64"),
                        str(tmp_path / "d.f64"), str(tmp_path / "s.i32"), str(use_plane), str(tmp_path / "plane.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    d = np.fromfile(tmp_path / "d.f64", np.float64)
    s = np.fromfile(tmp_path / "s.i32", np.int32)
    p = O.yaml_params()
    p.do_use_ransac_plane = use_plane
    orc = O.Oracle(p)
    cam = synth.kitti_camera()
    orc.initialize(1241, 376, cam.focal_length_, cam.principal_point_x_, cam.principal_point_y_, synth.KITTI_T_LIDAR_TO_CAM)
    orc.set_cloud(cloud)
    plane = None
    if use_plane:
        raw = np.fromfile(tmp_path / "plane.bin", np.uint8)
        coeffs = raw[:16].view(np.float32)
        inl = raw[16:].view(np.int32)
        plane = (coeffs, inl)
        rc, c_ref, inl_ref, _ = O.ransac_plane(p, cloud, 99)
        assert rc == 0 and np.array_equal(inl, inl_ref)
    d_ref, s_ref = orc.calculate_depth(uv, plane)
    PU.assert_depth_status_equal(d, s, d_ref, s_ref, f"shim plane={use_plane}")


@pytest.mark.gpu
def test_shim_semantic_plane_like_the_production_caller(tmp_path):
    """tracklets_depth builds a SemanticPlane from the label image and hands it to CalculateDepth
    (tracklet_depth_module.cpp:269-284, 318-330): same call sequence through the C++ shim, checked against the numpy
    restatement (plane) and the oracle (depths with that plane)."""
    import importlib.util
    import sys

    sys.path.insert(0, str(ROOT / "oracle"))
    import semantic_plane_np as SP

    spec = importlib.util.spec_from_file_location("make_ref_golden", ROOT / "tests" / "golden" / "make_ref_golden.py")
    MK = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(MK)
    subprocess.check_call(["make", "-C", str(ROOT / "shim")])
    cloud, labels, gl, thr = MK.semantic_case(0)
    uv = synth.features_host(synth.default_config(), 77, 0, 2000)
    cloud.tofile(tmp_path / "pts.f32")
    uv.tofile(tmp_path / "uv.f64")
    labels.tofile(tmp_path / "labels.u8")
    r = subprocess.run([str(ROOT / "shim" / "shim_selftest"), str(tmp_path / "pts.f32"), str(tmp_path / "uv.f64"),
                        str(tmp_path / "d.f64"), str(tmp_path / "s.i32"), "2", str(tmp_path / "plane.bin"),
                        str(tmp_path / "labels.u8"), repr(thr)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    raw = np.fromfile(tmp_path / "plane.bin", np.uint8)
    coeffs, inl = raw[:16].view(np.float32), raw[16:].view(np.int32)
    c_ref, inl_ref, kept, first = SP.semantic_plane(cloud, labels, 718.856, 607.1928, 185.2157, synth.KITTI_T_LIDAR_TO_CAM, gl, thr)
    assert np.all(np.abs(coeffs - c_ref) < 2e-3) and len(np.setxor1d(inl, inl_ref)) <= 0.01 * len(inl_ref)
    p = O.yaml_params()
    orc = O.Oracle(p)
    cam = synth.kitti_camera()
    orc.initialize(1241, 376, cam.focal_length_, cam.principal_point_x_, cam.principal_point_y_, synth.KITTI_T_LIDAR_TO_CAM)
    orc.set_cloud(cloud)
    d_ref, s_ref = orc.calculate_depth(uv, (coeffs, inl))
    PU.assert_depth_status_equal(np.fromfile(tmp_path / "d.f64", np.float64), np.fromfile(tmp_path / "s.i32", np.int32), d_ref, s_ref, "shim semantic")
