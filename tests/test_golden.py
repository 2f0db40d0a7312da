"""End-to-end golden fixture (tests/golden/golden_small.npz, made by tests/golden/make_golden.py)."""
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
import parity_util as PU
from mono_lidar_depth_b200 import CameraPinhole, DepthEstimator, GroundPlane

G = np.load(Path(__file__).resolve().parent / "golden" / "golden_small.npz")


def _camera():
    W, H, f, cx, cy = G["camera"]
    return CameraPinhole(int(W), int(H), float(f), float(cx), float(cy))


def test_oracle_reproduces_the_golden_fixture():
    cam = _camera()
    o = O.Oracle(O.yaml_params())
    o.initialize(cam.width_, cam.height_, cam.focal_length_, cam.principal_point_x_, cam.principal_point_y_, G["T"])
    o.set_cloud(G["cloud"])
    assert np.array_equal(o.pixel_map_raw(), G["pixel_map"])
    assert np.array_equal(o.point_index(), G["point_index"])
    d, s = o.calculate_depth(G["uv"])
    assert np.array_equal(s, G["status_noplane"]) and np.array_equal(d, G["depth_noplane"])
    d, s = o.calculate_depth(G["uv"], (G["plane_coeffs"], G["plane_inliers"]))
    assert np.array_equal(s, G["status_plane"]) and np.array_equal(d, G["depth_plane"])
    rc, c, inl, it = O.ransac_plane(O.yaml_params(), G["cloud"], 77)
    assert rc == int(G["ransac_rc"]) and it == int(G["ransac_iterations"])
    assert np.array_equal(c, G["ransac_coeffs"]) and np.array_equal(inl, G["ransac_inliers"])
    assert len(set(G["status_noplane"].tolist())) >= 5 and (G["status_plane"] == 16).sum() > 0


@pytest.mark.gpu
def test_gpu_matches_the_golden_fixture():
    est = DepthEstimator()
    est.InitConfig(PU.params_from_c(O.yaml_params()))
    est.Initialize(_camera(), G["T"])
    est.setInputCloud(G["cloud"], GroundPlane(G["plane_coeffs"], G["plane_inliers"]))
    assert np.array_equal(est.getPixelMap(), G["pixel_map"])
    d, s = est.CalculateDepth(G["uv"])
    PU.assert_depth_status_equal(d, s, G["depth_noplane"], G["status_noplane"], "golden no plane")
    d, s = est.CalculateDepth(G["uv"], GroundPlane(G["plane_coeffs"], G["plane_inliers"]))
    PU.assert_depth_status_equal(d, s, G["depth_plane"], G["status_plane"], "golden plane")
    pl = est.estimateGroundPlane(G["cloud"], 77)
    assert pl.iterations == int(G["ransac_iterations"]) and np.array_equal(pl.getInlinersIndex(), G["ransac_inliers"])
    assert np.allclose(pl.getModelCoeffs(), G["ransac_coeffs"], rtol=1e-6, atol=1e-7)
