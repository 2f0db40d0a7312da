"""Pins the CPU oracle against the golden vectors / properties of the reference's own tests
(/root/reference/monolidar_fusion/test/test_monolidar_fusion.cpp). CPU only."""
import ctypes as C

import numpy as np
import pytest

import kat_data
import oracle_lib as O


def test_histogram_filter_points_min_dist_blob_golden():
    """Histogram.FilterPointsMinDistBlob (:306-374): exact output {8.2, 8.3, 8.4}."""
    ok, pos, lo, hi = O.histogram_filter(kat_data.HIST_DEPTHS, kat_data.HIST_BIN_WIDTH, kat_data.HIST_MIN_COUNT)
    assert ok
    out = [kat_data.HIST_DEPTHS[i] for i in pos]
    assert out == kat_data.HIST_EXPECTED
    for z in out:
        assert lo <= z <= hi
    assert (lo, hi) == (8.0, 9.0)


def test_histogram_edge_cases():
    # no bin reaches the minimum count before an empty bin follows an occupied one -> fail (:82-84)
    ok, pos, _, _ = O.histogram_filter([1.1, 3.2, 3.3, 3.4], 1.0, 3)
    assert not ok
    # first local maximum wins, a later bigger blob is ignored (:75-80)
    ok, pos, lo, hi = O.histogram_filter([2.1, 2.2, 2.3, 3.5, 5.1, 5.2, 5.3, 5.4], 1.0, 3)
    assert ok and list(pos) == [0, 1, 2] and (lo, hi) == (2.0, 3.0)
    # maxDist = 0 -> binCount 1 -> fail (:53)
    ok, *_ = O.histogram_filter([], 1.0, 3)
    assert not ok
    # depth capped into the last bin (Histogram.cpp:29-30) does not crash
    ok, pos, _, _ = O.histogram_filter([998.5, 998.6, 998.7], 0.5, 3)
    assert ok and len(pos) == 3


def test_neighbor_finder_find_by_pixel_property():
    """NeigborFinder.findByPixel (:82-171): every neighbour lies inside the search rectangle and
    re-projects onto its own pixel."""
    rng = np.random.RandomState(0)
    W = H = 100
    f, cu, cv = 600.0, 50.0, 50.0
    sw, sh = 3, 5
    n = 50
    img = np.stack([rng.randint(0, 10, n), rng.randint(0, 10, n)], 1).astype(np.float64)
    cam = np.zeros((n, 3))
    proj = np.zeros((n, 2))
    for i in range(n):
        d = np.zeros(3)
        O.lib().orc_viewing_ray(W, H, f, cu, cv, img[i, 0], img[i, 1], d.ctypes.data)
        assert abs(np.linalg.norm(d) - 1.0) < 1e-12
        cam[i] = float(rng.randint(1, 11)) * d
        uv = np.zeros(2)
        O.lib().orc_image_point(W, H, f, cu, cv, cam[i].ctypes.data, uv.ctypes.data)
        proj[i] = uv
    total = 0
    for i in range(n):
        out = np.empty(256, np.int32)
        k = O.lib().orc_neighbor_finder(W, H, sw, sh, img.ctypes.data, cam.ctypes.data, n, img[i, 0], img[i, 1], out.ctypes.data, 256)
        assert k >= 1  # the pixel of the feature itself holds a point
        for idx in out[:k]:
            assert np.linalg.norm(proj[idx] - img[idx]) < 0.01
            assert abs(img[idx, 0] - img[i, 0]) <= np.ceil(sw * 0.5) + 0.01
            assert abs(img[idx, 1] - img[i, 1]) <= np.ceil(sh * 0.5) + 0.01
        total += k
    assert total > n


def test_first_point_wins_pixel_map():
    """NeighborFinderPixel::InitializeLidarProjection (NeighborFinderPixel.cpp:40-55): the first visible
    point in cloud order with z > 0 owns the pixel; a later, nearer point does not replace it."""
    p = O.yaml_params()
    o = O.Oracle(p)
    T = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], np.float64)
    o.initialize(64, 48, 100.0, 32.0, 24.0, T)
    cloud = np.array(
        [
            [0.0, 0.0, -5.0, 0],   # behind the camera: projects to (32,24) and is "visible" but never enters the map
            [0.0, 0.0, 10.0, 0],   # first in-front point on pixel (32,24)
            [0.0, 0.0, 2.0, 0],    # nearer, later -> loses
            [np.nan, 0.0, 1.0, 0], # NaN never visible
            [1.0, 0.0, 10.0, 0],   # pixel (42,24)
        ],
        np.float32,
    )
    o.set_cloud(cloud)
    assert list(o.point_index()) == [0, 1, 2, 4]
    m = o.pixel_map_raw()
    assert m[24, 32] == 1 and m[24, 42] == 4
    assert (m >= 0).sum() == 2
    mv = o.pixel_map_visible()
    assert mv[24, 32] == 1 and mv[24, 42] == 3  # visible numbering counts the behind-camera point


def test_ransac_plane_reference_kat():
    """RansacPlane.CalculateInlersPlane (:376-441): coefficients within +-0.2 of n=(0,0,1), d=1.6."""
    cloud = kat_data.ransac_kat_cloud()
    p = O.default_params()
    p.ransac_plane_distance_treshold = 0.2
    p.ransac_plane_max_iterations = 600
    p.ransac_plane_use_refinement = 1
    p.ransac_plane_refinement_treshold = 0.05
    p.ransac_plane_probability = 0.99
    for seed in (0, 1, 2, 1234):
        rc, coeffs, inl, iters = O.ransac_plane(p, cloud, seed)
        assert rc == 0
        sign = 1.0 if coeffs[2] > 0 else -1.0
        assert abs(coeffs[0] - 0.0) < 0.2 and abs(coeffs[1] - 0.0) < 0.2
        assert abs(coeffs[2] - sign * 1.0) < 0.2 and abs(coeffs[3] - sign * 1.6) < 0.2
        assert 1 <= iters <= 601
        assert len(inl) > 0 and np.all(np.diff(inl) > 0)  # order-preserving subsample


def test_ransac_too_few_points_is_pcl_invalid():
    p = O.default_params()
    rc, *_ = O.ransac_plane(p, np.zeros((2, 4), np.float32), 0)
    assert rc == -1


def test_status_precedence_and_region_growing():
    p = O.yaml_params()
    p.do_use_depth_segmentation = 1
    o = O.Oracle(p)
    o.initialize(64, 48, 100.0, 32.0, 24.0, np.eye(4)[:3])
    o.set_cloud(np.array([[0, 0, 5, 0]], np.float32))
    with pytest.raises(RuntimeError):
        o.calculate_depth(np.array([[32.0, 24.0]]))
    p.do_use_depth_segmentation = 0
    p.set_all_depths_to_zero = 1
    o = O.Oracle(p)
    o.initialize(64, 48, 100.0, 32.0, 24.0, np.eye(4)[:3])
    o.set_cloud(np.array([[0, 0, 5, 0]], np.float32))
    d, s = o.calculate_depth(np.array([[32.0, 24.0], [1.0, 1.0]]))
    assert list(s) == [1, 1] and list(d) == [-1, -1]  # DepthEstimator.cpp:448-453


def test_planar_wall_depth_matches_geometry():
    """A fronto-parallel wall at z = 7 m sampled densely: every feature on it must come out as Success
    with depth 7 (ray/plane intersection returns the z of the hit point, LinePlaneIntersectionNormal.cpp:28)."""
    p = O.yaml_params()
    p.do_use_ransac_plane = 0
    o = O.Oracle(p)
    W, H, f = 128, 96, 200.0
    o.initialize(W, H, f, 64.0, 48.0, np.eye(4)[:3])
    us, vs = np.meshgrid(np.arange(2, W - 2, 2) + 0.5, np.arange(2, H - 2, 3) + 0.5)
    z = 7.0
    x = (us.ravel() - 64.0) / f * z
    y = (vs.ravel() - 48.0) / f * z
    cloud = np.stack([x, y, np.full_like(x, z), np.zeros_like(x)], 1).astype(np.float32)
    o.set_cloud(cloud)
    uv = np.array([[40.0, 40.0], [64.0, 48.0], [90.0, 60.0]])
    d, s = o.calculate_depth(uv)
    assert list(s) == [1, 1, 1]
    np.testing.assert_allclose(d, 7.0, rtol=1e-6)


def test_passthrough_drops_points_with_non_finite_x_or_y():
    """pcl::PassThrough::applyFilterIndices removes every point with a non-finite x, y or z before it tests the field
    (RansacPlane.cpp:58-64): a point with NaN x and a z inside the limits never becomes a candidate or an inlier."""
    p = O.yaml_params()
    p.ransac_plane_min_z = -3.0
    p.ransac_plane_max_z = 0.0
    rng = np.random.RandomState(2)
    cloud = np.zeros((400, 4), np.float32)
    cloud[:, :2] = rng.uniform(-10, 10, (400, 2))
    cloud[:, 2] = -1.7 + rng.normal(0, 0.01, 400)
    cloud[::7, 0] = np.nan
    cloud[3::11, 1] = np.inf
    rc, coeffs, inl, _ = O.ransac_plane(p, cloud, 5)
    assert rc == 0 and np.isfinite(coeffs).all()
    bad = ~np.isfinite(cloud[:, :3]).all(axis=1)
    assert not bad[inl].any() and len(inl) == (~bad).sum()
