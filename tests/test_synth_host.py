"""Host side of the synthetic generator (no GPU needed)."""
import numpy as np

from mono_lidar_depth_b200 import synth


def test_shapes_and_determinism():
    cfg = synth.default_config()
    assert synth.points_per_frame(cfg) == 120000
    a = synth.points_host(cfg, 7, 3)
    b = synth.points_host(cfg, 7, 3)
    c = synth.points_host(cfg, 7, 4)
    assert a.shape == (120000, 4) and a.dtype == np.float32
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert not np.array_equal(a.view(np.uint32), c.view(np.uint32))
    nan = np.isnan(a[:, 0])
    assert 0.01 < nan.mean() < 0.2  # dropouts + no-returns
    r = np.linalg.norm(a[~nan, :3], axis=1)
    assert r.min() > 0.3 and r.max() < 121.0
    assert np.isclose(np.median(a[~nan, 2]), -1.73, atol=0.2)  # most returns come from the ground
    dcfg = synth.default_config(dense=True)
    assert synth.points_per_frame(dcfg) == 260096


def test_features_are_integer_pixels_inside_the_image():
    cfg = synth.default_config()
    uv = synth.features_host(cfg, 1, 0, 2000)
    assert uv.shape == (2000, 2)
    assert np.array_equal(uv, np.floor(uv))
    assert uv[:, 0].min() >= 0 and uv[:, 0].max() < 1241 and uv[:, 1].min() >= 0 and uv[:, 1].max() < 376
    band = (uv[:, 1] >= int(0.4 * 376)).mean()
    assert 0.3 < band < 0.55  # 58 % of the features sit above the lidar-covered band (include/mld_synth.h)


def test_road_mix_and_pointxyzi_layout():
    cfg = synth.default_config(road=True)
    uv = synth.features_host(cfg, 1, 0, 4000)
    lower_third = (uv[:, 1] >= (2 * 376) // 3).mean()
    assert 0.5 < lower_third < 0.7  # half of the features by construction + what the other classes put there
    k = synth.default_config()
    a = synth.points_host(k, 7, 3)
    b = synth.points_host_xyzi32(k, 7, 3)
    assert b.shape == (120000, 8)
    assert np.array_equal(a[:, :3].view(np.uint32), b[:, :3].view(np.uint32)) and np.array_equal(a[:, 3].view(np.uint32), b[:, 4].view(np.uint32))
