#!/usr/bin/env python
"""Regenerates tests/golden/golden_small.npz: a small end-to-end fixture (inputs + the CPU oracle's outputs).

The reference ships no end-to-end golden data and cannot be built here, so the fixture is
produced by the oracle restatement; it freezes the oracle's behaviour (any later change to oracle/ or to the
synthetic generator shows up as a diff) and gives the GPU tests inputs/outputs that do not depend on the
oracle library being rebuilt. Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle_lib as O  # noqa: E402
import parity_util as PU  # noqa: E402
from mono_lidar_depth_b200 import synth  # noqa: E402

W, H, F_, CX, CY = 320, 240, 300.0, 160.3, 119.6


def main():
    rng = np.random.RandomState(2026)
    T = synth.KITTI_T_LIDAR_TO_CAM
    cloud = PU.random_scene_cloud(rng, 5000, W, H, F_, CX, CY, T, dense_patches=25)
    # a road: dense returns on the lidar-frame plane z = -1.6 in front of the sensor (x forward, y left)
    gx, gy = np.meshgrid(np.arange(4.0, 30.0, 0.25), np.arange(-6.0, 6.0, 0.2))
    ground = np.stack([gx.ravel(), gy.ravel(), -1.6 + rng.normal(0, 0.01, gx.size), np.zeros(gx.size)], 1).astype(np.float32)
    cloud = np.concatenate([cloud, ground], 0)
    rng.shuffle(cloud)
    uv = np.stack([rng.uniform(-5, W + 5, 800), rng.uniform(-5, H + 5, 800)], 1)
    uv[400:, 1] = rng.uniform(H * 0.55, H, 400)  # half of the features on the road
    uv[:100] = np.floor(uv[:100])  # integer pixels like the real caller's features
    p = O.yaml_params()
    o = O.Oracle(p)
    o.initialize(W, H, F_, CX, CY, T)
    o.set_cloud(cloud)
    d0, s0 = o.calculate_depth(uv)  # plane nullptr
    # road path: a synthetic ground plane in the lidar frame and every third near-plane point as inlier
    coeffs = np.array([0.0, 0.0, 1.0, 1.6], np.float32)
    dist = np.abs(cloud[:, 2] + 1.6)
    inl = np.nonzero(dist < 0.05)[0][::2].astype(np.int32)
    d1, s1 = o.calculate_depth(uv, (coeffs, inl))
    rc, rcoef, rinl, rit = O.ransac_plane(p, cloud, 77)
    np.savez_compressed(Path(__file__).with_name("golden_small.npz"), cloud=cloud, uv=uv, pixel_map=o.pixel_map_raw(),
                        point_index=o.point_index(), depth_noplane=d0, status_noplane=s0, plane_coeffs=coeffs, plane_inliers=inl,
                        depth_plane=d1, status_plane=s1, ransac_rc=rc, ransac_coeffs=rcoef, ransac_inliers=rinl, ransac_iterations=rit,
                        camera=np.array([W, H, F_, CX, CY]), T=T)
    print("status mix (no plane):", np.bincount(s0, minlength=17), "with plane:", np.bincount(s1, minlength=17))


if __name__ == "__main__":
    main()
