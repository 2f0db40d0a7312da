#!/usr/bin/env python
"""Regenerates tests/golden/ref_golden.npz: outputs of the REFERENCE ITSELF (oracle/_ref/libmld_ref.so = the
reference's own monolidar_fusion sources compiled on the stand-in headers of oracle/ref_standin) on seeded inputs.

/root/reference does not exist on the GPU box, so these fixtures are how reference outputs travel: the GPU tests
(tests/test_ref_golden.py) compare the CUDA path with them directly, without the oracle in between.
Run from the repo root in a container that has /root/reference:  make -C oracle && python tests/golden/make_ref_golden.py

Cases
  small_*  320x240 random scene with a road (the inputs of golden_small.npz), yaml parameters, plane nullptr and an
           injected plane (M-estimator road path)
  kitti_*  two KITTI-shaped synthetic frames (inputs are regenerated from the seed by mono_lidar_depth_b200.synth, only
           the reference's outputs are stored), plane nullptr; frame 1 also with an injected plane
  sem*_    SemanticPlane coefficients and inlier indices for three label images over KITTI-shaped sweeps
  var_*    the 256x192 scene of the parameter-variant tests under every variant of parity_util.VARIANTS except pca
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle_lib as O  # noqa: E402
import parity_util as PU  # noqa: E402
import ref_lib as R  # noqa: E402
from mono_lidar_depth_b200 import synth  # noqa: E402

KCAM = (1241, 376, 718.856, 607.1928, 185.2157)
KITTI_SEED, KITTI_FRAMES, KITTI_F = 4242, (0, 1), 2000
VAR_CAM = (256, 192, 250.0, 128.0, 96.0)


def variant_inputs():
    rng = np.random.RandomState(42)
    W, H, f, cx, cy = VAR_CAM
    cloud = PU.random_scene_cloud(rng, 10000, W, H, f, cx, cy, synth.KITTI_T_LIDAR_TO_CAM, dense_patches=45)
    uv = np.stack([rng.uniform(0, W, 2000), rng.uniform(0, H, 2000)], 1)
    return cloud, uv


def kitti_plane(cloud):
    coeffs = np.array([0.0, 0.0, 1.0, 1.73], np.float32)
    dist = np.abs(cloud[:, 2] + 1.73)
    return coeffs, np.nonzero(np.isfinite(dist) & (dist < 0.1))[0].astype(np.int32)


def semantic_case(case):
    """(cloud, label image, ground labels, inlier threshold) of the SemanticPlane fixtures."""
    cfg = synth.default_config()
    cloud = synth.points_host(cfg, 77, case)
    W, H = 1241, 376
    lab = np.zeros((H, W), np.uint8)
    rng = np.random.RandomState(100 + case)
    lab[200 + 10 * case:, :] = 7                      # road
    lab[300:, 400:800] = 9                            # another ground class
    lab[250:280, 100:300] = 2                         # a non-ground object on the road
    lab[:40, :] = 6                                   # ground label in the sky: only stray / behind-camera projections
    noise = rng.rand(H, W) < 0.01
    lab[noise] = rng.randint(0, 12, int(noise.sum())).astype(np.uint8)
    gl = [6, 7, 8, 9] if case != 2 else [7, 300, -1]  # labels outside 0..255 never match a pixel
    thr = (0.1, 0.25, 0.05)[case]
    return cloud, lab, gl, thr


def semantic_boundary_cloud(m=60000):
    """(cloud, label image): lidar points whose projection sits on every decision boundary of the SemanticPlane label test --
    integer pixel coordinates +- 1e-7, the image frame, the camera plane, points behind the camera, huge / non-finite values.
    Everything is ground except a lattice, so an off-by-one pixel changes the kept set."""
    F_, CU, CV = KCAM[2], KCAM[3], KCAM[4]
    rng = np.random.RandomState(12)
    labels = np.full((376, 1241), 7, np.uint8)
    labels[::2, ::3] = 1
    u = np.concatenate([rng.randint(-3, 1245, m // 2) + rng.choice([0.0, 1e-7, -1e-7, 1e-4, -1e-4, 0.5], m // 2), rng.uniform(-5, 1246, m // 2)])
    v = np.concatenate([rng.randint(-3, 380, m // 2) + rng.choice([0.0, 1e-7, -1e-7, 1e-4, -1e-4, 0.5], m // 2), rng.uniform(-5, 381, m // 2)])
    z = rng.choice([0.26, 0.24, 0.01, 1e-4, 1.0, 7.3, 55.0, 400.0, -0.3, -2.0, -40.0], m) * rng.uniform(0.9, 1.1, m)
    camp = np.stack([(u - CU) / F_ * z, (v - CV) / F_ * z, z], 1)
    Tm = np.vstack([synth.KITTI_T_LIDAR_TO_CAM[:3], [0, 0, 0, 1]])
    Ti = np.linalg.inv(Tm)
    lid = (Ti[:3, :3] @ camp.T).T + Ti[:3, 3]
    cloud = np.zeros((m + 8, 4), np.float32)
    cloud[:m, :3] = lid.astype(np.float32)
    cloud[m:, :3] = [[np.nan, 1, 1], [np.inf, 1, 1], [1, -np.inf, 2], [3e38, 0, 0], [1e6, 2e5, -3e4], [0, 0, 0], [1e-30, 0, 0], [5, np.nan, np.nan]]
    return cloud, labels


def main():
    assert R.available(), "build oracle/_ref first (make -C oracle in a container with /root/reference)"
    T = synth.KITTI_T_LIDAR_TO_CAM
    out = {}
    # ---- small scene: same inputs as golden_small.npz ----
    G = np.load(Path(__file__).with_name("golden_small.npz"))
    cam = tuple(G["camera"])
    cam = (int(cam[0]), int(cam[1]), float(cam[2]), float(cam[3]), float(cam[4]))
    p = O.yaml_params()
    r = R.Reference(p)
    r.initialize(*cam, G["T"])
    r.set_cloud(G["cloud"], (G["plane_coeffs"], G["plane_inliers"]))
    out["small_pixel_map"] = r.pixel_map_raw()
    out["small_point_index"] = r.point_index()
    out["small_depth_noplane"], out["small_status_noplane"] = r.calculate_depth(G["uv"], with_plane=False)
    out["small_depth_plane"], out["small_status_plane"] = r.calculate_depth(G["uv"], with_plane=True)
    nb = [r.neighbors(float(u), float(v), sw, sh) for u, v in G["uv"][:64] for sw, sh in ((1.0, 1.0), (2.0, 1.5))]
    out["small_neighbors_flat"] = np.concatenate(nb) if nb else np.zeros(0, np.int32)
    out["small_neighbors_len"] = np.array([len(x) for x in nb], np.int32)
    # ---- KITTI-shaped frames ----
    cfg = synth.default_config()
    r = R.Reference(p)
    r.initialize(*KCAM, T)
    for fr in KITTI_FRAMES:
        cloud = synth.points_host(cfg, KITTI_SEED, fr)
        uv = synth.features_host(cfg, KITTI_SEED, fr, KITTI_F)
        plane = kitti_plane(cloud)
        r.set_cloud(cloud, plane)
        out[f"kitti{fr}_pixel_map"] = r.pixel_map_raw()
        out[f"kitti{fr}_point_index"] = r.point_index()
        out[f"kitti{fr}_depth_noplane"], out[f"kitti{fr}_status_noplane"] = r.calculate_depth(uv, with_plane=False)
        if fr == 1:
            rng = np.random.RandomState(1)
            uvr = np.stack([rng.randint(0, 1241, 3000), rng.randint(180, 376, 3000)], 1).astype(np.float64)
            out["kitti1_uv_road"] = uvr
            out["kitti1_depth_plane"], out["kitti1_status_plane"] = r.calculate_depth(uvr, with_plane=True)
    # ---- parameter variants ----
    cloud, uv = variant_inputs()
    for v in PU.VARIANTS:
        if v == "pca":
            continue
        r = R.Reference(PU.variant_params(v))
        r.initialize(*VAR_CAM, T)
        r.set_cloud(cloud)
        out[f"var_{v}_depth"], out[f"var_{v}_status"] = r.calculate_depth(uv)
    # ---- SemanticPlane (RansacPlane.cpp:159-274) run by the reference's own code ----
    for case in (0, 1, 2):
        cloud, lab, gl, thr = semantic_case(case)
        rc, c, inl = R.semantic_plane(lab, KCAM[2], KCAM[3], KCAM[4], T, gl, thr, cloud)
        assert rc == 0
        out[f"sem{case}_coeffs"], out[f"sem{case}_inliers"] = c, inl
        print("semantic case", case, c, len(inl))
    np.savez_compressed(Path(__file__).with_name("ref_golden.npz"), **out)
    for k in ("small_status_noplane", "small_status_plane", "kitti0_status_noplane", "kitti1_status_plane"):
        print(k, np.bincount(out[k], minlength=17))
    print("written", Path(__file__).with_name("ref_golden.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
