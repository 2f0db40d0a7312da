"""CPU restatement (numpy) of Mono_Lidar::SemanticPlane::CalculateInliersPlane, RansacPlane.cpp:159-274.

TEST INFRASTRUCTURE ONLY (oracle/): the checker of mld_semantic_ground_plane. Pinned against the reference's own code
(oracle/_ref, ref_semantic_plane) by tests/test_semantic_plane.py: ground-labelled set and inlier set bit-identical,
coefficients to float rounding (the only difference is the 3x3 eigen solver: numpy's eigh here, Jacobi in the stand-in).

PCL pieces restated (PCL 1.8, un-vendored upstream): transformPointCloud in the transform's scalar (double) rounded to
float; SampleConsensusModelPlane::optimizeModelCoefficients = computeMeanAndCovarianceMatrix with sequential FLOAT
accumulators over the finite inliers + eigenvector of the smallest eigenvalue, d = -n.centroid, input model returned
when there are not more than 3 inliers; selectWithinDistance = |a x + b y + c z + d| (float, left to right) < threshold
over every point of the cloud.
"""
import numpy as np


class PclInvalid(Exception):
    """GroundPlane::ExceptionPclInvalid (RansacPlane.h:48-52)."""


def _fit(cloud_xyz, idx, model):
    """optimizeModelCoefficients(idx, model): float accumulation in index order."""
    if len(idx) <= 3:
        return model.copy()
    p = cloud_xyz[idx]
    p = p[np.isfinite(p).all(axis=1)]
    if len(p) == 0:
        return model.copy()
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    cols = [x * x, x * y, x * z, y * y, y * z, z * z, x, y, z]  # float32 products, like accu[k] += p.a * p.b
    accu = np.array([np.cumsum(c, dtype=np.float32)[-1] for c in cols], np.float32)
    accu = accu / np.float32(len(p))
    cx, cy, cz = accu[6], accu[7], accu[8]
    cov = np.empty((3, 3), np.float32)
    cov[0, 0] = accu[0] - cx * cx
    cov[0, 1] = cov[1, 0] = accu[1] - cx * cy
    cov[0, 2] = cov[2, 0] = accu[2] - cx * cz
    cov[1, 1] = accu[3] - cy * cy
    cov[1, 2] = cov[2, 1] = accu[4] - cy * cz
    cov[2, 2] = accu[5] - cz * cz
    w, v = np.linalg.eigh(cov.astype(np.float64))
    n = v[:, 0].astype(np.float32)
    d = np.float32(-1.0) * (n[0] * cx + n[1] * cy + n[2] * cz)
    return np.array([n[0], n[1], n[2], d], np.float32)


def ground_labelled(cloud, labels, f, cu, cv, T_cam_lidar, ground_labels):
    """Indices of the points whose projection carries a ground label (RansacPlane.cpp:197-222)."""
    xyz = np.ascontiguousarray(cloud, np.float32)[:, :3]
    T = np.asarray(T_cam_lidar, np.float64)[:3, :4]
    H, W = labels.shape
    with np.errstate(all="ignore"):
        p = xyz.astype(np.float64)
        t = np.empty_like(p)
        for i in range(3):  # ((t0 x + t1 y) + t2 z) + t3 in double, then float
            t[:, i] = ((T[i, 0] * p[:, 0] + T[i, 1] * p[:, 1]) + T[i, 2] * p[:, 2]) + T[i, 3]
        t = t.astype(np.float32).astype(np.float64)
        q0 = (f * t[:, 0] + 0.0 * t[:, 1]) + cu * t[:, 2]
        q1 = (0.0 * t[:, 0] + f * t[:, 1]) + cv * t[:, 2]
        q2 = (0.0 * t[:, 0] + 0.0 * t[:, 1]) + 1.0 * t[:, 2]
        u, v = q0 / q2, q1 / q2
        ok = (np.abs(u) < 2147483648.0) & (np.abs(v) < 2147483648.0)  # cvttsd2si: NaN / overflow -> INT_MIN -> invalid
        px = np.where(ok, np.trunc(np.where(ok, u, 0.0)), -1).astype(np.int64)
        py = np.where(ok, np.trunc(np.where(ok, v, 0.0)), -1).astype(np.int64)
    # the reference accepts x == cols / y == rows and then reads out of bounds (undefined); treated as not ground
    inside = ok & (px >= 0) & (px < W) & (py >= 0) & (py < H)
    lab = np.zeros(len(xyz), np.int64)
    lab[inside] = labels[py[inside], px[inside]]
    return np.nonzero(inside & np.isin(lab, np.asarray(list(ground_labels), np.int64)))[0].astype(np.int32)


def semantic_plane(cloud, labels, f, cu, cv, T_cam_lidar, ground_labels, inlier_threshold):
    """Returns (coeffs float32[4], inlier indices int32[]) or raises PclInvalid."""
    xyz = np.ascontiguousarray(cloud, np.float32)[:, :3]
    kept = ground_labelled(cloud, np.ascontiguousarray(labels, np.uint8), f, cu, cv, T_cam_lidar, ground_labels)
    if len(kept) < 3:
        raise PclInvalid()
    model = _fit(xyz, kept, np.array([0, 0, 1, 0], np.float32))
    with np.errstate(all="ignore"):
        dist = np.abs(((model[0] * xyz[:, 0] + model[1] * xyz[:, 1]) + model[2] * xyz[:, 2]) + model[3])  # float32
        inl = np.nonzero(dist.astype(np.float64) < inlier_threshold)[0].astype(np.int32)
    return _fit(xyz, inl, model), inl, kept, model
