// mld_geometry.cuh -- scalar FP64 geometry of the per-feature tail, shared by the warp-per-feature
// and thread-per-feature kernels. Each function restates one reference routine (file:line in the
// comment above it); evaluation order follows Appendix A of SURVEY.md so that results are bit-identical
// to the CPU oracle (no FMA contraction: the translation units are compiled with -fmad=false).
#pragma once
#include "mld_common.cuh"

struct Plane {
    D3 n;
    double off;
};

// A8 PlaneEstimationCheckPlanar::CheckPlanar
static __device__ bool check_planar(const D3& c1, const D3& c2, const D3& c3, double treshold) {
    D3 e1 = normalized3(c2 - c1), e2 = normalized3(c3 - c1), e3 = normalized3(c3 - c2);
    double l12 = norm3(cross3(e1, e2)), l13 = norm3(cross3(e1, e3)), l23 = norm3(cross3(e2, e3));
    return (l12 >= treshold) && (l13 >= treshold) && (l23 >= treshold);
}

// Eigen::Hyperplane<double,3>::Through(p0,p1,p2)
static __device__ Plane plane_through(const D3& p0, const D3& p1, const D3& p2) {
    D3 v0 = p2 - p0, v1 = p1 - p0;
    D3 n = cross3(v0, v1);
    double nn = norm3(n);
    if (nn <= norm3(v0) * norm3(v1) * 2.220446049250313e-16) {
        // degenerate: null direction of [v0; v1] (Eigen: 2x3 JacobiSVD, column 2 of V)
        double w[3];
        D3 ev[3];
        eig3_sym_regs(v0.x * v0.x + v1.x * v1.x, v0.x * v0.y + v1.x * v1.y, v0.x * v0.z + v1.x * v1.z,
                      v0.y * v0.y + v1.y * v1.y, v0.y * v0.z + v1.y * v1.z, v0.z * v0.z + v1.z * v1.z, w, ev);
        int b = 0;
        if (w[1] < w[b]) b = 1;
        if (w[2] < w[b]) b = 2;
        n = (b == 0) ? ev[0] : (b == 1 ? ev[1] : ev[2]);
    } else {
        n = n / nn;
    }
    return Plane{n, -dot3(p0, n)};
}

// A10 LinePlaneIntersection{Normal,OrthogonalTreshold}::GetIntersection
static __device__ bool line_plane(const Plane& pl, const D3& n0, const D3& n1, double ortho_treshold, double& depth) {
    D3 dir = normalized3(n1 - n0);  // ParametrizedLine::Through
    if (ortho_treshold > 0) {
        D3 lineNormal = normalized3(n1);
        D3 planeNormal = normalized3(pl.n);
        if (!(fabs(dot3(planeNormal, lineNormal)) >= ortho_treshold)) return false;
    }
    double t = -(pl.off + dot3(pl.n, n0)) / dot3(pl.n, dir);
    D3 pt = n0 + dir * t;
    depth = pt.z;
    return true;
}

// A9 CameraPinhole::getViewingRays (+ the caller's z flip, DepthEstimator.cpp:938-939)
static __device__ D3 viewing_ray(const DevParams& P, double u, double v) {
    D3 d = D3{(P.Kinv[0] * u + P.Kinv[1] * v) + P.Kinv[2] * 1.0, (P.Kinv[3] * u + P.Kinv[4] * v) + P.Kinv[5] * 1.0,
              (P.Kinv[6] * u + P.Kinv[7] * v) + P.Kinv[8] * 1.0};
    d = normalized3(d);
    if (d.z < 0) d = d * -1.0;
    return d;
}

// A11 thresholds; returns 0 or the failing status, may clamp depth in Adjust mode
static __device__ int apply_tresholds(const DevParams& P, double& depth, double minZ, double maxZ) {
    if (P.glob_en) {  // TresholdDepthGlobal::CheckInDepth
        if (depth < P.glob_min) {
            if (P.glob_mode == 0) return ST_TresholdDepthGlobalSmallerMin;
            depth = P.glob_min;
        } else if (depth > P.glob_max) {
            if (P.glob_mode == 0) return ST_TresholdDepthGlobalGreaterMax;
            depth = P.glob_max;
        }
    }
    if (P.loc_en) {  // TresholdDepthLocal::CheckInBounds
        double depthInterval = maxZ - minZ;
        double lo, hi;
        if (P.loc_type == 1) {
            double r = depthInterval * P.loc_val;
            lo = minZ - r;
            hi = maxZ + r;
        } else {
            lo = minZ - P.loc_val;
            hi = maxZ + P.loc_val;
        }
        if (depth < lo) {
            if (P.loc_mode == 0) return ST_TresholdDepthLocalSmallerMin;
            depth = lo;
        } else if (depth > hi) {
            if (P.loc_mode == 0) return ST_TresholdDepthLocalGreaterMax;
            depth = hi;
        }
    }
    return 0;
}

