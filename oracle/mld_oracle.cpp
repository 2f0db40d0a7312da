/*
 * mld_oracle.cpp -- CPU parity oracle (TEST INFRASTRUCTURE, never shipped, never on the product path).
 *
 * Dependency-free restatement of Mono_Lidar::DepthEstimator's per-frame hot path. Citations are
 * relative to /root/reference (MF = monolidar_fusion). Built with
 *     g++ -O2 -std=c++17 -ffp-contract=off -fopenmp -shared -fPIC
 * (-ffp-contract=off: no FMA contraction, so that every (int) cast and threshold compare sees the
 * same IEEE double value as a plain SSE2 build of the reference and as the CUDA kernels, which are
 * compiled with -fmad=false).
 *
 * Eigen / PCL expressions are restated by hand (those libraries are not vendored in the reference
 * and are absent here; versions are unpinned upstream -- Eigen >= 3.3 semantics are assumed:
 * normalize() leaves a zero vector untouched, Hyperplane::Through normalises the cross product,
 * 3-term reductions are evaluated left to right).
 *
 * Parity status is stated per function: "pinned" = checked against a golden vector / property of
 * the reference's own tests in tests/test_oracle_kat.py; "unpinned" = restatement only.
 */
#include "mld_oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ----------------------------------------------------------------------------------------------
// tiny 3-vector toolkit (Eigen::Vector3d restated; all reductions left-to-right)
// ----------------------------------------------------------------------------------------------
struct V3 {
    double x, y, z;
};
inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator*(const V3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(const V3& a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline double sqnorm(const V3& a) { return dot(a, a); }
inline double norm(const V3& a) { return std::sqrt(sqnorm(a)); }
inline V3 cross(const V3& a, const V3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen 3.3 MatrixBase::normalized(): z = squaredNorm(); z > 0 ? v / sqrt(z) : v
inline V3 normalized(const V3& a) {
    double z = sqnorm(a);
    if (z > 0) return a / std::sqrt(z);
    return a;
}

// Eigen compute_inverse<Matrix3d>: cofactor expansion, result(i,j) = cofactor(j,i) * (1/det).
// Used by CameraPinhole::getViewingRays (camera_pinhole.h:65) and Affine3d::inverse()
// (DepthEstimator.cpp:44). m and out are row-major 3x3.
inline double cof3(const double* m, int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
void inverse3(const double* m, double* out) {
    double c0 = cof3(m, 0, 0), c1 = cof3(m, 1, 0), c2 = cof3(m, 2, 0);
    double det = (c0 * m[0] + c1 * m[3]) + c2 * m[6];
    double invdet = 1.0 / det;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) out[i * 3 + j] = cof3(m, j, i) * invdet;
}

// Cyclic Jacobi eigen-decomposition of a symmetric 3x3 (row-major a). Eigenvalues ascending in w,
// eigenvectors in the columns of v (v[r*3+c]). Restates Eigen::SelfAdjointEigenSolver's contract
// (ascending eigenvalues, orthonormal eigenvectors), MF/src/PCA.cpp:52.
void eig3_sym(const double* a_in, double* w, double* v) {
    double a[9];
    std::memcpy(a, a_in, sizeof(a));
    for (int i = 0; i < 9; i++) v[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
        double diag = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
        if (!(off > 1e-32 * diag) || !(off > 0)) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double apq = a[p * 3 + q];
                if (apq == 0.0) continue;
                double app = a[p * 3 + p], aqq = a[q * 3 + q];
                double theta = (aqq - app) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                if (!std::isfinite(theta)) t = 0.0;
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; k++) {  // A <- A * J
                    double akp = a[k * 3 + p], akq = a[k * 3 + q];
                    a[k * 3 + p] = c * akp - s * akq;
                    a[k * 3 + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) {  // A <- J^T * A
                    double apk = a[p * 3 + k], aqk = a[q * 3 + k];
                    a[p * 3 + k] = c * apk - s * aqk;
                    a[q * 3 + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    double vkp = v[k * 3 + p], vkq = v[k * 3 + q];
                    v[k * 3 + p] = c * vkp - s * vkq;
                    v[k * 3 + q] = s * vkp + c * vkq;
                }
            }
    }
    int idx[3] = {0, 1, 2};
    double d[3] = {a[0], a[4], a[8]};
    std::sort(idx, idx + 3, [&](int i, int j) { return d[i] < d[j]; });
    double vv[9];
    for (int c = 0; c < 3; c++) {
        w[c] = d[idx[c]];
        for (int r = 0; r < 3; r++) vv[r * 3 + c] = v[r * 3 + idx[c]];
    }
    std::memcpy(v, vv, sizeof(vv));
}

// One-sided (Hestenes) Jacobi SVD of the k x 3 matrix whose rows are `rows`: returns the right
// singular vector of the smallest singular value == the last left-singular vector of the 3 x k
// matrix the reference hands to Eigen::JacobiSVD (MF/src/PlaneEstimationMEstimator.cpp:49-50).
// Sign is arbitrary (the ray/plane intersection is sign invariant).
V3 smallest_left_singular_vector(const std::vector<V3>& cols) {
    size_t k = cols.size();
    std::vector<double> A(k * 3);
    for (size_t i = 0; i < k; i++) {
        A[i * 3 + 0] = cols[i].x;
        A[i * 3 + 1] = cols[i].y;
        A[i * 3 + 2] = cols[i].z;
    }
    double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int sweep = 0; sweep < 60; sweep++) {
        bool rotated = false;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (size_t i = 0; i < k; i++) {
                    alpha += A[i * 3 + p] * A[i * 3 + p];
                    beta += A[i * 3 + q] * A[i * 3 + q];
                    gamma += A[i * 3 + p] * A[i * 3 + q];
                }
                if (!(std::fabs(gamma) > 1e-17 * std::sqrt(alpha * beta)) || gamma == 0.0) continue;
                rotated = true;
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                if (!std::isfinite(zeta)) t = 0.0;
                double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (size_t i = 0; i < k; i++) {
                    double ap = A[i * 3 + p], aq = A[i * 3 + q];
                    A[i * 3 + p] = c * ap - s * aq;
                    A[i * 3 + q] = s * ap + c * aq;
                }
                for (int i = 0; i < 3; i++) {
                    double vp = V[i * 3 + p], vq = V[i * 3 + q];
                    V[i * 3 + p] = c * vp - s * vq;
                    V[i * 3 + q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double best = std::numeric_limits<double>::infinity();
    int bi = 2;
    for (int c = 0; c < 3; c++) {
        double s2 = 0;
        for (size_t i = 0; i < k; i++) s2 += A[i * 3 + c] * A[i * 3 + c];
        if (s2 < best) {
            best = s2;
            bi = c;
        }
    }
    return {V[0 * 3 + bi], V[1 * 3 + bi], V[2 * 3 + bi]};
}

// MF/include/monolidar_fusion/eDepthResultType.h:9-31
enum Status {
    Unspecified = 0,
    Success = 1,
    RadiusSearchInsufficientPoints = 2,
    HistogramNoLocalMax = 3,
    TresholdDepthGlobalGreaterMax = 4,
    TresholdDepthGlobalSmallerMin = 5,
    TresholdDepthLocalGreaterMax = 6,
    TresholdDepthLocalSmallerMin = 7,
    TriangleNotPlanar = 8,
    TriangleNotPlanarInsufficientPoints = 9,
    CornerBehindCamera = 10,
    PlaneViewrayNotOrthogonal = 11,
    PcaIsPoint = 12,
    PcaIsLine = 13,
    PcaIsCubic = 14,
    InsufficientRoadPoints = 15,
    SuccessRoad = 16
};

struct Hyperplane {
    V3 n;
    double off;
};

// ----------------------------------------------------------------------------------------------
// A6  PointHistogram::FilterPointsMinDistBlob, MF/src/HistogramPointDepth.cpp:15-123 and
//     Histogram::AddElement, MF/src/Histogram.cpp:19-33.        PINNED (golden vector {8.2,8.3,8.4})
// ----------------------------------------------------------------------------------------------
bool histogram_filter(const std::vector<double>& depths, double binWitdh, int minimalMaximumSize,
                      std::vector<int>& keep, double& lowerBorder, double& higherBorder) {
    lowerBorder = -1;
    higherBorder = -1;
    keep.clear();
    int depthCount = (int)depths.size();
    int maxDist = 0;
    for (int i = 0; i < depthCount; i++)
        if (depths[i] > maxDist) maxDist = (int)std::ceil(depths[i]);  // :38-41
    int binCount = (int)((maxDist) / binWitdh + 1);                      // :43
    if (binCount <= 1) return false;                                     // :53
    std::vector<int> bins((size_t)binCount, 0);
    for (int i = 0; i < depthCount; i++) {
        double value = std::min(depths[i], 1e10);  // Histogram.cpp:29
        int binIndex = static_cast<int>(std::min(std::abs(value / binWitdh), static_cast<double>(bins.size()) - 1.));
        bins[(size_t)binIndex]++;
    }
    int binMaxId = -1, binMaxVal = -1, binValue = 0;
    for (int i = 0; i < binCount; i++) {  // :70-85
        float lastBinValue = (float)binValue;
        binValue = bins[(size_t)i];
        if ((binValue > binMaxVal) && (binValue >= minimalMaximumSize)) {
            binMaxVal = binValue;
            binMaxId = i;
        } else if (binValue < binMaxVal)
            break;
        if ((lastBinValue > 0) && (binValue == 0)) return false;
    }
    if (binMaxId < 0) return false;  // :95
    lowerBorder = binMaxId * binWitdh - 0.0f * binWitdh;   // :99
    higherBorder = (binMaxId)*binWitdh + 1.0f * binWitdh;  // :100
    for (int i = 0; i < depthCount; i++)
        if ((depths[i] >= lowerBorder) && (depths[i] < higherBorder)) keep.push_back(i);  // :115-120
    return true;
}

// ----------------------------------------------------------------------------------------------
// A7  PlaneEstimationCalcMaxSpanningTriangle::CalculatePlaneCorners,
//     MF/src/PlaneEstimationCalcMaxSpanningTriangle.cpp:37-145 (bool ctor => _distTreshold = 0, :11-13)
//     UNPINNED (no reference test)
// ----------------------------------------------------------------------------------------------
bool max_spanning_triangle(const std::vector<V3>& points, double distTreshold, int& ci, int& cj, int& ck) {
    int pointsCount = (int)points.size();
    if (pointsCount < 3) return false;
    int maxDist_i = -1, maxDist_j = -1;
    double maxdist = -1;
    for (int i = 0; i < pointsCount - 1; i++)
        for (int j = i + 1; j < pointsCount; j++) {
            double dist = sqnorm(points[(size_t)i] - points[(size_t)j]);
            if (dist > maxdist) {
                maxdist = dist;
                maxDist_i = i;
                maxDist_j = j;
            }
        }
    if (maxdist <= distTreshold) return false;
    double maxdist2 = -1;
    double maxDist_k = -1;
    for (int k = 0; k < pointsCount - 1; k++) {  // note: the last point is never eligible (:71)
        if (k == maxDist_i || k == maxDist_j) continue;
        double dist1 = sqnorm(points[(size_t)k] - points[(size_t)maxDist_i]);
        if (dist1 <= distTreshold) continue;
        double dist2 = sqnorm(points[(size_t)k] - points[(size_t)maxDist_j]);
        if (dist2 <= distTreshold) continue;
        double dist = dist1 + dist2;
        if (dist > maxdist2) {
            maxdist2 = dist;
            maxDist_k = k;
        }
    }
    if ((maxDist_i == -1) || (maxDist_j == -1) || (maxDist_k == -1)) return false;
    ci = maxDist_i;
    cj = maxDist_j;
    ck = (int)maxDist_k;
    return true;
}

// A8  PlaneEstimationCheckPlanar::CheckPlanar, MF/src/PlaneEstimationCheckPlanar.cpp:18-44. UNPINNED
bool check_planar(const V3& c1, const V3& c2, const V3& c3, double treshold) {
    V3 e1 = normalized(c2 - c1), e2 = normalized(c3 - c1), e3 = normalized(c3 - c2);
    double l12 = norm(cross(e1, e2)), l13 = norm(cross(e1, e3)), l23 = norm(cross(e2, e3));
    return (l12 >= treshold) && (l13 >= treshold) && (l23 >= treshold);
}

// Eigen::Hyperplane<double,3>::Through(p0,p1,p2) (Eigen 3.3 Geometry/Hyperplane.h): v0 = p2-p0,
// v1 = p1-p0, n = v0 x v1, normalised; degenerate (n ~ 0) falls back to the null direction of
// [v0;v1] (Eigen uses a 2x3 JacobiSVD; restated as the smallest eigenvector of v0 v0^T + v1 v1^T).
Hyperplane plane_through(const V3& p0, const V3& p1, const V3& p2) {
    V3 v0 = p2 - p0, v1 = p1 - p0;
    V3 n = cross(v0, v1);
    double nn = norm(n);
    if (nn <= norm(v0) * norm(v1) * DBL_EPSILON) {
        double m[9] = {v0.x * v0.x + v1.x * v1.x, v0.x * v0.y + v1.x * v1.y, v0.x * v0.z + v1.x * v1.z,
                       v0.x * v0.y + v1.x * v1.y, v0.y * v0.y + v1.y * v1.y, v0.y * v0.z + v1.y * v1.z,
                       v0.x * v0.z + v1.x * v1.z, v0.y * v0.z + v1.y * v1.z, v0.z * v0.z + v1.z * v1.z};
        double w[3], v[9];
        eig3_sym(m, w, v);
        n = {v[0], v[3], v[6]};
    } else {
        n = n / nn;
    }
    return {n, -dot(p0, n)};
}

// A10  LinePlaneIntersection{Normal,OrthogonalTreshold}::GetIntersection,
//      MF/src/LinePlaneIntersectionNormal.cpp:11-31, MF/src/LinePlaneIntersectionOrthogonalTreshold.cpp:16-48.
//      ParametrizedLine::Through(n0,n1) = (origin n0, direction (n1-n0).normalized());
//      intersectionParameter = -(off + n.origin)/(n.direction); depth = z of the point.  UNPINNED
bool line_plane(const Hyperplane& pl, const V3& n0, const V3& n1, double ortho_treshold, double& depth) {
    V3 dir = normalized(n1 - n0);
    if (ortho_treshold > 0) {
        V3 lineNormal = normalized(n1);
        V3 planeNormal = normalized(pl.n);
        if (!(std::fabs(dot(planeNormal, lineNormal)) >= ortho_treshold)) return false;
    }
    double t = -(pl.off + dot(pl.n, n0)) / dot(pl.n, dir);
    V3 pt = n0 + dir * t;
    depth = pt.z;
    return true;
}

// A11  TresholdDepthGlobal::CheckInDepth, MF/src/TresholdDepthGlobal.cpp:16-36 (0 in bounds, 1 <min, 2 >max)
int treshold_global(int mode, double minV, double maxV, double& depth) {
    if (depth < minV) {
        if (mode == 0) {
            depth = -1;
            return 1;
        }
        depth = minV;
    } else if (depth > maxV) {
        if (mode == 0) {
            depth = -1;
            return 2;
        }
        depth = maxV;
    }
    return 0;
}
// A11  TresholdDepthLocal::CheckInBounds, MF/src/TresholdDepthLocal.cpp:18-66
int treshold_local(int mode, int tolType, double tolValue, const std::vector<V3>& pts, double& depth) {
    double minZ = std::numeric_limits<double>::max();
    double maxZ = std::numeric_limits<double>::lowest();
    for (const auto& p : pts) {
        if (p.z < minZ) minZ = p.z;
        if (p.z > maxZ) maxZ = p.z;
    }
    double depthInterval = maxZ - minZ;
    double lo, hi;
    if (tolType == 1) {  // relative
        double r = depthInterval * tolValue;
        lo = minZ - r;
        hi = maxZ + r;
    } else {  // absolute
        lo = minZ - tolValue;
        hi = maxZ + tolValue;
    }
    if (depth < lo) {
        if (mode == 0) {
            depth = -1;
            return 1;
        }
        depth = lo;
    } else if (depth > hi) {
        if (mode == 0) {
            depth = -1;
            return 2;
        }
        depth = hi;
    }
    return 0;
}

// counter-based RNG shared (by definition, not by code) with the CUDA RANSAC kernel
inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint64_t hash3(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
    return mix64(mix64(mix64(mix64(seed) ^ a) ^ b) ^ c);
}

}  // namespace

// ==============================================================================================
struct orc_estimator {
    orc_params P;
    bool initialized = false, have_cloud = false;
    int W = 0, H = 0;
    double f = 0, cx = 0, cy = 0;
    double R[9], t[3];        // lidar -> camera
    double Ri[9], ti[3];      // camera -> lidar (Affine3d::inverse())
    double Kinv[9];           // CameraPinhole::makeIntrinsics().inverse()
    int64_t n = 0;
    std::vector<double> cam;  // 3 x n col-major (_points_cs_camera)
    std::vector<double> img_vis;     // 2 x nvis (_points_cs_image_visible)
    std::vector<int32_t> pointIndex; // visible -> raw (_pointIndex)
    std::vector<int32_t> map;        // W*H, offset x + y*W, visible index or -1 (_img_points_lidar)

    V3 cam_pt(int raw) const { return {cam[(size_t)raw * 3], cam[(size_t)raw * 3 + 1], cam[(size_t)raw * 3 + 2]}; }

    // A9 CameraPinhole::getViewingRays, MF/include/monolidar_fusion/camera_pinhole.h:52-69
    V3 viewing_ray(double u, double v) const {
        V3 d = {(Kinv[0] * u + Kinv[1] * v) + Kinv[2] * 1.0, (Kinv[3] * u + Kinv[4] * v) + Kinv[5] * 1.0,
                (Kinv[6] * u + Kinv[7] * v) + Kinv[8] * 1.0};
        return normalized(d);
    }

    // A5 NeighborFinderPixel::getNeighbors, MF/src/NeighborFinderPixel.cpp:60-95 (visible indices)
    void neighbors_visible(double u, double v, float scaleW, float scaleH, std::vector<int>& out) const {
        double halfSizeX = static_cast<double>(P.pixelarea_search_witdh) * 0.5 * static_cast<double>(scaleW);
        double halfSizeY = static_cast<double>(P.pixelarea_search_height) * 0.5 * static_cast<double>(scaleH);
        double leftEdgeX = std::max(u - halfSizeX, 0.);
        double rightEdgeX = std::min(u + halfSizeX, static_cast<double>(W - 1));
        double topEdgeY = std::max(v - halfSizeY, 0.);
        double bottomEdgeY = std::min(v + halfSizeY, static_cast<double>(H - 1));
        for (int i = static_cast<int>(topEdgeY); i <= static_cast<int>(bottomEdgeY); i++)
            for (int j = static_cast<int>(leftEdgeX); j <= static_cast<int>(rightEdgeX); j++) {
                // the reference indexes an Eigen::MatrixXi without bounds checks; i, j are inside
                // [0,H-1] x [0,W-1] for finite features, NaN features make the casts UB upstream.
                if (i < 0 || j < 0 || i >= H || j >= W) continue;
                int idx = map[(size_t)j + (size_t)i * (size_t)W];
                if (idx != -1) out.push_back(idx);
            }
    }

    // DepthEstimator::CalculateNeighbors, MF/src/DepthEstimator.cpp:636-684
    bool calculate_neighbors(double u, double v, float sW, float sH, std::vector<int>& idxCut, std::vector<V3>& nb) const {
        neighbors_visible(u, v, sW, sH, idxCut);
        for (int index : idxCut) nb.push_back(cam_pt(pointIndex[(size_t)index]));  // NeighborFinderBase.cpp:15-27
        if (nb.size() < (unsigned)P.radiusSearch_count_min) return false;            // :680
        return true;
    }

    // DepthEstimator::CalculateDepthSegmentation, MF/src/DepthEstimator.cpp:726-780
    bool depth_segmentation(const std::vector<V3>& nb, std::vector<V3>& seg) const {
        seg.clear();
        if (P.do_use_histogram_segmentation) {
            std::vector<double> d(nb.size());
            for (size_t i = 0; i < nb.size(); i++) d[i] = std::min(nb[i].z, 999.);
            std::vector<int> keep;
            double lo, hi;
            if (!histogram_filter(d, P.histogram_segmentation_bin_witdh, P.histogram_segmentation_min_pointcount, keep, lo, hi))
                return false;
            for (int i : keep) seg.push_back(nb[(size_t)i]);
        } else {
            seg = nb;
        }
        return true;
    }

    // A13 Mono_LidarPipeline::PCA, MF/src/PCA.cpp:11-62 (0 plane, 12 point, 13 line, 14 cubic)  UNPINNED
    int pca(const std::vector<V3>& pts, V3& normal, V3& anchor) const {
        size_t n_ = pts.size();
        V3 mean = {0, 0, 0};
        for (const auto& p : pts) mean = mean + p;
        mean = mean / (double)n_;
        double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (const auto& p : pts) {
            V3 d = p - mean;
            c[0] += d.x * d.x; c[1] += d.x * d.y; c[2] += d.x * d.z;
            c[4] += d.y * d.y; c[5] += d.y * d.z; c[8] += d.z * d.z;
        }
        c[3] = c[1]; c[6] = c[2]; c[7] = c[5];
        double w[3], v[9];
        eig3_sym(c, w, v);
        double ev1 = w[0], ev2 = w[1], ev3 = w[2];
        float planarity = (float)((ev2 - ev1) / ev3);
        float linearity = (float)((ev3 - ev2) / ev3);
        V3 e0 = {v[0], v[3], v[6]};
        normal = e0 / norm(e0);
        anchor = mean;
        if (planarity < P.pca_treshold_2_1_rel_min) return PcaIsCubic;
        if (linearity > P.pca_treshold_3_2_rel_max) return PcaIsLine;
        if (ev3 < P.pca_treshold_3_abs_min) return PcaIsPoint;
        return 0;
    }

    // A12 DepthEstimator::CalculateDepthSegmented, MF/src/DepthEstimator.cpp:903-1037
    std::pair<int, double> depth_segmented(double u, double v, const std::vector<V3>& seg, bool checkPlanar) const {
        V3 c1{}, c2{}, c3{};
        if (!P.do_use_PCA && P.do_use_triangle_size_maximation) {
            int i, j, k;
            if (!max_spanning_triangle(seg, 0.0, i, j, k)) return {TriangleNotPlanarInsufficientPoints, -1};
            c1 = seg[(size_t)i]; c2 = seg[(size_t)j]; c3 = seg[(size_t)k];
        } else {
            if (seg.size() < 3) return {HistogramNoLocalMax, -1};
            c1 = seg[0]; c2 = seg[1]; c3 = seg[2];
        }
        if (!P.do_use_PCA && checkPlanar)
            if (!check_planar(c1, c2, c3, P.triangleplanar_crossnorm_treshold)) return {TriangleNotPlanar, -1};

        V3 support = {0, 0, 0};
        V3 dir = viewing_ray(u, v);
        if (dir.z < 0) dir = dir * -1.0;
        double depth;
        double ortho = P.viewray_plane_orthoganality_treshold;  // >0 selects the OrthogonalTreshold module (:77-81)
        if (P.do_use_PCA) {
            V3 normal, anchor;
            int r = pca(seg, normal, anchor);
            if (r != 0) return {r, -1};
            Hyperplane pl = {normal, -dot(normal, anchor)};  // Hyperplane(normal, point)
            if (!line_plane(pl, support, dir, ortho, depth)) return {PlaneViewrayNotOrthogonal, -1};
        } else {
            Hyperplane pl = plane_through(c1, c2, c3);
            if (!line_plane(pl, support, dir, ortho, depth)) return {PlaneViewrayNotOrthogonal, -1};
        }
        if (P.treshold_depth_enabled) {
            int r = treshold_global(P.treshold_depth_mode, (double)P.treshold_depth_min, (double)P.treshold_depth_max, depth);
            if (r == 1) return {TresholdDepthGlobalSmallerMin, -1};
            if (r == 2) return {TresholdDepthGlobalGreaterMax, -1};
        }
        if (P.treshold_depth_local_enabled) {
            int r = treshold_local(P.treshold_depth_local_mode, P.treshold_depth_local_valuetype, P.treshold_depth_local_value, seg, depth);
            if (r == 1) return {TresholdDepthLocalSmallerMin, -1};
            if (r == 2) return {TresholdDepthLocalGreaterMax, -1};
        }
        if (depth < 0 && P.do_use_cut_behind_camera) return {CornerBehindCamera, -1};
        return {Success, depth};
    }

    // R2 DepthEstimator::CalculateDepthSegmentationPlane, MF/src/DepthEstimator.cpp:782-900.
    // pcl::pointToPlaneDistance(PointXYZ, Vector4f) evaluates a*x+b*y+c*z+d in FLOAT, left to right.
    bool segmentation_plane(const std::vector<V3>& nb, const std::vector<int>& idxCut, const orc_plane& plane,
                            const std::vector<uint8_t>& inlier_mask, std::vector<V3>& seg) const {
        seg.clear();
        double treshold = P.ransac_plane_point_distance_treshold;
        for (size_t i = 0; i < nb.size(); i++) {
            int raw = pointIndex[(size_t)idxCut[i]];
            const V3& p = nb[i];
            double lx = ((Ri[0] * p.x + Ri[1] * p.y) + Ri[2] * p.z) + ti[0];
            double ly = ((Ri[3] * p.x + Ri[4] * p.y) + Ri[5] * p.z) + ti[1];
            double lz = ((Ri[6] * p.x + Ri[7] * p.y) + Ri[8] * p.z) + ti[2];
            float fx = (float)lx, fy = (float)ly, fz = (float)lz;
            float s = ((plane.coeffs[0] * fx + plane.coeffs[1] * fy) + plane.coeffs[2] * fz) + plane.coeffs[3];
            double distance = std::fabs((double)s);
            if (distance > treshold) return false;
            if (raw >= 0 && (size_t)raw < inlier_mask.size() && inlier_mask[(size_t)raw]) seg.push_back(p);
        }
        if (seg.size() < 3) return false;
        return true;
    }

    // R3 RoadDepthEstimatorMEstimator::CalculateDepth (MF/src/RoadDepthEstimatorMEstimator.cpp:28-74) with
    //    PlaneEstimationMEstimator::EstimatePlane (MF/src/PlaneEstimationMEstimator.cpp:18-55); weighted==false
    //    gives the unweighted total-least-squares plane used for R4 (the reference's Ceres variant is UB).
    std::pair<int, double> road_mestimator(double u, double v, const std::vector<V3>& pts, const Hyperplane& prior, bool weighted) const {
        size_t k = pts.size();
        V3 center = {0, 0, 0};
        std::vector<double> w(k);
        double wsum = 0;
        for (size_t i = 0; i < k; i++) {
            w[i] = weighted ? 1 / std::fabs(dot(prior.n, pts[i]) + prior.off) : 1.0;
            center = center + pts[i] * w[i];
            wsum += w[i];
        }
        center = center / wsum;
        std::vector<V3> cols(k);
        for (size_t i = 0; i < k; i++) cols[i] = (pts[i] - center) * std::sqrt(w[i]);
        V3 nrm = normalized(smallest_left_singular_vector(cols));
        Hyperplane pl = {nrm, -dot(nrm, center)};
        return road_finish(u, v, pl, pts);
    }

    // shared tail of the road estimators: ray with swapped arguments (RoadDepthEstimatorMEstimator.cpp:52-53),
    // LinePlaneIntersectionNormal (no orthogonality gate), thresholds, SuccessRoad.
    std::pair<int, double> road_finish(double u, double v, const Hyperplane& pl, const std::vector<V3>& pts) const {
        V3 support = {0, 0, 0};
        V3 dir = viewing_ray(u, v);
        if (dir.z < 0) dir = dir * -1.0;
        double depth;
        line_plane(pl, dir, support, 0.0, depth);
        if (P.treshold_depth_enabled) {
            int r = treshold_global(P.treshold_depth_mode, (double)P.treshold_depth_min, (double)P.treshold_depth_max, depth);
            if (r == 1) return {TresholdDepthGlobalSmallerMin, -1};
            if (r == 2) return {TresholdDepthGlobalGreaterMax, -1};
        }
        if (P.treshold_depth_local_enabled) {
            int r = treshold_local(P.treshold_depth_local_mode, P.treshold_depth_local_valuetype, P.treshold_depth_local_value, pts, depth);
            if (r == 1) return {TresholdDepthLocalSmallerMin, -1};
            if (r == 2) return {TresholdDepthLocalGreaterMax, -1};
        }
        return {SuccessRoad, depth};
    }

    // R5 RoadDepthEstimatorMaxSpanningTriangle::CalculateDepth, MF/src/RoadDepthEstimatorMaxSpanningTriangle.cpp:24-75
    //    + LinePlaneIntersectionCeckXZTreshold::Check, MF/src/LinePlaneIntersectionCeckXZTreshold.cpp:15-45
    std::pair<int, double> road_triangle(double u, double v, const std::vector<V3>& pts) const {
        int i, j, k;
        if (!max_spanning_triangle(pts, 0.0, i, j, k)) return {RadiusSearchInsufficientPoints, -1};
        double minX = std::numeric_limits<double>::max(), maxX = std::numeric_limits<double>::lowest();
        double minZ = minX, maxZ = maxX;
        for (const auto& p : pts) {
            if (p.x < minX) minX = p.x;
            if (p.x > maxX) maxX = p.x;
            if (p.z < minZ) minZ = p.z;
            if (p.z > maxZ) maxZ = p.z;
        }
        double relation = (maxZ - minZ) / (maxX - minX);
        if (!(relation >= P.plane_estimator_z_x_min_relation)) return {InsufficientRoadPoints, -1};
        Hyperplane pl = plane_through(pts[(size_t)i], pts[(size_t)j], pts[(size_t)k]);
        return road_finish(u, v, pl, pts);
    }

    // per-feature driver, MF/src/DepthEstimator.cpp:491-600
    std::pair<int, double> feature_depth(double u, double v, const orc_plane* plane, const std::vector<uint8_t>& inlier_mask,
                                         const Hyperplane& prior) const {
        std::vector<int> idxCut;
        std::vector<V3> nb;
        std::pair<int, double> result = {Unspecified, -1};
        if (!calculate_neighbors(u, v, 1.0f, 1.0f, idxCut, nb)) return {RadiusSearchInsufficientPoints, -1};
        std::vector<V3> seg;
        if (!depth_segmentation(nb, seg)) result = {HistogramNoLocalMax, -1};
        bool checkPlanar = P.do_check_triangleplanar_condition != 0;
        if (result.first != HistogramNoLocalMax) {
            result = depth_segmented(u, v, seg, checkPlanar);
            if (result.first == Success) return result;
        }
        int resultOld = result.first;
        if (plane != nullptr && P.do_use_ransac_plane) {  // ransacPlane != nullptr && _roadDepthEstimator != NULL
            idxCut.clear();
            nb.clear();
            if (!calculate_neighbors(u, v, 2.0f, 1.5f, idxCut, nb)) return {RadiusSearchInsufficientPoints, -1};
            if (!segmentation_plane(nb, idxCut, *plane, inlier_mask, seg)) return {resultOld, -1};
            if (P.plane_estimator_use_triangle_maximation)
                result = road_triangle(u, v, seg);
            else if (P.plane_estimator_use_leastsquares)
                result = road_mestimator(u, v, seg, prior, false);
            else
                result = road_mestimator(u, v, seg, prior, true);
        }
        return result;
    }
};

extern "C" {

void orc_default_params(orc_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->neighbor_search_mode = 0;
    p->pixelarea_search_witdh = 12;
    p->pixelarea_search_height = 15;
    p->radiusSearch_count_min = 3;
    p->do_use_histogram_segmentation = 1;
    p->histogram_segmentation_bin_witdh = 0.5;
    p->histogram_segmentation_min_pointcount = 3;
    p->do_use_depth_segmentation = 0;
    p->treshold_depth_enabled = 1;
    p->treshold_depth_mode = 0;
    p->treshold_depth_max = 100;
    p->treshold_depth_min = 0;
    p->treshold_depth_local_enabled = 1;
    p->treshold_depth_local_mode = 0;
    p->treshold_depth_local_valuetype = 1;
    p->treshold_depth_local_value = 0.5;
    p->do_use_PCA = 0;
    p->pca_debug = 0;
    p->pca_treshold_3_abs_min = 0.005;
    p->pca_treshold_3_2_rel_max = 15;
    p->pca_treshold_2_1_rel_min = 0.5;
    p->do_use_ransac_plane = 1;
    p->ransac_plane_distance_treshold = 0.2;
    p->ransac_plane_min_z = -10000;
    p->ransac_plane_max_z = 10000;
    p->ransac_plane_max_iterations = 10000;
    p->ransac_plane_use_refinement = 1;
    p->ransac_plane_refinement_treshold = 10.2;
    p->ransac_plane_use_camx_treshold = 0;
    p->ransac_plane_treshold_camx = 2.0;
    p->ransac_plane_point_distance_treshold = 0.2;
    p->ransac_plane_probability = 0.999;
    p->plane_estimator_use_triangle_maximation = 0;
    p->plane_estimator_z_x_min_relation = 0;
    p->plane_estimator_use_leastsquares = 0;
    p->plane_estimator_use_mestimator = 1;
    p->do_use_cut_behind_camera = 1;
    p->do_use_triangle_size_maximation = 1;
    p->do_check_triangleplanar_condition = 1;
    p->triangleplanar_crossnorm_treshold = 0.1;
    p->viewray_plane_orthoganality_treshold = 1.0;  // `{01}` is octal 1 (DepthEstimatorParameters.h:155)
    p->set_all_depths_to_zero = 0;
}

void orc_yaml_params(orc_params* p) {
    orc_default_params(p);
    p->pixelarea_search_witdh = 6;
    p->pixelarea_search_height = 9;
    p->radiusSearch_count_min = 1;
    p->histogram_segmentation_bin_witdh = 0.3;
    p->histogram_segmentation_min_pointcount = 3;
    p->do_use_depth_segmentation = 0;  // shipped value 1 makes the reference throw (DepthEstimator.cpp:608)
    p->pca_treshold_2_1_rel_min = 1.5;
    p->ransac_plane_distance_treshold = 0.3;
    p->viewray_plane_orthoganality_treshold = 0.03;
    // keys absent from the yaml read as 0 through cv::FileStorage; ransac_plane_min_z/max_z are such
    // keys, which would make PassThrough keep only z == 0 -- the struct defaults are kept instead.
}

orc_estimator* orc_create(const orc_params* p) {
    auto* e = new orc_estimator();
    e->P = *p;
    return e;
}
void orc_destroy(orc_estimator* e) { delete e; }

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n <= 0) n = omp_get_num_procs();
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_get_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_initialize(orc_estimator* e, int W, int H, double f, double cx, double cy, const double* T) {
    if (e->P.neighbor_search_mode != 0) return -2;  // only the pixel finder exists (DepthEstimator.cpp:47-57)
    e->W = W; e->H = H; e->f = f; e->cx = cx; e->cy = cy;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) e->R[r * 3 + c] = T[r * 4 + c];
        e->t[r] = T[r * 4 + 3];
    }
    // Eigen Transform::inverse(Affine): linear^-1 by cofactors, translation = (-linear^-1) * t
    inverse3(e->R, e->Ri);
    for (int r = 0; r < 3; r++)
        e->ti[r] = ((-e->Ri[r * 3 + 0]) * e->t[0] + (-e->Ri[r * 3 + 1]) * e->t[1]) + (-e->Ri[r * 3 + 2]) * e->t[2];
    double K[9] = {f, 0, cx, 0, f, cy, 0, 0, 1};  // camera_pinhole.h:100-106
    inverse3(K, e->Kinv);
    e->map.assign((size_t)W * (size_t)H, -1);
    e->initialized = true;
    return 0;
}

// A2/A3/A4: Transform_Cloud_LidarToCamera (DepthEstimator.cpp:156-217), CameraPinhole::getImagePoints
// (camera_pinhole.h:85-97), NeighborFinderPixel::InitializeLidarProjection (NeighborFinderPixel.cpp:29-58).
// Serial like the reference. A4/A5 PINNED by the window-extent property; projection numerics UNPINNED.
int orc_set_cloud(orc_estimator* e, const float* pts, int64_t n, int stride_floats) {
    if (!e->initialized) return -1;  // throw "call of 'setInputCloud' without 'initialize'"
    const int W = e->W, H = e->H;
    const double f = e->f, cx = e->cx, cy = e->cy;
    e->n = n;
    e->cam.resize((size_t)n * 3);
    std::vector<double> img((size_t)n * 2);
    std::vector<uint8_t> inRange((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        double x = (double)pts[i * stride_floats + 0], y = (double)pts[i * stride_floats + 1], z = (double)pts[i * stride_floats + 2];
        double X = ((e->R[0] * x + e->R[1] * y) + e->R[2] * z) + e->t[0];
        double Y = ((e->R[3] * x + e->R[4] * y) + e->R[5] * z) + e->t[1];
        double Z = ((e->R[6] * x + e->R[7] * y) + e->R[8] * z) + e->t[2];
        e->cam[(size_t)i * 3 + 0] = X; e->cam[(size_t)i * 3 + 1] = Y; e->cam[(size_t)i * 3 + 2] = Z;
        double q0 = (f * X + 0.0 * Y) + cx * Z;
        double q1 = (0.0 * X + f * Y) + cy * Z;
        double q2 = (0.0 * X + 0.0 * Y) + 1.0 * Z;
        double u = q0 / q2, v = q1 / q2;  // colwise().hnormalized()
        img[(size_t)i * 2] = u; img[(size_t)i * 2 + 1] = v;
        inRange[(size_t)i] = (u >= 0.) && (u <= static_cast<double>(W)) && (v >= 0.) && (v <= static_cast<double>(H));
    }
    e->img_vis.clear();
    e->pointIndex.clear();
    for (int64_t i = 0; i < n; i++) {
        if (inRange[(size_t)i]) {
            double u = img[(size_t)i * 2], v = img[(size_t)i * 2 + 1];
            if ((u > 0) && (u < W) && (v > 0) && (v < H)) {
                e->img_vis.push_back(u);
                e->img_vis.push_back(v);
                e->pointIndex.push_back((int32_t)i);
            }
        }
    }
    std::fill(e->map.begin(), e->map.end(), -1);
    int pointCount = (int)e->pointIndex.size();
    for (int i = 0; i < pointCount; i++) {
        int x_img = (int)e->img_vis[(size_t)i * 2];
        int y_img = (int)e->img_vis[(size_t)i * 2 + 1];
        int indexRaw = e->pointIndex[(size_t)i];
        double zc = e->cam[(size_t)indexRaw * 3 + 2];
        int32_t& cell = e->map[(size_t)x_img + (size_t)y_img * (size_t)W];
        if ((cell == -1) && (zc > 0)) cell = i;
    }
    e->have_cloud = true;
    return 0;
}

int orc_calculate_depth(orc_estimator* e, const double* uv, int F, double* depth, int32_t* status, const orc_plane* plane) {
    if (!e->have_cloud) return -1;  // throw "call of 'CalculateDepth' without 'SetInputCloud'"
    if (e->P.set_all_depths_to_zero) {  // DepthEstimator.cpp:448-453
        for (int i = 0; i < F; i++) {
            status[i] = 1;
            depth[i] = -1;
        }
        return 0;
    }
    if (e->P.do_use_depth_segmentation) return -3;  // region growing throws (DepthEstimator.cpp:608)
    std::vector<uint8_t> mask;
    Hyperplane prior = {{0, 0, 0}, 0};
    if (plane != nullptr) {
        mask.assign((size_t)e->n, 0);
        for (int64_t i = 0; i < plane->n_inliers; i++) {
            int32_t r = plane->inlier_idx[i];
            if (r >= 0 && r < e->n) mask[(size_t)r] = 1;
        }
        // prior for the M-estimator, DepthEstimator.cpp:286-292 (lidar-frame coefficients, as is)
        V3 pn = {(double)plane->coeffs[0], (double)plane->coeffs[1], (double)plane->coeffs[2]};
        prior = {normalized(pn), (double)plane->coeffs[3]};
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < F; i++) {
        auto r = e->feature_depth(uv[(size_t)i * 2], uv[(size_t)i * 2 + 1], plane, mask, prior);
        depth[i] = r.second;
        status[i] = r.first;
    }
    return 0;
}

int64_t orc_visible_count(const orc_estimator* e) { return (int64_t)e->pointIndex.size(); }
void orc_get_point_index(const orc_estimator* e, int32_t* out) { std::memcpy(out, e->pointIndex.data(), e->pointIndex.size() * sizeof(int32_t)); }
void orc_get_image_points_visible(const orc_estimator* e, double* out) { std::memcpy(out, e->img_vis.data(), e->img_vis.size() * sizeof(double)); }
void orc_get_points_camera(const orc_estimator* e, double* out) { std::memcpy(out, e->cam.data(), e->cam.size() * sizeof(double)); }
void orc_get_pixel_map_visible(const orc_estimator* e, int32_t* out) { std::memcpy(out, e->map.data(), e->map.size() * sizeof(int32_t)); }
void orc_get_pixel_map_raw(const orc_estimator* e, int32_t* out) {
    for (size_t i = 0; i < e->map.size(); i++) out[i] = e->map[i] < 0 ? -1 : e->pointIndex[(size_t)e->map[i]];
}
int orc_get_neighbors(const orc_estimator* e, double u, double v, double scale_w, double scale_h, int32_t* out_raw, int cap) {
    std::vector<int> idx;
    e->neighbors_visible(u, v, (float)scale_w, (float)scale_h, idx);
    int k = (int)idx.size();
    for (int i = 0; i < k && i < cap; i++) out_raw[i] = e->pointIndex[(size_t)idx[(size_t)i]];
    return k;
}

int orc_histogram_filter(const double* depths, int n, double bin_width, int min_count, int32_t* out_pos, int* n_out,
                         double* lower, double* higher) {
    std::vector<double> d(depths, depths + n);
    std::vector<int> keep;
    bool ok = histogram_filter(d, bin_width, min_count, keep, *lower, *higher);
    *n_out = (int)keep.size();
    for (size_t i = 0; i < keep.size(); i++) out_pos[i] = keep[i];
    return ok ? 1 : 0;
}

int orc_neighbor_finder(int W, int H, int search_w, int search_h, const double* img, const double* cam, int n, double u,
                        double v, int32_t* out_idx, int cap) {
    // NeighborFinderPixel used standalone with pointIndex = identity, as in the reference's test (:139-145)
    orc_estimator e;
    orc_default_params(&e.P);
    e.P.pixelarea_search_witdh = search_w;
    e.P.pixelarea_search_height = search_h;
    e.W = W; e.H = H;
    e.map.assign((size_t)W * (size_t)H, -1);
    for (int i = 0; i < n; i++) {
        int x_img = (int)img[(size_t)i * 2], y_img = (int)img[(size_t)i * 2 + 1];
        if (x_img < 0 || y_img < 0 || x_img >= W || y_img >= H) continue;
        int32_t& cell = e.map[(size_t)x_img + (size_t)y_img * (size_t)W];
        if (cell == -1 && cam[(size_t)i * 3 + 2] > 0) cell = i;
    }
    std::vector<int> idx;
    e.neighbors_visible(u, v, 1.0f, 1.0f, idx);
    int k = (int)idx.size();
    for (int i = 0; i < k && i < cap; i++) out_idx[i] = idx[(size_t)i];
    return k;
}

void orc_viewing_ray(int W, int H, double f, double cx, double cy, double u, double v, double* dir3) {
    orc_estimator e;
    (void)W; (void)H;
    double K[9] = {f, 0, cx, 0, f, cy, 0, 0, 1};
    inverse3(K, e.Kinv);
    V3 d = e.viewing_ray(u, v);
    dir3[0] = d.x; dir3[1] = d.y; dir3[2] = d.z;
}
int orc_image_point(int W, int H, double f, double cx, double cy, const double* p, double* uv2) {
    double q0 = (f * p[0] + 0.0 * p[1]) + cx * p[2];
    double q1 = (0.0 * p[0] + f * p[1]) + cy * p[2];
    double q2 = (0.0 * p[0] + 0.0 * p[1]) + 1.0 * p[2];
    uv2[0] = q0 / q2;
    uv2[1] = q1 / q2;
    return (uv2[0] >= 0.) && (uv2[0] <= (double)W) && (uv2[1] >= 0.) && (uv2[1] <= (double)H);
}

// ----------------------------------------------------------------------------------------------
// R1  RansacPlane::CalculateInliersPlane, MF/src/RansacPlane.cpp:41-140.  PARITY UNPINNED except for the
// reference's +-0.2 coefficient test. PCL (un-vendored, unpinned; 1.8 semantics assumed) restated:
//   PassThrough("z", min_z, max_z)            -> keep finite points with min_z <= z <= max_z      (:58-64)
//   RandomSample(6000)                        -> order-preserving subsample; PCL's is time-seeded
//                                                selection sampling, restated as stratified sampling
//                                                driven by a counter-based hash                    (:66-74)
//   SampleConsensusModelPerpendicularPlane    -> axis (0,0,1), eps 10 deg                          (:94-100)
//   RandomSampleConsensus::computeModel       -> adaptive loop k = log(1-p)/log(1-w^3)             (:102-108)
//   optimizeModelCoefficients                 -> centroid + smallest eigenvector of the covariance
//                                                (double accumulation here; PCL uses float)        (:117-126)
//   selectWithinDistance(UNREFINED coeffs, refinement threshold)                                   (:121)
// ----------------------------------------------------------------------------------------------
namespace {
struct F3 {
    float x, y, z;
};
inline bool sample_good(const F3& p0, const F3& p1, const F3& p2) {
    // SampleConsensusModelPlane::isSampleGood: dy1dy2 = (p1-p0)/(p2-p0) componentwise
    float r0 = (p1.x - p0.x) / (p2.x - p0.x), r1 = (p1.y - p0.y) / (p2.y - p0.y), r2 = (p1.z - p0.z) / (p2.z - p0.z);
    return (r0 != r1) || (r2 != r1);
}
inline void plane_from_sample(const F3& p0, const F3& p1, const F3& p2, float* c) {
    float ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
    float bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
    float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    float z = (nx * nx + ny * ny) + nz * nz;
    if (z > 0) {
        float s = std::sqrt(z);
        nx /= s; ny /= s; nz /= s;
    }
    c[0] = nx; c[1] = ny; c[2] = nz;
    c[3] = -1 * ((nx * p0.x + ny * p0.y) + nz * p0.z);
}
inline bool model_valid(const float* c, double cos_eps) {
    // isModelValid: angle(axis z, normal) folded to [0,pi/2] <= eps  <=>  |n_z| / |n| >= cos(eps)
    float z = (c[0] * c[0] + c[1] * c[1]) + c[2] * c[2];
    float nz = c[2];
    if (z > 0) nz = c[2] / std::sqrt(z);
    return std::fabs((double)nz) >= cos_eps;
}
inline double plane_dist(const float* c, const F3& p) {
    float s = ((c[0] * p.x + c[1] * p.y) + c[2] * p.z) + c[3];
    return std::fabs((double)s);
}
}  // namespace

int orc_ransac_plane(const orc_params* P, const float* pts, int64_t n, int stride_floats, uint64_t seed, float* coeffs4,
                     int32_t* inlier_idx, int64_t* n_inliers, int32_t* iterations_out) {
    *n_inliers = 0;
    if (iterations_out) *iterations_out = 0;
    if (n < 3) return -1;  // ExceptionPclInvalid
    auto P3 = [&](int64_t i) { return F3{pts[i * stride_floats], pts[i * stride_floats + 1], pts[i * stride_floats + 2]}; };
    std::vector<int32_t> cand;
    cand.reserve((size_t)n);
    if (P->ransac_plane_min_z > -1001.) {
        for (int64_t i = 0; i < n; i++) {
            float z = pts[i * stride_floats + 2];
            // pcl::PassThrough::applyFilterIndices first drops every point with a non-finite x, y or z, then tests the field
            const float x = pts[i * stride_floats], y = pts[i * stride_floats + 1];
            if (std::isfinite(x) && std::isfinite(y) && std::isfinite(z) && !((double)z < P->ransac_plane_min_z) && !((double)z > P->ransac_plane_max_z))
                cand.push_back((int32_t)i);
        }
    } else {
        for (int64_t i = 0; i < n; i++) cand.push_back((int32_t)i);
    }
    const int64_t S = 6000;  // _numberRandomSamplePoints, RansacPlane.cpp:32
    std::vector<int32_t> sub;
    int64_t L = (int64_t)cand.size();
    if (L <= S) {
        sub = cand;
    } else {
        sub.resize((size_t)S);
        for (int64_t j = 0; j < S; j++) {
            int64_t lo = (j * L) / S, hi = ((j + 1) * L) / S;
            uint64_t len = (uint64_t)(hi - lo);
            sub[(size_t)j] = cand[(size_t)(lo + (int64_t)(hash3(seed, 0x5A17, (uint64_t)j, 0) % len))];
        }
    }
    const int64_t M = (int64_t)sub.size();
    const double cos_eps = std::cos(M_PI / 18.);
    const double threshold = P->ransac_plane_distance_treshold;
    const int max_iterations = P->ransac_plane_max_iterations;
    float best[4] = {0, 0, 0, 0};
    bool have_model = false;
    if (M >= 3) {
        int iterations = 0;
        int n_best = -INT_MAX;
        double k = 1.0;
        double log_probability = std::log(1.0 - P->ransac_plane_probability);
        double one_over_indices = 1.0 / (double)M;
        unsigned skipped = 0;
        const unsigned max_skip = (unsigned)max_iterations * 10u;
        while (iterations < k && skipped < max_skip) {
            uint64_t draw = (uint64_t)iterations + (uint64_t)skipped;
            bool got = false;
            F3 p0{}, p1{}, p2{};
            for (int a = 0; a < 1000 && !got; a++) {  // max_sample_checks_
                uint64_t h0 = hash3(seed, draw, (uint64_t)a, 0), h1 = hash3(seed, draw, (uint64_t)a, 1), h2 = hash3(seed, draw, (uint64_t)a, 2);
                int64_t i0 = (int64_t)(h0 % (uint64_t)M);
                int64_t i1 = (int64_t)(h1 % (uint64_t)(M - 1));
                if (i1 >= i0) i1++;
                int64_t i2 = (int64_t)(h2 % (uint64_t)(M - 2));
                int64_t lo = std::min(i0, i1), hi = std::max(i0, i1);
                if (i2 >= lo) i2++;
                if (i2 >= hi) i2++;
                p0 = P3(sub[(size_t)i0]); p1 = P3(sub[(size_t)i1]); p2 = P3(sub[(size_t)i2]);
                got = sample_good(p0, p1, p2);
            }
            if (!got) break;  // "No samples could be selected!"
            float c[4];
            plane_from_sample(p0, p1, p2, c);
            int count = 0;
            if (model_valid(c, cos_eps))
                for (int64_t j = 0; j < M; j++)
                    if (plane_dist(c, P3(sub[(size_t)j])) < threshold) count++;
            if (count > n_best) {
                n_best = count;
                std::memcpy(best, c, sizeof(best));
                have_model = true;
                double w = (double)n_best * one_over_indices;
                double p_no_outliers = 1.0 - std::pow(w, 3.0);
                p_no_outliers = std::max(std::numeric_limits<double>::epsilon(), p_no_outliers);
                p_no_outliers = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no_outliers);
                k = log_probability / std::log(p_no_outliers);
            }
            ++iterations;
            if (iterations > max_iterations) break;
        }
        if (iterations_out) *iterations_out = iterations;
    }
    if (!have_model) return -2;
    std::vector<int32_t> inl;
    if (model_valid(best, cos_eps))
        for (int64_t j = 0; j < M; j++)
            if (plane_dist(best, P3(sub[(size_t)j])) < threshold) inl.push_back(sub[(size_t)j]);
    float out[4] = {best[0], best[1], best[2], best[3]};
    if (P->ransac_plane_use_refinement) {
        float refined[4] = {best[0], best[1], best[2], best[3]};
        if (inl.size() >= 4) {
            double sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
            for (int32_t r : inl) {
                F3 p = P3(r);
                double x = p.x, y = p.y, z = p.z;
                sx += x; sy += y; sz += z;
                sxx += x * x; sxy += x * y; sxz += x * z; syy += y * y; syz += y * z; szz += z * z;
            }
            double m = (double)inl.size();
            double mx = sx / m, my = sy / m, mz = sz / m;
            double c[9] = {sxx / m - mx * mx, sxy / m - mx * my, sxz / m - mx * mz,
                           sxy / m - mx * my, syy / m - my * my, syz / m - my * mz,
                           sxz / m - mx * mz, syz / m - my * mz, szz / m - mz * mz};
            double w[3], v[9];
            eig3_sym(c, w, v);
            float ex = (float)v[0], ey = (float)v[3], ez = (float)v[6];
            float cand4[4] = {ex, ey, ez, 0};
            cand4[3] = -1 * ((ex * (float)mx + ey * (float)my) + ez * (float)mz);
            if (model_valid(cand4, cos_eps)) std::memcpy(refined, cand4, sizeof(refined));
        }
        // selectWithinDistance(modelCoeffs /*un-refined*/, _planeRefinementDistance, _inliersIndex)
        inl.clear();
        if (model_valid(best, cos_eps))
            for (int64_t j = 0; j < M; j++)
                if (plane_dist(best, P3(sub[(size_t)j])) < P->ransac_plane_refinement_treshold) inl.push_back(sub[(size_t)j]);
        std::memcpy(out, refined, sizeof(out));
    }
    std::memcpy(coeffs4, out, sizeof(out));
    *n_inliers = (int64_t)inl.size();
    for (size_t i = 0; i < inl.size(); i++) inlier_idx[i] = inl[i];
    return 0;
}

}  // extern "C"
