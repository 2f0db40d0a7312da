// mld_feature_warp.cuh -- the per-feature driver with one WARP per feature (device functions) of the warp-per-feature kernel
// (mld_feature.cu: the general path for dense clouds / large windows and the overflow pass of the chunked pipelines). See
// mld_feature.cu for the reference routines restated and the mapping.
#pragma once
#include "mld_common.cuh"
#include "mld_geometry.cuh"

namespace {

struct WarpSlab {
    double* x;
    double* y;
    double* z;
    int* raw;  // raw point index in the normal path; reused for bin ids by the histogram
};

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---- A5: window scan + gather --------------------------------------------------------------
// Returns the neighbour count k; slab[0..k) holds the camera-frame points in scan order.
__device__ __forceinline__ int gather_window(const DevParams& P, const MapCode& mc, const unsigned int* __restrict__ map,
                                             const float* __restrict__ pts, int stride_f, double u, double v, double hx, double hy, int lane,
                                             const WarpSlab& s, const int KCAP) {
    // NaN / out-of-int-range features are undefined behaviour in the reference ((int) casts of the
    // window edges); they are defined here as "empty window".
    if (!(fabs(u) < 1e9) || !(fabs(v) < 1e9)) return 0;
    double leftEdgeX = fmax(u - hx, 0.);
    double rightEdgeX = fmin(u + hx, (double)(P.W - 1));
    double topEdgeY = fmax(v - hy, 0.);
    double bottomEdgeY = fmin(v + hy, (double)(P.H - 1));
    int x0 = (int)leftEdgeX, x1 = (int)rightEdgeX, y0 = (int)topEdgeY, y1 = (int)bottomEdgeY;
    int wc = x1 - x0 + 1, wr = y1 - y0 + 1;
    if (wc <= 0 || wr <= 0) return 0;
    int area = wc * wr;
    int k = 0;
    const unsigned lt = lanemask_lt();
    for (int base = 0; base < area; base += 32) {
        int idx = base + lane;
        unsigned int cell = MLD_EMPTY;
        if (idx < area) {
            int ry = idx / wc;
            int rx = idx - ry * wc;
            cell = __ldg(&map[(long long)(y0 + ry) * P.W + (x0 + rx)]);
        }
        const bool hit = (idx < area) && map_cell_valid(mc, cell);
        unsigned m = __ballot_sync(MLD_FULL_MASK, hit);
        if (hit) {
            int pos = k + __popc(m & lt);
            if (pos < KCAP) s.raw[pos] = (int)map_cell_index(mc, cell);
        }
        k += __popc(m);
    }
    if (k > KCAP) k = KCAP;  // cannot happen: mld_create rejects windows with area > KCAP
    __syncwarp();
    for (int i = lane; i < k; i += 32) {
        const float* p = pts + (long long)s.raw[i] * stride_f;
        float4 q = __ldg(reinterpret_cast<const float4*>(p));
        D3 c = lidar_to_cam(P, q.x, q.y, q.z);
        s.x[i] = c.x;
        s.y[i] = c.y;
        s.z[i] = c.z;
    }
    __syncwarp();
    return k;
}

// in-place, order-preserving compaction of slab entries with keep flag; flags are evaluated by
// `pred(i)` for i in [0,n). Chunks of 32 are read before they are overwritten.
template <typename Pred>
__device__ int compact_slab(int n, int lane, const WarpSlab& s, bool with_raw, Pred pred) {
    int out = 0;
    const unsigned lt = lanemask_lt();
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        bool keep = false;
        double x = 0, y = 0, z = 0;
        int r = 0;
        if (i < n) {
            x = s.x[i]; y = s.y[i]; z = s.z[i];
            if (with_raw) r = s.raw[i];
            keep = pred(i, x, y, z, r);
        }
        unsigned m = __ballot_sync(MLD_FULL_MASK, keep);
        __syncwarp();
        if (keep) {
            int pos = out + __popc(m & lt);
            s.x[pos] = x; s.y[pos] = y; s.z[pos] = z;
            if (with_raw) s.raw[pos] = r;
        }
        out += __popc(m);
        __syncwarp();
    }
    return out;
}

// ---- A6: histogram foreground segmentation ---------------------------------------------------
// Returns the segmented count (slab compacted in place) or -1 for "no local maximum".
__device__ int histogram_segment(const DevParams& P, int k, int lane, const WarpSlab& s) {
    // depth = min(z, 999.) (DepthEstimator.cpp:741-744); maxDist = running (int)ceil(depth) maximum
    // (HistogramPointDepth.cpp:36-41) == (int)ceil(max depth) for positive depths, else 0.
    double dmax = -1.0;
    for (int i = lane; i < k; i += 32) {
        double d = fmin(s.z[i], 999.);
        dmax = (d > dmax) ? d : dmax;
    }
    dmax = warp_max_d(dmax);
    int maxDist = 0;
    if (dmax > 0.0) maxDist = (int)ceil(dmax);
    int binCount = (int)((maxDist) / P.bin_w + 1);  // :43
    if (binCount <= 1) return -1;                  // :53
    // bin ids (Histogram.cpp:29-30) parked in s.raw (the normal path does not need raw ids any more)
    int bmin = 0x7fffffff;
    for (int i = lane; i < k; i += 32) {
        double value = fmin(fmin(s.z[i], 999.), 1e10);
        int b = (int)fmin(fabs(value / P.bin_w), (double)binCount - 1.);
        s.raw[i] = b;
        bmin = min(bmin, b);
    }
    bmin = warp_min_i(bmin);
    __syncwarp();
    if (k == 0) bmin = binCount;  // no occupied bin at all
    // sequential first-local-maximum scan (:66-85). Bins before the first occupied one hold 0
    // elements and can only register a "maximum" when the minimum count is <= 0.
    int binMaxId = -1, binMaxVal = -1, binValue = 0;
    if (bmin > 0 && 0 >= P.hist_min) {
        binMaxVal = 0;
        binMaxId = 0;
    }
    bool fail = false;
    for (int b = bmin; b < binCount; b++) {
        int lastBinValue = binValue;
        int cnt = 0;
        for (int base = 0; base < k; base += 32) {
            int i = base + lane;
            bool hit = (i < k) && (s.raw[i] == b);
            cnt += __popc(__ballot_sync(MLD_FULL_MASK, hit));
        }
        binValue = cnt;
        if ((binValue > binMaxVal) && (binValue >= P.hist_min)) {
            binMaxVal = binValue;
            binMaxId = b;
        } else if (binValue < binMaxVal)
            break;
        if ((lastBinValue > 0) && (binValue == 0)) {
            fail = true;
            break;
        }
        // an empty bin that neither broke nor failed can only be followed by more empty bins
        if (binValue == 0) break;
    }
    if (fail || binMaxId < 0) return -1;
    double lowerBorder = binMaxId * P.bin_w - 0.0 * P.bin_w;   // :99
    double higherBorder = (binMaxId)*P.bin_w + 1.0 * P.bin_w;  // :100
    __syncwarp();
    return compact_slab(k, lane, s, false, [&](int, double, double, double z, int) {
        double d = fmin(z, 999.);
        return (d >= lowerBorder) && (d < higherBorder);  // :116
    });
}

// ---- A7: max spanning triangle -----------------------------------------------------------------
__device__ __forceinline__ D3 slab_pt(const WarpSlab& s, int i) { return D3{s.x[i], s.y[i], s.z[i]}; }

// linear pair index p (lexicographic over i<j) -> (i,j)
__device__ __forceinline__ void pair_from_index(int p, int n, int& i, int& j) {
    float fn = (float)(2 * n - 1);
    int ii = (int)((fn - sqrtf(fn * fn - 8.0f * (float)p)) * 0.5f);
    if (ii < 0) ii = 0;
    if (ii > n - 2) ii = n - 2;
    // row start S(i) = i*(2n-i-1)/2
    while (ii > 0 && (ii * (2 * n - ii - 1)) / 2 > p) ii--;
    while (ii < n - 2 && ((ii + 1) * (2 * n - ii - 2)) / 2 <= p) ii++;
    i = ii;
    j = p - (ii * (2 * n - ii - 1)) / 2 + ii + 1;
}

// returns false for the reference's `return false` sites; corners by slab index
__device__ bool max_spanning_triangle(int n, int lane, const WarpSlab& s, int& ci, int& cj, int& ck) {
    if (n < 3) return false;  // :44
    // farthest pair, strict '>' in lexicographic order == first maximum (:52-62)
    const int npairs = n * (n - 1) / 2;
    double best = -1.0;
    int bestp = 0x7fffffff;
    for (int p = lane; p < npairs; p += 32) {
        int i, j;
        pair_from_index(p, n, i, j);
        double dist = sqnorm3(slab_pt(s, i) - slab_pt(s, j));
        if (dist > best) {
            best = dist;
            bestp = p;
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        double ob = shfl_xor_d(best, m);
        int op = __shfl_xor_sync(MLD_FULL_MASK, bestp, m);
        if (ob > best || (ob == best && op < bestp)) {
            best = ob;
            bestp = op;
        }
    }
    if (best <= 0.0) return false;  // maxdist <= _distTreshold (== 0, bool ctor) (:65)
    int mi, mj;
    pair_from_index(bestp, n, mi, mj);
    // third corner: k in [0, n-2] (the last point is never eligible, :71), first maximum of d1+d2
    D3 pi = slab_pt(s, mi), pj = slab_pt(s, mj);
    double best2 = -1.0;
    int bestk = 0x7fffffff;
    for (int k = lane; k < n - 1; k += 32) {
        if (k == mi || k == mj) continue;
        D3 pk = slab_pt(s, k);
        double dist1 = sqnorm3(pk - pi);
        if (dist1 <= 0.0) continue;
        double dist2 = sqnorm3(pk - pj);
        if (dist2 <= 0.0) continue;
        double dist = dist1 + dist2;
        if (dist > best2) {
            best2 = dist;
            bestk = k;
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        double ob = shfl_xor_d(best2, m);
        int ok = __shfl_xor_sync(MLD_FULL_MASK, bestk, m);
        if (ob > best2 || (ob == best2 && ok < bestk)) {
            best2 = ob;
            bestk = ok;
        }
    }
    if (bestk == 0x7fffffff) return false;  // maxDist_k == -1 (:93)
    ci = mi;
    cj = mj;
    ck = bestk;
    return true;
}

__device__ void slab_z_range(int n, int lane, const WarpSlab& s, double& minZ, double& maxZ) {
    double lo = 1.7976931348623157e308, hi = -1.7976931348623157e308;
    for (int i = lane; i < n; i += 32) {
        double z = s.z[i];
        if (z < lo) lo = z;
        if (z > hi) hi = z;
    }
    minZ = warp_min_d(lo);
    maxZ = warp_max_d(hi);
}

// weighted centroid + scatter of slab[0..n); w_i = 1/|prior.n . p + prior.off| or 1. Every lane runs the SAME sequential sums in
// the reference's order (PCA.cpp:42-50, PlaneEstimationMEstimator.cpp:24-45) -- lane-strided partial sums would differ from the
// oracle and from the thread-per-feature kernels in the last bits, enough to flip a float-cast PCA ratio on a threshold; the
// warp path only sees the rare windows that overflow a thread's slab, so the redundant work does not matter.
__device__ void slab_weighted_scatter(int n, int lane, const WarpSlab& s, bool weighted, const Plane& prior, D3& center,
                                      double c[6]) {
    (void)lane;
    D3 acc = D3{0, 0, 0};
    double wsum = 0;
    for (int i = 0; i < n; i++) {
        D3 p = slab_pt(s, i);
        double w = weighted ? 1 / fabs(dot3(prior.n, p) + prior.off) : 1.0;
        acc = acc + p * w;
        wsum += w;
    }
    center = acc / wsum;
    c[0] = c[1] = c[2] = c[3] = c[4] = c[5] = 0;
    for (int i = 0; i < n; i++) {
        D3 p = slab_pt(s, i);
        double w = weighted ? 1 / fabs(dot3(prior.n, p) + prior.off) : 1.0;
        D3 d = p - center;
        c[0] += w * d.x * d.x; c[1] += w * d.x * d.y; c[2] += w * d.x * d.z;
        c[3] += w * d.y * d.y; c[4] += w * d.y * d.z; c[5] += w * d.z * d.z;
    }
}

// ---- A12: CalculateDepthSegmented ----------------------------------------------------------------
// corners9 (debug view, normally nullptr): the three triangle corners of a feature whose corner selection succeeded, camera frame
__device__ int depth_segmented(const DevParams& P, double u, double v, int n, int lane, const WarpSlab& s, double& depth_out,
                               double* corners9 = nullptr) {
    depth_out = -1;
    D3 c1{}, c2{}, c3{};
    if (!P.use_pca && P.use_tri_max) {
        int i, j, k;
        if (!max_spanning_triangle(n, lane, s, i, j, k)) return ST_TriangleNotPlanarInsufficientPoints;
        c1 = slab_pt(s, i); c2 = slab_pt(s, j); c3 = slab_pt(s, k);
    } else {
        if (n < 3) return ST_HistogramNoLocalMax;  // DepthEstimator.cpp:920-921
        c1 = slab_pt(s, 0); c2 = slab_pt(s, 1); c3 = slab_pt(s, 2);
    }
    if (corners9 != nullptr && lane == 0) {  // _points_triangle_corners (PlaneEstimationCalcMaxSpanningTriangle.cpp:27-33)
        corners9[0] = c1.x; corners9[1] = c1.y; corners9[2] = c1.z;
        corners9[3] = c2.x; corners9[4] = c2.y; corners9[5] = c2.z;
        corners9[6] = c3.x; corners9[7] = c3.y; corners9[8] = c3.z;
    }
    if (!P.use_pca && P.check_planar)
        if (!check_planar(c1, c2, c3, P.crossnorm_thr)) return ST_TriangleNotPlanar;

    D3 support = D3{0, 0, 0};
    D3 dir = viewing_ray(P, u, v);
    double depth;
    if (P.use_pca) {
        // Mono_LidarPipeline::PCA (PCA.cpp:42-62): mean, un-normalised scatter, ascending eigenvalues
        D3 mean;
        double c[6];
        Plane none{};
        slab_weighted_scatter(n, lane, s, false, none, mean, c);
        double w[3];
        D3 ev[3];
        eig3_sym_regs(c[0], c[1], c[2], c[3], c[4], c[5], w, ev);
        // sort ascending
        int i0 = 0, i1 = 1, i2 = 2, tmp;
        if (w[i1] < w[i0]) { tmp = i0; i0 = i1; i1 = tmp; }
        if (w[i2] < w[i1]) { tmp = i1; i1 = i2; i2 = tmp; }
        if (w[i1] < w[i0]) { tmp = i0; i0 = i1; i1 = tmp; }
        double ev1 = w[i0], ev2 = w[i1], ev3 = w[i2];
        float planarity = (float)((ev2 - ev1) / ev3);  // PCA.cpp:27-28
        float linearity = (float)((ev3 - ev2) / ev3);
        if (planarity < P.pca_2_1_rel_min) return ST_PcaIsCubic;
        if (linearity > P.pca_3_2_rel_max) return ST_PcaIsLine;
        if (ev3 < P.pca_3_abs_min) return ST_PcaIsPoint;
        D3 e0 = (i0 == 0) ? ev[0] : (i0 == 1 ? ev[1] : ev[2]);
        D3 normal = e0 / norm3(e0);
        Plane pl{normal, -dot3(normal, mean)};
        if (!line_plane(pl, support, dir, P.ortho_thr, depth)) return ST_PlaneViewrayNotOrthogonal;
    } else {
        Plane pl = plane_through(c1, c2, c3);
        if (!line_plane(pl, support, dir, P.ortho_thr, depth)) return ST_PlaneViewrayNotOrthogonal;
    }
    double minZ, maxZ;
    slab_z_range(n, lane, s, minZ, maxZ);
    int r = apply_tresholds(P, depth, minZ, maxZ);
    if (r) return r;
    if (depth < 0 && P.cut_behind) return ST_CornerBehindCamera;
    depth_out = depth;
    return ST_Success;
}

// ---- road path: R2 + R3/R4/R5 -------------------------------------------------------------------
__device__ int road_depth(const DevParams& P, double u, double v, int k2, int lane, const WarpSlab& s, const float* coeffs,
                          const unsigned int* __restrict__ inlier_bits, int old_status, double& depth_out) {
    depth_out = -1;
    const float a = coeffs[0], b = coeffs[1], c = coeffs[2], d = coeffs[3];
    // R2 gate: any neighbour farther than the threshold from the plane rejects the feature
    // (DepthEstimator.cpp:803-815). pcl::pointToPlaneDistance on a PointXYZ evaluates in float.
    bool far = false;
    for (int i = lane; i < k2; i += 32) {
        D3 p = slab_pt(s, i);
        double lx = ((P.Ri[0] * p.x + P.Ri[1] * p.y) + P.Ri[2] * p.z) + P.ti[0];
        double ly = ((P.Ri[3] * p.x + P.Ri[4] * p.y) + P.Ri[5] * p.z) + P.ti[1];
        double lz = ((P.Ri[6] * p.x + P.Ri[7] * p.y) + P.Ri[8] * p.z) + P.ti[2];
        float fx = (float)lx, fy = (float)ly, fz = (float)lz;
        float sd = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, fx), __fmul_rn(b, fy)), __fmul_rn(c, fz)), d);
        double distance = fabs((double)sd);
        if (distance > P.road_dist_thr) far = true;
    }
    if (__any_sync(MLD_FULL_MASK, far)) return old_status;
    // keep the plane inliers (GroundPlane::CheckPointInPlane on the raw index, :817)
    int n = compact_slab(k2, lane, s, true, [&](int, double, double, double, int raw) {
        return ((inlier_bits[raw >> 5] >> (raw & 31)) & 1u) != 0;
    });
    if (n < 3) return old_status;  // :827-829

    Plane pl;
    if (P.road_mode == ROAD_TRIANGLE) {
        int i, j, k;
        if (!max_spanning_triangle(n, lane, s, i, j, k)) return ST_RadiusSearchInsufficientPoints;
        // LinePlaneIntersectionCeckXZTreshold::Check
        double loX = 1.7976931348623157e308, hiX = -1.7976931348623157e308, loZ = loX, hiZ = hiX;
        for (int q = lane; q < n; q += 32) {
            double x = s.x[q], z = s.z[q];
            if (x < loX) loX = x;
            if (x > hiX) hiX = x;
            if (z < loZ) loZ = z;
            if (z > hiZ) hiZ = z;
        }
        loX = warp_min_d(loX); hiX = warp_max_d(hiX); loZ = warp_min_d(loZ); hiZ = warp_max_d(hiZ);
        double relation = (hiZ - loZ) / (hiX - loX);
        if (!(relation >= P.zx_min_rel)) return ST_InsufficientRoadPoints;
        pl = plane_through(slab_pt(s, i), slab_pt(s, j), slab_pt(s, k));
    } else {
        // PlaneEstimationMEstimator::EstimatePlane; prior = Hyperplane(normalized(a,b,c), d) in the
        // lidar frame applied to camera-frame points, as the reference does (DepthEstimator.cpp:286-292).
        // The last left-singular vector of [sqrt(w_i)(p_i - c)] is the eigenvector of the smallest
        // eigenvalue of sum w_i (p_i-c)(p_i-c)^T, solved in registers.
        Plane prior{normalized3(D3{(double)a, (double)b, (double)c}), (double)d};
        D3 center;
        double cv[6];
        slab_weighted_scatter(n, lane, s, P.road_mode == ROAD_MESTIMATOR, prior, center, cv);
        double w[3];
        D3 ev[3];
        eig3_sym_regs(cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], w, ev);
        int bi = 0;
        if (w[1] < w[bi]) bi = 1;
        if (w[2] < w[bi]) bi = 2;
        D3 nrm = normalized3((bi == 0) ? ev[0] : (bi == 1 ? ev[1] : ev[2]));
        pl = Plane{nrm, -dot3(nrm, center)};
    }
    // ray with swapped arguments (origin = direction, RoadDepthEstimatorMEstimator.cpp:52-53), no orthogonality gate
    D3 support = D3{0, 0, 0};
    D3 dir = viewing_ray(P, u, v);
    double depth;
    line_plane(pl, dir, support, 0.0, depth);
    double minZ, maxZ;
    slab_z_range(n, lane, s, minZ, maxZ);
    int r = apply_tresholds(P, depth, minZ, maxZ);
    if (r) return r;
    depth_out = depth;
    return ST_SuccessRoad;
}

// ---- per-feature driver (DepthEstimator.cpp:491-600) --------------------------------------------
__device__ __noinline__ void feature_depth(const DevParams& P, const MapCode& mc, const unsigned int* __restrict__ map,
                                           const float* __restrict__ pts, int stride_f, double u, double v, const float* plane_coeffs,
                                           const unsigned int* __restrict__ inlier_bits, int lane, const WarpSlab& s, const int KCAP,
                                           int& status_out, double& depth_out, double* corners9 = nullptr) {
    depth_out = -1;
    int k = gather_window(P, mc, map, pts, stride_f, u, v, P.hx1, P.hy1, lane, s, KCAP);
    if ((unsigned)k < (unsigned)P.count_min) {  // neighbors.size() < (uint)radiusSearch_count_min (:680)
        status_out = ST_RadiusSearchInsufficientPoints;
        return;
    }
    int status = ST_Unspecified;
    int n = k;
    if (P.use_hist) {
        n = histogram_segment(P, k, lane, s);
        if (n < 0) status = ST_HistogramNoLocalMax;
    }
    if (status != ST_HistogramNoLocalMax) {
        double depth;
        status = depth_segmented(P, u, v, n, lane, s, depth, corners9);
        if (status == ST_Success) {
            status_out = status;
            depth_out = depth;
            return;
        }
    }
    if (plane_coeffs != nullptr && P.road_mode != ROAD_NONE) {
        __syncwarp();
        int k2 = gather_window(P, mc, map, pts, stride_f, u, v, P.hx2, P.hy2, lane, s, KCAP);
        if ((unsigned)k2 < (unsigned)P.count_min) {
            status_out = ST_RadiusSearchInsufficientPoints;
            return;
        }
        double depth;
        status = road_depth(P, u, v, k2, lane, s, plane_coeffs, inlier_bits, status, depth);
        depth_out = depth;
    }
    status_out = status;
}

}  // namespace
