// mld_thread_helpers.cuh -- per-thread (one feature per thread) restatements of the reference routines,
// shared by the split gather / solve / road kernels (mld_feature_split.cu). One thread per feature: the window of a lidar feature holds 2-9 points, so a warp per feature idles most
// lanes and runs the scalar FP64 tail 32x redundantly (measured in round 1: ~800 warp instructions per feature).
#pragma once
#include "mld_common.cuh"
#include "mld_geometry.cuh"

namespace {

// TCAP = neighbours a thread can hold, TBT = threads (= features) per block
template <int TCAP_, int TBT_>
struct TSlabT {
    static constexpr int TCAP = TCAP_;
    static constexpr int TBT = TBT_;
    double* x;
    double* y;
    double* z;
    int* aux;  // raw indices during the gather, bin ids during the histogram
    __device__ __forceinline__ D3 pt(int i) const { return D3{x[i * TBT], y[i * TBT], z[i * TBT]}; }
    __device__ __forceinline__ void set(int i, const D3& p) const {
        x[i * TBT] = p.x;
        y[i * TBT] = p.y;
        z[i * TBT] = p.z;
    }
    __device__ __forceinline__ double& Z(int i) const { return z[i * TBT]; }
    __device__ __forceinline__ double& X(int i) const { return x[i * TBT]; }
    __device__ __forceinline__ int& A(int i) const { return aux[i * TBT]; }
};

// A5: window scan (reference order: rows outer, columns inner) + gather. Returns k, or -1 when the
// window holds more than TCAP points. inlier_mask (bit i = neighbour i is a plane inlier) is filled
// when inlier_bits != nullptr.
//
// Three phases, each a batch of independent loads: (1) one occupancy word per window row (rows of up
// to 17 pixels; wider rows walk further words), set bits -> pixel offsets in scan order; (2) the map
// cells of the occupied pixels -> raw point indices; (3) the points themselves -> FP64 camera frame.
constexpr int T_PAIRS = 8;  // row pairs (= occupancy words per tile column) fetched up front: windows of up to 15-16 rows in one batch

__device__ __forceinline__ unsigned int row_mask(unsigned int w, int base_px, int x0, int x1) {
    // keep the bits of word `w` (covering pixels base_px .. base_px+31) that lie in [x0, x1]
    int lo = x0 - base_px, hi = x1 - base_px;
    if (lo < 0) lo = 0;
    if (hi > 31) hi = 31;
    if (hi < lo) return 0u;
    return w & (0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (31 - hi));
}

// Occupied pixels of the window [x0, x1] x [y0, y1] in the reference's scan order (rows outer, columns inner,
// NeighborFinderPixel.cpp:78-92): emit(pixel offset) per occupied pixel. The occupancy words of the first two tile
// columns are fetched for all row pairs of a batch before any is used (independent loads); wider windows walk on.
__device__ __forceinline__ unsigned int occ_load(const unsigned int* p) { return __ldg(p); }

template <typename Emit>
__device__ __forceinline__ void occ_scan_window(const unsigned int* __restrict__ occ, int W, int x0, int x1, int y0, int y1, Emit emit) {
    const int tiles_x = occ_tiles_x(W);
    const int tx0 = x0 >> 4, tx1 = x1 >> 4;
    const int yp_last = y1 >> 1;
    for (int ypb = y0 >> 1; ypb <= yp_last; ypb += T_PAIRS) {
        unsigned int w0[T_PAIRS], w1[T_PAIRS];
#pragma unroll
        for (int r = 0; r < T_PAIRS; r++) {
            const int yp = ypb + r;  // rows 2 yp and 2 yp + 1: tile row yp >> 3, word yp & 7 of the tile
            const bool on = yp <= yp_last;
            const long long base = ((long long)(yp >> 3) * tiles_x + tx0) * 8 + (yp & 7);
            w0[r] = on ? occ_load(occ + base) : 0u;
            w1[r] = (on && tx1 > tx0) ? occ_load(occ + base + 8) : 0u;
        }
#pragma unroll
        for (int r = 0; r < T_PAIRS; r++) {
            const int yp = ypb + r;
            if (yp > yp_last) break;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int y = 2 * yp + h;
                if (y < y0 || y > y1) continue;
                const int sh = h << 4;
                unsigned int m = row_mask(((w0[r] >> sh) & 0xFFFFu) | (((w1[r] >> sh) & 0xFFFFu) << 16), tx0 << 4, x0, x1);
                int tx = tx0;
                while (true) {
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        emit(y * W + (tx << 4) + b);
                    }
                    tx += 2;  // next 32-pixel span of a wide row
                    if ((tx << 4) > x1) break;
                    const long long base = ((long long)(yp >> 3) * tiles_x + tx) * 8 + (yp & 7);
                    const unsigned int a = occ_load(occ + base), c = (tx + 1 <= tx1) ? occ_load(occ + base + 8) : 0u;
                    m = row_mask(((a >> sh) & 0xFFFFu) | (((c >> sh) & 0xFFFFu) << 16), tx << 4, x0, x1);
                }
            }
        }
    }
}

template <typename TSlab>
__device__ int t_gather_window(const DevParams& P, const MapCode& mc, const unsigned int* __restrict__ map,
                               const unsigned int* __restrict__ occ, const float* __restrict__ pts, int stride_f, double u,
                               double v, double hx, double hy, const TSlab& s, const unsigned int* __restrict__ inlier_bits,
                               unsigned int& inlier_mask) {
    inlier_mask = 0u;
    if (!(fabs(u) < 1e9) || !(fabs(v) < 1e9)) return 0;  // see mld_feature.cu: UB upstream, empty window here
    double leftEdgeX = fmax(u - hx, 0.);
    double rightEdgeX = fmin(u + hx, (double)(P.W - 1));
    double topEdgeY = fmax(v - hy, 0.);
    double bottomEdgeY = fmin(v + hy, (double)(P.H - 1));
    const int x0 = (int)leftEdgeX, x1 = (int)rightEdgeX, y0 = (int)topEdgeY, y1 = (int)bottomEdgeY;
    if (x1 < x0 || y1 < y0) return 0;
    int k = 0;
    // ---- phase 1: occupancy tiles -> pixel offsets (row-major order) ----
    occ_scan_window(occ, P.W, x0, x1, y0, y1, [&](int off) {
        if (k < TSlab::TCAP) s.A(k) = off;
        k++;
    });
    if (k > TSlab::TCAP) return -1;
    // ---- phase 2: map cells -> raw indices ----
#pragma unroll 4
    for (int i = 0; i < k; i++) {
        unsigned int cell = __ldg(map + s.A(i));
        s.A(i) = (int)map_cell_index(mc, cell);
    }
    // ---- phase 3: points -> camera frame ----
#pragma unroll 2
    for (int i = 0; i < k; i++) {
        int raw = s.A(i);
        float4 q = __ldg(reinterpret_cast<const float4*>(pts + (long long)raw * stride_f));
        s.set(i, lidar_to_cam(P, q.x, q.y, q.z));
        if (inlier_bits && ((inlier_bits[raw >> 5] >> (raw & 31)) & 1u)) inlier_mask |= 1u << i;
    }
    return k;
}

// A6: PointHistogram::FilterPointsMinDistBlob, sequential like the reference. Returns the segmented
// count (slab compacted in place, order kept) or -1.
template <typename TSlab>
__device__ int t_histogram_segment(const DevParams& P, int k, const TSlab& s) {
    int maxDist = 0;
    for (int i = 0; i < k; i++) {
        double d = fmin(s.Z(i), 999.);
        if (d > maxDist) maxDist = (int)ceil(d);  // HistogramPointDepth.cpp:38-41
    }
    int binCount = (int)((maxDist) / P.bin_w + 1);  // :43
    if (binCount <= 1) return -1;
    int bmin = binCount;
    for (int i = 0; i < k; i++) {
        double value = fmin(fmin(s.Z(i), 999.), 1e10);  // Histogram.cpp:29
        int b = (int)fmin(fabs(value / P.bin_w), (double)binCount - 1.);
        s.A(i) = b;
        bmin = min(bmin, b);
    }
    // first-local-maximum scan (:66-85); only the first run of occupied bins can decide it
    int binMaxId = -1, binMaxVal = -1, binValue = 0;
    if (bmin > 0 && 0 >= P.hist_min) {
        binMaxVal = 0;
        binMaxId = 0;
    }
    bool fail = false;
    for (int b = bmin; b < binCount; b++) {
        int lastBinValue = binValue;
        int cnt = 0;
        for (int i = 0; i < k; i++) cnt += (s.A(i) == b) ? 1 : 0;
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
        if (binValue == 0) break;
    }
    if (fail || binMaxId < 0) return -1;
    double lowerBorder = binMaxId * P.bin_w - 0.0 * P.bin_w;   // :99
    double higherBorder = (binMaxId)*P.bin_w + 1.0 * P.bin_w;  // :100
    int n = 0;
    for (int i = 0; i < k; i++) {
        D3 p = s.pt(i);
        double d = fmin(p.z, 999.);
        if ((d >= lowerBorder) && (d < higherBorder)) {  // :116
            if (n != i) s.set(n, p);
            n++;
        }
    }
    return n;
}

// A7: PlaneEstimationCalcMaxSpanningTriangle::CalculatePlaneCorners, sequential
template <typename TSlab>
__device__ bool t_max_spanning_triangle(int n, const TSlab& s, int& ci, int& cj, int& ck) {
    if (n < 3) return false;
    int mi = -1, mj = -1;
    double maxdist = -1;
    for (int i = 0; i < n - 1; i++) {
        D3 pi = s.pt(i);
        for (int j = i + 1; j < n; j++) {
            double dist = sqnorm3(pi - s.pt(j));
            if (dist > maxdist) {
                maxdist = dist;
                mi = i;
                mj = j;
            }
        }
    }
    if (maxdist <= 0.0) return false;
    D3 pi = s.pt(mi), pj = s.pt(mj);
    double maxdist2 = -1;
    int mk = -1;
    for (int k = 0; k < n - 1; k++) {  // the last point is never eligible (:71)
        if (k == mi || k == mj) continue;
        D3 pk = s.pt(k);
        double dist1 = sqnorm3(pk - pi);
        if (dist1 <= 0.0) continue;
        double dist2 = sqnorm3(pk - pj);
        if (dist2 <= 0.0) continue;
        double dist = dist1 + dist2;
        if (dist > maxdist2) {
            maxdist2 = dist;
            mk = k;
        }
    }
    if (mi == -1 || mj == -1 || mk == -1) return false;
    ci = mi;
    cj = mj;
    ck = mk;
    return true;
}

template <typename TSlab>
__device__ void t_z_range(int n, const TSlab& s, double& minZ, double& maxZ) {
    minZ = 1.7976931348623157e308;
    maxZ = -1.7976931348623157e308;
    for (int i = 0; i < n; i++) {
        double z = s.Z(i);
        if (z < minZ) minZ = z;
        if (z > maxZ) maxZ = z;
    }
}

// weighted centroid + scatter in the reference's sequential order
template <typename TSlab>
__device__ void t_weighted_scatter(int n, const TSlab& s, bool weighted, const Plane& prior, D3& center, double c[6]) {
    D3 acc = D3{0, 0, 0};
    double wsum = 0;
    for (int i = 0; i < n; i++) {
        D3 p = s.pt(i);
        double w = weighted ? 1 / fabs(dot3(prior.n, p) + prior.off) : 1.0;
        acc = acc + p * w;
        wsum += w;
    }
    center = acc / wsum;
    c[0] = c[1] = c[2] = c[3] = c[4] = c[5] = 0;
    for (int i = 0; i < n; i++) {
        D3 p = s.pt(i);
        double w = weighted ? 1 / fabs(dot3(prior.n, p) + prior.off) : 1.0;
        D3 d = p - center;
        c[0] += w * d.x * d.x; c[1] += w * d.x * d.y; c[2] += w * d.x * d.z;
        c[3] += w * d.y * d.y; c[4] += w * d.y * d.z; c[5] += w * d.z * d.z;
    }
}

// A12 (first half): corner selection of CalculateDepthSegmented (DepthEstimator.cpp:915-926).
// Returns 0 and the corner indices, or the failing status.
template <typename TSlab>
__device__ int t_select_corners(const DevParams& P, int n, const TSlab& s, int& ci, int& cj, int& ck) {
    ci = 0; cj = 1; ck = 2;
    if (!P.use_pca && P.use_tri_max) {
        if (!t_max_spanning_triangle(n, s, ci, cj, ck)) return ST_TriangleNotPlanarInsufficientPoints;
    } else {
        if (n < 3) return ST_HistogramNoLocalMax;
    }
    return 0;
}

// A12 (second half): planarity, viewing ray, plane intersection, thresholds (DepthEstimator.cpp:928-1036)
template <typename TSlab>
__device__ int t_depth_from_corners(const DevParams& P, double u, double v, int n, const TSlab& s, int ci, int cj, int ck,
                                    double& depth_out) {
    depth_out = -1;
    D3 c1 = s.pt(ci), c2 = s.pt(cj), c3 = s.pt(ck);
    if (!P.use_pca && P.check_planar)
        if (!check_planar(c1, c2, c3, P.crossnorm_thr)) return ST_TriangleNotPlanar;
    D3 support = D3{0, 0, 0};
    D3 dir = viewing_ray(P, u, v);
    double depth;
    if (P.use_pca) {
        D3 mean;
        double c[6];
        Plane none{};
        t_weighted_scatter(n, s, false, none, mean, c);
        double w[3];
        D3 ev[3];
        eig3_sym_regs(c[0], c[1], c[2], c[3], c[4], c[5], w, ev);
        int i0 = 0, i1 = 1, i2 = 2, tmp;
        if (w[i1] < w[i0]) { tmp = i0; i0 = i1; i1 = tmp; }
        if (w[i2] < w[i1]) { tmp = i1; i1 = i2; i2 = tmp; }
        if (w[i1] < w[i0]) { tmp = i0; i0 = i1; i1 = tmp; }
        double ev1 = w[i0], ev2 = w[i1], ev3 = w[i2];
        float planarity = (float)((ev2 - ev1) / ev3);
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
    t_z_range(n, s, minZ, maxZ);
    int r = apply_tresholds(P, depth, minZ, maxZ);
    if (r) return r;
    if (depth < 0 && P.cut_behind) return ST_CornerBehindCamera;
    depth_out = depth;
    return ST_Success;
}

// R3/R4/R5: road estimators on the n plane-inlier points held in the slab (RoadDepthEstimator*.cpp)
template <typename TSlab>
__device__ int t_road_estimate(const DevParams& P, double u, double v, int n, const TSlab& s, const float* coeffs, double& depth_out) {
    constexpr int TBT = TSlab::TBT;
    depth_out = -1;
    const float a = coeffs[0], b = coeffs[1], c = coeffs[2], d = coeffs[3];
    Plane pl;
    if (P.road_mode == ROAD_TRIANGLE) {
        int i, j, k;
        if (!t_max_spanning_triangle(n, s, i, j, k)) return ST_RadiusSearchInsufficientPoints;
        double loX = 1.7976931348623157e308, hiX = -1.7976931348623157e308, loZ = loX, hiZ = hiX;
        for (int q = 0; q < n; q++) {
            double x = s.x[q * TBT], z = s.z[q * TBT];
            if (x < loX) loX = x;
            if (x > hiX) hiX = x;
            if (z < loZ) loZ = z;
            if (z > hiZ) hiZ = z;
        }
        double relation = (hiZ - loZ) / (hiX - loX);
        if (!(relation >= P.zx_min_rel)) return ST_InsufficientRoadPoints;
        pl = plane_through(s.pt(i), s.pt(j), s.pt(k));
    } else {
        Plane prior{normalized3(D3{(double)a, (double)b, (double)c}), (double)d};
        D3 center;
        double cv[6];
        t_weighted_scatter(n, s, P.road_mode == ROAD_MESTIMATOR, prior, center, cv);
        double w[3];
        D3 ev[3];
        eig3_sym_regs(cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], w, ev);
        int bi = 0;
        if (w[1] < w[bi]) bi = 1;
        if (w[2] < w[bi]) bi = 2;
        D3 nrm = normalized3((bi == 0) ? ev[0] : (bi == 1 ? ev[1] : ev[2]));
        pl = Plane{nrm, -dot3(nrm, center)};
    }
    D3 support = D3{0, 0, 0};
    D3 dir = viewing_ray(P, u, v);
    double depth;
    line_plane(pl, dir, support, 0.0, depth);
    double minZ, maxZ;
    t_z_range(n, s, minZ, maxZ);
    int r = apply_tresholds(P, depth, minZ, maxZ);
    if (r) return r;
    depth_out = depth;
    return ST_SuccessRoad;
}

// R2 gate: |a x + b y + c z + d| of the lidar-frame point, in float like pcl::pointToPlaneDistance(PointXYZ, Vector4f)
__device__ __forceinline__ bool road_point_far(const DevParams& P, const D3& p, float a, float b, float c, float d) {
    double lx = ((P.Ri[0] * p.x + P.Ri[1] * p.y) + P.Ri[2] * p.z) + P.ti[0];
    double ly = ((P.Ri[3] * p.x + P.Ri[4] * p.y) + P.Ri[5] * p.z) + P.ti[1];
    double lz = ((P.Ri[6] * p.x + P.Ri[7] * p.y) + P.Ri[8] * p.z) + P.ti[2];
    float fx = (float)lx, fy = (float)ly, fz = (float)lz;
    float sd = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, fx), __fmul_rn(b, fy)), __fmul_rn(c, fz)), d);
    return fabs((double)sd) > P.road_dist_thr;  // DepthEstimator.cpp:814-815
}

// R2 + R3/R4/R5 on a slab holding all k2 neighbours of the wide window
template <typename TSlab>
__device__ int t_road_depth(const DevParams& P, double u, double v, int k2, const TSlab& s, const float* coeffs,
                            unsigned int inlier_mask, int old_status, double& depth_out) {
    depth_out = -1;
    const float a = coeffs[0], b = coeffs[1], c = coeffs[2], d = coeffs[3];
    for (int i = 0; i < k2; i++)
        if (road_point_far(P, s.pt(i), a, b, c, d)) return old_status;
    int n = 0;
    for (int i = 0; i < k2; i++) {
        if ((inlier_mask >> i) & 1u) {
            if (n != i) s.set(n, s.pt(i));
            n++;
        }
    }
    if (n < 3) return old_status;
    return t_road_estimate(P, u, v, n, s, coeffs, depth_out);
}

constexpr int ST_OVERFLOW = -1;

// order-preserving block-wide compaction: threads with `flag` append `value` to list; returns the count
template <int TBT>
__device__ int block_compact(bool flag, int value, short* list, int* warp_tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(MLD_FULL_MASK, flag);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < TBT / 32; w++) {
        const int c = warp_tot[w];
        if (w < warp) base += c;
        total += c;
    }
    if (flag) list[base + __popc(m & ((1u << lane) - 1u))] = (short)value;
    __syncthreads();
    return total;
}


}  // namespace
