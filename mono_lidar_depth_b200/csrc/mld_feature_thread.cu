// mld_feature_thread.cu -- K2 fast path: per-feature depth estimation, one THREAD per feature.
//
// Same reference routines as mld_feature.cu (DepthEstimator.cpp:491-600 and the helpers it calls);
// the difference is the mapping. On lidar data the search window of a feature holds a handful of
// points (KITTI shape: 7x10 pixels, ~2.4 x 5.3 pixels between returns => k ~ 2-8), so a warp per
// feature leaves most lanes idle and executes the scalar FP64 tail 32 times redundantly (ncu of the
// warp kernel: ~800 warp instructions per feature, 17 % occupancy). Here every lane owns a feature
// and runs the reference's sequential algorithm on a small per-thread slab in shared memory
// (interleaved [entry][thread], conflict free), which is also the order the oracle uses -- sums,
// first-maximum scans and tie-breaks are literally sequential.
//
// Features whose window holds more than TCAP points (dense clouds / large windows) are not handled
// here: they are appended to an overflow list and finished by the warp-per-feature kernel.
#include "mld_common.cuh"
#include "mld_geometry.cuh"
#include "mld_kernels.h"

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
constexpr int T_ROWS = 16;  // window rows whose occupancy words are fetched up front

__device__ __forceinline__ unsigned int row_mask(unsigned int w, int base_px, int x0, int x1) {
    // keep the bits of word `w` (covering pixels base_px .. base_px+31) that lie in [x0, x1]
    int lo = x0 - base_px, hi = x1 - base_px;
    if (lo < 0) lo = 0;
    if (hi > 31) hi = 31;
    if (hi < lo) return 0u;
    return w & (0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (31 - hi));
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
    const int pitch = occ_words_per_row(P.W);
    const int wj0 = x0 >> 4;
    int k = 0;
    // ---- phase 1: occupancy words -> pixel offsets (row-major order) ----
    for (int yb = y0; yb <= y1; yb += T_ROWS) {
        unsigned int w[T_ROWS];
#pragma unroll
        for (int r = 0; r < T_ROWS; r++) {
            const int y = yb + r;
            w[r] = (y <= y1) ? __ldg(occ + (long long)y * pitch + wj0) : 0u;
        }
#pragma unroll
        for (int r = 0; r < T_ROWS; r++) {
            const int y = yb + r;
            if (y > y1) break;
            unsigned int m = row_mask(w[r], wj0 << 4, x0, x1);
            int wj = wj0;
            while (true) {
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    if (k < TSlab::TCAP) s.A(k) = y * P.W + (wj << 4) + b;
                    k++;
                }
                wj += 2;  // next non-overlapping 32-pixel span of a wide row
                if ((wj << 4) > x1) break;
                m = row_mask(__ldg(occ + (long long)y * pitch + wj), wj << 4, x0, x1);
            }
        }
    }
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

// R2 + R3/R4/R5
template <typename TSlab>
__device__ int t_road_depth(const DevParams& P, double u, double v, int k2, const TSlab& s, const float* coeffs,
                            unsigned int inlier_mask, int old_status, double& depth_out) {
    depth_out = -1;
    const float a = coeffs[0], b = coeffs[1], c = coeffs[2], d = coeffs[3];
    for (int i = 0; i < k2; i++) {
        D3 p = s.pt(i);
        double lx = ((P.Ri[0] * p.x + P.Ri[1] * p.y) + P.Ri[2] * p.z) + P.ti[0];
        double ly = ((P.Ri[3] * p.x + P.Ri[4] * p.y) + P.Ri[5] * p.z) + P.ti[1];
        double lz = ((P.Ri[6] * p.x + P.Ri[7] * p.y) + P.Ri[8] * p.z) + P.ti[2];
        float fx = (float)lx, fy = (float)ly, fz = (float)lz;
        float sd = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, fx), __fmul_rn(b, fy)), __fmul_rn(c, fz)), d);
        if (fabs((double)sd) > P.road_dist_thr) return old_status;  // DepthEstimator.cpp:814-815
    }
    int n = 0;
    for (int i = 0; i < k2; i++) {
        if ((inlier_mask >> i) & 1u) {
            if (n != i) s.set(n, s.pt(i));
            n++;
        }
    }
    if (n < 3) return old_status;
    Plane pl;
    if (P.road_mode == ROAD_TRIANGLE) {
        int i, j, k;
        if (!t_max_spanning_triangle(n, s, i, j, k)) return ST_RadiusSearchInsufficientPoints;
        double loX = 1.7976931348623157e308, hiX = -1.7976931348623157e308, loZ = loX, hiZ = hiX;
        for (int q = 0; q < n; q++) {
            double x = s.X(q), z = s.Z(q);
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

// Phases (a block owns TBT consecutive features of one frame; between phases the surviving features are
// compacted onto the low lanes so that the expensive later phases run on dense warps):
//   P1 every thread: window scan + gather of its own feature          -> status 2 / overflow / survivor
//   P2 survivors:    histogram segmentation + corner selection        -> status 3 / 9 / survivor
//   P3 survivors:    planarity, ray/plane intersection, thresholds    -> final status of the normal path
//   P4 (plane given) features without Success: road path (wide window, plane gate, M-estimator ...)
template <int TCAP, int TBT>
__global__ void __launch_bounds__(TBT)
feature_depth_thread_kernel(DevParams P, MapCode mc, const float* __restrict__ pts, int stride_f, long long pitch_pts,
                            const unsigned int* __restrict__ maps, const unsigned int* __restrict__ occs,
                            const double* __restrict__ uv, int F, double* __restrict__ depth, int* __restrict__ status,
                            const float* __restrict__ plane_coeffs, const unsigned int* __restrict__ inlier_bits,
                            long long inlier_words_per_frame, int* __restrict__ overflow_list, int* __restrict__ overflow_count) {
    using TSlab = TSlabT<TCAP, TBT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sx = reinterpret_cast<double*>(smem_raw);
    double* sy = sx + TCAP * TBT;
    double* sz = sy + TCAP * TBT;
    int* saux = reinterpret_cast<int*>(sz + TCAP * TBT);
    __shared__ double s_u[TBT], s_v[TBT], s_dp[TBT];
    __shared__ short s_list[TBT];
    __shared__ short s_cnt[TBT];
    __shared__ signed char s_st[TBT], s_ci[TBT], s_cj[TBT], s_ck[TBT];
    __shared__ int s_wtot[TBT / 32];

    const int tid = threadIdx.x;
    const int fi = blockIdx.x * TBT + tid;
    const bool valid = fi < F;
    const long long frame = blockIdx.y;
    const long long obase = frame * (long long)F + (long long)blockIdx.x * TBT;  // global id of this block's feature 0
    const float* fp = pts + frame * pitch_pts * (long long)stride_f;
    const unsigned int* map = maps + frame * (long long)P.W * (long long)P.H;
    const unsigned int* occ = occs + frame * (long long)occ_words_per_row(P.W) * (long long)P.H;
    const float* pc = plane_coeffs ? plane_coeffs + frame * 4 : nullptr;
    const unsigned int* bits = inlier_bits ? inlier_bits + frame * inlier_words_per_frame : nullptr;
    const bool road = pc != nullptr && P.road_mode != ROAD_NONE;
    auto slab_of = [&](int owner) { return TSlab{sx + owner, sy + owner, sz + owner, saux + owner}; };

    if (P.set_all_zero) {  // DepthEstimator.cpp:448-453
        if (valid) {
            status[obase + tid] = 1;
            depth[obase + tid] = -1;
        }
        return;
    }

    // ---- P1: gather ----
    bool surv = false;
    s_st[tid] = ST_Unspecified;
    s_dp[tid] = -1;
    if (valid) {
        const double2 f2 = __ldg(reinterpret_cast<const double2*>(uv) + obase + tid);
        s_u[tid] = f2.x;
        s_v[tid] = f2.y;
        unsigned int mask;
        const int k = t_gather_window(P, mc, map, occ, fp, stride_f, f2.x, f2.y, P.hx1, P.hy1, slab_of(tid), nullptr, mask);
        if (k < 0) {
            s_st[tid] = ST_OVERFLOW;
        } else if ((unsigned)k < (unsigned)P.count_min) {  // neighbors.size() < (uint)radiusSearch_count_min (:680)
            s_st[tid] = ST_RadiusSearchInsufficientPoints;
        } else {
            s_cnt[tid] = (short)k;
            surv = true;
        }
    }
    int n1 = block_compact<TBT>(surv, tid, s_list, s_wtot);

    // ---- P2: histogram + corner selection ----
    surv = false;
    int owner = -1;
    if (tid < n1) {
        owner = s_list[tid];
        const TSlab s = slab_of(owner);
        int n = s_cnt[owner];
        int st = ST_Unspecified;
        if (P.use_hist) {
            n = t_histogram_segment(P, n, s);
            if (n < 0) st = ST_HistogramNoLocalMax;
        }
        if (st != ST_HistogramNoLocalMax) {
            int ci, cj, ck;
            st = t_select_corners(P, n, s, ci, cj, ck);
            if (st == 0) {
                s_cnt[owner] = (short)n;
                s_ci[owner] = (signed char)ci;
                s_cj[owner] = (signed char)cj;
                s_ck[owner] = (signed char)ck;
                surv = true;
            }
        }
        if (!surv) s_st[owner] = (signed char)st;
    }
    __syncthreads();
    int n2 = block_compact<TBT>(surv, owner, s_list, s_wtot);

    // ---- P3: geometry tail ----
    if (tid < n2) {
        owner = s_list[tid];
        double dp;
        const int st = t_depth_from_corners(P, s_u[owner], s_v[owner], (int)s_cnt[owner], slab_of(owner), (int)s_ci[owner],
                                            (int)s_cj[owner], (int)s_ck[owner], dp);
        s_st[owner] = (signed char)st;
        s_dp[owner] = dp;
    }
    __syncthreads();

    // ---- P4: road path for features the normal path could not solve (DepthEstimator.cpp:579-597) ----
    if (road) {
        const int st0 = s_st[tid];
        const bool want = valid && st0 != ST_Success && st0 != ST_OVERFLOW && st0 != ST_RadiusSearchInsufficientPoints;
        const int n3 = block_compact<TBT>(want, tid, s_list, s_wtot);
        if (tid < n3) {
            owner = s_list[tid];
            const TSlab s = slab_of(owner);
            unsigned int mask;
            const int k2 = t_gather_window(P, mc, map, occ, fp, stride_f, s_u[owner], s_v[owner], P.hx2, P.hy2, s, bits, mask);
            if (k2 < 0) {
                s_st[owner] = ST_OVERFLOW;
            } else if ((unsigned)k2 < (unsigned)P.count_min) {
                s_st[owner] = ST_RadiusSearchInsufficientPoints;
                s_dp[owner] = -1;
            } else {
                double dp;
                const int st = t_road_depth(P, s_u[owner], s_v[owner], k2, s, pc, mask, (int)s_st[owner], dp);
                s_st[owner] = (signed char)st;
                s_dp[owner] = dp;
            }
        }
        __syncthreads();
    }

    // ---- results ----
    if (valid) {
        const int st = s_st[tid];
        const long long o = obase + tid;
        if (st == ST_OVERFLOW) {
            const int slot = atomicAdd(overflow_count, 1);
            overflow_list[slot] = (int)o;  // finished by the warp-per-feature kernel
        } else {
            status[o] = st;
            depth[o] = (st == ST_Success || st == ST_SuccessRoad) ? s_dp[tid] : -1.0;
        }
    }
}

template <int TCAP, int TBT>
cudaError_t launch_thread(const DevParams& P, const MapCode& mc, const float* d_pts, int stride_f, long long pitch_pts,
                          const unsigned int* d_maps, const unsigned int* d_occ, const double* d_uv, int F, double* d_depth,
                          int* d_status, const float* d_plane_coeffs, const unsigned int* d_inlier_bits, long long words_per_frame,
                          int nframes, int* d_overflow_list, int* d_overflow_count, cudaStream_t stream) {
    constexpr size_t smem = (size_t)TCAP * TBT * (3 * sizeof(double) + sizeof(int));
    dim3 grid((unsigned)((F + TBT - 1) / TBT), (unsigned)nframes);
    feature_depth_thread_kernel<TCAP, TBT><<<grid, TBT, smem, stream>>>(P, mc, d_pts, stride_f, pitch_pts, d_maps, d_occ, d_uv, F,
                                                                       d_depth, d_status, d_plane_coeffs, d_inlier_bits,
                                                                       words_per_frame, d_overflow_list, d_overflow_count);
    return cudaGetLastError();
}

}  // namespace

// two builds: MLD_T_TCAP neighbours x MLD_T_TBT features per block when only the normal window is scanned,
// 24 x 64 when the road path (window scale 2.0 x 1.5) may run
#ifndef MLD_T_TCAP
#define MLD_T_TCAP 12
#endif
#ifndef MLD_T_TBT
#define MLD_T_TBT 128
#endif
int mld_thread_feature_capacity(int road) { return road ? 24 : MLD_T_TCAP; }

cudaError_t mld_configure_feature_depth_thread(void) {
    cudaError_t e = cudaFuncSetAttribute(feature_depth_thread_kernel<MLD_T_TCAP, MLD_T_TBT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         MLD_T_TCAP * MLD_T_TBT * 28);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(feature_depth_thread_kernel<24, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 64 * 28);
}

cudaError_t mld_launch_feature_depth_thread(const DevParams& P, const MapCode& mc, const float* d_pts, int stride_f,
                                            long long pitch_pts, const unsigned int* d_maps, const unsigned int* d_occ,
                                            const double* d_uv, int F, double* d_depth, int* d_status, const float* d_plane_coeffs,
                                            const unsigned int* d_inlier_bits, long long words_per_frame, int nframes,
                                            int* d_overflow_list, int* d_overflow_count, cudaStream_t stream) {
    if (F <= 0 || nframes <= 0) return cudaSuccess;
    const bool road = d_plane_coeffs != nullptr && P.road_mode != ROAD_NONE;
    if (road)
        return launch_thread<24, 64>(P, mc, d_pts, stride_f, pitch_pts, d_maps, d_occ, d_uv, F, d_depth, d_status, d_plane_coeffs,
                                     d_inlier_bits, words_per_frame, nframes, d_overflow_list, d_overflow_count, stream);
    return launch_thread<MLD_T_TCAP, MLD_T_TBT>(P, mc, d_pts, stride_f, pitch_pts, d_maps, d_occ, d_uv, F, d_depth, d_status, d_plane_coeffs,
                                  d_inlier_bits, words_per_frame, nframes, d_overflow_list, d_overflow_count, stream);
}
