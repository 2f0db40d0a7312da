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
#include "mld_thread_helpers.cuh"

namespace {

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
    const unsigned int* occ = occs + frame * occ_words_per_frame(P.W, P.H);
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
