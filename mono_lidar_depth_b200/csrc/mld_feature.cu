// mld_feature.cu -- K2/K3: per-feature depth estimation, one warp per feature.
//
// Replaces (reference, /root/reference/monolidar_fusion/src unless noted):
//   DepthEstimator::CalculateDepth (batch loop + per-feature driver)   DepthEstimator.cpp:429-600
//   DepthEstimator::CalculateNeighbors                                 DepthEstimator.cpp:636-684
//   NeighborFinderPixel::getNeighbors                                  NeighborFinderPixel.cpp:60-95
//   NeighborFinderBase::getNeighbors                                   NeighborFinderBase.cpp:15-27
//   PointHistogram::FilterPointsMinDistBlob / Histogram::AddElement    HistogramPointDepth.cpp:15-123, Histogram.cpp:19-33
//   PlaneEstimationCalcMaxSpanningTriangle::CalculatePlaneCorners      PlaneEstimationCalcMaxSpanningTriangle.cpp:37-145
//   PlaneEstimationCheckPlanar::CheckPlanar                            PlaneEstimationCheckPlanar.cpp:18-44
//   CameraPinhole::getViewingRays                                      include/monolidar_fusion/camera_pinhole.h:52-69
//   LinePlaneIntersection{Base,Normal,OrthogonalTreshold}              LinePlaneIntersection*.cpp
//   TresholdDepthGlobal::CheckInDepth / TresholdDepthLocal::CheckInBounds
//   Mono_LidarPipeline::PCA                                            PCA.cpp:11-62
//   DepthEstimator::CalculateDepthSegmentationPlane                    DepthEstimator.cpp:782-900
//   RoadDepthEstimator{MEstimator,LeastSquares,MaxSpanningTriangle}    RoadDepthEstimator*.cpp
//   PlaneEstimationMEstimator::EstimatePlane                           PlaneEstimationMEstimator.cpp:18-55
//
// Mapping. A warp owns one feature. Lanes stride the search window in the reference's row-major
// scan order; ballot + popc compaction keeps that order (it decides ties in the farthest-pair
// search). Neighbour points are re-derived from the float cloud in FP64 (bit-identical to K1) and
// parked in a per-warp shared-memory slab; the histogram is evaluated with ballots over bin ids
// (only the first run of occupied bins can decide the reference's sequential scan), the
// farthest-pair / third-corner searches are lane-parallel arg-max reductions with the reference's
// first-wins tie-break, and the scalar geometry is evaluated redundantly by all lanes so every
// branch is warp-uniform.
//
// HBM traffic per feature: 16 B read (u,v), 12 B written (depth f64 + status i32). Window reads
// (<= 4*70 B) and the neighbour gather hit L2 (K1 just streamed the same frame).
#include "mld_common.cuh"
#include "mld_geometry.cuh"
#include "mld_kernels.h"
#include "mld_feature_warp.cuh"

namespace {

template <int KCAP, int K2_WARPS>
__global__ void __launch_bounds__(K2_WARPS * 32)
feature_depth_kernel(DevParams P, MapCode mc, const float* __restrict__ pts, int stride_f, long long pitch_pts,
                     const unsigned int* __restrict__ maps, const double* __restrict__ uv, int F,
                     double* __restrict__ depth, int* __restrict__ status, const float* __restrict__ plane_coeffs,
                     const unsigned int* __restrict__ inlier_bits, long long inlier_words_per_frame,
                     const int* __restrict__ list, const int* __restrict__ list_count, double* __restrict__ corners) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sx = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 3 * KCAP;
    WarpSlab s{sx, sx + KCAP, sx + 2 * KCAP,
               reinterpret_cast<int*>(reinterpret_cast<double*>(smem_raw) + (size_t)K2_WARPS * 3 * KCAP) + (size_t)warp * KCAP};

    // direct mode: grid = (ceil(F / warps), frames), one feature per warp.
    // list mode (overflow of the thread-per-feature kernel): warps stride a list of global feature ids.
    long long item, item_end, item_step;
    if (list != nullptr) {
        item = (long long)blockIdx.x * K2_WARPS + warp;
        item_end = *list_count;
        item_step = (long long)gridDim.x * K2_WARPS;
    } else {
        const int fi = blockIdx.x * K2_WARPS + warp;
        if (fi >= F) return;
        item = (long long)blockIdx.y * F + fi;
        item_end = item + 1;
        item_step = 1;
    }
    for (; item < item_end; item += item_step) {
        const long long o = list ? (long long)list[item] : item;
        const long long frame = o / F;
        const float* fp = pts + frame * pitch_pts * (long long)stride_f;
        const unsigned int* map = maps + frame * (long long)P.W * (long long)P.H;
        if (P.set_all_zero) {  // DepthEstimator.cpp:448-453
            if (lane == 0) {
                status[o] = 1;
                depth[o] = -1;
            }
            continue;
        }
        double u = uv[o * 2], v = uv[o * 2 + 1];
        const float* pc = plane_coeffs ? plane_coeffs + frame * 4 : nullptr;
        const unsigned int* bits = inlier_bits ? inlier_bits + frame * inlier_words_per_frame : nullptr;
        int st;
        double dp;
        feature_depth(P, mc, map, fp, stride_f, u, v, pc, bits, lane, s, KCAP, st, dp, corners ? corners + o * 9 : nullptr);
        if (lane == 0) {
            status[o] = st;
            depth[o] = (st == ST_Success || st == ST_SuccessRoad) ? dp : -1.0;
        }
        __syncwarp();
    }
}

// debug: neighbour list of one feature in scan order (raw indices)
__global__ void neighbors_debug_kernel(DevParams P, MapCode mc, const unsigned int* __restrict__ map, double u, double v, double hx,
                                       double hy, int* __restrict__ out, int cap, int* __restrict__ k_out) {
    const int lane = threadIdx.x;
    if (!(fabs(u) < 1e9) || !(fabs(v) < 1e9)) {
        if (lane == 0) *k_out = 0;
        return;
    }
    int x0 = (int)fmax(u - hx, 0.), x1 = (int)fmin(u + hx, (double)(P.W - 1));
    int y0 = (int)fmax(v - hy, 0.), y1 = (int)fmin(v + hy, (double)(P.H - 1));
    int wc = x1 - x0 + 1, wr = y1 - y0 + 1;
    int k = 0;
    if (wc > 0 && wr > 0) {
        int area = wc * wr;
        const unsigned lt = lanemask_lt();
        for (int base = 0; base < area; base += 32) {
            int idx = base + lane;
            unsigned int cell = MLD_EMPTY;
            if (idx < area) {
                int ry = idx / wc, rx = idx - (idx / wc) * wc;
                cell = map[(long long)(y0 + ry) * P.W + (x0 + rx)];
            }
            const bool hit = (idx < area) && map_cell_valid(mc, cell);
            unsigned m = __ballot_sync(MLD_FULL_MASK, hit);
            if (hit) {
                int pos = k + __popc(m & lt);
                if (pos < cap) out[pos] = (int)map_cell_index(mc, cell);
            }
            k += __popc(m);
        }
    }
    if (lane == 0) *k_out = k;
}

template <int KCAP, int K2_WARPS>
cudaError_t launch_feature(const DevParams& P, const MapCode& mc, const float* d_pts, int stride_f, long long pitch_pts,
                           const unsigned int* d_maps, const double* d_uv, int F, double* d_depth, int* d_status,
                           const float* d_plane_coeffs, const unsigned int* d_inlier_bits, long long words_per_frame,
                           int nframes, const int* d_list, const int* d_list_count, int list_blocks, cudaStream_t stream, double* d_corners) {
    constexpr size_t smem = (size_t)K2_WARPS * KCAP * (3 * sizeof(double) + sizeof(int));
    dim3 grid = d_list ? dim3((unsigned)list_blocks, 1) : dim3((unsigned)((F + K2_WARPS - 1) / K2_WARPS), (unsigned)nframes);
    feature_depth_kernel<KCAP, K2_WARPS><<<grid, K2_WARPS * 32, smem, stream>>>(
        P, mc, d_pts, stride_f, pitch_pts, d_maps, d_uv, F, d_depth, d_status, d_plane_coeffs, d_inlier_bits, words_per_frame,
        d_list, d_list_count, d_corners);
    return cudaGetLastError();
}

template <int KCAP, int K2_WARPS>
cudaError_t configure_feature() {
    constexpr size_t smem = (size_t)K2_WARPS * KCAP * (3 * sizeof(double) + sizeof(int));
    return cudaFuncSetAttribute(feature_depth_kernel<KCAP, K2_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

}  // namespace

int mld_feature_capacity_for(int max_area) {
    if (max_area <= 96) return 96;
    if (max_area <= 256) return 256;
    if (max_area <= 1024) return 1024;
    return -1;
}

cudaError_t mld_launch_feature_depth(const DevParams& P, const MapCode& mc, int kcap, const float* d_pts, int stride_f,
                                     long long pitch_pts, const unsigned int* d_maps, const double* d_uv, int F, double* d_depth,
                                     int* d_status, const float* d_plane_coeffs, const unsigned int* d_inlier_bits,
                                     long long words_per_frame, int nframes, const int* d_list, const int* d_list_count,
                                     int list_blocks, cudaStream_t stream, double* d_corners) {
    if (F <= 0 || nframes <= 0) return cudaSuccess;
    if (d_list != nullptr) {
        // list mode (overflow pass beside the fused launches): 2-warp blocks. A 256-thread block of this 126-register kernel needs
        // half of an SM's register file at once and was only ever placed in the tail of the co-running fused launch, so the
        // slot's "done" event -- which the launch after next waits for -- came a whole launch late (7 % of the step).
        const int blocks = list_blocks * 4;
        switch (kcap) {
            case 96:
                return launch_feature<96, 2>(P, mc, d_pts, stride_f, pitch_pts, d_maps, d_uv, F, d_depth, d_status, d_plane_coeffs,
                                             d_inlier_bits, words_per_frame, nframes, d_list, d_list_count, blocks, stream, d_corners);
            case 256:
                return launch_feature<256, 2>(P, mc, d_pts, stride_f, pitch_pts, d_maps, d_uv, F, d_depth, d_status, d_plane_coeffs,
                                              d_inlier_bits, words_per_frame, nframes, d_list, d_list_count, blocks, stream, d_corners);
            case 1024:
                return launch_feature<1024, 2>(P, mc, d_pts, stride_f, pitch_pts, d_maps, d_uv, F, d_depth, d_status, d_plane_coeffs,
                                               d_inlier_bits, words_per_frame, nframes, d_list, d_list_count, blocks, stream, d_corners);
            default:
                return cudaErrorInvalidValue;
        }
    }
    switch (kcap) {
        case 96:
            return launch_feature<96, 8>(P, mc, d_pts, stride_f, pitch_pts, d_maps, d_uv, F, d_depth, d_status, d_plane_coeffs,
                                         d_inlier_bits, words_per_frame, nframes, d_list, d_list_count, list_blocks, stream, d_corners);
        case 256:
            return launch_feature<256, 8>(P, mc, d_pts, stride_f, pitch_pts, d_maps, d_uv, F, d_depth, d_status, d_plane_coeffs,
                                          d_inlier_bits, words_per_frame, nframes, d_list, d_list_count, list_blocks, stream, d_corners);
        case 1024:
            return launch_feature<1024, 4>(P, mc, d_pts, stride_f, pitch_pts, d_maps, d_uv, F, d_depth, d_status, d_plane_coeffs,
                                           d_inlier_bits, words_per_frame, nframes, d_list, d_list_count, list_blocks, stream, d_corners);
        default:
            return cudaErrorInvalidValue;
    }
}

// opt in to the dynamic shared memory the chosen variant needs (once per device, at mld_create)
cudaError_t mld_configure_feature_depth(int kcap) {
    switch (kcap) {
        case 96: return configure_feature<96, 8>();
        case 256: return configure_feature<256, 8>();
        case 1024: {
            cudaError_t e = configure_feature<1024, 4>();
            return e != cudaSuccess ? e : configure_feature<1024, 2>();  // 57 KB: the only 2-warp variant past the 48 KB default
        }
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t mld_launch_neighbors_debug(const DevParams& P, const MapCode& mc, const unsigned int* d_map, double u, double v,
                                       double hx, double hy, int* d_out, int cap, int* d_k, cudaStream_t stream) {
    neighbors_debug_kernel<<<1, 32, 0, stream>>>(P, mc, d_map, u, v, hx, hy, d_out, cap, d_k);
    return cudaGetLastError();
}
