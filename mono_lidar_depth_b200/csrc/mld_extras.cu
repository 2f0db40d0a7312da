// mld_extras.cu -- small kernels around the hot path ("next" rows of SURVEY.md 8f):
//   * DepthCalculationStatistics: per-status counters of a result block
//     (reference: monolidar_fusion/src/DepthEstimator.cpp:1039-1090 LogDepthCalcStats,
//      include/monolidar_fusion/DepthCalculationStatistics.h) == histogram of the status array;
//   * matches_msg_depth_ros/FeaturePoint {float32 u, v, d} packing
//     (reference: matches_msg_depth_ros/msg/FeaturePoint.msg:1-5, tracklets_depth/src/tracklet_depth_module.cpp:209-259):
//     d = (float)depth, -1 for features without a depth.
#include "mld_common.cuh"
#include "mld_kernels.h"

namespace {

constexpr int NSTATUS = 21;  // DepthResultType 0..20

__global__ void status_histogram_kernel(const int* __restrict__ status, long long n, unsigned long long* __restrict__ hist) {
    __shared__ unsigned int sh[NSTATUS];
    if (threadIdx.x < NSTATUS) sh[threadIdx.x] = 0u;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int s = status[i];
        if (s >= 0 && s < NSTATUS) atomicAdd(&sh[s], 1u);
    }
    __syncthreads();
    if (threadIdx.x < NSTATUS && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

__global__ void pack_feature_points_kernel(const double* __restrict__ uv, const double* __restrict__ depth, long long n,
                                           float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 f = reinterpret_cast<const double2*>(uv)[i];
    out[i * 3 + 0] = (float)f.x;
    out[i * 3 + 1] = (float)f.y;
    out[i * 3 + 2] = (float)depth[i];
}

}  // namespace

cudaError_t mld_launch_status_histogram(const int* d_status, long long n, unsigned long long* d_hist21, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(d_hist21, 0, NSTATUS * sizeof(unsigned long long), stream);
    if (e != cudaSuccess || n <= 0) return e;
    const int blocks = (int)std::min<long long>(1184, (n + 255) / 256);
    status_histogram_kernel<<<blocks, 256, 0, stream>>>(d_status, n, d_hist21);
    return cudaGetLastError();
}

// 12-byte xyz records (host-packed, mld_process_frames_host) -> the float4 layout K1 streams
__global__ void unpack_xyz_kernel(const float* __restrict__ src, float4* __restrict__ dst, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
}
cudaError_t mld_launch_unpack_xyz(const float* d_xyz, float* d_out4, long long n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    unpack_xyz_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_xyz, reinterpret_cast<float4*>(d_out4), n);
    return cudaGetLastError();
}

cudaError_t mld_launch_pack_feature_points(const double* d_uv, const double* d_depth, long long n, float* d_out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    pack_feature_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_uv, d_depth, n, d_out);
    return cudaGetLastError();
}
