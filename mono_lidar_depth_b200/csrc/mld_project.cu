// mld_project.cu -- K1: lidar -> camera transform, pinhole projection, frustum cull and
// first-point-wins pixel map.
//
// Replaces (reference, /root/reference/monolidar_fusion):
//   Transform_Cloud_LidarToCamera                 src/DepthEstimator.cpp:156-217
//   CameraPinhole::getImagePoints                 include/monolidar_fusion/camera_pinhole.h:85-97
//   NeighborFinderPixel::InitializeLidarProjection src/NeighborFinderPixel.cpp:29-58
//
// The reference walks the visible points in cloud order and writes a pixel only while it is still
// empty and the point has z_cam > 0 (NeighborFinderPixel.cpp:51): the winner of a pixel is the point
// with the smallest raw index among those that qualify. That is an order-independent reduction, so
// it is done with atomicMin on the (epoch-tagged) raw index -- deterministic and bit-exact whatever
// the thread schedule.
//
// Roofline: HBM. 16 B (float4) per point are read once, streaming; ~80 % of a 360-degree sweep
// never reaches the image, so each point first goes through an FP32 pre-filter (5 linear forms with
// rigorous error bounds, ~25 FP32 ops) and only the survivors pay the exact FP64 transform and the
// two FP64 divisions whose truncated results index the map.
#include "mld_common.cuh"
#include "mld_kernels.h"
#include "mld_project.cuh"

namespace {

__global__ void __launch_bounds__(K1_THREADS, MLD_K1_MINBLOCKS)
project_scatter_kernel(DevParams P, MapCode mc, const float* __restrict__ pts, int stride_f, int n, long long pitch_pts,
                       unsigned int* __restrict__ maps, unsigned int* __restrict__ occ) {
    k1_tile(P, mc, pts, stride_f, n, pitch_pts, maps, occ, blockIdx.y, (int)blockIdx.x);
}

// debug view: Transform_Cloud_LidarToCamera's visibility cull (no z > 0 test, DepthEstimator.cpp:184-207)
// and the camera-frame coordinates (_points_cs_camera). Not on the hot path.
__global__ void visible_debug_kernel(DevParams P, const float* __restrict__ pts, int stride_f, long long n,
                                     unsigned char* __restrict__ visible, double* __restrict__ cam) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = pts[i * stride_f], y = pts[i * stride_f + 1], z = pts[i * stride_f + 2];
    int px, py;
    if (visible) visible[i] = project_pixel(P, x, y, z, false, px, py) ? 1 : 0;
    if (cam) {
        D3 c = lidar_to_cam(P, x, y, z);
        cam[i * 3] = c.x;
        cam[i * 3 + 1] = c.y;
        cam[i * 3 + 2] = c.z;
    }
}

// camera-frame coordinates of selected raw points (getCloudRansacPlane: the plane's inliers, DepthEstimator.cpp:294-308)
__global__ void points_camera_indexed_kernel(DevParams P, const float* __restrict__ pts, int stride_f, long long n, const int* __restrict__ idx,
                                             long long n_idx, double* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_idx) return;
    const long long r = idx[i];
    D3 c = D3{__longlong_as_double(0x7ff8000000000000ll), __longlong_as_double(0x7ff8000000000000ll), __longlong_as_double(0x7ff8000000000000ll)};
    if (r >= 0 && r < n) c = lidar_to_cam(P, pts[r * stride_f], pts[r * stride_f + 1], pts[r * stride_f + 2]);
    out[i * 3] = c.x;
    out[i * 3 + 1] = c.y;
    out[i * 3 + 2] = c.z;
}

// ---- visible-order views: _pointIndex, _points_cs_image_visible, getPointDepthCamVisible ----
// Transform_Cloud_LidarToCamera compacts the visible points in cloud order (DepthEstimator.cpp:189-207). On the GPU
// that is an order-preserving stream compaction: per-block counts, an exclusive scan of the counts, and a second
// pass that ranks the visible points inside each block with warp ballots. Off the hot path (debug publishers).
constexpr int VC_THREADS = 256;
constexpr int VC_PPT = 4;  // thread t of a block owns points base + 4 t .. + 3 (contiguous, so ranks follow cloud order)

__device__ __forceinline__ int vc_flags(const DevParams& P, const float* __restrict__ pts, int stride_f, long long n, long long first,
                                        double (&u)[VC_PPT], double (&v)[VC_PPT], double (&zc)[VC_PPT]) {
    int m = 0;
#pragma unroll
    for (int j = 0; j < VC_PPT; j++) {
        const long long i = first + j;
        if (i >= n) break;
        const float* q = pts + i * stride_f;
        D3 c;
        if (project_uv(P, q[0], q[1], q[2], false, u[j], v[j], c)) {  // no z > 0 test here (DepthEstimator.cpp:184-207)
            m |= 1 << j;
            zc[j] = c.z;
        }
    }
    return m;
}

__global__ void __launch_bounds__(VC_THREADS)
visible_count_kernel(DevParams P, const float* __restrict__ pts, int stride_f, long long n, unsigned int* __restrict__ block_counts) {
    __shared__ int s_w[VC_THREADS / 32];
    double u[VC_PPT], v[VC_PPT], zc[VC_PPT];
    const long long first = ((long long)blockIdx.x * VC_THREADS + threadIdx.x) * VC_PPT;
    int c = __popc(vc_flags(P, pts, stride_f, n, first, u, v, zc));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(MLD_FULL_MASK, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < VC_THREADS / 32; w++) t += s_w[w];
        block_counts[blockIdx.x] = (unsigned int)t;
    }
}

// exclusive scan of the block counts in place (one block; nblocks is N / 1024, i.e. small), total -> counts[nblocks]
__global__ void __launch_bounds__(1024) visible_scan_kernel(unsigned int* __restrict__ counts, int nblocks) {
    __shared__ unsigned int s_part[1024];
    const int per = (nblocks + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(lo + per, nblocks);
    unsigned int t = 0;
    for (int i = lo; i < hi; i++) t += counts[i];
    s_part[threadIdx.x] = t;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan of the per-thread sums
        const unsigned int add = (threadIdx.x >= o) ? s_part[threadIdx.x - o] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += add;
        __syncthreads();
    }
    unsigned int run = s_part[threadIdx.x] - t;
    for (int i = lo; i < hi; i++) {
        const unsigned int c = counts[i];
        counts[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) counts[nblocks] = s_part[1023];
}

__global__ void __launch_bounds__(VC_THREADS)
visible_compact_kernel(DevParams P, const float* __restrict__ pts, int stride_f, long long n, const unsigned int* __restrict__ block_offsets,
                       long long capacity, int* __restrict__ point_index, double* __restrict__ image_points, double* __restrict__ depth_cam) {
    __shared__ int s_w[VC_THREADS / 32];
    double u[VC_PPT], v[VC_PPT], zc[VC_PPT];
    const long long first = ((long long)blockIdx.x * VC_THREADS + threadIdx.x) * VC_PPT;
    const int m = vc_flags(P, pts, stride_f, n, first, u, v, zc);
    const int c = __popc(m);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = c;  // inclusive scan of the per-thread counts inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(MLD_FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; w++) base += s_w[w];
    long long slot = (long long)block_offsets[blockIdx.x] + base + (incl - c);
#pragma unroll
    for (int j = 0; j < VC_PPT; j++) {
        if (!((m >> j) & 1)) continue;
        if (slot < capacity) {
            if (point_index) point_index[slot] = (int)(first + j);
            if (image_points) {
                image_points[2 * slot] = u[j];
                image_points[2 * slot + 1] = v[j];
            }
            if (depth_cam) depth_cam[slot] = zc[j];
        }
        slot++;
    }
}

}  // namespace

size_t mld_visible_scratch_bytes(long long n) { return (size_t)((n + VC_THREADS * VC_PPT - 1) / (VC_THREADS * VC_PPT) + 2) * sizeof(unsigned int); }

// d_scratch: mld_visible_scratch_bytes(n); the visible count ends up in d_scratch[nblocks] (returned through d_count_out)
cudaError_t mld_launch_visible_compact(const DevParams& P, const float* d_pts, int stride_f, long long n, void* d_scratch, long long capacity,
                                       int* d_point_index, double* d_image_points, double* d_depth_cam, const unsigned int** d_count_out,
                                       cudaStream_t stream, int* launches) {
    unsigned int* counts = reinterpret_cast<unsigned int*>(d_scratch);
    const long long nb = (n + VC_THREADS * VC_PPT - 1) / (VC_THREADS * VC_PPT);
    if (nb > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (d_count_out) *d_count_out = counts + nb;
    if (n <= 0) return cudaMemsetAsync(counts, 0, sizeof(unsigned int), stream);
    visible_count_kernel<<<(unsigned)nb, VC_THREADS, 0, stream>>>(P, d_pts, stride_f, n, counts);
    visible_scan_kernel<<<1, 1024, 0, stream>>>(counts, (int)nb);
    visible_compact_kernel<<<(unsigned)nb, VC_THREADS, 0, stream>>>(P, d_pts, stride_f, n, counts, capacity, d_point_index, d_image_points,
                                                                  d_depth_cam);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

// Host side of the pre-filter: the five linear forms in double, rounded to float, with bounds that
// cover (a) rounding the coefficients to float, (b) the four float operations of the evaluation and
// (c) the (far smaller) rounding of the exact FP64 path itself. 16 * 2^-24 leaves a factor > 2 of slack.
void mld_setup_prefilter(DevParams& P) {
    P.map_cells = P.W * P.H;
    P.occ_words = (int)occ_words_per_frame(P.W, P.H);
    P.occ_tx = occ_tiles_x(P.W);
    const double W = (double)P.W, H = (double)P.H;
    double g[5][3], h[5];
    for (int j = 0; j < 3; j++) {
        g[0][j] = P.R[6 + j];
        g[1][j] = P.f * P.R[0 + j] + P.cx * P.R[6 + j];
        g[2][j] = P.f * P.R[0 + j] + (P.cx - W) * P.R[6 + j];
        g[3][j] = P.f * P.R[3 + j] + P.cy * P.R[6 + j];
        g[4][j] = P.f * P.R[3 + j] + (P.cy - H) * P.R[6 + j];
    }
    h[0] = P.t[2];
    h[1] = P.f * P.t[0] + P.cx * P.t[2];
    h[2] = P.f * P.t[0] + (P.cx - W) * P.t[2];
    h[3] = P.f * P.t[1] + P.cy * P.t[2];
    h[4] = P.f * P.t[1] + (P.cy - H) * P.t[2];
    const double eps = 16.0 / 16777216.0;
    for (int k = 0; k < 5; k++) {
        double gmax = 0;
        for (int j = 0; j < 3; j++) {
            P.pf_g[k][j] = (float)g[k][j];
            // the exact path evaluates f*X + c*Z from separately rounded terms: bound with the parts' magnitudes
            gmax = fmax(gmax, fabs(g[k][j]));
        }
        // |f*R0j| + |c*R2j| can exceed |f*R0j + c*R2j| when the terms cancel; use the un-cancelled magnitude
        double gabs = gmax;
        if (k > 0) {
            const double c = (k == 1) ? P.cx : (k == 2) ? (P.cx - W) : (k == 3) ? P.cy : (P.cy - H);
            const int row = (k <= 2) ? 0 : 3;
            double m = 0;
            for (int j = 0; j < 3; j++) m = fmax(m, fabs(P.f * P.R[row + j]) + fabs(c * P.R[6 + j]));
            gabs = m;
            P.pf_H[k] = (float)(eps * (fabs(P.f * P.t[row / 3]) + fabs(c * P.t[2])) * 1.0000002 + 1e-30);
        } else {
            P.pf_H[k] = (float)(eps * fabs(h[0]) * 1.0000002 + 1e-30);
        }
        P.pf_h[k] = (float)h[k];
        P.pf_G[k] = (float)(eps * gabs * 1.0000002 + 1e-30);
    }
}

cudaError_t mld_launch_project_scatter(const DevParams& P, const MapCode& mc, const float* d_pts, int stride_f, long long n,
                                       long long pitch_pts, unsigned int* d_maps, unsigned int* d_occ, int nframes,
                                       cudaStream_t stream) {
    if (n <= 0 || nframes <= 0) return cudaSuccess;
    if (n > 0x7fffffffLL / 8) return cudaErrorInvalidValue;  // 32-bit point indexing inside a frame
    dim3 grid((unsigned)((n + K1_THREADS * K1_PPT - 1) / (K1_THREADS * K1_PPT)), (unsigned)nframes);
    project_scatter_kernel<<<grid, K1_THREADS, 0, stream>>>(P, mc, d_pts, stride_f, (int)n, pitch_pts, d_maps, d_occ);
    return cudaGetLastError();
}

cudaError_t mld_launch_points_camera_indexed(const DevParams& P, const float* d_pts, int stride_f, long long n, const int* d_idx, long long n_idx,
                                             double* d_out, cudaStream_t stream) {
    if (n_idx <= 0) return cudaSuccess;
    points_camera_indexed_kernel<<<(unsigned)((n_idx + 255) / 256), 256, 0, stream>>>(P, d_pts, stride_f, n, d_idx, n_idx, d_out);
    return cudaGetLastError();
}

cudaError_t mld_launch_visible_debug(const DevParams& P, const float* d_pts, int stride_f, long long n,
                                     unsigned char* d_visible, double* d_cam, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    visible_debug_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(P, d_pts, stride_f, n, d_visible, d_cam);
    return cudaGetLastError();
}
