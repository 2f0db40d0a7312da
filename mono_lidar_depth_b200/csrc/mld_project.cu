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
// it is done with atomicMin(raw index) on a map pre-filled with 0xFFFFFFFF -- deterministic and
// bit-exact whatever the thread schedule.
//
// HBM traffic per frame: 16 B (float4) per point read once, streaming; the 4*W*H-byte map is cleared
// by a memset node and touched by ~N_visible L2 atomics.
#include "mld_common.cuh"
#include "mld_kernels.h"

namespace {

constexpr int K1_THREADS = 256;
constexpr int K1_PPT = 4;  // points per thread: four independent 16-byte loads in flight

// projection of one point; returns the pixel offset or -1 when the point does not enter the map
__device__ __forceinline__ int project_pixel(const DevParams& P, float x, float y, float z, bool need_front) {
    D3 c = lidar_to_cam(P, x, y, z);
    // the map only accepts points in front of the camera (NeighborFinderPixel.cpp:51); testing it
    // first skips both divisions for everything behind the image plane
    if (need_front && !(c.z > 0.0)) return -1;
    // K * p with K = [f 0 cx; 0 f cy; 0 0 1] evaluated term by term like Eigen's product
    // (camera_pinhole.h:88), then colwise().hnormalized() = division by the third row (:90)
    double q0 = __dadd_rn(__dadd_rn(__dmul_rn(P.f, c.x), __dmul_rn(0.0, c.y)), __dmul_rn(P.cx, c.z));
    double q1 = __dadd_rn(__dadd_rn(__dmul_rn(0.0, c.x), __dmul_rn(P.f, c.y)), __dmul_rn(P.cy, c.z));
    double q2 = __dadd_rn(__dadd_rn(__dmul_rn(0.0, c.x), __dmul_rn(0.0, c.y)), __dmul_rn(1.0, c.z));
    double u = __ddiv_rn(q0, q2);
    double v = __ddiv_rn(q1, q2);
    double Wd = (double)P.W, Hd = (double)P.H;
    bool in_range = (u >= 0.) && (u <= Wd) && (v >= 0.) && (v <= Hd);  // camera_pinhole.h:93-96
    bool visible = (u > 0.) && (u < Wd) && (v > 0.) && (v < Hd);       // DepthEstimator.cpp:186-187
    if (!(in_range && visible)) return -1;
    return (int)v * P.W + (int)u;  // int x_img = u; int y_img = v (NeighborFinderPixel.cpp:41-42)
}

__global__ void __launch_bounds__(K1_THREADS)
project_scatter_kernel(DevParams P, const float* __restrict__ pts, int stride_f, long long n, long long pitch_pts,
                       unsigned int* __restrict__ maps) {
    const long long frame = blockIdx.y;
    const float* fp = pts + frame * pitch_pts * (long long)stride_f;
    unsigned int* map = maps + frame * (long long)P.W * (long long)P.H;
    const long long base = (long long)blockIdx.x * (K1_THREADS * K1_PPT) + threadIdx.x;

    float4 p[K1_PPT];
#pragma unroll
    for (int j = 0; j < K1_PPT; j++) {
        long long i = base + (long long)j * K1_THREADS;
        if (i < n)
            p[j] = ld_stream_f4(fp + i * stride_f);
        else
            p[j] = make_float4(0.f, 0.f, -1.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < K1_PPT; j++) {
        long long i = base + (long long)j * K1_THREADS;
        if (i >= n) continue;
        int off = project_pixel(P, p[j].x, p[j].y, p[j].z, true);
        if (off >= 0) atomicMin(&map[off], (unsigned int)i);
    }
}

// debug view: Transform_Cloud_LidarToCamera's visibility cull (no z > 0 test, DepthEstimator.cpp:184-207)
// and the camera-frame coordinates (_points_cs_camera). Not on the hot path.
__global__ void visible_debug_kernel(DevParams P, const float* __restrict__ pts, int stride_f, long long n,
                                     unsigned char* __restrict__ visible, double* __restrict__ cam) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = pts[i * stride_f], y = pts[i * stride_f + 1], z = pts[i * stride_f + 2];
    if (visible) visible[i] = project_pixel(P, x, y, z, false) >= 0 ? 1 : 0;
    if (cam) {
        D3 c = lidar_to_cam(P, x, y, z);
        cam[i * 3] = c.x;
        cam[i * 3 + 1] = c.y;
        cam[i * 3 + 2] = c.z;
    }
}

}  // namespace

cudaError_t mld_launch_project_scatter(const DevParams& P, const float* d_pts, int stride_f, long long n,
                                       long long pitch_pts, unsigned int* d_maps, int nframes, cudaStream_t stream) {
    if (n <= 0 || nframes <= 0) return cudaSuccess;
    dim3 grid((unsigned)((n + K1_THREADS * K1_PPT - 1) / (K1_THREADS * K1_PPT)), (unsigned)nframes);
    project_scatter_kernel<<<grid, K1_THREADS, 0, stream>>>(P, d_pts, stride_f, n, pitch_pts, d_maps);
    return cudaGetLastError();
}

cudaError_t mld_launch_visible_debug(const DevParams& P, const float* d_pts, int stride_f, long long n,
                                     unsigned char* d_visible, double* d_cam, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    visible_debug_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(P, d_pts, stride_f, n, d_visible, d_cam);
    return cudaGetLastError();
}
