// mld_synth.cu -- device generators of the synthetic KITTI-shaped input (include/mld_synth.h; bench / tests only,
// not a reference component). The scene model is mld_synth_model.h, shared bit for bit with the host generators of
// libmld_synth.so.
#include <vector>

#include "mld_common.cuh"
#include "mld_kernels.h"
#include "mld_synth_model.h"

namespace {

__global__ void synth_points_kernel(mld_synth_config c, uint64_t seed, long long frame0, long long pitch_pts,
                                    const float* __restrict__ tables, float* __restrict__ out) {
    __shared__ SynthBox boxes[SYNTH_MAX_BOXES];
    const long long frame = frame0 + blockIdx.y;
    const long long n = (long long)c.rings * c.azimuth_steps;
    if (threadIdx.x < c.n_boxes) boxes[threadIdx.x] = synth_box(c, seed, frame, threadIdx.x, tables);
    __syncthreads();
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    float p[4];
    synth_point(c, seed, frame, idx, tables, boxes, p);
    reinterpret_cast<float4*>(out)[(long long)blockIdx.y * pitch_pts + idx] = make_float4(p[0], p[1], p[2], p[3]);
}

__global__ void synth_features_kernel(mld_synth_config c, uint64_t seed, long long frame0, int F, const float* __restrict__ tables,
                                      double* __restrict__ out) {
    const long long frame = frame0 + blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F) return;
    double uv[2];
    synth_feature(c, seed, frame, i, tables, uv);
    reinterpret_cast<double2*>(out)[(long long)blockIdx.y * F + i] = make_double2(uv[0], uv[1]);
}

}  // namespace

cudaError_t mld_launch_synth_points(const mld_synth_config& c, uint64_t seed, long long frame0, long long nframes,
                                    long long pitch_pts, const float* d_tables, float* d_out, cudaStream_t stream) {
    const long long n = (long long)c.rings * c.azimuth_steps;
    if (n <= 0 || nframes <= 0) return cudaSuccess;
    const long long max_y = 32768;
    for (long long f = 0; f < nframes; f += max_y) {
        long long cnt = nframes - f < max_y ? nframes - f : max_y;
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)cnt);
        synth_points_kernel<<<grid, 256, 0, stream>>>(c, seed, frame0 + f, pitch_pts, d_tables, d_out + f * pitch_pts * 4);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t mld_launch_synth_features(const mld_synth_config& c, uint64_t seed, long long frame0, long long nframes, int F,
                                      const float* d_tables, double* d_out, cudaStream_t stream) {
    if (F <= 0 || nframes <= 0) return cudaSuccess;
    const long long max_y = 32768;
    for (long long f = 0; f < nframes; f += max_y) {
        long long cnt = nframes - f < max_y ? nframes - f : max_y;
        dim3 grid((unsigned)((F + 255) / 256), (unsigned)cnt);
        synth_features_kernel<<<grid, 256, 0, stream>>>(c, seed, frame0 + f, F, d_tables, d_out + f * (long long)F * 2);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

void mld_synth_build_tables(const mld_synth_config& c, float* tables) { synth_build_tables(c, tables); }
size_t mld_synth_table_floats(const mld_synth_config& c) { return synth_table_floats(c); }
bool mld_synth_config_ok(const mld_synth_config* c) { return synth_config_ok(c); }
