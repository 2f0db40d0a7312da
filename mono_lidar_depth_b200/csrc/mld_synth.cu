// mld_synth.cu -- deterministic synthetic KITTI-shaped input (bench / tests only; not a reference
// component). One spinning lidar (rings x azimuth steps, point order azimuth-major then ring) over a
// ground plane with random axis-aligned boxes, range noise and NaN dropouts; features are integer
// pixel coordinates like the ones tracklets_depth hands to the estimator
// (/root/reference/tracklets_depth/src/tracklet_depth_module.cpp:75-76).
//
// Host and device run the SAME inline functions: only exactly rounded float operations (+ - * /,
// comparisons via ternaries) on hashes and on trig tables computed once on the host, no FMA
// contraction on either side (-fmad=false / -ffp-contract=off), so both produce identical bits.
#include <math.h>
#include <string.h>

#include <vector>

#include "mld_common.cuh"
#include "mld_kernels.h"

namespace {

struct SynthBox {
    float lox, hix, loy, hiy, loz, hiz;
};

__host__ __device__ __forceinline__ float u01(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }
__host__ __device__ __forceinline__ float u01b(uint64_t h) { return (float)((h >> 16) & 0xffffffull) * (1.0f / 16777216.0f); }
__host__ __device__ __forceinline__ float qnan_f() {
#ifdef __CUDA_ARCH__
    return __int_as_float(0x7fc00000);
#else
    uint32_t b = 0x7fc00000u;
    float f;
    memcpy(&f, &b, sizeof(f));
    return f;
#endif
}

__host__ __device__ inline SynthBox synth_box(const mld_synth_config& c, uint64_t seed, long long frame, int b,
                                              const float* tables) {
    const float* cos_az = tables + 2 * c.rings;
    const float* sin_az = cos_az + c.azimuth_steps;
    uint64_t key = seed ^ 0xB0C5B0C5ull;
    uint64_t h0 = mld_hash3(key, (uint64_t)frame, (uint64_t)b, 0);
    uint64_t h1 = mld_hash3(key, (uint64_t)frame, (uint64_t)b, 1);
    uint64_t h2 = mld_hash3(key, (uint64_t)frame, (uint64_t)b, 2);
    uint64_t h3 = mld_hash3(key, (uint64_t)frame, (uint64_t)b, 3);
    uint64_t h4 = mld_hash3(key, (uint64_t)frame, (uint64_t)b, 4);
    uint64_t h5 = mld_hash3(key, (uint64_t)frame, (uint64_t)b, 5);
    // two thirds of the boxes sit in the +-60 degree sector in front of the sensor (the camera looks along +x)
    int az;
    int sector = c.azimuth_steps / 6;
    if ((h0 % 3ull) != 0ull) {
        int off = (int)(h1 % (uint64_t)(2 * sector + 1)) - sector;
        az = (off + c.azimuth_steps) % c.azimuth_steps;
    } else {
        az = (int)(h1 % (uint64_t)c.azimuth_steps);
    }
    float r = 8.0f + 72.0f * u01(h2) * u01b(h2);  // denser near the sensor, never on top of it
    float cx = r * cos_az[az], cy = r * sin_az[az];
    float hx = 0.5f + 2.5f * u01(h3), hy = 0.5f + 2.5f * u01(h4), hh = 0.5f + 3.5f * u01(h5);
    SynthBox bx;
    bx.lox = cx - hx; bx.hix = cx + hx;
    bx.loy = cy - hy; bx.hiy = cy + hy;
    bx.loz = -c.sensor_height; bx.hiz = -c.sensor_height + hh;
    return bx;
}

__host__ __device__ __forceinline__ float fmin_t(float a, float b) { return (a < b) ? a : b; }
__host__ __device__ __forceinline__ float fmax_t(float a, float b) { return (a > b) ? a : b; }

// range to the nearest surface along direction (dx,dy,dz) or a negative value for "no return"
__host__ __device__ inline float synth_cast(const mld_synth_config& c, float dx, float dy, float dz, const SynthBox* boxes,
                                            int nb) {
    float best = c.max_range;
    bool hit = false;
    if (dz < 0.0f) {
        float t = (-c.sensor_height) / dz;
        if (t < best) {
            best = t;
            hit = true;
        }
    }
    for (int b = 0; b < nb; b++) {
        const SynthBox& bx = boxes[b];
        float tx1 = bx.lox / dx, tx2 = bx.hix / dx;
        float ty1 = bx.loy / dy, ty2 = bx.hiy / dy;
        float tz1 = bx.loz / dz, tz2 = bx.hiz / dz;
        float tn = fmax_t(fmax_t(fmin_t(tx1, tx2), fmin_t(ty1, ty2)), fmin_t(tz1, tz2));
        float tf = fmin_t(fmin_t(fmax_t(tx1, tx2), fmax_t(ty1, ty2)), fmax_t(tz1, tz2));
        if (tn <= tf && tn > 0.5f && tn < best) {
            best = tn;
            hit = true;
        }
    }
    return hit ? best : -1.0f;
}

__host__ __device__ inline void synth_point(const mld_synth_config& c, uint64_t seed, long long frame, long long idx,
                                            const float* tables, const SynthBox* boxes, float out[4]) {
    const float* cos_el = tables;
    const float* sin_el = tables + c.rings;
    const float* cos_az = tables + 2 * c.rings;
    const float* sin_az = cos_az + c.azimuth_steps;
    int a = (int)(idx / c.rings), e = (int)(idx % c.rings);
    float dx = cos_el[e] * cos_az[a], dy = cos_el[e] * sin_az[a], dz = sin_el[e];
    uint64_t key = seed ^ 0x9017C10Dull;
    uint64_t hd = mld_hash3(key, (uint64_t)frame, (uint64_t)idx, 0);
    uint64_t hn = mld_hash3(key, (uint64_t)frame, (uint64_t)idx, 1);
    const float qnan = qnan_f();
    float t = synth_cast(c, dx, dy, dz, boxes, c.n_boxes);
    bool drop = u01(hd) < c.dropout_prob;
    if (t < 0.0f || drop) {
        out[0] = qnan; out[1] = qnan; out[2] = qnan; out[3] = 0.0f;
        return;
    }
    // Irwin-Hall(4) noise, unit variance after scaling by sqrt(3)
    float s = (float)(hn & 0xffff) * (1.0f / 65536.0f) + (float)((hn >> 16) & 0xffff) * (1.0f / 65536.0f) +
              (float)((hn >> 32) & 0xffff) * (1.0f / 65536.0f) + (float)((hn >> 48) & 0xffff) * (1.0f / 65536.0f);
    float noise = (s - 2.0f) * 1.7320508f * c.range_noise_sigma;
    float tr = t + noise;
    out[0] = dx * tr; out[1] = dy * tr; out[2] = dz * tr;
    out[3] = u01b(hd);
}

__host__ __device__ inline void synth_feature(const mld_synth_config& c, uint64_t seed, long long frame, int i, double out[2]) {
    uint64_t key = seed ^ 0xFEA7FEA7ull;
    uint64_t h0 = mld_hash3(key, (uint64_t)frame, (uint64_t)i, 0);
    uint64_t h1 = mld_hash3(key, (uint64_t)frame, (uint64_t)i, 1);
    uint64_t h2 = mld_hash3(key, (uint64_t)frame, (uint64_t)i, 2);
    int W = c.image_width, H = c.image_height;
    int band_top = (int)(c.band_top_frac * (float)H);
    if (band_top < 1) band_top = 1;
    if (band_top > H - 1) band_top = H - 1;
    int u = (int)(h0 % (uint64_t)W);
    int v;
    if (u01(h1) < c.band_feature_frac)
        v = band_top + (int)(h2 % (uint64_t)(H - band_top));
    else
        v = (int)(h2 % (uint64_t)band_top);
    out[0] = (double)u;
    out[1] = (double)v;
}

constexpr int SYNTH_MAX_BOXES = 64;

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

__global__ void synth_features_kernel(mld_synth_config c, uint64_t seed, long long frame0, int F, double* __restrict__ out) {
    const long long frame = frame0 + blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F) return;
    double uv[2];
    synth_feature(c, seed, frame, i, uv);
    reinterpret_cast<double2*>(out)[(long long)blockIdx.y * F + i] = make_double2(uv[0], uv[1]);
}

}  // namespace

void mld_synth_build_tables(const mld_synth_config& c, float* tables) {
    const double deg = 3.14159265358979323846 / 180.0;
    for (int e = 0; e < c.rings; e++) {
        double el = (double)c.elev_top_deg +
                    ((double)c.elev_bottom_deg - (double)c.elev_top_deg) * (c.rings > 1 ? (double)e / (double)(c.rings - 1) : 0.0);
        tables[e] = (float)cos(el * deg);
        tables[c.rings + e] = (float)sin(el * deg);
    }
    for (int a = 0; a < c.azimuth_steps; a++) {
        // azimuth 0 looks along +x; the sweep starts behind the sensor so that frontal points are mid-cloud
        double az = -180.0 + 360.0 * (double)a / (double)c.azimuth_steps;
        tables[2 * c.rings + a] = (float)cos(az * deg);
        tables[2 * c.rings + c.azimuth_steps + a] = (float)sin(az * deg);
    }
}

void mld_synth_points_host_impl(const mld_synth_config& c, uint64_t seed, long long frame, const float* tables, float* out) {
    std::vector<SynthBox> boxes((size_t)c.n_boxes);
    for (int b = 0; b < c.n_boxes; b++) boxes[(size_t)b] = synth_box(c, seed, frame, b, tables);
    const long long n = (long long)c.rings * c.azimuth_steps;
    for (long long i = 0; i < n; i++) synth_point(c, seed, frame, i, tables, boxes.data(), out + i * 4);
}

void mld_synth_features_host_impl(const mld_synth_config& c, uint64_t seed, long long frame, int F, double* out) {
    for (int i = 0; i < F; i++) synth_feature(c, seed, frame, i, out + (size_t)i * 2);
}

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
                                      double* d_out, cudaStream_t stream) {
    if (F <= 0 || nframes <= 0) return cudaSuccess;
    const long long max_y = 32768;
    for (long long f = 0; f < nframes; f += max_y) {
        long long cnt = nframes - f < max_y ? nframes - f : max_y;
        dim3 grid((unsigned)((F + 255) / 256), (unsigned)cnt);
        synth_features_kernel<<<grid, 256, 0, stream>>>(c, seed, frame0 + f, F, d_out + f * (long long)F * 2);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
