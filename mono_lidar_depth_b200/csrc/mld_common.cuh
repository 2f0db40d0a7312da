// mld_common.cuh -- shared device/host definitions for libmld_cuda.so (sm_100a).
//
// All hot-path arithmetic is IEEE double (the reference works on Eigen::Matrix3Xd / Vector3d,
// monolidar_fusion/src/DepthEstimator.cpp:169) and every translation unit is compiled with
// -fmad=false so that no multiply-add is contracted: the (int) casts, bin indices and threshold
// compares then see exactly the values a plain SSE2 build of the reference computes.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mld_c_api.h"
#include "mld_hash.h"

#define MLD_FULL_MASK 0xffffffffu
#define MLD_EMPTY 0xffffffffu  // pixel-map cell without a point; reads as int32 -1 (POINT_NOT_DEFINED)

// Mono_Lidar::DepthResultType (eDepthResultType.h:9-31) as ST_<name>, from the C ABI's list
#define MLD_ST_ENTRY(name, value) ST_##name = value,
enum MldStatus : int { MLD_DEPTH_RESULT_TYPES(MLD_ST_ENTRY) };
#undef MLD_ST_ENTRY

enum MldRoadMode : int { ROAD_NONE = 0, ROAD_TRIANGLE = 1, ROAD_LEASTSQUARES = 2, ROAD_MESTIMATOR = 3 };

// Constant block handed to every kernel by value (DepthEstimator::Initialize's module selection,
// monolidar_fusion/src/DepthEstimator.cpp:35-127, flattened into flags).
struct DevParams {
    int W, H;
    double Wd, Hd;        // (double)W, (double)H
    double f, cx, cy;
    double R[9], t[3];    // transform_lidar_to_cam
    double Ri[9], ti[3];  // transform_cam_to_lidar (Affine3d::inverse(), :44)
    double Kinv[9];       // makeIntrinsics().inverse() (camera_pinhole.h:65)
    double hx1, hy1;      // half window, scale (1,1)      (NeighborFinderPixel.cpp:67-68)
    double hx2, hy2;      // half window, scale (2.0,1.5)  (DepthEstimator.cpp:585)
    int count_min;        // radiusSearch_count_min
    int use_hist;
    int hist_min;
    double bin_w;
    int use_tri_max;      // do_use_triangle_size_maximation
    int check_planar;     // do_check_triangleplanar_condition
    double crossnorm_thr;
    double ortho_thr;     // > 0 selects LinePlaneIntersectionOrthogonalTreshold (:77-81)
    int use_pca;
    double pca_3_abs_min, pca_3_2_rel_max, pca_2_1_rel_min;
    int glob_en, glob_mode;
    double glob_min, glob_max;
    int loc_en, loc_mode, loc_type;
    double loc_val;
    int cut_behind;
    int road_mode;        // MldRoadMode; ROAD_NONE when do_use_ransac_plane == 0
    double road_dist_thr; // ransac_plane_point_distance_treshold
    double zx_min_rel;    // plane_estimator_z_x_min_relation
    int set_all_zero;
    // FP32 pre-filter of K1 (see mld_project.cu): five linear forms g.(x,y,z)+h in the lidar point --
    // z_cam, f*X+cx*Z, f*X+(cx-W)*Z, f*Y+cy*Z, f*Y+(cy-H)*Z -- with a rigorous float error bound
    // G*(|x|+|y|+|z|)+H each. A point is dropped without any FP64 work only when one of the exact
    // visibility tests is violated by more than that bound.
    float pf_g[5][3], pf_h[5], pf_G[5], pf_H[5];
    // per-frame pitches of the map slots, precomputed (K1 used to derive them per thread: 4 % of its instructions, ncu r2o)
    int map_cells;  // W * H
    int occ_words;  // occ_words_per_frame(W, H)
    int occ_tx;     // occ_tiles_x(W)
};

// pixel-map cell encoding. Tagged mode (clouds of <= 2^18 points): key = (tag << 18) | raw index
// with tag = 0x3FFF - epoch; a newer epoch has a smaller tag, so atomicMin lets it overwrite every
// stale cell and the 4*W*H-byte clear per frame disappears (one clear per 16383 uses of a map slot).
// Plain mode (bigger clouds): key = raw index, map cleared to 0xFFFFFFFF before every use.
#define MLD_TAG_SHIFT 18
#define MLD_TAG_IDX_MASK 0x3FFFFu
#define MLD_TAG_MAX_EPOCH 16383u
struct MapCode {
    unsigned int tagged;  // 0 plain, 1 tagged
    unsigned int tag;     // current tag (tagged mode)
};
__host__ __device__ __forceinline__ bool map_cell_valid(const MapCode& mc, unsigned int cell) {
    return mc.tagged ? ((cell >> MLD_TAG_SHIFT) == mc.tag) : (cell != MLD_EMPTY);
}
__host__ __device__ __forceinline__ unsigned int map_cell_index(const MapCode& mc, unsigned int cell) {
    return mc.tagged ? (cell & MLD_TAG_IDX_MASK) : cell;
}

// Occupancy bitmap of the pixel map (one per in-flight frame, cleared per chunk), tiled: a 16 x 16 pixel tile is one
// 32-byte sector (8 words; word w holds rows 2w and 2w+1 of the tile, 16 bits each). A search window (7 x 10 pixels
// by default) touches 1-4 tiles = 1-4 sectors instead of one sector per window row, and K1 sets ONE bit per pixel it
// writes. K2 reads window rows from here instead of loading every pixel cell of the 4-byte map.
__host__ __device__ __forceinline__ int occ_tiles_x(int W) { return (W + 15) >> 4; }
__host__ __device__ __forceinline__ long long occ_words_per_frame(int W, int H) { return (long long)occ_tiles_x(W) * ((H + 15) >> 4) * 8; }
// word holding pixel (x, y) and the bit inside it
__host__ __device__ __forceinline__ long long occ_word_of(int tiles_x, int x, int y) {
    return ((long long)(y >> 4) * tiles_x + (x >> 4)) * 8 + ((y & 15) >> 1);
}
__host__ __device__ __forceinline__ unsigned int occ_bit_of(int x, int y) { return 1u << (((y & 1) << 4) | (x & 15)); }

struct D3 {
    double x, y, z;
};

__host__ __device__ __forceinline__ D3 d3(double x, double y, double z) { return D3{x, y, z}; }
__host__ __device__ __forceinline__ D3 operator-(const D3& a, const D3& b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ __forceinline__ D3 operator+(const D3& a, const D3& b) { return D3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ __forceinline__ D3 operator*(const D3& a, double s) { return D3{a.x * s, a.y * s, a.z * s}; }
__host__ __device__ __forceinline__ D3 operator/(const D3& a, double s) { return D3{a.x / s, a.y / s, a.z / s}; }
// 3-term reductions left to right, like Eigen's unrolled redux on Vector3d
__host__ __device__ __forceinline__ double dot3(const D3& a, const D3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__host__ __device__ __forceinline__ double sqnorm3(const D3& a) { return dot3(a, a); }
__host__ __device__ __forceinline__ double norm3(const D3& a) { return sqrt(sqnorm3(a)); }
__host__ __device__ __forceinline__ D3 cross3(const D3& a, const D3& b) {
    return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen 3.3 normalized(): untouched when the squared norm is not > 0
__host__ __device__ __forceinline__ D3 normalized3(const D3& a) {
    double z = sqnorm3(a);
    if (z > 0) return a / sqrt(z);
    return a;
}

// lidar -> camera, "((r0*x + r1*y) + r2*z) + t" with explicitly rounded operations
// (Eigen: res = t; res += R * p, monolidar_fusion/src/DepthEstimator.cpp:173)
__device__ __forceinline__ D3 lidar_to_cam(const DevParams& P, float x, float y, float z) {
    double dx = (double)x, dy = (double)y, dz = (double)z;
    D3 r;
    r.x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.R[0], dx), __dmul_rn(P.R[1], dy)), __dmul_rn(P.R[2], dz)), P.t[0]);
    r.y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.R[3], dx), __dmul_rn(P.R[4], dy)), __dmul_rn(P.R[5], dz)), P.t[1]);
    r.z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.R[6], dx), __dmul_rn(P.R[7], dy)), __dmul_rn(P.R[8], dz)), P.t[2]);
    return r;
}

// streaming 16-byte load that does not pollute L1 (points are read once per kernel)
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 v;
#ifdef MLD_DIAG_LDG
    return __ldg(reinterpret_cast<const float4*>(p));
#endif
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(MLD_FULL_MASK, v, m); }

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor_d(v, m);
    return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        double o = shfl_xor_d(v, m);
        v = (o > v) ? o : v;
    }
    return v;
}
__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        double o = shfl_xor_d(v, m);
        v = (o < v) ? o : v;
    }
    return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = min(v, __shfl_xor_sync(MLD_FULL_MASK, v, m));
    return v;
}

// Cyclic Jacobi eigen-decomposition of a symmetric 3x3 in registers, operation for operation the algorithm the CPU oracle pins
// against the reference for Eigen::SelfAdjointEigenSolver's contract (the test-side CPU restatement: full two-sided updates
// A <- A J, A <- J^T A in the pair order (0,1), (0,2), (1,2), the same stopping test, a stable ascending sort): the float-cast
// eigenvalue ratios of the PCA variant (PCA.cpp:27-37) and the RANSAC refit then agree with the oracle bit for bit.
// Input a = (a00,a01,a02,a11,a12,a22); output: w ascending, v[c] = eigenvector of w[c].
template <int p, int q>
__device__ __forceinline__ void mld_jacobi_rotate(double (&a)[9], double (&v)[9]) {
    const double apq = a[p * 3 + q];
    if (apq == 0.0) return;
    const double app = a[p * 3 + p], aqq = a[q * 3 + q];
    const double theta = (aqq - app) / (2.0 * apq);
    double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    if (!isfinite(theta)) t = 0.0;
    const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
    for (int k = 0; k < 3; k++) {  // A <- A * J
        const double akp = a[k * 3 + p], akq = a[k * 3 + q];
        a[k * 3 + p] = c * akp - s * akq;
        a[k * 3 + q] = s * akp + c * akq;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {  // A <- J^T * A
        const double apk = a[p * 3 + k], aqk = a[q * 3 + k];
        a[p * 3 + k] = c * apk - s * aqk;
        a[q * 3 + k] = s * apk + c * aqk;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double vkp = v[k * 3 + p], vkq = v[k * 3 + q];
        v[k * 3 + p] = c * vkp - s * vkq;
        v[k * 3 + q] = s * vkp + c * vkq;
    }
}

static __device__ __noinline__ void eig3_sym_regs(double a00, double a01, double a02, double a11, double a12, double a22, double w[3],
                                                  D3 v[3]) {
    double a[9] = {a00, a01, a02, a01, a11, a12, a02, a12, a22};
    double r[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int sweep = 0; sweep < 60; sweep++) {
        const double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
        const double diag = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
        if (!(off > 1e-32 * diag) || !(off > 0)) break;
        mld_jacobi_rotate<0, 1>(a, r);
        mld_jacobi_rotate<0, 2>(a, r);
        mld_jacobi_rotate<1, 2>(a, r);
    }
    // stable ascending order of the diagonal (what std::sort does on three elements)
    const double d0 = a[0], d1 = a[4], d2 = a[8];
    int i0 = 0, i1 = 1, i2 = 2;
    if (d1 < d0) { i0 = 1; i1 = 0; }
    {
        const double di1 = (i1 == 0) ? d0 : d1, di0 = (i0 == 0) ? d0 : d1;
        if (d2 < di1) {
            i2 = i1;
            i1 = 2;
            if (d2 < di0) {
                i1 = i0;
                i0 = 2;
            }
        }
    }
    auto diag_of = [&](int i) { return i == 0 ? d0 : (i == 1 ? d1 : d2); };
    auto col_of = [&](int i) { return i == 0 ? D3{r[0], r[3], r[6]} : (i == 1 ? D3{r[1], r[4], r[7]} : D3{r[2], r[5], r[8]}); };
    w[0] = diag_of(i0); w[1] = diag_of(i1); w[2] = diag_of(i2);
    v[0] = col_of(i0); v[1] = col_of(i1); v[2] = col_of(i2);
}

// cofactor inverse of a row-major 3x3, the way Eigen evaluates Matrix3d::inverse()
// (result(i,j) = cofactor(j,i) * (1/det)); host only, used once per Initialize.
inline void mld_inverse3_host(const double* m, double* out) {
    auto cof = [&](int i, int j) {
        int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
    };
    double c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    double det = (c0 * m[0] + c1 * m[3]) + c2 * m[6];
    double invdet = 1.0 / det;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) out[i * 3 + j] = cof(j, i) * invdet;
}
