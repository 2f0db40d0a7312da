// mld_semantic.cu -- SemanticPlane::CalculateInliersPlane on the GPU: the ground plane the
// production caller actually uses (tracklets_depth/src/tracklet_depth_module.cpp:269-284).
//
// Replaces (reference, /root/reference/monolidar_fusion/src/RansacPlane.cpp:159-274):
//   pcl::transformPointCloud(cloud, transformed, cam_.transform_cam_lidar)        :198
//   project(): p = K * (x,y,z); p /= p[2]; cv::Point(p[0], p[1])                  :170-181
//   keep points whose pixel carries a ground label                                 :201-222
//   < 3 kept -> ExceptionPclInvalid                                                :224-227
//   SampleConsensusModelPlane::optimizeModelCoefficients(kept, (0,0,1,0))          :236-242
//   selectWithinDistance(coeffs, inlier_threshold) over the WHOLE cloud            :248
//   optimizeModelCoefficients(inliers, coeffs)                                     :249
//   _modelCoeffs = refined, _inliersIndex = inliers                                :254-265
//
// Two streaming passes over the cloud (16 B per point each, HBM bound) with the 3x3 fit done by the last block of
// each pass (ticket counter), so a frame costs two launches and no host round trip. Thread i owns point i, hence
// a warp's ballot IS word i/32 of the inlier bitmask: plain coalesced stores, no atomics, no clear. In pass 1 a float
// quotient on the rounded camera-frame coordinates removes the ~85 % of a sweep that cannot hit the image before the two
// FP64 divisions; in a batch every block streams 8 tiles so that the ten-moment reduction is paid once per 8192 points
// (first version, one tile per block and no pre-filter: 224 + 129 us per 128 sweeps; round-1 captures r1b, git history).
//
// Arithmetic: the transform is evaluated in double and rounded to float per coordinate (PCL 1.8 computes
// transform(i,0)*x + ... in the transform's scalar and casts), the projection in double with true division and
// (int) truncation like cv::Point_<int>(double, double); the point-to-plane distance in float (PCL). The moments of
// the least-squares fit are accumulated in double (PCL: sequential float), so coefficients agree with the reference
// to ~1e-4 and the inlier sets differ only for points within that margin of the threshold (tests state both).
// Deviation: a pixel with x == cols or y == rows passes the reference's validity test ('>' instead of '>=',
// :205-206) and is then read out of bounds (undefined); here such a pixel is "not ground".
#include "mld_common.cuh"
#include "mld_kernels.h"

namespace {

#ifndef MLD_SP_THREADS
#define MLD_SP_THREADS 256
#endif
#ifndef MLD_SP_PPT
#define MLD_SP_PPT 4
#endif
constexpr int SP_THREADS = MLD_SP_THREADS;
constexpr int SP_PPT = MLD_SP_PPT;  // independent 16-byte loads a thread keeps in flight
constexpr int SP_NSUM = 10;  // n, x, y, z, xx, xy, xz, yy, yz, zz

__device__ __forceinline__ double sp_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(MLD_FULL_MASK, v, o);
    return v;
}

// block-reduce the 10 moments and add them to acc (double atomics); returns true in the LAST block of the frame
__device__ bool sp_commit(double (&v)[SP_NSUM], double* __restrict__ acc, unsigned int* __restrict__ ticket, unsigned int nblocks) {
    __shared__ double s_red[SP_NSUM][SP_THREADS / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < SP_NSUM; q++) {
        const double t = sp_warp_sum(v[q]);
        if (lane == 0) s_red[q][warp] = t;
    }
    __syncthreads();
    if (threadIdx.x < SP_NSUM) {
        double t = 0;
#pragma unroll
        for (int w = 0; w < SP_THREADS / 32; w++) t += s_red[threadIdx.x][w];
        if (t != 0.0) atomicAdd(&acc[threadIdx.x], t);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == nblocks - 1);
    __syncthreads();
    return s_last;
}

// SampleConsensusModelPlane::optimizeModelCoefficients: centroid + eigenvector of the smallest eigenvalue of the
// covariance; returns the input model unchanged when there are not more than 3 inliers.
__device__ void sp_fit(const double* __restrict__ acc, const float in[4], float out[4]) {
    out[0] = in[0]; out[1] = in[1]; out[2] = in[2]; out[3] = in[3];
    const double m = acc[0];
    if (!(m > 3.0)) return;
    const double mx = acc[1] / m, my = acc[2] / m, mz = acc[3] / m;
    double w[3];
    D3 ev[3];
    eig3_sym_regs(acc[4] / m - mx * mx, acc[5] / m - mx * my, acc[6] / m - mx * mz, acc[7] / m - my * my, acc[8] / m - my * mz,
                  acc[9] / m - mz * mz, w, ev);
    int bi = 0;
    if (w[1] < w[bi]) bi = 1;
    if (w[2] < w[bi]) bi = 2;
    const D3 e = (bi == 0) ? ev[0] : (bi == 1 ? ev[1] : ev[2]);
    const float ex = (float)e.x, ey = (float)e.y, ez = (float)e.z;
    out[0] = ex; out[1] = ey; out[2] = ez;
    out[3] = -1 * __fadd_rn(__fadd_rn(__fmul_rn(ex, (float)mx), __fmul_rn(ey, (float)my)), __fmul_rn(ez, (float)mz));
}

struct SemCam {
    double T[12];  // cam <- lidar, row-major 3x4
    double f, cu, cv;
    int W, H;      // label image size
    unsigned int ground[8];  // 256-bit set of ground labels
};

// pass 1: label lookup -> moments of the labelled points; the last block fits the first model.
// state per frame: acc1[10], acc2[10] doubles | coeffs1[4], coeffs2[4] floats | tickets[2], n_labelled, n_inliers, rc
template <bool WITH_FLAGS>
__global__ void __launch_bounds__(SP_THREADS)
semantic_label_kernel(SemCam C, const float* __restrict__ pts, int stride_f, int n, long long pitch_pts,
                      const unsigned char* __restrict__ labels, double* __restrict__ acc_all, float* __restrict__ coeff_all,
                      unsigned int* __restrict__ ctl_all, int iters, unsigned char* __restrict__ flags_out) {
    const long long frame = blockIdx.y;
    const float* fp = pts + frame * pitch_pts * (long long)stride_f;
    const unsigned char* lab = labels + frame * (long long)C.W * (long long)C.H;
    double* acc = acc_all + frame * 2 * SP_NSUM;
    float* coeff = coeff_all + frame * 8;
    unsigned int* ctl = ctl_all + frame * 8;
    double v[SP_NSUM] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const float ff = (float)C.f, fcu = (float)C.cu, fcv = (float)C.cv, fW = (float)C.W + 1.f, fH = (float)C.H + 1.f;
    // a block streams `iters` tiles of SP_THREADS x SP_PPT points: the ten-moment block reduction at the end costs about as
    // much as 12 points per thread, so batches give every thread 32 points (iters = 8) and a single sweep keeps 4 (iters = 1)
    for (int it = 0; it < iters; it++) {
        const int base = (blockIdx.x * iters + it) * (SP_THREADS * SP_PPT) + threadIdx.x;
        if (base - (int)threadIdx.x >= n) break;  // uniform per block
        float4 p[SP_PPT];
#pragma unroll
        for (int j = 0; j < SP_PPT; j++) {
            const int i = base + j * SP_THREADS;
            p[j] = (i < n) ? ld_stream_f4(fp + (long long)i * stride_f) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // phase A: pixel of every point (or -1); phase B: the label bytes, all loads in flight together; phase C: moments
        int off[SP_PPT];
#pragma unroll
        for (int j = 0; j < SP_PPT; j++) {
            off[j] = -1;
            const int i = base + j * SP_THREADS;
            if (i >= n) continue;
            const double x = p[j].x, y = p[j].y, z = p[j].z;
            // pcl::transformPointCloud: per coordinate in double, left to right, rounded to float
            const float tx = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(C.T[0], x), __dmul_rn(C.T[1], y)), __dmul_rn(C.T[2], z)), C.T[3]);
            const float ty = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(C.T[4], x), __dmul_rn(C.T[5], y)), __dmul_rn(C.T[6], z)), C.T[7]);
            const float tz = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(C.T[8], x), __dmul_rn(C.T[9], y)), __dmul_rn(C.T[10], z)), C.T[11]);
            // float pre-filter on the SAME rounded coordinates the exact path uses, u = f * (tx / tz) + cu: the absolute error of
            // uf, vf is < 0.01 pixel for any coordinate the exact path accepts (-1 < u < W), so such a point never falls outside
            // (-2, W + 1); NaN fails the test and is invalid in the exact path too; for |tz| > 2^126 __fdividef returns 0 and the
            // point simply goes on to the exact path (no product that could overflow is formed). About 85 % of a sweep leaves here.
            const float uf = fmaf(ff, __fdividef(tx, tz), fcu), vf = fmaf(ff, __fdividef(ty, tz), fcv);
            if (!(uf > -2.f && uf < fW && vf > -2.f && vf < fH)) continue;
            // project(): intrin * Vector3d{x,y,z}, p /= p[2], cv::Point(p[0], p[1]) (RansacPlane.cpp:174-177); Eigen's 0*X, 0*Y
            // terms are +-0 for the finite values that reach this point
            const double X = tx, Y = ty, Z = tz;
            const double q0 = __dadd_rn(__dmul_rn(C.f, X), __dmul_rn(C.cu, Z));
            const double q1 = __dadd_rn(__dmul_rn(C.f, Y), __dmul_rn(C.cv, Z));
            const double u = __ddiv_rn(q0, Z), w = __ddiv_rn(q1, Z);
            // double -> int like cvttsd2si: NaN and out-of-range values become INT_MIN, i.e. "x < 0" -> invalid
            if (!(fabs(u) < 2147483648.0) || !(fabs(w) < 2147483648.0)) continue;
            const int px = (int)u, py = (int)w;
            if (px < 0 || px >= C.W || py < 0 || py >= C.H) continue;  // see the deviation note in the header
            // the fit runs on the ORIGINAL cloud (model_p is built on `cloud`, RansacPlane.cpp:236); PCL skips non-finite points
            // (the coordinates are finite here: a non-finite one makes tx, ty or tz non-finite and fails stage 2)
            off[j] = py * C.W + px;
        }
        unsigned int lbl[SP_PPT];
#pragma unroll
        for (int j = 0; j < SP_PPT; j++) lbl[j] = off[j] >= 0 ? (unsigned int)__ldg(lab + off[j]) : 256u;
#pragma unroll
        for (int j = 0; j < SP_PPT; j++) {
            const unsigned int l = lbl[j];
            const bool ground = l <= 255u && ((C.ground[l >> 5] >> (l & 31)) & 1u);
            if (WITH_FLAGS && base + j * SP_THREADS < n) flags_out[frame * (long long)n + base + j * SP_THREADS] = ground ? 1 : 0;  // parity view
            if (!ground) continue;
            const double x = p[j].x, y = p[j].y, z = p[j].z;
            v[0] += 1.0; v[1] += x; v[2] += y; v[3] += z;
            v[4] += x * x; v[5] += x * y; v[6] += x * z; v[7] += y * y; v[8] += y * z; v[9] += z * z;
        }
    }
    if (sp_commit(v, acc, ctl + 0, gridDim.x) && threadIdx.x == 0) {
        volatile double* va = acc;
        double a[SP_NSUM];
        for (int q = 0; q < SP_NSUM; q++) a[q] = va[q];
        ctl[2] = (unsigned int)a[0];                 // labelled points
        const float dummy[4] = {0.f, 0.f, 1.f, 0.f};  // dummy_model_coeffs (RansacPlane.cpp:239-240)
        float c[4];
        sp_fit(a, dummy, c);
        for (int q = 0; q < 4; q++) coeff[q] = c[q];
        ctl[4] = (a[0] < 3.0) ? 1u : 0u;  // ExceptionPclInvalid (RansacPlane.cpp:224-227)
    }
}

// pass 2: selectWithinDistance over the whole cloud -> inlier bitmask + moments; the last block refits.
__global__ void __launch_bounds__(SP_THREADS)
semantic_select_kernel(const float* __restrict__ pts, int stride_f, int n, long long pitch_pts, double threshold,
                       double* __restrict__ acc_all, float* __restrict__ coeff_all, unsigned int* __restrict__ ctl_all,
                       unsigned int* __restrict__ bits_all, long long words_per_frame, float* __restrict__ coeffs_out,
                       int* __restrict__ n_inliers_out, int* __restrict__ rc_out, int iters) {
    const long long frame = blockIdx.y;
    const float* fp = pts + frame * pitch_pts * (long long)stride_f;
    double* acc = acc_all + frame * 2 * SP_NSUM + SP_NSUM;
    float* coeff = coeff_all + frame * 8;
    unsigned int* ctl = ctl_all + frame * 8;
    unsigned int* bits = bits_all + frame * words_per_frame;
    const float a = coeff[0], b = coeff[1], c = coeff[2], d = coeff[3];
    const bool invalid = ctl[4] != 0u;
    double v[SP_NSUM] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int it = 0; it < iters; it++) {
        const int base = (blockIdx.x * iters + it) * (SP_THREADS * SP_PPT) + threadIdx.x;
        if (base - (int)threadIdx.x >= n) break;  // uniform per block
        float4 p[SP_PPT];
#pragma unroll
        for (int j = 0; j < SP_PPT; j++) {
            const int i = base + j * SP_THREADS;
            p[j] = (i < n) ? ld_stream_f4(fp + (long long)i * stride_f) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < SP_PPT; j++) {
            const int i = base + j * SP_THREADS;  // warp-uniform tail: i - lane is a multiple of 32
            bool in = false;
            if (i < n && !invalid) {
                // pcl::SampleConsensusModelPlane::selectWithinDistance: fabs(dot((x,y,z,1), coeffs)) < threshold, float dot
                const float dist = fabsf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, p[j].x), __fmul_rn(b, p[j].y)), __fmul_rn(c, p[j].z)), d));
                in = (double)dist < threshold;
                if (in) {
                    const double x = p[j].x, y = p[j].y, z = p[j].z;
                    v[0] += 1.0; v[1] += x; v[2] += y; v[3] += z;
                    v[4] += x * x; v[5] += x * y; v[6] += x * z; v[7] += y * y; v[8] += y * z; v[9] += z * z;
                }
            }
            const unsigned int m = __ballot_sync(MLD_FULL_MASK, in);
            if ((threadIdx.x & 31) == 0 && (i - (int)(threadIdx.x & 31)) < n) bits[i >> 5] = m;
        }
    }
    if (sp_commit(v, acc, ctl + 1, gridDim.x) && threadIdx.x == 0) {
        volatile double* va = acc;
        double s[SP_NSUM];
        for (int q = 0; q < SP_NSUM; q++) s[q] = va[q];
        const float first[4] = {a, b, c, d};
        float r[4];
        sp_fit(s, first, r);
        for (int q = 0; q < 4; q++) {
            coeff[4 + q] = r[q];
            coeffs_out[frame * 4 + q] = invalid ? 0.f : r[q];
        }
        n_inliers_out[frame] = invalid ? 0 : (int)s[0];
        rc_out[frame] = invalid ? MLD_ERR_PCL_INVALID : 0;
    }
}

// ---- exact mode: PCL's own accumulation order -----------------------------------------------------------------------------
// SampleConsensusModelPlane::optimizeModelCoefficients = computeMeanAndCovarianceMatrix with nine sequential FLOAT accumulators
// over the (finite) inliers in index order, a float covariance, the eigenvector of the smallest eigenvalue (RansacPlane.cpp:
// 236-256 via PCL). Float addition is not associative, so bit-exact coefficients -- and with them a bit-exact inlier set in
// the select pass -- need exactly that order: one block per frame streams the cloud in tiles, compacts the flagged points of a
// tile in index order into shared memory, and lanes 0..8 each run ONE accumulator's dependent chain over the tile (the nine
// chains are independent of each other). ~0.2 ms per 120 k-point sweep and pass instead of ~15 us: the mode exists for parity
// with the reference's numbers, the double-precision moments of the default mode are the better fit.
// FLAG_BITS: flags are a bitmask over raw indices (pass 2: the inlier mask) or one byte per point (pass 1: the labelled set).
constexpr int SX_THREADS = 256, SX_PPT = 4, SX_TILE = SX_THREADS * SX_PPT;

template <bool FLAG_BITS>
__global__ void __launch_bounds__(SX_THREADS)
semantic_exact_fit_kernel(const float* __restrict__ pts, int stride_f, int n, long long pitch_pts, const void* __restrict__ flags_all,
                          long long flags_pitch, float* __restrict__ coeff_all, unsigned int* __restrict__ ctl_all,
                          float* __restrict__ coeffs_out) {
    __shared__ float sx[SX_TILE], sy[SX_TILE], sz[SX_TILE];
    __shared__ int s_warp[SX_THREADS / 32];
    __shared__ float s_acc[9];
    const long long frame = blockIdx.x;
    const float* fp = pts + frame * pitch_pts * (long long)stride_f;
    float* coeff = coeff_all + frame * 8;
    unsigned int* ctl = ctl_all + frame * 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // accumulator k of lane k: acc += a * b with (a, b) = (x,x) (x,y) (x,z) (y,y) (y,z) (z,z) (x,1) (y,1) (z,1); x * 1 is exact
    const int ia = (tid < 3) ? 0 : (tid < 5 ? 1 : (tid == 5 ? 2 : tid - 6));
    const int ib = (tid < 3) ? tid : (tid < 5 ? tid - 2 : (tid == 5 ? 2 : 3));
    float acc = 0.f;
    long long total = 0, flagged = 0;
    for (int base = 0; base < n; base += SX_TILE) {
        float4 p[SX_PPT];
        bool keep[SX_PPT];
        int cnt = 0, nflag = 0;
#pragma unroll
        for (int j = 0; j < SX_PPT; j++) {  // thread t owns the consecutive points base + 4 t .. + 3: ranks follow index order
            const int i = base + tid * SX_PPT + j;
            keep[j] = false;
            if (i < n) {
                bool f;
                if (FLAG_BITS)
                    f = (reinterpret_cast<const unsigned int*>(flags_all)[frame * flags_pitch + (i >> 5)] >> (i & 31)) & 1u;
                else
                    f = reinterpret_cast<const unsigned char*>(flags_all)[frame * flags_pitch + i] != 0;
                if (f) {
                    nflag++;
                    p[j] = __ldg(reinterpret_cast<const float4*>(fp + (long long)i * stride_f));
                    keep[j] = isfinite(p[j].x) && isfinite(p[j].y) && isfinite(p[j].z);  // PCL skips non-finite points
                    cnt += keep[j] ? 1 : 0;
                }
            }
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(MLD_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        // flagged count of the tile (all points the model was asked to fit, finite or not: inliers.size())
        int nf = nflag;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nf += __shfl_xor_sync(MLD_FULL_MASK, nf, o);
        __syncthreads();
        int pos = incl - cnt, tile_cnt = 0;
#pragma unroll
        for (int w = 0; w < SX_THREADS / 32; w++) {
            if (w < warp) pos += s_warp[w];
            tile_cnt += s_warp[w];
        }
#pragma unroll
        for (int j = 0; j < SX_PPT; j++) {
            if (!keep[j]) continue;
            sx[pos] = p[j].x;
            sy[pos] = p[j].y;
            sz[pos] = p[j].z;
            pos++;
        }
        __syncthreads();  // the tile is complete and every thread has read the scan counts
        if (lane == 0) s_warp[warp] = nf;  // now: flagged points per warp
        if (tid < 9) {
            for (int q = 0; q < tile_cnt; q++) {
                const float x = sx[q], y = sy[q], z = sz[q];
                const float a = ia == 0 ? x : (ia == 1 ? y : z);
                const float b = ib == 0 ? x : (ib == 1 ? y : (ib == 2 ? z : 1.f));
                acc = __fadd_rn(acc, __fmul_rn(a, b));
            }
        }
        total += tile_cnt;
        __syncthreads();  // the tile is consumed, the flagged counts are in place
        if (tid == 0)
            for (int w = 0; w < SX_THREADS / 32; w++) flagged += s_warp[w];
        __syncthreads();
    }
    if (tid < 9) s_acc[tid] = acc;
    __syncthreads();
    if (tid != 0) return;
    const bool second = FLAG_BITS;
    float in[4] = {0.f, 0.f, 1.f, 0.f};  // dummy_model_coeffs (RansacPlane.cpp:239-240)
    if (second)
        for (int q = 0; q < 4; q++) in[q] = coeff[q];
    float out[4] = {in[0], in[1], in[2], in[3]};
    if (flagged > 3 && total > 0) {  // inliers.size() <= 3: the input model is returned; no finite point: likewise
        float a9[9];
        const float cntf = (float)total;
        for (int q = 0; q < 9; q++) a9[q] = __fdiv_rn(s_acc[q], cntf);
        const float cx = a9[6], cy = a9[7], cz = a9[8];
        double w[3];
        D3 ev[3];
        eig3_sym_regs((double)__fsub_rn(a9[0], __fmul_rn(cx, cx)), (double)__fsub_rn(a9[1], __fmul_rn(cx, cy)),
                      (double)__fsub_rn(a9[2], __fmul_rn(cx, cz)), (double)__fsub_rn(a9[3], __fmul_rn(cy, cy)),
                      (double)__fsub_rn(a9[4], __fmul_rn(cy, cz)), (double)__fsub_rn(a9[5], __fmul_rn(cz, cz)), w, ev);
        out[0] = (float)ev[0].x;
        out[1] = (float)ev[0].y;
        out[2] = (float)ev[0].z;
        out[3] = __fmul_rn(-1.f, __fadd_rn(__fadd_rn(__fmul_rn(out[0], cx), __fmul_rn(out[1], cy)), __fmul_rn(out[2], cz)));
    }
    if (!second) {
        for (int q = 0; q < 4; q++) coeff[q] = out[q];
        ctl[2] = (unsigned int)flagged;
        ctl[4] = (flagged < 3) ? 1u : 0u;  // ExceptionPclInvalid (RansacPlane.cpp:224-227)
    } else {
        const bool invalid = ctl[4] != 0u;
        for (int q = 0; q < 4; q++) {
            coeff[4 + q] = out[q];
            coeffs_out[frame * 4 + q] = invalid ? 0.f : out[q];
        }
    }
}

}  // namespace

size_t mld_semantic_state_bytes(int nframes) { return (size_t)nframes * (2 * SP_NSUM * sizeof(double) + 8 * sizeof(float) + 8 * sizeof(unsigned int)); }

cudaError_t mld_launch_semantic_plane(const double* T_cam_lidar, double f, double cu, double cv, int label_w, int label_h,
                                      const unsigned int* ground_set8, double inlier_threshold, const float* d_pts, int stride_f,
                                      long long n_points, long long pitch_pts, const unsigned char* d_labels, int nframes,
                                      void* d_state, float* d_coeffs, unsigned int* d_inlier_bits, long long words_per_frame,
                                      int* d_n_inliers, int* d_rc, cudaStream_t stream, int* launches, unsigned char* d_flags_out,
                                      int exact) {
    if (nframes <= 0) return cudaSuccess;
    if (exact && !d_flags_out) return cudaErrorInvalidValue;  // the exact mode fits from the labelled flags
    if (n_points > 0x7fffffffLL / 8) return cudaErrorInvalidValue;
    SemCam C;
    for (int i = 0; i < 12; i++) C.T[i] = T_cam_lidar[i];
    C.f = f; C.cu = cu; C.cv = cv; C.W = label_w; C.H = label_h;
    for (int i = 0; i < 8; i++) C.ground[i] = ground_set8[i];
    cudaError_t e = cudaMemsetAsync(d_state, 0, mld_semantic_state_bytes(nframes), stream);
    if (e != cudaSuccess) return e;
    double* acc = reinterpret_cast<double*>(d_state);
    float* coeff = reinterpret_cast<float*>(acc + (size_t)nframes * 2 * SP_NSUM);
    unsigned int* ctl = reinterpret_cast<unsigned int*>(coeff + (size_t)nframes * 8);
    // tiles per block: a batch fills the machine with frames, so each block amortises its reduction over 8 tiles; one sweep
    // alone needs all the blocks it can get
    const int iters = nframes >= 8 ? 8 : 1;
    const long long per_block = (long long)SP_THREADS * SP_PPT * iters;
    const unsigned int gx = (unsigned int)std::max<long long>(1, (n_points + per_block - 1) / per_block);
    dim3 grid(gx, (unsigned)nframes);
    if (d_flags_out)
        semantic_label_kernel<true><<<grid, SP_THREADS, 0, stream>>>(C, d_pts, stride_f, (int)n_points, pitch_pts, d_labels, acc, coeff, ctl, iters, d_flags_out);
    else
        semantic_label_kernel<false><<<grid, SP_THREADS, 0, stream>>>(C, d_pts, stride_f, (int)n_points, pitch_pts, d_labels, acc, coeff, ctl, iters, nullptr);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (exact) {  // first model from PCL's sequential float moments of the labelled points (replaces the double-precision fit)
        semantic_exact_fit_kernel<false><<<(unsigned)nframes, SX_THREADS, 0, stream>>>(d_pts, stride_f, (int)n_points, pitch_pts, d_flags_out,
                                                                                     n_points, coeff, ctl, d_coeffs);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (launches) (*launches)++;
    }
    semantic_select_kernel<<<grid, SP_THREADS, 0, stream>>>(d_pts, stride_f, (int)n_points, pitch_pts, inlier_threshold, acc, coeff, ctl,
                                                           d_inlier_bits, words_per_frame, d_coeffs, d_n_inliers, d_rc, iters);
    if (launches) *launches += 2;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (exact) {  // refit from the inlier mask in the same order
        semantic_exact_fit_kernel<true><<<(unsigned)nframes, SX_THREADS, 0, stream>>>(d_pts, stride_f, (int)n_points, pitch_pts, d_inlier_bits,
                                                                                    words_per_frame, coeff, ctl, d_coeffs);
        if (launches) (*launches)++;
    }
    return cudaGetLastError();
}
