// mld_ransac.cu -- K4: per-frame RANSAC ground-plane fit, one thread-block cluster per frame.
//
// Replaces RansacPlane::CalculateInliersPlane (/root/reference/monolidar_fusion/src/RansacPlane.cpp:41-140)
// and the PCL pieces it drives (PCL is not vendored in the reference; 1.8 semantics restated):
//   PassThrough("z")                         -> candidate list                         (:58-64)
//   RandomSample(6000)                       -> order-preserving stratified subsample  (:66-74)
//   SampleConsensusModelPerpendicularPlane   -> axis (0,0,1), eps 10 degrees           (:94-100)
//   RandomSampleConsensus::computeModel      -> adaptive loop, k = log(1-p)/log(1-w^3) (:102-108)
//   optimizeModelCoefficients                -> centroid + smallest covariance eigvec  (:117-126)
//   selectWithinDistance(un-refined coeffs)  -> final inlier set                       (:121)
//
// PCL's RandomSample and sample draws are time/rand() seeded, so the reference's hypotheses cannot be
// reproduced by anyone ("parity unpinned"). The draws here come from a counter-based
// hash of (seed, draw index, attempt) -- the same definition the CPU oracle restates -- which makes
// the sequential semantics (first best hypothesis wins, adaptive stop) reproducible in parallel.
//
// Mapping. Two instantiations of one body. Batches (>= RS_BATCH_FRAMES frames per launch): ONE CTA per frame with the whole
// 6000-point subsample in its shared memory (96 KB, two frames per SM) -- no cluster barriers, every SM busy with its own
// frames (252 -> ~100 us per 512 frames). Few frames (the per-call drop-in path): a cluster of 8 CTAs owns one frame, below.
// A cluster of 8 CTAs owns one frame. CTA r keeps sample points [750 r, 750 (r+1)) in shared
// memory. Per round, thread t of every CTA derives hypothesis (round*256 + t), scores it against the
// CTA's slice with broadcast shared-memory reads, and the 8 partial counts are summed through
// distributed shared memory. Every CTA then replays PCL's sequential update over the 256 totals
// (identical data => identical decision), so the cluster agrees on "done" without another exchange.
#include <cooperative_groups.h>

#include "mld_common.cuh"
#include "mld_kernels.h"

namespace cg = cooperative_groups;

namespace {

constexpr int RS_CLUSTER = 8;
constexpr int RS_THREADS = 256;
constexpr int RS_BATCH_FRAMES = 64;  // launches of at least this many frames use one CTA per frame
constexpr int RS_MAX_SAMPLE_CHECKS = 1000;  // pcl::SampleConsensusModel::max_sample_checks_
// PCL's adaptive loop usually stops after 50-250 iterations, so a round scores only RS_HYP hypotheses;
// the RS_THREADS / RS_HYP threads that share a hypothesis split the CTA's slice of the sample.
#ifndef MLD_RS_HYP
#define MLD_RS_HYP 32
#endif
constexpr int RS_HYP = MLD_RS_HYP;
constexpr int RS_PARTS = RS_THREADS / RS_HYP;

struct F3 {
    float x, y, z;
};

// SampleConsensusModelPlane::isSampleGood: (p1-p0)/(p2-p0) componentwise, collinear when all equal
__device__ __forceinline__ bool sample_good(const F3& p0, const F3& p1, const F3& p2) {
    float r0 = __fdiv_rn(__fsub_rn(p1.x, p0.x), __fsub_rn(p2.x, p0.x));
    float r1 = __fdiv_rn(__fsub_rn(p1.y, p0.y), __fsub_rn(p2.y, p0.y));
    float r2 = __fdiv_rn(__fsub_rn(p1.z, p0.z), __fsub_rn(p2.z, p0.z));
    return (r0 != r1) || (r2 != r1);
}
// SampleConsensusModelPlane::computeModelCoefficients
__device__ __forceinline__ void plane_from_sample(const F3& p0, const F3& p1, const F3& p2, float c[4]) {
    float ax = __fsub_rn(p1.x, p0.x), ay = __fsub_rn(p1.y, p0.y), az = __fsub_rn(p1.z, p0.z);
    float bx = __fsub_rn(p2.x, p0.x), by = __fsub_rn(p2.y, p0.y), bz = __fsub_rn(p2.z, p0.z);
    float nx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
    float ny = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
    float nz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
    float z = __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz));
    if (z > 0) {
        float s = __fsqrt_rn(z);
        nx = __fdiv_rn(nx, s); ny = __fdiv_rn(ny, s); nz = __fdiv_rn(nz, s);
    }
    c[0] = nx; c[1] = ny; c[2] = nz;
    c[3] = -1 * __fadd_rn(__fadd_rn(__fmul_rn(nx, p0.x), __fmul_rn(ny, p0.y)), __fmul_rn(nz, p0.z));
}
// SampleConsensusModelPerpendicularPlane::isModelValid with axis z: |n_z|/|n| >= cos(eps)
__device__ __forceinline__ bool model_valid(const float c[4], double cos_eps) {
    float z = __fadd_rn(__fadd_rn(__fmul_rn(c[0], c[0]), __fmul_rn(c[1], c[1])), __fmul_rn(c[2], c[2]));
    float nz = c[2];
    if (z > 0) nz = __fdiv_rn(c[2], __fsqrt_rn(z));
    return fabs((double)nz) >= cos_eps;
}
// |a x + b y + c z + d| evaluated in float, left to right, no contraction (PCL's Vector4f arithmetic).
// PCL compares (double)dist < (double)threshold; with thr_lt = the largest float whose double value is
// below the threshold this is exactly  dist <= thr_lt  in float (see ransac_config()).
__device__ __forceinline__ float plane_dist(float a, float b, float c, float d, float x, float y, float z) {
    return fabsf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, x), __fmul_rn(b, y)), __fmul_rn(c, z)), d));
}

struct FrameView {
    const float* pts;
    int stride_f;
    long long n;
    const int* cand;  // nullptr: identity
    long long L;      // candidate count
    long long M;      // subsample size
    uint64_t seed;
};

__device__ __forceinline__ int sub_raw(const FrameView& fv, long long j) {
    long long pos;
    if (fv.L <= MLD_RANSAC_SAMPLE) {
        pos = j;
    } else {
        long long lo = (j * fv.L) / MLD_RANSAC_SAMPLE, hi = ((j + 1) * fv.L) / MLD_RANSAC_SAMPLE;
        unsigned long long len = (unsigned long long)(hi - lo);
        pos = lo + (long long)(mld_hash3(fv.seed, 0x5A17, (uint64_t)j, 0) % len);
    }
    return fv.cand ? fv.cand[pos] : (int)pos;
}
__device__ __forceinline__ F3 load_pt(const FrameView& fv, int raw) {
    const float* p = fv.pts + (long long)raw * fv.stride_f;
    return F3{__ldg(p), __ldg(p + 1), __ldg(p + 2)};
}

// PassThrough("z", min_z, max_z): order-preserving candidate list, one block per frame.
__global__ void __launch_bounds__(1024)
ransac_candidates_kernel(RansacConfig cfg, const float* __restrict__ pts, int stride_f, long long n, long long pitch_pts,
                         int* __restrict__ cand_all, int* __restrict__ cand_count) {
    __shared__ int warp_tot[32];
    __shared__ int running;
    const long long frame = blockIdx.x;
    const float* fp = pts + frame * pitch_pts * (long long)stride_f;
    int* cand = cand_all + frame * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
        long long i = base + threadIdx.x;
        bool keep = false;
        if (i < n) {
            // pcl::PassThrough::applyFilterIndices drops every point with a non-finite x, y or z before it tests the field
            const float x = fp[i * stride_f], y = fp[i * stride_f + 1], z = fp[i * stride_f + 2];
            keep = isfinite(x) && isfinite(y) && isfinite(z) && !((double)z < cfg.min_z) && !((double)z > cfg.max_z);
        }
        unsigned m = __ballot_sync(MLD_FULL_MASK, keep);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int off = running;
        for (int w = 0; w < warp; w++) off += warp_tot[w];
        if (keep) cand[off + __popc(m & ((1u << lane) - 1u))] = (int)i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < 32; w++) tot += warp_tot[w];
            running += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) cand_count[frame] = running;
}

__device__ __forceinline__ double block_sum_d(double v, double* scratch /* >= 8 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0;
    for (int w = 0; w < RS_THREADS / 32; w++) t += scratch[w];
    return t;
}

// NC = CTAs that share a frame (thread-block cluster of NC, or 1); s_pts = NC-th part of the subsample (dynamic shared memory)
template <int NC>
__device__ __forceinline__ void ransac_frame(const RansacConfig& cfg, const float* __restrict__ pts, int stride_f, long long n, long long pitch_pts,
                                             uint64_t seed, long long frame0, const int* __restrict__ cand_all, const int* __restrict__ cand_count,
                                             float* __restrict__ out_coeffs, unsigned int* __restrict__ out_bits, long long words_per_frame,
                                             int* __restrict__ out_n_inliers, int* __restrict__ out_iterations, int* __restrict__ out_rc,
                                             float4* s_pts, long long frame) {
    constexpr int RS_SLICE = (MLD_RANSAC_SAMPLE + NC - 1) / NC;  // 750 of the 6000 sample points per CTA of a cluster of 8
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = NC > 1 ? cluster.block_rank() : 0u;
    const int tid = threadIdx.x;
    auto cluster_sync = [&]() {
        if (NC > 1)
            cluster.sync();
        else
            __syncthreads();
    };
    // s_pts: x, y, z, raw index bits
    __shared__ int s_partial[RS_HYP];          // this CTA's counts for the round's hypotheses
    __shared__ int s_total[RS_HYP];
    __shared__ float s_hyp[RS_HYP][4];
    __shared__ unsigned char s_nosample[RS_HYP], s_valid[RS_HYP];
    __shared__ float s_best[4];
    __shared__ int s_state[4];                 // done, have_model, iterations, n_best
    __shared__ double s_k;
    __shared__ double s_khyp[RS_HYP];         // iteration bound if hypothesis t becomes the best
    __shared__ double s_red[8];
    __shared__ double s_sums[10];              // this CTA's refinement partial sums
    __shared__ double s_wsum[RS_THREADS / 32][10];

    FrameView fv;
    fv.pts = pts + frame * pitch_pts * (long long)stride_f;
    fv.stride_f = stride_f;
    fv.n = n;
    const bool pass = cfg.min_z > -1001.;
    fv.cand = pass ? cand_all + frame * n : nullptr;
    fv.L = pass ? (long long)cand_count[frame] : n;
    fv.M = fv.L < MLD_RANSAC_SAMPLE ? fv.L : MLD_RANSAC_SAMPLE;
    fv.seed = seed + (uint64_t)(frame0 + frame);
    const long long M = fv.M;

    if (n < 3 || M < 3) {  // ExceptionPclInvalid (RansacPlane.cpp:44-50) / no model
        if (rank == 0 && tid == 0) {
            out_rc[frame] = (n < 3) ? MLD_ERR_PCL_INVALID : MLD_ERR_NO_MODEL;
            out_n_inliers[frame] = 0;
            out_iterations[frame] = 0;
            for (int q = 0; q < 4; q++) out_coeffs[frame * 4 + q] = 0.f;
        }
        return;  // uniform across the cluster
    }

    // stage this CTA's slice of the subsample
    const long long j0 = (long long)rank * RS_SLICE;
    long long cnt_ll = M - j0;
    const int slice_n = cnt_ll <= 0 ? 0 : (cnt_ll < RS_SLICE ? (int)cnt_ll : RS_SLICE);
    // four points per pass: their index chains (hash, 64-bit divisions) and random 12-byte reads are independent, one after the other
    // a CTA that stages the whole subsample (24 points per thread) spent more time here than scoring its hypotheses
    for (int q0 = tid; q0 < slice_n; q0 += 4 * RS_THREADS) {
        int raw[4];
        F3 p[4];
#pragma unroll
        for (int u = 0; u < 4; u++) raw[u] = (q0 + u * RS_THREADS < slice_n) ? sub_raw(fv, j0 + q0 + u * RS_THREADS) : 0;
#pragma unroll
        for (int u = 0; u < 4; u++) p[u] = load_pt(fv, raw[u]);
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (q0 + u * RS_THREADS < slice_n) s_pts[q0 + u * RS_THREADS] = make_float4(p[u].x, p[u].y, p[u].z, __int_as_float(raw[u]));
    }
    if (tid == 0) {
        s_state[0] = 0; s_state[1] = 0; s_state[2] = 0; s_state[3] = -2147483647;
        s_k = 1.0;
        s_best[0] = s_best[1] = s_best[2] = s_best[3] = 0.f;
    }
    cluster_sync();  // every slice is staged: the hypotheses read their sample points from any CTA's slice

    const double cos_eps = cfg.cos_eps;  // cos(M_PI / 18.), evaluated on the host
    const double log_probability = cfg.log_probability;  // log(1 - probability), evaluated on the host
    const unsigned max_skip = (unsigned)cfg.max_iterations * 10u;
    const double one_over_indices = 1.0 / (double)M;

    const float thr_lt = cfg.thr_lt;
    const int hyp = tid & (RS_HYP - 1), part = tid / RS_HYP;
    const int per_part = (slice_n + RS_PARTS - 1) / RS_PARTS;
    const int q_begin = part * per_part, q_end = min(slice_n, q_begin + per_part);
    for (int round = 0;; round++) {
        // ---- hypothesis (round*RS_HYP + tid), tid < RS_HYP: draw, model, validity ------------------
        if (tid < RS_HYP) {
            const uint64_t draw = (uint64_t)round * RS_HYP + (uint64_t)tid;
            bool got = false;
            F3 p0{}, p1{}, p2{};
            for (int a = 0; a < RS_MAX_SAMPLE_CHECKS && !got; a++) {
                uint64_t h0 = mld_hash3(fv.seed, draw, (uint64_t)a, 0), h1 = mld_hash3(fv.seed, draw, (uint64_t)a, 1),
                         h2 = mld_hash3(fv.seed, draw, (uint64_t)a, 2);
                long long i0 = (long long)(h0 % (uint64_t)M);
                long long i1 = (long long)(h1 % (uint64_t)(M - 1));
                if (i1 >= i0) i1++;
                long long i2 = (long long)(h2 % (uint64_t)(M - 2));
                long long lo = i0 < i1 ? i0 : i1, hi = i0 < i1 ? i1 : i0;
                if (i2 >= lo) i2++;
                if (i2 >= hi) i2++;
                // sample point i is entry i % RS_SLICE of CTA i / RS_SLICE's staged slice (own or distributed shared memory)
                auto staged = [&](long long i) {
                    const int r = (int)(i / RS_SLICE), q = (int)(i - (long long)r * RS_SLICE);
                    const float4 v = NC > 1 ? *cluster.map_shared_rank(&s_pts[q], (unsigned)r) : s_pts[q];
                    return F3{v.x, v.y, v.z};
                };
                p0 = staged(i0);
                p1 = staged(i1);
                p2 = staged(i2);
                got = sample_good(p0, p1, p2);
            }
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            bool valid = false;
            if (got) {
                plane_from_sample(p0, p1, p2, c);
                valid = model_valid(c, cos_eps);
            }
            s_nosample[tid] = got ? 0 : 1;
            s_valid[tid] = valid ? 1 : 0;
            s_hyp[tid][0] = c[0]; s_hyp[tid][1] = c[1]; s_hyp[tid][2] = c[2]; s_hyp[tid][3] = c[3];
            s_partial[tid] = 0;
        }
        __syncthreads();
        // ---- score against this CTA's slice (countWithinDistance); RS_PARTS threads share a hypothesis ----
        if (s_valid[hyp]) {
            const float c0 = s_hyp[hyp][0], c1 = s_hyp[hyp][1], c2 = s_hyp[hyp][2], c3 = s_hyp[hyp][3];
            int count = 0;
            for (int q = q_begin; q < q_end; q++) {
                float4 p = s_pts[q];
                if (plane_dist(c0, c1, c2, c3, p.x, p.y, p.z) <= thr_lt) count++;
            }
            if (count) atomicAdd(&s_partial[hyp], count);
        }
        cluster_sync();
        if (tid < RS_HYP) {
            int total = 0;
            if (NC > 1) {
#pragma unroll
                for (unsigned r = 0; r < NC; r++) total += *cluster.map_shared_rank(&s_partial[tid], r);
            } else {
                total = s_partial[tid];
            }
            s_total[tid] = total;
            // the iteration bound PCL would compute if this hypothesis became the best one (RandomSampleConsensus::computeModel):
            // evaluated here by 64 threads at once -- pow and log in double on one thread, 15-20 times per frame, were the longest
            // serial stretch of the kernel -- and only picked up by the sequential replay below
            double w = (double)total * one_over_indices;
            double p_no_outliers = 1.0 - pow(w, 3.0);
            p_no_outliers = fmax(2.220446049250313e-16, p_no_outliers);
            p_no_outliers = fmin(1.0 - 2.220446049250313e-16, p_no_outliers);
            s_khyp[tid] = log_probability / log(p_no_outliers);
        }
        __syncthreads();
        // ---- PCL's sequential update over the round (RandomSampleConsensus::computeModel) ------
        if (tid == 0) {
            int iterations = s_state[2], n_best = s_state[3], have = s_state[1], done = 0, best_t = -1;
            double k = s_k;
            for (int t = 0; t < RS_HYP; t++) {
                if (!((double)iterations < k) || !(0u < max_skip)) { done = 1; break; }
                if (s_nosample[t]) { done = 1; break; }  // "No samples could be selected!"
                int cnt = s_total[t];
                if (cnt > n_best) {
                    n_best = cnt;
                    best_t = t;
                    have = 1;
                    k = s_khyp[t];
                }
                ++iterations;
                if (iterations > cfg.max_iterations) { done = 1; break; }
            }
            if (best_t >= 0) {
                s_best[0] = s_hyp[best_t][0]; s_best[1] = s_hyp[best_t][1]; s_best[2] = s_hyp[best_t][2]; s_best[3] = s_hyp[best_t][3];
            }
            s_state[0] = done; s_state[1] = have; s_state[2] = iterations; s_state[3] = n_best;
            s_k = k;
        }
        __syncthreads();
        const int done = s_state[0];
        cluster_sync();  // every CTA has consumed the partials before the next round overwrites them
        if (done) break;
    }

    const int have_model = s_state[1];
    if (!have_model) {
        if (rank == 0 && tid == 0) {
            out_rc[frame] = MLD_ERR_NO_MODEL;
            out_n_inliers[frame] = 0;
            out_iterations[frame] = s_state[2];
            for (int q = 0; q < 4; q++) out_coeffs[frame * 4 + q] = 0.f;
        }
        return;
    }
    const float b0 = s_best[0], b1 = s_best[1], b2 = s_best[2], b3 = s_best[3];
    const float best[4] = {b0, b1, b2, b3};
    const bool best_valid = model_valid(best, cos_eps);
    float outc[4] = {b0, b1, b2, b3};

    if (cfg.use_refinement) {
        // optimizeModelCoefficients on the RANSAC inliers: centroid + covariance in double
        double v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (best_valid) {
            for (int q = tid; q < slice_n; q += RS_THREADS) {
                float4 p = s_pts[q];
                if (plane_dist(b0, b1, b2, b3, p.x, p.y, p.z) <= thr_lt) {
                    double x = p.x, y = p.y, z = p.z;
                    v[0] += 1.0; v[1] += x; v[2] += y; v[3] += z;
                    v[4] += x * x; v[5] += x * y; v[6] += x * z; v[7] += y * y; v[8] += y * z; v[9] += z * z;
                }
            }
        }
        // block sums of the ten moments: butterfly inside the warp, then the warps in order (the order block_sum_d uses), with
        // two barriers for all ten instead of two each
        {
            const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
            for (int q = 0; q < 10; q++) {
                const double t = warp_sum_d(v[q]);
                if (lane == 0) s_wsum[warp][q] = t;
            }
            __syncthreads();
            if (tid < 10) {
                double t = 0;
                for (int w = 0; w < RS_THREADS / 32; w++) t += s_wsum[w][tid];
                s_sums[tid] = t;
            }
        }
        cluster_sync();
        // cluster totals and the eigen solve: only the thread that writes the coefficients needs them
        if (rank == 0 && tid == 0) {
            double tot[10];
            for (int q = 0; q < 10; q++) {
                double t = 0;
                if (NC > 1) {
                    for (unsigned r = 0; r < NC; r++) t += *cluster.map_shared_rank(&s_sums[q], r);
                } else {
                    t = s_sums[q];
                }
                tot[q] = t;
            }
            if (tot[0] >= 4.0) {
                double m = tot[0];
                double mx = tot[1] / m, my = tot[2] / m, mz = tot[3] / m;
                double w[3];
                D3 ev[3];
                eig3_sym_regs(tot[4] / m - mx * mx, tot[5] / m - mx * my, tot[6] / m - mx * mz, tot[7] / m - my * my,
                              tot[8] / m - my * mz, tot[9] / m - mz * mz, w, ev);
                int bi = 0;
                if (w[1] < w[bi]) bi = 1;
                if (w[2] < w[bi]) bi = 2;
                D3 e = (bi == 0) ? ev[0] : (bi == 1 ? ev[1] : ev[2]);
                float ex = (float)e.x, ey = (float)e.y, ez = (float)e.z;
                float cand4[4] = {ex, ey, ez, 0.f};
                cand4[3] = -1 * __fadd_rn(__fadd_rn(__fmul_rn(ex, (float)mx), __fmul_rn(ey, (float)my)), __fmul_rn(ez, (float)mz));
                if (model_valid(cand4, cos_eps)) {
                    outc[0] = cand4[0]; outc[1] = cand4[1]; outc[2] = cand4[2]; outc[3] = cand4[3];
                }
            }
        }
        cluster_sync();  // the other CTAs keep their partial sums alive until rank 0 has read them
    }
    // final inlier set: selectWithinDistance with the UN-refined coefficients (RansacPlane.cpp:121)
    const float final_lt = cfg.use_refinement ? cfg.refine_lt : thr_lt;
    unsigned int* bits = out_bits + frame * words_per_frame;
    int my_inl = 0;
    if (best_valid) {
        for (int q = tid; q < slice_n; q += RS_THREADS) {
            float4 p = s_pts[q];
            if (plane_dist(b0, b1, b2, b3, p.x, p.y, p.z) <= final_lt) {
                int raw = __float_as_int(p.w);
                atomicOr(&bits[raw >> 5], 1u << (raw & 31));
                my_inl++;
            }
        }
    }
    double blk = block_sum_d((double)my_inl, s_red);
    if (tid == 0) atomicAdd(&out_n_inliers[frame], (int)blk);
    if (rank == 0 && tid == 0) {
        out_rc[frame] = 0;
        out_iterations[frame] = s_state[2];
        for (int q = 0; q < 4; q++) out_coeffs[frame * 4 + q] = outc[q];
    }
}

#ifndef MLD_RS_MINBLOCKS
#define MLD_RS_MINBLOCKS 4
#endif
__global__ void __cluster_dims__(RS_CLUSTER, 1, 1) __launch_bounds__(RS_THREADS, MLD_RS_MINBLOCKS)
ransac_cluster_kernel(RansacConfig cfg, const float* __restrict__ pts, int stride_f, long long n, long long pitch_pts,
                      uint64_t seed, long long frame0, const int* __restrict__ cand_all, const int* __restrict__ cand_count,
                      float* __restrict__ out_coeffs, unsigned int* __restrict__ out_bits, long long words_per_frame,
                      int* __restrict__ out_n_inliers, int* __restrict__ out_iterations, int* __restrict__ out_rc) {
    extern __shared__ float4 rs_dyn_pts[];
    ransac_frame<RS_CLUSTER>(cfg, pts, stride_f, n, pitch_pts, seed, frame0, cand_all, cand_count, out_coeffs, out_bits, words_per_frame,
                             out_n_inliers, out_iterations, out_rc, rs_dyn_pts, (long long)blockIdx.y);
}

__global__ void __launch_bounds__(RS_THREADS, 2)
ransac_frame_kernel(RansacConfig cfg, const float* __restrict__ pts, int stride_f, long long n, long long pitch_pts,
                    uint64_t seed, long long frame0, const int* __restrict__ cand_all, const int* __restrict__ cand_count,
                    float* __restrict__ out_coeffs, unsigned int* __restrict__ out_bits, long long words_per_frame,
                    int* __restrict__ out_n_inliers, int* __restrict__ out_iterations, int* __restrict__ out_rc) {
    extern __shared__ float4 rs_dyn_pts[];
    ransac_frame<1>(cfg, pts, stride_f, n, pitch_pts, seed, frame0, cand_all, cand_count, out_coeffs, out_bits, words_per_frame,
                    out_n_inliers, out_iterations, out_rc, rs_dyn_pts, (long long)blockIdx.x);
}

}  // namespace

size_t mld_ransac_scratch_bytes(long long n_points, int nframes) {
    // candidate list (int32 per point) + candidate count per frame
    return (size_t)nframes * ((size_t)n_points * sizeof(int) + 64);
}

cudaError_t mld_launch_ransac(const RansacConfig& cfg, const float* d_pts, int stride_f, long long n_points,
                              long long pitch_pts, int nframes, uint64_t seed, long long frame0, void* d_scratch,
                              float* d_coeffs, unsigned int* d_inlier_bits, long long words_per_frame, int* d_n_inliers,
                              int* d_iterations, int* d_rc, cudaStream_t stream, int* launches) {
    if (nframes <= 0) return cudaSuccess;
    cudaError_t e;
    e = cudaMemsetAsync(d_inlier_bits, 0, (size_t)nframes * (size_t)words_per_frame * sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(d_n_inliers, 0, (size_t)nframes * sizeof(int), stream);
    if (e != cudaSuccess) return e;
    int* cand_count = reinterpret_cast<int*>(d_scratch);
    int* cand_all = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(d_scratch) + (size_t)nframes * 64);
    if (cfg.min_z > -1001. && n_points > 0) {
        ransac_candidates_kernel<<<(unsigned)nframes, 1024, 0, stream>>>(cfg, d_pts, stride_f, n_points, pitch_pts, cand_all,
                                                                       cand_count);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (launches) (*launches)++;
    }
    if (nframes >= RS_BATCH_FRAMES) {
        constexpr size_t smem1 = (size_t)MLD_RANSAC_SAMPLE * sizeof(float4);  // 96 KB: opt in once per device
        static thread_local int configured_device = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (configured_device != dev) {
            e = cudaFuncSetAttribute(ransac_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
            if (e != cudaSuccess) return e;
            configured_device = dev;
        }
        ransac_frame_kernel<<<(unsigned)nframes, RS_THREADS, smem1, stream>>>(cfg, d_pts, stride_f, n_points, pitch_pts, seed, frame0, cand_all,
                                                                            cand_count, d_coeffs, d_inlier_bits, words_per_frame, d_n_inliers,
                                                                            d_iterations, d_rc);
    } else {
        dim3 grid(RS_CLUSTER, (unsigned)nframes);
        constexpr size_t smem8 = (size_t)((MLD_RANSAC_SAMPLE + RS_CLUSTER - 1) / RS_CLUSTER) * sizeof(float4);
        ransac_cluster_kernel<<<grid, RS_THREADS, smem8, stream>>>(cfg, d_pts, stride_f, n_points, pitch_pts, seed, frame0, cand_all,
                                                                  cand_count, d_coeffs, d_inlier_bits, words_per_frame, d_n_inliers,
                                                                  d_iterations, d_rc);
    }
    e = cudaGetLastError();
    if (launches) (*launches)++;
    return e;
}
