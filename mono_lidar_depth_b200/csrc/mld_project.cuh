// mld_project.cuh -- device side of K1 (lidar -> camera transform, projection, frustum cull, first-point-wins scatter),
// shared by project_scatter_kernel (mld_project.cu) and the fused K1 + gather launch (mld_feature_split.cu).
// See mld_project.cu for the reference routines restated and the roofline notes.
#pragma once
#include "mld_common.cuh"

namespace {

#ifndef MLD_K1_THREADS
#define MLD_K1_THREADS 128
#endif
#ifndef MLD_K1_PPT
#define MLD_K1_PPT 8
#endif
#ifndef MLD_K1_MINBLOCKS
#define MLD_K1_MINBLOCKS 8
#endif
constexpr int K1_THREADS = MLD_K1_THREADS;
constexpr int K1_PPT = MLD_K1_PPT;  // points per thread: independent 16-byte loads in flight

// exact projection of one point; returns false when the point is culled, else its image coordinates and camera-frame point
__device__ __forceinline__ bool project_uv(const DevParams& P, float x, float y, float z, bool need_front, double& u, double& v, D3& c);

// exact projection of one point; returns false when the point does not enter the map, else its pixel
__device__ __forceinline__ bool project_pixel(const DevParams& P, float x, float y, float z, bool need_front, int& px, int& py) {
    double u, v;
    D3 c;
    if (!project_uv(P, x, y, z, need_front, u, v, c)) return false;
    px = (int)u;  // int x_img = u; int y_img = v (NeighborFinderPixel.cpp:41-42)
    py = (int)v;
    return true;
}

__device__ __forceinline__ bool project_uv(const DevParams& P, float x, float y, float z, bool need_front, double& u, double& v, D3& c) {
    c = lidar_to_cam(P, x, y, z);
    // the map only accepts points in front of the camera (NeighborFinderPixel.cpp:51)
    if (need_front && !(c.z > 0.0)) return false;
    // K * p with K = [f 0 cx; 0 f cy; 0 0 1] (camera_pinhole.h:88), then colwise().hnormalized() = division by the third
    // row (:90). Eigen's product also adds the terms 0*X, 0*Y: they are +-0 for finite coordinates and change no value
    // (at most the sign of a zero numerator, which fails u > 0 / v > 0 either way); for non-finite coordinates they make
    // the quotient NaN, and so does the division below (inf/inf) or the point fails the bounds as +-inf. Left out.
    const double q0 = __dadd_rn(__dmul_rn(P.f, c.x), __dmul_rn(P.cx, c.z));
    const double q1 = __dadd_rn(__dmul_rn(P.f, c.y), __dmul_rn(P.cy, c.z));
    u = __ddiv_rn(q0, c.z);
    v = __ddiv_rn(q1, c.z);
    bool in_range = (u >= 0.) && (u <= P.Wd) && (v >= 0.) && (v <= P.Hd);  // camera_pinhole.h:93-96
    bool visible = (u > 0.) && (u < P.Wd) && (v > 0.) && (v < P.Hd);       // DepthEstimator.cpp:186-187
    return in_range && visible;
}

// FP32 pre-filter: true when the point certainly fails one of  z_cam > 0, u > 0, u < W, v > 0, v < H
// (or is not finite, which fails all of them in the exact path as well).
__device__ __forceinline__ float pf_form(const DevParams& P, int k, float x, float y, float z) {
    return fmaf(P.pf_g[k][0], x, fmaf(P.pf_g[k][1], y, fmaf(P.pf_g[k][2], z, P.pf_h[k])));
}
__device__ __forceinline__ bool surely_outside(const DevParams& P, float x, float y, float z) {
    const float S = fabsf(x) + fabsf(y) + fabsf(z);
    // tests ordered by how much of a 360-degree sweep they remove; a sweep is azimuth ordered, so the early exits are
    // nearly warp uniform. Each test is written so that a NaN (dropout) is rejected by the first one; an infinite
    // coordinate is rejected here or, failing that, by the exact path -- a rejection is only ever a shortcut.
    if (!(pf_form(P, 0, x, y, z) + fmaf(P.pf_G[0], S, P.pf_H[0]) >= 0.f)) return true;  // z_cam < 0
    if (!(pf_form(P, 1, x, y, z) + fmaf(P.pf_G[1], S, P.pf_H[1]) >= 0.f)) return true;  // f*X + cx*Z < 0      <=> u < 0
    if (!(pf_form(P, 2, x, y, z) - fmaf(P.pf_G[2], S, P.pf_H[2]) <= 0.f)) return true;  // f*X + (cx-W)*Z > 0  <=> u > W
    if (!(pf_form(P, 3, x, y, z) + fmaf(P.pf_G[3], S, P.pf_H[3]) >= 0.f)) return true;  // f*Y + cy*Z < 0      <=> v < 0
    if (!(pf_form(P, 4, x, y, z) - fmaf(P.pf_G[4], S, P.pf_H[4]) <= 0.f)) return true;  // f*Y + (cy-H)*Z > 0  <=> v > H
    return false;
}

// pre-filter, exact projection and scatter of the K1_PPT points a thread holds in registers.
// Stage 1: the five pre-filter tests, one test at a time over all K1_PPT points (branch-free inside a test: the test's six
// constants stay in registers), leaving as soon as no point of the thread is alive -- the common case, since a thread's points
// span a few degrees of azimuth: half of a sweep leaves after the first test (z_cam < 0), the front-left / front-right sectors
// outside the image after the second / third. (One test after the other per point cost ~14 instructions per point and test,
// 4 of them constant loads: ncu r1c, 604 warp instructions per 8 points.)
// Stage 2: ONE copy of the exact FP64 path, looped over the surviving points (a set bit per point): eight unrolled copies
// cost 5.6 k instructions of I-cache and ~30 registers of hoisted FP64 constants (spills once the kernel carries a second role).
template <bool FULL_TILE>
__device__ __forceinline__ void scatter_points(const DevParams& P, const float4 (&p)[K1_PPT], int base, int n, unsigned int hi,
                                               unsigned int* __restrict__ map, unsigned int* __restrict__ ob, int occ_pitch) {
    unsigned int alive = 0u;
    {
        const float g0 = P.pf_g[0][0], g1 = P.pf_g[0][1], g2 = P.pf_g[0][2], h0 = P.pf_h[0], G0 = P.pf_G[0], H0 = P.pf_H[0];
#pragma unroll
        for (int j = 0; j < K1_PPT; j++) {
            const float S = fabsf(p[j].x) + fabsf(p[j].y) + fabsf(p[j].z);
            const float f = fmaf(g0, p[j].x, fmaf(g1, p[j].y, fmaf(g2, p[j].z, h0))) + fmaf(G0, S, H0);
            // NaN fails f >= 0 and is rejected here (never visible); points past the end of a ragged tile were loaded as zeros
            if (f >= 0.f && (FULL_TILE || base + j * K1_THREADS < n)) alive |= 1u << j;
        }
    }
    if (!alive) return;
#pragma unroll
    for (int t = 1; t < 5; t++) {
        // t = 1: f*X + cx*Z < 0 <=> u < 0;  2: f*X + (cx-W)*Z > 0 <=> u > W;  3: f*Y + cy*Z < 0 <=> v < 0;  4: f*Y + (cy-H)*Z > 0 <=> v > H
        const float g0 = P.pf_g[t][0], g1 = P.pf_g[t][1], g2 = P.pf_g[t][2], h0 = P.pf_h[t], G0 = P.pf_G[t], H0 = P.pf_H[t];
        unsigned int keep = 0u;
#pragma unroll
        for (int j = 0; j < K1_PPT; j++) {
            const float S = fabsf(p[j].x) + fabsf(p[j].y) + fabsf(p[j].z);
            const float f = fmaf(g0, p[j].x, fmaf(g1, p[j].y, fmaf(g2, p[j].z, h0)));
            const float e = fmaf(G0, S, H0);
            // a rejection is only ever a shortcut: an infinite coordinate that slips through fails the exact tests below
            const bool in = (t & 1) ? (f + e >= 0.f) : (f - e <= 0.f);
            if (in) keep |= 1u << j;
        }
        alive &= keep;
        if (!alive) return;
    }
#ifdef MLD_DIAG_NOFP64
    if (p[0].x == 123456.f) map[0] = alive;  // diagnostic build: keep the loads alive, skip the exact path
    return;
#endif
#pragma unroll 1
    while (alive) {
        const int j = __ffs((int)alive) - 1;
        alive &= alive - 1u;
        float x = p[0].x, y = p[0].y, z = p[0].z;
#pragma unroll
        for (int q = 1; q < K1_PPT; q++) {
            if (j == q) {
                x = p[q].x;
                y = p[q].y;
                z = p[q].z;
            }
        }
        int px, py;
        if (!project_pixel(P, x, y, z, true, px, py)) continue;
#ifdef MLD_DIAG_NOATOM
        if (px == -5) map[0] = 1;  // diagnostic build: no scatter
        continue;
#endif
        atomicMin(&map[py * P.W + px], hi | (unsigned int)(base + j * K1_THREADS));
        if (ob) atomicOr(ob + occ_word_of(occ_pitch, px, py), occ_bit_of(px, py));  // occ_pitch = tiles per image row
    }
}

// one K1 tile: K1_THREADS x K1_PPT consecutive points of one frame, starting at point tile * K1_THREADS * K1_PPT of the cloud at
// `cloud` (n points), scattered into `map` / `ob` (the frame's pixel map and occupancy bitmap). STRIDE_F = 4 (float4) or 8
// (pcl::PointXYZI) makes the eight load offsets immediates; a tile that lies completely inside the frame (all but the last)
// loads without per-point bounds predicates.
template <int STRIDE_F>
__device__ __forceinline__ void k1_tile_at(const DevParams& P, unsigned int hi, const float* __restrict__ cloud, int stride_rt, int n,
                                           unsigned int* __restrict__ map, unsigned int* __restrict__ ob, int tile) {
    const int stride_f = STRIDE_F > 0 ? STRIDE_F : stride_rt;
    const int occ_pitch = P.occ_tx;
    const int base = tile * (K1_THREADS * K1_PPT) + threadIdx.x;
    const float* src = cloud + (size_t)base * (size_t)stride_f;
    const int step = K1_THREADS * stride_f;  // floats between this thread's consecutive points

    float4 p[K1_PPT];
    if ((tile + 1) * (K1_THREADS * K1_PPT) <= n) {  // uniform per block
#pragma unroll
        for (int j = 0; j < K1_PPT; j++) p[j] = ld_stream_f4(src + j * step);
        scatter_points<true>(P, p, base, n, hi, map, ob, occ_pitch);
    } else {
#pragma unroll
        for (int j = 0; j < K1_PPT; j++) {
            if (base + j * K1_THREADS < n)
                p[j] = ld_stream_f4(src + j * step);
            else
                p[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        scatter_points<false>(P, p, base, n, hi, map, ob, occ_pitch);
    }
}

template <int STRIDE_F>
__device__ __forceinline__ void k1_tile_s(const DevParams& P, const MapCode& mc, const float* __restrict__ pts, int stride_rt, int n,
                                          long long pitch_pts, unsigned int* __restrict__ maps, unsigned int* __restrict__ occ,
                                          unsigned int frame, int tile) {
    const int stride_f = STRIDE_F > 0 ? STRIDE_F : stride_rt;
    unsigned int* map = maps + (size_t)frame * (size_t)P.map_cells;
    unsigned int* ob = occ ? occ + (size_t)frame * (size_t)P.occ_words : nullptr;
    const float* cloud = pts + (size_t)frame * (size_t)pitch_pts * (size_t)stride_f;
    const unsigned int hi = mc.tagged ? (mc.tag << MLD_TAG_SHIFT) : 0u;
    k1_tile_at<STRIDE_F>(P, hi, cloud, stride_rt, n, map, ob, tile);
}

__device__ __forceinline__ void k1_tile(const DevParams& P, const MapCode& mc, const float* __restrict__ pts, int stride_f, int n,
                                        long long pitch_pts, unsigned int* __restrict__ maps, unsigned int* __restrict__ occ,
                                        unsigned int frame, int tile) {
    if (stride_f == 4)
        k1_tile_s<4>(P, mc, pts, stride_f, n, pitch_pts, maps, occ, frame, tile);
    else
        k1_tile_s<0>(P, mc, pts, stride_f, n, pitch_pts, maps, occ, frame, tile);
}

}  // namespace
