// mld_pipeline.cu -- the batched hot path as ONE persistent kernel per sequence: projection + first-point-wins pixel map
// (K1) and the per-feature depth estimation (window gather, histogram segmentation, plane fit, ray intersection, thresholds,
// road path) run as two ROLES of the same grid, pipelined a few frames apart.
//
// Why. With one launch per stage and 512 frames per launch (the chunked pipelines of mld_capi.cu) the pixel maps of a chunk
// (0.96 GB) are written to DRAM by K1 and read back by the gather, the survivor coordinates make a DRAM round trip between
// gather and solve, and the solve kernel does not overlap the DRAM-bound launches (ncu r1f: 1.53x the algorithmic bytes in the
// fused K1 + gather kernel, 7.2x in the solve kernel, step == serialised sum). Here a frame's features are processed
// `delay` (~12) frames after its points were scattered: the map cells, occupancy words and points a window touches are
// still in the 126 MB L2, the survivors' coordinates never leave shared memory, and the latency-bound feature work shares
// every SM with the streaming role for the whole launch.
//
// Work distribution. Two in-order queues, K1 items (a few consecutive tiles of one frame) and feature blocks, each a global
// counter. A block that needs work looks at the head of both queues and CLAIMS (compare-and-swap on the counter) only an item
// whose dependency is already satisfied: a feature block of frame f once all K1 items of f have finished, a K1 item of frame f
// once the features of frame f - R (the previous user of its map slot, R = ring size) have released the slot. Ready feature
// blocks are preferred, so K1 runs only as far ahead of the features as it must (a few frames: everything a window touches is
// still in L2) and no resident block ever sits on an item it cannot run. Progress: the head of the feature queue depends on K1
// items that are already claimed by running blocks, the head of the K1 queue on feature blocks that are already claimed; claimed
// items never wait, so one of the two heads always becomes ready. The idle spin is bounded (error flag + exit) so that a bug
// cannot hang the GPU. (A single ticket queue with blocking waits was measured first: blocks parked on not-yet-ready items held
// 50-70 % of the resident slots.)
//
// Pixel maps live in a ring of R slots with the epoch-tagged encoding of mld_common.cuh (never cleared); the occupancy
// bitmap of a slot is cleared by the last feature block of the frame that used it. Map and occupancy are read with ld.cg
// (L2): a slot is rewritten during the launch, so L1 may hold stale lines. The point stream is loaded with an L2
// evict_first policy so that it does not push the ring out of L2.
//
// Reference routines restated by the roles: see mld_project.cu (K1) and mld_feature.cu / mld_thread_helpers.cuh (features).
#include <algorithm>

#include "mld_common.cuh"
#include "mld_geometry.cuh"
#include "mld_kernels.h"
#include "mld_thread_helpers.cuh"
#include "mld_project.cuh"
#include "mld_feature_warp.cuh"

namespace {

constexpr int PT = K1_THREADS;  // threads per block, both roles
constexpr int PSCAP = 9;        // neighbours a thread's slab holds (normal window: 3 rings x 3 returns); fuller windows take the warp path
constexpr int PRCAP = 24;       // neighbours of the road window (scale 2.0 x 1.5)
constexpr int PRB = 16;         // road survivors solved per batch
constexpr int K1_TILE_BYTES = K1_THREADS * K1_PPT * 16;  // one K1 tile of float4 points: 16 KB
constexpr int POOL_BYTES = (PSCAP * PT * (3 * 8 + 4) > 2 * K1_TILE_BYTES) ? PSCAP * PT * (3 * 8 + 4) : 2 * K1_TILE_BYTES;  // slabs x, y, z (double) + aux (int) of the feature role; two staging tiles of the K1 role
#ifndef MLD_PIPE_MINBLOCKS
#define MLD_PIPE_MINBLOCKS 6
#endif
static_assert(PT == 128, "the feature role is written for 4 warps");
static_assert(PRCAP * PT * 4 + PRB * PRCAP * (3 * 8 + 4) <= POOL_BYTES, "road phase must fit the pool");

struct PipeArgs {
    const float* pts;
    int stride_f;
    int n;
    long long pitch_pts;
    const double* uv;
    int F;
    double* depth;
    int* status;
    int nframes;
    unsigned int* maps;  // R slots of W x H cells
    unsigned int* occ;   // R slots of occ_words_per_frame words (zero on entry)
    int R;
    unsigned int epoch0;  // use g of a slot carries epoch epoch0 + g + 1
    int* sync;            // [0] K1 queue head, [1] error, [2] feature queue head, [8..24] profiling, then one 128-byte line per slot: {k1_done, feat_done, slot_free}; zero on entry
    const float* coeffs;  // per frame, lidar frame; nullptr: no road path
    const unsigned int* bits;
    long long words;
    int tiles;    // K1 tiles per frame
    int k1_group; // consecutive tiles of one K1 work item (bulk-copy pipeline inside the item)
    int k1_items; // K1 work items per frame = ceil(tiles / k1_group)
    int gblocks;  // feature blocks per frame
    int delay;    // K1 may run this many frames ahead of the feature queue
    int kcap;     // capacity of the warp path (96 / 256 / 1024)
    int hint;     // evict_first policy on the point stream
    int timing;   // accumulate clock64() per role / phase into sync[2..15] (profiling runs only)
    unsigned int total_k1;    // K1 items of the sequence
    unsigned int total_feat;  // feature blocks of the sequence
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- 1-D bulk async copies (TMA) global -> shared with mbarrier completion ------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long* bar, unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "MLD_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra MLD_DONE;\n"
        "bra MLD_WAIT;\n"
        "MLD_DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ unsigned long long l2_policy_evict_normal() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// Thread 0 claims the next runnable work item (see "Work distribution" above). Returns the item: bit 31 set = K1 item (low
// bits = index in the K1 queue), bit 31 clear = feature block (index in the feature queue), 0xFFFFFFFF = nothing left or the
// launch was aborted. sync[0] = K1 queue head, sync[1] = error flag, sync[2] = feature queue head.
constexpr unsigned int PIPE_DONE = 0xFFFFFFFFu, PIPE_K1 = 0x80000000u;
__device__ __forceinline__ int* sync_k1_done(int* sync, int slot) { return sync + 32 + slot * 32; }
__device__ __forceinline__ int* sync_feat_done(int* sync, int slot) { return sync + 32 + slot * 32 + 1; }
__device__ __forceinline__ int* sync_slot_free(int* sync, int slot) { return sync + 32 + slot * 32 + 2; }

// bounded spin of thread 0 until *p >= target; false when the launch was (or has to be) abandoned
__device__ __forceinline__ bool spin_until_ge(const PipeArgs& a, const int* p, int target) {
    unsigned int spins = 0;
    while (ld_acquire(p) < target) {
        if ((++spins & 15u) == 0u) {
            if (ld_acquire(a.sync + 1) != 0) return false;
            if (spins > (1u << 22)) {  // ~ a second of polling: abandon the launch instead of hanging the GPU
                atomicExch(a.sync + 1, 1);
                return false;
            }
        }
        __nanosleep(100);
    }
    return true;
}

__device__ __noinline__ unsigned int claim_item(const PipeArgs& a) {
    unsigned int* k1_head = reinterpret_cast<unsigned int*>(a.sync);
    unsigned int* feat_head = reinterpret_cast<unsigned int*>(a.sync + 2);
    unsigned int spins = 0;
    while (true) {
        const unsigned int fi = (unsigned int)ld_acquire(a.sync + 2), ki = (unsigned int)ld_acquire(a.sync);
        if (fi >= a.total_feat && ki >= a.total_k1) return PIPE_DONE;
        if (fi < a.total_feat) {
            const int f = (int)(fi / (unsigned)a.gblocks);
            const int slot = f % a.R, g = f / a.R;
            if (ld_acquire(sync_k1_done(a.sync, slot)) >= (g + 1) * a.k1_items) {
                // the head is ready: take a ticket (fetch-add: claims of different blocks do not serialise). Blocks that looked at
                // the same moment may push the ticket past the ready frames; such a ticket waits for its frame right here -- its
                // K1 items have lower tickets, all of them claimed or claimable by the blocks that are not idle like this one.
                const unsigned int t = atomicAdd(feat_head, 1u);
                if (t >= a.total_feat) continue;
                const int ft = (int)(t / (unsigned)a.gblocks);
                if (ft != f && !spin_until_ge(a, sync_k1_done(a.sync, ft % a.R), (ft / a.R + 1) * a.k1_items)) return PIPE_DONE;
                return t;
            }
        }
        if (ki < a.total_k1) {
            const int f = (int)(ki / (unsigned)a.k1_items);
            const int slot = f % a.R, g = f / a.R;
            // K1 runs at most `delay` frames ahead of the feature queue (L2 locality), and only into a free map slot
            const bool lead_ok = fi >= a.total_feat || f - (int)(fi / (unsigned)a.gblocks) < a.delay;
            if (lead_ok && (g == 0 || ld_acquire(sync_slot_free(a.sync, slot)) >= g)) {
                const unsigned int t = atomicAdd(k1_head, 1u);
                if (t >= a.total_k1) continue;
                const int ft = (int)(t / (unsigned)a.k1_items);
                if (ft != f && ft / a.R > 0 && !spin_until_ge(a, sync_slot_free(a.sync, ft % a.R), ft / a.R)) return PIPE_DONE;
                return PIPE_K1 | t;
            }
        }
        // nothing runnable right now: the items both heads depend on are running on other blocks
        if ((++spins & 15u) == 0u) {
            if (ld_acquire(a.sync + 1) != 0) return PIPE_DONE;
            if (spins > (1u << 22)) {
                atomicExch(a.sync + 1, 1);
                return PIPE_DONE;
            }
        }
        __nanosleep(100 + ((blockIdx.x * 37u) & 127u));  // jittered: idle blocks do not all look at the same instant
    }
}

// profiling counters (64-bit, sync[8..21] viewed as unsigned long long[7]): 0 K1 items, 1 looking for a runnable item (idle +
// claim latency), 2 unused, 3 feature phases A+B (window scan + gather), 4 phase C (solve), 5 road, 6 warp path; sync[24] = features
// on the warp path
__device__ __forceinline__ void tick(const PipeArgs& a, int which, long long& t0) {
    if (!a.timing) return;
    if (threadIdx.x == 0) {
        const long long t1 = clock64();
        atomicAdd(reinterpret_cast<unsigned long long*>(a.sync + 8) + which, (unsigned long long)(t1 - t0));
        t0 = t1;
    }
}

// ---- the feature role: 128 features of one frame ------------------------------------------------------------------
struct FeatCtx {
    const float* fp;            // the frame's cloud
    const unsigned int* map;    // its pixel map (ring slot)
    const unsigned int* occ;    // its occupancy bitmap
    MapCode mc;
    const float* coeffs;        // the frame's plane or nullptr
    const unsigned int* bits;
};

template <typename Release>
__device__ __forceinline__ void feature_block(const DevParams& P, const PipeArgs& a, const FeatCtx& c, long long frame, int bx,
                                              unsigned char* pool, Release release) {
    using TSlab = TSlabT<PSCAP, PT>;
    double* sx = reinterpret_cast<double*>(pool);
    double* sy = sx + PSCAP * PT;
    double* sz = sy + PSCAP * PT;
    int* saux = reinterpret_cast<int*>(sz + PSCAP * PT);  // pixel offsets by thread (gather), bin ids by rank (solve)
    __shared__ double s_resd[PT];
    __shared__ int s_hist[PSCAP + 1], s_start[PSCAP + 2], s_off[PSCAP + 1];
    __shared__ int s_wtot[PT / 32];
    __shared__ int s_novf, s_nroad;
    __shared__ short s_list[PT], s_cnt[PT];
    __shared__ unsigned char s_order[PT], s_k[PT], s_ovf[PT], s_road[PT];
    __shared__ signed char s_resst[PT], s_st[PT], s_ci[PT], s_cj[PT], s_ck[PT];

    const int tid = threadIdx.x;
    const int F = a.F;
    const int fi = bx * PT + tid;
    const bool valid = fi < F;
    const long long o = frame * (long long)F + fi;
    const double2* uv2 = reinterpret_cast<const double2*>(a.uv) + frame * (long long)F + (long long)bx * PT;  // this block's features

    if (P.set_all_zero) {  // DepthEstimator.cpp:448-453
        if (valid) {
            a.status[o] = 1;
            a.depth[o] = -1;
        }
        release();
        return;
    }
    if (tid == 0) {
        s_novf = 0;
        s_nroad = 0;
    }
    if (tid <= PSCAP) s_hist[tid] = 0;
    long long tt = a.timing ? clock64() : 0;

    // phase A: occupancy words -> pixel offsets of the window's points in the reference's scan order
    int k = 0;
    if (valid) {
        const double2 f2 = __ldg(uv2 + tid);
        const double u = f2.x, v = f2.y;
        if ((fabs(u) < 1e9) && (fabs(v) < 1e9)) {  // NaN / huge coordinates: empty window (see mld_feature_warp.cuh)
            const int x0 = (int)fmax(u - P.hx1, 0.), x1 = (int)fmin(u + P.hx1, (double)(P.W - 1));
            const int y0 = (int)fmax(v - P.hy1, 0.), y1 = (int)fmin(v + P.hy1, (double)(P.H - 1));
            if (x1 >= x0 && y1 >= y0) {
                occ_scan_window<true>(c.occ, P.W, x0, x1, y0, y1, [&](int off) {
                    if (k < PSCAP) saux[k * PT + tid] = off;
                    k++;
                });
            }
        }
    }
    const bool overflow = valid && k > PSCAP;
    const bool surv = valid && !overflow && (unsigned)k >= (unsigned)P.count_min;
    // result of this thread's feature; -1 = decided later (survivor or warp path)
    s_resst[tid] = (valid && !surv && !overflow) ? (signed char)ST_RadiusSearchInsufficientPoints : (signed char)-1;  // DepthEstimator.cpp:680
    s_resd[tid] = -1.0;
    __syncthreads();
    if (overflow) s_ovf[atomicAdd(&s_novf, 1)] = (unsigned char)tid;
    // survivors ordered by neighbour count (counting sort): the solve warps hold features of nearly equal k
    int r = 0;
    if (surv) r = atomicAdd(&s_hist[k], 1);
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int j = 0; j <= PSCAP; j++) {
            s_start[j] = acc;
            acc += s_hist[j];
        }
        s_start[PSCAP + 1] = acc;
        int pairs = 0;
        for (int i = 0; i < PSCAP; i++) {
            s_off[i] = pairs;
            pairs += acc - s_start[i + 1];  // survivors with k > i own an entry i
        }
        s_off[PSCAP] = pairs;
    }
    __syncthreads();
    const int S = s_start[PSCAP + 1];
    if (surv) {
        const int rank = s_start[k] + r;
        s_order[rank] = (unsigned char)tid;
        s_k[rank] = (unsigned char)k;
    }
    __syncthreads();
    // phase B over the flattened (entry, survivor rank) pairs, entry-major: map cell -> raw index -> point -> FP64 camera frame
    // into the survivor's slab; two pairs per thread in flight
    {
        const int T = s_off[PSCAP];
        for (int p0 = tid; p0 < T; p0 += 2 * PT) {
            const unsigned int* cellp[2];
            int dsti[2];
            bool ok[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int p = p0 + q * PT;
                ok[q] = p < T;
                int i = 0;
#pragma unroll
                for (int j = 1; j < PSCAP; j++) i += (p >= s_off[j]) ? 1 : 0;
                const int rank = ok[q] ? s_start[i + 1] + (p - s_off[i]) : 0;
                cellp[q] = c.map + saux[i * PT + s_order[rank]];
                dsti[q] = i * PT + rank;
            }
            unsigned int raw[2];
#pragma unroll
            for (int q = 0; q < 2; q++) raw[q] = ok[q] ? map_cell_index(c.mc, __ldcg(cellp[q])) : 0u;
            float4 pt[2];
#pragma unroll
            for (int q = 0; q < 2; q++)
                pt[q] = ok[q] ? __ldg(reinterpret_cast<const float4*>(c.fp + (long long)raw[q] * a.stride_f)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (!ok[q]) continue;
                const D3 cc = lidar_to_cam(P, pt[q].x, pt[q].y, pt[q].z);
                sx[dsti[q]] = cc.x;
                sy[dsti[q]] = cc.y;
                sz[dsti[q]] = cc.z;
            }
        }
    }
    __syncthreads();  // slabs complete; the pixel offsets in saux are dead from here on
    tick(a, 3, tt);
    // The slot (map, occupancy) is needed again only by the road path's wide window and by the warp path; without them it is
    // handed back right here, before the long solve phase: K1 of frame + R waits on this, and its latency bounds the pipeline.
    const bool need_slot_later = (c.coeffs != nullptr && P.road_mode != ROAD_NONE) || s_novf > 0;  // uniform per block
    if (!need_slot_later) release();

    // phase C: histogram segmentation + corner selection, one thread per survivor rank
    auto slab_of = [&](int rank) { return TSlab{sx + rank, sy + rank, sz + rank, saux + rank}; };
    bool stage2 = false;
    s_st[tid] = ST_Unspecified;
    if (tid < S) {
        const TSlab s = slab_of(tid);
        int n = (int)s_k[tid];
        int st = ST_Unspecified;
        if (P.use_hist) {
            n = t_histogram_segment(P, n, s);
            if (n < 0) st = ST_HistogramNoLocalMax;
        }
        if (st != ST_HistogramNoLocalMax) {
            int ci, cj, ck;
            st = t_select_corners(P, n, s, ci, cj, ck);
            if (st == 0) {
                s_cnt[tid] = (short)n;
                s_ci[tid] = (signed char)ci;
                s_cj[tid] = (signed char)cj;
                s_ck[tid] = (signed char)ck;
                stage2 = true;
            }
        }
        if (!stage2) s_st[tid] = (signed char)st;
    }
    __syncthreads();
    const int n2 = block_compact<PT>(stage2, tid, s_list, s_wtot);
    // geometry tail on dense lanes
    if (tid < n2) {
        const int rank = s_list[tid];
        const double2 f2 = __ldg(uv2 + s_order[rank]);
        double dp;
        const int st = t_depth_from_corners(P, f2.x, f2.y, (int)s_cnt[rank], slab_of(rank), (int)s_ci[rank], (int)s_cj[rank],
                                            (int)s_ck[rank], dp);
        s_st[rank] = (signed char)st;
        if (st == ST_Success) s_resd[s_order[rank]] = dp;
    }
    __syncthreads();
    if (tid < S) s_resst[s_order[tid]] = s_st[tid];
    __syncthreads();

    tick(a, 4, tt);
    // road path (DepthEstimator.cpp:579-597): survivors of the neighbour search that did not succeed
    if (c.coeffs != nullptr && P.road_mode != ROAD_NONE) {
        using RSlab = TSlabT<PRCAP, PRB>;
        int* raux = reinterpret_cast<int*>(pool);                               // [PRCAP][PT] pixel offsets / raw indices by candidate
        double* rx = reinterpret_cast<double*>(pool + PRCAP * PT * 4);          // [PRCAP][PRB] road slabs
        double* ry = rx + PRCAP * PRB;
        double* rz = ry + PRCAP * PRB;
        int* rax = reinterpret_cast<int*>(rz + PRCAP * PRB);
        __shared__ unsigned char s_rs[PT];      // road survivors: candidate index
        __shared__ unsigned char s_rn[PT];      // their inlier counts
        __shared__ unsigned int s_rmask[PT];    // their inlier masks
        __shared__ int s_nrs;
        const bool cand = valid && surv && s_resst[tid] != ST_Success;
        if (cand) s_road[atomicAdd(&s_nroad, 1)] = (unsigned char)tid;
        if (tid == 0) s_nrs = 0;
        __syncthreads();
        const int ncand = s_nroad;
        if (tid < ncand) {
            const int owner = s_road[tid];
            const double2 f2 = __ldg(uv2 + owner);
            const double u = f2.x, v = f2.y;
            int k2 = 0;
            if ((fabs(u) < 1e9) && (fabs(v) < 1e9)) {
                const int x0 = (int)fmax(u - P.hx2, 0.), x1 = (int)fmin(u + P.hx2, (double)(P.W - 1));
                const int y0 = (int)fmax(v - P.hy2, 0.), y1 = (int)fmin(v + P.hy2, (double)(P.H - 1));
                if (x1 >= x0 && y1 >= y0) {
                    occ_scan_window<true>(c.occ, P.W, x0, x1, y0, y1, [&](int off) {
                        if (k2 < PRCAP) raux[k2 * PT + tid] = off;
                        k2++;
                    });
                }
            }
            if (k2 > PRCAP) {
                s_ovf[atomicAdd(&s_novf, 1)] = (unsigned char)owner;  // the warp path redoes the feature from scratch
                s_resst[owner] = -1;
            } else if ((unsigned)k2 < (unsigned)P.count_min) {  // DepthEstimator.cpp:585-586
                s_resst[owner] = ST_RadiusSearchInsufficientPoints;
            } else {
                const float pa = c.coeffs[0], pb = c.coeffs[1], pc = c.coeffs[2], pd = c.coeffs[3];
#pragma unroll 4
                for (int i = 0; i < k2; i++) raux[i * PT + tid] = (int)map_cell_index(c.mc, __ldcg(c.map + raux[i * PT + tid]));
                bool far = false;
                int n_inl = 0;
                unsigned int inl_mask = 0u;
                for (int i = 0; i < k2 && !far; i++) {
                    const int raw = raux[i * PT + tid];
                    const float4 q = __ldg(reinterpret_cast<const float4*>(c.fp + (long long)raw * a.stride_f));
                    far = road_point_far(P, lidar_to_cam(P, q.x, q.y, q.z), pa, pb, pc, pd);
                    if ((c.bits[raw >> 5] >> (raw & 31)) & 1u) {
                        inl_mask |= 1u << i;
                        n_inl++;
                    }
                }
                // a far neighbour or fewer than 3 inliers: the normal path's status stands (DepthEstimator.cpp:589-591)
                if (!far && n_inl >= 3) {
                    const int e = atomicAdd(&s_nrs, 1);
                    s_rs[e] = (unsigned char)tid;
                    s_rn[e] = (unsigned char)n_inl;
                    s_rmask[e] = inl_mask;
                }
            }
        }
        __syncthreads();
        const int nrs = s_nrs;
        for (int b0 = 0; b0 < nrs; b0 += PRB) {  // road estimator on batches of PRB survivors (their inlier points in slabs)
            if (tid < PRB && b0 + tid < nrs) {
                const int e = b0 + tid;
                const int ct = s_rs[e];
                const int owner = s_road[ct];
                const RSlab s{rx + tid, ry + tid, rz + tid, rax + tid};
                unsigned int m = s_rmask[e];
                int w = 0;
                for (int i = 0; m; i++, m >>= 1) {
                    if (!(m & 1u)) continue;
                    const float4 q = __ldg(reinterpret_cast<const float4*>(c.fp + (long long)raux[i * PT + ct] * a.stride_f));
                    s.set(w++, lidar_to_cam(P, q.x, q.y, q.z));
                }
                const double2 f2 = __ldg(uv2 + owner);
                double dp;
                const int st = t_road_estimate(P, f2.x, f2.y, (int)s_rn[e], s, c.coeffs, dp);
                s_resst[owner] = (signed char)st;
                s_resd[owner] = (st == ST_SuccessRoad) ? dp : -1.0;
            }
            __syncthreads();
        }
    }

    // windows with more points than a slab holds: one warp per feature, from scratch (mld_feature_warp.cuh)
    __syncthreads();
    tick(a, 5, tt);
    const int novf = s_novf;
    if (a.timing && tid == 0 && novf) atomicAdd(a.sync + 24, novf);
    if (novf > 0) {  // uniform per block
        const int warp = tid >> 5, lane = tid & 31;
        const int slab_bytes = a.kcap * (3 * 8 + 4);
        const int nw = min(PT / 32, POOL_BYTES / slab_bytes);  // warps that fit the pool side by side (kcap 1024: one)
        if (warp < nw) {
            double* wx = reinterpret_cast<double*>(pool + (size_t)warp * slab_bytes);
            const WarpSlab ws{wx, wx + a.kcap, wx + 2 * a.kcap, reinterpret_cast<int*>(wx + 3 * a.kcap)};
            for (int i = warp; i < novf; i += nw) {
                const int owner = s_ovf[i];
                const double2 f2 = __ldg(uv2 + owner);
                int st;
                double dp;
                feature_depth(P, c.mc, c.map, c.fp, a.stride_f, f2.x, f2.y, c.coeffs, c.bits, lane, ws, a.kcap, st, dp);
                if (lane == 0) {
                    s_resst[owner] = (signed char)st;
                    s_resd[owner] = (st == ST_Success || st == ST_SuccessRoad) ? dp : -1.0;
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }
    tick(a, 6, tt);
    if (need_slot_later) release();
    if (valid) {  // one coalesced write of every feature's result
        a.status[o] = (int)s_resst[tid];
        a.depth[o] = s_resd[tid];
    }
}

// ---- the persistent kernel ----------------------------------------------------------------------------------------------
template <int STRIDE_F>
__global__ void __launch_bounds__(PT, MLD_PIPE_MINBLOCKS) depth_pipeline_kernel(const __grid_constant__ DevParams P, const __grid_constant__ PipeArgs a) {
    __shared__ __align__(16) unsigned char pool[POOL_BYTES];
    __shared__ unsigned int s_tk;
    __shared__ int s_flag;
    __shared__ __align__(8) unsigned long long s_bar[2];  // "tile landed" barriers of the two staging buffers
    const int tid = threadIdx.x;
    unsigned int bar_parity = 0u;  // bit b = parity the next wait on s_bar[b] expects (uniform across the block)
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // per-slot counters on their own 128-byte line (pollers of one slot do not slow the atomics of another)
    auto k1_done = [&](int slot) { return sync_k1_done(a.sync, slot); };
    auto feat_done = [&](int slot) { return sync_feat_done(a.sync, slot); };
    auto slot_free = [&](int slot) { return sync_slot_free(a.sync, slot); };
    const unsigned long long pol = a.hint ? l2_policy_evict_first() : l2_policy_evict_normal();
    const size_t WH = (size_t)P.W * (size_t)P.H;
    const size_t OW = (size_t)occ_words_per_frame(P.W, P.H);

    while (true) {
        long long tt = a.timing ? clock64() : 0;
        if (tid == 0) s_tk = claim_item(a);
        __syncthreads();
        const unsigned int t = s_tk;
        if (t == PIPE_DONE) break;
        tick(a, 1, tt);  // time spent looking for a runnable item (idle + claim latency)
        if (t & PIPE_K1) {
            // ---- K1: tiles [i * k1_group, ...) of frame r; its map slot is free (checked by claim_item) ----
            const unsigned int ki = t & ~PIPE_K1;
            const int r = (int)(ki / (unsigned)a.k1_items);
            const int i = (int)(ki - (unsigned)r * (unsigned)a.k1_items);
            {
                const int slot = r % a.R, g = r / a.R;
                {
                    const unsigned int hi = (MLD_TAG_MAX_EPOCH - (a.epoch0 + (unsigned)g + 1u)) << MLD_TAG_SHIFT;
                    const float* cloud = a.pts + (size_t)r * (size_t)a.pitch_pts * (size_t)a.stride_f;
                    unsigned int* map = a.maps + slot * WH;
                    unsigned int* ob = a.occ + slot * OW;
                    const int t0 = i * a.k1_group, t1 = min(a.tiles, t0 + a.k1_group);
                    if (STRIDE_F == 4) {
                        // float4 clouds: the item's tiles stream through two 16 KB staging buffers in shared memory, filled by 1-D
                        // bulk async copies (one elected thread issues, the hardware copies): the next tile is in flight while this
                        // one is filtered, independent of how many K1 blocks happen to be resident.
                        const int occ_pitch = occ_tiles_x(P.W);
                        auto issue = [&](int tile, int b) {  // thread 0 only
                            const int first = tile * (K1_THREADS * K1_PPT);
                            const unsigned int bytes = (unsigned int)min(K1_THREADS * K1_PPT, a.n - first) * 16u;
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic accesses to the buffer are done
                            mbar_expect_tx(&s_bar[b], bytes);
                            bulk_g2s(pool + b * K1_TILE_BYTES, cloud + (size_t)first * 4, bytes, &s_bar[b], pol);
                        };
                        if (tid == 0) {
                            issue(t0, 0);
                            if (t0 + 1 < t1) issue(t0 + 1, 1);
                        }
                        for (int tile = t0; tile < t1; tile++) {
                            const int b = (tile - t0) & 1;
                            mbar_wait(&s_bar[b], (bar_parity >> b) & 1u);
                            bar_parity ^= 1u << b;
                            const int base = tile * (K1_THREADS * K1_PPT) + tid;
                            const float4* src = reinterpret_cast<const float4*>(pool + b * K1_TILE_BYTES) + tid;
                            float4 p[K1_PPT];
                            const bool full = (tile + 1) * (K1_THREADS * K1_PPT) <= a.n;
#pragma unroll
                            for (int j = 0; j < K1_PPT; j++)
                                p[j] = (full || base + j * K1_THREADS < a.n) ? src[j * K1_THREADS] : make_float4(0.f, 0.f, 0.f, 0.f);
                            __syncthreads();  // every thread holds its points: the buffer can be refilled
                            if (tid == 0 && tile + 2 < t1) issue(tile + 2, b);
                            if (full)
                                scatter_points<true>(P, p, base, a.n, hi, map, ob, occ_pitch);
                            else
                                scatter_points<false>(P, p, base, a.n, hi, map, ob, occ_pitch);
                        }
                    } else {
                        for (int tile = t0; tile < t1; tile++) {
                            if (a.hint)
                                k1_tile_at<STRIDE_F, true>(P, hi, cloud, a.stride_f, a.n, map, ob, tile, pol);
                            else
                                k1_tile_at<STRIDE_F, false>(P, hi, cloud, a.stride_f, a.n, map, ob, tile, pol);
                        }
                    }
                    __threadfence();  // this thread's map / occupancy atomics are visible before the item is counted
                    __syncthreads();
                    if (tid == 0) {
                        __threadfence();
                        atomicAdd(k1_done(slot), 1);
                    }
                    tick(a, 0, tt);
                }
            }
        } else {
            // ---- features: block i of frame f; all K1 items of the frame have finished (checked by claim_item) ----
            const int f = (int)(t / (unsigned)a.gblocks);
            const int i = (int)(t - (unsigned)f * (unsigned)a.gblocks);
            {
                const int slot = f % a.R, g = f / a.R;
                {
                    FeatCtx c;
                    c.fp = a.pts + (size_t)f * (size_t)a.pitch_pts * (size_t)a.stride_f;
                    c.map = a.maps + slot * WH;
                    c.occ = a.occ + slot * OW;
                    c.mc = MapCode{1u, MLD_TAG_MAX_EPOCH - (a.epoch0 + (unsigned)g + 1u)};
                    c.coeffs = a.coeffs ? a.coeffs + (size_t)f * 4 : nullptr;
                    c.bits = a.bits ? a.bits + (size_t)f * (size_t)a.words : nullptr;
                    // release(): every thread of the block has finished reading the slot's map / occupancy. The last block of the
                    // frame hands the slot back to K1 (of frame f + R) with a clean occupancy bitmap.
                    auto release = [&]() {
                        __syncthreads();
                        if (tid == 0) {
                            __threadfence();
                            s_flag = (atomicAdd(feat_done(slot), 1) == (g + 1) * a.gblocks - 1) ? 1 : 0;
                            __threadfence();  // the other blocks' reads of the slot precede the clear below
                        }
                        __syncthreads();
                        if (s_flag != 0) {
                            uint4* ob = reinterpret_cast<uint4*>(a.occ + slot * OW);
                            for (size_t q = tid; q < OW / 4; q += PT) __stcg(ob + q, make_uint4(0u, 0u, 0u, 0u));
                            __threadfence();
                            __syncthreads();
                            if (tid == 0) {
                                __threadfence();
                                atomicExch(slot_free(slot), g + 1);
                            }
                        }
                    };
                    feature_block(P, a, c, (long long)f, i, pool, release);
                }
            }
        }
        __syncthreads();  // s_tk is rewritten by the next claim
    }
}

}  // namespace

// ring geometry of the persistent pipeline for a given image: R slots; bytes of maps / occupancy / sync words
int mld_pipeline_ring_slots(void) { return 128; }
size_t mld_pipeline_sync_bytes(int R) { return (size_t)(32 + 32 * R) * sizeof(int); }

// One persistent launch over nframes device-resident frames. d_sync: mld_pipeline_sync_bytes(R); d_occ_ring is zeroed here, d_sync
// too. epoch0: uses of the ring's slots so far (tagged maps; the caller clears the ring when the 14-bit epoch space runs out).
cudaError_t mld_launch_depth_pipeline(const DevParams& P, const float* d_pts, int stride_f, long long n_points, long long pitch_pts,
                                      const double* d_uv, int F, double* d_depth, int* d_status, long long nframes,
                                      unsigned int* d_map_ring, unsigned int* d_occ_ring, int R, unsigned int epoch0, int* d_sync,
                                      const float* d_plane_coeffs, const unsigned int* d_inlier_bits, long long words_per_frame, int kcap,
                                      int delay, int hint, int timing, int k1_group, int grid_blocks, cudaStream_t stream, int* launches) {
    if (nframes <= 0 || F <= 0 || n_points <= 0) return cudaSuccess;
    if (n_points > (long long)(MLD_TAG_IDX_MASK + 1u) || R < 2) return cudaErrorInvalidValue;
    PipeArgs a{};
    a.pts = d_pts; a.stride_f = stride_f; a.n = (int)n_points; a.pitch_pts = pitch_pts;
    a.uv = d_uv; a.F = F; a.depth = d_depth; a.status = d_status; a.nframes = (int)nframes;
    a.maps = d_map_ring; a.occ = d_occ_ring; a.R = R; a.epoch0 = epoch0; a.sync = d_sync;
    a.coeffs = d_plane_coeffs; a.bits = d_inlier_bits; a.words = words_per_frame;
    a.tiles = (int)((n_points + K1_THREADS * K1_PPT - 1) / (K1_THREADS * K1_PPT));
    a.k1_group = std::max(1, std::min(k1_group, a.tiles));
    a.k1_items = (a.tiles + a.k1_group - 1) / a.k1_group;
    a.gblocks = (F + PT - 1) / PT;
    a.delay = std::max(1, std::min(delay, R));
    a.kcap = kcap;
    a.hint = hint;
    a.timing = timing;
    const long long total = nframes * (long long)(a.k1_items + a.gblocks);
    if (nframes * (long long)std::max(a.k1_items, a.gblocks) >= 0x7fffff00LL) return cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(d_pts) & 15u) != 0 || (stride_f == 4 && (pitch_pts * 16) % 16 != 0)) return cudaErrorInvalidValue;
    a.total_k1 = (unsigned int)(nframes * a.k1_items);
    a.total_feat = (unsigned int)(nframes * a.gblocks);
    cudaError_t e = cudaMemsetAsync(d_sync, 0, mld_pipeline_sync_bytes(R), stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(d_occ_ring, 0, (size_t)R * (size_t)occ_words_per_frame(P.W, P.H) * sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(grid_blocks, total));
    if (stride_f == 4)
        depth_pipeline_kernel<4><<<grid, PT, 0, stream>>>(P, a);
    else
        depth_pipeline_kernel<0><<<grid, PT, 0, stream>>>(P, a);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

// resident blocks per SM of the pipeline kernel (grid = this x SM count)
int mld_pipeline_blocks_per_sm(void) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, depth_pipeline_kernel<4>, PT, 0) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}
