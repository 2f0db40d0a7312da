// mld_feature_split.cu -- K2 as two (three with a ground plane) thread-per-feature kernels with a
// chunk-wide compaction in between.
//
// ncu of the fused thread-per-feature kernel (round-1 captures r1_v4, git history): 45 % of the stall samples wait on the
// three dependent load rounds of the window gather, 23 % on the block barriers of the in-block
// compaction, occupancy pinned at 16 warps/SM by the per-thread XYZ slabs that only the later phases
// need, and 44 % of the features (empty windows) idle through those phases. Splitting fixes all three:
//
//   K2a gather  every feature: occupancy words -> pixel offsets -> map cells -> raw point indices. No XYZ
//               slab (8 KB of shared memory per block), so it runs at register-limited occupancy where
//               the load latency is hidden. Features with k >= radiusSearch_count_min are appended to a
//               chunk-wide survivor list (one atomic per block) with the raw indices of their neighbours
//               as [entry][slot]; empty windows get status 2 right here. (Round 1 handed the FP64
//               camera-frame points over instead: 24 bytes per neighbour written by the DRAM-bound fused
//               launch and read back by the solve -- 7.2x the solve's algorithmic bytes, ncu r1f.)
//   K2b solve   one thread per SURVIVOR (dense warps over the whole chunk): raw indices -> points ->
//               FP64 camera frame (bit-identical to K1), histogram segmentation, corner selection,
//               in-block compaction, geometry tail. With a plane, unsolved features go to a second list.
//   K2c road    one thread per road candidate: wide-window gather + plane gate + road estimator.
//
// Reference routines restated: see mld_feature.cu / mld_thread_helpers.cuh (DepthEstimator.cpp:491-600).
#include "mld_common.cuh"
#include "mld_geometry.cuh"
#include "mld_kernels.h"
#include "mld_thread_helpers.cuh"
#include "mld_project.cuh"

namespace {

// Entries of the solve kernel's per-thread slab (28 bytes of shared memory each). Round 1 (uniform ring spacing: 2 rings x 3-4
// returns per 7 x 10 window): 8 entries, 31 KB per block, 6-7 blocks per SM (12 entries: 4 blocks, 1.22 instead of 1.31 M frames/s).
// Round 2 (HDL-64E ring layout: the upper block's rings are 4 pixels apart, 3 rings x 3 returns fit a window): 9 entries, 35 KB,
// 6 blocks per SM; with 8 entries 3.5 % of the features took the warp-per-feature overflow pass (0.56 ms per 512 frames).
#ifndef MLD_SCAP
#define MLD_SCAP 9  // 3 rings x 3 returns fit (round 2 scene: HDL-64E ring layout); 8 sent 3.5 % of the features to the overflow pass
#endif
#ifndef MLD_SOLVE_MINBLOCKS
#define MLD_SOLVE_MINBLOCKS 5  // 5 blocks per SM let the kernel keep its FP64 state in registers (6: capped at 80, 60 + 160 bytes of spills per thread): 1.41 -> 1.47 M frames/s; 4 and 3 measure the same
#endif
#ifndef MLD_SBT_B
#define MLD_SBT_B 128
#endif
#ifndef MLD_GATHER_ILP
#define MLD_GATHER_ILP 2
#endif
#ifndef MLD_GCAP
#define MLD_GCAP 16
#endif
constexpr int GILP = MLD_GATHER_ILP;  // (entry, survivor) pairs a gather thread keeps in flight
constexpr int SCAP = MLD_SCAP;  // neighbours per feature the main solve kernel's slabs hold
// neighbours per feature the gather lists hold (more -> warp-kernel overflow list). Windows of SCAP + 1 .. GCAP points (one
// feature in 2000 on the KITTI shape) are solved by a second, small instantiation of the solve kernel with bigger slabs: the
// warp-per-feature overflow pass they took before cost 7 % of the step (126 blocks x 256 threads x 25 us per 512 frames beside
// the fused launch; measured by leaving the pass out).
constexpr int GCAP = MLD_GCAP > MLD_SCAP ? MLD_GCAP : MLD_SCAP;
constexpr int SBT_A = 128;  // threads per block, gather
constexpr int SBT_B = MLD_SBT_B;  // threads per block, solve
constexpr int SBT_B2 = 64;  // threads per block, solve of the classes SCAP + 1 .. GCAP
constexpr int RCAP = 24;    // neighbours per feature in the road window
constexpr int SBT_C = 64;   // threads per block, road

// road-survivor record: (k << 27) | global feature id
__device__ __forceinline__ unsigned int pack_rec(int k, long long o) { return ((unsigned int)k << 27) | (unsigned int)o; }
// survivor class k (= neighbour count, 0..GCAP): records at surv_rec[k][slot], raw neighbour indices at
// surv_idx[class_row(k) + entry][slot]; rows of all classes: class_row(GCAP + 1)
__host__ __device__ constexpr int class_row(int k) { return k * (k - 1) / 2; }
constexpr int CLASS_COUNT_AT = 8;  // class counters: ints 8 .. 8 + GCAP of the scratch header (0..2: road / road-survivor counters)
static_assert(CLASS_COUNT_AT + GCAP + 1 <= 32, "counters live in the 128-byte header");

// ---- K2a ------------------------------------------------------------------------------------------
// one block of SBT_A features of `frame` (bx = block index inside the frame)
__device__ __forceinline__ void gather_block(const DevParams& P, const MapCode& mc, const float* __restrict__ pts, int stride_f,
                                             long long pitch_pts, const unsigned int* __restrict__ maps,
                                             const unsigned int* __restrict__ occs, const double* __restrict__ uv, int F,
                                             double* __restrict__ depth, int* __restrict__ status, int* __restrict__ overflow_list,
                                             int* __restrict__ overflow_count, unsigned int* __restrict__ surv_rec,
                                             unsigned int* __restrict__ surv_idx, int* __restrict__ class_count, long long cap, int bx,
                                             long long frame) {
    __shared__ int s_aux[GCAP * SBT_A];
    __shared__ int s_hist[GCAP + 1], s_start[GCAP + 2], s_off[GCAP + 1];
    __shared__ unsigned char s_order[SBT_A], s_kof[SBT_A];
    __shared__ int s_cbase[GCAP + 1];
    static_assert(SBT_A <= 256, "s_order holds thread ids in a byte");
    const int tid = threadIdx.x;
    const int fi = bx * SBT_A + tid;
    const bool valid = fi < F;
    const long long o = frame * (long long)F + fi;
    const unsigned int* map = maps + frame * (long long)P.map_cells;
    const unsigned int* occ = occs + frame * (long long)P.occ_words;
    int* aux = s_aux + tid;

    if (P.set_all_zero) {  // DepthEstimator.cpp:448-453
        if (valid) {
            status[o] = 1;
            depth[o] = -1;
        }
        return;
    }

    // phase 1: occupancy words -> pixel offsets of the window's points in scan order
    int k = 0;
    if (valid) {
        const double2 f2 = __ldg(reinterpret_cast<const double2*>(uv) + o);
        const double u = f2.x, v = f2.y;
        if ((fabs(u) < 1e9) && (fabs(v) < 1e9)) {  // NaN / huge coordinates: empty window (see mld_feature.cu)
            const int x0 = (int)fmax(u - P.hx1, 0.), x1 = (int)fmin(u + P.hx1, (double)(P.W - 1));
            const int y0 = (int)fmax(v - P.hy1, 0.), y1 = (int)fmin(v + P.hy1, (double)(P.H - 1));
            if (x1 >= x0 && y1 >= y0) {
                occ_scan_window(occ, P.W, x0, x1, y0, y1, [&](int off) {
                    if (k < GCAP) aux[k * SBT_A] = off;
                    k++;
                });
            }
        }
    }
    const bool overflow = valid && k > GCAP;
    const bool surv = valid && !overflow && (unsigned)k >= (unsigned)P.count_min;
    if (valid && !surv) {
        if (overflow) {
            overflow_list[atomicAdd(overflow_count, 1)] = (int)o;  // finished by the warp-per-feature kernel
        } else {  // neighbors.size() < (uint)radiusSearch_count_min (DepthEstimator.cpp:680)
            status[o] = ST_RadiusSearchInsufficientPoints;
            depth[o] = -1;
        }
    }
    // chunk-wide survivor lists, ONE PER NEIGHBOUR COUNT k (class): the solve kernel's blocks then hold survivors of equal k, so
    // its per-thread loops over the neighbours do not diverge (ncu r2q with one list sorted by k inside each gather block only:
    // 15 of 32 lanes active in the histogram, 10 in the triangle search -- a block's ~50 survivors spread over every k). The
    // block's survivors are counting-sorted by k in shared memory; one global atomic per class reserves their slots.
    if (tid <= GCAP) s_hist[tid] = 0;
    __syncthreads();
    int r = 0;
    if (surv) r = atomicAdd(&s_hist[k], 1);  // rank among the block's survivors with the same k
    __syncthreads();
    if (tid == 0) {
        // s_start[j] = survivors with k < j; s_off[i] = (entry, survivor) pairs with entry < i
        int acc = 0;
        for (int j = 0; j <= GCAP; j++) {
            s_start[j] = acc;
            acc += s_hist[j];
        }
        s_start[GCAP + 1] = acc;
        int pairs = 0;
        for (int i = 0; i < GCAP; i++) {
            s_off[i] = pairs;
            pairs += acc - s_start[i + 1];  // survivors with k > i own an entry i
        }
        s_off[GCAP] = pairs;
    }
    if (tid <= GCAP && s_hist[tid] > 0) s_cbase[tid] = atomicAdd(class_count + tid, s_hist[tid]);  // first slot of the block in class tid
    __syncthreads();
    const int S = s_start[GCAP + 1];
    if (S == 0) return;  // uniform per block
    if (surv) {
        const int rank = s_start[k] + r;
        s_order[rank] = (unsigned char)tid;
        s_kof[rank] = (unsigned char)k;
        surv_rec[(long long)k * cap + (s_cbase[k] + r)] = (unsigned int)o;
    }
    __syncthreads();
    // phase 2 over the flattened (entry i, survivor rank) pairs, entry-major: every lane owns one neighbour (dense warps
    // whatever the spread of k), lanes of equal i store to consecutive slots. map cell -> raw index, stored [entry][slot]
    const int T = s_off[GCAP];
    // GILP pairs per thread and pass: the map loads of all of them are issued before the first store
    for (int p0 = tid; p0 < T; p0 += GILP * SBT_A) {
        long long dsti[GILP];
        const unsigned int* cellp[GILP];
        bool ok[GILP];
#pragma unroll
        for (int u = 0; u < GILP; u++) {
            const int p = p0 + u * SBT_A;
            ok[u] = p < T;
            int i = 0;
#pragma unroll
            for (int j = 1; j < GCAP; j++) i += (p >= s_off[j]) ? 1 : 0;
            const int rank = ok[u] ? s_start[i + 1] + (p - s_off[i]) : 0;
            const int owner = s_order[rank];
            const int kc = s_kof[rank];  // class k owns the index rows tri(k) .. tri(k) + k - 1 of [row][slot]
            cellp[u] = map + s_aux[i * SBT_A + owner];
            dsti[u] = (long long)(class_row(kc) + i) * cap + (s_cbase[kc] + (rank - s_start[kc]));
        }
        unsigned int raw[GILP];
#pragma unroll
        for (int u = 0; u < GILP; u++) raw[u] = ok[u] ? map_cell_index(mc, __ldg(cellp[u])) : 0u;
#pragma unroll
        for (int u = 0; u < GILP; u++)
            if (ok[u]) surv_idx[dsti[u]] = raw[u];
    }
}

__global__ void __launch_bounds__(SBT_A)
feature_gather_kernel(DevParams P, MapCode mc, const float* __restrict__ pts, int stride_f, long long pitch_pts,
                      const unsigned int* __restrict__ maps, const unsigned int* __restrict__ occs,
                      const double* __restrict__ uv, int F, double* __restrict__ depth, int* __restrict__ status,
                      int* __restrict__ overflow_list, int* __restrict__ overflow_count, unsigned int* __restrict__ surv_rec,
                      unsigned int* __restrict__ surv_idx, int* __restrict__ class_count, long long cap) {
    gather_block(P, mc, pts, stride_f, pitch_pts, maps, occs, uv, F, depth, status, overflow_list, overflow_count, surv_rec, surv_idx,
                 class_count, cap, (int)blockIdx.x, (long long)blockIdx.y);
}

// ---- K1 of one chunk and K2a of the previous chunk in ONE launch ------------------------------------------------
// Both are latency bound and independent of each other (different map slots); on separate streams the hardware runs
// them mostly back to back because K1's grid fills every SM. Here a frame's gather blocks follow its K1 tiles in launch order,
// so the two kinds are co-resident in a fixed ratio for the whole launch and cover each other's memory stalls.
struct FusedK1 {
    MapCode mc;
    const float* pts;
    int n;
    long long pitch_pts;
    unsigned int* maps;
    unsigned int* occ;
    int tiles_per_frame;
};
struct FusedGather {
    MapCode mc;
    const float* pts;
    long long pitch_pts;
    const unsigned int* maps;
    const unsigned int* occs;
    const double* uv;
    int F;
    double* depth;
    int* status;
    int* overflow_list;
    int* overflow_count;
    unsigned int* surv_rec;
    unsigned int* surv_idx;
    int* class_count;
    long long cap;
    int blocks_per_frame;
};
static_assert(SBT_A == K1_THREADS, "the fused launch uses one block size for both roles");

#ifndef MLD_FUSED_MINBLOCKS
#define MLD_FUSED_MINBLOCKS 8
#endif
// grid = (K1 tiles per frame + gather blocks per frame, frames): block (x, y) is K1 tile x of frame y of the new chunk or gather
// block x - tiles of frame y of the previous chunk, so a frame's 16 gather blocks follow its 118 K1 tiles in launch order (the two
// roles stay co-resident in that ratio) and the role costs two compares. (Round 2 before this: a 1-D grid with every `period`-th
// block a gather block; the three integer divisions that took a block to its role and frame were 8 % of the kernel's warp
// instructions, ncu r2o.)
__global__ void __launch_bounds__(SBT_A, MLD_FUSED_MINBLOCKS)
fused_project_gather_kernel(DevParams P, int stride_f, FusedK1 a, FusedGather g, int frames_k1, int frames_g) {
    const int x = (int)blockIdx.x, frame = (int)blockIdx.y;
    if (x >= a.tiles_per_frame) {  // uniform per block
        const int gb = x - a.tiles_per_frame;
        if (frame < frames_g && gb < g.blocks_per_frame)
            gather_block(P, g.mc, g.pts, stride_f, g.pitch_pts, g.maps, g.occs, g.uv, g.F, g.depth, g.status, g.overflow_list, g.overflow_count,
                         g.surv_rec, g.surv_idx, g.class_count, g.cap, gb, (long long)frame);
        return;
    }
    if (frame >= frames_k1) return;
    k1_tile(P, a.mc, a.pts, stride_f, a.n, a.pitch_pts, a.maps, a.occ, (unsigned int)frame, x);
}

// ---- K2b ------------------------------------------------------------------------------------------
// One block = TBT survivors of ONE class (equal neighbour count k); classes are laid out heaviest first. Phases:
//   load     raw indices -> points -> FP64 camera frame into the thread's shared-memory slab
//   P2a      histogram segmentation per thread -> n points kept (or a final status)
//   sort     block-wide counting sort of the still-alive survivors by n, largest first: the triangle search is O(n^2) and ran
//            with 10 of 32 lanes active while every warp held every n
//   P2b+P3   corner selection and the geometry tail for the sorted survivors (dense lanes, equal n inside a warp)
// CAP = slab entries = largest class of this instantiation, KLO = its smallest class, TBT = threads per block
template <int CAP, int KLO, int TBT, int MINB>
__global__ void __launch_bounds__(TBT, MINB)
feature_solve_kernel(DevParams P, const float* __restrict__ pts, int stride_f, long long pitch_pts, int F,
                     const double* __restrict__ uv, double* __restrict__ depth, int* __restrict__ status,
                     const unsigned int* __restrict__ surv_rec, const unsigned int* __restrict__ surv_idx,
                     const int* __restrict__ class_count, long long cap, int road, int* __restrict__ road_list,
                     int* __restrict__ road_count, int strided) {
    using TSlab = TSlabT<CAP, TBT>;
    __shared__ double sx[CAP * TBT], sy[CAP * TBT], sz[CAP * TBT];
    __shared__ int saux[CAP * TBT];
    __shared__ double s_u[TBT], s_v[TBT];
    __shared__ int s_o[TBT];
    __shared__ short s_list[TBT];
    __shared__ signed char s_st[TBT], s_cnt[TBT];
    __shared__ int s_hist[CAP + 1], s_start[CAP + 1];
    __shared__ int s_wtot[TBT / 32];
    __shared__ int s_base, s_n2;
    const int tid = threadIdx.x;
    // one work item (TBT survivors of one class) per block; a grid smaller than the item count strides over the items
    for (int item = (int)blockIdx.x;; item += (int)gridDim.x) {
    // which class, and which TBT survivors of it (uniform per block)
    int k = CAP, b = item, count = 0;
    for (; k >= KLO; k--) {
        count = __ldg(class_count + k);
        const int nb = (count + TBT - 1) / TBT;
        if (b < nb) break;
        b -= nb;
    }
    if (k < KLO) return;
    const long long slot = (long long)b * TBT + tid;
    const bool valid = slot < count;
    auto slab_of = [&](int owner) { return TSlab{sx + owner, sy + owner, sz + owner, saux + owner}; };

    // load + P2a
    int n = 0;
    bool alive = false;
    s_st[tid] = ST_Unspecified;
    if (tid <= CAP) s_hist[tid] = 0;
    if (valid) {
        const int o = (int)surv_rec[(long long)k * cap + slot];
        s_o[tid] = o;
        const double2 f2 = __ldg(reinterpret_cast<const double2*>(uv) + o);
        s_u[tid] = f2.x;
        s_v[tid] = f2.y;
        const TSlab s = slab_of(tid);
        // raw indices (coalesced: [row][slot]) -> points -> FP64 camera frame, the same expression K1 evaluated
        const float* fp = pts + (long long)(o / F) * pitch_pts * (long long)stride_f;
        const unsigned int* idx = surv_idx + (long long)class_row(k) * cap + slot;
        for (int i = 0; i < k; i++) {
            const int raw = (int)__ldg(idx + (long long)i * cap);
            s.A(i) = raw;
            // the points are random 16-byte reads of a cloud that left L2 long ago, and the slabs cap this kernel at 24 warps per SM:
            // ask L2 for all of a survivor's points at once (no registers held), the loads below then find them there
            asm volatile("prefetch.global.L2 [%0];" ::"l"(fp + (long long)raw * stride_f));
        }
#pragma unroll 2
        for (int i = 0; i < k; i++) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(fp + (long long)s.A(i) * stride_f));
            s.set(i, lidar_to_cam(P, q.x, q.y, q.z));
        }
        n = k;
        int st = ST_Unspecified;
        if (P.use_hist) {
            n = t_histogram_segment(P, k, s);
            if (n < 0) st = ST_HistogramNoLocalMax;
        }
        if (st == ST_Unspecified && n < 3)  // t_select_corners without the search (DepthEstimator.cpp:915-926)
            st = (!P.use_pca && P.use_tri_max) ? ST_TriangleNotPlanarInsufficientPoints : ST_HistogramNoLocalMax;
        if (st == ST_Unspecified) {
            alive = true;
            s_cnt[tid] = (signed char)n;
        } else {
            s_st[tid] = (signed char)st;
        }
    }
    __syncthreads();
    // counting sort of the alive survivors by n, largest n first
    int r = 0;
    if (alive) r = atomicAdd(&s_hist[n], 1);
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int j = CAP; j >= 0; j--) {
            s_start[j] = acc;
            acc += s_hist[j];
        }
        s_n2 = acc;
    }
    __syncthreads();
    if (alive) s_list[s_start[n] + r] = (short)tid;
    __syncthreads();

    // P2b + P3 on dense lanes
    if (tid < s_n2) {
        const int owner = s_list[tid];
        const TSlab s = slab_of(owner);
        const int no = (int)s_cnt[owner];
        int ci, cj, ck;
        int st = t_select_corners(P, no, s, ci, cj, ck);
        if (st == 0) {
            double dp;
            st = t_depth_from_corners(P, s_u[owner], s_v[owner], no, s, ci, cj, ck, dp);
            sx[owner] = dp;  // park the depth in the owner's slab (entry 0 is no longer needed once the tail is done)
        }
        s_st[owner] = (signed char)st;
    }
    __syncthreads();
    if (valid) {
        const int st = s_st[tid];
        const int o = s_o[tid];
        // the road path overwrites status/depth of its candidates later; everything gets a normal-path result first
        status[o] = st;
        depth[o] = (st == ST_Success) ? sx[tid] : -1.0;
    }
    if (road) {
        const bool want = valid && s_st[tid] != ST_Success;
        const int lane = tid & 31, warp = tid >> 5;
        const unsigned bm = __ballot_sync(MLD_FULL_MASK, want);
        if (lane == 0) s_wtot[warp] = __popc(bm);
        __syncthreads();
        int base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < TBT / 32; w++) {
            const int c = s_wtot[w];
            if (w < warp) base += c;
            total += c;
        }
        if (tid == 0) s_base = total ? atomicAdd(road_count, total) : 0;
        __syncthreads();
        if (want) road_list[s_base + base + __popc(bm & ((1u << lane) - 1u))] = s_o[tid];
    }
    if (!strided) return;  // the grid covers every item: one per block
    __syncthreads();  // the shared arrays are reused by the block's next item
    }
}

// ---- K2c: road gather ---------------------------------------------------------------------------------
// One thread per unsolved survivor. Most candidates die at the plane gate or have fewer than 3 plane inliers
// (the inlier set is restricted to the 6000-point RANSAC subsample, RansacPlane.cpp:100), so nothing is parked
// in shared memory here: the gate is evaluated while the neighbours stream by, and only the rare candidates
// that pass get their inlier points written to the road-survivor arrays for K2d.
__global__ void __launch_bounds__(SBT_A)
feature_road_gather_kernel(DevParams P, MapCode mc, const float* __restrict__ pts, int stride_f, long long pitch_pts,
                           const unsigned int* __restrict__ maps, const unsigned int* __restrict__ occs,
                           const double* __restrict__ uv, int F, double* __restrict__ depth, int* __restrict__ status,
                           const float* __restrict__ plane_coeffs, const unsigned int* __restrict__ inlier_bits,
                           long long inlier_words_per_frame, const int* __restrict__ road_list, const int* __restrict__ road_count,
                           int* __restrict__ overflow_list, int* __restrict__ overflow_count, unsigned int* __restrict__ rs_rec,
                           double* __restrict__ rs_xyz, int* __restrict__ rs_count, long long cap) {
    __shared__ int s_aux[RCAP * SBT_A];
    __shared__ int s_wtot[SBT_A / 32];
    __shared__ int s_base;
    const int tid = threadIdx.x;
    const int count = *road_count;
    const long long item = (long long)blockIdx.x * SBT_A + tid;
    if ((long long)blockIdx.x * SBT_A >= count) return;  // uniform per block
    const bool valid = item < count;
    int* aux = s_aux + tid;
    bool surv = false;
    int n_inl = 0;
    unsigned int inl_mask = 0u;
    long long o = 0;
    const float* fp = nullptr;
    const unsigned int* map = nullptr;
    if (valid) {
        o = road_list[item];
        const long long frame = o / F;
        fp = pts + frame * pitch_pts * (long long)stride_f;
        map = maps + frame * (long long)P.W * (long long)P.H;
        const unsigned int* occ = occs + frame * occ_words_per_frame(P.W, P.H);
        const float* pc = plane_coeffs + frame * 4;
        const unsigned int* bits = inlier_bits + frame * inlier_words_per_frame;
        const double2 f2 = __ldg(reinterpret_cast<const double2*>(uv) + o);
        const double u = f2.x, v = f2.y;
        // wide window (scale 2.0 x 1.5): occupancy words -> pixel offsets in scan order
        int k2 = 0;
        if ((fabs(u) < 1e9) && (fabs(v) < 1e9)) {
            const int x0 = (int)fmax(u - P.hx2, 0.), x1 = (int)fmin(u + P.hx2, (double)(P.W - 1));
            const int y0 = (int)fmax(v - P.hy2, 0.), y1 = (int)fmin(v + P.hy2, (double)(P.H - 1));
            if (x1 >= x0 && y1 >= y0) {
                occ_scan_window(occ, P.W, x0, x1, y0, y1, [&](int off) {
                    if (k2 < RCAP) aux[k2 * SBT_A] = off;
                    k2++;
                });
            }
        }
        if (k2 > RCAP) {
            overflow_list[atomicAdd(overflow_count, 1)] = (int)o;  // the warp kernel redoes the feature from scratch
        } else if ((unsigned)k2 < (unsigned)P.count_min) {           // DepthEstimator.cpp:585-586
            status[o] = ST_RadiusSearchInsufficientPoints;
            depth[o] = -1;
        } else {
            const float a = pc[0], b = pc[1], c = pc[2], d = pc[3];
#pragma unroll 4
            for (int i = 0; i < k2; i++) aux[i * SBT_A] = (int)map_cell_index(mc, __ldg(map + aux[i * SBT_A]));
            bool far = false;
            for (int i = 0; i < k2 && !far; i++) {
                const int raw = aux[i * SBT_A];
                const float4 q = __ldg(reinterpret_cast<const float4*>(fp + (long long)raw * stride_f));
                far = road_point_far(P, lidar_to_cam(P, q.x, q.y, q.z), a, b, c, d);
                if ((bits[raw >> 5] >> (raw & 31)) & 1u) {
                    inl_mask |= 1u << i;
                    n_inl++;
                }
            }
            // a far neighbour or fewer than 3 inliers: the normal path's status stands (already written by K2b)
            surv = !far && n_inl >= 3;
        }
    }
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned bm = __ballot_sync(MLD_FULL_MASK, surv);
    if (lane == 0) s_wtot[warp] = __popc(bm);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SBT_A / 32; w++) {
        const int c = s_wtot[w];
        if (w < warp) base += c;
        total += c;
    }
    if (tid == 0) s_base = total ? atomicAdd(rs_count, total) : 0;
    __syncthreads();
    if (!surv) return;
    const long long slot = (long long)s_base + base + __popc(bm & ((1u << lane) - 1u));
    rs_rec[slot] = pack_rec(n_inl, o);
    int e = 0;
    for (int i = 0; inl_mask; i++, inl_mask >>= 1) {
        if (!(inl_mask & 1u)) continue;
        const float4 q = __ldg(reinterpret_cast<const float4*>(fp + (long long)aux[i * SBT_A] * stride_f));
        const D3 c = lidar_to_cam(P, q.x, q.y, q.z);
        double* dst = rs_xyz + (long long)e * 3 * cap + slot;
        dst[0] = c.x;
        dst[cap] = c.y;
        dst[2 * cap] = c.z;
        e++;
    }
}

// ---- K2d: road solve ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(SBT_C)
feature_road_solve_kernel(DevParams P, const double* __restrict__ uv, int F, double* __restrict__ depth, int* __restrict__ status,
                          const float* __restrict__ plane_coeffs, const unsigned int* __restrict__ rs_rec,
                          const double* __restrict__ rs_xyz, const int* __restrict__ rs_count, long long cap) {
    using TSlab = TSlabT<RCAP, SBT_C>;
    __shared__ double sx[RCAP * SBT_C], sy[RCAP * SBT_C], sz[RCAP * SBT_C];
    __shared__ int saux[RCAP * SBT_C];
    const int tid = threadIdx.x;
    const long long slot = (long long)blockIdx.x * SBT_C + tid;
    if (slot >= *rs_count) return;
    const unsigned int rec = rs_rec[slot];
    const int n = (int)(rec >> 27);
    const long long o = (long long)(rec & 0x7FFFFFFu);
    const TSlab s{sx + tid, sy + tid, sz + tid, saux + tid};
    for (int i = 0; i < n; i++) {
        const double* src = rs_xyz + (long long)i * 3 * cap + slot;
        s.set(i, D3{src[0], src[cap], src[2 * cap]});
    }
    const double2 f2 = __ldg(reinterpret_cast<const double2*>(uv) + o);
    double dp;
    const int st = t_road_estimate(P, f2.x, f2.y, n, s, plane_coeffs + (o / F) * 4, dp);
    status[o] = st;
    depth[o] = (st == ST_SuccessRoad) ? dp : -1.0;
}

}  // namespace

size_t mld_split_scratch_bytes(long long features, int road) {
    // header (128 B: counters) + survivor records [class][features] + road list + the survivors' raw neighbour indices
    // [class rows][features]; with a plane the same area holds the road pass's (rarer, up to RCAP-entry) survivor points
    // [RCAP][3][features] once K2b has finished. A class can hold every feature, so most of this is never touched.
    const size_t idx_area = (size_t)features * class_row(GCAP + 1) * sizeof(unsigned int);
    const size_t road_area = road ? (size_t)features * RCAP * 3 * sizeof(double) : 0;
    return 128 + (size_t)features * ((GCAP + 1) * sizeof(unsigned int) + sizeof(int)) + (idx_area > road_area ? idx_area : road_area) + 256;
}

namespace {
struct SplitLayout {
    int *header, *class_count, *road_count, *rs_count, *road_list;
    unsigned int* surv_rec;
    unsigned int* surv_idx;  // normal path: raw neighbour indices [class rows][features]
    double* surv_xyz;        // road pass: survivor points [RCAP][3][features] (same area)
};
SplitLayout split_layout(void* d_scratch, long long features) {
    unsigned char* base = reinterpret_cast<unsigned char*>(d_scratch);
    SplitLayout L;
    L.header = reinterpret_cast<int*>(base);
    L.road_count = L.header + 1;
    L.rs_count = L.header + 2;
    L.class_count = L.header + CLASS_COUNT_AT;
    L.surv_rec = reinterpret_cast<unsigned int*>(base + 128);
    L.road_list = reinterpret_cast<int*>(L.surv_rec + (size_t)features * (GCAP + 1));
    size_t off = 128 + (size_t)features * ((GCAP + 1) * sizeof(unsigned int) + sizeof(int));
    off = (off + 255) & ~(size_t)255;
    L.surv_xyz = reinterpret_cast<double*>(base + off);
    L.surv_idx = reinterpret_cast<unsigned int*>(base + off);
    return L;
}
}  // namespace

// K1 of (frames_k1 frames at d_pts_k1 into maps_k1 / occ_k1) together with the gather of a previous chunk (frames_g frames
// whose maps are complete). Either part may be empty (frames_* == 0): the first launch of a sequence has no gather, the
// last no K1. The gather's counters are zeroed here.
cudaError_t mld_launch_fused_project_gather(const DevParams& P, int stride_f, const MapCode& mc_k1, const float* d_pts_k1, long long n_points,
                                            long long pitch_pts, unsigned int* d_maps_k1, unsigned int* d_occ_k1, int frames_k1,
                                            const MapCode& mc_g, const float* d_pts_g, const unsigned int* d_maps_g,
                                            const unsigned int* d_occ_g, const double* d_uv_g, int F, double* d_depth_g, int* d_status_g,
                                            int frames_g, int* d_overflow_list, int* d_overflow_count, void* d_scratch_g,
                                            cudaStream_t stream, int* launches) {
    const long long tiles = (n_points + K1_THREADS * K1_PPT - 1) / (K1_THREADS * K1_PPT);
    const long long k1_blocks = (frames_k1 > 0 && n_points > 0) ? tiles * frames_k1 : 0;
    const long long gbpf = (F + SBT_A - 1) / SBT_A;
    const long long g_blocks = (frames_g > 0 && F > 0) ? gbpf * frames_g : 0;
    if (k1_blocks + g_blocks == 0) return cudaSuccess;
    if (k1_blocks + g_blocks > 0x7fffffffLL || n_points > 0x7fffffffLL / 8) return cudaErrorInvalidValue;
    const long long features = (long long)frames_g * F;
    if (features >= (1ll << 27)) return cudaErrorInvalidValue;
    FusedK1 a{mc_k1, d_pts_k1, (int)n_points, pitch_pts, d_maps_k1, d_occ_k1, (int)std::max<long long>(tiles, 1)};
    FusedGather g{};
    g.blocks_per_frame = (int)std::max<long long>(gbpf, 1);
    if (g_blocks > 0) {
        const SplitLayout L = split_layout(d_scratch_g, features);
        cudaError_t e = cudaMemsetAsync(L.header, 0, 128, stream);
        if (e != cudaSuccess) return e;
        g = FusedGather{mc_g, d_pts_g, pitch_pts, d_maps_g, d_occ_g, d_uv_g, F, d_depth_g, d_status_g, d_overflow_list, d_overflow_count,
                        L.surv_rec, L.surv_idx, L.class_count, features, (int)gbpf};
    }
    const long long gx = (k1_blocks > 0 ? tiles : 0) + (g_blocks > 0 ? gbpf : 0);
    const int gy = std::max(k1_blocks > 0 ? frames_k1 : 0, g_blocks > 0 ? frames_g : 0);
    if (gx > 0x7fffffffLL || gy > 65535) return cudaErrorInvalidValue;
    a.tiles_per_frame = k1_blocks > 0 ? (int)tiles : 0;
    fused_project_gather_kernel<<<dim3((unsigned)gx, (unsigned)gy), SBT_A, 0, stream>>>(P, stride_f, a, g, k1_blocks > 0 ? frames_k1 : 0,
                                                                                       g_blocks > 0 ? frames_g : 0);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

// K2b on the survivors a gather (fused or not) left in d_scratch, followed by the road kernels when a plane is given
// (d_plane_coeffs != nullptr and a road estimator is configured)
cudaError_t mld_launch_feature_solve(const DevParams& P, const MapCode& mc, const float* d_pts, int stride_f, long long pitch_pts,
                                     const unsigned int* d_maps, const unsigned int* d_occ, const double* d_uv, int F, double* d_depth,
                                     int* d_status, const float* d_plane_coeffs, const unsigned int* d_inlier_bits,
                                     long long words_per_frame, int nframes, int* d_overflow_list, int* d_overflow_count, void* d_scratch,
                                     cudaStream_t stream, int* launches, cudaEvent_t* ev_after_solve) {
    if (F <= 0 || nframes <= 0) {
        if (ev_after_solve) return cudaEventRecord(*ev_after_solve, stream);
        return cudaSuccess;
    }
    const long long features = (long long)nframes * F;
    const SplitLayout L = split_layout(d_scratch, features);
    const bool road = d_plane_coeffs != nullptr && P.road_mode != ROAD_NONE;
    cudaError_t e;
    // every class rounds its survivors up to whole blocks; blocks past the last class leave at once
    const unsigned gb = (unsigned)((features + SBT_B - 1) / SBT_B) + SCAP + 1;
    feature_solve_kernel<SCAP, 0, SBT_B, MLD_SOLVE_MINBLOCKS><<<gb, SBT_B, 0, stream>>>(
        P, d_pts, stride_f, pitch_pts, F, d_uv, d_depth, d_status, L.surv_rec, L.surv_idx, L.class_count, features, road ? 1 : 0, L.road_list,
        L.road_count, 0);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (GCAP > SCAP) {
        // the rare fuller windows: 64-thread blocks with GCAP-entry slabs; a small grid strides over however many there are
        const unsigned gb2 = (unsigned)std::min<long long>((features + SBT_B2 - 1) / SBT_B2 + (GCAP - SCAP), 2 * 148);
        feature_solve_kernel<GCAP, (GCAP > SCAP ? SCAP + 1 : 0), SBT_B2, 4><<<gb2, SBT_B2, 0, stream>>>(
            P, d_pts, stride_f, pitch_pts, F, d_uv, d_depth, d_status, L.surv_rec, L.surv_idx, L.class_count, features, road ? 1 : 0,
            L.road_list, L.road_count, 1);
        if (launches) (*launches)++;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (ev_after_solve && (e = cudaEventRecord(*ev_after_solve, stream)) != cudaSuccess) return e;  // profiling: end of the solve
    if (launches) (*launches)++;
    if (road) {
        // the survivor arrays are free again once K2b has finished: the road pass reuses them
        const unsigned gc = (unsigned)((features + SBT_A - 1) / SBT_A);
        feature_road_gather_kernel<<<gc, SBT_A, 0, stream>>>(P, mc, d_pts, stride_f, pitch_pts, d_maps, d_occ, d_uv, F, d_depth, d_status,
                                                            d_plane_coeffs, d_inlier_bits, words_per_frame, L.road_list, L.road_count,
                                                            d_overflow_list, d_overflow_count, L.surv_rec, L.surv_xyz, L.rs_count, features);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        const unsigned gd = (unsigned)((features + SBT_C - 1) / SBT_C);
        feature_road_solve_kernel<<<gd, SBT_C, 0, stream>>>(P, d_uv, F, d_depth, d_status, d_plane_coeffs, L.surv_rec, L.surv_xyz, L.rs_count,
                                                           features);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (launches) *launches += 2;
    }
    return cudaSuccess;
}

cudaError_t mld_launch_feature_depth_split(const DevParams& P, const MapCode& mc, const float* d_pts, int stride_f,
                                           long long pitch_pts, const unsigned int* d_maps, const unsigned int* d_occ,
                                           const double* d_uv, int F, double* d_depth, int* d_status, const float* d_plane_coeffs,
                                           const unsigned int* d_inlier_bits, long long words_per_frame, int nframes,
                                           int* d_overflow_list, int* d_overflow_count, void* d_scratch, cudaStream_t stream,
                                           int* launches, cudaEvent_t* ev_mid) {
    if (F <= 0 || nframes <= 0) return cudaSuccess;
    const long long features = (long long)nframes * F;
    if (features >= (1ll << 27)) return cudaErrorInvalidValue;  // survivor records hold 27-bit feature ids
    const SplitLayout L = split_layout(d_scratch, features);
    cudaError_t e = cudaMemsetAsync(L.header, 0, 128, stream);
    if (e != cudaSuccess) return e;
    dim3 ga((unsigned)((F + SBT_A - 1) / SBT_A), (unsigned)nframes);
    feature_gather_kernel<<<ga, SBT_A, 0, stream>>>(P, mc, d_pts, stride_f, pitch_pts, d_maps, d_occ, d_uv, F, d_depth, d_status,
                                                   d_overflow_list, d_overflow_count, L.surv_rec, L.surv_idx, L.class_count, features);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (launches) (*launches)++;
    if (ev_mid && (e = cudaEventRecord(ev_mid[0], stream)) != cudaSuccess) return e;  // profiling: end of the gather
    return mld_launch_feature_solve(P, mc, d_pts, stride_f, pitch_pts, d_maps, d_occ, d_uv, F, d_depth, d_status, d_plane_coeffs, d_inlier_bits,
                                    words_per_frame, nframes, d_overflow_list, d_overflow_count, d_scratch, stream, launches,
                                    ev_mid ? ev_mid + 1 : nullptr);
}
