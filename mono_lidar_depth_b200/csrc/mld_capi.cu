// mld_capi.cu -- the C ABI of libmld_cuda.so (include/mld_c_api.h): handle, parameter loading,
// device buffers, stream pipelines and kernel launches. No arithmetic of the hot path lives here and
// there is no CPU fallback: without a CUDA device every compute entry point fails with MLD_ERR_CUDA.
#include <cuda_runtime.h>
#include <immintrin.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <fstream>
#include <functional>
#include <mutex>
#include <thread>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "mld_c_api.h"
#include "mld_common.cuh"
#include "mld_host_pack.h"
#include "mld_kernels.h"

namespace {

thread_local std::string g_create_error;

constexpr int MLD_PIPE_SLOTS = 6;   // slots (streams + buffers) a handle owns
constexpr int MLD_HOST_SLOTS = 3;   // of which the host-buffer pipeline uses
constexpr int MLD_PREV_SLOT = 5;    // holds the previous cloud of mld_calculate_depth_pair_resident (batched paths use slots 0..2)

// everything one in-flight chunk of frames needs on the device
struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    cudaEvent_t ev_k1 = nullptr, ev_k2 = nullptr;  // priority mode: K1 of the slot's chunk finished / its K2 finished
    void* d_pts = nullptr;      size_t pts_bytes = 0;
    double* d_uv = nullptr;     size_t uv_bytes = 0;
    double* d_depth = nullptr;  size_t depth_bytes = 0;
    int* d_status = nullptr;    size_t status_bytes = 0;
    unsigned int* d_maps = nullptr; size_t maps_bytes = 0;
    unsigned int* d_bits = nullptr; size_t bits_bytes = 0;
    float* d_coeffs = nullptr;  size_t coeffs_bytes = 0;
    void* d_scratch = nullptr;  size_t scratch_bytes = 0;
    int* d_small = nullptr;     size_t small_bytes = 0;  // n_inliers | iterations | rc, per frame
    int* d_ovf = nullptr;       size_t ovf_bytes = 0;    // [0] overflow count, [1..] global feature ids
    unsigned char* d_labels = nullptr; size_t labels_bytes = 0; // semantic label image (mld_semantic_ground_plane)
    unsigned char* d_sem = nullptr; size_t sem_bytes = 0;       // SemanticPlane scratch
    unsigned char* d_flags = nullptr; size_t flags_bytes = 0;   // ground-labelled flag per point (SemanticPlane exact mode)
    unsigned int* d_occ = nullptr; size_t occ_bytes = 0; // occupancy bitmaps of the maps
    void* d_split = nullptr;    size_t split_bytes = 0;  // survivor / road lists of the split K2 kernels
    unsigned char* h_stage = nullptr; size_t h_stage_bytes = 0;  // pinned staging of the packed (12-byte xyz) host pipeline
    float* d_pack = nullptr;    size_t d_pack_bytes = 0;          // its device copy, expanded to float4 in d_pts
    cudaEvent_t ev_stage = nullptr;                                // the H2D copy out of h_stage has finished
    bool stage_busy = false;                                       // ev_stage has been recorded in the current call
    int* h_ovf_seen = nullptr;  // pinned: overflow count of this slot's previous chunk (sizes the next overflow launch)
    unsigned epoch = 0;         // uses of d_maps since its last clear (tagged mode), 0 = never cleared
    size_t occ_clean_bytes = 0; // leading bytes of d_occ known to be zero once the slot's `done` event has fired (fused pipeline)
    MapCode mc = {0u, 0u};      // encoding of the maps currently held by this slot
};

template <typename T>
cudaError_t ensure(T*& p, size_t& cap, size_t need, bool* changed = nullptr) {
    if (need <= cap && p != nullptr) return cudaSuccess;
    if (changed) *changed = true;
    if (p) {
        cudaError_t e = cudaFree(p);
        if (e != cudaSuccess) return e;
        p = nullptr;
        cap = 0;
    }
    size_t alloc = need + need / 4 + 256;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, alloc);
    if (e != cudaSuccess) return e;
    p = static_cast<T*>(q);
    cap = alloc;
    return cudaSuccess;
}

}  // namespace

// Host worker threads that strip pcl::PointXYZI records (32 bytes: x y z pad | intensity pad pad pad) down to 12-byte xyz in
// pinned staging buffers: the host-buffer pipeline is PCIe bound, and only 12 of the 32 bytes are ever used
// (DepthEstimator.cpp:169 casts topRows<3>). run() blocks until every item is done; the calling thread works too.
class HostPool {
public:
    explicit HostPool(int workers) {
        for (int i = 0; i < workers; i++) threads_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    void run(int items, const std::function<void(int)>& fn) {
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn;
            items_ = items;
            next_.store(0);
            done_.store(0);
            generation_++;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(m_);
        // every item is done AND no worker is still inside work() with this job's function pointer
        cv_done_.wait(lk, [this] { return done_.load() >= items_ && active_ == 0; });
        fn_ = nullptr;
    }
    int workers() const { return (int)threads_.size(); }

private:
    void work() {
        const std::function<void(int)>* fn = fn_;
        const int items = items_;
        while (fn) {
            const int i = next_.fetch_add(1);
            if (i >= items) break;
            (*fn)(i);
            if (done_.fetch_add(1) + 1 >= items) {
                std::lock_guard<std::mutex> lk(m_);
                cv_done_.notify_all();
            }
        }
    }
    void loop() {
        unsigned long long seen = 0;
        while (true) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                if (!fn_) continue;  // the job finished before this worker woke up
                active_++;
            }
            work();
            {
                std::lock_guard<std::mutex> lk(m_);
                active_--;
            }
            cv_done_.notify_all();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, cv_done_;
    const std::function<void(int)>* fn_ = nullptr;
    int items_ = 0;
    std::atomic<int> next_{0}, done_{0};
    unsigned long long generation_ = 0;
    int active_ = 0;
    bool stop_ = false;
};

struct mld_handle {
    mld_params params;
    DevParams dp;
    int device = 0;
    bool initialized = false;
    bool have_cloud = false;
    int kcap = 0;
    int chunk_frames = 128;
    // 2: split gather/solve(/road) thread-per-feature kernels + warp-per-feature overflow pass (default); 0: warp per feature only
    int feature_mode = 2;
    int overflow_blocks = 296;
    bool use_tagged_maps = true;
    bool fuse_k1_gather = true;     // K1 of chunk j and the gather of chunk j-1 in one heterogeneous launch (MLD_FUSE=0: off)
    int fuse_chunk = 512;           // frames per fused launch (MLD_FUSE_CHUNK)
    bool fuse_road = true;          // the road / SemanticPlane / external-plane sequences use the fused pipeline too (MLD_FUSE=2: no)
    int overlap_slots = 5;          // chunks of a device-resident sequence alternate over this many slots/streams (MLD_OVERLAP). The solve and overflow pass of a chunk are starved by the fused launch that runs beside them and finish near its end; with 3 slots the launch after next waited for them (1.47 M frames/s), 4 / 5 / 6 slots: 1.50 / 1.52 / 1.49 M
    cudaEvent_t ev_fork = nullptr;
    // priority mode: every K1 of a sequence runs on a low-priority stream, every K2 on a high-priority one, so the
    // latency-bound K2 blocks are placed first and the streaming K1 fills what is left
    int overlap_mode = 0;           // 0: whole chunks alternate over slot streams (faster, measured), 1: priority streams
    int sm_count = 148;
    cudaStream_t st_lo = nullptr, st_hi = nullptr;
    cudaEvent_t ev_join = nullptr;
    long long cur_n = 0;
    int cur_stride_f = 4;
    bool cloud_in_flight = false;   // the last mld_set_cloud returned without waiting for slot 0's stream (packed upload)
    bool have_prev = false;         // slots[MLD_PREV_SLOT] holds the previous cloud of mld_calculate_depth_pair_resident
    long long prev_n = 0;
    int prev_stride_f = 4;
    Slot slots[MLD_PIPE_SLOTS];
    bool semantic_exact = false;    // mld_set_semantic_exact / MLD_SEMANTIC_EXACT=1: PCL's sequential float accumulation order
    bool stats_on = false;          // mld_set_statistics: status histogram of every mld_calculate_depth call
    unsigned long long* d_hist = nullptr;  // 21 counters on the device
    unsigned long long* h_hist = nullptr;  // pinned copy
    int64_t last_hist[21] = {0};
    bool last_hist_valid = false;
    // host-buffer pipeline: pack xyz on host threads (always for strides > 16 bytes; MLD_HOST_PACK=1/0 forces it on / off)
    int host_pack = -1;
    int host_pack_threads = 0;      // MLD_PACK_THREADS (0 = hardware concurrency - 2, at most 14)
    HostPool* pool = nullptr;
    double pcie_gbs = 42.0;         // link rate the pack / copy choice is modelled with (MLD_PCIE_GBS): below the ~55 GB/s of an idle link because the copy engine shares the host's memory system with the pack threads (25 / 30 / 35 / 40 / 45 / 50 / 60 / 75: 20.1 / 20.0 / 20.8 / 21.7 / 21.5 / 20.0 / 19.7 / 18.9 k frames/s on 32-byte records, 16-core host)
    int64_t host_stats[4] = {0, 0, 0, 0};  // since creation: H2D bytes, D2H bytes, frames packed, frames copied as they are
    int* d_dbg = nullptr;           // neighbour debug buffer
    float* d_synth_tables = nullptr;
    mld_synth_config synth_cfg_cached;
    bool synth_tables_valid = false;
    long long launches = 0;
    std::string error;
    // profiling (mld_profile_enable): events around each kernel class of a chunk
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_events;  // 5 per sampled chunk
    std::vector<int> prof_frames;          // frames of each sampled chunk
    std::vector<int> prof_ransac_launches;
    size_t prof_used = 0;                  // sampled chunks
};
constexpr size_t MLD_PROF_MAX_CHUNKS = 2048;
constexpr int MLD_PROF_EVENTS = 8;  // per sampled chunk: start, clear, K1, K2-stream start, K4, end, gather end, solve end

namespace {

int fail(mld_handle* h, int code, const std::string& msg) {
    if (h) h->error = msg;
    else g_create_error = msg;
    return code;
}
int fail_cuda(mld_handle* h, cudaError_t e, const char* where) {
    return fail(h, MLD_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}

#define CK(call)                                               \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return fail_cuda(h, e__, #call); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

RansacConfig ransac_config(const mld_params& p) {
    RansacConfig c;
    c.distance_treshold = p.ransac_plane_distance_treshold;
    c.refinement_treshold = p.ransac_plane_refinement_treshold;
    c.probability = p.ransac_plane_probability;
    c.min_z = p.ransac_plane_min_z;
    c.max_z = p.ransac_plane_max_z;
    c.max_iterations = p.ransac_plane_max_iterations;
    c.use_refinement = p.ransac_plane_use_refinement;
    c.cos_eps = cos(M_PI / 18.);
    c.log_probability = log(1.0 - p.ransac_plane_probability);
    // (double)dist < threshold  <=>  dist <= largest float strictly below the threshold (dist is a float)
    auto below = [](double thr) {
        float f = (float)thr;
        if (!((double)f < thr)) f = nextafterf(f, -INFINITY);
        return f;
    };
    c.thr_lt = below(p.ransac_plane_distance_treshold);
    c.refine_lt = below(p.ransac_plane_refinement_treshold);
    return c;
}

// largest number of pixels NeighborFinderPixel::getNeighbors can scan for half sizes (hx, hy)
int max_window_area(double hx, double hy, int W, int H) {
    double cx = floor(2.0 * hx) + 2.0, cy = floor(2.0 * hy) + 2.0;
    if (cx > W) cx = W;
    if (cy > H) cy = H;
    if (cx < 0) cx = 0;
    if (cy < 0) cy = 0;
    double a = cx * cy;
    return a > 1e9 ? 1000000000 : (int)a;
}

int slot_reserve(mld_handle* h, Slot& s, long long n_points, int stride_bytes, int F, int frames, bool own_io, bool road) {
    const size_t WH = (size_t)h->dp.W * (size_t)h->dp.H;
    if (own_io) {
        CK(ensure(s.d_pts, s.pts_bytes, (size_t)frames * (size_t)n_points * (size_t)stride_bytes));
        CK(ensure(s.d_uv, s.uv_bytes, (size_t)frames * (size_t)F * 2 * sizeof(double)));
        CK(ensure(s.d_depth, s.depth_bytes, (size_t)frames * (size_t)F * sizeof(double)));
        CK(ensure(s.d_status, s.status_bytes, (size_t)frames * (size_t)F * sizeof(int)));
    }
    bool maps_changed = false;
    CK(ensure(s.d_maps, s.maps_bytes, (size_t)frames * WH * sizeof(unsigned int), &maps_changed));
    if (maps_changed) s.epoch = 0;  // fresh memory holds no valid tags
    CK(ensure(s.d_ovf, s.ovf_bytes, ((size_t)frames * (size_t)std::max(F, 1) + 1) * sizeof(int)));
    bool occ_changed = false;
    CK(ensure(s.d_occ, s.occ_bytes, (size_t)frames * (size_t)occ_words_per_frame(h->dp.W, h->dp.H) * sizeof(unsigned int), &occ_changed));
    if (occ_changed) s.occ_clean_bytes = 0;
    if (road) {
        const size_t words = (size_t)((n_points + 31) / 32);
        CK(ensure(s.d_bits, s.bits_bytes, (size_t)frames * words * sizeof(unsigned int)));
        CK(ensure(s.d_coeffs, s.coeffs_bytes, (size_t)frames * 4 * sizeof(float)));
        CK(ensure(s.d_scratch, s.scratch_bytes, mld_ransac_scratch_bytes(n_points, frames)));
        CK(ensure(s.d_small, s.small_bytes, (size_t)frames * 3 * sizeof(int)));
    }
    return MLD_OK;
}

// Prepare the slot's maps for `frames` new frames: tagged mode bumps the epoch (and clears only when the
// 14-bit epoch space is exhausted or the memory is fresh), plain mode clears every time.
int begin_maps(mld_handle* h, Slot& s, int frames, long long n_points, cudaStream_t st, MapCode& mc, bool occ_may_be_clean = false) {
    const size_t WH = (size_t)h->dp.W * (size_t)h->dp.H;
    const bool tagged = h->use_tagged_maps && n_points <= (long long)(MLD_TAG_IDX_MASK + 1u);
    if (!tagged) {
        CK(cudaMemsetAsync(s.d_maps, 0xFF, (size_t)frames * WH * sizeof(unsigned int), st));
        s.epoch = 0;
        mc = MapCode{0u, 0u};
    } else {
        if (s.epoch == 0 || s.epoch >= MLD_TAG_MAX_EPOCH) {
            // whole buffer: frames beyond this chunk may hold stale tags from a previous epoch cycle
            CK(cudaMemsetAsync(s.d_maps, 0xFF, s.maps_bytes, st));
            s.epoch = 0;
        }
        s.epoch++;
        mc = MapCode{1u, MLD_TAG_MAX_EPOCH - s.epoch};
    }
    s.mc = mc;
    if (h->feature_mode >= 1) {
        // the fused pipeline clears a slot's occupancy bitmaps on the slot's own stream behind its last reader (off the front
        // stream's critical path); every other path clears here, right before use
        const size_t need = (size_t)frames * (size_t)occ_words_per_frame(h->dp.W, h->dp.H) * sizeof(unsigned int);
        if (!(occ_may_be_clean && s.occ_clean_bytes >= need)) CK(cudaMemsetAsync(s.d_occ, 0, need, st));
        s.occ_clean_bytes = 0;  // about to be written
    }
    return MLD_OK;
}

// The overflow pass runs the (slow) warp-per-feature kernel over a list whose length is only known on the
// device. Its grid is sized from the overflow count this slot saw on its previous chunk (read back
// asynchronously into pinned memory, never waited for): an empty list costs a 4-block launch instead of a
// machine-filling one. Any grid size is correct -- the warps stride the list.
int overflow_grid(mld_handle* h, Slot& s) {
    const int seen = *reinterpret_cast<volatile int*>(s.h_ovf_seen);
    const long long want = ((long long)seen * 2 + 7) / 8 + 4;  // 8 warps per block, 2x head-room
    return (int)std::min<long long>(h->overflow_blocks, std::max<long long>(4, want));
}

// K2: thread-per-feature kernel + warp-per-feature pass over its overflow list, or warp-per-feature only
int launch_features(mld_handle* h, Slot& s, const MapCode& mc, cudaStream_t st, const float* d_pts, int stride_f, long long pitch_pts,
                    const double* d_uv, int F, double* d_depth, int* d_status, const float* coeffs, const unsigned int* bits,
                    long long words, int frames, cudaEvent_t* ev_mid = nullptr) {
    if (F <= 0 || frames <= 0) {
        if (ev_mid) {
            CK(cudaEventRecord(ev_mid[0], st));
            CK(cudaEventRecord(ev_mid[1], st));
        }
        return MLD_OK;
    }
    if (h->feature_mode == 2) {
        CK(ensure(s.d_split, s.split_bytes, mld_split_scratch_bytes((long long)frames * F, coeffs != nullptr)));
        CK(cudaMemsetAsync(s.d_ovf, 0, sizeof(int), st));
        int nl = 0;
        CK(mld_launch_feature_depth_split(h->dp, mc, d_pts, stride_f, pitch_pts, s.d_maps, s.d_occ, d_uv, F, d_depth, d_status, coeffs, bits,
                                          words, frames, s.d_ovf + 1, s.d_ovf, s.d_split, st, &nl, ev_mid));
        CK(mld_launch_feature_depth(h->dp, mc, h->kcap, d_pts, stride_f, pitch_pts, s.d_maps, d_uv, F, d_depth, d_status, coeffs,
                                    bits, words, frames, s.d_ovf + 1, s.d_ovf, overflow_grid(h, s), st));
        CK(cudaMemcpyAsync(s.h_ovf_seen, s.d_ovf, sizeof(int), cudaMemcpyDeviceToHost, st));
        h->launches += nl + 1;
    } else {
        if (ev_mid) {
            CK(cudaEventRecord(ev_mid[0], st));
            CK(cudaEventRecord(ev_mid[1], st));
        }
        CK(mld_launch_feature_depth(h->dp, mc, h->kcap, d_pts, stride_f, pitch_pts, s.d_maps, d_uv, F, d_depth, d_status, coeffs,
                                    bits, words, frames, nullptr, nullptr, 0, st));
        h->launches++;
    }
    return MLD_OK;
}

// where the ground plane of a batched frame comes from (road path): fitted by RANSAC on the device (the road flag of
// mld_process_frames_*), fitted from a semantic label image on the device (SemanticPlane), or handed in by the caller
struct PlaneSrc {
    enum Kind { RANSAC = 0, SEMANTIC = 1, EXTERNAL = 2 } kind = RANSAC;
    // SEMANTIC (pointers already offset to the chunk's first frame by the caller of enqueue_chunk)
    const unsigned char* d_labels = nullptr;
    int label_w = 0, label_h = 0;
    double f = 0, cu = 0, cv = 0, T[12] = {0}, inlier_threshold = 0;
    unsigned int ground[8] = {0};
    int* d_rc_out = nullptr;
    // EXTERNAL
    const float* d_coeffs = nullptr;
    const unsigned int* d_bits = nullptr;
};

// exact mode of the SemanticPlane fit: the labelled flags of `frames` clouds live in the slot
int sem_flags(mld_handle* h, Slot& s, long long n_points, long long frames, unsigned char** out) {
    *out = nullptr;
    if (!h->semantic_exact) return MLD_OK;
    CK(ensure(s.d_flags, s.flags_bytes, (size_t)std::max<long long>(n_points * frames, 1)));
    *out = s.d_flags;
    return MLD_OK;
}

void ground_label_set(const int32_t* labels, int n, unsigned int set8[8]) {
    for (int i = 0; i < 8; i++) set8[i] = 0u;
    for (int i = 0; i < n; i++)
        if (labels[i] >= 0 && labels[i] <= 255) set8[labels[i] >> 5] |= 1u << (labels[i] & 31);
}


// profiling: the event set of the next sampled chunk (nullptr when profiling is off or the pool is exhausted)
int prof_acquire(mld_handle* h, int frames, cudaEvent_t** out) {
    *out = nullptr;
    if (!h->prof_on || h->prof_used >= MLD_PROF_MAX_CHUNKS) return MLD_OK;
    if (h->prof_events.size() < (h->prof_used + 1) * MLD_PROF_EVENTS) {
        for (int q = 0; q < MLD_PROF_EVENTS; q++) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            h->prof_events.push_back(e);
        }
        h->prof_frames.push_back(0);
        h->prof_ransac_launches.push_back(0);
    }
    *out = &h->prof_events[h->prof_used * MLD_PROF_EVENTS];
    h->prof_frames[h->prof_used] = frames;
    h->prof_ransac_launches[h->prof_used] = 0;
    return MLD_OK;
}
// one chunk of frames on one stream: [clear maps], K1, [K4], K2
int enqueue_chunk(mld_handle* h, Slot& s, cudaStream_t st, const float* d_pts, long long n_points, long long pitch_pts,
                  int stride_f, const double* d_uv, int F, double* d_depth, int* d_status, int frames, int road, uint64_t seed,
                  long long frame0, float* d_coeffs_out, cudaStream_t st_k2 = nullptr, const PlaneSrc* src = nullptr) {
    // st: stream of the map clear + K1 (and of everything when st_k2 is null); st_k2: stream of K4 + K2
    const bool two = st_k2 != nullptr && st_k2 != st;
    cudaStream_t sb = two ? st_k2 : st;
    cudaEvent_t* ev = nullptr;
    int rcp = prof_acquire(h, frames, &ev);
    if (rcp) return rcp;
    if (two) CK(cudaStreamWaitEvent(st, s.ev_k2, 0));  // the slot's maps are free once its previous chunk's K2 has finished
    if (ev) CK(cudaEventRecord(ev[0], st));
    MapCode mc;
    int rcm = begin_maps(h, s, frames, n_points, st, mc);
    if (rcm) return rcm;
    if (ev) CK(cudaEventRecord(ev[1], st));
    CK(mld_launch_project_scatter(h->dp, mc, d_pts, stride_f, n_points, pitch_pts, s.d_maps, h->feature_mode >= 1 ? s.d_occ : nullptr, frames, st));
    if (n_points > 0) h->launches++;
    if (ev) CK(cudaEventRecord(ev[2], st));
    if (two) {
        CK(cudaEventRecord(s.ev_k1, st));
        CK(cudaStreamWaitEvent(sb, s.ev_k1, 0));
    }
    if (ev) CK(cudaEventRecord(ev[3], sb));
    const float* coeffs = nullptr;
    const unsigned int* bits = nullptr;
    const long long words = (n_points + 31) / 32;
    if (road && h->dp.road_mode != ROAD_NONE) {
        float* cdst = d_coeffs_out ? d_coeffs_out : s.d_coeffs;
        int nl = 0;
        if (src && src->kind == PlaneSrc::EXTERNAL) {
            coeffs = src->d_coeffs;
            bits = src->d_bits;
        } else if (src && src->kind == PlaneSrc::SEMANTIC) {
            // SemanticPlane::CalculateInliersPlane per frame (what TrackletDepthModule::process does before CalculateDepth)
            CK(ensure(s.d_sem, s.sem_bytes, mld_semantic_state_bytes(frames)));
            unsigned char* fl = nullptr;
            int rcs = sem_flags(h, s, n_points, frames, &fl);
            if (rcs) return rcs;
            CK(mld_launch_semantic_plane(src->T, src->f, src->cu, src->cv, src->label_w, src->label_h, src->ground, src->inlier_threshold,
                                         d_pts, stride_f, n_points, pitch_pts, src->d_labels, frames, s.d_sem, cdst, s.d_bits, words,
                                         s.d_small, src->d_rc_out ? src->d_rc_out : s.d_small + 2 * frames, sb, &nl, fl, fl ? 1 : 0));
            coeffs = cdst;
            bits = s.d_bits;
        } else {
            CK(mld_launch_ransac(ransac_config(h->params), d_pts, stride_f, n_points, pitch_pts, frames, seed, frame0, s.d_scratch,
                                 cdst, s.d_bits, words, s.d_small, s.d_small + frames, s.d_small + 2 * frames, sb, &nl));
            coeffs = cdst;
            bits = s.d_bits;
        }
        h->launches += nl;
        if (ev) h->prof_ransac_launches[h->prof_used] = nl;
    }
    if (ev) CK(cudaEventRecord(ev[4], sb));
    int rcf = launch_features(h, s, mc, sb, d_pts, stride_f, pitch_pts, d_uv, F, d_depth, d_status, coeffs, bits, words, frames,
                              ev ? ev + 6 : nullptr);
    if (rcf) return rcf;
    if (two) CK(cudaEventRecord(s.ev_k2, sb));
    if (ev) {
        CK(cudaEventRecord(ev[5], sb));
        h->prof_used++;
    }
    return MLD_OK;
}

// ---- flat OpenCV-YAML reader (cv::FileStorage subset used by DepthEstimatorParameters::fromFile) ----
bool parse_yaml(const char* path, std::map<std::string, std::string>& kv) {
    std::ifstream in(path);
    if (!in.is_open()) return false;
    std::string line;
    while (std::getline(in, line)) {
        size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        if (line.empty() || line[0] == '%') continue;
        size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        auto trim = [](std::string s) {
            size_t a = s.find_first_not_of(" \t\r\n\"'"), b = s.find_last_not_of(" \t\r\n\"'");
            return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
        };
        std::string k = trim(line.substr(0, colon)), v = trim(line.substr(colon + 1));
        if (!k.empty()) kv[k] = v;
    }
    return true;
}
// cv::FileNode -> double: absent 0, int node -> (double)i, real node -> f
double yaml_double(const std::map<std::string, std::string>& kv, const char* key) {
    auto it = kv.find(key);
    if (it == kv.end() || it->second.empty()) return 0.0;
    char* end = nullptr;
    double d = strtod(it->second.c_str(), &end);
    return end == it->second.c_str() ? 0.0 : d;
}
// cv::FileNode -> int: absent 0, int node -> i, real node -> cvRound(f) (round half to even)
int yaml_int(const std::map<std::string, std::string>& kv, const char* key) {
    auto it = kv.find(key);
    if (it == kv.end() || it->second.empty()) return 0;
    const std::string& s = it->second;
    char* end = nullptr;
    long li = strtol(s.c_str(), &end, 0);
    if (end != s.c_str() && *end == '\0') return (int)li;
    double d = strtod(s.c_str(), &end);
    if (end == s.c_str()) return 0x7fffffff;
    return (int)nearbyint(d);
}

}  // namespace

extern "C" {

int mld_sizeof_params(void) { return (int)sizeof(mld_params); }

void mld_default_params(mld_params* p) {
    memset(p, 0, sizeof(*p));
    // DepthEstimatorParameters.h:12-172 member initialisers
    p->neighbor_search_mode = 0;
    p->pixelarea_search_witdh = 12;
    p->pixelarea_search_height = 15;
    p->radiusSearch_count_min = 3;
    p->do_use_histogram_segmentation = 1;
    p->histogram_segmentation_bin_witdh = 0.5;
    p->histogram_segmentation_min_pointcount = 3;
    p->do_use_depth_segmentation = 0;
    p->treshold_depth_enabled = 1;
    p->treshold_depth_mode = 0;
    p->treshold_depth_max = 100;
    p->treshold_depth_min = 0;
    p->treshold_depth_local_enabled = 1;
    p->treshold_depth_local_mode = 0;
    p->treshold_depth_local_valuetype = 1;
    p->treshold_depth_local_value = 0.5;
    p->do_use_PCA = 0;
    p->pca_debug = 0;
    p->pca_treshold_3_abs_min = 0.005;
    p->pca_treshold_3_2_rel_max = 15;
    p->pca_treshold_2_1_rel_min = 0.5;
    p->do_use_ransac_plane = 1;
    p->ransac_plane_distance_treshold = 0.2;
    p->ransac_plane_min_z = -10000;
    p->ransac_plane_max_z = 10000;
    p->ransac_plane_max_iterations = 10000;
    p->ransac_plane_use_refinement = 1;
    p->ransac_plane_refinement_treshold = 10.2;
    p->ransac_plane_use_camx_treshold = 0;
    p->ransac_plane_treshold_camx = 2.0;
    p->ransac_plane_point_distance_treshold = 0.2;
    p->ransac_plane_probability = 0.999;
    p->plane_estimator_use_triangle_maximation = 0;
    p->plane_estimator_z_x_min_relation = 0;
    p->plane_estimator_use_leastsquares = 0;
    p->plane_estimator_use_mestimator = 1;
    p->do_use_cut_behind_camera = 1;
    p->do_use_triangle_size_maximation = 1;
    p->do_check_triangleplanar_condition = 1;
    p->triangleplanar_crossnorm_treshold = 0.1;
    p->viewray_plane_orthoganality_treshold = 1.0;  // the header's `{01}` is an octal literal
    p->set_all_depths_to_zero = 0;
}

int mld_params_from_yaml(const char* path, mld_params* p) {
    std::map<std::string, std::string> kv;
    if (!path || !p) return fail(nullptr, MLD_ERR_INVALID_ARG, "mld_params_from_yaml: null argument");
    if (!parse_yaml(path, kv)) return fail(nullptr, MLD_ERR_IO, std::string("Cant find settings file: ") + path);
    memset(p, 0, sizeof(*p));
    // every read below mirrors one line of DepthEstimatorParameters::fromFile; bools are read through (int)
    p->neighbor_search_mode = yaml_int(kv, "neighbor_search_mode");
    p->pixelarea_search_witdh = yaml_int(kv, "pixelarea_search_witdh");
    p->pixelarea_search_height = yaml_int(kv, "pixelarea_search_height");
    p->radiusSearch_count_min = yaml_int(kv, "radiusSearch_count_min");
    p->do_use_histogram_segmentation = yaml_int(kv, "do_use_histogram_segmentation") != 0;
    p->histogram_segmentation_bin_witdh = yaml_double(kv, "histogram_segmentation_bin_witdh");
    p->histogram_segmentation_min_pointcount = yaml_int(kv, "histogram_segmentation_min_pointcount");
    p->do_use_depth_segmentation = yaml_int(kv, "do_use_depth_segmentation") != 0;
    p->treshold_depth_enabled = yaml_int(kv, "treshold_depth_enabled") != 0;
    p->treshold_depth_mode = yaml_int(kv, "treshold_depth_mode");
    p->treshold_depth_max = yaml_int(kv, "treshold_depth_max");
    p->treshold_depth_min = yaml_int(kv, "treshold_depth_min");
    p->treshold_depth_local_enabled = yaml_int(kv, "treshold_depth_local_enabled") != 0;
    p->treshold_depth_local_mode = yaml_int(kv, "treshold_depth_local_mode");
    p->treshold_depth_local_valuetype = yaml_int(kv, "treshold_depth_local_valuetype");
    p->treshold_depth_local_value = yaml_double(kv, "treshold_depth_local_value");
    p->do_use_PCA = yaml_double(kv, "do_use_PCA") != 0.0;  // read through (double) upstream (:76)
    p->pca_debug = yaml_int(kv, "pca_debug") != 0;
    p->pca_treshold_3_abs_min = yaml_double(kv, "pca_treshold_3_abs_min");
    p->pca_treshold_3_2_rel_max = yaml_double(kv, "pca_treshold_3_2_rel_max");
    p->pca_treshold_2_1_rel_min = yaml_double(kv, "pca_treshold_2_1_rel_min");
    p->do_use_ransac_plane = yaml_int(kv, "do_use_ransac_plane") != 0;
    p->ransac_plane_distance_treshold = yaml_double(kv, "ransac_plane_distance_treshold");
    p->ransac_plane_min_z = yaml_double(kv, "ransac_plane_min_z");
    p->ransac_plane_max_z = yaml_double(kv, "ransac_plane_max_z");
    p->ransac_plane_max_iterations = yaml_int(kv, "ransac_plane_max_iterations");
    p->ransac_plane_use_refinement = yaml_int(kv, "ransac_plane_use_refinement") != 0;
    p->ransac_plane_refinement_treshold = yaml_double(kv, "ransac_plane_refinement_treshold");
    p->ransac_plane_use_camx_treshold = yaml_int(kv, "ransac_plane_use_camx_treshold") != 0;
    p->ransac_plane_treshold_camx = yaml_double(kv, "ransac_plane_treshold_camx");
    p->ransac_plane_point_distance_treshold = yaml_double(kv, "ransac_plane_point_distance_treshold");
    p->ransac_plane_probability = yaml_double(kv, "ransac_plane_probability");
    p->plane_estimator_use_triangle_maximation = yaml_int(kv, "plane_estimator_use_triangle_maximation") != 0;
    p->plane_estimator_z_x_min_relation = yaml_double(kv, "plane_estimator_z_x_min_relation");
    p->plane_estimator_use_leastsquares = yaml_int(kv, "plane_estimator_use_leastsquares") != 0;
    p->plane_estimator_use_mestimator = yaml_int(kv, "plane_estimator_use_mestimator") != 0;
    p->do_use_cut_behind_camera = yaml_int(kv, "do_use_cut_behind_camera") != 0;
    p->do_use_triangle_size_maximation = yaml_int(kv, "do_use_triangle_size_maximation") != 0;
    p->do_check_triangleplanar_condition = yaml_int(kv, "do_check_triangleplanar_condition") != 0;
    p->triangleplanar_crossnorm_treshold = yaml_double(kv, "triangleplanar_crossnorm_treshold");
    p->viewray_plane_orthoganality_treshold = yaml_double(kv, "viewray_plane_orthoganality_treshold");
    p->set_all_depths_to_zero = yaml_int(kv, "set_all_depths_to_zero") != 0;
    return MLD_OK;
}

int mld_yaml_int(const char* path, const char* key, int32_t* out, int32_t* found) {
    if (!path || !key || !out) return MLD_ERR_INVALID_ARG;
    std::map<std::string, std::string> kv;
    if (!parse_yaml(path, kv)) return fail(nullptr, MLD_ERR_IO, std::string("Cant find settings file: ") + path);
    if (found) *found = kv.count(key) ? 1 : 0;
    *out = yaml_int(kv, key);  // absent keys read as 0, like cv::FileStorage
    return MLD_OK;
}

const char* mld_status_name(int status) {
    switch (status) {
#define MLD_STATUS_NAME_CASE(name, value) \
    case value: return #name;
        MLD_DEPTH_RESULT_TYPES(MLD_STATUS_NAME_CASE)
#undef MLD_STATUS_NAME_CASE
        default: return "unknown";
    }
}

const char* mld_last_error(const mld_handle* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int mld_create(const mld_params* p, int device, mld_handle** out) {
    if (!p || !out) return fail(nullptr, MLD_ERR_INVALID_ARG, "mld_create: null argument");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(nullptr, MLD_ERR_CUDA,
                    std::string("mld_create: no CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count 0") +
                        "); this library has no CPU fallback");
    if (device < 0) {
        e = cudaGetDevice(&device);
        if (e != cudaSuccess) return fail_cuda(nullptr, e, "cudaGetDevice");
    }
    if (device >= count) return fail(nullptr, MLD_ERR_INVALID_ARG, "mld_create: device index out of range");
    if (p->treshold_depth_enabled && (p->treshold_depth_mode < 0 || p->treshold_depth_mode > 1))
        return fail(nullptr, MLD_ERR_INVALID_ARG, "Undefined treshold depth mode in config");
    if (p->treshold_depth_local_enabled && (p->treshold_depth_local_mode < 0 || p->treshold_depth_local_mode > 1))
        return fail(nullptr, MLD_ERR_INVALID_ARG, "Undefined treshold depth mode in config (local)");
    if (p->treshold_depth_local_enabled && (p->treshold_depth_local_valuetype < 0 || p->treshold_depth_local_valuetype > 1))
        return fail(nullptr, MLD_ERR_INVALID_ARG, "Undefined treshold tolerance mode for tresholdDepthLocal");
    if (p->do_use_histogram_segmentation && !(p->histogram_segmentation_bin_witdh > 1e-6))
        return fail(nullptr, MLD_ERR_INVALID_ARG, "histogram_segmentation_bin_witdh must be > 1e-6");
    mld_handle* h = new mld_handle();
    h->params = *p;
    h->device = device;
    const char* env = getenv("MLD_CHUNK_FRAMES");
    if (env && atoi(env) > 0) h->chunk_frames = atoi(env);
    env = getenv("MLD_FEATURE_MODE");  // "warp" / "split": K2 variants (A/B measurements)
    if (env && strcmp(env, "warp") == 0) h->feature_mode = 0;
    if (env && strcmp(env, "split") == 0) h->feature_mode = 2;
    env = getenv("MLD_TAGGED_MAPS");   // "0": clear the pixel maps before every use instead of epoch tags
    if (env && strcmp(env, "0") == 0) h->use_tagged_maps = false;
    env = getenv("MLD_FUSE");          // "0": separate K1 / gather launches for device-resident non-road sequences too
    if (env) h->fuse_k1_gather = atoi(env) != 0;
    if (env) h->fuse_road = atoi(env) == 1;
    env = getenv("MLD_FUSE_CHUNK");
    if (env && atoi(env) > 0) h->fuse_chunk = atoi(env);
    env = getenv("MLD_OVERLAP_MODE");  // "slots": whole chunks alternate over slot streams; "prio": K1 / K2 priority streams
    if (env && strcmp(env, "slots") == 0) h->overlap_mode = 0;
    if (env && strcmp(env, "prio") == 0) h->overlap_mode = 1;
    env = getenv("MLD_OVERLAP");       // "1": run the chunks of a sequence on one stream (no K1/K2 overlap), up to 3
    if (env && atoi(env) >= 1 && atoi(env) <= MLD_PIPE_SLOTS) h->overlap_slots = atoi(env);
    DeviceGuard g(device);
    if (!g.ok) {
        delete h;
        return fail(nullptr, MLD_ERR_CUDA, "mld_create: cudaSetDevice failed");
    }
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lo = least priority (numerically greatest)
        e = cudaStreamCreateWithPriority(&h->st_lo, cudaStreamNonBlocking, lo);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->st_hi, cudaStreamNonBlocking, hi);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            delete h;
            return fail_cuda(nullptr, e, "priority stream creation");
        }
    }
    e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete h;
        return fail_cuda(nullptr, e, "event creation");
    }
    env = getenv("MLD_SEMANTIC_EXACT");
    if (env) h->semantic_exact = atoi(env) != 0;
    env = getenv("MLD_HOST_PACK");
    if (env) h->host_pack = atoi(env) != 0 ? 1 : 0;
    env = getenv("MLD_PCIE_GBS");
    if (env && atof(env) > 0) h->pcie_gbs = atof(env);
    env = getenv("MLD_PACK_THREADS");
    if (env && atoi(env) > 0) h->host_pack_threads = atoi(env);
    for (int i = 0; i < MLD_PIPE_SLOTS; i++) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        e = cudaStreamCreateWithPriority(&h->slots[i].stream, cudaStreamNonBlocking, lo);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->slots[i].done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->slots[i].ev_k1, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->slots[i].ev_k2, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&h->slots[i].h_ovf_seen), sizeof(int), cudaHostAllocDefault);
        if (e == cudaSuccess) *h->slots[i].h_ovf_seen = 1 << 30;  // unknown yet: launch the full overflow grid
        if (e != cudaSuccess) {
            mld_destroy(h);
            return fail_cuda(nullptr, e, "stream/event creation");
        }
    }
    *out = h;
    return MLD_OK;
}

int mld_destroy(mld_handle* h) {
    if (!h) return MLD_OK;
    DeviceGuard g(h->device);
    for (auto& s : h->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        cudaFree(s.d_pts); cudaFree(s.d_uv); cudaFree(s.d_depth); cudaFree(s.d_status); cudaFree(s.d_maps);
        cudaFree(s.d_bits); cudaFree(s.d_coeffs); cudaFree(s.d_scratch); cudaFree(s.d_small); cudaFree(s.d_ovf); cudaFree(s.d_occ); cudaFree(s.d_split); cudaFree(s.d_labels); cudaFree(s.d_sem); cudaFree(s.d_flags);
        if (s.done) cudaEventDestroy(s.done);
        if (s.ev_k1) cudaEventDestroy(s.ev_k1);
        if (s.ev_k2) cudaEventDestroy(s.ev_k2);
        if (s.h_ovf_seen) cudaFreeHost(s.h_ovf_seen);
        if (s.h_stage) cudaFreeHost(s.h_stage);
        cudaFree(s.d_pack);
        if (s.ev_stage) cudaEventDestroy(s.ev_stage);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    delete h->pool;
    cudaFree(h->d_dbg);
    cudaFree(h->d_hist);
    if (h->h_hist) cudaFreeHost(h->h_hist);
    cudaFree(h->d_synth_tables);
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->st_lo) cudaStreamDestroy(h->st_lo);
    if (h->st_hi) cudaStreamDestroy(h->st_hi);
    delete h;
    return MLD_OK;
}

int mld_initialize(mld_handle* h, int W, int H, double f, double cx, double cy, const double* T) {
    if (!h) return fail(nullptr, MLD_ERR_NOT_CONFIGURED, "Call 'InitConfig' before calling 'Initialize'.");
    if (!T || W <= 0 || H <= 0) return fail(h, MLD_ERR_INVALID_ARG, "mld_initialize: bad camera or transform");
    const mld_params& p = h->params;
    if (p.neighbor_search_mode != 0)
        return fail(h, MLD_ERR_BAD_SEARCH_MODE, "neighbor_search_mode has the invalid value: " + std::to_string(p.neighbor_search_mode));
    DevParams& d = h->dp;
    memset(&d, 0, sizeof(d));
    d.W = W; d.H = H; d.Wd = (double)W; d.Hd = (double)H; d.f = f; d.cx = cx; d.cy = cy;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) d.R[r * 3 + c] = T[r * 4 + c];
        d.t[r] = T[r * 4 + 3];
    }
    // Eigen::Affine3d::inverse(): linear part by cofactors, translation = (-linear^-1) * t
    mld_inverse3_host(d.R, d.Ri);
    for (int r = 0; r < 3; r++)
        d.ti[r] = ((-d.Ri[r * 3 + 0]) * d.t[0] + (-d.Ri[r * 3 + 1]) * d.t[1]) + (-d.Ri[r * 3 + 2]) * d.t[2];
    const double K[9] = {f, 0, cx, 0, f, cy, 0, 0, 1};
    mld_inverse3_host(K, d.Kinv);
    // NeighborFinderPixel::getNeighbors: half = size * 0.5 * (double)(float)scale
    d.hx1 = static_cast<double>(p.pixelarea_search_witdh) * 0.5 * static_cast<double>(1.0f);
    d.hy1 = static_cast<double>(p.pixelarea_search_height) * 0.5 * static_cast<double>(1.0f);
    d.hx2 = static_cast<double>(p.pixelarea_search_witdh) * 0.5 * static_cast<double>(2.0f);
    d.hy2 = static_cast<double>(p.pixelarea_search_height) * 0.5 * static_cast<double>(1.5f);
    d.count_min = p.radiusSearch_count_min;
    d.use_hist = p.do_use_histogram_segmentation != 0;
    d.hist_min = p.histogram_segmentation_min_pointcount;
    d.bin_w = p.histogram_segmentation_bin_witdh;
    d.use_tri_max = p.do_use_triangle_size_maximation != 0;
    d.check_planar = p.do_check_triangleplanar_condition != 0;
    d.crossnorm_thr = p.triangleplanar_crossnorm_treshold;
    d.ortho_thr = p.viewray_plane_orthoganality_treshold;
    d.use_pca = p.do_use_PCA != 0;
    d.pca_3_abs_min = p.pca_treshold_3_abs_min;
    d.pca_3_2_rel_max = p.pca_treshold_3_2_rel_max;
    d.pca_2_1_rel_min = p.pca_treshold_2_1_rel_min;
    d.glob_en = p.treshold_depth_enabled != 0;
    d.glob_mode = p.treshold_depth_mode;
    d.glob_min = (double)p.treshold_depth_min;
    d.glob_max = (double)p.treshold_depth_max;
    d.loc_en = p.treshold_depth_local_enabled != 0;
    d.loc_mode = p.treshold_depth_local_mode;
    d.loc_type = p.treshold_depth_local_valuetype;
    d.loc_val = p.treshold_depth_local_value;
    d.cut_behind = p.do_use_cut_behind_camera != 0;
    d.road_mode = ROAD_NONE;
    if (p.do_use_ransac_plane) {  // estimator priority, DepthEstimator.cpp:84-94
        if (p.plane_estimator_use_triangle_maximation) d.road_mode = ROAD_TRIANGLE;
        else if (p.plane_estimator_use_leastsquares) d.road_mode = ROAD_LEASTSQUARES;
        else if (p.plane_estimator_use_mestimator) d.road_mode = ROAD_MESTIMATOR;
        else return fail(h, MLD_ERR_NO_ROAD_ESTIMATOR, "No road depth estimator selected.");
    }
    d.road_dist_thr = p.ransac_plane_point_distance_treshold;
    d.zx_min_rel = p.plane_estimator_z_x_min_relation;
    d.set_all_zero = p.set_all_depths_to_zero != 0;

    int area = max_window_area(d.hx1, d.hy1, W, H);
    if (d.road_mode != ROAD_NONE) area = std::max(area, max_window_area(d.hx2, d.hy2, W, H));
    h->kcap = mld_feature_capacity_for(area);
    if (h->kcap < 0)
        return fail(h, MLD_ERR_CAPACITY, "search window of " + std::to_string(area) + " pixels exceeds the neighbour capacity " +
                                             std::to_string(mld_neighbor_capacity()));
    mld_setup_prefilter(d);
    DeviceGuard g(h->device);
    CK(mld_configure_feature_depth(h->kcap));
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device) == cudaSuccess && sms > 0) {
        h->overflow_blocks = 2 * sms;
        h->sm_count = sms;
    }
    for (auto& sl : h->slots) sl.epoch = 0;  // image size may have changed
    h->initialized = true;
    h->have_cloud = false;
    h->have_prev = false;
    return MLD_OK;
}

int mld_neighbor_capacity(void) { return 1024; }
int mld_chunk_frames(const mld_handle* h) { return h ? h->chunk_frames : 0; }
int mld_fused_chunk_frames(const mld_handle* h) { return (h && h->fuse_k1_gather && h->feature_mode == 2 && h->overlap_slots >= 2) ? h->fuse_chunk : 0; }

int mld_profile_enable(mld_handle* h, int on) {
    if (!h) return MLD_ERR_INVALID_ARG;
    h->prof_on = on != 0;
    return MLD_OK;
}

int mld_profile_read(mld_handle* h, double* ms7, int64_t* launches7, int64_t* frames_sampled) {
    if (!h || !ms7 || !launches7) return MLD_ERR_INVALID_ARG;
    DeviceGuard g(h->device);
    for (int q = 0; q < 7; q++) {
        ms7[q] = 0.0;
        launches7[q] = 0;
    }
    int64_t frames = 0;
    for (size_t c = 0; c < h->prof_used; c++) {
        cudaEvent_t* ev = &h->prof_events[c * MLD_PROF_EVENTS];
        CK(cudaEventSynchronize(ev[5]));
        // clear, K1, K4, K2 (all of it), K2 gather, K2 solve, K2 rest (road kernels + overflow pass)
        const int from[7] = {0, 1, 3, 4, 4, 6, 7}, to[7] = {1, 2, 4, 5, 6, 7, 5};
        for (int q = 0; q < 7; q++) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ev[from[q]], ev[to[q]]));
            ms7[q] += (double)ms;
        }
        for (int q = 0; q < 7; q++) launches7[q] += (q == 2) ? h->prof_ransac_launches[c] : 1;
        frames += h->prof_frames[c];
    }
    if (frames_sampled) *frames_sampled = frames;
    h->prof_used = 0;
    return MLD_OK;
}
int64_t mld_kernel_launch_count(const mld_handle* h) { return h ? h->launches : 0; }
int mld_host_pipeline_stats(const mld_handle* h, int64_t* out4) {
    if (!h || !out4) return MLD_ERR_INVALID_ARG;
    for (int i = 0; i < 4; i++) out4[i] = h->host_stats[i];
    return MLD_OK;
}

static int check_stride(mld_handle* h, int stride_bytes) {
    if (stride_bytes < 16 || stride_bytes % 16 != 0)
        return fail(h, MLD_ERR_INVALID_ARG, "point stride must be a positive multiple of 16 bytes (float4 / pcl::PointXYZI)");
    return MLD_OK;
}

static int ensure_pool(mld_handle* h) {
    if (!h->pool) {
        int t = h->host_pack_threads;
        if (t <= 0) t = std::max(1, std::min(14, (int)std::thread::hardware_concurrency() - 2));
        h->pool = new HostPool(t - 1);  // the calling thread is the t-th worker
    }
    return MLD_OK;
}

// One cloud of the per-call entry points (setInputCloud and friends) from host memory into the slot's point buffer. A cudaMemcpy
// out of PAGEABLE memory -- what a pcl::PointCloud is -- is staged by the driver at ~12 GB/s: 0.16 ms for 120 000 float4 points,
// 0.32 ms for 32-byte PointXYZI records, most of a frame's latency. Here the host workers strip such a cloud to 12-byte xyz
// straight out of the caller's buffer into the slot's pinned staging buffer (the same squeeze the batched host pipeline uses),
// one pinned copy moves 12 bytes per point and a kernel expands them to the float4 layout. Pinned (registered) sources and small
// clouds are copied as they are. *stride_f_out = floats per point of the device copy.
static int upload_cloud(mld_handle* h, Slot& s, const void* pts, int64_t n, int stride_bytes, int* stride_f_out, bool* consumed = nullptr) {
    *stride_f_out = stride_bytes / 4;
    if (consumed) *consumed = false;  // true: the caller's buffer has been read completely when this returns
    if (n <= 0) return MLD_OK;
    bool pack = h->host_pack != 0 && n >= 16384;
    if (pack && h->host_pack != 1) {
        cudaPointerAttributes at;
        const cudaError_t e = cudaPointerGetAttributes(&at, pts);
        if (e != cudaSuccess) (void)cudaGetLastError();
        pack = e != cudaSuccess || at.type == cudaMemoryTypeUnregistered;
    }
    if (!pack) {
        CK(cudaMemcpyAsync(s.d_pts, pts, (size_t)n * (size_t)stride_bytes, cudaMemcpyHostToDevice, s.stream));
        return MLD_OK;
    }
    int rc = ensure_pool(h);
    if (rc) return rc;
    const size_t need = (size_t)n * 12;
    if (need > s.h_stage_bytes) {
        CK(cudaStreamSynchronize(s.stream));
        if (s.h_stage) CK(cudaFreeHost(s.h_stage));
        s.h_stage = nullptr;
        s.h_stage_bytes = 0;
        CK(cudaHostAlloc(reinterpret_cast<void**>(&s.h_stage), need, cudaHostAllocDefault));
        s.h_stage_bytes = need;
        s.stage_busy = false;
    }
    CK(ensure(s.d_pack, s.d_pack_bytes, need));
    CK(ensure(s.d_pts, s.pts_bytes, (size_t)n * 16));
    if (!s.ev_stage) CK(cudaEventCreateWithFlags(&s.ev_stage, cudaEventDisableTiming));
    if (s.stage_busy) CK(cudaEventSynchronize(s.ev_stage));  // the staging buffer's previous copy has left the host
    const int pieces = 2 * (h->pool->workers() + 1);
    const long long per = ((n + pieces - 1) / pieces + 15) & ~15LL;  // whole groups of 16 records: the squeeze works on cache lines
    const unsigned char* src = reinterpret_cast<const unsigned char*>(pts);
    unsigned char* stage = s.h_stage;
    const std::function<void(int)> job = [&](int pc) {
        const long long lo = (long long)pc * per, hi = std::min<long long>(n, lo + per);
        if (lo < hi) mld_host_pack_xyz(src + (size_t)lo * (size_t)stride_bytes, stride_bytes, reinterpret_cast<float*>(stage + (size_t)lo * 12), hi - lo, 0);
    };
    h->pool->run(pieces, job);
    CK(cudaMemcpyAsync(s.d_pack, s.h_stage, need, cudaMemcpyHostToDevice, s.stream));
    CK(cudaEventRecord(s.ev_stage, s.stream));
    s.stage_busy = true;
    CK(mld_launch_unpack_xyz(s.d_pack, reinterpret_cast<float*>(s.d_pts), n, s.stream));
    h->launches++;
    *stride_f_out = 4;
    if (consumed) *consumed = true;
    return MLD_OK;
}

static void bits_to_plane(const std::vector<unsigned int>& bits, long long n, mld_plane* pl) {
    int64_t cnt = 0;
    for (long long i = 0; i < n; i++)
        if ((bits[(size_t)(i >> 5)] >> (i & 31)) & 1u) {
            if (pl->inlier_idx && cnt < pl->inlier_capacity) pl->inlier_idx[cnt] = (int32_t)i;
            cnt++;
        }
    pl->n_inliers = cnt;
}

static int run_ransac_single(mld_handle* h, Slot& s, long long n, int stride_f, uint64_t seed, mld_plane* pl, int32_t* iters) {
    const long long words = (n + 31) / 32;
    int rc = slot_reserve(h, s, n, stride_f * 4, 0, 1, false, true);
    if (rc) return rc;
    int nl = 0;
    CK(mld_launch_ransac(ransac_config(h->params), reinterpret_cast<const float*>(s.d_pts), stride_f, n, n, 1, seed, 0, s.d_scratch,
                         s.d_coeffs, s.d_bits, words, s.d_small, s.d_small + 1, s.d_small + 2, s.stream, &nl));
    h->launches += nl;
    std::vector<unsigned int> bits((size_t)words);
    int small[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(pl->coeffs, s.d_coeffs, 4 * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(bits.data(), s.d_bits, (size_t)words * sizeof(unsigned int), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(small, s.d_small, 3 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    if (iters) *iters = small[1];
    if (small[2] == MLD_ERR_PCL_INVALID) return fail(h, MLD_ERR_PCL_INVALID, "In GroundPlane: Input pointcloud is invalid");
    if (small[2] != 0) return fail(h, MLD_ERR_NO_MODEL, "RANSAC found no plane model");
    bits_to_plane(bits, n, pl);
    pl->segmented = 1;
    return MLD_OK;
}

int mld_set_cloud(mld_handle* h, const void* points_host, int64_t n, int stride_bytes, mld_plane* inout_plane, uint64_t ransac_seed) {
    if (!h) return MLD_ERR_INVALID_ARG;
    if (!h->initialized) return fail(h, MLD_ERR_NOT_INITIALIZED, "call of 'setInputCloud' without 'initialize'");
    if (n < 0 || (n > 0 && !points_host)) return fail(h, MLD_ERR_INVALID_ARG, "mld_set_cloud: bad cloud");
    int rc = check_stride(h, stride_bytes);
    if (rc) return rc;
    DeviceGuard g(h->device);
    Slot& s = h->slots[0];
    const bool want_ransac = h->params.do_use_ransac_plane && inout_plane && !inout_plane->segmented;
    // the reference throws ExceptionPclInvalid before doing anything else with the plane (RansacPlane.cpp:44-50)
    if (want_ransac && n < 3) return fail(h, MLD_ERR_PCL_INVALID, "In GroundPlane: Input pointcloud is invalid");
    rc = slot_reserve(h, s, std::max<int64_t>(n, 1), stride_bytes, 0, 1, true, false);
    if (rc) return rc;
    int sf = stride_bytes / 4;  // floats per point of the device copy (4 when the host workers packed the cloud)
    bool consumed = false;
    rc = upload_cloud(h, s, points_host, n, stride_bytes, &sf, &consumed);
    if (rc) return rc;
    MapCode mc;
    rc = begin_maps(h, s, 1, n, s.stream, mc);
    if (rc) return rc;
    CK(mld_launch_project_scatter(h->dp, mc, reinterpret_cast<const float*>(s.d_pts), sf, n, n, s.d_maps,
                                  h->feature_mode >= 1 ? s.d_occ : nullptr, 1, s.stream));
    if (n > 0) h->launches++;
    h->cur_n = n;
    h->cur_stride_f = sf;
    h->have_cloud = true;
    if (want_ransac) {
        rc = run_ransac_single(h, s, n, sf, ransac_seed, inout_plane, nullptr);
        if (rc) return rc;
    }
    // A packed upload has read the caller's buffer completely: the copy out of the staging buffer, the expansion and the projection
    // go on behind the caller's back and every later call on this handle is ordered behind them on the slot's stream (the
    // reference's setInputCloud reports nothing either). Otherwise the source may still be in use by the copy engine: wait.
    if (!consumed) CK(cudaStreamSynchronize(s.stream));
    h->cloud_in_flight = consumed;
    return MLD_OK;
}

int mld_estimate_ground_plane(mld_handle* h, const void* points_host, int64_t n, int stride_bytes, uint64_t seed,
                              mld_plane* out_plane, int32_t* iterations_out) {
    if (!h || !out_plane) return MLD_ERR_INVALID_ARG;
    if (n < 0 || (n > 0 && !points_host)) return fail(h, MLD_ERR_INVALID_ARG, "mld_estimate_ground_plane: bad cloud");
    int rc = check_stride(h, stride_bytes);
    if (rc) return rc;
    if (n < 3) return fail(h, MLD_ERR_PCL_INVALID, "In GroundPlane: Input pointcloud is invalid");
    DeviceGuard g(h->device);
    Slot& s = h->slots[1];  // does not disturb the current cloud of slot 0
    CK(ensure(s.d_pts, s.pts_bytes, (size_t)n * (size_t)stride_bytes));
    int sf = stride_bytes / 4;
    rc = upload_cloud(h, s, points_host, n, stride_bytes, &sf);
    if (rc) return rc;
    return run_ransac_single(h, s, n, sf, seed, out_plane, iterations_out);
}

int mld_semantic_ground_plane_device(mld_handle* h, const void* d_points, int64_t n_points, int64_t frame_pitch_points,
                                     int stride_bytes, const uint8_t* d_labels, int label_w, int label_h, double f, double cu,
                                     double cv, const double* T_cam_lidar, const int32_t* ground_labels, int n_ground_labels,
                                     double inlier_threshold, int64_t nframes, float* d_coeffs_out, uint32_t* d_inlier_bits_out,
                                     int32_t* d_n_inliers_out, int32_t* d_rc_out, void* stream) {
    if (!h) return MLD_ERR_INVALID_ARG;
    if (nframes < 0 || n_points < 0 || label_w <= 0 || label_h <= 0 || !T_cam_lidar || (n_ground_labels > 0 && !ground_labels) ||
        (nframes > 0 && (!d_labels || !d_coeffs_out || !d_inlier_bits_out || !d_n_inliers_out || !d_rc_out || (n_points > 0 && !d_points))))
        return fail(h, MLD_ERR_INVALID_ARG, "mld_semantic_ground_plane_device: bad arguments");
    int rc = check_stride(h, stride_bytes);
    if (rc) return rc;
    if (nframes == 0) return MLD_OK;
    DeviceGuard g(h->device);
    Slot& s = h->slots[1];
    CK(ensure(s.d_sem, s.sem_bytes, mld_semantic_state_bytes((int)nframes)));
    unsigned int set8[8];
    ground_label_set(ground_labels, n_ground_labels, set8);
    int nl = 0;
    unsigned char* fl = nullptr;
    rc = sem_flags(h, s, n_points, nframes, &fl);
    if (rc) return rc;
    CK(mld_launch_semantic_plane(T_cam_lidar, f, cu, cv, label_w, label_h, set8, inlier_threshold, reinterpret_cast<const float*>(d_points),
                                 stride_bytes / 4, n_points, frame_pitch_points, d_labels, (int)nframes, s.d_sem, d_coeffs_out,
                                 d_inlier_bits_out, (n_points + 31) / 32, d_n_inliers_out, d_rc_out, reinterpret_cast<cudaStream_t>(stream), &nl, fl,
                                 fl ? 1 : 0));
    h->launches += nl;
    return MLD_OK;
}

int mld_semantic_ground_plane(mld_handle* h, const void* points_host, int64_t n, int stride_bytes, const uint8_t* labels_host,
                              int label_w, int label_h, double f, double cu, double cv, const double* T_cam_lidar,
                              const int32_t* ground_labels, int n_ground_labels, double inlier_threshold, mld_plane* out_plane) {
    if (!h || !out_plane) return MLD_ERR_INVALID_ARG;
    if (n < 0 || (n > 0 && !points_host) || !labels_host || label_w <= 0 || label_h <= 0)
        return fail(h, MLD_ERR_INVALID_ARG, "mld_semantic_ground_plane: bad arguments");
    int rc = check_stride(h, stride_bytes);
    if (rc) return rc;
    DeviceGuard g(h->device);
    Slot& s = h->slots[1];  // does not disturb the current cloud of slot 0
    const long long words = (n + 31) / 32;
    CK(ensure(s.d_pts, s.pts_bytes, (size_t)std::max<int64_t>(n, 1) * (size_t)stride_bytes));
    CK(ensure(s.d_labels, s.labels_bytes, (size_t)label_w * (size_t)label_h));
    CK(ensure(s.d_bits, s.bits_bytes, (size_t)std::max<long long>(words, 1) * sizeof(unsigned int)));
    CK(ensure(s.d_coeffs, s.coeffs_bytes, 4 * sizeof(float)));
    CK(ensure(s.d_small, s.small_bytes, 3 * sizeof(int)));
    int sf = stride_bytes / 4;
    rc = upload_cloud(h, s, points_host, n, stride_bytes, &sf);
    if (rc) return rc;
    CK(cudaMemcpyAsync(s.d_labels, labels_host, (size_t)label_w * (size_t)label_h, cudaMemcpyHostToDevice, s.stream));
    rc = mld_semantic_ground_plane_device(h, s.d_pts, n, n, sf * 4, s.d_labels, label_w, label_h, f, cu, cv, T_cam_lidar,
                                          ground_labels, n_ground_labels, inlier_threshold, 1, s.d_coeffs, s.d_bits, s.d_small,
                                          s.d_small + 2, s.stream);
    if (rc) return rc;
    std::vector<unsigned int> bits((size_t)words);
    int small[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(out_plane->coeffs, s.d_coeffs, 4 * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    if (words > 0) CK(cudaMemcpyAsync(bits.data(), s.d_bits, (size_t)words * sizeof(unsigned int), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(small, s.d_small, 3 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    if (small[2] == MLD_ERR_PCL_INVALID) return fail(h, MLD_ERR_PCL_INVALID, "In GroundPlane: Input pointcloud is invalid");
    bits_to_plane(bits, n, out_plane);
    out_plane->segmented = 1;
    return MLD_OK;
}

int mld_semantic_ground_labelled(mld_handle* h, const void* points_host, int64_t n, int stride_bytes, const uint8_t* labels_host,
                                 int label_w, int label_h, double f, double cu, double cv, const double* T_cam_lidar,
                                 const int32_t* ground_labels, int n_ground_labels, uint8_t* out_flags_host) {
    if (!h || !out_flags_host) return MLD_ERR_INVALID_ARG;
    if (n < 0 || (n > 0 && !points_host) || !labels_host || label_w <= 0 || label_h <= 0 || !T_cam_lidar)
        return fail(h, MLD_ERR_INVALID_ARG, "mld_semantic_ground_labelled: bad arguments");
    int rc = check_stride(h, stride_bytes);
    if (rc) return rc;
    if (n == 0) return MLD_OK;
    DeviceGuard g(h->device);
    Slot& s = h->slots[1];
    const long long words = (n + 31) / 32;
    CK(ensure(s.d_pts, s.pts_bytes, (size_t)n * (size_t)stride_bytes));
    CK(ensure(s.d_labels, s.labels_bytes, (size_t)label_w * (size_t)label_h));
    CK(ensure(s.d_bits, s.bits_bytes, (size_t)words * sizeof(unsigned int)));
    CK(ensure(s.d_coeffs, s.coeffs_bytes, 4 * sizeof(float)));
    CK(ensure(s.d_small, s.small_bytes, 3 * sizeof(int)));
    CK(ensure(s.d_sem, s.sem_bytes, mld_semantic_state_bytes(1)));
    unsigned char* d_flags = nullptr;
    CK(cudaMalloc(&d_flags, (size_t)n));
    unsigned int set8[8];
    ground_label_set(ground_labels, n_ground_labels, set8);
    int nl = 0;
    cudaError_t e = cudaMemcpyAsync(s.d_pts, points_host, (size_t)n * (size_t)stride_bytes, cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.d_labels, labels_host, (size_t)label_w * (size_t)label_h, cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_flags, 0, (size_t)n, s.stream);
    if (e == cudaSuccess)
        e = mld_launch_semantic_plane(T_cam_lidar, f, cu, cv, label_w, label_h, set8, 0.0, reinterpret_cast<const float*>(s.d_pts), stride_bytes / 4, n,
                                      n, s.d_labels, 1, s.d_sem, s.d_coeffs, s.d_bits, words, s.d_small, s.d_small + 2, s.stream, &nl, d_flags);
    h->launches += nl;
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_flags_host, d_flags, (size_t)n, cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    cudaFree(d_flags);
    if (e != cudaSuccess) return fail_cuda(h, e, "mld_semantic_ground_labelled");
    return MLD_OK;
}

int mld_calculate_depth(mld_handle* h, const double* uv_host, int F, double* depth_host, int32_t* status_host,
                        const mld_plane* plane) {
    if (!h) return MLD_ERR_INVALID_ARG;
    if (!h->have_cloud) return fail(h, MLD_ERR_NO_CLOUD, "call of 'CalculateDepth' without 'SetInputCloud'");
    if (F < 0 || (F > 0 && (!uv_host || !depth_host || !status_host))) return fail(h, MLD_ERR_INVALID_ARG, "mld_calculate_depth: bad buffers");
    if (h->params.do_use_depth_segmentation && !h->params.set_all_depths_to_zero)
        return fail(h, MLD_ERR_REGION_GROWING, "DepthEstimator: Region growing not supported!");
    if (F == 0) return MLD_OK;
    DeviceGuard g(h->device);
    Slot& s = h->slots[0];
    CK(ensure(s.d_uv, s.uv_bytes, (size_t)F * 2 * sizeof(double)));
    CK(ensure(s.d_depth, s.depth_bytes, (size_t)F * sizeof(double)));
    CK(ensure(s.d_status, s.status_bytes, (size_t)F * sizeof(int)));
    const long long n = h->cur_n;
    const long long words = (n + 31) / 32;
    const float* coeffs = nullptr;
    const unsigned int* bits = nullptr;
    std::vector<unsigned int> hb;
    int rc = MLD_OK;
    if (plane && h->dp.road_mode != ROAD_NONE) {
        CK(ensure(s.d_bits, s.bits_bytes, (size_t)std::max<long long>(words, 1) * sizeof(unsigned int)));
        CK(ensure(s.d_coeffs, s.coeffs_bytes, 4 * sizeof(float)));
        hb.assign((size_t)std::max<long long>(words, 1), 0u);
        for (int64_t i = 0; i < plane->n_inliers; i++) {
            int32_t r = plane->inlier_idx[i];
            if (r >= 0 && r < n) hb[(size_t)(r >> 5)] |= 1u << (r & 31);
        }
        CK(cudaMemcpyAsync(s.d_bits, hb.data(), hb.size() * sizeof(unsigned int), cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.d_coeffs, plane->coeffs, 4 * sizeof(float), cudaMemcpyHostToDevice, s.stream));
        coeffs = s.d_coeffs;
        bits = s.d_bits;
    }
    CK(cudaMemcpyAsync(s.d_uv, uv_host, (size_t)F * 2 * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    CK(ensure(s.d_ovf, s.ovf_bytes, ((size_t)F + 1) * sizeof(int)));
    rc = launch_features(h, s, s.mc, s.stream, reinterpret_cast<const float*>(s.d_pts), h->cur_stride_f, n, s.d_uv, F, s.d_depth,
                         s.d_status, coeffs, bits, words, 1);
    if (rc) return rc;
    CK(cudaMemcpyAsync(depth_host, s.d_depth, (size_t)F * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(status_host, s.d_status, (size_t)F * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    if (h->stats_on) {  // DepthCalculationStatistics: counters of this call, from the status array that is still on the device
        if (!h->d_hist) CK(cudaMalloc(&h->d_hist, 21 * sizeof(unsigned long long)));
        if (!h->h_hist) CK(cudaHostAlloc(reinterpret_cast<void**>(&h->h_hist), 21 * sizeof(unsigned long long), cudaHostAllocDefault));
        CK(mld_launch_status_histogram(s.d_status, F, h->d_hist, s.stream));
        h->launches++;
        CK(cudaMemcpyAsync(h->h_hist, h->d_hist, 21 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
    }
    CK(cudaStreamSynchronize(s.stream));
    h->cloud_in_flight = false;
    if (h->stats_on) {
        for (int i = 0; i < 21; i++) h->last_hist[i] = (int64_t)h->h_hist[i];
        h->last_hist_valid = true;
    }
    return MLD_OK;
}

int mld_set_semantic_exact(mld_handle* h, int on) {
    if (!h) return MLD_ERR_INVALID_ARG;
    h->semantic_exact = on != 0;
    return MLD_OK;
}

int mld_set_statistics(mld_handle* h, int on) {
    if (!h) return MLD_ERR_INVALID_ARG;
    h->stats_on = on != 0;
    return MLD_OK;
}

int mld_last_status_histogram(mld_handle* h, int64_t* hist21_out) {
    if (!h || !hist21_out) return MLD_ERR_INVALID_ARG;
    for (int i = 0; i < 21; i++) hist21_out[i] = h->last_hist_valid ? h->last_hist[i] : 0;
    return MLD_OK;
}

// debug view: the max-spanning-triangle corners CalculateDepthSegmented used per feature (DepthEstimator.cpp:915-926), camera frame
int mld_get_triangle_corners(mld_handle* h, const double* uv_host, int F, double* corners_out_host, uint8_t* valid_out_host) {
    if (!h || F < 0 || (F > 0 && (!uv_host || !corners_out_host || !valid_out_host))) return MLD_ERR_INVALID_ARG;
    if (!h->have_cloud) return fail(h, MLD_ERR_NO_CLOUD, "call of 'CalculateDepth' without 'SetInputCloud'");
    if (F == 0) return MLD_OK;
    DeviceGuard g(h->device);
    Slot& s = h->slots[0];
    CK(ensure(s.d_uv, s.uv_bytes, (size_t)F * 2 * sizeof(double)));
    CK(ensure(s.d_depth, s.depth_bytes, (size_t)F * sizeof(double)));
    CK(ensure(s.d_status, s.status_bytes, (size_t)F * sizeof(int)));
    double* d_corners = nullptr;
    CK(cudaMalloc(&d_corners, (size_t)F * 9 * sizeof(double)));
    cudaError_t e = cudaMemsetAsync(d_corners, 0xFF, (size_t)F * 9 * sizeof(double), s.stream);  // all-ones = NaN: "no triangle"
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.d_uv, uv_host, (size_t)F * 2 * sizeof(double), cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess)
        e = mld_launch_feature_depth(h->dp, s.mc, h->kcap, reinterpret_cast<const float*>(s.d_pts), h->cur_stride_f, h->cur_n, s.d_maps, s.d_uv, F,
                                     s.d_depth, s.d_status, nullptr, nullptr, 0, 1, nullptr, nullptr, 0, s.stream, d_corners);
    h->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(corners_out_host, d_corners, (size_t)F * 9 * sizeof(double), cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    cudaFree(d_corners);
    if (e != cudaSuccess) return fail_cuda(h, e, "mld_get_triangle_corners");
    for (int i = 0; i < F; i++) valid_out_host[i] = (corners_out_host[(size_t)i * 9] == corners_out_host[(size_t)i * 9]) ? 1 : 0;
    return MLD_OK;
}

// Device-resident non-road sequence with K1 of chunk j and the gather of chunk j-1 in ONE launch on the front stream
// (slot 0's ev_k1/ev_k2 streams are not involved): front stream = clears + fused launches in chunk order; the solve and the
// overflow pass of a chunk run on its slot's stream under the next fused launch. A slot (maps, occupancy, survivor lists)
// is reused by chunk j once chunk j - nslots has finished its overflow pass (which still reads the maps).
static int process_frames_device_fused(mld_handle* h, const float* pts, int64_t n_points, int64_t frame_pitch_points, int stride_f,
                                       const double* d_uv, int F, double* d_depth, int32_t* d_status, int64_t nframes, int chunk,
                                       cudaStream_t st, bool use_road, uint64_t seed, float* d_plane_coeffs_out, const PlaneSrc* src) {
    const int nslots = h->overlap_slots < 2 ? 2 : h->overlap_slots;
    const int64_t nchunks = (nframes + chunk - 1) / chunk;
    cudaStream_t front = h->st_lo;
    CK(cudaEventRecord(h->ev_fork, st));
    CK(cudaStreamWaitEvent(front, h->ev_fork, 0));
    for (int i = 0; i < nslots; i++) CK(cudaStreamWaitEvent(h->slots[i].stream, h->ev_fork, 0));
    for (int i = 0; i < nslots; i++)
        CK(ensure(h->slots[i].d_split, h->slots[i].split_bytes, mld_split_scratch_bytes((long long)chunk * F, use_road ? 1 : 0)));
    const long long words = (n_points + 31) / 32;
    for (int64_t j = 0; j <= nchunks; j++) {
        const bool have_k1 = j < nchunks, have_g = j >= 1;
        Slot* sk = have_k1 ? &h->slots[j % nslots] : nullptr;
        Slot* sg = have_g ? &h->slots[(j - 1) % nslots] : nullptr;
        const int64_t f0k = j * chunk, f0g = (j - 1) * chunk;
        const int ck = have_k1 ? (int)std::min<int64_t>(chunk, nframes - f0k) : 0;
        const int cg = have_g ? (int)std::min<int64_t>(chunk, nframes - f0g) : 0;
        MapCode mck{0u, 0u};
        // profiling brackets of this launch group: [0,1] clears, [1,2] the fused launch (reported as the K1 class), [6,7] the
        // solve, [7,5] the overflow pass; the gather has no bracket of its own ([4,6] is empty)
        // (every 4th full launch group is sampled: the timing events cost ~4 % of the step when every group carries them)
        cudaEvent_t* ev = nullptr;
        if (have_k1 && have_g && (j & 3) == 1) {
            int rcp = prof_acquire(h, ck, &ev);
            if (rcp) return rcp;
        }
        if (have_k1 && j >= nslots) CK(cudaStreamWaitEvent(front, sk->done, 0));  // chunk j - nslots has left the slot
        if (ev) CK(cudaEventRecord(ev[0], front));
        if (have_k1) {
            int rcm = begin_maps(h, *sk, ck, n_points, front, mck, true);
            if (rcm) return rcm;
            sk->mc = mck;
        }
        if (have_g) CK(cudaMemsetAsync(sg->d_ovf, 0, sizeof(int), front));
        if (ev) CK(cudaEventRecord(ev[1], front));
        if (have_k1 && use_road && !(src && src->kind == PlaneSrc::EXTERNAL)) {
            // the ground plane of chunk j only needs the points: fitted on the slot's stream (behind chunk j - nslots, whose road
            // kernels were the last readers of the slot's plane buffers) while the fused launches go on
            float* cdst = d_plane_coeffs_out ? d_plane_coeffs_out + f0k * 4 : sk->d_coeffs;
            const float* cp = pts + f0k * frame_pitch_points * stride_f;
            int nlp = 0;
            if (src && src->kind == PlaneSrc::SEMANTIC) {
                CK(ensure(sk->d_sem, sk->sem_bytes, mld_semantic_state_bytes(ck)));
                unsigned char* fl = nullptr;
                int rcs = sem_flags(h, *sk, n_points, ck, &fl);
                if (rcs) return rcs;
                CK(mld_launch_semantic_plane(src->T, src->f, src->cu, src->cv, src->label_w, src->label_h, src->ground, src->inlier_threshold, cp,
                                             stride_f, n_points, frame_pitch_points, src->d_labels + f0k * (int64_t)src->label_w * src->label_h, ck,
                                             sk->d_sem, cdst, sk->d_bits, words, sk->d_small,
                                             src->d_rc_out ? src->d_rc_out + f0k : sk->d_small + 2 * ck, sk->stream, &nlp, fl, fl ? 1 : 0));
            } else {
                CK(mld_launch_ransac(ransac_config(h->params), cp, stride_f, n_points, frame_pitch_points, ck, seed, f0k, sk->d_scratch, cdst,
                                     sk->d_bits, words, sk->d_small, sk->d_small + ck, sk->d_small + 2 * ck, sk->stream, &nlp));
            }
            h->launches += nlp;
        }
        int nl = 0;
        CK(mld_launch_fused_project_gather(h->dp, stride_f, mck, have_k1 ? pts + f0k * frame_pitch_points * stride_f : nullptr, n_points,
                                           frame_pitch_points, have_k1 ? sk->d_maps : nullptr, have_k1 ? sk->d_occ : nullptr, ck,
                                           have_g ? sg->mc : mck, have_g ? pts + f0g * frame_pitch_points * stride_f : nullptr,
                                           have_g ? sg->d_maps : nullptr, have_g ? sg->d_occ : nullptr,
                                           have_g ? d_uv + f0g * (int64_t)F * 2 : nullptr, F, have_g ? d_depth + f0g * (int64_t)F : nullptr,
                                           have_g ? d_status + f0g * (int64_t)F : nullptr, cg, have_g ? sg->d_ovf + 1 : nullptr,
                                           have_g ? sg->d_ovf : nullptr, have_g ? sg->d_split : nullptr, front, &nl));
        h->launches += nl;
        if (ev) CK(cudaEventRecord(ev[2], front));
        cudaStream_t s2 = have_g ? sg->stream : front;
        if (have_g) {
            CK(cudaEventRecord(sg->ev_k1, front));
            CK(cudaStreamWaitEvent(sg->stream, sg->ev_k1, 0));
        }
        if (ev) {
            CK(cudaEventRecord(ev[3], s2));
            CK(cudaEventRecord(ev[4], s2));
            CK(cudaEventRecord(ev[6], s2));
        }
        if (have_g) {
            int nl2 = 0;
            const float* gp = pts + f0g * frame_pitch_points * stride_f;
            const float* coeffs = nullptr;
            const unsigned int* bits = nullptr;
            if (use_road) {
                if (src && src->kind == PlaneSrc::EXTERNAL) {
                    coeffs = src->d_coeffs + f0g * 4;
                    bits = src->d_bits + f0g * words;
                } else {
                    coeffs = d_plane_coeffs_out ? d_plane_coeffs_out + f0g * 4 : sg->d_coeffs;
                    bits = sg->d_bits;
                }
            }
            CK(mld_launch_feature_solve(h->dp, sg->mc, gp, stride_f, frame_pitch_points, sg->d_maps, sg->d_occ, d_uv + f0g * (int64_t)F * 2, F,
                                        d_depth + f0g * (int64_t)F, d_status + f0g * (int64_t)F, coeffs, bits, words, cg, sg->d_ovf + 1,
                                        sg->d_ovf, sg->d_split, s2, &nl2, ev ? &ev[7] : nullptr));
            CK(mld_launch_feature_depth(h->dp, sg->mc, h->kcap, gp, stride_f, frame_pitch_points, sg->d_maps, d_uv + f0g * (int64_t)F * 2, F,
                                        d_depth + f0g * (int64_t)F, d_status + f0g * (int64_t)F, coeffs, bits, words, cg, sg->d_ovf + 1,
                                        sg->d_ovf, overflow_grid(h, *sg), s2));
            CK(cudaMemcpyAsync(sg->h_ovf_seen, sg->d_ovf, sizeof(int), cudaMemcpyDeviceToHost, s2));
            if (ev) CK(cudaEventRecord(ev[5], s2));
            {  // the overflow pass was the slot's last reader: clear its occupancy bitmaps for the next chunk here
                const size_t ob = (size_t)chunk * (size_t)occ_words_per_frame(h->dp.W, h->dp.H) * sizeof(unsigned int);
                CK(cudaMemsetAsync(sg->d_occ, 0, std::min(ob, sg->occ_bytes), s2));
                sg->occ_clean_bytes = std::min(ob, sg->occ_bytes);
            }
            CK(cudaEventRecord(sg->done, s2));
            h->launches += nl2 + 1;
        } else if (ev) {
            CK(cudaEventRecord(ev[7], s2));
            CK(cudaEventRecord(ev[5], s2));
        }
        if (ev) h->prof_used++;
    }
    for (int i = 0; i < nslots; i++) CK(cudaStreamWaitEvent(st, h->slots[i].done, 0));
    CK(cudaEventRecord(h->ev_join, front));
    CK(cudaStreamWaitEvent(st, h->ev_join, 0));
    h->have_cloud = false;
    return MLD_OK;
}

static int process_frames_device_impl(mld_handle* h, const void* d_points, int64_t n_points, int64_t frame_pitch_points, int stride_bytes,
                                      const double* d_uv, int F, double* d_depth, int32_t* d_status, int64_t nframes, int road,
                                      uint64_t seed, float* d_plane_coeffs_out, void* stream, const PlaneSrc* src) {
    if (!h) return MLD_ERR_INVALID_ARG;
    if (!h->initialized) return fail(h, MLD_ERR_NOT_INITIALIZED, "call of 'setInputCloud' without 'initialize'");
    int rc = check_stride(h, stride_bytes);
    if (rc) return rc;
    if (nframes < 0 || n_points < 0 || frame_pitch_points < n_points || F < 0)
        return fail(h, MLD_ERR_INVALID_ARG, "mld_process_frames_device: bad sizes");
    if (nframes > 0 && ((n_points > 0 && !d_points) || (F > 0 && (!d_uv || !d_depth || !d_status))))
        return fail(h, MLD_ERR_INVALID_ARG, "mld_process_frames_device: null buffer");
    if ((reinterpret_cast<uintptr_t>(d_points) & 15u) != 0)
        return fail(h, MLD_ERR_INVALID_ARG, "mld_process_frames_device: points must be 16-byte aligned");
    if (h->params.do_use_depth_segmentation && !h->params.set_all_depths_to_zero)
        return fail(h, MLD_ERR_REGION_GROWING, "DepthEstimator: Region growing not supported!");
    if (road && n_points < 3 && !(src && src->kind == PlaneSrc::EXTERNAL))
        return fail(h, MLD_ERR_PCL_INVALID, "In GroundPlane: Input pointcloud is invalid");
    if (road && h->dp.road_mode == ROAD_NONE && src)
        return fail(h, MLD_ERR_NO_ROAD_ESTIMATOR, "a ground plane was given but do_use_ransac_plane is off: no road depth estimator (DepthEstimator.cpp:84-103)");
    DeviceGuard g(h->device);
    if (h->cloud_in_flight) {  // the batch writes slot 0's maps from other streams: let the per-call projection finish first
        CK(cudaStreamSynchronize(h->slots[0].stream));
        h->cloud_in_flight = false;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // short sequences are cut into at least `overlap_slots` chunks so that the streams still overlap
    const bool use_road = road && h->dp.road_mode != ROAD_NONE;
    // MLD_FUSE=2 restricts the fused pipeline to the non-road path
    const bool fused_ok = h->fuse_k1_gather && (!(use_road || src) || h->fuse_road) && h->feature_mode == 2 && h->overlap_slots >= 2 && F > 0 &&
                          n_points > 0;
    const int chunk = (int)std::max<int64_t>(1, std::min<int64_t>(fused_ok ? h->fuse_chunk : h->chunk_frames,
                                                                  std::max<int64_t>(8, (nframes + h->overlap_slots - 1) / h->overlap_slots)));
    const int64_t nchunks = (nframes + chunk - 1) / chunk;
    // K1 is issue/DRAM bound, K2 latency bound: chunks alternate over `nslots` slots (own maps + stream) so
    // that K1 of one chunk runs under K2 of the previous one. Fork from / join into the caller's stream.
    const int nslots = (int)std::max<int64_t>(1, std::min<int64_t>(h->overlap_slots, nchunks));
    for (int i = 0; i < nslots; i++) {
        rc = slot_reserve(h, h->slots[i], std::max<int64_t>(n_points, 1), stride_bytes, F, chunk, false, use_road);
        if (rc) return rc;
    }
    const int stride_f = stride_bytes / 4;
    const float* pts = reinterpret_cast<const float*>(d_points);
    if (fused_ok && nchunks >= 2 && nslots >= 2)
        return process_frames_device_fused(h, pts, n_points, frame_pitch_points, stride_f, d_uv, F, d_depth, d_status, nframes, chunk, st,
                                           use_road, seed, d_plane_coeffs_out, src);
    const bool prio = nslots > 1 && h->overlap_mode == 1;
    if (nslots > 1) {
        CK(cudaEventRecord(h->ev_fork, st));
        if (prio) {
            CK(cudaStreamWaitEvent(h->st_lo, h->ev_fork, 0));
            CK(cudaStreamWaitEvent(h->st_hi, h->ev_fork, 0));
        } else {
            for (int i = 0; i < nslots; i++) CK(cudaStreamWaitEvent(h->slots[i].stream, h->ev_fork, 0));
        }
    }
    int64_t ci = 0;
    for (int64_t f0 = 0; f0 < nframes; f0 += chunk, ci++) {
        int c = (int)std::min<int64_t>(chunk, nframes - f0);
        Slot& s = h->slots[ci % nslots];
        PlaneSrc cs;
        if (src) {  // the chunk's view of the per-frame plane inputs
            cs = *src;
            if (cs.d_labels) cs.d_labels += f0 * (int64_t)cs.label_w * (int64_t)cs.label_h;
            if (cs.d_rc_out) cs.d_rc_out += f0;
            if (cs.d_coeffs) cs.d_coeffs += f0 * 4;
            if (cs.d_bits) cs.d_bits += f0 * ((n_points + 31) / 32);
        }
        rc = enqueue_chunk(h, s, prio ? h->st_lo : (nslots > 1 ? s.stream : st), pts + f0 * frame_pitch_points * stride_f, n_points,
                           frame_pitch_points, stride_f, d_uv + f0 * (int64_t)F * 2, F, d_depth + f0 * (int64_t)F,
                           d_status + f0 * (int64_t)F, c, use_road, seed, f0, d_plane_coeffs_out ? d_plane_coeffs_out + f0 * 4 : nullptr,
                           prio ? h->st_hi : nullptr, src ? &cs : nullptr);
        if (rc) return rc;
    }
    if (prio) {
        // every K1 precedes a K2 on st_hi, so the end of st_hi is the end of the sequence
        CK(cudaEventRecord(h->ev_join, h->st_hi));
        CK(cudaStreamWaitEvent(st, h->ev_join, 0));
    } else if (nslots > 1) {
        for (int i = 0; i < nslots; i++) {
            CK(cudaEventRecord(h->slots[i].done, h->slots[i].stream));
            CK(cudaStreamWaitEvent(st, h->slots[i].done, 0));
        }
    }
    h->have_cloud = false;  // slot 0's map now belongs to the batch
    return MLD_OK;
}

int mld_process_frames_device(mld_handle* h, const void* d_points, int64_t n_points, int64_t frame_pitch_points, int stride_bytes,
                              const double* d_uv, int F, double* d_depth, int32_t* d_status, int64_t nframes, int road,
                              uint64_t seed, float* d_plane_coeffs_out, void* stream) {
    return process_frames_device_impl(h, d_points, n_points, frame_pitch_points, stride_bytes, d_uv, F, d_depth, d_status, nframes, road, seed,
                                      d_plane_coeffs_out, stream, nullptr);
}

int mld_process_frames_device_semantic(mld_handle* h, const void* d_points, int64_t n_points, int64_t frame_pitch_points, int stride_bytes,
                                       const uint8_t* d_labels, int label_w, int label_h, double f, double cu, double cv,
                                       const double* T_cam_lidar, const int32_t* ground_labels, int n_ground_labels,
                                       double inlier_threshold, const double* d_uv, int F, double* d_depth, int32_t* d_status,
                                       int64_t nframes, float* d_plane_coeffs_out, int32_t* d_plane_rc_out, void* stream) {
    if (!h) return MLD_ERR_INVALID_ARG;
    if (!d_labels || label_w <= 0 || label_h <= 0 || !T_cam_lidar || (n_ground_labels > 0 && !ground_labels))
        return fail(h, MLD_ERR_INVALID_ARG, "mld_process_frames_device_semantic: bad label image / camera");
    PlaneSrc src;
    src.kind = PlaneSrc::SEMANTIC;
    src.d_labels = d_labels;
    src.label_w = label_w;
    src.label_h = label_h;
    src.f = f;
    src.cu = cu;
    src.cv = cv;
    for (int i = 0; i < 12; i++) src.T[i] = T_cam_lidar[i];
    src.inlier_threshold = inlier_threshold;
    ground_label_set(ground_labels, n_ground_labels, src.ground);
    src.d_rc_out = d_plane_rc_out;
    return process_frames_device_impl(h, d_points, n_points, frame_pitch_points, stride_bytes, d_uv, F, d_depth, d_status, nframes, 1, 0,
                                      d_plane_coeffs_out, stream, &src);
}

int mld_process_frames_device_planes(mld_handle* h, const void* d_points, int64_t n_points, int64_t frame_pitch_points, int stride_bytes,
                                     const float* d_plane_coeffs, const uint32_t* d_inlier_bits, const double* d_uv, int F,
                                     double* d_depth, int32_t* d_status, int64_t nframes, void* stream) {
    if (!h) return MLD_ERR_INVALID_ARG;
    if (nframes > 0 && (!d_plane_coeffs || !d_inlier_bits)) return fail(h, MLD_ERR_INVALID_ARG, "mld_process_frames_device_planes: null plane buffers");
    PlaneSrc src;
    src.kind = PlaneSrc::EXTERNAL;
    src.d_coeffs = d_plane_coeffs;
    src.d_bits = d_inlier_bits;
    return process_frames_device_impl(h, d_points, n_points, frame_pitch_points, stride_bytes, d_uv, F, d_depth, d_status, nframes, 1, 0,
                                      nullptr, stream, &src);
}

int mld_process_frames_host(mld_handle* h, const void* points_host, int64_t n_points, int64_t frame_pitch_points, int stride_bytes,
                            const double* uv_host, int F, double* depth_host, int32_t* status_host, int64_t nframes, int road,
                            uint64_t seed, float* plane_coeffs_out_host) {
    if (!h) return MLD_ERR_INVALID_ARG;
    if (!h->initialized) return fail(h, MLD_ERR_NOT_INITIALIZED, "call of 'setInputCloud' without 'initialize'");
    int rc = check_stride(h, stride_bytes);
    if (rc) return rc;
    if (nframes < 0 || n_points < 0 || frame_pitch_points < n_points || F < 0)
        return fail(h, MLD_ERR_INVALID_ARG, "mld_process_frames_host: bad sizes");
    if (nframes > 0 && ((n_points > 0 && !points_host) || (F > 0 && (!uv_host || !depth_host || !status_host))))
        return fail(h, MLD_ERR_INVALID_ARG, "mld_process_frames_host: null buffer");
    if (h->params.do_use_depth_segmentation && !h->params.set_all_depths_to_zero)
        return fail(h, MLD_ERR_REGION_GROWING, "DepthEstimator: Region growing not supported!");
    if (road && n_points < 3) return fail(h, MLD_ERR_PCL_INVALID, "In GroundPlane: Input pointcloud is invalid");
    DeviceGuard g(h->device);
    h->cloud_in_flight = false;  // slot 0's work is ordered on its own stream here
    // H2D copies, kernels and D2H copies of different chunks overlap: chunks are kept small (<= 32 frames, ~1 ms of
    // PCIe time each) and a sequence is cut into at least two chunks per slot
    const int chunk = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(h->chunk_frames, 32), std::max<int64_t>(8, (nframes + 2 * MLD_HOST_SLOTS - 1) / (2 * MLD_HOST_SLOTS))));
    const bool use_road = road && h->dp.road_mode != ROAD_NONE;
    const int stride_f = stride_bytes / 4;
    const size_t frame_bytes = (size_t)n_points * (size_t)stride_bytes;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(points_host);
    for (int i = 0; i < MLD_HOST_SLOTS; i++) {
        rc = slot_reserve(h, h->slots[i], std::max<int64_t>(n_points, 1), stride_bytes, std::max(F, 1), chunk, true, use_road);
        if (rc) return rc;
    }
    // Only x, y, z of a point are used (12 of the 16 / 32 bytes of a record) and the pipeline is PCIe bound: host threads pack
    // the chunk's points to 12-byte xyz in a pinned staging buffer, one contiguous copy moves them, a kernel expands them to the
    // float4 layout the projection kernel streams. (A strided 2-D DMA copy that skips the padding runs at 11 GB/s and a kernel
    // reading mapped pinned memory transfers every byte of the 32-byte sectors anyway: measured, scripts/pcie_probe.cu.)
    // 16-byte float4 records are packed as well where the host squeezes whole cache lines (AVX-512): 12 instead of 16 bytes cross the link
    const bool pack = n_points > 0 && (h->host_pack == 1 || (h->host_pack != 0 && (stride_bytes > 16 || (stride_bytes == 16 && mld_host_pack_level() == 512))));
    if (pack) ensure_pool(h);
    if (pack) {
        for (int i = 0; i < MLD_HOST_SLOTS; i++) {
            Slot& s = h->slots[i];
            const size_t need = (size_t)chunk * (size_t)n_points * 12;
            if (need > s.h_stage_bytes) {
                CK(cudaStreamSynchronize(s.stream));
                if (s.h_stage) CK(cudaFreeHost(s.h_stage));
                s.h_stage = nullptr;
                s.h_stage_bytes = 0;
                CK(cudaHostAlloc(reinterpret_cast<void**>(&s.h_stage), need, cudaHostAllocDefault));
                s.h_stage_bytes = need;
            }
            CK(ensure(s.d_pack, s.d_pack_bytes, need));
            CK(ensure(s.d_pts, s.pts_bytes, (size_t)chunk * (size_t)n_points * 16));
            if (!s.ev_stage) CK(cudaEventCreateWithFlags(&s.ev_stage, cudaEventDisableTiming));
        }
    }
    int64_t ci = 0;
    cudaEvent_t last_h2d = nullptr;  // the most recently queued H2D copy of the cloud
    const auto t_origin = std::chrono::steady_clock::now();
    double link_free_at = 0.0;       // modelled time (s since t_origin) at which the queued H2D copies will have left the host
    double pack_s_per_frame = 0.0;   // measured, smoothed
    const double link_bytes_per_s = h->pcie_gbs * 1e9;
    auto queue_copy = [&](double bytes) {
        const double now = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_origin).count();
        link_free_at = std::max(link_free_at, now) + bytes / link_bytes_per_s;
    };
    for (int i = 0; i < MLD_HOST_SLOTS; i++) h->slots[i].stage_busy = false;
    for (int64_t f0 = 0; f0 < nframes; f0 += chunk, ci++) {
        Slot& s = h->slots[ci % MLD_HOST_SLOTS];
        int c = (int)std::min<int64_t>(chunk, nframes - f0);
        // stream order makes the slot's buffers safe to reuse: the previous chunk on this stream is complete
        // (its D2H copies included) before these copies start
        // Packing costs host time, copying whole records costs PCIe time (measured on a 16-core host: 14.2 k frames/s copying 32-byte
        // records, 20.3 k packing everything with 14 threads). A chunk is packed when the copies already queued keep the link busy until
        // the packed data is ready, and goes out as whole records when the link would otherwise idle: both resources stay busy.
        // The link's backlog is modelled on the host (bytes queued / link rate); it only steers the choice, never correctness.
        bool pack_this = pack;
        if (pack && h->host_pack != 1) {
            const double now = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_origin).count();
            if (last_h2d != nullptr && cudaEventQuery(last_h2d) == cudaSuccess) link_free_at = std::min(link_free_at, now);  // backlog drained
            if (now + pack_s_per_frame * c > link_free_at) pack_this = false;
        }
        if (pack_this) {
            if (s.stage_busy) CK(cudaEventSynchronize(s.ev_stage));  // the staging buffer's previous copy has left the host
            const int pieces = 8;                                          // per frame: load balance over the workers
            const long long per = (n_points + pieces - 1) / pieces;
            unsigned char* stage = s.h_stage;
            const std::function<void(int)> job = [&](int item) {
                const int fr = item / pieces, pc = item % pieces;
                const long long lo = (long long)pc * per, hi = std::min<long long>(n_points, lo + per);
                const unsigned char* p = src + ((size_t)(f0 + fr) * (size_t)frame_pitch_points + (size_t)lo) * (size_t)stride_bytes;
                float* q = reinterpret_cast<float*>(stage + ((size_t)fr * (size_t)n_points + (size_t)lo) * 12);
                mld_host_pack_xyz(p, stride_bytes, q, hi - lo, 0);  // AVX-512 line-at-a-time squeeze where the host has it (mld_host_pack.cpp)
            };
            const auto tp0 = std::chrono::steady_clock::now();
            h->pool->run(c * pieces, job);
            const double took = std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count() / c;
            pack_s_per_frame = pack_s_per_frame > 0 ? 0.7 * pack_s_per_frame + 0.3 * took : took;
            CK(cudaMemcpyAsync(s.d_pack, s.h_stage, (size_t)c * (size_t)n_points * 12, cudaMemcpyHostToDevice, s.stream));
            CK(cudaEventRecord(s.ev_stage, s.stream));
            s.stage_busy = true;
            last_h2d = s.ev_stage;
            h->host_stats[0] += (int64_t)c * n_points * 12;
            h->host_stats[2] += c;
            queue_copy((double)c * (double)n_points * 12.0);
            CK(mld_launch_unpack_xyz(s.d_pack, reinterpret_cast<float*>(s.d_pts), (long long)c * n_points, s.stream));
            h->launches++;
        } else if (n_points > 0) {
            CK(cudaMemcpy2DAsync(s.d_pts, frame_bytes, src + (size_t)f0 * (size_t)frame_pitch_points * (size_t)stride_bytes,
                                 (size_t)frame_pitch_points * (size_t)stride_bytes, frame_bytes, (size_t)c, cudaMemcpyHostToDevice,
                                 s.stream));
            if (pack) {
                CK(cudaEventRecord(s.ev_k1, s.stream));
                last_h2d = s.ev_k1;
            }
            h->host_stats[0] += (int64_t)c * (int64_t)frame_bytes;
            h->host_stats[3] += c;
            queue_copy((double)c * (double)frame_bytes);
        }
        h->host_stats[0] += (int64_t)c * F * 16;
        h->host_stats[1] += (int64_t)c * F * 12;
        if (F > 0)
            CK(cudaMemcpyAsync(s.d_uv, uv_host + f0 * (int64_t)F * 2, (size_t)c * (size_t)F * 2 * sizeof(double),
                               cudaMemcpyHostToDevice, s.stream));
        rc = enqueue_chunk(h, s, s.stream, reinterpret_cast<const float*>(s.d_pts), n_points, n_points, pack_this ? 4 : stride_f, s.d_uv, F, s.d_depth,
                           s.d_status, c, use_road, seed, f0, nullptr);
        if (rc) return rc;
        if (F > 0) {
            CK(cudaMemcpyAsync(depth_host + f0 * (int64_t)F, s.d_depth, (size_t)c * (size_t)F * sizeof(double), cudaMemcpyDeviceToHost,
                               s.stream));
            CK(cudaMemcpyAsync(status_host + f0 * (int64_t)F, s.d_status, (size_t)c * (size_t)F * sizeof(int), cudaMemcpyDeviceToHost,
                               s.stream));
        }
        if (use_road && plane_coeffs_out_host)
            CK(cudaMemcpyAsync(plane_coeffs_out_host + f0 * 4, s.d_coeffs, (size_t)c * 4 * sizeof(float), cudaMemcpyDeviceToHost,
                               s.stream));
    }
    for (int i = 0; i < MLD_HOST_SLOTS; i++) CK(cudaStreamSynchronize(h->slots[i].stream));
    h->have_cloud = false;
    return MLD_OK;
}

// ---- tracklets_depth batch adaptor: previous + current cloud in one call ----
// resident: the slot already holds this cloud (points, pixel map, occupancy) from the call in which it was the current cloud; only
// the features (and the plane) go to the device
static int pair_side_begin(mld_handle* h, Slot& s, const void* pts, int64_t n, int stride_bytes, const double* uv, int F,
                           const mld_plane* plane, bool want_ransac, uint64_t seed, std::vector<unsigned int>& hb, bool resident = false,
                           int* stride_f_out = nullptr) {
    int rc = MLD_OK;
    int stride_f = stride_bytes / 4;  // resident: the stride the cloud was stored with; else set by the upload
    MapCode mc = s.mc;
    if (!resident) {
        rc = slot_reserve(h, s, std::max<int64_t>(n, 1), stride_bytes, std::max(F, 1), 1, true, want_ransac || plane != nullptr);
        if (rc) return rc;
        rc = upload_cloud(h, s, pts, n, stride_bytes, &stride_f);
        if (rc) return rc;
        if (stride_f_out) *stride_f_out = stride_f;
        rc = begin_maps(h, s, 1, n, s.stream, mc);
        if (rc) return rc;
        CK(mld_launch_project_scatter(h->dp, mc, reinterpret_cast<const float*>(s.d_pts), stride_f, n, n, s.d_maps,
                                      h->feature_mode >= 1 ? s.d_occ : nullptr, 1, s.stream));
        if (n > 0) h->launches++;
    } else {
        CK(ensure(s.d_uv, s.uv_bytes, (size_t)std::max(F, 1) * 2 * sizeof(double)));
        CK(ensure(s.d_depth, s.depth_bytes, (size_t)std::max(F, 1) * sizeof(double)));
        CK(ensure(s.d_status, s.status_bytes, (size_t)std::max(F, 1) * sizeof(int)));
        CK(ensure(s.d_ovf, s.ovf_bytes, ((size_t)std::max(F, 1) + 1) * sizeof(int)));
        if (want_ransac || plane != nullptr) {
            const size_t words = (size_t)((n + 31) / 32);
            CK(ensure(s.d_bits, s.bits_bytes, std::max<size_t>(words, 1) * sizeof(unsigned int)));
            CK(ensure(s.d_coeffs, s.coeffs_bytes, 4 * sizeof(float)));
            CK(ensure(s.d_scratch, s.scratch_bytes, mld_ransac_scratch_bytes(std::max<int64_t>(n, 1), 1)));
            CK(ensure(s.d_small, s.small_bytes, 3 * sizeof(int)));
        }
    }
    if (F > 0) CK(cudaMemcpyAsync(s.d_uv, uv, (size_t)F * 2 * sizeof(double), cudaMemcpyHostToDevice, s.stream));
    const long long words = (n + 31) / 32;
    const float* coeffs = nullptr;
    const unsigned int* bits = nullptr;
    if (want_ransac) {
        int nl = 0;
        CK(mld_launch_ransac(ransac_config(h->params), reinterpret_cast<const float*>(s.d_pts), stride_f, n, n, 1, seed, 0, s.d_scratch,
                             s.d_coeffs, s.d_bits, words, s.d_small, s.d_small + 1, s.d_small + 2, s.stream, &nl));
        h->launches += nl;
        coeffs = s.d_coeffs;
        bits = s.d_bits;
    } else if (plane != nullptr && h->dp.road_mode != ROAD_NONE) {
        hb.assign((size_t)std::max<long long>(words, 1), 0u);
        for (int64_t i = 0; i < plane->n_inliers; i++) {
            int32_t r = plane->inlier_idx[i];
            if (r >= 0 && r < n) hb[(size_t)(r >> 5)] |= 1u << (r & 31);
        }
        CK(cudaMemcpyAsync(s.d_bits, hb.data(), hb.size() * sizeof(unsigned int), cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.d_coeffs, plane->coeffs, 4 * sizeof(float), cudaMemcpyHostToDevice, s.stream));
        coeffs = s.d_coeffs;
        bits = s.d_bits;
    }
    if (F > 0) {
        rc = launch_features(h, s, mc, s.stream, reinterpret_cast<const float*>(s.d_pts), stride_f, n, s.d_uv, F, s.d_depth, s.d_status,
                             coeffs, bits, words, 1);
        if (rc) return rc;
    }
    return MLD_OK;
}

static int calculate_depth_pair_impl(mld_handle* h, const void* pts_prev, int64_t n_prev, const double* uv_prev, int F_prev,
                                     double* depth_prev, int32_t* status_prev, mld_plane* plane_prev, const void* pts_cur, int64_t n_cur,
                                     const double* uv_cur, int F_cur, double* depth_cur, int32_t* status_cur, mld_plane* plane_cur,
                                     int stride_bytes, uint64_t ransac_seed, bool prev_resident) {
    if (!h) return MLD_ERR_INVALID_ARG;
    if (!h->initialized) return fail(h, MLD_ERR_NOT_INITIALIZED, "call of 'setInputCloud' without 'initialize'");
    int rc = check_stride(h, stride_bytes);
    if (rc) return rc;
    if ((!prev_resident && n_prev < 0) || n_cur < 0 || F_prev < 0 || F_cur < 0 || !pts_cur || (F_prev > 0 && (!uv_prev || !depth_prev)) ||
        (F_cur > 0 && (!uv_cur || !depth_cur)))
        return fail(h, MLD_ERR_INVALID_ARG, "mld_calculate_depth_pair: bad arguments");
    if (h->params.do_use_depth_segmentation && !h->params.set_all_depths_to_zero)
        return fail(h, MLD_ERR_REGION_GROWING, "DepthEstimator: Region growing not supported!");
    DeviceGuard g(h->device);
    int prev_stride_bytes = stride_bytes;
    if (prev_resident) {
        // the cloud that was current in the previous call becomes the previous cloud: its slot (points, pixel map, occupancy)
        // moves aside as it is, and only the new cloud crosses PCIe
        if (h->have_cloud) {
            std::swap(h->slots[0], h->slots[MLD_PREV_SLOT]);
            h->prev_n = h->cur_n;
            h->prev_stride_f = h->cur_stride_f;
            h->have_prev = true;
            h->have_cloud = false;
        } else {
            h->have_prev = false;
        }
        n_prev = h->have_prev ? h->prev_n : 0;
        prev_stride_bytes = h->prev_stride_f * 4;
    }
    const bool have_prev = prev_resident ? h->have_prev : pts_prev != nullptr;
    const bool road = h->params.do_use_ransac_plane != 0;
    const bool ransac_prev = have_prev && road && plane_prev && !plane_prev->segmented;
    const bool ransac_cur = road && plane_cur && !plane_cur->segmented;
    if ((ransac_prev && n_prev < 3) || (ransac_cur && n_cur < 3)) return fail(h, MLD_ERR_PCL_INVALID, "In GroundPlane: Input pointcloud is invalid");
    Slot& sc = h->slots[0];  // current cloud stays the handle's cloud
    Slot& sp = h->slots[prev_resident ? MLD_PREV_SLOT : 1];
    std::vector<unsigned int> hb_prev, hb_cur;
    std::vector<int32_t> st_prev_tmp, st_cur_tmp;
    if (have_prev) {
        rc = pair_side_begin(h, sp, pts_prev, n_prev, prev_stride_bytes, uv_prev, F_prev, (road && plane_prev && !ransac_prev) ? plane_prev : nullptr,
                             ransac_prev, ransac_seed, hb_prev, prev_resident);
        if (rc) return rc;
    } else {
        for (int i = 0; i < F_prev; i++) {  // depths.setConstant(-1) (tracklet_depth_module.cpp:97-100)
            depth_prev[i] = -1;
            if (status_prev) status_prev[i] = 0;
        }
    }
    int cur_sf = stride_bytes / 4;
    rc = pair_side_begin(h, sc, pts_cur, n_cur, stride_bytes, uv_cur, F_cur, (road && plane_cur && !ransac_cur) ? plane_cur : nullptr, ransac_cur,
                         ransac_seed + 1, hb_cur, false, &cur_sf);
    if (rc) return rc;
    auto finish = [&](Slot& s, int64_t n, int F, double* depth, int32_t* status, mld_plane* plane, bool ransac) -> int {
        if (F > 0) {
            CK(cudaMemcpyAsync(depth, s.d_depth, (size_t)F * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
            if (status) CK(cudaMemcpyAsync(status, s.d_status, (size_t)F * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        }
        if (ransac) {
            const long long words = (n + 31) / 32;
            std::vector<unsigned int> bits((size_t)words);
            int small[3] = {0, 0, 0};
            CK(cudaMemcpyAsync(plane->coeffs, s.d_coeffs, 4 * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(bits.data(), s.d_bits, (size_t)words * sizeof(unsigned int), cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(small, s.d_small, 3 * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            CK(cudaStreamSynchronize(s.stream));
            if (small[2] != 0) return fail(h, MLD_ERR_NO_MODEL, "RANSAC found no plane model");
            bits_to_plane(bits, n, plane);
            plane->segmented = 1;
        } else {
            CK(cudaStreamSynchronize(s.stream));
        }
        return MLD_OK;
    };
    if (have_prev) {
        rc = finish(sp, n_prev, F_prev, depth_prev, status_prev, plane_prev, ransac_prev);
        if (rc) return rc;
    }
    rc = finish(sc, n_cur, F_cur, depth_cur, status_cur, plane_cur, ransac_cur);
    if (rc) return rc;
    h->cur_n = n_cur;
    h->cur_stride_f = cur_sf;
    h->have_cloud = true;
    return MLD_OK;
}

int mld_calculate_depth_pair(mld_handle* h, const void* pts_prev, int64_t n_prev, const double* uv_prev, int F_prev,
                             double* depth_prev, int32_t* status_prev, mld_plane* plane_prev, const void* pts_cur, int64_t n_cur,
                             const double* uv_cur, int F_cur, double* depth_cur, int32_t* status_cur, mld_plane* plane_cur,
                             int stride_bytes, uint64_t ransac_seed) {
    return calculate_depth_pair_impl(h, pts_prev, n_prev, uv_prev, F_prev, depth_prev, status_prev, plane_prev, pts_cur, n_cur, uv_cur, F_cur,
                                     depth_cur, status_cur, plane_cur, stride_bytes, ransac_seed, false);
}

int mld_calculate_depth_pair_resident(mld_handle* h, const double* uv_prev, int F_prev, double* depth_prev, int32_t* status_prev,
                                      mld_plane* plane_prev, const void* pts_cur, int64_t n_cur, const double* uv_cur, int F_cur,
                                      double* depth_cur, int32_t* status_cur, mld_plane* plane_cur, int stride_bytes, uint64_t ransac_seed) {
    return calculate_depth_pair_impl(h, nullptr, 0, uv_prev, F_prev, depth_prev, status_prev, plane_prev, pts_cur, n_cur, uv_cur, F_cur, depth_cur,
                                     status_cur, plane_cur, stride_bytes, ransac_seed, true);
}

int mld_has_resident_cloud(const mld_handle* h) { return (h && h->have_cloud) ? 1 : 0; }

// ---- DepthCalculationStatistics / FeaturePoint packing ----
int mld_status_histogram_device(mld_handle* h, const int32_t* d_status, int64_t n, int64_t* hist21_out_host, void* stream) {
    if (!h || !hist21_out_host || n < 0 || (n > 0 && !d_status)) return MLD_ERR_INVALID_ARG;
    DeviceGuard g(h->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned long long* d_hist = nullptr;
    CK(cudaMalloc(&d_hist, 21 * sizeof(unsigned long long)));
    cudaError_t e = mld_launch_status_histogram(d_status, n, d_hist, st);
    h->launches++;
    unsigned long long hh[21];
    if (e == cudaSuccess) e = cudaMemcpyAsync(hh, d_hist, sizeof(hh), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_hist);
    if (e != cudaSuccess) return fail_cuda(h, e, "mld_status_histogram_device");
    for (int i = 0; i < 21; i++) hist21_out_host[i] = (int64_t)hh[i];
    return MLD_OK;
}

int mld_status_histogram_host(mld_handle* h, const int32_t* status_host, int64_t n, int64_t* hist21_out) {
    if (!h || !hist21_out || n < 0 || (n > 0 && !status_host)) return MLD_ERR_INVALID_ARG;
    DeviceGuard g(h->device);
    Slot& s = h->slots[2];
    CK(ensure(s.d_status, s.status_bytes, (size_t)std::max<int64_t>(n, 1) * sizeof(int)));
    if (n > 0) CK(cudaMemcpyAsync(s.d_status, status_host, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s.stream));
    return mld_status_histogram_device(h, s.d_status, n, hist21_out, s.stream);
}

int mld_pack_feature_points_device(mld_handle* h, const double* d_uv, const double* d_depth, int64_t n, float* d_out_uvd, void* stream) {
    if (!h || n < 0 || (n > 0 && (!d_uv || !d_depth || !d_out_uvd))) return MLD_ERR_INVALID_ARG;
    DeviceGuard g(h->device);
    CK(mld_launch_pack_feature_points(d_uv, d_depth, n, d_out_uvd, reinterpret_cast<cudaStream_t>(stream)));
    if (n > 0) h->launches++;
    return MLD_OK;
}

// ---- debug / parity views ----
int mld_get_pixel_map(mld_handle* h, int32_t* out_host) {
    if (!h || !out_host) return MLD_ERR_INVALID_ARG;
    if (!h->have_cloud) return fail(h, MLD_ERR_NO_CLOUD, "no cloud set");
    DeviceGuard g(h->device);
    Slot& s = h->slots[0];
    const size_t WH = (size_t)h->dp.W * (size_t)h->dp.H;
    CK(cudaMemcpyAsync(out_host, s.d_maps, WH * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    for (size_t i = 0; i < WH; i++) {  // decode the epoch-tagged cells into raw indices / -1
        unsigned int cell = (unsigned int)out_host[i];
        out_host[i] = map_cell_valid(s.mc, cell) ? (int32_t)map_cell_index(s.mc, cell) : -1;
    }
    return MLD_OK;
}

int mld_get_neighbors(mld_handle* h, double u, double v, double scale_w, double scale_h, int32_t* out_raw, int cap, int* k_out) {
    if (!h || !k_out || cap < 0) return MLD_ERR_INVALID_ARG;
    if (!h->have_cloud) return fail(h, MLD_ERR_NO_CLOUD, "no cloud set");
    DeviceGuard g(h->device);
    Slot& s = h->slots[0];
    if (!h->d_dbg) CK(cudaMalloc(&h->d_dbg, (size_t)(mld_neighbor_capacity() + 1) * sizeof(int)));
    int dcap = std::min(cap, mld_neighbor_capacity());
    double hx = static_cast<double>(h->params.pixelarea_search_witdh) * 0.5 * static_cast<double>((float)scale_w);
    double hy = static_cast<double>(h->params.pixelarea_search_height) * 0.5 * static_cast<double>((float)scale_h);
    CK(mld_launch_neighbors_debug(h->dp, s.mc, s.d_maps, u, v, hx, hy, h->d_dbg + 1, dcap, h->d_dbg, s.stream));
    h->launches++;
    int k = 0;
    CK(cudaMemcpyAsync(&k, h->d_dbg, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    int ncopy = std::min(k, dcap);
    if (ncopy > 0 && out_raw) CK(cudaMemcpy(out_raw, h->d_dbg + 1, (size_t)ncopy * sizeof(int), cudaMemcpyDeviceToHost));
    *k_out = k;
    return MLD_OK;
}

static int debug_views(mld_handle* h, uint8_t* vis_host, int64_t* n_vis, double* cam_host) {
    if (!h->have_cloud) return fail(h, MLD_ERR_NO_CLOUD, "no cloud set");
    DeviceGuard g(h->device);
    Slot& s = h->slots[0];
    const long long n = h->cur_n;
    if (n == 0) {
        if (n_vis) *n_vis = 0;
        return MLD_OK;
    }
    unsigned char* d_vis = nullptr;
    double* d_cam = nullptr;
    if (vis_host) CK(cudaMalloc(&d_vis, (size_t)n));
    if (cam_host) CK(cudaMalloc(&d_cam, (size_t)n * 3 * sizeof(double)));
    cudaError_t e = mld_launch_visible_debug(h->dp, reinterpret_cast<const float*>(s.d_pts), h->cur_stride_f, n, d_vis, d_cam, s.stream);
    h->launches++;
    if (e == cudaSuccess && vis_host) e = cudaMemcpyAsync(vis_host, d_vis, (size_t)n, cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess && cam_host) e = cudaMemcpyAsync(cam_host, d_cam, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    cudaFree(d_vis);
    cudaFree(d_cam);
    if (e != cudaSuccess) return fail_cuda(h, e, "debug_views");
    if (vis_host && n_vis) {
        int64_t c = 0;
        for (long long i = 0; i < n; i++) c += vis_host[i] ? 1 : 0;
        *n_vis = c;
    }
    return MLD_OK;
}

int mld_get_visible(mld_handle* h, uint8_t* out_visible_host, int64_t* n_visible_out) {
    if (!h || !out_visible_host) return MLD_ERR_INVALID_ARG;
    return debug_views(h, out_visible_host, n_visible_out, nullptr);
}
int mld_get_visible_points(mld_handle* h, int32_t* point_index_out, double* image_points_out, double* depth_cam_out,
                           int64_t capacity, int64_t* n_visible_out) {
    if (!h || capacity < 0) return MLD_ERR_INVALID_ARG;
    if (!h->have_cloud) return fail(h, MLD_ERR_NO_CLOUD, "no cloud set");
    DeviceGuard g(h->device);
    Slot& s = h->slots[0];
    const long long n = h->cur_n;
    const long long cap = std::min<long long>(capacity, n);
    unsigned char* d_buf = nullptr;
    const size_t scratch = (mld_visible_scratch_bytes(n) + 15) / 16 * 16;
    const size_t idx_b = point_index_out ? ((size_t)cap * sizeof(int) + 15) / 16 * 16 : 0;
    const size_t img_b = image_points_out ? (size_t)cap * 2 * sizeof(double) : 0;
    const size_t dep_b = depth_cam_out ? (size_t)cap * sizeof(double) : 0;
    CK(cudaMalloc(&d_buf, scratch + idx_b + img_b + dep_b + 16));
    int* d_idx = point_index_out ? reinterpret_cast<int*>(d_buf + scratch) : nullptr;
    double* d_img = image_points_out ? reinterpret_cast<double*>(d_buf + scratch + idx_b) : nullptr;
    double* d_dep = depth_cam_out ? reinterpret_cast<double*>(d_buf + scratch + idx_b + img_b) : nullptr;
    const unsigned int* d_count = nullptr;
    int nl = 0;
    cudaError_t e = mld_launch_visible_compact(h->dp, reinterpret_cast<const float*>(s.d_pts), h->cur_stride_f, n, d_buf, cap, d_idx, d_img,
                                               d_dep, &d_count, s.stream, &nl);
    h->launches += nl;
    unsigned int nvis = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&nvis, d_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    const size_t w = (size_t)std::min<long long>(cap, (long long)nvis);
    if (e == cudaSuccess && d_idx && w) e = cudaMemcpy(point_index_out, d_idx, w * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && d_img && w) e = cudaMemcpy(image_points_out, d_img, w * 2 * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && d_dep && w) e = cudaMemcpy(depth_cam_out, d_dep, w * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_buf);
    if (e != cudaSuccess) return fail_cuda(h, e, "mld_get_visible_points");
    if (n_visible_out) *n_visible_out = (int64_t)nvis;
    return MLD_OK;
}
int mld_get_points_camera_indexed(mld_handle* h, const int32_t* idx_host, int64_t n_idx, double* out_host) {
    if (!h || n_idx < 0 || (n_idx > 0 && (!idx_host || !out_host))) return MLD_ERR_INVALID_ARG;
    if (!h->have_cloud) return fail(h, MLD_ERR_NO_CLOUD, "no cloud set");
    if (n_idx == 0) return MLD_OK;
    DeviceGuard g(h->device);
    Slot& s = h->slots[0];
    int* d_idx = nullptr;
    double* d_out = nullptr;
    CK(cudaMalloc(&d_idx, (size_t)n_idx * sizeof(int)));
    cudaError_t e = cudaMalloc(&d_out, (size_t)n_idx * 3 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, idx_host, (size_t)n_idx * sizeof(int), cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess)
        e = mld_launch_points_camera_indexed(h->dp, reinterpret_cast<const float*>(s.d_pts), h->cur_stride_f, h->cur_n, d_idx, n_idx, d_out, s.stream);
    h->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d_out, (size_t)n_idx * 3 * sizeof(double), cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    cudaFree(d_idx);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail_cuda(h, e, "mld_get_points_camera_indexed");
    return MLD_OK;
}
int mld_get_points_camera(mld_handle* h, double* out_host) {
    if (!h || !out_host) return MLD_ERR_INVALID_ARG;
    return debug_views(h, nullptr, nullptr, out_host);
}

// ---- synthetic input (device generators; the host generators live in libmld_synth.so) ----
static int synth_tables(mld_handle* h, const mld_synth_config* c, cudaStream_t st) {
    const size_t tn = mld_synth_table_floats(*c);
    if (h->synth_tables_valid && memcmp(&h->synth_cfg_cached, c, sizeof(*c)) == 0) return MLD_OK;
    std::vector<float> tables(tn);
    mld_synth_build_tables(*c, tables.data());
    CK(cudaStreamSynchronize(st));
    if (h->d_synth_tables) CK(cudaFree(h->d_synth_tables));
    h->d_synth_tables = nullptr;
    h->synth_tables_valid = false;
    CK(cudaMalloc(&h->d_synth_tables, tn * sizeof(float)));
    CK(cudaMemcpy(h->d_synth_tables, tables.data(), tn * sizeof(float), cudaMemcpyHostToDevice));
    h->synth_cfg_cached = *c;
    h->synth_tables_valid = true;
    return MLD_OK;
}

int mld_synth_points_device(mld_handle* h, const mld_synth_config* c, uint64_t seed, int64_t frame0, int64_t nframes,
                            int64_t frame_pitch_points, float* d_out_xyzi, void* stream) {
    if (!h || !mld_synth_config_ok(c) || !d_out_xyzi) return MLD_ERR_INVALID_ARG;
    if (frame_pitch_points < (int64_t)c->rings * c->azimuth_steps) return fail(h, MLD_ERR_INVALID_ARG, "frame pitch smaller than the cloud");
    DeviceGuard g(h->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = synth_tables(h, c, st);
    if (rc) return rc;
    CK(mld_launch_synth_points(*c, seed, frame0, nframes, frame_pitch_points, h->d_synth_tables, d_out_xyzi, st));
    return MLD_OK;
}

int mld_synth_features_device(mld_handle* h, const mld_synth_config* c, uint64_t seed, int64_t frame0, int64_t nframes, int F,
                              double* d_out_uv, void* stream) {
    if (!h || !mld_synth_config_ok(c) || !d_out_uv || F < 0) return MLD_ERR_INVALID_ARG;
    DeviceGuard g(h->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = synth_tables(h, c, st);
    if (rc) return rc;
    CK(mld_launch_synth_features(*c, seed, frame0, nframes, F, h->d_synth_tables, d_out_uv, st));
    return MLD_OK;
}

}  // extern "C"
