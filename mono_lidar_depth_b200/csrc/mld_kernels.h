// mld_kernels.h -- host-callable launchers of the sm_100a kernels (internal to libmld_cuda.so).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mld_c_api.h"

struct DevParams;
struct MapCode;

// K1 (mld_project.cu)
void mld_setup_prefilter(DevParams& P);
// d_occ: occupancy bitmaps (occ_words_per_row(W) * H words per frame, zeroed by the caller) or nullptr
cudaError_t mld_launch_project_scatter(const DevParams& P, const MapCode& mc, const float* d_pts, int stride_f, long long n,
                                       long long pitch_pts, unsigned int* d_maps, unsigned int* d_occ, int nframes,
                                       cudaStream_t stream);
// visible-order compaction: _pointIndex, _points_cs_image_visible (2 x nvis), camera-frame depth
size_t mld_visible_scratch_bytes(long long n);
cudaError_t mld_launch_visible_compact(const DevParams& P, const float* d_pts, int stride_f, long long n, void* d_scratch, long long capacity,
                                       int* d_point_index, double* d_image_points, double* d_depth_cam, const unsigned int** d_count_out,
                                       cudaStream_t stream, int* launches);
cudaError_t mld_launch_points_camera_indexed(const DevParams& P, const float* d_pts, int stride_f, long long n, const int* d_idx, long long n_idx,
                                             double* d_out, cudaStream_t stream);
cudaError_t mld_launch_visible_debug(const DevParams& P, const float* d_pts, int stride_f, long long n,
                                     unsigned char* d_visible, double* d_cam, cudaStream_t stream);

// K2/K3, warp per feature (mld_feature.cu). d_list == nullptr: every feature of every frame;
// otherwise the warps of list_blocks blocks stride the *d_list_count global feature ids in d_list.
int mld_feature_capacity_for(int max_area);
cudaError_t mld_configure_feature_depth(int kcap);
cudaError_t mld_launch_feature_depth(const DevParams& P, const MapCode& mc, int kcap, const float* d_pts, int stride_f,
                                     long long pitch_pts, const unsigned int* d_maps, const double* d_uv, int F, double* d_depth,
                                     int* d_status, const float* d_plane_coeffs, const unsigned int* d_inlier_bits,
                                     long long words_per_frame, int nframes, const int* d_list, const int* d_list_count,
                                     int list_blocks, cudaStream_t stream, double* d_corners = nullptr /* debug: 9 doubles per feature */);
cudaError_t mld_launch_neighbors_debug(const DevParams& P, const MapCode& mc, const unsigned int* d_map, double u, double v,
                                       double hx, double hy, int* d_out, int cap, int* d_k, cudaStream_t stream);

// K2/K3 split into gather / solve / road kernels with a chunk-wide compaction (mld_feature_split.cu)
size_t mld_split_scratch_bytes(long long features, int road);
cudaError_t mld_launch_feature_depth_split(const DevParams& P, const MapCode& mc, const float* d_pts, int stride_f,
                                           long long pitch_pts, const unsigned int* d_maps, const unsigned int* d_occ,
                                           const double* d_uv, int F, double* d_depth, int* d_status, const float* d_plane_coeffs,
                                           const unsigned int* d_inlier_bits, long long words_per_frame, int nframes,
                                           int* d_overflow_list, int* d_overflow_count, void* d_scratch, cudaStream_t stream,
                                           int* launches, cudaEvent_t* ev_mid = nullptr);

// K1 of one chunk + K2a (gather) of the previous chunk as one heterogeneous launch, and K2b (+ road kernels) alone
cudaError_t mld_launch_fused_project_gather(const DevParams& P, int stride_f, const MapCode& mc_k1, const float* d_pts_k1, long long n_points,
                                            long long pitch_pts, unsigned int* d_maps_k1, unsigned int* d_occ_k1, int frames_k1,
                                            const MapCode& mc_g, const float* d_pts_g, const unsigned int* d_maps_g,
                                            const unsigned int* d_occ_g, const double* d_uv_g, int F, double* d_depth_g, int* d_status_g,
                                            int frames_g, int* d_overflow_list, int* d_overflow_count, void* d_scratch_g,
                                            cudaStream_t stream, int* launches);
cudaError_t mld_launch_feature_solve(const DevParams& P, const MapCode& mc, const float* d_pts, int stride_f, long long pitch_pts,
                                     const unsigned int* d_maps, const unsigned int* d_occ, const double* d_uv, int F, double* d_depth,
                                     int* d_status, const float* d_plane_coeffs, const unsigned int* d_inlier_bits,
                                     long long words_per_frame, int nframes, int* d_overflow_list, int* d_overflow_count, void* d_scratch,
                                     cudaStream_t stream, int* launches, cudaEvent_t* ev_after_solve = nullptr);


// K4 (mld_ransac.cu): per-frame ground-plane RANSAC.
struct RansacConfig {
    double distance_treshold;     // ransac_plane_distance_treshold
    double refinement_treshold;   // ransac_plane_refinement_treshold
    double probability;           // ransac_plane_probability
    double min_z, max_z;          // ransac_plane_min_z / max_z (PassThrough when min_z > -1001)
    int max_iterations;           // ransac_plane_max_iterations
    int use_refinement;           // ransac_plane_use_refinement
    double cos_eps;               // cos(M_PI / 18.): SampleConsensusModelPerpendicularPlane eps angle (RansacPlane.cpp:99)
    double log_probability;       // log(1 - probability)
    float thr_lt, refine_lt;      // largest floats whose double value is < distance / refinement threshold
};
constexpr int MLD_RANSAC_SAMPLE = 6000;  // _numberRandomSamplePoints, RansacPlane.cpp:32

// Scratch per frame (device): see mld_ransac.cu. Sizes in bytes for nframes.
size_t mld_ransac_scratch_bytes(long long n_points, int nframes);
// SemanticPlane (mld_semantic.cu): d_state holds mld_semantic_state_bytes(nframes) bytes of scratch
size_t mld_semantic_state_bytes(int nframes);
cudaError_t mld_launch_semantic_plane(const double* T_cam_lidar, double f, double cu, double cv, int label_w, int label_h,
                                      const unsigned int* ground_set8, double inlier_threshold, const float* d_pts, int stride_f,
                                      long long n_points, long long pitch_pts, const unsigned char* d_labels, int nframes,
                                      void* d_state, float* d_coeffs, unsigned int* d_inlier_bits, long long words_per_frame,
                                      int* d_n_inliers, int* d_rc, cudaStream_t stream, int* launches,
                                      unsigned char* d_flags_out = nullptr /* ground-labelled flag per point (debug view; input of the exact mode) */,
                                      int exact = 0 /* PCL's sequential float accumulation order: bit-exact coefficients and inlier set */);
// Fits one plane per frame. Outputs per frame: coeffs[4] (float), inlier bitmask over raw indices
// (words_per_frame uint32), n_inliers, iterations, rc (0 ok, MLD_ERR_PCL_INVALID, MLD_ERR_NO_MODEL).
cudaError_t mld_launch_ransac(const RansacConfig& cfg, const float* d_pts, int stride_f, long long n_points,
                              long long pitch_pts, int nframes, uint64_t seed, long long frame0, void* d_scratch,
                              float* d_coeffs, unsigned int* d_inlier_bits, long long words_per_frame, int* d_n_inliers,
                              int* d_iterations, int* d_rc, cudaStream_t stream, int* launches);
// extras (mld_extras.cu)
cudaError_t mld_launch_status_histogram(const int* d_status, long long n, unsigned long long* d_hist21, cudaStream_t stream);
cudaError_t mld_launch_unpack_xyz(const float* d_xyz, float* d_out4, long long n, cudaStream_t stream);
cudaError_t mld_launch_pack_feature_points(const double* d_uv, const double* d_depth, long long n, float* d_out, cudaStream_t stream);

// synthetic data (mld_synth.cu; scene model in mld_synth_model.h, host generators in libmld_synth.so)
cudaError_t mld_launch_synth_points(const mld_synth_config& c, uint64_t seed, long long frame0, long long nframes,
                                    long long pitch_pts, const float* d_tables, float* d_out, cudaStream_t stream);
cudaError_t mld_launch_synth_features(const mld_synth_config& c, uint64_t seed, long long frame0, long long nframes, int F,
                                      const float* d_tables, double* d_out, cudaStream_t stream);
void mld_synth_build_tables(const mld_synth_config& c, float* tables);
size_t mld_synth_table_floats(const mld_synth_config& c);
bool mld_synth_config_ok(const mld_synth_config* c);
