/*
 * mld_c_api.h -- C ABI of libmld_cuda.so, the B200 (sm_100a) implementation of the
 * monolidar_fusion depth-estimation hot path.
 *
 * This is the drop-in boundary: the reference's C++ class Mono_Lidar::DepthEstimator
 * (monolidar_fusion/include/monolidar_fusion/DepthEstimator.h:39-359) keeps its public
 * signature and forwards to these entry points (see shim/ and INTEGRATION.md). Plain C
 * types only; no torch / Eigen / PCL types cross this boundary. There is no CPU fallback:
 * every compute entry point returns MLD_ERR_CUDA when no usable device is present.
 *
 * Conventions
 *   - every function returns 0 (MLD_OK) or a negative mld_error; mld_last_error() gives text.
 *   - "host" pointers are ordinary (ideally pinned) host memory, "device" pointers are CUDA
 *     device memory on the handle's device.
 *   - points: x,y,z as the first three floats of each element, stride_bytes apart
 *     (16 for float4, 32 for pcl::PointXYZI).
 *   - features: 2 x F column-major doubles (u0,v0,u1,v1,...) == Eigen::Matrix2Xd::data().
 *   - status codes are Mono_Lidar::DepthResultType (eDepthResultType.h:9-31).
 */
#ifndef MLD_C_API_H
#define MLD_C_API_H

#include <stdint.h>

#include "mld_synth.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum mld_error {
    MLD_OK = 0,
    MLD_ERR_INVALID_ARG = -1,
    MLD_ERR_NOT_CONFIGURED = -2,   /* reference: throw "Call 'InitConfig' before calling 'Initialize'." (DepthEstimator.cpp:38) */
    MLD_ERR_NOT_INITIALIZED = -3,  /* reference: throw "call of 'setInputCloud' without 'initialize'" (:227) */
    MLD_ERR_NO_CLOUD = -4,         /* reference: throw "call of 'CalculateDepth' without 'SetInputCloud'" (:439) */
    MLD_ERR_BAD_SEARCH_MODE = -5,  /* reference: throw "neighbor_search_mode has the invalid value" (:57) */
    MLD_ERR_REGION_GROWING = -6,   /* reference: runtime_error "Region growing not supported!" (:608) */
    MLD_ERR_PCL_INVALID = -7,      /* reference: GroundPlane::ExceptionPclInvalid, cloud < 3 points (RansacPlane.cpp:44-50) */
    MLD_ERR_NO_ROAD_ESTIMATOR = -8,/* reference: throw "No road depth estimator selected." (:94) */
    MLD_ERR_CAPACITY = -9,         /* search window larger than the kernels' neighbour capacity */
    MLD_ERR_CUDA = -10,            /* CUDA runtime error or no device */
    MLD_ERR_NO_MODEL = -11,        /* RANSAC found no model */
    MLD_ERR_IO = -12               /* settings file cannot be read (reference: throw "Cant find settings file") */
} mld_error;

/* Status codes written to the status outputs: Mono_Lidar::DepthResultType (eDepthResultType.h:9-31), one list for the
 * C ABI, the kernels, mld_status_name() and the C++ shim's enum. X(name, value). */
#define MLD_DEPTH_RESULT_TYPES(X)                                                                                              \
    X(Unspecified, 0) X(Success, 1) X(RadiusSearchInsufficientPoints, 2) X(HistogramNoLocalMax, 3)                             \
    X(TresholdDepthGlobalGreaterMax, 4) X(TresholdDepthGlobalSmallerMin, 5) X(TresholdDepthLocalGreaterMax, 6)                 \
    X(TresholdDepthLocalSmallerMin, 7) X(TriangleNotPlanar, 8) X(TriangleNotPlanarInsufficientPoints, 9)                       \
    X(CornerBehindCamera, 10) X(PlaneViewrayNotOrthogonal, 11) X(PcaIsPoint, 12) X(PcaIsLine, 13) X(PcaIsCubic, 14)            \
    X(InsufficientRoadPoints, 15) X(SuccessRoad, 16) X(RegionGrowingNearestSeedNotAvailable, 17)                               \
    X(RegionGrowingSeedsOutOfRange, 18) X(RegionGrowingInsufficientPoints, 19) X(SuccessRegionGrowing, 20)
#define MLD_STATUS_ENUM_ENTRY(name, value) MLD_STATUS_##name = value,
typedef enum mld_status { MLD_DEPTH_RESULT_TYPES(MLD_STATUS_ENUM_ENTRY) MLD_STATUS_COUNT = 21 } mld_status;

/* Mono_Lidar::DepthEstimatorParameters (DepthEstimatorParameters.h:12-172): same field names,
 * bools as int32 (the reference's loader reads them as (int), DepthEstimatorParameters.cpp:27 ff.).
 * Only the fields the hot path reads are present. */
typedef struct mld_params {
    int32_t neighbor_search_mode;
    int32_t pixelarea_search_witdh;
    int32_t pixelarea_search_height;
    int32_t radiusSearch_count_min;

    int32_t do_use_histogram_segmentation;
    int32_t histogram_segmentation_min_pointcount;
    double histogram_segmentation_bin_witdh;

    int32_t do_use_depth_segmentation;

    int32_t treshold_depth_enabled;
    int32_t treshold_depth_mode;
    int32_t treshold_depth_max;
    int32_t treshold_depth_min;

    int32_t treshold_depth_local_enabled;
    int32_t treshold_depth_local_mode;
    int32_t treshold_depth_local_valuetype;
    double treshold_depth_local_value;

    int32_t do_use_PCA;
    int32_t pca_debug;
    double pca_treshold_3_abs_min;
    double pca_treshold_3_2_rel_max;
    double pca_treshold_2_1_rel_min;

    int32_t do_use_ransac_plane;
    int32_t ransac_plane_max_iterations;
    double ransac_plane_distance_treshold;
    double ransac_plane_min_z;
    double ransac_plane_max_z;
    int32_t ransac_plane_use_refinement;
    int32_t ransac_plane_use_camx_treshold;
    double ransac_plane_refinement_treshold;
    double ransac_plane_treshold_camx;
    double ransac_plane_point_distance_treshold;
    double ransac_plane_probability;

    int32_t plane_estimator_use_triangle_maximation;
    int32_t plane_estimator_use_leastsquares;
    int32_t plane_estimator_use_mestimator;
    int32_t do_use_cut_behind_camera;
    double plane_estimator_z_x_min_relation;

    int32_t do_use_triangle_size_maximation;
    int32_t do_check_triangleplanar_condition;
    double triangleplanar_crossnorm_treshold;
    double viewray_plane_orthoganality_treshold;
    int32_t set_all_depths_to_zero;
    int32_t reserved0;
} mld_params;

/* Host view of Mono_Lidar::GroundPlane (RansacPlane.h:38-126): getModelCoeffs(), getInlinersIndex()
 * / CheckPointInPlane(), isSegmented(). Coefficients are in the LIDAR frame. The library never
 * frees inlier_idx: on output (mld_set_cloud with RANSAC, mld_estimate_ground_plane) the caller
 * provides inlier_idx with room for inlier_capacity entries. */
typedef struct mld_plane {
    float coeffs[4];
    int32_t* inlier_idx;
    int64_t n_inliers;
    int64_t inlier_capacity;
    int32_t segmented;
    int32_t reserved0;
} mld_plane;

typedef struct mld_handle mld_handle;

/* ---- configuration (DepthEstimator::InitConfig, DepthEstimator.cpp:129-154) ---- */
int mld_sizeof_params(void);
void mld_default_params(mld_params* p);                       /* C++ member defaults */
int mld_params_from_yaml(const char* path, mld_params* p);    /* DepthEstimatorParameters::fromFile (DepthEstimatorParameters.cpp:16-114):
                                                                 flat "key: value # comment" OpenCV-YAML; absent keys read as 0 */
/* one integer key of the same flat yaml (DepthEstimatorParameters' debug switches that are not part of mld_params); absent -> 0 */
int mld_yaml_int(const char* path, const char* key, int32_t* out, int32_t* found);
const char* mld_status_name(int status);                      /* DepthResultTypeMap (DepthEstimator.h:45-60) */
const char* mld_last_error(const mld_handle* h);              /* h may be NULL: error of the last failed mld_create on this thread */

int mld_create(const mld_params* p, int device, mld_handle** out);   /* device < 0: current device */
int mld_destroy(mld_handle* h);

/* ---- DepthEstimator::Initialize (DepthEstimator.cpp:35-127) ----
 * camera = CameraPinhole(W,H,f,cx,cy) (camera_pinhole.h:21-26); T = row-major 3x4 [R|t] of
 * transform_lidar_to_cam (Eigen::Affine3d::matrix().topRows<3>()). */
int mld_initialize(mld_handle* h, int W, int H, double f, double cx, double cy, const double* T_lidar_to_cam);

/* ---- DepthEstimator::setInputCloud (DepthEstimator.cpp:220-312) ----
 * Host points in, one H2D copy, projection + pixel map on the device. When do_use_ransac_plane is
 * set and inout_plane is non-NULL with segmented == 0, the ground plane is fitted on the device
 * (RansacPlane::CalculateInliersPlane) with ransac_seed and written back to *inout_plane.
 * points_host may be pageable (a pcl::PointCloud) or pinned and is free for reuse when the call returns. A large cloud in
 * pageable memory is stripped to 12-byte xyz by host worker threads into a pinned staging buffer and the call returns while the
 * copy and the projection are still running; every later call on the handle is ordered behind them (an error they raise is
 * reported by that call). MLD_HOST_PACK=0 restores the plain copy + wait. */
int mld_set_cloud(mld_handle* h, const void* points_host, int64_t n, int stride_bytes, mld_plane* inout_plane,
                  uint64_t ransac_seed);

/* ---- DepthEstimator::CalculateDepth (DepthEstimator.cpp:429-488) on the current cloud ----
 * plane == NULL is the reference's ransacPlane == nullptr (road path skipped). */
int mld_calculate_depth(mld_handle* h, const double* uv_host, int F, double* depth_host, int32_t* status_host,
                        const mld_plane* plane);

/* ---- RansacPlane::CalculateInliersPlane (RansacPlane.cpp:41-140), stand-alone ---- */
int mld_estimate_ground_plane(mld_handle* h, const void* points_host, int64_t n, int stride_bytes, uint64_t seed,
                              mld_plane* out_plane, int32_t* iterations_out);

/* ---- SemanticPlane::CalculateInliersPlane (RansacPlane.cpp:159-274), stand-alone ----
 * The ground plane the production caller builds from a semantic label image
 * (tracklets_depth/src/tracklet_depth_module.cpp:269-284): points whose projection carries a ground label ->
 * least-squares plane -> inliers within inlier_threshold over the whole cloud -> refit.
 * labels_host: label_h x label_w uint8, row-major (the cv::Mat); f/cu/cv and T_cam_lidar (row-major 3x4) are
 * SemanticPlane::Camera; ground_labels: the std::set<int> (values outside 0..255 never match a uint8 pixel).
 * Fails with MLD_ERR_PCL_INVALID when fewer than 3 points carry a ground label (ExceptionPclInvalid, :224-227).
 * out_plane receives getModelCoeffs() / getInlinersIndex() and segmented = 1; it can be passed to mld_set_cloud /
 * mld_calculate_depth like any GroundPlane. Does not disturb the handle's current cloud. */
int mld_semantic_ground_plane(mld_handle* h, const void* points_host, int64_t n, int stride_bytes, const uint8_t* labels_host,
                              int label_w, int label_h, double f, double cu, double cv, const double* T_cam_lidar,
                              const int32_t* ground_labels, int n_ground_labels, double inlier_threshold, mld_plane* out_plane);
/* Fit mode of every SemanticPlane entry point of this handle. 0 (default): the moments of the labelled points / inliers are
 * accumulated in double by a parallel reduction -- the better-conditioned fit; coefficients agree with the reference's to ~2e-3,
 * inlier sets differ only for points within that margin of the threshold. 1: PCL's own order -- nine sequential FLOAT
 * accumulators over the points in index order (computeMeanAndCovarianceMatrix), which loses 4-5 digits over 1e4 points but is what
 * the reference computes: coefficients and inlier set bit-identical to it (about 0.2 ms per sweep and pass instead of 15 us).
 * Env MLD_SEMANTIC_EXACT=1 sets it at mld_create. */
int mld_set_semantic_exact(mld_handle* h, int on);
/* debug / parity view of the first stage: out_flags_host[i] = 1 when point i projects onto a ground-labelled pixel
 * (the points kept by RansacPlane.cpp:201-222), 0 otherwise. */
int mld_semantic_ground_labelled(mld_handle* h, const void* points_host, int64_t n, int stride_bytes, const uint8_t* labels_host,
                                 int label_w, int label_h, double f, double cu, double cv, const double* T_cam_lidar,
                                 const int32_t* ground_labels, int n_ground_labels, uint8_t* out_flags_host);
/* Device-resident, batched form: nframes clouds (frame_pitch_points apart) and nframes label images back to back.
 * Outputs per frame: 4 coefficients, an inlier bitmask over raw indices ((n_points + 31) / 32 words, the format the
 * road path consumes), the inlier count and a return code (0 or MLD_ERR_PCL_INVALID). Enqueued on `stream`. */
int mld_semantic_ground_plane_device(mld_handle* h, const void* d_points, int64_t n_points, int64_t frame_pitch_points,
                                     int stride_bytes, const uint8_t* d_labels, int label_w, int label_h, double f, double cu,
                                     double cv, const double* T_cam_lidar, const int32_t* ground_labels, int n_ground_labels,
                                     double inlier_threshold, int64_t nframes, float* d_coeffs_out, uint32_t* d_inlier_bits_out,
                                     int32_t* d_n_inliers_out, int32_t* d_rc_out, void* stream);

/* ---- batched, device-resident sequence (the benchmark path; frames are independent) ----
 * d_points: nframes clouds, frame_pitch_points elements apart, n_points valid in each.
 * d_uv: nframes x (2 x F) doubles; d_depth: nframes x F; d_status: nframes x F.
 * road: 0 = plane nullptr for every frame; 1 = fit a RANSAC ground plane per frame on the device
 * (seed + frame index) and run the road path with it.
 * d_plane_coeffs_out (nullable): nframes x 4 floats, the fitted coefficients.
 * Work is enqueued on `stream` (a cudaStream_t, may be 0); the call does not synchronise. */
int mld_process_frames_device(mld_handle* h, const void* d_points, int64_t n_points, int64_t frame_pitch_points,
                              int stride_bytes, const double* d_uv, int F, double* d_depth, int32_t* d_status,
                              int64_t nframes, int road, uint64_t seed, float* d_plane_coeffs_out, void* stream);

/* Same sequence with the ground plane of every frame fitted from a semantic label image on the device
 * (SemanticPlane::CalculateInliersPlane, RansacPlane.cpp:195-274) before the road path runs: the per-frame work of
 * TrackletDepthModule::process (tracklets_depth/src/tracklet_depth_module.cpp:269-330) in one call.
 * d_labels: nframes images of label_h x label_w uint8. d_plane_rc_out (nullable): 0 or MLD_ERR_PCL_INVALID per frame
 * (fewer than 3 ground-labelled points; such a frame gets no road depths). Requires a road estimator
 * (do_use_ransac_plane != 0 in the parameters), else MLD_ERR_NO_ROAD_ESTIMATOR. */
int mld_process_frames_device_semantic(mld_handle* h, const void* d_points, int64_t n_points, int64_t frame_pitch_points,
                                       int stride_bytes, const uint8_t* d_labels, int label_w, int label_h, double f, double cu,
                                       double cv, const double* T_cam_lidar, const int32_t* ground_labels, int n_ground_labels,
                                       double inlier_threshold, const double* d_uv, int F, double* d_depth, int32_t* d_status,
                                       int64_t nframes, float* d_plane_coeffs_out, int32_t* d_plane_rc_out, void* stream);
/* Same sequence with caller-provided, device-resident planes: nframes x 4 coefficients (lidar frame) and nframes inlier
 * bitmasks over raw indices ((n_points + 31) / 32 words each) -- any GroundPlane computed elsewhere. */
int mld_process_frames_device_planes(mld_handle* h, const void* d_points, int64_t n_points, int64_t frame_pitch_points,
                                     int stride_bytes, const float* d_plane_coeffs, const uint32_t* d_inlier_bits,
                                     const double* d_uv, int F, double* d_depth, int32_t* d_status, int64_t nframes, void* stream);

/* ---- batched, host-resident sequence: same as above with pinned (or pageable) host buffers;
 * H2D / kernels / D2H are pipelined over internal streams; returns after the results are in host
 * memory. plane_coeffs_out_host nullable. */
int mld_process_frames_host(mld_handle* h, const void* points_host, int64_t n_points, int64_t frame_pitch_points,
                            int stride_bytes, const double* uv_host, int F, double* depth_host, int32_t* status_host,
                            int64_t nframes, int road, uint64_t seed, float* plane_coeffs_out_host);

/* counters of mld_process_frames_host since mld_create: [0] bytes copied host -> device, [1] device -> host, [2] frames whose
 * points were packed to 12-byte xyz by the host threads, [3] frames copied as whole records. Records wider than 16 bytes
 * (pcl::PointXYZI) are packed whenever the copy engine still has work queued and copied as they are when it would idle; env
 * MLD_HOST_PACK=1 / 0 forces packing on / off, MLD_PACK_THREADS sets the worker count. */
int mld_host_pipeline_stats(const mld_handle* h, int64_t* out4);

/* ---- tracklets_depth batch adaptor ----
 * TrackletDepthModule::process calls CalculateDepth twice per frame, for the previous and the current cloud
 * with different feature sets (tracklets_depth/src/tracklet_depth_module.cpp:318, :330). Both clouds go through
 * the device concurrently (own streams, maps and staging buffers); a NULL cloud (no previous frame yet,
 * :97-100) yields depth -1 for its features. Planes follow mld_set_cloud / mld_calculate_depth semantics;
 * the status outputs may be NULL (the 4-argument overload discards them). After the call the handle's
 * current cloud is the `cur` cloud. */
int mld_calculate_depth_pair(mld_handle* h, const void* pts_prev, int64_t n_prev, const double* uv_prev, int F_prev,
                             double* depth_prev, int32_t* status_prev, mld_plane* plane_prev, const void* pts_cur,
                             int64_t n_cur, const double* uv_cur, int F_cur, double* depth_cur, int32_t* status_cur,
                             mld_plane* plane_cur, int stride_bytes, uint64_t ransac_seed);

/* The same per-frame work when the caller walks a sequence (TrackletDepthModule keeps _cloud_last_frame = the cloud of its previous
 * callback, tracklet_depth_module.cpp:318-354): the previous cloud is the one that was `cur` in the last call of either pair entry
 * point (or the handle's current cloud after mld_set_cloud) and is STILL ON THE DEVICE with its pixel map -- only the new cloud
 * crosses PCIe, one upload and one projection per frame instead of two. Without a resident cloud (first frame, or a batched call
 * in between) the previous side yields depth -1 like a NULL cloud. mld_has_resident_cloud tells which case the next call is. */
int mld_calculate_depth_pair_resident(mld_handle* h, const double* uv_prev, int F_prev, double* depth_prev, int32_t* status_prev,
                                      mld_plane* plane_prev, const void* pts_cur, int64_t n_cur, const double* uv_cur, int F_cur,
                                      double* depth_cur, int32_t* status_cur, mld_plane* plane_cur, int stride_bytes, uint64_t ransac_seed);
int mld_has_resident_cloud(const mld_handle* h);

/* ---- DepthCalculationStatistics (DepthEstimator.cpp:1039-1090): counters per DepthResultType (0..20) ----
 * mld_set_statistics(on): every mld_calculate_depth call also reduces its status array (still on the device) to the 21
 * counters (the reference's per-feature LogDepthCalcStats, DepthEstimator.cpp:470-479); mld_last_status_histogram returns the
 * counters of the last call (zeros before the first). */
int mld_set_statistics(mld_handle* h, int on);
int mld_last_status_histogram(mld_handle* h, int64_t* hist21_out);
int mld_status_histogram_host(mld_handle* h, const int32_t* status_host, int64_t n, int64_t* hist21_out);
int mld_status_histogram_device(mld_handle* h, const int32_t* d_status, int64_t n, int64_t* hist21_out_host, void* stream);

/* ---- matches_msg_depth_ros/FeaturePoint {float32 u, v, d} packing (FeaturePoint.msg:1-5) ----
 * device buffers in, device buffer of 3 floats per feature out; enqueued on `stream`. */
int mld_pack_feature_points_device(mld_handle* h, const double* d_uv, const double* d_depth, int64_t n, float* d_out_uvd,
                                   void* stream);

/* number of kernels this handle has launched since creation (for bench.py's gpu_launches) */
int64_t mld_kernel_launch_count(const mld_handle* h);
/* per-kernel device timing with CUDA events on the launching stream (bench.py's roofline figures).
 * Classes (7): 0 = pixel-map / occupancy clear (memset nodes), 1 = project_scatter, 2 = ransac, 3 = feature_depth (all
 * K2 kernels of the chunk), 4 = feature_gather, 5 = feature_solve, 6 = the rest of K2 (road kernels + overflow pass).
 * While enabled, every chunk of frames records events around each class (bounded pool; chunks beyond
 * the pool are not sampled). mld_profile_read synchronises, returns the accumulated milliseconds,
 * launch counts and frames covered per class since the last read, and resets the accumulators. */
int mld_profile_enable(mld_handle* h, int on);
int mld_profile_read(mld_handle* h, double* ms7, int64_t* launches7, int64_t* frames_sampled);
/* frames processed per kernel launch in the batched paths (env MLD_CHUNK_FRAMES, default 16) */
int mld_chunk_frames(const mld_handle* h);
/* frames per fused K1 + gather launch of device-resident non-road sequences (env MLD_FUSE_CHUNK, default 512), or 0 when
 * the fused pipeline is off (MLD_FUSE=0 or another K2 mode): such sequences then use mld_chunk_frames() like the rest */
int mld_fused_chunk_frames(const mld_handle* h);
/* largest neighbour count per feature the kernels were built for */
int mld_neighbor_capacity(void);

/* ---- debug / parity views of the current cloud (NeighborFinderPixel::_img_points_lidar etc.) ---- */
/* H x W row-major (offset x + y*W), RAW point index of the first point (in cloud order) that
 * projects into the pixel with z > 0, -1 = empty (NeighborFinderPixel.cpp:40-55). */
int mld_get_pixel_map(mld_handle* h, int32_t* out_host);
/* neighbours of one feature in the reference's scan order (rows outer, columns inner), RAW indices;
 * *k_out = number found (may exceed cap; only cap are written) (NeighborFinderPixel.cpp:60-95). */
int mld_get_neighbors(mld_handle* h, double u, double v, double scale_w, double scale_h, int32_t* out_raw, int cap,
                      int* k_out);
/* visible flags per raw point (Transform_Cloud_LidarToCamera's cull, DepthEstimator.cpp:184-207):
 * out_visible_host[n] bytes; *n_visible_out = count. */
int mld_get_visible(mld_handle* h, uint8_t* out_visible_host, int64_t* n_visible_out);
/* visible-order views, compacted on the device in cloud order:
 *   point_index_out[j]  = raw index of visible point j                  (_pointIndex, DepthEstimator.cpp:197-207)
 *   image_points_out    = 2 x nvis doubles, column-major (u_j, v_j)     (getPointsCloudImageCs / _points_cs_image_visible)
 *   depth_cam_out[j]    = camera-frame z of visible point j              (getPointDepthCamVisible(j))
 * Any output may be NULL; at most `capacity` entries are written; *n_visible_out = number of visible points. */
int mld_get_visible_points(mld_handle* h, int32_t* point_index_out, double* image_points_out, double* depth_cam_out,
                           int64_t capacity, int64_t* n_visible_out);
/* camera-frame coordinates of raw point i (3 doubles each), _points_cs_camera */
int mld_get_points_camera(mld_handle* h, double* out_host);
/* the same for selected raw indices (getCloudRansacPlane: the ground plane's inliers, DepthEstimator.cpp:294-308); an index
 * outside the cloud yields NaN */
int mld_get_points_camera_indexed(mld_handle* h, const int32_t* idx_host, int64_t n_idx, double* out_host);
/* the three triangle corners CalculateDepthSegmented picked for each feature of the current cloud (camera frame, 9 doubles per
 * feature; getCloudTriangleCorners / _points_triangle_corners, DepthEstimator.cpp:915-926); valid_out[i] = 0 and NaN corners when
 * the feature never reached a corner selection (empty window, no histogram maximum, no triangle) */
int mld_get_triangle_corners(mld_handle* h, const double* uv_host, int F, double* corners_out_host, uint8_t* valid_out_host);

/* ---- synthetic KITTI-shaped input, device generators (bench / tests; include/mld_synth.h has the configuration and the
 * host generators of libmld_synth.so, which agree with these bit for bit) ----
 * frames [frame0, frame0 + nframes) written frame_pitch_points apart */
int mld_synth_points_device(mld_handle* h, const mld_synth_config* c, uint64_t seed, int64_t frame0, int64_t nframes,
                            int64_t frame_pitch_points, float* d_out_xyzi, void* stream);
int mld_synth_features_device(mld_handle* h, const mld_synth_config* c, uint64_t seed, int64_t frame0, int64_t nframes,
                              int F, double* d_out_uv, void* stream);

#ifdef __cplusplus
}
#endif
#endif
