/*
 * mld_oracle.h -- C interface of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 *
 * The oracle is a dependency-free C++17 restatement of the reference's
 * Mono_Lidar::DepthEstimator hot path (monolidar_fusion/src/DepthEstimator.cpp and
 * the helper classes it calls). The reference itself cannot be compiled in this
 * environment (Eigen, PCL, Ceres, OpenCV and catkin are absent), so every Eigen /
 * PCL expression is restated by hand; see the function headers in mld_oracle.cpp
 * for the reference file:line each one follows.
 *
 * Parity status: pinned against the reference itself. oracle/_ref/libmld_ref.so is the
 * reference's own sources compiled on stand-in Eigen/PCL headers (oracle/ref_standin,
 * oracle/ref_bridge.cpp); tests/test_ref_pin.py diffs this restatement against it on
 * seeded inputs (visible set, pixel map, neighbour lists, status codes bit-exact; depths
 * bit-exact on the main path, <= 1e-9 on the M-estimator road path) for every parameter
 * variant, and tests/golden/ref_golden.npz freezes reference outputs for the GPU box.
 * Also pinned by the reference's own KATs: A6 golden vector (test_monolidar_fusion.cpp:
 * 306-374), A4/A5 window property (:82-171), R1 +-0.2 coefficient test (:376-441).
 * NOT pinned: arithmetic inside Eigen/PCL calls (the stand-in's, same assumptions as
 * here), the RANSAC hypothesis stream (PCL is time-seeded), R4 (undefined upstream).
 */
#ifndef MLD_ORACLE_H
#define MLD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Field names follow Mono_Lidar::DepthEstimatorParameters
 * (monolidar_fusion/include/monolidar_fusion/DepthEstimatorParameters.h:12-172).
 * bool fields are int (the yaml loader reads them as (int), DepthEstimatorParameters.cpp:27 ff.).
 * Layout is intentionally identical to mld_params in include/mld_c_api.h so the tests
 * can hand the same bytes to both sides. */
typedef struct orc_params {
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
} orc_params;

/* Ground plane handed to CalculateDepth (GroundPlane interface, RansacPlane.h:38-126):
 * model coefficients a,b,c,d in the LIDAR frame (float, like Eigen::Vector4f) and the
 * inlier set as raw point indices (getInlinersIndex / CheckPointInPlane). */
typedef struct orc_plane {
    float coeffs[4];
    const int32_t* inlier_idx;
    int64_t n_inliers;
} orc_plane;

typedef struct orc_estimator orc_estimator;

void orc_default_params(orc_params* p);   /* C++ struct defaults (DepthEstimatorParameters.h) */
void orc_yaml_params(orc_params* p);      /* values of monolidar_fusion/parameters.yaml with
                                             do_use_depth_segmentation forced to 0 */

orc_estimator* orc_create(const orc_params* p);
void orc_destroy(orc_estimator* e);
void orc_set_num_threads(int n);          /* OpenMP threads of CalculateDepth; <=0 -> all */
int orc_get_max_threads(void);

/* DepthEstimator::Initialize (DepthEstimator.cpp:35-127). T = row-major 3x4 [R|t] lidar->camera. */
int orc_initialize(orc_estimator* e, int W, int H, double f, double cx, double cy, const double* T);

/* DepthEstimator::setInputCloud without the RANSAC step (DepthEstimator.cpp:220-272).
 * pts: n points, stride_floats floats apart (4 for float4, 8 for pcl::PointXYZI), x,y,z first. */
int orc_set_cloud(orc_estimator* e, const float* pts, int64_t n, int stride_floats);

/* DepthEstimator::CalculateDepth (DepthEstimator.cpp:429-600). uv = 2xF column-major (u0,v0,u1,v1,..).
 * plane may be NULL (== ransacPlane nullptr). Returns 0, or <0 for the reference's throw sites. */
int orc_calculate_depth(orc_estimator* e, const double* uv, int F, double* depth, int32_t* status,
                        const orc_plane* plane);

/* ---- debug / parity views ---- */
int64_t orc_visible_count(const orc_estimator* e);
/* _pointIndex (visible -> raw), DepthEstimator.cpp:197-207 */
void orc_get_point_index(const orc_estimator* e, int32_t* out);
/* _points_cs_image_visible (2 x nvis col-major) */
void orc_get_image_points_visible(const orc_estimator* e, double* out);
/* camera-frame points 3 x n col-major (_points_cs_camera) */
void orc_get_points_camera(const orc_estimator* e, double* out);
/* pixel map, row-major H x W (offset x + y*W), VISIBLE indices, -1 = empty (NeighborFinderPixel.cpp:29-58) */
void orc_get_pixel_map_visible(const orc_estimator* e, int32_t* out);
/* same map translated to RAW indices via _pointIndex */
void orc_get_pixel_map_raw(const orc_estimator* e, int32_t* out);
/* neighbours of one feature in scan order as RAW indices; returns k (NeighborFinderPixel.cpp:60-95) */
int orc_get_neighbors(const orc_estimator* e, double u, double v, double scale_w, double scale_h,
                      int32_t* out_raw, int cap);

/* ---- unit-level entry points used to pin the oracle against the reference's own tests ---- */
/* PointHistogram::FilterPointsMinDistBlob (HistogramPointDepth.cpp:15-123): returns 1/0,
 * writes the positions (into the input) of the kept elements. */
int orc_histogram_filter(const double* depths, int n, double bin_width, int min_count,
                         int32_t* out_pos, int* n_out, double* lower, double* higher);
/* NeighborFinderPixel standalone (InitializeLidarProjection + getNeighbors), used by the
 * restated NeigborFinder.findByPixel test: img = 2 x n image coords, cam = 3 x n points. */
int orc_neighbor_finder(int W, int H, int search_w, int search_h, const double* img, const double* cam,
                        int n, double u, double v, int32_t* out_idx, int cap);
/* CameraPinhole::getViewingRays / getImagePoints (camera_pinhole.h:52-97) for single points */
void orc_viewing_ray(int W, int H, double f, double cx, double cy, double u, double v, double* dir3);
int orc_image_point(int W, int H, double f, double cx, double cy, const double* p3, double* uv2);

/* RansacPlane::CalculateInliersPlane (RansacPlane.cpp:41-140) with the PCL pieces restated and a
 * counter-based RNG (seed) instead of PCL's time seed. inlier_idx must hold >= n entries.
 * Returns 0, -1 for ExceptionPclInvalid (<3 points), -2 when no model was found. */
int orc_ransac_plane(const orc_params* p, const float* pts, int64_t n, int stride_floats, uint64_t seed,
                     float* coeffs4, int32_t* inlier_idx, int64_t* n_inliers, int32_t* iterations_out);

#ifdef __cplusplus
}
#endif
#endif
