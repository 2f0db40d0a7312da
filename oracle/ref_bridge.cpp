/*
 * ref_bridge.cpp -- C bridge onto the REFERENCE's own Mono_Lidar::DepthEstimator, compiled from the sources
 * where they lie under /root/reference (never copied) against the stand-in Eigen/PCL/OpenCV headers of
 * oracle/ref_standin. Output: oracle/_ref/libmld_ref.so (git-ignored). TEST INFRASTRUCTURE ONLY.
 *
 * What this pins: the reference's control flow, status precedence, call order, argument order, index
 * bookkeeping (visible <-> raw), first-wins map, window scan order, histogram scan, triangle search,
 * threshold logic, road gate -- all executed by the reference's unmodified code. What it does NOT pin:
 * the arithmetic INSIDE Eigen/PCL calls (dot/cross/normalize/inverse/SVD/RANSAC), which is the stand-in's.
 *
 * The functions mirror the orc_* entry points of mld_oracle.h one to one so tests can diff both sides.
 * Compiled with -fno-access-control: the parity views read private members of the reference classes.
 */
#include <cstdint>
#include <cstring>
#include <memory>
#include <set>
#include <string>

#include "DepthEstimator.h"
#include "HistogramPointDepth.h"
#include "NeighborFinderPixel.h"
#include "PlaneEstimationLeastSquares.h"
#include "RansacPlane.h"
#include "mld_oracle.h"

namespace Mono_Lidar {
// RoadDepthEstimatorLeastSquares links against this; the reference's Ceres implementation is undefined
// behaviour and is not part of oracle/_ref.
PlaneEstimationLeastSquares::PlaneEstimationLeastSquares() {}
bool PlaneEstimationLeastSquares::EstimatePlane(const VecOfVec3d&, Eigen::Vector3d&, double&) {
    throw std::runtime_error("oracle/_ref: PlaneEstimationLeastSquares (Ceres) is not built");
}
}  // namespace Mono_Lidar

namespace {
using namespace Mono_Lidar;
using Cloud = DepthEstimator::Cloud;

// a plane computed elsewhere, handed in through the reference's GroundPlane interface (RansacPlane.h:38-126)
struct InjectedPlane : GroundPlane {
    InjectedPlane(const orc_plane& p) {
        for (int i = 0; i < 4; i++) _modelCoeffs[i] = p.coeffs[i];
        _inliersIndex.assign(p.inlier_idx, p.inlier_idx + p.n_inliers);
        for (int i : _inliersIndex) _pointIsInPlane.insert(std::pair<int, bool>(i, true));
        is_segmented_ = true;
    }
    void CalculateInliersPlane(const Cloud::ConstPtr&, double, double) override {}
};

std::shared_ptr<DepthEstimatorParameters> to_ref_params(const orc_params& p) {
    auto q = std::make_shared<DepthEstimatorParameters>();
#define CP(f) q->f = p.f
    CP(neighbor_search_mode); CP(pixelarea_search_witdh); CP(pixelarea_search_height); CP(radiusSearch_count_min);
    CP(do_use_histogram_segmentation); CP(histogram_segmentation_min_pointcount); CP(histogram_segmentation_bin_witdh);
    CP(do_use_depth_segmentation); CP(treshold_depth_enabled); CP(treshold_depth_mode); CP(treshold_depth_max);
    CP(treshold_depth_min); CP(treshold_depth_local_enabled); CP(treshold_depth_local_mode);
    CP(treshold_depth_local_valuetype); CP(treshold_depth_local_value); CP(do_use_PCA); CP(pca_debug);
    CP(pca_treshold_3_abs_min); CP(pca_treshold_3_2_rel_max); CP(pca_treshold_2_1_rel_min); CP(do_use_ransac_plane);
    CP(ransac_plane_max_iterations); CP(ransac_plane_distance_treshold); CP(ransac_plane_min_z); CP(ransac_plane_max_z);
    CP(ransac_plane_use_refinement); CP(ransac_plane_use_camx_treshold); CP(ransac_plane_refinement_treshold);
    CP(ransac_plane_treshold_camx); CP(ransac_plane_point_distance_treshold); CP(ransac_plane_probability);
    CP(plane_estimator_use_triangle_maximation); CP(plane_estimator_use_leastsquares); CP(plane_estimator_use_mestimator);
    CP(do_use_cut_behind_camera); CP(plane_estimator_z_x_min_relation); CP(do_use_triangle_size_maximation);
    CP(do_check_triangleplanar_condition); CP(triangleplanar_crossnorm_treshold);
    CP(viewray_plane_orthoganality_treshold); CP(set_all_depths_to_zero);
#undef CP
    q->do_publish_points = false;  // debug clouds only
    q->do_depth_calc_statistics = false;
    q->do_debug_singleFeatures = false;
    return q;
}
void from_ref_params(const DepthEstimatorParameters& q, orc_params& p) {
    std::memset(&p, 0, sizeof(p));
#define CP(f) p.f = q.f
    CP(neighbor_search_mode); CP(pixelarea_search_witdh); CP(pixelarea_search_height); CP(radiusSearch_count_min);
    CP(do_use_histogram_segmentation); CP(histogram_segmentation_min_pointcount); CP(histogram_segmentation_bin_witdh);
    CP(do_use_depth_segmentation); CP(treshold_depth_enabled); CP(treshold_depth_mode); CP(treshold_depth_max);
    CP(treshold_depth_min); CP(treshold_depth_local_enabled); CP(treshold_depth_local_mode);
    CP(treshold_depth_local_valuetype); CP(treshold_depth_local_value); CP(do_use_PCA); CP(pca_debug);
    CP(pca_treshold_3_abs_min); CP(pca_treshold_3_2_rel_max); CP(pca_treshold_2_1_rel_min); CP(do_use_ransac_plane);
    CP(ransac_plane_max_iterations); CP(ransac_plane_distance_treshold); CP(ransac_plane_min_z); CP(ransac_plane_max_z);
    CP(ransac_plane_use_refinement); CP(ransac_plane_use_camx_treshold); CP(ransac_plane_refinement_treshold);
    CP(ransac_plane_treshold_camx); CP(ransac_plane_point_distance_treshold); CP(ransac_plane_probability);
    CP(plane_estimator_use_triangle_maximation); CP(plane_estimator_use_leastsquares); CP(plane_estimator_use_mestimator);
    CP(do_use_cut_behind_camera); CP(plane_estimator_z_x_min_relation); CP(do_use_triangle_size_maximation);
    CP(do_check_triangleplanar_condition); CP(triangleplanar_crossnorm_treshold);
    CP(viewray_plane_orthoganality_treshold); CP(set_all_depths_to_zero);
#undef CP
}

Cloud::Ptr make_cloud(const float* pts, int64_t n, int stride_floats) {
    Cloud::Ptr c(new Cloud());
    c->points.resize(size_t(n));
    for (int64_t i = 0; i < n; i++) {
        const float* f = pts + i * stride_floats;
        auto& p = c->points[size_t(i)];
        p.x = f[0]; p.y = f[1]; p.z = f[2];
        p.intensity = stride_floats >= 8 ? f[4] : (stride_floats >= 4 ? f[3] : 0.f);
    }
    c->width = uint32_t(n); c->height = 1;
    return c;
}

thread_local std::string g_err;
template <class F>
int guarded(F&& f) {
    try { f(); return 0; }
    catch (const char* s) { g_err = s; return -1; }
    catch (const std::string& s) { g_err = s; return -2; }
    catch (const GroundPlane::ExceptionPclInvalid& e) { g_err = e.what(); return -4; }
    catch (const std::exception& e) { g_err = e.what(); return -3; }
}
}  // namespace

struct ref_estimator {
    DepthEstimator est;
    std::shared_ptr<CameraPinhole> cam;
    GroundPlane::Ptr plane;
    Cloud::Ptr cloud;
    int W = 0, H = 0;
};

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }

// DepthEstimator::InitConfig(shared_ptr<DepthEstimatorParameters>) (DepthEstimator.cpp:141-154)
ref_estimator* ref_create(const orc_params* p) {
    auto* e = new ref_estimator();
    e->est.InitConfig(to_ref_params(*p), false);
    return e;
}
// DepthEstimator::InitConfig(path) through the reference's own yaml loader (DepthEstimatorParameters.cpp:16-114)
ref_estimator* ref_create_from_yaml(const char* path, orc_params* loaded) {
    auto* e = new ref_estimator();
    int rc = guarded([&] { e->est.InitConfig(std::string(path), false); });
    if (rc) { delete e; return nullptr; }
    if (loaded) from_ref_params(*e->est.getParameters(), *loaded);
    return e;
}
void ref_default_params(orc_params* p) { DepthEstimatorParameters q; from_ref_params(q, *p); }
void ref_destroy(ref_estimator* e) { delete e; }

int ref_initialize(ref_estimator* e, int W, int H, double f, double cx, double cy, const double* T) {
    return guarded([&] {
        e->cam = std::make_shared<CameraPinhole>(W, H, f, cx, cy);
        Eigen::Affine3d tf = Eigen::Affine3d::Identity();
        for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) tf.matrix()(i, j) = T[i * 4 + j];
        e->W = W; e->H = H;
        e->est.Initialize(e->cam, tf);
    });
}

// DepthEstimator::setInputCloud (DepthEstimator.cpp:220-312). plane != NULL: injected GroundPlane (already
// segmented); plane == NULL: nullptr in, i.e. the reference creates and fits its RansacPlane when
// do_use_ransac_plane is set (stand-in PCL; seed through ref_set_seed).
int ref_set_cloud(ref_estimator* e, const float* pts, int64_t n, int stride_floats, const orc_plane* plane) {
    return guarded([&] {
        e->cloud = make_cloud(pts, n, stride_floats);
        e->plane = plane ? GroundPlane::Ptr(new InjectedPlane(*plane)) : GroundPlane::Ptr();
        e->est.setInputCloud(e->cloud, e->plane);
    });
}
// with_plane == 0 passes a null GroundPlane::Ptr to CalculateDepth (non-road path only)
int ref_calculate_depth(ref_estimator* e, const double* uv, int F, double* depth, int32_t* status, int with_plane) {
    return guarded([&] {
        Eigen::Matrix2Xd feats(2, F);
        for (int i = 0; i < F; i++) { feats(0, i) = uv[2 * i]; feats(1, i) = uv[2 * i + 1]; }
        Eigen::VectorXd d;
        Eigen::VectorXi s;
        GroundPlane::Ptr none;
        e->est.CalculateDepth(feats, d, s, with_plane ? e->plane : none);
        for (int i = 0; i < F; i++) { depth[i] = d[i]; status[i] = s[i]; }
    });
}
// the single-feature overload (DepthEstimator.cpp:491-600)
int ref_calculate_depth_single(ref_estimator* e, double u, double v, double* depth, int32_t* status, int with_plane) {
    return guarded([&] {
        GroundPlane::Ptr none;
        auto r = e->est.CalculateDepth(Eigen::Vector2d(u, v), with_plane ? e->plane : none);
        *depth = r.second; *status = int32_t(r.first);
    });
}
int ref_get_plane(ref_estimator* e, float* coeffs4, int32_t* inlier_idx, int64_t* n_inliers) {
    if (!e->plane) return -1;
    for (int i = 0; i < 4; i++) coeffs4[i] = e->plane->getModelCoeffs()[i];
    const auto& idx = e->plane->getInlinersIndex();
    *n_inliers = int64_t(idx.size());
    if (inlier_idx) std::copy(idx.begin(), idx.end(), inlier_idx);
    return 0;
}

// ---- parity views (private members; -fno-access-control) ----
int64_t ref_visible_count(ref_estimator* e) { return int64_t(e->est._points._pointIndex.size()); }
void ref_get_point_index(ref_estimator* e, int32_t* out) { const auto& v = e->est._points._pointIndex; std::copy(v.begin(), v.end(), out); }
void ref_get_image_points_visible(ref_estimator* e, double* out) {
    Eigen::Matrix2Xd m;
    e->est.getPointsCloudImageCs(m);
    std::copy(m.data(), m.data() + m.size(), out);
}
void ref_get_points_camera(ref_estimator* e, double* out) { const auto& m = e->est._points._points_cs_camera; std::copy(m.data(), m.data() + m.size(), out); }
double ref_get_point_depth_cam_visible(ref_estimator* e, int i) { return e->est.getPointDepthCamVisible(i); }
// NeighborFinderPixel::_img_points_lidar is (W, H) indexed (x, y); out is row-major H x W, visible indices
void ref_get_pixel_map_visible(ref_estimator* e, int32_t* out) {
    auto nf = std::dynamic_pointer_cast<NeighborFinderPixel>(e->est._neighborFinder);
    for (int y = 0; y < e->H; y++) for (int x = 0; x < e->W; x++) out[size_t(y) * e->W + x] = nf->_img_points_lidar(x, y);
}
void ref_get_pixel_map_raw(ref_estimator* e, int32_t* out) {
    ref_get_pixel_map_visible(e, out);
    const auto& pi = e->est._points._pointIndex;
    for (size_t i = 0; i < size_t(e->W) * e->H; i++) if (out[i] >= 0) out[i] = pi[size_t(out[i])];
}
int ref_get_neighbors(ref_estimator* e, double u, double v, double scale_w, double scale_h, int32_t* out_raw, int cap) {
    std::vector<int> cut;
    e->est._neighborFinder->getNeighbors(Eigen::Vector2d(u, v), e->est._points._points_cs_camera, e->est._points._pointIndex, cut,
                                         nullptr, float(scale_w), float(scale_h));
    int k = 0;
    for (int c : cut) { if (k < cap) out_raw[k] = e->est._points._pointIndex[size_t(c)]; k++; }
    return k;
}

// ---- unit-level entry points ----
int ref_histogram_filter(const double* depths, int n, double bin_width, int min_count, int32_t* out_pos, int* n_out,
                         double* lower, double* higher) {
    VecOfVec3d in, out;
    std::vector<int> idx, out_idx;
    Eigen::VectorXd d(n);
    for (int i = 0; i < n; i++) { in.push_back(Eigen::Vector3d(0, 0, depths[i])); idx.push_back(i); d[i] = depths[i]; }
    bool ok = PointHistogram::FilterPointsMinDistBlob(in, idx, d, bin_width, min_count, out, out_idx, *lower, *higher);
    *n_out = int(out_idx.size());
    for (size_t i = 0; i < out_idx.size(); i++) out_pos[i] = out_idx[i];
    return ok ? 1 : 0;
}
int ref_neighbor_finder(int W, int H, int search_w, int search_h, const double* img, const double* cam, int n, double u,
                        double v, int32_t* out_idx, int cap) {
    NeighborFinderPixel nf(W, H, search_w, search_h);
    Eigen::Matrix2Xd im(2, n);
    Eigen::Matrix3Xd cm(3, n);
    std::vector<int> pi(static_cast<size_t>(n));
    for (int i = 0; i < n; i++) { im(0, i) = img[2 * i]; im(1, i) = img[2 * i + 1]; for (int r = 0; r < 3; r++) cm(r, i) = cam[3 * i + r]; pi[size_t(i)] = i; }
    nf.InitializeLidarProjection(im, cm, pi);
    std::vector<int> cut;
    nf.getNeighbors(Eigen::Vector2d(u, v), cm, pi, cut, nullptr, 1.0f, 1.0f);
    int k = 0;
    for (int c : cut) { if (k < cap) out_idx[k] = c; k++; }
    return k;
}
void ref_viewing_ray(int W, int H, double f, double cx, double cy, double u, double v, double* dir3) {
    CameraPinhole cam(W, H, f, cx, cy);
    Eigen::Vector3d sp, dir;
    cam.getViewingRays(Eigen::Vector2d(u, v), sp, dir);
    for (int i = 0; i < 3; i++) dir3[i] = dir[i];
}
int ref_image_point(int W, int H, double f, double cx, double cy, const double* p3, double* uv2) {
    CameraPinhole cam(W, H, f, cx, cy);
    Eigen::Matrix<double, 3, 1> p(p3[0], p3[1], p3[2]);
    Eigen::Matrix<double, 2, 1> q;
    auto in = cam.getImagePoints(p, q);
    uv2[0] = q[0]; uv2[1] = q[1];
    return in[0] ? 1 : 0;
}

void ref_set_seed(unsigned s) { pcl::standin::seed() = s; }
// RansacPlane::CalculateInliersPlane (RansacPlane.cpp:41-140) run by the reference's own code on stand-in PCL
int ref_ransac_plane(const orc_params* p, const float* pts, int64_t n, int stride_floats, unsigned seed, float* coeffs4,
                     int32_t* inlier_idx, int64_t* n_inliers) {
    return guarded([&] {
        pcl::standin::seed() = seed;
        RansacPlane rp(to_ref_params(*p));
        Cloud::Ptr c = make_cloud(pts, n, stride_floats);
        rp.CalculateInliersPlane(c, p->ransac_plane_min_z, p->ransac_plane_max_z);
        for (int i = 0; i < 4; i++) coeffs4[i] = rp.getModelCoeffs()[i];
        const auto& idx = rp.getInlinersIndex();
        *n_inliers = int64_t(idx.size());
        std::copy(idx.begin(), idx.end(), inlier_idx);
    });
}
// SemanticPlane::CalculateInliersPlane (RansacPlane.cpp:195-274): labels = H x W u8 row-major; T = cam<-lidar 3x4
int ref_semantic_plane(const uint8_t* labels, int W, int H, double f, double cu, double cv, const double* T,
                       const int32_t* ground_labels, int n_labels, double inlier_threshold, const float* pts, int64_t n,
                       int stride_floats, float* coeffs4, int32_t* inlier_idx, int64_t* n_inliers) {
    return guarded([&] {
        SemanticPlane::Camera cam;
        cam.f = f; cam.cu = cu; cam.cv = cv;
        cam.transform_cam_lidar = Eigen::Affine3d::Identity();
        for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) cam.transform_cam_lidar.matrix()(i, j) = T[i * 4 + j];
        cv::Mat img(H, W, labels);
        SemanticPlane sp(img, cam, std::set<int>(ground_labels, ground_labels + n_labels), inlier_threshold);
        Cloud::Ptr c = make_cloud(pts, n, stride_floats);
        sp.CalculateInliersPlane(Cloud::ConstPtr(c));
        for (int i = 0; i < 4; i++) coeffs4[i] = sp.getModelCoeffs()[i];
        const auto& idx = sp.getInlinersIndex();
        *n_inliers = int64_t(idx.size());
        std::copy(idx.begin(), idx.end(), inlier_idx);
    });
}

}  // extern "C"
