// Host shim: the reference's Mono_Lidar::DepthEstimator / RansacPlane member functions implemented
// as buffer plumbing over the C ABI. Reference behaviour mirrored per method:
//   Initialize     monolidar_fusion/src/DepthEstimator.cpp:35-127
//   InitConfig     :129-154
//   setInputCloud  :220-312
//   CalculateDepth :404-488 (+ the single-point overload :491-600)
//   RansacPlane::CalculateInliersPlane  monolidar_fusion/src/RansacPlane.cpp:41-140
#include "monolidar_fusion/DepthEstimator.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <vector>

namespace Mono_Lidar {

namespace {

// host view of a GroundPlane for the C ABI; keeps the index buffer alive for the call
struct PlaneView {
    mld_plane c;
    std::vector<int32_t> idx;
};

void fill_plane(PlaneView& v, Eigen::Vector4f& coeffs, const std::vector<int>& inliers, bool segmented, int64_t capacity) {
    std::memset(&v.c, 0, sizeof(v.c));
    for (int i = 0; i < 4; i++) v.c.coeffs[i] = coeffs[i];
    v.idx.assign(inliers.begin(), inliers.end());
    if ((int64_t)v.idx.size() < capacity) v.idx.resize((size_t)capacity);
    v.c.inlier_idx = v.idx.empty() ? nullptr : v.idx.data();
    v.c.n_inliers = (int64_t)inliers.size();
    v.c.inlier_capacity = (int64_t)v.idx.size();
    v.c.segmented = segmented ? 1 : 0;
}

}  // namespace

void DepthEstimatorParameters::print() {
    std::cout << "DepthEstimator parameters: " << std::endl << std::endl;
    std::cout << "pixelarea_search_witdh: " << pixelarea_search_witdh << std::endl;
    std::cout << "pixelarea_search_height: " << pixelarea_search_height << std::endl;
    std::cout << "radiusSearch_count_min: " << radiusSearch_count_min << std::endl;
    std::cout << "do_use_histogram_segmentation: " << do_use_histogram_segmentation << std::endl;
    std::cout << "histogram_segmentation_bin_witdh: " << histogram_segmentation_bin_witdh << std::endl;
    std::cout << "histogram_segmentation_min_pointcount: " << histogram_segmentation_min_pointcount << std::endl;
    std::cout << "do_use_ransac_plane: " << do_use_ransac_plane << std::endl;
    std::cout << "viewray_plane_orthoganality_treshold: " << viewray_plane_orthoganality_treshold << std::endl;
}

// ---------------------------------------------------------------------------------------------
RansacPlane::RansacPlane() {
    params_.ransac_plane_distance_treshold = 0;
    params_.ransac_plane_max_iterations = 0;
    params_.ransac_plane_probability = 0.999;
    params_.ransac_plane_refinement_treshold = 10000;
    params_.ransac_plane_use_refinement = 0;
}

RansacPlane::RansacPlane(const std::shared_ptr<DepthEstimatorParameters>& parameters) : params_(*parameters) {}

RansacPlane::~RansacPlane() { mld_destroy(handle_); }

void RansacPlane::CalculateInliersPlane(const Cloud::ConstPtr& pointCloud) { CalculateInliersPlane(pointCloud, -1000, 1000); }

void RansacPlane::CalculateInliersPlane(const Cloud::ConstPtr& pointCloud, double min_z, double max_z) {
    if (pointCloud->points.size() < 3) throw ExceptionPclInvalid();
    if (!handle_) {
        if (mld_create(&params_, -1, &handle_) != MLD_OK) throw std::runtime_error(mld_last_error(nullptr));
    }
    // min_z / max_z are call arguments in the reference (RansacPlane.cpp:41); forward them through the parameter block
    DepthEstimatorParameters p = params_;
    p.ransac_plane_min_z = min_z;
    p.ransac_plane_max_z = max_z;
    if (p.ransac_plane_min_z != params_.ransac_plane_min_z || p.ransac_plane_max_z != params_.ransac_plane_max_z) {
        mld_destroy(handle_);
        handle_ = nullptr;
        params_ = p;
        if (mld_create(&params_, -1, &handle_) != MLD_OK) throw std::runtime_error(mld_last_error(nullptr));
    }
    PlaneView v;
    std::vector<int> none;
    fill_plane(v, _modelCoeffs, none, false, (int64_t)pointCloud->points.size());
    int rc = mld_estimate_ground_plane(handle_, pointCloud->points.data(), (int64_t)pointCloud->points.size(), (int)sizeof(Point), seed_,
                                       &v.c, nullptr);
    if (rc == MLD_ERR_PCL_INVALID) throw ExceptionPclInvalid();
    if (rc != MLD_OK) throw std::runtime_error(mld_last_error(handle_));
    for (int i = 0; i < 4; i++) _modelCoeffs[i] = v.c.coeffs[i];
    _inliersIndex.assign(v.idx.begin(), v.idx.begin() + v.c.n_inliers);
    _pointIsInPlane.clear();
    for (const auto& index : _inliersIndex) _pointIsInPlane.insert(std::pair<int, bool>(index, true));
    is_segmented_ = true;
}

// ---------------------------------------------------------------------------------------------
// SemanticPlane (monolidar_fusion/src/RansacPlane.cpp:159-274) on mld_semantic_ground_plane
SemanticPlane::SemanticPlane(const cv::Mat& img, Camera cam, std::set<int> groundplane_label, double inlier_threshold)
        : rows_(img.rows), cols_(img.cols), cam_(cam), inlier_threshold_(inlier_threshold), groundplane_label_(groundplane_label) {
    labels_.resize((size_t)rows_ * (size_t)cols_);
    for (int y = 0; y < rows_; y++) std::memcpy(labels_.data() + (size_t)y * (size_t)cols_, img.ptr<unsigned char>(y), (size_t)cols_);
}

SemanticPlane::~SemanticPlane() { mld_destroy(handle_); }

void SemanticPlane::CalculateInliersPlane(const Cloud::ConstPtr& cloud) {
    if (!handle_) {
        DepthEstimatorParameters p;
        if (mld_create(&p, -1, &handle_) != MLD_OK) throw std::runtime_error(mld_last_error(nullptr));
    }
    double T[12];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) T[r * 4 + c] = cam_.transform_cam_lidar.matrix()(r, c);
    std::vector<int32_t> labels(groundplane_label_.begin(), groundplane_label_.end());
    PlaneView v;
    std::vector<int> none;
    const int64_t n = (int64_t)cloud->points.size();
    fill_plane(v, _modelCoeffs, none, false, n);
    int rc = mld_semantic_ground_plane(handle_, cloud->points.data(), n, (int)sizeof(Point), labels_.data(), cols_, rows_, cam_.f, cam_.cu,
                                       cam_.cv, T, labels.data(), (int)labels.size(), inlier_threshold_, &v.c);
    if (rc == MLD_ERR_PCL_INVALID) throw ExceptionPclInvalid();
    if (rc != MLD_OK) throw std::runtime_error(mld_last_error(handle_));
    for (int i = 0; i < 4; i++) _modelCoeffs[i] = v.c.coeffs[i];
    _inliersIndex.assign(v.idx.begin(), v.idx.begin() + v.c.n_inliers);
    _pointIsInPlane.clear();
    for (const auto& index : _inliersIndex) _pointIsInPlane.insert(std::pair<int, bool>(index, true));
    is_segmented_ = true;
}

// ---------------------------------------------------------------------------------------------
DepthEstimator::DepthEstimator() {
    for (int s = 1; s <= 15; s++) DepthResultTypeMap[(DepthResultType)s] = mld_status_name(s);
}

DepthEstimator::~DepthEstimator() { mld_destroy(_handle); }

void DepthEstimator::rethrow(int rc) {
    const char* msg = mld_last_error(_handle);
    switch (rc) {
        case MLD_ERR_PCL_INVALID: throw GroundPlane::ExceptionPclInvalid();
        case MLD_ERR_REGION_GROWING: throw std::runtime_error("DepthEstimator: Region growing not supported!");
        case MLD_ERR_BAD_SEARCH_MODE: throw std::string(msg);
        case MLD_ERR_NOT_CONFIGURED: throw "Call 'InitConfig' before calling 'Initialize'.";
        case MLD_ERR_NOT_INITIALIZED: throw "call of 'setInputCloud' without 'initialize'";
        case MLD_ERR_NO_CLOUD: throw "call of 'CalculateDepth' without 'SetInputCloud'";
        case MLD_ERR_NO_ROAD_ESTIMATOR: throw "No road depth estimator selected.";
        default: throw std::runtime_error(std::string("mld: ") + msg);
    }
}

bool DepthEstimator::InitConfig(const std::string& filePath, const bool printparams) {
    auto p = std::make_shared<DepthEstimatorParameters>();
    p->fromFile(filePath);
    return InitConfig(p, printparams);
}

bool DepthEstimator::InitConfig(std::shared_ptr<DepthEstimatorParameters> parameters, const bool printparams) {
    _parameters = parameters ? parameters : std::make_shared<DepthEstimatorParameters>();
    if (printparams) _parameters->print();
    mld_destroy(_handle);
    _handle = nullptr;
    int rc = mld_create(_parameters.get(), -1, &_handle);
    if (rc != MLD_OK) throw std::runtime_error(std::string("mld: ") + mld_last_error(nullptr));
    mld_set_statistics(_handle, _parameters->do_depth_calc_statistics ? 1 : 0);
    std::memcpy(&_paramsOnDevice, static_cast<const mld_params*>(_parameters.get()), sizeof(mld_params));
    _isInitializedConfig = true;
    _isInitialized = false;
    _isInitializedPointCloud = false;
    return true;
}

bool DepthEstimator::Initialize(const std::shared_ptr<CameraPinhole>& camera, const Eigen::Affine3d& transform_lidar_to_cam) {
    if (!_isInitializedConfig) throw "Call 'InitConfig' before calling 'Initialize'.";
    // The reference builds its modules from the live parameter block HERE (DepthEstimator.cpp:46-127); the device handle holds the
    // block as it was at InitConfig. A caller that changed it through the shared pointer in between gets the new values, like upstream.
    if (std::memcmp(static_cast<const mld_params*>(_parameters.get()), &_paramsOnDevice, sizeof(mld_params)) != 0) InitConfig(_parameters, false);
    _camera = camera;
    _transform_lidar_to_cam = transform_lidar_to_cam;
    int W, H;
    _camera->getImageSize(W, H);
    double T[12];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) T[r * 4 + c] = transform_lidar_to_cam.matrix()(r, c);
    int rc = mld_initialize(_handle, W, H, camera->focalLength(), camera->principalPointX(), camera->principalPointY(), T);
    if (rc != MLD_OK) rethrow(rc);
    _isInitialized = true;
    return true;
}

void DepthEstimator::setInputCloud(const Cloud::ConstPtr& cloud, GroundPlane::Ptr& groundPlane) {
    if (!_isInitialized) throw "call of 'setInputCloud' without 'initialize'";
    const int64_t n = (int64_t)cloud->points.size();
    int rc;
    if (_parameters->do_use_ransac_plane) {
        if (groundPlane == nullptr) groundPlane = std::make_shared<RansacPlane>(_parameters);  // DepthEstimator.cpp:275-278
        if (!groundPlane->isSegmented()) {
            if (dynamic_cast<RansacPlane*>(groundPlane.get()) != nullptr) {
                // the RANSAC fit shares the H2D copy of the cloud with the projection
                if (n < 3) throw GroundPlane::ExceptionPclInvalid();
                PlaneView v;
                std::vector<int> none;
                fill_plane(v, groundPlane->_modelCoeffs, none, false, n);
                rc = mld_set_cloud(_handle, cloud->points.data(), n, (int)sizeof(Point), &v.c, _ransacSeed);
                if (rc != MLD_OK) rethrow(rc);
                for (int i = 0; i < 4; i++) groundPlane->_modelCoeffs[i] = v.c.coeffs[i];
                groundPlane->_inliersIndex.assign(v.idx.begin(), v.idx.begin() + v.c.n_inliers);
                groundPlane->_pointIsInPlane.clear();
                for (const auto& index : groundPlane->_inliersIndex) groundPlane->_pointIsInPlane.insert(std::pair<int, bool>(index, true));
                groundPlane->is_segmented_ = true;
                _pointCount = n;
                _isInitializedPointCloud = true;
                _groundInliers = groundPlane->_inliersIndex;
                _depthCamVisibleValid = false;
                _residentCloud = cloud;
                return;
            }
            // any other GroundPlane segments itself (SemanticPlane: on the GPU through mld_semantic_ground_plane)
            groundPlane->CalculateInliersPlane(cloud, _parameters->ransac_plane_min_z, _parameters->ransac_plane_max_z);
        }
    }
    rc = mld_set_cloud(_handle, cloud->points.data(), n, (int)sizeof(Point), nullptr, 0);
    if (rc != MLD_OK) rethrow(rc);
    _pointCount = n;
    _isInitializedPointCloud = true;
    _depthCamVisibleValid = false;
    _residentCloud = cloud;
    _groundInliers.clear();  // _points_groundplane is rebuilt per cloud, from the plane's inliers (DepthEstimator.cpp:234, :294-308)
    if (_parameters->do_use_ransac_plane && groundPlane != nullptr) _groundInliers = groundPlane->_inliersIndex;
}

void DepthEstimator::CalculateDepth(const Cloud::ConstPtr& pointCloud, const Eigen::Matrix2Xd& points_image_cs,
                                    Eigen::VectorXd& points_depths, GroundPlane::Ptr& ransacPlane) {
    setInputCloud(pointCloud, ransacPlane);
    CalculateDepth(points_image_cs, points_depths, ransacPlane);
}

void DepthEstimator::CalculateDepth(const Cloud::ConstPtr& pointCloud, const Eigen::Matrix2Xd& points_image_cs,
                                    Eigen::VectorXd& points_depths, Eigen::VectorXi& resultType, GroundPlane::Ptr& ransacPlane) {
    setInputCloud(pointCloud, ransacPlane);
    CalculateDepth(points_image_cs, points_depths, resultType, ransacPlane);
}

void DepthEstimator::CalculateDepth(const Eigen::Matrix2Xd& featurePoints_image_cs, Eigen::VectorXd& points_depths,
                                    const GroundPlane::Ptr& ransacPlane) {
    Eigen::VectorXi depthTypes(featurePoints_image_cs.cols());
    CalculateDepth(featurePoints_image_cs, points_depths, depthTypes, ransacPlane);
}

void DepthEstimator::CalculateDepth(const Eigen::Matrix2Xd& featurePoints_image_cs, Eigen::VectorXd& points_depths,
                                    Eigen::VectorXi& resultType, const GroundPlane::Ptr& ransacPlane) {
    if (!_isInitializedPointCloud) throw "call of 'CalculateDepth' without 'SetInputCloud'";
    int imgPointCount = featurePoints_image_cs.cols();
    points_depths.resize(imgPointCount);
    resultType.resize(imgPointCount);
    static_assert(sizeof(int) == sizeof(int32_t), "Eigen::VectorXi holds 32-bit ints");
    PlaneView v;
    const mld_plane* pl = nullptr;
    if (ransacPlane != nullptr) {
        fill_plane(v, ransacPlane->getModelCoeffs(), ransacPlane->getInlinersIndex(), ransacPlane->isSegmented(), 0);
        pl = &v.c;
    }
    int rc = mld_calculate_depth(_handle, featurePoints_image_cs.data(), imgPointCount, points_depths.data(),
                                 reinterpret_cast<int32_t*>(resultType.data()), pl);
    if (rc != MLD_OK) rethrow(rc);
    _lastFeatures.assign(featurePoints_image_cs.data(), featurePoints_image_cs.data() + 2 * (size_t)imgPointCount);
    if (_parameters->do_depth_calc_statistics) {  // DepthEstimator.cpp:445-446, :482-485
        int64_t hist[21];
        long long h2[21];
        rc = mld_last_status_histogram(_handle, hist);
        if (rc != MLD_OK) rethrow(rc);
        for (int i = 0; i < 21; i++) h2[i] = (long long)hist[i];
        _depthCalcStats.SetFromHistogram(h2, imgPointCount);
    }
}

std::pair<DepthResultType, double> DepthEstimator::CalculateDepth(const Eigen::Vector2d& featurePoint_image_cs,
                                                                  const GroundPlane::Ptr& ransacPlane) {
    Eigen::Matrix2Xd f(2, 1);
    f(0, 0) = featurePoint_image_cs[0];
    f(1, 0) = featurePoint_image_cs[1];
    Eigen::VectorXd d;
    Eigen::VectorXi t;
    CalculateDepth(f, d, t, ransacPlane);
    return std::pair<DepthResultType, double>((DepthResultType)t(0), d(0));
}

void DepthEstimator::CalculateDepthPair(const Cloud::ConstPtr& cloudLast, const Eigen::Matrix2Xd& featuresLast, Eigen::VectorXd& depthsLast,
                                        GroundPlane::Ptr& planeLast, const Cloud::ConstPtr& cloudCur, const Eigen::Matrix2Xd& featuresCur,
                                        Eigen::VectorXd& depthsCur, GroundPlane::Ptr& planeCur) {
    if (!_isInitialized) throw "call of 'setInputCloud' without 'initialize'";
    depthsLast.resize(featuresLast.cols());
    depthsCur.resize(featuresCur.cols());
    const bool road = _parameters->do_use_ransac_plane != 0;
    GroundPlane::Ptr* planes[2] = {&planeLast, &planeCur};
    const Cloud::ConstPtr* clouds[2] = {&cloudLast, &cloudCur};
    PlaneView views[2];
    mld_plane* cp[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; i++) {
        if (!road || *clouds[i] == nullptr) continue;
        if (*planes[i] == nullptr) *planes[i] = std::make_shared<RansacPlane>(_parameters);  // DepthEstimator.cpp:275-278
        GroundPlane& gp = **planes[i];
        if (!gp.isSegmented() && dynamic_cast<RansacPlane*>(&gp) == nullptr)  // e.g. SemanticPlane: segments itself
            gp.CalculateInliersPlane(*clouds[i], _parameters->ransac_plane_min_z, _parameters->ransac_plane_max_z);
        fill_plane(views[i], gp._modelCoeffs, gp._inliersIndex, gp.isSegmented(), gp.isSegmented() ? 0 : (int64_t)(*clouds[i])->points.size());
        cp[i] = &views[i].c;
    }
    const bool have_last = cloudLast != nullptr;
    int rc;
    if (have_last && cloudLast == _residentCloud && mld_has_resident_cloud(_handle)) {
        // walking a sequence: the previous cloud is the one uploaded as current by the last call and is still on the device with its
        // pixel map (the shim holds the shared_ptr, so the identity test cannot be fooled by a recycled address)
        rc = mld_calculate_depth_pair_resident(_handle, featuresLast.data(), featuresLast.cols(), depthsLast.data(), nullptr, cp[0],
                                               cloudCur->points.data(), (int64_t)cloudCur->points.size(), featuresCur.data(), featuresCur.cols(),
                                               depthsCur.data(), nullptr, cp[1], (int)sizeof(Point), _ransacSeed);
    } else {
        rc = mld_calculate_depth_pair(_handle, have_last ? cloudLast->points.data() : nullptr, have_last ? (int64_t)cloudLast->points.size() : 0,
                                      featuresLast.data(), featuresLast.cols(), depthsLast.data(), nullptr, cp[0], cloudCur->points.data(),
                                      (int64_t)cloudCur->points.size(), featuresCur.data(), featuresCur.cols(), depthsCur.data(), nullptr, cp[1],
                                      (int)sizeof(Point), _ransacSeed);
    }
    if (rc != MLD_OK) rethrow(rc);
    _residentCloud = cloudCur;
    for (int i = 0; i < 2; i++) {
        if (!cp[i] || (*planes[i])->isSegmented()) continue;
        GroundPlane& gp = **planes[i];
        for (int q = 0; q < 4; q++) gp._modelCoeffs[q] = cp[i]->coeffs[q];
        gp._inliersIndex.assign(views[i].idx.begin(), views[i].idx.begin() + cp[i]->n_inliers);
        gp._pointIsInPlane.clear();
        for (const auto& index : gp._inliersIndex) gp._pointIsInPlane.insert(std::pair<int, bool>(index, true));
        gp.is_segmented_ = true;
    }
    _pointCount = (long long)cloudCur->points.size();
    _isInitializedPointCloud = true;
    _depthCamVisibleValid = false;
    _groundInliers.clear();
    if (road && planeCur != nullptr) _groundInliers = planeCur->_inliersIndex;
    _lastFeatures.assign(featuresCur.data(), featuresCur.data() + 2 * (size_t)featuresCur.cols());
}

void DepthEstimator::getDepthCalcStats(const Eigen::VectorXi& resultType, long long counters[21]) {
    int64_t hist[21];
    int rc = mld_status_histogram_host(_handle, reinterpret_cast<const int32_t*>(resultType.data()), resultType.size(), hist);
    if (rc != MLD_OK) rethrow(rc);
    for (int i = 0; i < 21; i++) counters[i] = (long long)hist[i];
}

void DepthEstimator::getCloudCameraCs(Cloud::Ptr& pointCloud_cam_cs) {
    std::vector<double> cam((size_t)std::max<long long>(_pointCount, 1) * 3);
    int rc = mld_get_points_camera(_handle, cam.data());
    if (rc != MLD_OK) rethrow(rc);
    pointCloud_cam_cs->clear();
    for (long long i = 0; i < _pointCount; i++) {
        pcl::PointXYZI point;
        point.x = (float)cam[(size_t)i * 3];
        point.y = (float)cam[(size_t)i * 3 + 1];
        point.z = (float)cam[(size_t)i * 3 + 2];
        point.intensity = 1;
        pointCloud_cam_cs->points.push_back(point);
    }
    pointCloud_cam_cs->width = (uint32_t)pointCloud_cam_cs->points.size();
    pointCloud_cam_cs->height = 1;
    pointCloud_cam_cs->is_dense = false;
}

void DepthEstimator::getPointsCloudImageCs(Eigen::Matrix2Xd& visiblePointsImageCs) {
    // _points_cs_image_visible: compacted in cloud order on the device (mld_get_visible_points)
    int64_t nvis = 0;
    int rc = mld_get_visible_points(_handle, nullptr, nullptr, nullptr, 0, &nvis);
    if (rc != MLD_OK) rethrow(rc);
    visiblePointsImageCs.resize(2, (int)nvis);
    if (nvis == 0) return;
    rc = mld_get_visible_points(_handle, nullptr, visiblePointsImageCs.data(), nullptr, nvis, &nvis);
    if (rc != MLD_OK) rethrow(rc);
}

double DepthEstimator::getPointDepthCamVisible(int index) {
    if (!_depthCamVisibleValid) {
        int64_t nvis = 0;
        int rc = mld_get_visible_points(_handle, nullptr, nullptr, nullptr, 0, &nvis);
        if (rc != MLD_OK) rethrow(rc);
        _depthCamVisible.assign((size_t)std::max<int64_t>(nvis, 1), 0.0);
        if (nvis > 0) {
            rc = mld_get_visible_points(_handle, nullptr, nullptr, _depthCamVisible.data(), nvis, &nvis);
            if (rc != MLD_OK) rethrow(rc);
        }
        _depthCamVisible.resize((size_t)nvis);
        _depthCamVisibleValid = true;
    }
    return _depthCamVisible.at((size_t)index);
}

namespace {
// DepthEstimator::FillCloud (DepthEstimator.cpp:354-374): xyz triples -> PointXYZI with intensity 1
void fill_cloud(const double* xyz, size_t n, const std::vector<unsigned char>* keep, DepthEstimator::Cloud::Ptr& cloud) {
    cloud->clear();
    for (size_t i = 0; i < n; i++) {
        if (keep && !(*keep)[i]) continue;
        pcl::PointXYZI point;
        point.x = (float)xyz[i * 3];
        point.y = (float)xyz[i * 3 + 1];
        point.z = (float)xyz[i * 3 + 2];
        point.intensity = 1;
        cloud->points.push_back(point);
    }
    cloud->width = (uint32_t)cloud->points.size();
    cloud->height = 1;
    cloud->is_dense = false;
}
}  // namespace

void DepthEstimator::getCloudRansacPlane(Cloud::Ptr& pointCloud_plane_ransac) {
    std::vector<int32_t> idx(_groundInliers.begin(), _groundInliers.end());
    std::vector<double> cam(std::max<size_t>(idx.size(), 1) * 3);
    if (!idx.empty()) {
        int rc = mld_get_points_camera_indexed(_handle, idx.data(), (int64_t)idx.size(), cam.data());
        if (rc != MLD_OK) rethrow(rc);
    }
    std::vector<unsigned char> keep(idx.size(), 1);
    if (_parameters->ransac_plane_use_camx_treshold) {  // DepthEstimator.cpp:300-305
        const double treshold = _parameters->ransac_plane_treshold_camx;
        for (size_t i = 0; i < idx.size(); i++) keep[i] = std::fabs(cam[i * 3]) <= treshold ? 1 : 0;
    }
    fill_cloud(cam.data(), idx.size(), &keep, pointCloud_plane_ransac);
}

void DepthEstimator::getCloudTriangleCorners(Cloud::Ptr& pointCloud_triangle_corner) {
    const int F = (int)(_lastFeatures.size() / 2);
    std::vector<double> corners((size_t)std::max(F, 1) * 9);
    std::vector<uint8_t> valid((size_t)std::max(F, 1), 0);
    if (F > 0 && _isInitializedPointCloud) {
        int rc = mld_get_triangle_corners(_handle, _lastFeatures.data(), F, corners.data(), valid.data());
        if (rc != MLD_OK) rethrow(rc);
    }
    std::vector<unsigned char> keep((size_t)F * 3, 0);
    for (int i = 0; i < F; i++) keep[(size_t)i * 3] = keep[(size_t)i * 3 + 1] = keep[(size_t)i * 3 + 2] = valid[(size_t)i];
    fill_cloud(corners.data(), (size_t)F * 3, &keep, pointCloud_triangle_corner);
}

void DepthEstimator::getCloudInterpolated(Cloud::Ptr& pointCloud_interpolated) { fill_cloud(nullptr, 0, nullptr, pointCloud_interpolated); }
void DepthEstimator::getCloudInterpolatedPlane(Cloud::Ptr& pointCloud_interpolated_plane) {
    fill_cloud(nullptr, 0, nullptr, pointCloud_interpolated_plane);
}
void DepthEstimator::getCloudNeighbors(Cloud::Ptr& pointCloud_neighbors) { fill_cloud(nullptr, 0, nullptr, pointCloud_neighbors); }

}  // namespace Mono_Lidar
