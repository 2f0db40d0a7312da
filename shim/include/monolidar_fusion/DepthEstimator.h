// Mono_Lidar::DepthEstimator -- source-compatible host shim over the C ABI (include/mld_c_api.h).
//
// Public interface = the reference's (monolidar_fusion/include/monolidar_fusion/DepthEstimator.h:73-222):
// InitConfig (file or struct), Initialize, setInputCloud, the five CalculateDepth overloads and the
// getters tracklets_depth uses (tracklets_depth/src/tracklet_depth_module.cpp:80,115,401,413). The shim
// holds no arithmetic: every method moves buffers and calls libmld_cuda.so. Errors are rethrown with
// the reference's exception types (const char*, std::string, std::runtime_error,
// GroundPlane::ExceptionPclInvalid).
#pragma once
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include <Eigen/Eigen>
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "DepthCalculationStatistics.h"
#include "DepthEstimatorParameters.h"
#include "RansacPlane.h"
#include "camera_pinhole.h"
#include "eDepthResultType.h"

namespace Mono_Lidar {

class DepthEstimator {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW

    using Point = pcl::PointXYZI;
    using Cloud = pcl::PointCloud<Point>;
    using UniquePtr = std::unique_ptr<DepthEstimator>;
    using SharedPtr = std::shared_ptr<DepthEstimator>;

    std::map<DepthResultType, std::string> DepthResultTypeMap;

    DepthEstimator();
    ~DepthEstimator();
    DepthEstimator(const DepthEstimator&) = delete;
    DepthEstimator& operator=(const DepthEstimator&) = delete;

    bool Initialize(const std::shared_ptr<CameraPinhole>& camera, const Eigen::Affine3d& transform_lidar_to_cam);
    bool InitConfig(const std::string& filePath, const bool printparams = true);
    bool InitConfig(std::shared_ptr<DepthEstimatorParameters> parameters = nullptr, const bool printparams = false);

    void setInputCloud(const Cloud::ConstPtr& pointCloud, GroundPlane::Ptr& ransacPlane);

    std::shared_ptr<DepthEstimatorParameters> getParameters() { return _parameters; }
    std::shared_ptr<CameraPinhole> getCamera() { return _camera; }
    Eigen::Affine3d getTransformLidarToCam() { return _transform_lidar_to_cam; }

    void CalculateDepth(const Cloud::ConstPtr& pointCloud, const Eigen::Matrix2Xd& points_image_cs, Eigen::VectorXd& points_depths,
                        GroundPlane::Ptr& ransacPlane);
    void CalculateDepth(const Cloud::ConstPtr& pointCloud, const Eigen::Matrix2Xd& points_image_cs, Eigen::VectorXd& points_depths,
                        Eigen::VectorXi& resultType, GroundPlane::Ptr& ransacPlane);
    void CalculateDepth(const Eigen::Matrix2Xd& points_image_cs, Eigen::VectorXd& points_depths, const GroundPlane::Ptr& ransacPlane);
    void CalculateDepth(const Eigen::Matrix2Xd& points_image_cs, Eigen::VectorXd& points_depths, Eigen::VectorXi& resultType,
                        const GroundPlane::Ptr& ransacPlane);
    std::pair<DepthResultType, double> CalculateDepth(const Eigen::Vector2d& point_image_cs, const GroundPlane::Ptr& ransacPlane);

    // tracklets_depth batch adaptor: previous + current cloud of one frame in a single call
    // (TrackletDepthModule::process issues them back to back, tracklet_depth_module.cpp:318, :330). A null
    // pointCloudLast (first frame) sets depthsLast to -1 like CalculateFeatureDepthsLastFrame (:97-100). When pointCloudLast is the
    // very cloud the previous call (or setInputCloud) put on the device, it is not uploaded or projected again.
    void CalculateDepthPair(const Cloud::ConstPtr& pointCloudLast, const Eigen::Matrix2Xd& featuresLast, Eigen::VectorXd& depthsLast,
                            GroundPlane::Ptr& planeLast, const Cloud::ConstPtr& pointCloudCur, const Eigen::Matrix2Xd& featuresCur,
                            Eigen::VectorXd& depthsCur, GroundPlane::Ptr& planeCur);

    // ---- statistics and debug views (DepthEstimator.h:116-164), served from device buffers on demand ----
    // counters of the last CalculateDepth call (kept when the parameters' do_depth_calc_statistics is set, as upstream)
    const DepthCalculationStatistics& getDepthCalcStats() { return _depthCalcStats; }
    // the same counters for any result vector (not in the reference)
    void getDepthCalcStats(const Eigen::VectorXi& resultType, long long counters[21]);
    // camera-frame depth of visible point `index` (_points_cs_camera(2, _pointIndex[index]))
    double getPointDepthCamVisible(int index);
    // _points_cs_image_visible: image coordinates of the visible points in cloud order (device stream compaction)
    void getPointsCloudImageCs(Eigen::Matrix2Xd& visiblePointsImageCs);
    // the whole cloud in the camera frame (intensity 1, like upstream)
    void getCloudCameraCs(Cloud::Ptr& pointCloud_cam_cs);
    // the ground plane's inliers in the camera frame, optionally cut at |x| <= ransac_plane_treshold_camx (:294-308)
    void getCloudRansacPlane(Cloud::Ptr& pointCloud_plane_ransac);
    // the triangle corners CalculateDepthSegmented used for the features of the last CalculateDepth call (three per feature
    // that reached a corner selection, in feature order; upstream the order is whatever the OpenMP threads produced)
    void getCloudTriangleCorners(Cloud::Ptr& pointCloud_triangle_corner);
    // Upstream never fills these lists any more (the push_backs are commented out, DepthEstimator.cpp:663, :1032): empty clouds
    void getCloudInterpolated(Cloud::Ptr& pointCloud_interpolated);
    void getCloudInterpolatedPlane(Cloud::Ptr& pointCloud_interpolated_plane);
    void getCloudNeighbors(Cloud::Ptr& pointCloud_neighbors);

    void setRansacSeed(unsigned long long seed) { _ransacSeed = seed; }

private:
    [[noreturn]] void rethrow(int rc);

    std::shared_ptr<DepthEstimatorParameters> _parameters;
    std::shared_ptr<CameraPinhole> _camera;
    Eigen::Affine3d _transform_lidar_to_cam;
    mld_handle* _handle{nullptr};
    mld_params _paramsOnDevice{};            // the parameter block the device handle was created from (re-checked in Initialize)
    bool _isInitializedConfig{false};
    bool _isInitialized{false};
    bool _isInitializedPointCloud{false};
    long long _pointCount{0};
    unsigned long long _ransacSeed{0};
    DepthCalculationStatistics _depthCalcStats;
    std::vector<int> _groundInliers;         // inlier indices of the plane seen by the last setInputCloud (getCloudRansacPlane)
    std::vector<double> _lastFeatures;       // 2 x F features of the last CalculateDepth call (getCloudTriangleCorners)
    std::vector<double> _depthCamVisible;    // lazily fetched per cloud (getPointDepthCamVisible)
    bool _depthCamVisibleValid{false};
    Cloud::ConstPtr _residentCloud;          // the cloud whose projection is on the device (CalculateDepthPair reuses it as 'previous')
};

}  // namespace Mono_Lidar
