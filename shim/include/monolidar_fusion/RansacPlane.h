// Mono_Lidar::GroundPlane / RansacPlane with the reference's interface
// (monolidar_fusion/include/monolidar_fusion/RansacPlane.h:38-164). RansacPlane::CalculateInliersPlane
// runs the fit on the GPU through mld_estimate_ground_plane; a caller-provided plane (e.g. the
// reference's SemanticPlane, computed on the host) plugs in through the same base class.
#pragma once
#include <exception>
#include <map>
#include <memory>
#include <vector>

#include <Eigen/Eigen>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "DepthEstimatorParameters.h"

namespace Mono_Lidar {

class DepthEstimator;

class GroundPlane {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    using Ptr = std::shared_ptr<GroundPlane>;
    using Point = pcl::PointXYZI;
    using Cloud = pcl::PointCloud<Point>;

    struct ExceptionPclInvalid : public std::exception {
        virtual const char* what() const throw() { return "In GroundPlane: Input pointcloud is invalid"; }
    };

    explicit GroundPlane() = default;
    virtual ~GroundPlane() = default;

    virtual void CalculateInliersPlane(const Cloud::ConstPtr& pointCloud) { CalculateInliersPlane(pointCloud, -1000., 1000.); }
    virtual void CalculateInliersPlane(const Cloud::ConstPtr& pointCloud, double min_z, double max_z) = 0;

    bool isSegmented() const { return is_segmented_; }
    inline Eigen::Vector4f& getModelCoeffs() { return _modelCoeffs; }
    bool CheckPointInPlane(const int index) const { return _pointIsInPlane.count(index) != 0; }
    const std::vector<int>& getInlinersIndex() { return _inliersIndex; }

protected:
    friend class DepthEstimator;
    bool is_segmented_{false};
    Eigen::Vector4f _modelCoeffs;
    std::map<int, bool> _pointIsInPlane;
    std::vector<int> _inliersIndex;
};

class RansacPlane : public GroundPlane {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    using Ptr = std::shared_ptr<RansacPlane>;

    RansacPlane();
    explicit RansacPlane(const std::shared_ptr<DepthEstimatorParameters>& parameters);
    ~RansacPlane() override;

    void CalculateInliersPlane(const Cloud::ConstPtr& pointCloud) override;
    void CalculateInliersPlane(const Cloud::ConstPtr& pointCloud, double min_z, double max_z) override;

    void setSeed(unsigned long long seed) { seed_ = seed; }  // PCL seeds from the clock; the GPU fit is counter based

private:
    DepthEstimatorParameters params_;
    mld_handle* handle_{nullptr};
    unsigned long long seed_{0};
};

}  // namespace Mono_Lidar
