// Mono_Lidar::GroundPlane / RansacPlane / SemanticPlane with the reference's interface
// (monolidar_fusion/include/monolidar_fusion/RansacPlane.h:38-216). RansacPlane::CalculateInliersPlane runs the fit on
// the GPU through mld_estimate_ground_plane, SemanticPlane::CalculateInliersPlane through mld_semantic_ground_plane; any
// other caller-provided plane plugs in through the same base class.
#pragma once
#include <exception>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include <Eigen/Eigen>
#include <opencv2/core/core.hpp>
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

// RansacPlane.h:166-216. The label image is copied at construction (the reference keeps a cv::Mat header copy).
class SemanticPlane : public GroundPlane {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW

    struct Camera {
        double f;
        double cu;
        double cv;
        Eigen::Affine3d transform_cam_lidar;
    };

    explicit SemanticPlane(const cv::Mat& img, Camera cam, std::set<int> groundplane_label, double inlier_threshold);
    ~SemanticPlane() override;

    void CalculateInliersPlane(const Cloud::ConstPtr& pointCloud, double /*min_z*/, double /*max_z*/) override {
        CalculateInliersPlane(pointCloud);  // the reference's override ignores the z range too (RansacPlane.h:205-207)
    }
    void CalculateInliersPlane(const Cloud::ConstPtr& pointCloud) override;

private:
    std::vector<unsigned char> labels_;  // rows x cols, row-major
    int rows_{0}, cols_{0};
    Camera cam_;
    double inlier_threshold_{0.1};
    std::set<int> groundplane_label_{6, 7, 8, 9};
    mld_handle* handle_{nullptr};
};

}  // namespace Mono_Lidar
