// Minimal stand-in for pcl/point_cloud.h.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
template <typename PointT>
class PointCloud {
public:
    using Ptr = std::shared_ptr<PointCloud<PointT>>;
    using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    void clear() { points.clear(); }
    size_t size() const { return points.size(); }
};
}  // namespace pcl
