// Minimal stand-in for pcl/point_types.h: pcl::PointXYZI with PCL's 32-byte layout
// (x,y,z,pad | intensity,pad,pad,pad), see SURVEY.md 8a row A3.
#pragma once
namespace pcl {
struct alignas(16) PointXYZI {
    float x = 0, y = 0, z = 0, _pad0 = 1.f;
    float intensity = 0, _pad1 = 0, _pad2 = 0, _pad3 = 0;
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI is 32 bytes");
}  // namespace pcl
