#pragma once
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct CameraInfo {
    using ConstPtr = std::shared_ptr<const CameraInfo>;
    std_msgs::Header header;
    uint32_t height = 0, width = 0;
    double K[9] = {0};  // row-major intrinsics
};
}  // namespace sensor_msgs
