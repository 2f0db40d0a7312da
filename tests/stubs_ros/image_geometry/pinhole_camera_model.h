#pragma once
#include <sensor_msgs/CameraInfo.h>
namespace image_geometry {
class PinholeCameraModel {
public:
    bool fromCameraInfo(const sensor_msgs::CameraInfo::ConstPtr& info) {
        for (int i = 0; i < 9; i++) K_[i] = info->K[i];
        return true;
    }
    double fx() const { return K_[0]; }
    double fy() const { return K_[4]; }
    double cx() const { return K_[2]; }
    double cy() const { return K_[5]; }

private:
    double K_[9] = {0};
};
}  // namespace image_geometry
