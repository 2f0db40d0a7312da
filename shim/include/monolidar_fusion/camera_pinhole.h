// CameraPinhole with the constructor and accessors the reference exposes
// (monolidar_fusion/include/monolidar_fusion/camera_pinhole.h:21-47). The projection / viewing-ray
// arithmetic of the reference class (:52-106) runs on the GPU behind the C ABI, not here.
#pragma once
#include <memory>

class CameraPinhole final {
public:
    using Ptr = std::shared_ptr<CameraPinhole>;
    using ConstPtr = std::shared_ptr<const CameraPinhole>;

    explicit CameraPinhole(int width, int height, double focal_length, double principal_point_x, double principal_point_y)
            : width_(width), height_(height), focal_length_(focal_length), principal_point_x_(principal_point_x),
              principal_point_y_(principal_point_y) {}

    void getImageSize(int& width, int& height) const {
        width = width_;
        height = height_;
    }
    double focalLength() const { return focal_length_; }
    double principalPointX() const { return principal_point_x_; }
    double principalPointY() const { return principal_point_y_; }

private:
    int width_, height_;
    double focal_length_, principal_point_x_, principal_point_y_;
};
