#pragma once
#include <opencv2/core/core.hpp>
#include <sensor_msgs/Image.h>
namespace cv_bridge {
struct CvImage {
    std_msgs::Header header;
    std::string encoding;
    cv::Mat image;
};
using CvImageConstPtr = std::shared_ptr<const CvImage>;
inline CvImageConstPtr toCvShare(const sensor_msgs::Image::ConstPtr& img, const std::string& encoding) {
    auto out = std::make_shared<CvImage>();
    out->header = img->header;
    out->encoding = encoding;
    out->image = cv::Mat((int)img->height, (int)img->width, CV_8UC1, const_cast<uint8_t*>(img->data.data()));
    return out;
}
}  // namespace cv_bridge
