#pragma once
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct Image {
    using ConstPtr = std::shared_ptr<const Image>;
    std_msgs::Header header;
    uint32_t height = 0, width = 0, step = 0;
    std::string encoding;
    std::vector<uint8_t> data;
};
namespace image_encodings {
const std::string MONO8 = "mono8";
}
}  // namespace sensor_msgs
