// matches_msg_depth_ros/msg/{FeaturePoint,Tracklet,MatchesMsg}.msg as the generated C++ message structs
#pragma once
#include <std_msgs/Header.h>
namespace matches_msg_depth_ros {
struct FeaturePoint {
    float u = 0, v = 0, d = 0;
};
struct Tracklet {
    uint64_t id = 0;
    uint32_t age = 0;
    std::vector<FeaturePoint> feature_points;
};
struct MatchesMsg {
    std_msgs::Header header;
    std::vector<ros::Time> stamps;
    std::vector<Tracklet> tracks;
};
}  // namespace matches_msg_depth_ros
