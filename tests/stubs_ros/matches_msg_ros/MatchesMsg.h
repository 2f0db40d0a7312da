// matches_msg_ros (external package, not in the reference repo): tracklets of 2-D feature points as the feature matcher sends them
#pragma once
#include <std_msgs/Header.h>
namespace matches_msg_ros {
struct FeaturePoint {
    float u = 0, v = 0;
};
struct Tracklet {
    uint64_t id = 0;
    uint32_t age = 0;
    std::vector<FeaturePoint> feature_points;
};
struct MatchesMsg {
    using ConstPtr = std::shared_ptr<const MatchesMsg>;
    std_msgs::Header header;
    std::vector<ros::Time> stamps;
    std::vector<Tracklet> tracks;
};
using MatchesMsgConstPtr = std::shared_ptr<const MatchesMsg>;
}  // namespace matches_msg_ros
