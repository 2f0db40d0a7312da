// feature_tracking_core (external package, not in the reference repo): the members tracklets_depth uses
#pragma once
#include <deque>
#include <memory>
namespace feature_tracking {
struct ImagePoint {
    float u_ = 0, v_ = 0;
};
struct WorldPoint {
    double data[3] = {0, 0, 0};
};
struct Match {
    ImagePoint p1_;
    std::shared_ptr<WorldPoint> x_;
    Match() = default;
    Match(float u, float v) { p1_.u_ = u; p1_.v_ = v; }
};
struct Tracklet : public std::deque<Match> {
    uint64_t id_ = 0;
    int age_ = 0;
};
}  // namespace feature_tracking
