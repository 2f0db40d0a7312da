// Minimal stand-in for <ros/ros.h> as tracklets_depth touches it (ros::Time and the logging macros). ROS is absent from this
// image; these stubs exist only so that the reference's OWN caller sources (tracklets_depth) can be compiled, unmodified,
// against the B200 shim (tests/test_caller_dropin.py).
#pragma once
#include <cassert>
#include <chrono>
#include <cstdint>
#include <deque>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace ros {
struct Time {
    uint32_t sec = 0, nsec = 0;
    Time() = default;
    Time(uint32_t s, uint32_t ns) : sec(s), nsec(ns) {}
    double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
    uint64_t toNSec() const { return (uint64_t)sec * 1000000000ull + nsec; }
};
}  // namespace ros
#define ROS_DEBUG_STREAM(x) do { std::stringstream mld_ss; mld_ss << x; } while (0)
#define ROS_INFO_STREAM(x) do { std::cout << "[INFO] " << x << std::endl; } while (0)
#define ROS_WARN_STREAM(x) do { std::cout << "[WARN] " << x << std::endl; } while (0)
#define ROS_ERROR_STREAM(x) do { std::cout << "[ERROR] " << x << std::endl; } while (0)
