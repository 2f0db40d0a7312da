// Mono_Lidar::DepthEstimatorParameters with the reference's field names and defaults
// (monolidar_fusion/include/monolidar_fusion/DepthEstimatorParameters.h:12-172). The object IS the
// C ABI's mld_params (same names), so no translation layer exists between the two.
#pragma once
#include <string>

#include "mld_c_api.h"

namespace Mono_Lidar {

class DepthEstimatorParameters : public mld_params {
public:
    DepthEstimatorParameters() { mld_default_params(this); }
    // DepthEstimatorParameters::fromFile (src/DepthEstimatorParameters.cpp:16-114); throws std::string like the reference
    void fromFile(const std::string& filePath) {
        if (mld_params_from_yaml(filePath.c_str(), this) != MLD_OK) throw("Cant find settings file: " + filePath);
        int32_t v = 0;  // the debug switches are not part of the kernels' parameter block (:109-113 upstream)
        if (mld_yaml_int(filePath.c_str(), "do_debug_singleFeatures", &v, nullptr) == MLD_OK) do_debug_singleFeatures = v != 0;
        if (mld_yaml_int(filePath.c_str(), "do_publish_points", &v, nullptr) == MLD_OK) do_publish_points = v != 0;
        if (mld_yaml_int(filePath.c_str(), "do_depth_calc_statistics", &v, nullptr) == MLD_OK) do_depth_calc_statistics = v != 0;
    }
    void print();

    // Debug switches (DepthEstimatorParameters.h:166-168). do_depth_calc_statistics: keep DepthCalculationStatistics up to date.
    // do_publish_points / do_debug_singleFeatures: upstream they gate CPU-side debug vectors that its own code no longer fills.
    bool do_debug_singleFeatures{false};
    bool do_publish_points{true};
    bool do_depth_calc_statistics{true};
};

}  // namespace Mono_Lidar
