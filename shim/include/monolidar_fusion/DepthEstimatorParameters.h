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
    }
    void print();
};

}  // namespace Mono_Lidar
