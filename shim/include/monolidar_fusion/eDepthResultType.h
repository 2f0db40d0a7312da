// Mono_Lidar::DepthResultType for callers of the shim: generated from the C ABI's status list
// (MLD_DEPTH_RESULT_TYPES in include/mld_c_api.h), which is value-compatible with the reference's enum
// (monolidar_fusion/include/monolidar_fusion/eDepthResultType.h:9-31).
#pragma once
#include "mld_c_api.h"

namespace Mono_Lidar {

#define MLD_SHIM_RESULT_TYPE(name, value) name = value,
enum DepthResultType { MLD_DEPTH_RESULT_TYPES(MLD_SHIM_RESULT_TYPE) };
#undef MLD_SHIM_RESULT_TYPE

}  // namespace Mono_Lidar
