// Per-feature debug record of the reference (monolidar_fusion/include/monolidar_fusion/DepthCalcStatsSinglePoint.h:20-67). The
// kernels do not export per-feature intermediates (do_debug_singleFeatures is a CPU-side debugging aid upstream); the record
// type exists so that code touching DepthCalculationStatistics::getPointStats() / AddPoint() keeps compiling.
#pragma once
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include <Eigen/Eigen>

#include "eDepthResultType.h"

namespace Mono_Lidar {
struct DepthCalcStatsSinglePoint {
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    DepthResultType _calcResult{Unspecified};
    std::vector<std::pair<float, float>> _neighbors2d, _pointsSegmented2d;
    std::vector<std::tuple<float, float, float>> _neighbors3d, _pointsSegmented3d;
    std::pair<int, int> _searchRectTopLeft{0, 0}, _searchRectBottomRight{0, 0};
    int _histBinCount{0}, _histMinDist{0}, _histMaxDist{0}, _histFound{0};
    double _histBinWitdh{0}, _histLowerBorder{0}, _histHigherBorder{0};
    std::vector<float> _histDepthEntryCount;
    std::string _pcaResult;
    int _featureX{0}, _featureY{0};
    double _featureDepth{-1};
    Eigen::Vector2d _pointInterpolated2d;
    Eigen::Vector3d _pointInterpolated3d;
};
}  // namespace Mono_Lidar
