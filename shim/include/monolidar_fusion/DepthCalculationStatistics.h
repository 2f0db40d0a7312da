// Mono_Lidar::DepthCalculationStatistics with the reference's interface
// (monolidar_fusion/include/monolidar_fusion/DepthCalculationStatistics.h:17-305): per-frame and accumulated counters per
// DepthResultType. Here the counters are ONE table indexed by the status value, filled from the status histogram the GPU
// reduces (mld_last_status_histogram); the reference's named Add*/get* members are thin views of that table.
//
// Two upstream quirks are kept because callers print these numbers side by side with the reference's:
//   getTresholdDepthLocalGreaterMax() / ...SmallerMin() return the GLOBAL counters (DepthCalculationStatistics.h:134-139);
//   AddRegionGrowingSeedsOutOfRange() only touches the accumulated counter (:104-107), which is also what
//   getRegionGrowingSeedsOutOfRange() returns (:164-166).
// One deliberate difference: upstream the per-feature LogDepthCalcStats call is commented out (DepthEstimator.cpp:470-479), so
// its counters other than the point count stay 0; here they are filled, since the histogram costs one tiny kernel.
#pragma once
#include <iostream>
#include <memory>
#include <vector>

#include <Eigen/Eigen>

#include "DepthCalcStatsSinglePoint.h"
#include "eDepthResultType.h"

namespace Mono_Lidar {

class DepthCalculationStatistics {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    static constexpr int kTypes = 21;  // DepthResultType values 0..20

    DepthCalculationStatistics() { Clear(); _acc_frames = 0; }

    void SetPointCount(int value) { _pointCount = value; _acc_pointCount += value; }
    // one feature of status `t` (what LogDepthCalcStats / AddPoint do upstream)
    void Add(DepthResultType t) {
        if ((int)t < 0 || (int)t >= kTypes) return;
        if (t != RegionGrowingSeedsOutOfRange) _cur[t]++;  // upstream quirk, see the header comment
        _acc[t]++;
    }
    // all counters of one frame at once: hist[t] = number of features with status t
    void SetFromHistogram(const long long* hist, int pointCount) {
        Clear();
        for (int t = 0; t < kTypes; t++) {
            if (t != RegionGrowingSeedsOutOfRange) _cur[t] = (int)hist[t];
            _acc[t] += (int)hist[t];
        }
        SetPointCount(pointCount);
    }
    void AddPoint(const std::shared_ptr<DepthCalcStatsSinglePoint>& point) { Add(point->_calcResult); _points.push_back(point); }
    void Clear() {
        _acc_frames++;
        _pointCount = 0;
        for (int t = 0; t < kTypes; t++) _cur[t] = 0;
        _points.clear();
    }

#define MLD_SHIM_STAT(AddName, getName, type) \
    void AddName() { Add(type); }             \
    int getName() { return _cur[type]; }
    MLD_SHIM_STAT(AddSuccess, getSuccess, Success)
    MLD_SHIM_STAT(AddRadiusSearchInsufficientPoints, getRadiusSearchInsufficientPoints, RadiusSearchInsufficientPoints)
    MLD_SHIM_STAT(AddHistogramNoLocalMax, getHistogramNoLocalMax, HistogramNoLocalMax)
    MLD_SHIM_STAT(AddTresholdDepthGlobalGreaterMax, getTresholdDepthGlobalGreaterMax, TresholdDepthGlobalGreaterMax)
    MLD_SHIM_STAT(AddTresholdDepthGlobalSmallerMin, getTresholdDepthGlobalSmallerMin, TresholdDepthGlobalSmallerMin)
    MLD_SHIM_STAT(AddTriangleNotPlanar, getTriangleNotPlanar, TriangleNotPlanar)
    MLD_SHIM_STAT(AddTriangleNotPlanarInsufficientPoints, getTriangleNotPlanarInsufficientPoints, TriangleNotPlanarInsufficientPoints)
    MLD_SHIM_STAT(AddCornerBehindCamera, getCornerBehindCamera, CornerBehindCamera)
    MLD_SHIM_STAT(AddPlaneViewrayNotOrthogonal, getPlaneViewrayNotOrthogonal, PlaneViewrayNotOrthogonal)
    MLD_SHIM_STAT(AddPCAIsPoint, getPCAIsPoint, PcaIsPoint)
    MLD_SHIM_STAT(AddPCAIsLine, getPCAIsLine, PcaIsLine)
    MLD_SHIM_STAT(AddPCAIsCubic, getPCAIsCubic, PcaIsCubic)
    MLD_SHIM_STAT(AddSuccessRoad, getSuccessRoad, SuccessRoad)
    MLD_SHIM_STAT(AddInsufficientRoadPoints, getInsufficientRoadPoints, InsufficientRoadPoints)
    MLD_SHIM_STAT(AddRegionGrowingInsufficientPoints, getRegionGrowingInsufficientPoints, RegionGrowingInsufficientPoints)
    MLD_SHIM_STAT(AddRegionGrowingNearestSeedNotAvailable, getRegionGrowingNearestSeedNotAvailable, RegionGrowingNearestSeedNotAvailable)
    MLD_SHIM_STAT(AddSuccessRegionGrowing, getSuccessRegionGrowing, SuccessRegionGrowing)
    MLD_SHIM_STAT(AddUnspecified, getUnspecified, Unspecified)
#undef MLD_SHIM_STAT
    void AddTresholdDepthLocalGreaterMax() { Add(TresholdDepthLocalGreaterMax); }
    void AddTresholdDepthLocalSmallerMin() { Add(TresholdDepthLocalSmallerMin); }
    int getTresholdDepthLocalGreaterMax() { return _cur[TresholdDepthGlobalGreaterMax]; }  // sic (upstream :134-136)
    int getTresholdDepthLocalSmallerMin() { return _cur[TresholdDepthGlobalSmallerMin]; }  // sic (upstream :137-139)
    void AddRegionGrowingSeedsOutOfRange() { Add(RegionGrowingSeedsOutOfRange); }
    int getRegionGrowingSeedsOutOfRange() { return _acc[RegionGrowingSeedsOutOfRange]; }   // sic (upstream :164-166)
    int getPointCount() { return _pointCount; }
    // the table itself (not in the reference): counter of any status, this frame / accumulated over all frames
    int count(DepthResultType t) const { return _cur[t]; }
    int accumulated(DepthResultType t) const { return _acc[t]; }
    int accumulatedFrames() const { return _acc_frames; }
    int accumulatedPointCount() const { return _acc_pointCount; }

    std::vector<std::shared_ptr<DepthCalcStatsSinglePoint>>& getPointStats() { return _points; }

    void ToFile(std::ostream& os) { os << *this; }
    friend std::ostream& operator<<(std::ostream& os, const DepthCalculationStatistics& d) {
        os << "--- DepthCalcStats : " << std::endl << "Current Frame: " << std::endl << "Points Count: " << d._pointCount << std::endl;
        for (int t = 0; t < kTypes; t++) os << mld_status_name(t) << ": " << d._cur[t] << std::endl;
        os << "Accumulated over " << d._acc_frames << " frames: " << std::endl << "Points Count: " << d._acc_pointCount << std::endl;
        for (int t = 0; t < kTypes; t++) os << mld_status_name(t) << ": " << d._acc[t] << std::endl;
        return os;
    }

private:
    std::vector<std::shared_ptr<DepthCalcStatsSinglePoint>> _points;
    int _pointCount{0};
    int _cur[kTypes];
    int _acc[kTypes] = {0};
    int _acc_frames{0};
    int _acc_pointCount{0};
};

}  // namespace Mono_Lidar
