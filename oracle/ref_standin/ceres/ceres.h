// Stand-in for Ceres (absent here). TEST INFRASTRUCTURE for oracle/_ref only. The reference's Ceres user,
// PlaneEstimationLeastSquares.cpp, is NOT compiled into oracle/_ref (its behaviour is undefined: all-zero start
// with a 0/0 residual and an out-of-bounds read of ext[4]); these declarations only let
// ErrorPlane.h parse.
#pragma once
namespace ceres {
class CostFunction { public: virtual ~CostFunction() {} };
template <class F, int R, int N0>
class AutoDiffCostFunction : public CostFunction { public: explicit AutoDiffCostFunction(F* f) : f_(f) {} ~AutoDiffCostFunction() override { delete f_; } private: F* f_; };
}  // namespace ceres
