// Stand-in (test infrastructure for oracle/_ref): see ceres/ceres.h.
#pragma once
#include <ceres/ceres.h>
