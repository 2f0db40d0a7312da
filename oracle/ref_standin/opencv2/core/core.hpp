// Stand-in (test infrastructure for oracle/_ref): see opencv2/opencv.hpp.
#pragma once
#include <opencv2/opencv.hpp>
