// REFERENCE-ARM INFRASTRUCTURE.  The library itself only needs cv::viz::isNan
// (projective_icp.cpp:199).
#pragma once
#include <cmath>
namespace cv { namespace viz {
inline bool isNan(double x) { return std::isnan(x); }
inline bool isNan(float x) { return std::isnan(x); }
} }
