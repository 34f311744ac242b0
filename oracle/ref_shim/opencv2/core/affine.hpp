#pragma once
#include <opencv2/core/core.hpp>
namespace cv {
template <typename T> struct Affine3 { Matx<T, 4, 4> matrix; };
typedef Affine3<float> Affine3f;
}
