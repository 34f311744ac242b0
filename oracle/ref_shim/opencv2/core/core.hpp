// TEST INFRASTRUCTURE.  Stand-in for the five OpenCV templates the reference's public
// headers name (/root/reference/tfusion/include/tfusion/types.hpp:15-18).  OpenCV C++ is
// not installed in this image; only the storage layout is needed to compile the
// reference's host+device shared headers for oracle/_ref.
#pragma once
namespace cv {
template <typename T, int m, int n> struct Matx {
    T val[m * n];
    T& operator()(int r, int c) { return val[r * n + c]; }
    const T& operator()(int r, int c) const { return val[r * n + c]; }
};
typedef Matx<float, 3, 3> Matx33f;
typedef Matx<float, 4, 4> Matx44f;
template <typename T, int n> struct Vec { T val[n]; T& operator[](int i) { return val[i]; } const T& operator[](int i) const { return val[i]; } };
typedef Vec<float, 3> Vec3f;
typedef Vec<int, 3> Vec3i;
template <typename T> struct Ptr { T* p; Ptr() : p(0) {} Ptr(T* q) : p(q) {} T* operator->() const { return p; } T& operator*() const { return *p; } };
}
