// The tfusion public API is OpenCV-typed (reference: include/tfusion/types.hpp:15-18 — cv::Matx33f, cv::Vec3f,
// cv::Vec3i, cv::Affine3f, cv::Ptr).  When OpenCV headers are available they are used as is; otherwise this header
// supplies the five templates with OpenCV's storage layout (row-major val[]) and the handful of members the
// library and apps/demo.cpp touch.  Define TFUSION_FORCE_CV_COMPAT to use the shim even when OpenCV exists.
#pragma once
#if !defined(TFUSION_FORCE_CV_COMPAT) && defined(__has_include)
#if __has_include(<opencv2/core/affine.hpp>)
#define TFUSION_HAVE_OPENCV 1
#endif
#endif

#ifdef TFUSION_HAVE_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/core/affine.hpp>
#else
#include <cmath>
#include <memory>

namespace cv {

template <typename T, int m, int n> struct Matx {
    T val[m * n];
    Matx() { for (int i = 0; i < m * n; ++i) val[i] = T(0); }
    static Matx eye() { Matx r; for (int i = 0; i < (m < n ? m : n); ++i) r(i, i) = T(1); return r; }
    T& operator()(int r, int c) { return val[r * n + c]; }
    const T& operator()(int r, int c) const { return val[r * n + c]; }
};
template <typename T, int m, int k, int n> inline Matx<T, m, n> operator*(const Matx<T, m, k>& a, const Matx<T, k, n>& b) {
    Matx<T, m, n> c;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) { T s = 0; for (int l = 0; l < k; ++l) s += a(i, l) * b(l, j); c(i, j) = s; }
    return c;
}
typedef Matx<float, 3, 3> Matx33f;
typedef Matx<float, 4, 4> Matx44f;
typedef Matx<float, 6, 6> Matx66f;

template <typename T, int n> struct Vec {
    T val[n];
    Vec() { for (int i = 0; i < n; ++i) val[i] = T(0); }
    Vec(T a, T b, T c) { static_assert(n == 3, "3 components"); val[0] = a; val[1] = b; val[2] = c; }
    explicit Vec(const T* p) { for (int i = 0; i < n; ++i) val[i] = p[i]; }
    static Vec all(T v) { Vec r; for (int i = 0; i < n; ++i) r.val[i] = v; return r; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    template <typename T2> operator Vec<T2, n>() const { Vec<T2, n> r; for (int i = 0; i < n; ++i) r.val[i] = (T2)val[i]; return r; }   // OpenCV: Matx::operator Matx<T2,m,n>()
};
template <typename T, int n> inline Vec<T, n> operator/(const Vec<T, n>& a, T s) { Vec<T, n> r; for (int i = 0; i < n; ++i) r[i] = a[i] / s; return r; }
typedef Vec<float, 3> Vec3f;
typedef Vec<int, 3> Vec3i;
typedef Vec<float, 6> Vec6f;

template <typename T> struct Affine3 {
    typedef Matx<T, 4, 4> Mat4;
    typedef Matx<T, 3, 3> Mat3;
    typedef Vec<T, 3> Vec3;
    Mat4 matrix;
    Affine3() : matrix(Mat4::eye()) {}
    explicit Affine3(const Mat4& m) : matrix(m) {}
    // Rodrigues vector + translation, evaluated in double like OpenCV's affine.hpp
    Affine3(const Vec3& rvec, const Vec3& t) : matrix(Mat4::eye()) {
        double th = std::sqrt((double)rvec[0] * rvec[0] + (double)rvec[1] * rvec[1] + (double)rvec[2] * rvec[2]);
        if (th >= 2.220446049250313e-16) {
            double c = std::cos(th), s = std::sin(th), c1 = 1. - c, it = 1. / th;
            T r[3] = {(T)(rvec[0] * it), (T)(rvec[1] * it), (T)(rvec[2] * it)};
            double rx[9] = {0, -(double)r[2], (double)r[1], (double)r[2], 0, -(double)r[0], -(double)r[1], (double)r[0], 0};
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    matrix(i, j) = (T)(c * (i == j) + c1 * (double)r[i] * r[j] + s * rx[i * 3 + j]);
        }
        for (int i = 0; i < 3; ++i) matrix(i, 3) = t[i];
    }
    static Affine3 Identity() { return Affine3(); }
    template <typename Y> operator Affine3<Y>() const { Affine3<Y> r; for (int i = 0; i < 16; ++i) r.matrix.val[i] = (Y)matrix.val[i]; return r; }   // affine.hpp: cast to another element type
    Mat3 rotation() const { Mat3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = matrix(i, j); return r; }
    Vec3 translation() const { return Vec3(matrix(0, 3), matrix(1, 3), matrix(2, 3)); }
    Affine3 translate(const Vec3& t) const { Affine3 r(*this); for (int i = 0; i < 3; ++i) r.matrix(i, 3) += t[i]; return r; }
    // general inverse in double (OpenCV: matrix.inv(DECOMP_SVD))
    Affine3 inv() const {
        double m[4][8];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { m[i][j] = matrix(i, j); m[i][4 + j] = (i == j); }
        for (int c = 0; c < 4; ++c) {
            int p = c;
            for (int r = c + 1; r < 4; ++r) if (std::fabs(m[r][c]) > std::fabs(m[p][c])) p = r;
            if (p != c) for (int j = 0; j < 8; ++j) { double t = m[c][j]; m[c][j] = m[p][j]; m[p][j] = t; }
            double d = m[c][c];
            for (int j = 0; j < 8; ++j) m[c][j] /= d;
            for (int r = 0; r < 4; ++r) if (r != c) { double f = m[r][c]; for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j]; }
        }
        Affine3 o;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) o.matrix(i, j) = (T)m[i][4 + j];
        return o;
    }
};
template <typename T> inline Affine3<T> operator*(const Affine3<T>& a, const Affine3<T>& b) { return Affine3<T>(a.matrix * b.matrix); }
typedef Affine3<float> Affine3f;

template <typename T> struct Ptr : public std::shared_ptr<T> {
    Ptr() {}
    Ptr(T* p) : std::shared_ptr<T>(p) {}
};

}  // namespace cv
#endif
