// REFERENCE-ARM INFRASTRUCTURE (not product code).  cv::Affine3<T> as the reference uses it
// (types.hpp:18; topfu.cpp:27,243,281; projective_icp.cpp:208-209), after OpenCV's
// core/affine.hpp: storage is a row-major Matx44; (rvec, t) constructor = Rodrigues in double;
// a * b = rotate-then-translate in T; inv() = general 4x4 inverse (OpenCV: Matx::inv(DECOMP_SVD)).
#pragma once
#include <opencv2/core/core.hpp>
namespace cv {
template <typename T> struct Affine3 {
    typedef Matx<T, 3, 3> Mat3;
    typedef Matx<T, 4, 4> Mat4;
    typedef Vec<T, 3> Vec3;
    Mat4 matrix;
    Affine3() : matrix(Mat4::eye()) {}
    Affine3(const Mat4& m) : matrix(m) {}
    static Affine3 Identity() { return Affine3(); }
    Affine3(const Vec3& rvec, const Vec3& t) : matrix(Mat4::eye()) {
        double rx = rvec[0], ry = rvec[1], rz = rvec[2];
        double theta = std::sqrt(rx * rx + ry * ry + rz * rz);
        if (theta >= DBL_EPSILON) {
            double c = std::cos(theta), s = std::sin(theta), c1 = 1. - c, it = 1. / theta;
            T r[3] = {(T)(rx * it), (T)(ry * it), (T)(rz * it)};
            double rrt[9] = {(double)r[0] * r[0], (double)r[0] * r[1], (double)r[0] * r[2], (double)r[0] * r[1], (double)r[1] * r[1],
                             (double)r[1] * r[2], (double)r[0] * r[2], (double)r[1] * r[2], (double)r[2] * r[2]};
            double rxm[9] = {0, -(double)r[2], (double)r[1], (double)r[2], 0, -(double)r[0], -(double)r[1], (double)r[0], 0};
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
                matrix(i, j) = (T)(c * (i == j ? 1.0 : 0.0) + c1 * rrt[i * 3 + j] + s * rxm[i * 3 + j]);
        }
        matrix(0, 3) = t[0]; matrix(1, 3) = t[1]; matrix(2, 3) = t[2];
    }
    Mat3 rotation() const { Mat3 R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = matrix(i, j); return R; }
    Vec3 translation() const { return Vec3(matrix(0, 3), matrix(1, 3), matrix(2, 3)); }
    Affine3 translate(const Vec3& t) const { Mat4 m = matrix; m(0, 3) += t[0]; m(1, 3) += t[1]; m(2, 3) += t[2]; return Affine3(m); }
    Affine3 rotate(const Mat3& R) const {
        Mat4 res;
        res(3, 3) = 1;
        for (int j = 0; j < 3; ++j) {
            for (int i = 0; i < 3; ++i) {
                T v = 0;
                for (int k = 0; k < 3; ++k) v += R(j, k) * matrix(k, i);
                res(j, i) = v;
            }
            T d = 0;
            for (int k = 0; k < 3; ++k) d += R(j, k) * matrix(k, 3);
            res(j, 3) = d;
        }
        return Affine3(res);
    }
    Affine3 concatenate(const Affine3& a) const { return rotate(a.rotation()).translate(a.translation()); }
    Affine3 inv(int method = DECOMP_SVD) const {
        (void)method;
        double m[4][8];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { m[i][j] = matrix(i, j); m[i][4 + j] = (i == j) ? 1.0 : 0.0; }
        for (int c = 0; c < 4; ++c) {
            int piv = c;
            for (int r = c + 1; r < 4; ++r) if (std::fabs(m[r][c]) > std::fabs(m[piv][c])) piv = r;
            if (piv != c) for (int j = 0; j < 8; ++j) std::swap(m[c][j], m[piv][j]);
            double d = m[c][c];
            for (int j = 0; j < 8; ++j) m[c][j] /= d;
            for (int r = 0; r < 4; ++r) if (r != c) { double f = m[r][c]; for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j]; }
        }
        Mat4 o;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) o(i, j) = (T)m[i][4 + j];
        return Affine3(o);
    }
};
template <typename T> inline Affine3<T> operator*(const Affine3<T>& a, const Affine3<T>& b) { return b.concatenate(a); }
typedef Affine3<float> Affine3f;
}  // namespace cv
