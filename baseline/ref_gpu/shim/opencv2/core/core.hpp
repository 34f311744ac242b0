// REFERENCE-ARM INFRASTRUCTURE (not product code).
// Minimal stand-in for the parts of OpenCV `core` that the reference library
// (/root/reference/tfusion) names.  OpenCV C++ is not installed in this image
// (SURVEY.md §8b), so the patched reference GPU build under baseline/_ref links
// against this header instead.  Only what the reference's hot path calls exists:
//   Matx / Vec / Ptr / Mat (storage only), determinant, solve(DECOMP_SVD),
//   tick counters, CV_Assert, the CV_xxCn type codes.
// Arithmetic follows OpenCV's published behaviour (core/matx.hpp, core/affine.hpp,
// lapack.cpp): fp32 LU determinant returned as double; SVD solve with the
// back-substitution cut 2*FLT_EPSILON*sum(w); tests/test_refgpu_shim.py pins both
// against cv2 4.13.
#pragma once
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <cfloat>
#include <chrono>
#include <memory>
#include <vector>
#include <stdexcept>
#include <algorithm>

#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)
#define CV_Assert(expr) do { if (!(expr)) { std::fprintf(stderr, "CV_Assert failed: %s (%s:%d)\n", #expr, __FILE__, __LINE__); std::abort(); } } while (0)

namespace cv {

enum { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_EIG = 2, DECOMP_CHOLESKY = 3, DECOMP_QR = 4 };

template <typename T, int m, int n> struct Matx {
    enum { rows = m, cols = n, channels = m * n };
    T val[m * n];
    Matx() { for (int i = 0; i < m * n; ++i) val[i] = T(0); }
    static Matx all(T a) { Matx r; for (int i = 0; i < m * n; ++i) r.val[i] = a; return r; }
    static Matx eye() { Matx r; for (int i = 0; i < (m < n ? m : n); ++i) r.val[i * n + i] = T(1); return r; }
    T& operator()(int r, int c) { return val[r * n + c]; }
    const T& operator()(int r, int c) const { return val[r * n + c]; }
};
typedef Matx<float, 3, 3> Matx33f;
typedef Matx<float, 4, 4> Matx44f;
typedef Matx<float, 6, 6> Matx66f;

template <typename T, int n> struct Vec : Matx<T, n, 1> {
    Vec() {}
    Vec(T a, T b, T c) { static_assert(n == 3, "Vec3 only"); this->val[0] = a; this->val[1] = b; this->val[2] = c; }
    explicit Vec(const T* p) { for (int i = 0; i < n; ++i) this->val[i] = p[i]; }
    static Vec all(T a) { Vec r; for (int i = 0; i < n; ++i) r.val[i] = a; return r; }
    T& operator[](int i) { return this->val[i]; }
    const T& operator[](int i) const { return this->val[i]; }
};
typedef Vec<float, 3> Vec3f;
typedef Vec<int, 3> Vec3i;
typedef Vec<float, 6> Vec6f;

// cv::Ptr: shared ownership (OpenCV >= 3 is std::shared_ptr based).
template <typename T> struct Ptr : std::shared_ptr<T> {
    Ptr() {}
    Ptr(T* p) : std::shared_ptr<T>(p) {}
};

// cv::Mat: the reference only uses it as a host landing buffer for debug downloads
// (topfu.cpp:212-223, :284-288).
struct Mat {
    int rows, cols, type_;
    size_t step;
    unsigned char* data;
    std::shared_ptr<std::vector<unsigned char>> store;
    Mat() : rows(0), cols(0), type_(0), step(0), data(0) {}
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type) {
        static const int depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
        size_t esz = (size_t)depth_bytes[type & 7] * ((type >> 3) + 1);
        step = esz * (size_t)c;
        store = std::make_shared<std::vector<unsigned char>>(step * (size_t)r);
        data = store->data();
    }
};

inline long long getTickCount() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline double getTickFrequency() { return 1e9; }

// cv::determinant on a Matx66f: LU with partial pivoting on a float copy, product in double
// (lapack.cpp LUImpl + determinant()).
inline double determinant(const Matx66f& A) {
    float a[36];
    std::memcpy(a, A.val, sizeof(a));
    double p = 1;
    for (int i = 0; i < 6; ++i) {
        int k = i;
        for (int j = i + 1; j < 6; ++j) if (std::fabs(a[j * 6 + i]) > std::fabs(a[k * 6 + i])) k = j;
        if (std::fabs(a[k * 6 + i]) < FLT_EPSILON * 10) return 0;
        if (k != i) { for (int j = i; j < 6; ++j) std::swap(a[i * 6 + j], a[k * 6 + j]); p = -p; }
        float d = -1 / a[i * 6 + i];
        for (int j = i + 1; j < 6; ++j) {
            float alpha = a[j * 6 + i] * d;
            for (int c = i + 1; c < 6; ++c) a[j * 6 + c] += alpha * a[i * 6 + c];
        }
    }
    for (int i = 0; i < 6; ++i) p *= a[i * 6 + i];
    return p;
}

// cv::solve(A, b, x, DECOMP_SVD) for the symmetric positive semi-definite 6x6 normal
// matrix of the ICP: one-sided Jacobi SVD, then SVBkSb with the cut 2*eps*sum(w).
// For a symmetric matrix the SVD is the eigen-decomposition up to signs; it is evaluated in
// double and rounded at the end, which stays inside the fp32 Jacobi's own error.
inline bool solve(const Matx66f& A, const Vec6f& b, Vec6f& x, int method) {
    (void)method;
    double a[6][6], v[6][6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { a[i][j] = A.val[i * 6 + j]; v[i][j] = (i == j); }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int i = 0; i < 6; ++i) for (int j = i + 1; j < 6; ++j) off += a[i][j] * a[i][j];
        if (off < 1e-300) break;
        for (int p = 0; p < 5; ++p) for (int q = p + 1; q < 6; ++q) {
            if (a[p][q] == 0) continue;
            double th = (a[q][q] - a[p][p]) / (2 * a[p][q]);
            double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1));
            double c = 1 / std::sqrt(t * t + 1), s = t * c;
            for (int k = 0; k < 6; ++k) { double x0 = a[k][p], x1 = a[k][q]; a[k][p] = c * x0 - s * x1; a[k][q] = s * x0 + c * x1; }
            for (int k = 0; k < 6; ++k) { double x0 = a[p][k], x1 = a[q][k]; a[p][k] = c * x0 - s * x1; a[q][k] = s * x0 + c * x1; }
            for (int k = 0; k < 6; ++k) { double x0 = v[k][p], x1 = v[k][q]; v[k][p] = c * x0 - s * x1; v[k][q] = s * x0 + c * x1; }
        }
    }
    double wsum = 0;
    for (int i = 0; i < 6; ++i) wsum += std::fabs(a[i][i]);
    double thr = wsum * 2 * (double)FLT_EPSILON, r[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; ++i) {
        double w = a[i][i];
        if (std::fabs(w) <= thr) continue;
        double s = 0;
        for (int k = 0; k < 6; ++k) s += v[k][i] * b.val[k];
        s /= w;
        for (int k = 0; k < 6; ++k) r[k] += s * v[k][i];
    }
    for (int k = 0; k < 6; ++k) x.val[k] = (float)r[k];
    return true;
}

}  // namespace cv
