// Pose algebra executed by ONE device thread (tail of the last ICP iteration, tfb_pose_set).  cv::Affine3f product
// and inverse are OpenCV calls in the reference (src/topfu.cpp:243,281-282,306); Matrix4::inv is
// include/Matrix.hpp:173-234.  Every loop is fully unrolled with compile-time indices so all matrices stay in
// registers — a single thread indexing local-memory arrays costs ~10 us per ICP iteration, 19 times per frame.
// Operation order is identical to the oracle (oracle/tfo_oracle.cpp pose_mul / pose_inv, tfo_kernels_port.cpp
// mat4_inv); translation units including this header are compiled with --fmad=false.
#pragma once
#include <cstddef>
#include "tfb_common.cuh"

namespace tfb {

__device__ __forceinline__ void pose_mul(const float (&a)[16], const float (&b)[16], float (&c)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            c[i * 4 + j] = s;
        }
}

// general 4x4 inverse, Gauss-Jordan with partial pivoting in fp64, rounded to fp32
__device__ __forceinline__ void pose_inv(const float (&a)[16], float (&o)[16]) {
    double m[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { m[i][j] = a[i * 4 + j]; m[i][4 + j] = (i == j) ? 1.0 : 0.0; }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        double best = fabs(m[c][c]);
#pragma unroll
        for (int r = c + 1; r < 4; ++r) {
            double v = fabs(m[r][c]);
            if (v > best) { best = v; piv = r; }
        }
#pragma unroll
        for (int r = c + 1; r < 4; ++r)
            if (piv == r) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { double t = m[c][j]; m[c][j] = m[r][j]; m[r][j] = t; }
            }
        const double d = m[c][c];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[c][j] /= d;
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (r != c) {
                const double f = m[r][c];
#pragma unroll
                for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j];
            }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i * 4 + j] = (float)m[i][4 + j];
}

__device__ __forceinline__ void to_colmajor(const float (&p)[16], float (&m)[16]) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) m[c * 4 + r] = p[r * 4 + c];
}

// cofactor inverse with the reference's operand order (Matrix.hpp:173-234)
__device__ __forceinline__ bool mat4_inv_cof(const float (&in)[16], float (&out)[16]) {
    constexpr unsigned char P1[12][2] = {{10, 15}, {11, 14}, {9, 15}, {11, 13}, {9, 14}, {10, 13},
                                         {8, 15},  {11, 12}, {8, 14}, {10, 12}, {8, 13}, {9, 12}};
    constexpr unsigned char P2[12][2] = {{2, 7}, {3, 6}, {1, 7}, {3, 5}, {1, 6}, {2, 5}, {0, 7}, {3, 4}, {0, 6}, {2, 4}, {0, 5}, {1, 4}};
    constexpr unsigned char C[16][12] = {
        {0, 5, 3, 6, 4, 7, 1, 5, 2, 6, 5, 7},         {1, 4, 6, 6, 9, 7, 0, 4, 7, 6, 8, 7},
        {2, 4, 7, 5, 10, 7, 3, 4, 6, 5, 11, 7},       {5, 4, 8, 5, 11, 6, 4, 4, 9, 5, 10, 6},
        {1, 1, 2, 2, 5, 3, 0, 1, 3, 2, 4, 3},         {0, 0, 7, 2, 8, 3, 1, 0, 6, 2, 9, 3},
        {3, 0, 6, 1, 11, 3, 2, 0, 7, 1, 10, 3},       {4, 0, 9, 1, 10, 2, 5, 0, 8, 1, 11, 2},
        {0, 13, 3, 14, 4, 15, 1, 13, 2, 14, 5, 15},   {1, 12, 6, 14, 9, 15, 0, 12, 7, 14, 8, 15},
        {2, 12, 7, 13, 10, 15, 3, 12, 6, 13, 11, 15}, {5, 12, 8, 13, 11, 14, 4, 12, 9, 13, 10, 14},
        {2, 10, 5, 11, 1, 9, 4, 11, 0, 9, 3, 10},     {8, 11, 0, 8, 7, 10, 6, 10, 9, 11, 1, 8},
        {6, 9, 11, 11, 3, 8, 10, 11, 2, 8, 7, 9},     {10, 10, 4, 8, 9, 9, 8, 9, 11, 10, 5, 8}};
    float s[16], t[12];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i + 4 * j] = in[i * 4 + j];
#pragma unroll
    for (int i = 0; i < 12; ++i) t[i] = s[P1[i][0]] * s[P1[i][1]];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        out[i] = (t[C[i][0]] * s[C[i][1]] + t[C[i][2]] * s[C[i][3]] + t[C[i][4]] * s[C[i][5]]) -
                 (t[C[i][6]] * s[C[i][7]] + t[C[i][8]] * s[C[i][9]] + t[C[i][10]] * s[C[i][11]]);
    const float det = s[0] * out[0] + s[1] * out[1] + s[2] * out[2] + s[3] * out[3];
    if (det == 0.0f) return false;
#pragma unroll
    for (int i = 0; i < 12; ++i) t[i] = s[P2[i][0]] * s[P2[i][1]];
#pragma unroll
    for (int i = 8; i < 16; ++i)
        out[i] = (t[C[i][0]] * s[C[i][1]] + t[C[i][2]] * s[C[i][3]] + t[C[i][4]] * s[C[i][5]]) -
                 (t[C[i][6]] * s[C[i][7]] + t[C[i][8]] * s[C[i][9]] + t[C[i][10]] * s[C[i][11]]);
    const float r = 1 / det;
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] *= r;
    return true;
}

// given pose_c2w (row-major) in registers: write it and everything derived from it into the state block
__device__ __forceinline__ void store_pose_c2w(DevState* ds, const float (&c2w)[16]) {
    float w2c[16], M[16], invM[16], Mc[16];
    pose_inv(c2w, w2c);
    to_colmajor(w2c, M);
    mat4_inv_cof(M, invM);
    to_colmajor(c2w, Mc);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        ds->pose_c2w[i] = c2w[i];
        ds->pose_w2c[i] = w2c[i];
        ds->M_w2c[i] = M[i];
        ds->invM_w2c[i] = invM[i];
        ds->M_c2w[i] = Mc[i];
    }
}

// pose_inv on a warp: lane j (mod 8) holds column j of the augmented matrix [A | I].  The pivot column is broadcast, every lane
// swaps, divides and eliminates its own column — per element exactly the operations, in exactly the order, of pose_inv above
// (one fp64 division and three multiply-subtracts per pivot instead of thirty-two of each in one thread), so the result is the
// same to the bit.  Every lane returns the whole inverse.  Must be called by all 32 lanes.
__device__ __forceinline__ void pose_inv_warp(const float (&a)[16], float (&o)[16], int lane) {
    const int j = lane & 7;
    double col[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double v = (i == j - 4) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (j == k) v = (double)a[i * 4 + k];
        col[i] = v;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double cc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) cc[i] = __shfl_sync(0xffffffffu, col[i], c);
        int piv = c;
        double best = fabs(cc[c]);
#pragma unroll
        for (int r = c + 1; r < 4; ++r) {
            const double v = fabs(cc[r]);
            if (v > best) { best = v; piv = r; }
        }
#pragma unroll
        for (int r = c + 1; r < 4; ++r)
            if (piv == r) {
                double t = col[c]; col[c] = col[r]; col[r] = t;
                t = cc[c]; cc[c] = cc[r]; cc[r] = t;
            }
        const double d = cc[c];
        col[c] /= d;
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (r != c) col[r] -= cc[r] * col[c];
    }
    float f[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) f[i] = (float)col[i];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) o[i * 4 + k] = __shfl_sync(0xffffffffu, f[i], 4 + k);
}

// store_pose_c2w by a whole warp: the inverse spread over the lanes, the 80 words of the state block written 16 at a time
__device__ __forceinline__ void store_pose_c2w_warp(DevState* ds, const float (&c2w)[16], int lane) {
    float w2c[16], M[16], invM[16], Mc[16];
    pose_inv_warp(c2w, w2c, lane);
    to_colmajor(w2c, M);
    mat4_inv_cof(M, invM);
    to_colmajor(c2w, Mc);
    // the five matrices are the first 80 floats of the state block, in this order: twenty 16-byte stores, one per lane
    static_assert(offsetof(DevState, pose_c2w) == 0 && offsetof(DevState, pose_w2c) == 64 && offsetof(DevState, M_w2c) == 128 &&
                  offsetof(DevState, invM_w2c) == 192 && offsetof(DevState, M_c2w) == 256, "layout of the pose block");
    float4* dst = reinterpret_cast<float4*>(ds);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (lane == q) dst[q] = make_float4(c2w[4 * q], c2w[4 * q + 1], c2w[4 * q + 2], c2w[4 * q + 3]);
        if (lane == 4 + q) dst[4 + q] = make_float4(w2c[4 * q], w2c[4 * q + 1], w2c[4 * q + 2], w2c[4 * q + 3]);
        if (lane == 8 + q) dst[8 + q] = make_float4(M[4 * q], M[4 * q + 1], M[4 * q + 2], M[4 * q + 3]);
        if (lane == 12 + q) dst[12 + q] = make_float4(invM[4 * q], invM[4 * q + 1], invM[4 * q + 2], invM[4 * q + 3]);
        if (lane == 16 + q) dst[16 + q] = make_float4(Mc[4 * q], Mc[4 * q + 1], Mc[4 * q + 2], Mc[4 * q + 3]);
    }
}

__device__ __forceinline__ void store_pose_w2c(DevState* ds, const float (&w2c)[16]) {
    float c2w[16], M[16], invM[16], Mc[16];
    pose_inv(w2c, c2w);
    to_colmajor(w2c, M);
    mat4_inv_cof(M, invM);
    to_colmajor(c2w, Mc);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        ds->pose_c2w[i] = c2w[i];
        ds->pose_w2c[i] = w2c[i];
        ds->M_w2c[i] = M[i];
        ds->invM_w2c[i] = invM[i];
        ds->M_c2w[i] = Mc[i];
    }
}

}  // namespace tfb
