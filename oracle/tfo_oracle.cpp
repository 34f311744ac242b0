// TEST INFRASTRUCTURE (see tfo_types.h).  CPU oracle for topfusion's per-frame hot path
// tfusion::TopFu::operator() (/root/reference/tfusion/src/topfu.cpp:161-330): stage
// functions, the kernel glue of the two engines, and the frame orchestrator.  Serial and
// deterministic; OpenMP is used only where iterations are independent and write disjoint
// outputs (never in allocation, which depends on raster order — SURVEY.md F4).
// Compile with -ffp-contract=off.
#include "tfo_types.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace tfo {

static const float QNAN = std::numeric_limits<float>::quiet_NaN();

// ---------------------------------------------------------------------------------------
// Image stages — restatement of src/cuda/imgproc.cu (device-only in the reference)
// ---------------------------------------------------------------------------------------

// imgproc.cu:263-280.  Note the reference ignores its lambda / intrinsics arguments.
void compute_dists(const uint16_t* depth, float* dists, int w, int h, int cutoff_mm) {
#pragma omp parallel for
    for (int i = 0; i < w * h; ++i) {
        int d = depth[i];
        dists[i] = (d >= cutoff_mm || d <= 0) ? -1.0f : d * 0.001f;
    }
}

// imgproc.cu:10-61.  Window [x-k/2, min(x-k/2+k, cols-1)) — right/bottom edge exclusive,
// raw zeros take part.  __expf/approximate division on the device -> tolerance-compared.
void bilateral(const uint16_t* src, uint16_t* dst, int w, int h, int ksz, float sigma_spatial, float sigma_depth_m) {
    float sigma_depth = sigma_depth_m * 1000;
    float ss = 0.5f / (sigma_spatial * sigma_spatial);
    float sd = 0.5f / (sigma_depth * sigma_depth);
#pragma omp parallel for
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int value = src[y * w + x];
            int tx = std::min(x - ksz / 2 + ksz, w - 1);
            int ty = std::min(y - ksz / 2 + ksz, h - 1);
            float sum1 = 0, sum2 = 0;
            for (int cy = std::max(y - ksz / 2, 0); cy < ty; ++cy)
                for (int cx = std::max(x - ksz / 2, 0); cx < tx; ++cx) {
                    int depth = src[cy * w + cx];
                    float space2 = (float)((x - cx) * (x - cx) + (y - cy) * (y - cy));
                    // 32-bit wrap like the device's integer multiply
                    unsigned diff = (unsigned)(value - depth);
                    float color2 = (float)(int)(diff * diff);
                    float weight = expf(-(space2 * ss + color2 * sd));
                    sum1 += depth * weight;
                    sum2 += weight;
                }
            dst[y * w + x] = (uint16_t)(int)rintf(sum1 / sum2);
        }
}

// imgproc.cu:70-89
void truncate_depth(uint16_t* depth, int w, int h, float max_dist_m) {
    uint16_t lim = (uint16_t)(max_dist_m * 1000.f);
    for (int i = 0; i < w * h; ++i)
        if (depth[i] > lim) depth[i] = 0;
}

// imgproc.cu:98-140; dst is (w/2, h/2)
void depth_pyr(const uint16_t* src, uint16_t* dst, int sw, int sh, float sigma_depth_m) {
    float thr = sigma_depth_m * 1000 * 3;
    int dw = sw / 2, dh = sh / 2;
    const int D = 5;
#pragma omp parallel for
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            int center = src[2 * y * sw + 2 * x];
            int tx = std::min(2 * x - D / 2 + D, sw - 1);
            int ty = std::min(2 * y - D / 2 + D, sh - 1);
            int sum = 0, count = 0;
            for (int cy = std::max(0, 2 * y - D / 2); cy < ty; ++cy)
                for (int cx = std::max(0, 2 * x - D / 2); cx < tx; ++cx) {
                    int val = src[cy * sw + cx];
                    if (std::abs(val - center) < thr) { sum += val; ++count; }
                }
            dst[y * dw + x] = (uint16_t)((count == 0) ? 0 : sum / count);
        }
}

// imgproc.cu:214-254, Reprojector device.hpp:43-48 (finv = 1/f, precomp.cpp:55)
void points_normals(const uint16_t* depth, float* points, float* normals, int w, int h, float fx, float fy, float cx,
                    float cy) {
    float finvx = 1.f / fx, finvy = 1.f / fy;
#pragma omp parallel for
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float* p = points + 4 * (y * w + x);
            float* n = normals + 4 * (y * w + x);
            for (int i = 0; i < 4; ++i) p[i] = n[i] = QNAN;
            if (x >= w - 1 || y >= h - 1) continue;
            float z00 = depth[y * w + x] * 0.001f;
            float z01 = depth[y * w + x + 1] * 0.001f;
            float z10 = depth[(y + 1) * w + x] * 0.001f;
            if (z00 * z01 * z10 != 0) {
                float v00[3] = {z00 * (x - cx) * finvx, z00 * (y - cy) * finvy, z00};
                float v01[3] = {z01 * (x + 1 - cx) * finvx, z01 * (y - cy) * finvy, z01};
                float v10[3] = {z10 * (x - cx) * finvx, z10 * (y + 1 - cy) * finvy, z10};
                float a[3] = {v01[0] - v00[0], v01[1] - v00[1], v01[2] - v00[2]};
                float b[3] = {v10[0] - v00[0], v10[1] - v00[1], v10[2] - v00[2]};
                float c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
                // temp_utils.hpp:27-30,101-104: dot via fma chain, rsqrt (approximate on device)
                float d = fmaf(c[0], c[0], fmaf(c[1], c[1], c[2] * c[2]));
                float r = 1.0f / sqrtf(d);
                n[0] = -(c[0] * r); n[1] = -(c[1] * r); n[2] = -(c[2] * r); n[3] = 1.0f;
                p[0] = v00[0]; p[1] = v00[1]; p[2] = v00[2]; p[3] = 1.0f;
            }
        }
}

// imgproc.cu:355-401; dst is (sw/2, sh/2); normals are NOT renormalised
void resize_points_normals(const float* vsrc, const float* nsrc, float* vdst, float* ndst, int sw, int sh) {
    int dw = sw / 2, dh = sh / 2;
#pragma omp parallel for
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            float* vo = vdst + 4 * (y * dw + x);
            float* no = ndst + 4 * (y * dw + x);
            vo[0] = vo[1] = vo[2] = no[0] = no[1] = no[2] = QNAN;
            vo[3] = no[3] = 0.f;
            int xs = 2 * x, ys = 2 * y;
            const float* d00 = vsrc + 4 * (ys * sw + xs);
            const float* d01 = d00 + 4;
            const float* d10 = vsrc + 4 * ((ys + 1) * sw + xs);
            const float* d11 = d10 + 4;
            if (!std::isnan(d00[0] * d01[0] * d10[0] * d11[0])) {
                for (int i = 0; i < 3; ++i) vo[i] = (d00[i] + d01[i] + d10[i] + d11[i]) * 0.25f;
                vo[3] = 1.0f;
                const float* n00 = nsrc + 4 * (ys * sw + xs);
                const float* n01 = n00 + 4;
                const float* n10 = nsrc + 4 * ((ys + 1) * sw + xs);
                const float* n11 = n10 + 4;
                for (int i = 0; i < 3; ++i) no[i] = (n00[i] + n01[i] + n10[i] + n11[i]) * 0.25f;
                no[3] = 0.f;
            }
        }
}

// ---------------------------------------------------------------------------------------
// Projective point-to-plane ICP — src/cuda/proj_icp.cu + src/projective_icp.cpp
// ---------------------------------------------------------------------------------------

struct Pose {  // row-major 4x4 like cv::Matx44f inside cv::Affine3f
    float m[16];
    static Pose identity() { Pose p; memset(p.m, 0, sizeof(p.m)); p.m[0] = p.m[5] = p.m[10] = p.m[15] = 1.f; return p; }
};

// Affine3f * Affine3f = Matx44f product, fp32, k-sequential (OpenCV matx.hpp MatxMulOp)
Pose pose_mul(const Pose& a, const Pose& b) {
    Pose c;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0;
            for (int k = 0; k < 4; ++k) s += a.m[i * 4 + k] * b.m[k * 4 + j];
            c.m[i * 4 + j] = s;
        }
    return c;
}

// cv::Affine3f::inv() calls Matx44f::inv(DECOMP_SVD) — OpenCV is outside the reference tree
// (SURVEY.md Appendix B).  Restated as a general 4x4 Gauss-Jordan inverse in fp64 rounded to
// fp32; tests/test_oracle_opencv.py pins it against cv2.invert(DECOMP_SVD) to 1e-6.
Pose pose_inv(const Pose& a) {
    double m[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { m[i][j] = a.m[i * 4 + j]; m[i][4 + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r) if (fabs(m[r][c]) > fabs(m[piv][c])) piv = r;
        if (piv != c) for (int j = 0; j < 8; ++j) std::swap(m[c][j], m[piv][j]);
        double d = m[c][c];
        for (int j = 0; j < 8; ++j) m[c][j] /= d;
        for (int r = 0; r < 4; ++r) if (r != c) { double f = m[r][c]; for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j]; }
    }
    Pose o;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) o.m[i * 4 + j] = (float)m[i][4 + j];
    return o;
}

// cv::Affine3f(rvec, t): Rodrigues evaluated in double, stored as float (Appendix B);
// pinned against cv2.Rodrigues in tests/test_oracle_opencv.py.
Pose pose_from_rvec_t(const float rv[3], const float t[3]) {
    Pose p = Pose::identity();
    double theta = sqrt((double)rv[0] * rv[0] + (double)rv[1] * rv[1] + (double)rv[2] * rv[2]);
    if (theta >= std::numeric_limits<double>::epsilon()) {
        double c = cos(theta), s = sin(theta), c1 = 1. - c;
        double it = 1. / theta;
        float r[3] = {(float)(rv[0] * it), (float)(rv[1] * it), (float)(rv[2] * it)};
        double rrt[9] = {(double)r[0] * r[0], (double)r[0] * r[1], (double)r[0] * r[2], (double)r[0] * r[1], (double)r[1] * r[1],
                         (double)r[1] * r[2], (double)r[0] * r[2], (double)r[1] * r[2], (double)r[2] * r[2]};
        double rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                int k = i * 3 + j;
                p.m[i * 4 + j] = (float)(c * (i == j ? 1.0 : 0.0) + c1 * rrt[k] + s * rx[k]);
            }
    }
    p.m[3] = t[0]; p.m[7] = t[1]; p.m[11] = t[2];
    return p;
}

// cv::determinant(Matx66f): LU with partial pivoting in fp32, product in double.
double det6_f32(const float A[36]) {
    float a[36];
    memcpy(a, A, sizeof(a));
    double p = 1;
    for (int i = 0; i < 6; ++i) {
        int k = i;
        for (int j = i + 1; j < 6; ++j) if (fabsf(a[j * 6 + i]) > fabsf(a[k * 6 + i])) k = j;
        if (fabsf(a[k * 6 + i]) < std::numeric_limits<float>::epsilon() * 10) return 0;
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

// cv::solve(A, b, r, DECOMP_SVD) on a symmetric 6x6: least-norm solution through the
// eigen-decomposition (cyclic Jacobi, fp64) with OpenCV's back-substitution threshold
// 2*FLT_EPSILON*sum(w) (SVBkSb).  Pinned against cv2.solve in tests/test_oracle_opencv.py.
void solve6_svd(const float A[36], const float b[6], float x[6]) {
    double a[6][6], v[6][6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) { a[i][j] = A[i * 6 + j]; v[i][j] = (i == j); }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int i = 0; i < 6; ++i) for (int j = i + 1; j < 6; ++j) off += a[i][j] * a[i][j];
        if (off < 1e-300) break;
        for (int p = 0; p < 5; ++p)
            for (int q = p + 1; q < 6; ++q) {
                if (a[p][q] == 0) continue;
                double th = (a[q][q] - a[p][p]) / (2 * a[p][q]);
                double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
                double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 6; ++k) { double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 6; ++k) { double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk; }
                for (int k = 0; k < 6; ++k) { double vkp = v[k][p], vkq = v[k][q]; v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq; }
            }
    }
    double wsum = 0;
    for (int i = 0; i < 6; ++i) wsum += fabs(a[i][i]);
    double thr = wsum * 2 * (double)std::numeric_limits<float>::epsilon();
    double r[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; ++i) {
        double w = a[i][i];
        if (fabs(w) <= thr) continue;
        double s = 0;
        for (int k = 0; k < 6; ++k) s += v[k][i] * b[k];
        s /= w;
        for (int k = 0; k < 6; ++k) r[k] += s * v[k][i];
    }
    for (int k = 0; k < 6; ++k) x[k] = (float)r[k];
}

struct IcpSetup {
    float min_cosine, dist2_thres;
    float fx, fy, cx, cy;
    int w, h;
};

static inline float dot3(const float a[3], const float b[3]) { return fmaf(a[0], b[0], fmaf(a[1], b[1], a[2] * b[2])); }

// proj_icp.cu:80-117 (points variant).  aff is the 3x4 [R|t] of the current estimate.
static inline int find_coresp(const IcpSetup& s, const float R[9], const float t[3], const float* vcurr, const float* ncurr,
                              const float* vprev, const float* nprev, int x, int y, float n[3], float d[3], float sp[3]) {
    const float* v = vcurr + 4 * (y * s.w + x);
    if (std::isnan(v[0])) return 40;
    float q[3] = {v[0], v[1], v[2]};
    sp[0] = dot3(R + 0, q) + t[0];
    sp[1] = dot3(R + 3, q) + t[1];
    sp[2] = dot3(R + 6, q) + t[2];
    float cooX = fmaf(s.fx, sp[0] / sp[2], s.cx);
    float cooY = fmaf(s.fy, sp[1] / sp[2], s.cy);
    if (sp[2] <= 0 || cooX < 0 || cooY < 0 || cooX >= s.w || cooY >= s.h) return 80;
    // point-sampled, unnormalised texture fetch: texel (floor(x), floor(y))
    int tx = (int)floorf(cooX), ty = (int)floorf(cooY);
    const float* dv = vprev + 4 * (ty * s.w + tx);
    if (std::isnan(dv[0])) return 120;
    d[0] = dv[0]; d[1] = dv[1]; d[2] = dv[2];
    float df[3] = {sp[0] - d[0], sp[1] - d[1], sp[2] - d[2]};
    if (dot3(df, df) > s.dist2_thres) return 160;
    const float* nc = ncurr + 4 * (y * s.w + x);
    float ncv[3] = {nc[0], nc[1], nc[2]};
    float ns[3] = {dot3(R + 0, ncv), dot3(R + 3, ncv), dot3(R + 6, ncv)};
    const float* nd = nprev + 4 * (ty * s.w + tx);
    n[0] = nd[0]; n[1] = nd[1]; n[2] = nd[2];
    if (fabsf(dot3(ns, n)) < s.min_cosine) return 200;
    return 0;
}

// temp_utils.hpp:503-523 — the 256-wide shared-memory tree, in the order the device sums
static float tree256(float* buf) {
    for (int half = 128; half >= 1; half >>= 1)
        for (int t = 0; t < half; ++t) buf[t] = buf[t] + buf[t + half];
    return buf[0];
}

// proj_icp.cu:120-403: per-CTA (32x8 tile) tree partials for the 27 products, then the
// strided final reduce.  out27 is in the reference's packed order A00..A05,b0,A11..,b5.
void icp_reduce(const IcpSetup& s, const Pose& aff, const float* vcurr, const float* ncurr, const float* vprev,
                const float* nprev, float out27[27], int* n_corresp) {
    float R[9] = {aff.m[0], aff.m[1], aff.m[2], aff.m[4], aff.m[5], aff.m[6], aff.m[8], aff.m[9], aff.m[10]};
    float t[3] = {aff.m[3], aff.m[7], aff.m[11]};
    int gx = (s.w + 31) / 32, gy = (s.h + 7) / 8;
    int ncta = gx * gy;
    std::vector<float> partial((size_t)27 * ncta);
    int count = 0;
#pragma omp parallel for reduction(+ : count)
    for (int cta = 0; cta < ncta; ++cta) {
        int bx = cta % gx, by = cta / gx;
        float rows[256][7];
        for (int tid = 0; tid < 256; ++tid) {
            int x = bx * 32 + (tid & 31), y = by * 8 + (tid >> 5);
            float n[3], d[3], sp[3];
            int filtered = (x < s.w && y < s.h) ? find_coresp(s, R, t, vcurr, ncurr, vprev, nprev, x, y, n, d, sp) : 1;
            float* r = rows[tid];
            if (!filtered) {
                r[0] = sp[1] * n[2] - sp[2] * n[1];
                r[1] = sp[2] * n[0] - sp[0] * n[2];
                r[2] = sp[0] * n[1] - sp[1] * n[0];
                r[3] = n[0]; r[4] = n[1]; r[5] = n[2];
                float ds[3] = {d[0] - sp[0], d[1] - sp[1], d[2] - sp[2]};
                r[6] = dot3(n, ds);
                ++count;
            } else {
                for (int i = 0; i < 7; ++i) r[i] = 0.f;
            }
        }
        float buf[256];
        int k = 0;
        for (int i = 0; i < 6; ++i)
            for (int j = i; j < 7; ++j) {
                for (int tid = 0; tid < 256; ++tid) buf[tid] = rows[tid][i] * rows[tid][j];
                partial[(size_t)k * ncta + cta] = tree256(buf);
                ++k;
            }
    }
    for (int k = 0; k < 27; ++k) {
        float buf[256];
        for (int tid = 0; tid < 256; ++tid) {
            float sum = 0.f;
            for (int i = tid; i < ncta; i += 256) sum += partial[(size_t)k * ncta + i];
            buf[tid] = sum;
        }
        out27[k] = tree256(buf);
    }
    if (n_corresp) *n_corresp = count;
}

// projective_icp.cpp:43-62 unpack + :197-209 solve/update
bool icp_solve_update(const float v27[27], Pose& affine) {
    float A[36], b[6];
    int shift = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 7; ++j) {
            float value = v27[shift++];
            if (j == 6) b[i] = value;
            else A[j * 6 + i] = A[i * 6 + j] = value;
        }
    double det = det6_f32(A);
    if (fabs(det) < 1e-15 || std::isnan(det)) return false;
    float r[6];
    solve6_svd(A, b, r);
    Pose tinc = pose_from_rvec_t(r, r + 3);
    affine = pose_mul(tinc, affine);
    return true;
}

// ---------------------------------------------------------------------------------------
// The oracle object: scene (hash + block pool), render state, frames, poses
// ---------------------------------------------------------------------------------------

struct Level {
    int w = 0, h = 0;
    std::vector<uint16_t> depth;
    std::vector<float> points, normals;
};

static void to_colmajor(const Pose& p, float m[16]) {
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m[c * 4 + r] = p.m[r * 4 + c];
}

struct Oracle {
    tfo_params p;
    HashGeom g;
    // Scene<Voxel_s, VoxelBlockHash>
    std::vector<HashEntry> table;
    std::vector<Voxel> vba;
    std::vector<int> vba_free, excess_free;
    int last_free_block = 0, last_free_excess = 0;
    // SceneReconstructionEngine_CUDA scratch (SceneReconstructionEngine_host.hpp:59-66)
    std::vector<uint8_t> alloc_type;
    std::vector<int16_t> block_coords;
    // RenderState_VH (RenderState_VH.hpp:35-46); zero-initialised here
    std::vector<uint8_t> vis_type;
    std::vector<int> visible_ids;
    int n_visible = 0;
    std::vector<float> minmax;   // Vector2f [rows x cols], only the (cols/8 x rows/8) corner is read
    std::vector<float> raycast;  // Vector4f [rows x cols]
    std::vector<RenderTile> tiles;
    int n_tiles = 0;
    // TopFu
    Level curr[4], prev[4];
    std::vector<float> dists;
    std::vector<Pose> poses;
    int frame_counter = 0;
    // statistics of the last frame
    long long voxel_updates = 0;
    int icp_corresp_last = 0;
    int resets = 0;

    explicit Oracle(const tfo_params& pp) : p(pp) {
        g.num_buckets = p.num_buckets; g.hash_mask = p.num_buckets - 1; g.excess_size = p.excess_size;
        int total = g.total_entries();
        table.resize(total); vba.resize((size_t)p.num_blocks * BLOCK3);
        vba_free.resize(p.num_blocks); excess_free.resize(p.excess_size);
        alloc_type.assign(total, 0); block_coords.assign((size_t)4 * total, 0);
        vis_type.assign(total, 0); visible_ids.assign(total, 0);
        minmax.resize((size_t)2 * p.cols * p.rows);
        // RenderState ctor fills the range image with (vf_min, vf_max) (RenderState.hpp:67-73)
        for (size_t i = 0; i < (size_t)p.cols * p.rows; ++i) { minmax[2 * i] = p.view_frustum_min; minmax[2 * i + 1] = p.view_frustum_max; }
        raycast.assign((size_t)4 * p.cols * p.rows, 0.f);
        tiles.resize(MAX_TILES);
        int w = p.cols, h = p.rows;
        for (int l = 0; l < 4; ++l) {
            for (Level* L : {&curr[l], &prev[l]}) {
                L->w = w; L->h = h;
                L->depth.assign((size_t)w * h, 0);
                L->points.assign((size_t)4 * w * h, 0.f);
                L->normals.assign((size_t)4 * w * h, 0.f);
            }
            w /= 2; h /= 2;
        }
        dists.assign((size_t)p.cols * p.rows, 0.f);
        reset_scene();
        reset();
    }

    int levels() const {  // ProjectiveICP::getUsedLevelsNum, projective_icp.cpp:110-115
        int i = 3;
        for (; i >= 0 && !p.icp_iters[i]; --i) {}
        return i + 1;
    }

    // SceneReconstructionEngine_host.cu:52-73
    void reset_scene() {
        Voxel v0; v0.sdf = 32767; v0.w_depth = 0; v0.pad_ = 0;
        std::fill(vba.begin(), vba.end(), v0);
        for (int i = 0; i < p.num_blocks; ++i) vba_free[i] = i;
        last_free_block = p.num_blocks - 1;
        HashEntry e; memset(&e, 0, sizeof(e)); e.ptr = -2;
        std::fill(table.begin(), table.end(), e);
        for (int i = 0; i < p.excess_size; ++i) excess_free[i] = i;
        last_free_excess = p.excess_size - 1;
    }

    // TopFu::reset, topfu.cpp:141-152 (render state is NOT cleared — kept as in the reference)
    void reset() {
        if (frame_counter) ++resets;
        frame_counter = 0;
        poses.clear();
        poses.push_back(Pose::identity());
        reset_scene();
    }

    bool owns(const int16_t* bc) const {
        if (p.shard_count <= 1) return true;
        unsigned hsh = ((unsigned)(int)bc[0] * 0x9E3779B1u) ^ ((unsigned)(int)bc[1] * 0x85EBCA77u) ^ ((unsigned)(int)bc[2] * 0xC2B2AE3Du);
        hsh ^= hsh >> 15; hsh *= 0x2C1B3C6Du; hsh ^= hsh >> 12;
        return (int)(hsh % (unsigned)p.shard_count) == p.shard_rank;
    }

    // AllocateSceneFromDepth, SceneReconstructionEngine_host.cu:76-195.  pose_w2c is what the
    // reference passes as `pose` (camera<-world); dists is the raw-depth metres image.
    void allocate(const Pose& pose_w2c, const float* dd) {
        float M[16], invM[16];
        to_colmajor(pose_w2c, M);
        k::mat4_inv(M, invM);
        float proj[4] = {p.fx, p.fy, p.cx, p.cy};
        float inv_proj[4] = {1.0f / p.fx, 1.0f / p.fy, p.cx, p.cy};
        float one_over_block = 1.0f / (p.voxel_size * BLOCK);
        int total = g.total_entries();

        std::fill(alloc_type.begin(), alloc_type.end(), 0);
        for (int i = 0; i < n_visible; ++i) vis_type[visible_ids[i]] = 3;  // setToType3, :343-348

        for (int y = 0; y < p.rows; ++y)  // raster order = the serial resolution of the F4 race
            for (int x = 0; x < p.cols; ++x)
                k::mark_pixel(alloc_type.data(), vis_type.data(), x, y, block_coords.data(), dd, invM, inv_proj, p.mu, p.cols,
                              p.rows, one_over_block, table.data(), p.view_frustum_min, p.view_frustum_max, g);

        // allocateVoxelBlocksList_device, :350-415 (atomicSub returns the old value)
        for (int s = 0; s < total; ++s) {
            int at = alloc_type[s];
            if (at == 0) continue;
            const int16_t* bc = &block_coords[4 * s];
            // Sharded scene (new): the index is replicated on every rank so the
            // admission order is identical everywhere; only the owner stores voxels.  A foreign
            // block gets ptr = -1, the reference's "swapped out" state, which integrate and
            // findVoxel already skip (VoxelBlockHash.hpp:38-43).
            bool mine = owns(bc);
            if (at == 1) {
                int vi = mine ? last_free_block-- : 0;
                if (vi >= 0) {
                    HashEntry e; e.pos[0] = bc[0]; e.pos[1] = bc[1]; e.pos[2] = bc[2]; e.pad_ = 0;
                    e.ptr = mine ? vba_free[vi] : -1; e.offset = 0;
                    table[s] = e;
                } else {
                    vis_type[s] = 0;
                    ++last_free_block;
                }
            } else if (at == 2) {
                int vi = mine ? last_free_block-- : 0;
                int ei = last_free_excess--;
                if (vi >= 0 && ei >= 0) {
                    HashEntry e; e.pos[0] = bc[0]; e.pos[1] = bc[1]; e.pos[2] = bc[2]; e.pad_ = 0;
                    e.ptr = mine ? vba_free[vi] : -1; e.offset = 0;
                    int off = excess_free[ei];
                    table[s].offset = off + 1;
                    table[g.num_buckets + off] = e;
                    vis_type[g.num_buckets + off] = 1;
                } else {
                    if (mine) ++last_free_block;
                    ++last_free_excess;
                }
            }
        }

        // buildVisibleList_device<false>, :434-479 (list order = ascending slot here)
        int nv = 0;
        for (int s = 0; s < total; ++s) {
            uint8_t t = vis_type[s];
            if (t == 3) {
                if (!k::block_visible(table[s].pos, M, proj, p.voxel_size, p.cols, p.rows)) t = 0;
                vis_type[s] = t;
            }
            if (t > 0) visible_ids[nv++] = s;
        }
        n_visible = nv;
    }

    // IntegrateIntoScene + integrateIntoScene_device, :198-251, :297-329
    void integrate(const Pose& pose_w2c, const float* dd) {
        voxel_updates = 0;
        if (n_visible == 0) return;
        float M[16];
        to_colmajor(pose_w2c, M);
        float proj[4] = {p.fx, p.fy, p.cx, p.cy};
        long long upd = 0;
#pragma omp parallel for reduction(+ : upd) schedule(dynamic, 64)
        for (int i = 0; i < n_visible; ++i) {
            const HashEntry& e = table[visible_ids[i]];
            if (e.ptr < 0) continue;
            int gx = e.pos[0] * BLOCK, gy = e.pos[1] * BLOCK, gz = e.pos[2] * BLOCK;
            Voxel* blk = &vba[(size_t)e.ptr * BLOCK3];
            for (int z = 0; z < BLOCK; ++z)
                for (int y = 0; y < BLOCK; ++y)
                    for (int x = 0; x < BLOCK; ++x) {
                        int id = x + y * BLOCK + z * BLOCK * BLOCK;
                        if (p.stop_integrating_at_max_w && blk[id].w_depth == p.max_w) continue;
                        float pt[4] = {(float)(gx + x) * p.voxel_size, (float)(gy + y) * p.voxel_size, (float)(gz + z) * p.voxel_size, 1.0f};
                        k::update_voxel(blk[id], pt, M, proj, p.mu, p.max_w, dd, p.cols, p.rows);
                    }
            upd += BLOCK3;
        }
        voxel_updates = upd;
    }

    // CreateExpectedDepths, VisualisationEngine_CUDA.cu:120-173 with projectAndSplitBlocks_device /
    // fillBlocks_device (VisualisationHelper.cu:52-77,105-121)
    void expected_depths(const Pose& pose_w2c) {
        size_t npx = (size_t)p.cols * p.rows;
        for (size_t i = 0; i < npx; ++i) { minmax[2 * i] = FAR_AWAY_F; minmax[2 * i + 1] = VERY_CLOSE_F; }
        n_tiles = 0;
        if (n_visible == 0) return;
        float M[16];
        to_colmajor(pose_w2c, M);
        float proj[4] = {p.fx, p.fy, p.cx, p.cy};
        int off = 0;
        for (int i = 0; i < n_visible; ++i) {
            const HashEntry& e = table[visible_ids[i]];
            if (e.ptr < 0) continue;
            int ul[2], lr[2]; float z[2];
            if (!k::project_block(e.pos, M, proj, p.cols, p.rows, p.voxel_size, ul, lr, z)) continue;
            int nx = (int)ceilf((float)(lr[0] - ul[0] + 1) / TILE), ny = (int)ceilf((float)(lr[1] - ul[1] + 1) / TILE);
            if (off + nx * ny > MAX_TILES) continue;  // :73, whole block dropped when it would overflow
            off = k::split_tiles(tiles.data(), off, ul, lr, z);
        }
        n_tiles = std::min(off, (int)MAX_TILES);
        for (int t = 0; t < n_tiles; ++t) {
            const RenderTile& b = tiles[t];
            for (int y = b.ul[1]; y <= b.lr[1] && y < b.ul[1] + TILE; ++y)
                for (int x = b.ul[0]; x <= b.lr[0] && x < b.ul[0] + TILE; ++x) {
                    float* px = &minmax[2 * ((size_t)x + (size_t)y * p.cols)];
                    px[0] = fminf(px[0], b.z[0]);
                    px[1] = fmaxf(px[1], b.z[1]);
                }
        }
    }

    // GenericRaycast(updateVisibleList) + genericRaycast_device, VisualisationEngine_CUDA.cu:176-218,
    // VisualisationHelper.hpp:33-46.  pose_c2w is the matrix the reference passes as invM.
    void raycast_pass(const Pose& pose_c2w, bool update_visible) {
        float invM[16];
        to_colmajor(pose_c2w, invM);
        float inv_proj[4] = {1.0f / p.fx, 1.0f / p.fy, -p.cx, -p.cy};  // InvertProjectionParams, Shared.hpp:28-31
        float one_over_voxel = 1.0f / p.voxel_size;
        uint8_t* vt = update_visible ? vis_type.data() : nullptr;
        // rays only ever store the constant 1 into vis_type, so the parallel loop is race-free in effect
#pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < p.rows; ++y)
            for (int x = 0; x < p.cols; ++x) {
                int id2 = (int)floorf((float)x / MINMAX_SUBSAMPLE) + (int)floorf((float)y / MINMAX_SUBSAMPLE) * p.cols;
                k::cast_ray(&raycast[4 * ((size_t)x + (size_t)y * p.cols)], vt, x, y, vba.data(), table.data(), invM, inv_proj,
                            one_over_voxel, p.mu, &minmax[2 * (size_t)id2], g);
            }
    }

    // CreateICPMaps_common, VisualisationEngine_CUDA.cu:324-360
    void icp_maps(const Pose& pose_c2w, float* points, float* normals) {
        raycast_pass(pose_c2w, true);
        float light[3] = {-pose_c2w.m[2], -pose_c2w.m[6], -pose_c2w.m[10]};  // -(column 2 of invM)
#pragma omp parallel for
        for (int y = 0; y < p.rows; ++y)
            for (int x = 0; x < p.cols; ++x)
                k::icp_map_pixel(points, normals, raycast.data(), p.cols, p.rows, x, y, p.voxel_size, light);
        if (p.corrected_mode) {
            // opt-in fix for SURVEY.md F1: express the model maps in the camera frame they were cast from
            Pose inv = pose_inv(pose_c2w);
            for (size_t i = 0; i < (size_t)p.cols * p.rows; ++i) {
                float* q = points + 4 * i; float* n = normals + 4 * i;
                if (std::isnan(q[0])) continue;
                float a[3] = {q[0], q[1], q[2]}, b[3] = {n[0], n[1], n[2]};
                for (int r = 0; r < 3; ++r) {
                    q[r] = inv.m[r * 4 + 0] * a[0] + inv.m[r * 4 + 1] * a[1] + inv.m[r * 4 + 2] * a[2] + inv.m[r * 4 + 3];
                    n[r] = inv.m[r * 4 + 0] * b[0] + inv.m[r * 4 + 1] * b[1] + inv.m[r * 4 + 2] * b[2];
                }
            }
        }
    }

    // TopFu::renderImage -> RenderImage_common(RENDER_SHADED_GREYSCALE, RENDER_FROM_NEW_RAYCAST)
    // (topfu.cpp:332-356, VisualisationEngine_CUDA.cu:220-291): raycast without touching visibility, then renderGrey_device
    void render_image(const Pose& pose_c2w, uint8_t* out_rgba) {
        raycast_pass(pose_c2w, false);
        float light[3] = {-pose_c2w.m[2], -pose_c2w.m[6], -pose_c2w.m[10]};
#pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < p.rows; ++y)
            for (int x = 0; x < p.cols; ++x) {
                size_t id = (size_t)x + (size_t)y * p.cols;
                k::shade_pixel_grey(out_rgba + 4 * id, &raycast[4 * id], vba.data(), table.data(), light, g);
            }
    }

    // renderPointCloud_device (include/tfusion/cuda/VisualisationHelper.hpp:150-198; dormant in the reference, its caller
    // take_cloud is a stub): per pixel of a raycast that leaves visibility alone — foundPoint = w > 0 (:164-166);
    // computeNormalAndAngle clears it when the SDF-gradient normal faces away from the light (:168, the same call
    // processPixelGrey makes: a shaded value of 0 <=> foundPoint false, drawPixelGrey writes >= 51 otherwise); skipPoints keeps
    // pixels with odd x and odd y (:170); locations = point * voxelSize, w = 1 (:189-192).  Raster order here; the reference's
    // order is the arrival order of its CTAs.  Returns the number of points.
    int render_point_cloud(const Pose& pose_c2w, bool skip_points, float* out_xyzw) {
        raycast_pass(pose_c2w, false);
        float light[3] = {-pose_c2w.m[2], -pose_c2w.m[6], -pose_c2w.m[10]};
        int n = 0;
        for (int y = 0; y < p.rows; ++y)
            for (int x = 0; x < p.cols; ++x) {
                size_t id = (size_t)x + (size_t)y * p.cols;
                uint8_t px[4];
                k::shade_pixel_grey(px, &raycast[4 * id], vba.data(), table.data(), light, g);
                bool found = px[0] != 0;
                if (skip_points && ((x % 2 == 0) || (y % 2 == 0))) found = false;
                if (!found) continue;
                for (int c = 0; c < 3; ++c) out_xyzw[4 * n + c] = raycast[4 * id + c] * p.voxel_size;
                out_xyzw[4 * n + 3] = 1.0f;
                ++n;
            }
        return n;
    }

    // ProjectiveICP::estimateTransform (points variant), projective_icp.cpp:169-212
    bool estimate_transform(Pose& affine) {
        IcpSetup s;
        s.min_cosine = cosf(p.icp_angle_thres);
        s.dist2_thres = p.icp_dist_thres * p.icp_dist_thres;
        affine = Pose::identity();
        for (int l = levels() - 1; l >= 0; --l) {
            int div = 1 << l;
            s.fx = p.fx / div; s.fy = p.fy / div; s.cx = p.cx / div; s.cy = p.cy / div;
            s.w = prev[l].w; s.h = prev[l].h;
            for (int it = 0; it < p.icp_iters[l]; ++it) {
                float v27[27];
                icp_reduce(s, affine, curr[l].points.data(), curr[l].normals.data(), prev[l].points.data(),
                           prev[l].normals.data(), v27, &icp_corresp_last);
                if (!icp_solve_update(v27, affine)) return false;
            }
        }
        return true;
    }

    void preprocess(const uint16_t* depth) {
        int L = levels();
        compute_dists(depth, dists.data(), p.cols, p.rows, p.depth_cutoff_mm);
        bilateral(depth, curr[0].depth.data(), p.cols, p.rows, p.bilateral_kernel_size, p.bilateral_sigma_spatial,
                  p.bilateral_sigma_depth);
        if (p.icp_truncate_depth_dist > 0) truncate_depth(curr[0].depth.data(), p.cols, p.rows, p.icp_truncate_depth_dist);
        for (int i = 1; i < L; ++i)
            depth_pyr(curr[i - 1].depth.data(), curr[i].depth.data(), curr[i - 1].w, curr[i - 1].h, p.bilateral_sigma_depth);
        for (int i = 0; i < L; ++i) {
            int div = 1 << i;  // Intr::operator()(level), precomp.cpp:10-14
            points_normals(curr[i].depth.data(), curr[i].points.data(), curr[i].normals.data(), curr[i].w, curr[i].h,
                           p.fx / div, p.fy / div, p.cx / div, p.cy / div);
        }
    }

    // TopFu::operator(), topfu.cpp:161-330 (debug downloads / renders omitted)
    bool process(const uint16_t* depth) {
        int L = levels();
        preprocess(depth);
        if (frame_counter == 0) {
            allocate(poses.back(), dists.data());
            integrate(poses.back(), dists.data());
            for (int i = 0; i < 4; ++i) { curr[i].points.swap(prev[i].points); curr[i].normals.swap(prev[i].normals); }
            ++frame_counter;
            return true;
        }
        Pose affine;
        bool ok = estimate_transform(affine);
        poses.push_back(pose_mul(poses.back(), affine));
        if (!ok) { reset(); return false; }
        Pose pose = poses.back();
        Pose inv = pose_inv(pose);
        allocate(inv, dists.data());
        integrate(inv, dists.data());
        expected_depths(inv);
        icp_maps(pose, prev[0].points.data(), prev[0].normals.data());
        for (int i = 1; i < L; ++i)
            resize_points_normals(prev[i - 1].points.data(), prev[i - 1].normals.data(), prev[i].points.data(),
                                  prev[i].normals.data(), prev[i - 1].w, prev[i - 1].h);
        ++frame_counter;
        return true;
    }
};

}  // namespace tfo

// ---------------------------------------------------------------------------------------
// C API (ctypes).  Poses cross this boundary as row-major float[16].
// ---------------------------------------------------------------------------------------
using namespace tfo;

extern "C" {

const char* tfo_impl_name() { return k::impl_name(); }

// TopFuParams::default_params, topfu.cpp:12-53; hash geometry VoxelBlockHash.hpp:10-18
void tfo_default_params(tfo_params* p) {
    memset(p, 0, sizeof(*p));
    p->cols = 640; p->rows = 480;
    p->fx = 504.261f; p->fy = 503.905f; p->cx = 352.457f; p->cy = 272.202f;
    p->bilateral_sigma_depth = 0.04f; p->bilateral_sigma_spatial = 4.5f; p->bilateral_kernel_size = 7;
    p->icp_truncate_depth_dist = 2.0f; p->icp_dist_thres = 0.1f; p->icp_angle_thres = 30.f * 0.017453293f;
    p->icp_iters[0] = 10; p->icp_iters[1] = 5; p->icp_iters[2] = 4; p->icp_iters[3] = 0;
    p->mu = 0.02f; p->max_w = 100; p->voxel_size = 0.005f; p->view_frustum_min = 0.2f; p->view_frustum_max = 3.0f;
    p->stop_integrating_at_max_w = 0;
    p->num_blocks = 0x10000; p->num_buckets = 0x100000; p->excess_size = 0x20000;
    p->depth_cutoff_mm = 2047; p->corrected_mode = 0; p->shard_rank = 0; p->shard_count = 1;
}

void* tfo_create(const tfo_params* p) { return new Oracle(*p); }
void tfo_destroy(void* h) { delete (Oracle*)h; }
void tfo_reset(void* h) { ((Oracle*)h)->reset(); }

// stateless image stages
void tfo_compute_dists(const uint16_t* depth, float* dists, int w, int h, int cutoff) { compute_dists(depth, dists, w, h, cutoff); }
void tfo_bilateral(const uint16_t* src, uint16_t* dst, int w, int h, int ksz, float ss, float sd) { bilateral(src, dst, w, h, ksz, ss, sd); }
void tfo_truncate_depth(uint16_t* d, int w, int h, float m) { truncate_depth(d, w, h, m); }
void tfo_depth_pyr(const uint16_t* src, uint16_t* dst, int sw, int sh, float sd) { depth_pyr(src, dst, sw, sh, sd); }
void tfo_points_normals(const uint16_t* d, float* pts, float* nrm, int w, int h, float fx, float fy, float cx, float cy) { points_normals(d, pts, nrm, w, h, fx, fy, cx, cy); }
void tfo_resize_points_normals(const float* v, const float* n, float* vo, float* no, int sw, int sh) { resize_points_normals(v, n, vo, no, sw, sh); }

// ICP pieces
void tfo_icp_reduce(int w, int h, float fx, float fy, float cx, float cy, float dist_thres, float angle_thres, const float* aff16,
                    const float* vcurr, const float* ncurr, const float* vprev, const float* nprev, float* out27, int* ncorr) {
    IcpSetup s; s.w = w; s.h = h; s.fx = fx; s.fy = fy; s.cx = cx; s.cy = cy;
    s.min_cosine = cosf(angle_thres); s.dist2_thres = dist_thres * dist_thres;
    Pose a; memcpy(a.m, aff16, sizeof(a.m));
    icp_reduce(s, a, vcurr, ncurr, vprev, nprev, out27, ncorr);
}
int tfo_icp_solve_update(const float* v27, float* aff16) {
    Pose a; memcpy(a.m, aff16, sizeof(a.m));
    bool ok = icp_solve_update(v27, a);
    memcpy(aff16, a.m, sizeof(a.m));
    return ok ? 1 : 0;
}
void tfo_solve6(const float* A, const float* b, float* x) { solve6_svd(A, b, x); }
double tfo_det6(const float* A) { return det6_f32(A); }
void tfo_rodrigues(const float* rv, const float* t, float* out16) { Pose p = pose_from_rvec_t(rv, t); memcpy(out16, p.m, sizeof(p.m)); }
void tfo_pose_inv(const float* in16, float* out16) { Pose a; memcpy(a.m, in16, 64); Pose b = pose_inv(a); memcpy(out16, b.m, 64); }
void tfo_pose_mul(const float* a16, const float* b16, float* out16) { Pose a, b; memcpy(a.m, a16, 64); memcpy(b.m, b16, 64); Pose c = pose_mul(a, b); memcpy(out16, c.m, 64); }
int tfo_mat4_inv_colmajor(const float* in16, float* out16) { return k::mat4_inv(in16, out16) ? 1 : 0; }

// scene stages with injected poses (row-major)
void tfo_allocate(void* h, const float* pose_w2c, const float* dists) { Pose p; memcpy(p.m, pose_w2c, 64); ((Oracle*)h)->allocate(p, dists); }
void tfo_integrate(void* h, const float* pose_w2c, const float* dists) { Pose p; memcpy(p.m, pose_w2c, 64); ((Oracle*)h)->integrate(p, dists); }
void tfo_expected_depths(void* h, const float* pose_w2c) { Pose p; memcpy(p.m, pose_w2c, 64); ((Oracle*)h)->expected_depths(p); }
void tfo_icp_maps(void* h, const float* pose_c2w, float* points, float* normals) { Pose p; memcpy(p.m, pose_c2w, 64); ((Oracle*)h)->icp_maps(p, points, normals); }
int tfo_render_point_cloud(void* h, const float* pose_c2w, int skip_points, float* out_xyzw) { Pose p; memcpy(p.m, pose_c2w, 64); return ((Oracle*)h)->render_point_cloud(p, skip_points != 0, out_xyzw); }
void tfo_render_image(void* h, const float* pose_c2w, uint8_t* out_rgba) { Pose p; memcpy(p.m, pose_c2w, 64); ((Oracle*)h)->render_image(p, out_rgba); }
void tfo_raycast(void* h, const float* pose_c2w, int update_visible) { Pose p; memcpy(p.m, pose_c2w, 64); ((Oracle*)h)->raycast_pass(p, update_visible != 0); }

// full frame
int tfo_process_frame(void* h, const uint16_t* depth) { return ((Oracle*)h)->process(depth) ? 1 : 0; }
int tfo_num_poses(void* h) { return (int)((Oracle*)h)->poses.size(); }
void tfo_get_pose(void* h, int idx, float* out16) {
    Oracle* o = (Oracle*)h;
    if (idx < 0 || idx >= (int)o->poses.size()) idx = (int)o->poses.size() - 1;
    memcpy(out16, o->poses[idx].m, 64);
}

// state export.  counters: [n_visible, last_free_block, last_free_excess, n_tiles, frame_counter, resets,
//                           icp_corresp_last, n_allocated]
void tfo_get_counters(void* h, long long* out8) {
    Oracle* o = (Oracle*)h;
    out8[0] = o->n_visible; out8[1] = o->last_free_block; out8[2] = o->last_free_excess; out8[3] = o->n_tiles;
    out8[4] = o->frame_counter; out8[5] = o->resets; out8[6] = o->icp_corresp_last;
    out8[7] = (long long)(o->p.num_blocks - 1 - o->last_free_block);
}
long long tfo_voxel_updates(void* h) { return ((Oracle*)h)->voxel_updates; }
int tfo_total_entries(void* h) { return ((Oracle*)h)->g.total_entries(); }
void tfo_export_table(void* h, void* out) { Oracle* o = (Oracle*)h; memcpy(out, o->table.data(), o->table.size() * sizeof(HashEntry)); }
void tfo_export_vis_type(void* h, uint8_t* out) { Oracle* o = (Oracle*)h; memcpy(out, o->vis_type.data(), o->vis_type.size()); }
void tfo_export_visible_ids(void* h, int* out) { Oracle* o = (Oracle*)h; memcpy(out, o->visible_ids.data(), sizeof(int) * o->n_visible); }
// one block's 512 voxels (int16 sdf, u8 w, u8 pad) by pool pointer
void tfo_export_block(void* h, int ptr, void* out) { Oracle* o = (Oracle*)h; memcpy(out, &o->vba[(size_t)ptr * BLOCK3], BLOCK3 * sizeof(Voxel)); }
void tfo_export_minmax(void* h, float* out) { Oracle* o = (Oracle*)h; memcpy(out, o->minmax.data(), o->minmax.size() * sizeof(float)); }
void tfo_export_raycast(void* h, float* out) { Oracle* o = (Oracle*)h; memcpy(out, o->raycast.data(), o->raycast.size() * sizeof(float)); }
void tfo_export_dists(void* h, float* out) { Oracle* o = (Oracle*)h; memcpy(out, o->dists.data(), o->dists.size() * sizeof(float)); }
// which: 0 curr depth (u16), 1 curr points, 2 curr normals, 3 prev points, 4 prev normals
void tfo_export_level(void* h, int which, int level, void* out) {
    Oracle* o = (Oracle*)h;
    Level& c = o->curr[level]; Level& pv = o->prev[level];
    switch (which) {
        case 0: memcpy(out, c.depth.data(), c.depth.size() * 2); break;
        case 1: memcpy(out, c.points.data(), c.points.size() * 4); break;
        case 2: memcpy(out, c.normals.data(), c.normals.size() * 4); break;
        case 3: memcpy(out, pv.points.data(), pv.points.size() * 4); break;
        case 4: memcpy(out, pv.normals.data(), pv.normals.size() * 4); break;
    }
}
// test hook: install model maps for ICP-only tests
void tfo_import_level(void* h, int which, int level, const void* in) {
    Oracle* o = (Oracle*)h;
    Level& c = o->curr[level]; Level& pv = o->prev[level];
    switch (which) {
        case 0: memcpy(c.depth.data(), in, c.depth.size() * 2); break;
        case 1: memcpy(c.points.data(), in, c.points.size() * 4); break;
        case 2: memcpy(c.normals.data(), in, c.normals.size() * 4); break;
        case 3: memcpy(pv.points.data(), in, pv.points.size() * 4); break;
        case 4: memcpy(pv.normals.data(), in, pv.normals.size() * 4); break;
    }
}
int tfo_estimate_transform(void* h, float* aff16) {
    Pose a; bool ok = ((Oracle*)h)->estimate_transform(a); memcpy(aff16, a.m, 64); return ok ? 1 : 0;
}
void tfo_preprocess(void* h, const uint16_t* depth) { ((Oracle*)h)->preprocess(depth); }

}  // extern "C"
