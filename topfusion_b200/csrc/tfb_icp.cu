// Projective point-to-plane ICP.  Replaces ComputeIcpHelper (icp_helper_kernel + icp_final_reduce_kernel,
// /root/reference/tfusion/src/cuda/proj_icp.cu:80-117,359-455) and the host half of
// ProjectiveICP::estimateTransform (src/projective_icp.cpp:43-62,169-212: stream sync, unpack, OpenCV
// determinant / SVD solve / Rodrigues, pose update).
//
// One launch per iteration, no host round trip: each thread accumulates the 27 unique A^T A / A^T b
// products of its pixels in fp32 registers, warps combine them with shuffles, the CTA writes one partial
// row, and the last CTA to finish (atomic ticket) folds the partials in a fixed order in fp64, solves the
// 6x6 system in fp64 and left-multiplies the increment onto the running transform kept in device memory.
// The reference spends 27 x (store, barrier, 8-step shared-memory tree, barrier) per CTA plus a second
// kernel, a 108-byte D2H copy and a stream sync per iteration (19 per frame).
#include "tfb_common.cuh"

namespace tfb {

struct IcpArgs {
    const float4* vcurr;
    const float4* ncurr;
    const float4* vprev;
    const float4* nprev;
    int w, h;
    float fx, fy, cx, cy;
    float min_cosine, dist2_thres;
    int rows_per_thread;
    int nblk;
};

constexpr int ICP_TX = 32, ICP_TY = 8, ICP_THREADS = ICP_TX * ICP_TY, ICP_WARPS = ICP_THREADS / 32;
constexpr int ICP_ACC = ICP_TERMS + 1;  // + correspondence count

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    // temp_utils.hpp:27-30
    return __fmaf_rn(ax, bx, __fmaf_rn(ay, by, az * bz));
}

// ---- small dense algebra run by one thread of the last CTA ----------------------------------

// cv::determinant(Matx66f): LU with partial pivoting in fp32, product in fp64 (projective_icp.cpp:197)
__device__ double det6_f32(const float* A) {
    float a[36];
    for (int i = 0; i < 36; ++i) a[i] = A[i];
    double p = 1;
    for (int i = 0; i < 6; ++i) {
        int k = i;
        for (int j = i + 1; j < 6; ++j)
            if (fabsf(a[j * 6 + i]) > fabsf(a[k * 6 + i])) k = j;
        if (fabsf(a[k * 6 + i]) < 1.1920929e-06f) return 0;
        if (k != i) {
            for (int j = i; j < 6; ++j) { float t = a[i * 6 + j]; a[i * 6 + j] = a[k * 6 + j]; a[k * 6 + j] = t; }
            p = -p;
        }
        float d = __fdiv_rn(-1.f, a[i * 6 + i]);
        for (int j = i + 1; j < 6; ++j) {
            float alpha = __fmul_rn(a[j * 6 + i], d);
            for (int cidx = i + 1; cidx < 6; ++cidx) a[j * 6 + cidx] = __fadd_rn(a[j * 6 + cidx], __fmul_rn(alpha, a[i * 6 + cidx]));
        }
    }
    for (int i = 0; i < 6; ++i) p *= a[i * 6 + i];
    return p;
}

// least-norm solve of the symmetric system through a cyclic-Jacobi eigen-decomposition (fp64) with the
// back-substitution threshold of cv::solve(DECOMP_SVD) (2*FLT_EPSILON*sum|w|).  Only used when the
// LDL^T fast path meets a tiny pivot, i.e. for (near) rank-deficient systems.
__device__ __noinline__ void solve6_jacobi(const double* A, const double* b, double* x) {
    double a[6][6], v[6][6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) { a[i][j] = A[i * 6 + j]; v[i][j] = (i == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int i = 0; i < 6; ++i)
            for (int j = i + 1; j < 6; ++j) off += a[i][j] * a[i][j];
        if (off < 1e-300) break;
        for (int p = 0; p < 5; ++p)
            for (int q = p + 1; q < 6; ++q) {
                if (a[p][q] == 0) continue;
                double th = (a[q][q] - a[p][p]) / (2 * a[p][q]);
                double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
                double cs = 1 / sqrt(t * t + 1), sn = t * cs;
                for (int k = 0; k < 6; ++k) { double akp = a[k][p], akq = a[k][q]; a[k][p] = cs * akp - sn * akq; a[k][q] = sn * akp + cs * akq; }
                for (int k = 0; k < 6; ++k) { double apk = a[p][k], aqk = a[q][k]; a[p][k] = cs * apk - sn * aqk; a[q][k] = sn * apk + cs * aqk; }
                for (int k = 0; k < 6; ++k) { double vkp = v[k][p], vkq = v[k][q]; v[k][p] = cs * vkp - sn * vkq; v[k][q] = sn * vkp + cs * vkq; }
            }
    }
    double wsum = 0;
    for (int i = 0; i < 6; ++i) wsum += fabs(a[i][i]);
    double thr = wsum * 2 * 1.1920928955078125e-07;
    for (int k = 0; k < 6; ++k) x[k] = 0;
    for (int i = 0; i < 6; ++i) {
        double w = a[i][i];
        if (fabs(w) <= thr) continue;
        double s = 0;
        for (int k = 0; k < 6; ++k) s += v[k][i] * b[k];
        s /= w;
        for (int k = 0; k < 6; ++k) x[k] += s * v[k][i];
    }
}

// LDL^T in fp64; false when a pivot is too small relative to the largest diagonal entry
__device__ bool solve6_ldlt(const double* A, const double* b, double* x) {
    double L[6][6], d[6];
    double amax = 0;
    for (int i = 0; i < 6; ++i) amax = fmax(amax, fabs(A[i * 6 + i]));
    const double tiny = amax * 4e-6;
    for (int j = 0; j < 6; ++j) {
        double dj = A[j * 6 + j];
        for (int k = 0; k < j; ++k) dj -= L[j][k] * L[j][k] * d[k];
        if (!(dj > tiny)) return false;
        d[j] = dj;
        for (int i = j + 1; i < 6; ++i) {
            double s = A[i * 6 + j];
            for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k] * d[k];
            L[i][j] = s / dj;
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
        y[i] = s;
    }
    for (int i = 5; i >= 0; --i) {
        double s = y[i] / d[i];
        for (int k = i + 1; k < 6; ++k) s -= L[k][i] * x[k];
        x[i] = s;
    }
    return true;
}

// StreamHelper::get unpack (projective_icp.cpp:43-62), nullspace test (:197-203), solve (:206),
// Tinc = Affine3f(rvec, t) and affine = Tinc * affine (:208-209)
__device__ void icp_solve_update(const double* v27, DevState* ds) {
    float Af[36], bf[6];
    int shift = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 7; ++j) {
            float value = (float)v27[shift++];
            if (j == 6) bf[i] = value;
            else Af[j * 6 + i] = Af[i * 6 + j] = value;
        }
    double det = det6_f32(Af);
    if (fabs(det) < 1e-15 || det != det) {
        ds->icp_failed = 1;
        return;
    }
    double A[36], b[6], r[6];
    for (int i = 0; i < 36; ++i) A[i] = Af[i];
    for (int i = 0; i < 6; ++i) b[i] = bf[i];
    if (!solve6_ldlt(A, b, r)) solve6_jacobi(A, b, r);
    float rf[6];
    for (int i = 0; i < 6; ++i) rf[i] = (float)r[i];

    // cv::Affine3f(rvec, t): Rodrigues evaluated in double, stored as float
    float T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    double theta = sqrt((double)rf[0] * rf[0] + (double)rf[1] * rf[1] + (double)rf[2] * rf[2]);
    if (theta >= 2.220446049250313e-16) {
        double sn, cs;
        sincos(theta, &sn, &cs);
        double c1 = 1. - cs, it = 1. / theta;
        float rx = (float)(rf[0] * it), ry = (float)(rf[1] * it), rz = (float)(rf[2] * it);
        double rrt[9] = {(double)rx * rx, (double)rx * ry, (double)rx * rz, (double)rx * ry, (double)ry * ry,
                         (double)ry * rz, (double)rx * rz, (double)ry * rz, (double)rz * rz};
        double rc[9] = {0, -(double)rz, (double)ry, (double)rz, 0, -(double)rx, -(double)ry, (double)rx, 0};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                int k = i * 3 + j;
                T[i * 4 + j] = (float)(cs * (i == j ? 1.0 : 0.0) + c1 * rrt[k] + sn * rc[k]);
            }
    }
    T[3] = rf[3]; T[7] = rf[4]; T[11] = rf[5];

    float old[16], nw[16];
    for (int i = 0; i < 16; ++i) old[i] = ds->affine[i];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0;
            for (int k = 0; k < 4; ++k) s = __fadd_rn(s, __fmul_rn(T[i * 4 + k], old[k * 4 + j]));
            nw[i * 4 + j] = s;
        }
    for (int i = 0; i < 16; ++i) ds->affine[i] = nw[i];
}

// ---- the iteration kernel ---------------------------------------------------------------------
__global__ void __launch_bounds__(ICP_THREADS)
    k_icp_iteration(IcpArgs a, DevState* __restrict__ ds, float* __restrict__ partial, int solve, float* __restrict__ out27) {
    if (ds->icp_failed) return;  // estimateTransform returned false earlier in this frame

    __shared__ float s_warp[ICP_WARPS][ICP_ACC];
    __shared__ double s_part[ICP_WARPS][32];
    __shared__ double s_tot[ICP_ACC];
    __shared__ int s_last;

    const int tid = threadIdx.y * ICP_TX + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    // aff = device_cast<Aff3f>(affine): rows of R and t (projective_icp.cpp:190)
    const float* af = ds->affine;
    const float r00 = af[0], r01 = af[1], r02 = af[2], t0 = af[3];
    const float r10 = af[4], r11 = af[5], r12 = af[6], t1 = af[7];
    const float r20 = af[8], r21 = af[9], r22 = af[10], t2 = af[11];

    float acc[ICP_ACC];
#pragma unroll
    for (int i = 0; i < ICP_ACC; ++i) acc[i] = 0.f;

    const int x = blockIdx.x * ICP_TX + threadIdx.x;
    const int ybase = (blockIdx.y * ICP_TY + threadIdx.y) * a.rows_per_thread;
    for (int ry = 0; ry < a.rows_per_thread; ++ry) {
        const int y = ybase + ry;
        if (x >= a.w || y >= a.h) continue;
        // find_coresp, proj_icp.cu:80-117 (points variant)
        const float4 v = __ldg(a.vcurr + y * a.w + x);
        if (isnan(v.x)) continue;
        const float sx = dot3(r00, r01, r02, v.x, v.y, v.z) + t0;
        const float sy = dot3(r10, r11, r12, v.x, v.y, v.z) + t1;
        const float sz = dot3(r20, r21, r22, v.x, v.y, v.z) + t2;
        // IEEE division where the reference uses __fdividef (proj_icp.cu:33-34): at the identity transform every
        // point projects exactly onto a pixel centre, so floor(coo) would otherwise hinge on the approximation error
        const float cox = __fmaf_rn(a.fx, __fdiv_rn(sx, sz), a.cx);
        const float coy = __fmaf_rn(a.fy, __fdiv_rn(sy, sz), a.cy);
        if (sz <= 0 || cox < 0 || coy < 0 || cox >= a.w || coy >= a.h) continue;
        const int pidx = (int)coy * a.w + (int)cox;  // point-sampled texel (floor, coordinates are >= 0)
        const float4 d = __ldg(a.vprev + pidx);
        if (isnan(d.x)) continue;
        const float ex = sx - d.x, ey = sy - d.y, ez = sz - d.z;
        if (dot3(ex, ey, ez, ex, ey, ez) > a.dist2_thres) continue;
        const float4 nc = __ldg(a.ncurr + y * a.w + x);
        const float nsx = dot3(r00, r01, r02, nc.x, nc.y, nc.z);
        const float nsy = dot3(r10, r11, r12, nc.x, nc.y, nc.z);
        const float nsz = dot3(r20, r21, r22, nc.x, nc.y, nc.z);
        const float4 nd = __ldg(a.nprev + pidx);
        if (fabsf(dot3(nsx, nsy, nsz, nd.x, nd.y, nd.z)) < a.min_cosine) continue;
        // row = [s x n, n, n.(d - s)], icp_helper_kernel proj_icp.cu:369-376
        float row[7];
        row[0] = sy * nd.z - sz * nd.y;
        row[1] = sz * nd.x - sx * nd.z;
        row[2] = sx * nd.y - sy * nd.x;
        row[3] = nd.x; row[4] = nd.y; row[5] = nd.z;
        row[6] = dot3(nd.x, nd.y, nd.z, d.x - sx, d.y - sy, d.z - sz);
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 7; ++j) acc[k++] += row[i] * row[j];
        acc[ICP_TERMS] += 1.f;
    }

    // warp shuffle tree, then one row per warp in shared memory
#pragma unroll
    for (int i = 0; i < ICP_ACC; ++i) {
        float vsum = acc[i];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
        if (lane == 0) s_warp[warp][i] = vsum;
    }
    __syncthreads();
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid < ICP_ACC) {
        float s = 0.f;
#pragma unroll
        for (int wi = 0; wi < ICP_WARPS; ++wi) s += s_warp[wi][tid];
        partial[tid * a.nblk + blk] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int t = atomicAdd(&ds->icp_ticket, 1u);
        s_last = (t == (unsigned)a.nblk - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // last CTA: fold the per-CTA partials in a fixed order, fp64
    {
        const int k = lane, part = warp;
        double s = 0;
        if (k < ICP_ACC)
            for (int i = part; i < a.nblk; i += ICP_WARPS) s += (double)__ldcg(partial + k * a.nblk + i);
        s_part[part][k] = s;
    }
    __syncthreads();
    if (tid < ICP_ACC) {
        double s = 0;
#pragma unroll
        for (int p = 0; p < ICP_WARPS; ++p) s += s_part[p][tid];
        s_tot[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        ds->icp_ticket = 0;
        ds->icp_corresp = (int)s_tot[ICP_TERMS];
        if (out27)
            for (int i = 0; i < ICP_TERMS; ++i) out27[i] = (float)s_tot[i];
        if (solve) icp_solve_update(s_tot, ds);
    }
}

__global__ void k_icp_begin(DevState* ds) {
    // estimateTransform starts from identity (projective_icp.cpp:174)
    if (threadIdx.x < 16) ds->affine[threadIdx.x] = ((threadIdx.x % 5) == 0) ? 1.f : 0.f;
    if (threadIdx.x == 0) { ds->icp_failed = 0; ds->icp_ticket = 0; ds->icp_corresp = 0; }
}

int launch_icp_begin(tfb_ctx* c) {
    TFB_KT(c, K_ICP_BEGIN);
    k_icp_begin<<<1, 32, 0, c->stream>>>(c->ds);
    TFB_LAUNCH_CHECK(c);
    return TFB_OK;
}

int launch_icp_iteration(tfb_ctx* c, int level, const float4* vcurr, const float4* ncurr, const float4* vprev, const float4* nprev,
                         int w, int h, float fx, float fy, float cx, float cy, bool solve, float* out27_dev) {
    IcpArgs a;
    a.vcurr = vcurr; a.ncurr = ncurr; a.vprev = vprev; a.nprev = nprev;
    a.w = w; a.h = h; a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy;
    a.min_cosine = cosf(c->p.icp_angle_thres);                 // ComputeIcpHelper ctor, projective_icp.cpp:11-15
    a.dist2_thres = c->p.icp_dist_thres * c->p.icp_dist_thres;
    // enough CTAs to cover the SMs at the coarse levels, four rows per thread at full resolution
    a.rows_per_thread = ((long long)w * h >= 200000) ? 4 : 1;
    dim3 block(ICP_TX, ICP_TY), grid(div_up(w, ICP_TX), div_up(h, ICP_TY * a.rows_per_thread));
    a.nblk = grid.x * grid.y;
    if (a.nblk > c->icp_max_blocks) return set_err(c, TFB_ERR_ARG, "icp: image larger than the partial buffer");
    TFB_KT(c, K_ICP_L0 + level);
    k_icp_iteration<<<grid, block, 0, c->stream>>>(a, c->ds, c->icp_partial, solve ? 1 : 0, out27_dev);
    TFB_LAUNCH_CHECK(c);
    return TFB_OK;
}

}  // namespace tfb
