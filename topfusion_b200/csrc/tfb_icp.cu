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
#include "tfb_pose.cuh"

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
    int first_iter;   // estimateTransform starts from the identity (projective_icp.cpp:174)
    int last_iter;    // also do poses_.push_back(poses_.back() * affine) and derive the frame's matrices (topfu.cpp:243)
};

constexpr int ICP_TX = 32, ICP_TY = 8, ICP_THREADS = ICP_TX * ICP_TY, ICP_WARPS = ICP_THREADS / 32;
constexpr int ICP_ACC = ICP_TERMS + 1;  // + correspondence count

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    // temp_utils.hpp:27-30
    return __fmaf_rn(ax, bx, __fmaf_rn(ay, by, az * bz));
}

// ---- small dense algebra run by one thread of the last CTA ----------------------------------
// Everything below is straight-line code over register arrays (all loops unrolled, compile-time indices).

// LU with partial pivoting in fp32 — the algorithm behind cv::determinant(Matx66f) (projective_icp.cpp:197).
// The factors are kept: they also precondition the solve below.  a = U above/on the diagonal, the multipliers
// (alpha, already negated as in OpenCV's LUImpl) below it; perm[i] = source row of pivot row i.
struct Lu6 {
    float a[6][6];
    int perm[6];
    double det;       // sign * prod(diag), 0 when a pivot falls below 10*FLT_EPSILON
};

__device__ __forceinline__ void lu6_f32(const float (&A)[36], Lu6& f) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        f.perm[i] = i;
#pragma unroll
        for (int j = 0; j < 6; ++j) f.a[i][j] = A[i * 6 + j];
    }
    double p = 1;
    bool singular = false;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        int k = i;
        float best = fabsf(f.a[i][i]);
#pragma unroll
        for (int j = i + 1; j < 6; ++j) {
            float v = fabsf(f.a[j][i]);
            if (v > best) { best = v; k = j; }
        }
        if (best < 1.1920929e-06f) singular = true;   // FLT_EPSILON * 10
#pragma unroll
        for (int j = i + 1; j < 6; ++j)
            if (k == j) {
#pragma unroll
                for (int c = 0; c < 6; ++c) { float t = f.a[i][c]; f.a[i][c] = f.a[j][c]; f.a[j][c] = t; }
                int tp = f.perm[i]; f.perm[i] = f.perm[j]; f.perm[j] = tp;
                p = -p;
            }
        const float d = __fdiv_rn(-1.f, f.a[i][i]);
#pragma unroll
        for (int j = i + 1; j < 6; ++j) {
            const float alpha = f.a[j][i] * d;
#pragma unroll
            for (int c = i + 1; c < 6; ++c) f.a[j][c] = f.a[j][c] + alpha * f.a[i][c];
            f.a[j][i] = alpha;   // keep the (negated) multiplier
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) p *= f.a[i][i];
    f.det = singular ? 0.0 : p;
}

// least-norm solve of the symmetric system through a cyclic-Jacobi eigen-decomposition (fp64) with the
// back-substitution threshold of cv::solve(DECOMP_SVD) (2*FLT_EPSILON*sum|w|).  Only reached when the
// LDL^T fast path meets a tiny pivot, i.e. for (near) rank-deficient systems; kept out of line.
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

// Fast path: LDL^T in fp32 (no pivoting, A is symmetric positive definite when tracking is healthy) used as a
// preconditioner, solution refined with fp64 residuals: error shrinks by ~cond(A)*6e-8 per step, three solves
// give fp64-grade accuracy for cond(A) up to ~1e5.  A pivot below 4e-6 * max(diag) or a refinement that does not
// contract sends the system to the exact reference path (pivoted-LU determinant test + least-norm eigen solve).
struct Ldl6 {
    float l[6][6];
    float dinv[6];
    float d[6];       // the pivots; their product is the determinant
    float dmin;
    bool ok;
};

__device__ __forceinline__ void ldl6_f32(const float (&A)[36], Ldl6& f) {
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) amax = fmaxf(amax, fabsf(A[i * 6 + i]));
    const float tiny = amax * 4e-6f;
    float d[6];
    f.ok = true;
    f.dmin = 3.0e38f;
    // this runs on ONE thread, 19 times per frame, on the critical path of every iteration: fused multiply-adds, the
    // products l[j][k] * d[k] shared by the column below, and a Newton-refined MUFU reciprocal instead of a division
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        float ld[6];
        float dj = A[j * 6 + j];
#pragma unroll
        for (int k = 0; k < j; ++k) { ld[k] = f.l[j][k] * d[k]; dj = __fmaf_rn(-f.l[j][k], ld[k], dj); }
        if (!(dj > tiny)) f.ok = false;
        d[j] = dj;
        f.d[j] = dj;
        f.dmin = fminf(f.dmin, dj);
        float y0;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(dj));
        f.dinv[j] = __fmaf_rn(y0, __fmaf_rn(-dj, y0, 1.0f), y0);
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            float sacc = A[i * 6 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) sacc = __fmaf_rn(-f.l[i][k], ld[k], sacc);
            f.l[i][j] = sacc * f.dinv[j];
        }
    }
}

template <typename T>
__device__ __forceinline__ void ldl6_solve(const Ldl6& f, const T (&r)[6], float (&x)[6]) {
    float y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        float v = (float)r[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v = __fmaf_rn(-f.l[i][k], y[k], v);
        y[i] = v;
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        float v = y[i] * f.dinv[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) v = __fmaf_rn(-f.l[k][i], x[k], v);
        x[i] = v;
    }
}

// One solve plus one step of fixed-precision refinement, everything in fp32: the accuracy class of the reference's own
// fp32 Jacobi SVD (cv::solve(DECOMP_SVD) on Matx66f), error ~ cond(A) * 6e-8 of an increment that itself shrinks from
// iteration to iteration.  (An earlier version refined with fp64 residuals; it was more exact than the reference and cost
// 2.2 us of single-thread latency in each of the 19 iterations of a frame.)  Returns false when the correction does not
// contract, which sends the system to the reference path below.
__device__ __forceinline__ bool solve6_refine(const float (&A)[36], const float (&b)[6], const Ldl6& f, float (&x)[6]) {
    ldl6_solve(f, b, x);
    float r[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        float sacc = b[i];
#pragma unroll
        for (int j = 0; j < 6; ++j) sacc = __fmaf_rn(-A[i * 6 + j], x[j], sacc);
        r[i] = sacc;
    }
    float dx[6];
    ldl6_solve(f, r, dx);
    float nx = 0.f, nd = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) { nx = fmaxf(nx, fabsf(x[i])); nd = fmaxf(nd, fabsf(dx[i])); x[i] += dx[i]; }
    return nd <= 0.05f * nx || nx == 0.f;
}

// the reference's exact path, used when the fast path declines: cv::determinant's pivoted fp32 LU for the nullspace
// test, then the least-norm solve
__device__ __noinline__ bool solve6_reference_path(const float* Af_, const float* bf_, double* r_out) {
    float Af[36];
    for (int i = 0; i < 36; ++i) Af[i] = Af_[i];
    Lu6 lu;
    lu6_f32(Af, lu);
    const double det = lu.det;
    if (fabs(det) < 1e-15 || det != det) return false;
    double A[36], b[6];
    for (int i = 0; i < 36; ++i) A[i] = Af[i];
    for (int i = 0; i < 6; ++i) b[i] = bf_[i];
    solve6_jacobi(A, b, r_out);
    return true;
}

// Tinc = Affine3f(rvec, t) and affine = Tinc * affine (projective_icp.cpp:208-209); aff is the running estimate (row-major,
// rows 0..2 — row 3 is (0, 0, 0, 1) from the identity the loop starts with and every product keeps it, so neither operand's
// fourth row is multiplied out).
__device__ __forceinline__ void icp_apply_increment(const float (&rf)[6], float* aff_io) {
    // cv::Affine3f(rvec, t) = Rodrigues: R = cos I + (1 - cos) r r^T + sin [r]x.  OpenCV evaluates it
    // in double and stores float; the increments here are fractions of a degree, where fp32 series for sin, cos and
    // (1 - cos) — the latter summed directly, without the cancellation of 1 - cos — are exact to the last float bit or two.
    float T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    const float t2 = rf[0] * rf[0] + rf[1] * rf[1] + rf[2] * rf[2];
    if (t2 > 0.f) {
        const float theta = sqrtf(t2);
        float sn, cs, c1;
        if (theta < 0.5f) {
            sn = theta * (1.0f + t2 * (-1.0f / 6 + t2 * (1.0f / 120 + t2 * (-1.0f / 5040 + t2 * (1.0f / 362880)))));
            c1 = t2 * (0.5f + t2 * (-1.0f / 24 + t2 * (1.0f / 720 + t2 * (-1.0f / 40320 + t2 * (1.0f / 3628800)))));
            cs = 1.0f - c1;
        } else {
            sincosf(theta, &sn, &cs);
            c1 = 1.0f - cs;
        }
        const float it = 1.0f / theta;
        const float rx = rf[0] * it, ry = rf[1] * it, rz = rf[2] * it;
        T[0] = cs + c1 * rx * rx;      T[1] = c1 * rx * ry - sn * rz; T[2] = c1 * rx * rz + sn * ry;
        T[4] = c1 * rx * ry + sn * rz; T[5] = cs + c1 * ry * ry;      T[6] = c1 * ry * rz - sn * rx;
        T[8] = c1 * rx * rz - sn * ry; T[9] = c1 * ry * rz + sn * rx; T[10] = cs + c1 * rz * rz;
    }
    T[3] = rf[3]; T[7] = rf[4]; T[11] = rf[5];
    float a[12], nw[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = aff_io[i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v = T[i * 4] * a[c];
            v += T[i * 4 + 1] * a[4 + c];
            v += T[i * 4 + 2] * a[8 + c];
            if (c == 3) v += T[i * 4 + 3];
            nw[i * 4 + c] = v;
        }
#pragma unroll
    for (int i = 0; i < 12; ++i) aff_io[i] = nw[i];
}

// StreamHelper::get unpack (projective_icp.cpp:43-62): the 27 sums -> symmetric A (6x6) and b
__device__ __forceinline__ void icp_unpack(const float* v27, float (&Af)[36], float (&bf)[6]) {
    int shift = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 7; ++j) {
            const float value = v27[shift++];
            if (j == 6) bf[i] = value;
            else { Af[j * 6 + i] = value; Af[i * 6 + j] = value; }
        }
}

// the system the fast path declined: the reference's nullspace test and least-norm solve, out of line (its arrays live in
// local memory; the fast path's stay in registers because nothing takes their address)
__device__ __noinline__ bool icp_solve_slow(const float* v27, float* aff_io) {
    float Af[36], bf[6];
    icp_unpack(v27, Af, bf);
    double r[6];
    if (!solve6_reference_path(Af, bf, r)) return false;
    float rf[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) rf[i] = (float)r[i];
    icp_apply_increment(rf, aff_io);
    return true;
}

// Unpack, nullspace test (projective_icp.cpp:197-203), solve (:206), Tinc * affine (:208-209).  v27: the folded sums as fp32
// (shared memory), aff_io: the running estimate, updated in place on success.  Returns false when tracking failed.
// One thread, 19 times a frame, on the critical path of every iteration: straight-line register code.  The determinant of
// the nullspace test is the product of the six pivots; when the smallest pivot is above (1e-15)^(1/6) = 3.17e-3 — always, for
// sums over 10^5 pixels — the product need not be formed to know it passes (it was a chain of six fp64 multiplies).
__device__ __forceinline__ bool icp_solve_update(const float* v27, float* aff_io) {
    float Af[36], bf[6];
    icp_unpack(v27, Af, bf);
    float rf[6];
    Ldl6 f;
    ldl6_f32(Af, f);
    bool solved = false;
    if (f.ok) {
        bool det_ok = f.dmin >= 3.2e-3f;
        if (!det_ok) {
            double det = 1.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) det *= (double)f.d[i];
            det_ok = fabs(det) >= 1e-15;
        }
        if (det_ok) solved = solve6_refine(Af, bf, f, rf);
    }
    if (!solved) return icp_solve_slow(v27, aff_io);
    icp_apply_increment(rf, aff_io);
    return true;
}

// ---- the iteration kernel ---------------------------------------------------------------------
__global__ void __launch_bounds__(ICP_THREADS)
    k_icp_iteration(IcpArgs a, DevState* __restrict__ ds, float* __restrict__ partial, int solve, float* __restrict__ out27) {
    if (!a.first_iter && ds->icp_failed) return;  // estimateTransform returned false earlier in this frame

    __shared__ float s_warp[ICP_WARPS][ICP_ACC];
    __shared__ double s_part[ICP_WARPS][32];
    __shared__ double s_tot[ICP_ACC];
    __shared__ int s_last;

    const int tid = threadIdx.y * ICP_TX + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    // aff = device_cast<Aff3f>(affine): rows of R and t (projective_icp.cpp:190)
    float r00 = 1.f, r01 = 0.f, r02 = 0.f, t0 = 0.f, r10 = 0.f, r11 = 1.f, r12 = 0.f, t1 = 0.f, r20 = 0.f, r21 = 0.f, r22 = 1.f, t2 = 0.f;
    if (!a.first_iter) {
        const float* af = ds->affine;
        r00 = af[0]; r01 = af[1]; r02 = af[2]; t0 = af[3];
        r10 = af[4]; r11 = af[5]; r12 = af[6]; t1 = af[7];
        r20 = af[8]; r21 = af[9]; r22 = af[10]; t2 = af[11];
    }

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
        if (solve) {
            float aff[16] = {r00, r01, r02, t0, r10, r11, r12, t1, r20, r21, r22, t2, 0.f, 0.f, 0.f, 1.f};
            float v27[ICP_TERMS];
#pragma unroll
            for (int i = 0; i < ICP_TERMS; ++i) v27[i] = (float)s_tot[i];
            const bool ok = icp_solve_update(v27, aff);
            if (a.first_iter || !ok) ds->icp_failed = ok ? 0 : 1;
            if (ok) {
#pragma unroll
                for (int i = 0; i < 16; ++i) ds->affine[i] = aff[i];
                if (a.last_iter) {
                    float prev[16], nw[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) prev[i] = ds->pose_c2w[i];
                    pose_mul(prev, aff, nw);
                    store_pose_c2w(ds, nw);
                }
            }
        } else if (a.first_iter) {
            ds->icp_failed = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// The whole coarse-to-fine loop in ONE cooperative launch (ProjectiveICP::estimateTransform,
// projective_icp.cpp:169-212: 19 x {kernel, kernel, D2H copy, stream sync, host SVD} in the reference).
//
//   * persistent grid (one 512-thread CTA per SM, co-resident by cooperative launch), grid-stride over pixels with four
//     independent pixels in flight per thread so the dependent load chain v -> (d, n_d) overlaps across pixels;
//   * per iteration one CTA-level reduction (shuffles + shared memory), one row of partials per CTA, ONE grid
//     synchronisation point — the epoch-stamped rows themselves, polled by the fold; then every CTA folds all partials in the
//     same fixed order (fp64) and solves redundantly, which makes the result bit-identical in every CTA and saves the
//     second barrier a broadcast would need;
//   * the last iteration composes poses_.back() * affine and derives the frame's matrices (topfu.cpp:243,281).
// ---------------------------------------------------------------------------------------------------------
struct IcpLevelArgs {
    const float4* vcurr;
    const float4* ncurr;
    const float4* vprev;
    const float4* nprev;
    int w, h, iters;
    float fx, fy, cx, cy;
    const int* vlist;            // the level's pixels that have a vertex, ascending (tfb_imgproc.cu), or null: the CTA compacts its own slots
    const unsigned int* vlist_n;
};

struct IcpAllArgs {
    IcpLevelArgs lv[MAX_LEVELS];
    int levels;
    float min_cosine, dist2_thres;
    int update_pose;
    unsigned int epoch_base;   // this launch's range of row epochs (64 per launch)
};

constexpr int ICPA_THREADS = 512, ICPA_WARPS = ICPA_THREADS / 32, ICPA_UNROLL = 5;   // one CTA per SM
constexpr int ICP_LIST_SLOTS = 16;   // pixel slots per thread the valid-pixel list covers (16 x 75 776 = 1.2 M pixels; 32 KB of smem)

// A word of a partial row is 64 bits: the fp32 partial sum in the low half, the epoch of the iteration that produced it in the
// high half, written with ONE st.relaxed.gpu.b64 and read with ONE ld.relaxed.gpu.b64.  An aligned 64-bit scalar access is
// single-copy atomic in the PTX memory model, so a reader that sees the epoch in a word holds that epoch's sum — per word,
// with no assumption about sectors, warps or store coalescing, and with no fence: on sm_100a an acquire is MEMBAR + CCTL.IVALL,
// which throws away the SM's L1 — and with it the vertex / normal maps this CTA re-reads in every one of the 10/5/4 iterations
// of a level.  (Round 1 packed seven sums and one epoch per 32-byte sector and relied on the sector being written as a unit.)
__device__ __forceinline__ unsigned long long ld_partial(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_partial(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

#ifdef TFB_ICP_PROFILE
__device__ long long g_icp_prof[64 * 8];
__device__ long long g_icp_cta[256 * 4];   // iteration 12 (level 0): per CTA globaltimer at pixel start, pixel end, row stored, fold done
__device__ int g_icp_pix[256 * 24];        // per CTA: SM id, list length at level 0, then per iteration 0..18 the cycles of its pixel phase (clock64 of thread 0)
#define ICP_STAMP(slot) do { if (tid == 0 && blockIdx.x < 256 && iter_global < 19) { if ((slot) == 0) pix_t0 = clock64(); if ((slot) == 1) g_icp_pix[blockIdx.x * 24 + 2 + iter_global] = (int)(clock64() - pix_t0); } \
    if (blockIdx.x == 0 && tid == 0 && iter_global < 64) g_icp_prof[iter_global * 8 + (slot)] = clock64(); \
    if (iter_global == 12 && tid == 0 && blockIdx.x < 256 && ((slot) == 0 || (slot) == 1 || (slot) == 2 || (slot) == 4)) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); \
        g_icp_cta[blockIdx.x * 4 + ((slot) == 0 ? 0 : (slot) == 1 ? 1 : (slot) == 2 ? 2 : 3)] = gt; } } while (0)
#else
#define ICP_STAMP(slot) do { } while (0)
#endif

// Pixel phase of one iteration: U independent pixels in flight per thread so the dependent chain own pixel -> projection ->
// gathered model pixel overlaps across pixels (find_coresp + the row of icp_helper_kernel, proj_icp.cu:80-117,369-376).
// LIST: the thread's pixels come from the CTA's list of valid pixels (built once per level, see build_valid_list) instead of
// its static slots; `npx` is then the length of the list, `gtid` the thread index and `gstride` the CTA size.
template <int U, bool LIST>
__device__ __forceinline__ void icp_pixels(const IcpLevelArgs& L, const IcpAllArgs& a, int npx, int gtid, int gstride,
                                           const float* __restrict__ s_aff, float (&acc)[ICP_ACC], const int* __restrict__ s_list) {
    const float r00 = s_aff[0], r01 = s_aff[1], r02 = s_aff[2], t0 = s_aff[3];
    const float r10 = s_aff[4], r11 = s_aff[5], r12 = s_aff[6], t1 = s_aff[7];
    const float r20 = s_aff[8], r21 = s_aff[9], r22 = s_aff[10], t2 = s_aff[11];
    for (int base = gtid; base < npx; base += gstride * U) {
        // stage 1: own pixel loads for U independent pixels
        float4 v[U], nc[U];
        int idx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            idx[u] = base + u * gstride;
            const bool in = idx[u] < npx;
            if (LIST) idx[u] = in ? s_list[idx[u]] : 0;
            const float qn = __int_as_float(0x7fffffff);
            v[u] = in ? __ldg(L.vcurr + idx[u]) : make_float4(qn, qn, qn, qn);
            nc[u] = in ? __ldg(L.ncurr + idx[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // stage 2: project, issue the gathers
        float sx[U], sy[U], sz[U];
        float4 d[U], nd[U];
        bool valid[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            // find_coresp, proj_icp.cu:80-117 (points variant)
            sx[u] = dot3(r00, r01, r02, v[u].x, v[u].y, v[u].z) + t0;
            sy[u] = dot3(r10, r11, r12, v[u].x, v[u].y, v[u].z) + t1;
            sz[u] = dot3(r20, r21, r22, v[u].x, v[u].y, v[u].z) + t2;
            const float cox = __fmaf_rn(L.fx, __fdiv_rn(sx[u], sz[u]), L.cx);
            const float coy = __fmaf_rn(L.fy, __fdiv_rn(sy[u], sz[u]), L.cy);
            valid[u] = !isnan(v[u].x) && !(sz[u] <= 0 || cox < 0 || coy < 0 || cox >= L.w || coy >= L.h);
            const int pidx = valid[u] ? ((int)coy * L.w + (int)cox) : 0;   // point-sampled texel
            d[u] = __ldg(L.vprev + pidx);
            nd[u] = __ldg(L.nprev + pidx);
        }
        // stage 3: tests + accumulation
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!valid[u] || isnan(d[u].x)) continue;
            const float ex = sx[u] - d[u].x, ey = sy[u] - d[u].y, ez = sz[u] - d[u].z;
            if (dot3(ex, ey, ez, ex, ey, ez) > a.dist2_thres) continue;
            const float nsx = dot3(r00, r01, r02, nc[u].x, nc[u].y, nc[u].z);
            const float nsy = dot3(r10, r11, r12, nc[u].x, nc[u].y, nc[u].z);
            const float nsz = dot3(r20, r21, r22, nc[u].x, nc[u].y, nc[u].z);
            if (fabsf(dot3(nsx, nsy, nsz, nd[u].x, nd[u].y, nd[u].z)) < a.min_cosine) continue;
            float row[7];
            row[0] = sy[u] * nd[u].z - sz[u] * nd[u].y;
            row[1] = sz[u] * nd[u].x - sx[u] * nd[u].z;
            row[2] = sx[u] * nd[u].y - sy[u] * nd[u].x;
            row[3] = nd[u].x; row[4] = nd[u].y; row[5] = nd[u].z;
            row[6] = dot3(nd[u].x, nd[u].y, nd[u].z, d[u].x - sx[u], d[u].y - sy[u], d[u].z - sz[u]);
            int k = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = i; j < 7; ++j) { acc[k] = __fmaf_rn(row[i], row[j], acc[k]); ++k; }
            acc[ICP_TERMS] += 1.f;
        }
    }

}

__global__ void __launch_bounds__(ICPA_THREADS, 1)
    k_icp_all(IcpAllArgs a, DevState* __restrict__ ds, float* __restrict__ partial, unsigned int* host_state, unsigned int host_seq,
              int* __restrict__ vis, const int* list0, const int* list1) {
    __shared__ float s_warp[ICPA_WARPS][ICP_ACC];
    __shared__ double s_part[ICPA_WARPS][32];
    __shared__ double s_tot[ICP_ACC];
    __shared__ __align__(16) float s_totf[ICP_ACC];
    __shared__ __align__(16) float s_aff[16];
    __shared__ int s_ok;
    __shared__ int s_list[ICP_LIST_SLOTS * ICPA_THREADS];       // the CTA's valid pixels of the current level
    __shared__ int s_cnt[ICP_LIST_SLOTS * ICPA_WARPS + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = gridDim.x;
    // pixel slot of this thread: warps are dealt round-robin over the CTAs, so a CTA's 16 warps (x up to 5 pixel slots) sample
    // the whole image instead of five contiguous runs of 512 pixels — the density of valid correspondences, and with it the
    // length of the pixel phase, is then the same for every CTA (they all wait for the slowest one, 19 times a frame)
    const int gtid = (warp * (int)gridDim.x + (int)blockIdx.x) * 32 + lane;
    const int gstride = nblk * ICPA_THREADS;

    if (tid < 16) s_aff[tid] = ((tid % 5) == 0) ? 1.f : 0.f;   // affine = Identity, projective_icp.cpp:174
    if (tid == 0) s_ok = 1;
    __syncthreads();

    int iter_global = 0;
    bool ok = true;
#ifdef TFB_ICP_PROFILE
    if (blockIdx.x == 0 && tid == 0) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_icp_prof[60 * 8] = gt; }
    long long pix_t0 = 0;
    if (tid == 0 && blockIdx.x < 256) { unsigned int smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); g_icp_pix[blockIdx.x * 24] = (int)smid; }
#endif
    for (int l = a.levels - 1; l >= 0 && ok; --l) {
        const IcpLevelArgs L = a.lv[l];
        const int npx = L.w * L.h;
        // Which of this CTA's pixels have a vertex at all does not change during the level (the current maps are fixed, only
        // the transform moves): the CTA compacts them once, in slot order (deterministic), into shared memory, and every
        // iteration of the level walks the list — at 640x480 with 44 % of the image empty that is 2.3 instead of 4.05 pixels
        // per thread, 10 times over.  The round-robin dealing of warps makes the lists of all CTAs equally long.
        const int slots = (npx + gstride - 1) / gstride;
        int n_list = -1;
        // Level 0 of the frame path: the list of valid pixels exists already (built inside the preprocessing launches on the second
        // stream, tfb_imgproc.cu), and every CTA takes an equal share of it.  Compacting its own slots, a CTA ends up with 896..1092 entries at
        // 640x480 (the content is spatially coherent), and the dozen CTAs past 1024 entries — a third pixel per thread in one
        // warp, 2.0 us against a median of 1.25 — set the length of every one of the ten level-0 iterations.
        if (L.vlist != nullptr && L.iters > 0) {
            const int n_total = (int)__ldcg(L.vlist_n);
            const int share = (n_total + nblk - 1) / nblk;
            if (share <= ICP_LIST_SLOTS * ICPA_THREADS) {
                __syncthreads();   // the previous level's list is no longer read
                const int begin = min((int)blockIdx.x * share, n_total);
                n_list = min(share, n_total - begin);
                for (int i = tid; i < n_list; i += ICPA_THREADS) s_list[i] = __ldcg(L.vlist + begin + i);
                __syncthreads();
            }
        }
        if (n_list < 0 && slots <= ICP_LIST_SLOTS && L.iters > 0) {
            __syncthreads();   // the previous level's list is no longer read
            int cnt = 0;
            for (int u = 0; u < slots; ++u) {
                const int idx = gtid + u * gstride;
                const bool valid = idx < npx && !isnan(__ldg(&L.vcurr[idx].x));
                const unsigned int m = __ballot_sync(0xffffffffu, valid);
                if (lane == 0) s_cnt[u * ICPA_WARPS + warp] = __popc(m);
                cnt |= valid ? (1 << u) : 0;
            }
            __syncthreads();
            if (tid == 0) {   // exclusive prefix over (slot, warp): <= 16 x 16 counters
                int run = 0;
                for (int i = 0; i < slots * ICPA_WARPS; ++i) { const int c = s_cnt[i]; s_cnt[i] = run; run += c; }
                s_cnt[ICP_LIST_SLOTS * ICPA_WARPS] = run;
            }
            __syncthreads();
            for (int u = 0; u < slots; ++u) {
                const bool valid = (cnt >> u) & 1;
                const unsigned int m = __ballot_sync(0xffffffffu, valid);
                if (valid) s_list[s_cnt[u * ICPA_WARPS + warp] + __popc(m & ((1u << lane) - 1u))] = gtid + u * gstride;
            }
            __syncthreads();
            n_list = s_cnt[ICP_LIST_SLOTS * ICPA_WARPS];
#ifdef TFB_ICP_PROFILE
            if (tid == 0 && blockIdx.x < 256 && l == 0) g_icp_pix[blockIdx.x * 24 + 1] = n_list;
#endif
        }
        for (int it = 0; it < L.iters && ok; ++it, ++iter_global) {
            float acc[ICP_ACC];
#pragma unroll
            for (int i = 0; i < ICP_ACC; ++i) acc[i] = 0.f;
            ICP_STAMP(0);
#ifdef TFB_ICP_PROFILE
            if (iter_global == 0 && blockIdx.x == 0 && tid == 0) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_icp_prof[60 * 8 + 1] = gt; }
#endif

            // pixels in flight per thread sized to the level: 5 at 640x480 (4.05 pixels per thread), 2 and 1 on the coarse levels
            if (n_list >= 0) {
                // Pixels in flight per WARP: the lists of the 148 CTAs differ by +-10 % (896..1092 entries at 640x480), and a CTA
                // whose list is a few entries past a multiple of the CTA size used to run a whole extra pixel in every thread,
                // predicated off in all but one warp — the same dozen CTAs took 2.15 us for a pixel phase whose median is 1.18 us,
                // in every iteration, and everybody waits for them.  A warp now runs exactly as many pixels as ITS threads have
                // (the first rem / 32 warps one more than the others, all of them in flight at once).  Pixel -> thread assignment
                // and the order of a thread's accumulations are unchanged, so the sums are bit-identical.
                const int full = n_list / ICPA_THREADS, rem = n_list - full * ICPA_THREADS;
                const int per_thread = full + ((warp * 32 < rem) ? 1 : 0);
                if (per_thread == 1) icp_pixels<1, true>(L, a, n_list, tid, ICPA_THREADS, s_aff, acc, s_list);
                else if (per_thread == 2) icp_pixels<2, true>(L, a, n_list, tid, ICPA_THREADS, s_aff, acc, s_list);
                else if (per_thread == 3) icp_pixels<3, true>(L, a, n_list, tid, ICPA_THREADS, s_aff, acc, s_list);
                else if (per_thread > 3) icp_pixels<ICPA_UNROLL, true>(L, a, n_list, tid, ICPA_THREADS, s_aff, acc, s_list);
            } else {
                icp_pixels<ICPA_UNROLL, false>(L, a, npx, gtid, gstride, s_aff, acc, nullptr);
            }

            ICP_STAMP(1);
            // CTA reduction -> one row of partials.  Warp level: a transposing butterfly — at each of the five steps a lane
            // keeps one half of its values and trades the other half with its partner, so the 28 sums cost 31 shuffles
            // instead of 28 x 5, and lane L ends up holding the warp total of term L.
            {
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = (i < ICP_ACC) ? acc[i] : 0.f;
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    const bool upper = (lane & o) != 0;
#pragma unroll
                    for (int i = 0; i < o; ++i) {
                        const float send = upper ? v[i] : v[i + o];
                        const float keep = upper ? v[i + o] : v[i];
                        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                    }
                }
                if (lane < ICP_ACC) s_warp[warp][lane] = v[0];
            }
            __syncthreads();
            // The CTA's row of partials IS its barrier arrival: 32 words of 64 bits, {sum, epoch} each (28 sums + 4 words of
            // padding so every lane of the fold runs the same code); the epoch is unique across iterations and launches.  No
            // counter, no fence, no second round trip: the fold below re-reads a word until it carries the epoch.  Rows are
            // double buffered by iteration parity; a CTA can only write iteration i+2 after every CTA has finished reading i.
            const unsigned int epoch = a.epoch_base + (unsigned)iter_global + 1u;
            unsigned long long* prow = reinterpret_cast<unsigned long long*>(partial) + (size_t)(iter_global & 1) * nblk * 32;
            if (tid < 32) {
                float sacc = 0.f;
                if (tid < ICP_ACC) {
#pragma unroll
                    for (int wi = 0; wi < ICPA_WARPS; ++wi) sacc += s_warp[wi][tid];
                }
                st_partial(prow + blockIdx.x * 32 + tid, ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(sacc));
            }
            ICP_STAMP(2);
            ICP_STAMP(3);

            // every CTA: fold all rows in the same order (fp64), solve, update its copy of the transform
            {
                // warp p takes CTAs p, p+16, ...; lane = word of the row; all of a warp's loads are in flight before the first add
                const int k = lane, part = warp;
                double sd = 0;
                for (int b0 = part; b0 < nblk; b0 += ICPA_WARPS * 10) {
                    unsigned long long tmp[10];
                    // poll: every round re-reads ALL words that are still missing at once (one L2 round trip per round, not one
                    // per late row), until the last of this warp's rows carries the epoch in every word
                    unsigned int pending = 0;
#pragma unroll
                    for (int j = 0; j < 10; ++j) {
                        tmp[j] = (unsigned long long)epoch << 32;
                        if (b0 + j * ICPA_WARPS < nblk) pending |= 1u << j;
                    }
                    while (pending) {
#pragma unroll
                        for (int j = 0; j < 10; ++j)
                            if (pending & (1u << j)) tmp[j] = ld_partial(prow + (b0 + j * ICPA_WARPS) * 32 + k);
                        unsigned int still = 0;
#pragma unroll
                        for (int j = 0; j < 10; ++j)
                            if ((pending & (1u << j)) && __any_sync(0xffffffffu, (unsigned int)(tmp[j] >> 32) != epoch)) still |= 1u << j;
                        pending = still;
                    }
#pragma unroll
                    for (int j = 0; j < 10; ++j) sd += (double)__uint_as_float((unsigned int)tmp[j]);
                }
                s_part[part][k] = sd;   // lanes 28..31 sum the padding words (zeros); nobody reads them
            }
            __syncthreads();
            if (tid < ICP_ACC) {
                double sd = 0;
#pragma unroll
                for (int p = 0; p < ICPA_WARPS; ++p) sd += s_part[p][tid];
                s_tot[tid] = sd;
                s_totf[tid] = (float)sd;   // StreamHelper::get hands the host floats: 28 conversions side by side, not 27 in the solving thread
            }
            __syncthreads();
            ICP_STAMP(4);
            if (tid == 0 && !icp_solve_update(s_totf, s_aff)) s_ok = 0;   // s_aff is updated in place, and only on success
            ICP_STAMP(5);
            __syncthreads();
            ok = (s_ok != 0);
        }
    }

    // setToType3 (SceneReconstructionEngine_host.cu:343-348) for the allocation stage that follows a tracked frame: every entry of
    // the current visible list becomes "visible in the previous frame, to be re-tested".  It used to be the first launch of the
    // frame's tail (k_set_type3: 3 us of kernel and a launch boundary on the critical path between this kernel and k_mark); here
    // the 147 CTAs that have nothing left to do take it — at the same point of the stream order, under the same condition
    // (skipped when tracking failed, as k_set_type3 is), so the allocation stage sees exactly what it saw before.
#ifdef TFB_ICP_PROFILE
    if (blockIdx.x == 0 && tid == 0) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_icp_prof[60 * 8 + 2] = gt; }
#endif
    if (vis != nullptr && ok) {
        // CTA 0 has the frame's result to compose and publish (below): the others share the list among themselves
        const int workers = gridDim.x > 1 ? (int)gridDim.x - 1 : 1, me = gridDim.x > 1 ? (int)blockIdx.x - 1 : 0;
        if (me >= 0) {
            const int* __restrict__ list = ds->cur_list ? list1 : list0;
            const int n = ds->n_visible;
            for (int i = me * ICPA_THREADS + tid; i < n; i += workers * ICPA_THREADS) vis[list[i]] = 3;
        }
    }
    if (blockIdx.x != 0) return;
#ifdef TFB_ICP_PROFILE
    if (tid == 0) g_icp_prof[61 * 8] = clock64();
#endif
    if (warp == 0) {   // the first warp composes the frame's pose: the fp64 inverse spread over its lanes (tfb_pose.cuh)
        if (lane == 0) {
            ds->icp_failed = ok ? 0 : 1;
            ds->icp_corresp = (int)s_tot[ICP_TERMS];
        }
        // ds->affine: the running product — on failure the product of the iterations before the failing one, which is what
        // ProjectiveICP::estimateTransform leaves in its `affine` argument when it returns false (projective_icp.cpp:197-203)
        if (lane < 16) ds->affine[lane] = s_aff[lane];
        if (ok && a.update_pose) {
            float aff[16], prev[16], nw[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { aff[i] = s_aff[i]; prev[i] = ds->pose_c2w[i]; }
            pose_mul(prev, aff, nw);
            store_pose_c2w_warp(ds, nw, lane);
        }
    }
    // The frame's result — pose, verdict, counters: the whole state block — goes straight into the host's pinned mirror
    // (zero-copy), followed by a sequence number the host spins on: no D2H copy to enqueue, no stream synchronise to wake
    // up from (the reference: 19 stream syncs + a blocking cudaMemcpy per frame just for ICP).
#ifdef TFB_ICP_PROFILE
    if (tid == 0) g_icp_prof[61 * 8 + 1] = clock64();
#endif
    if (host_state != nullptr) {
        __syncthreads();
        constexpr int WORDS = (int)(sizeof(DevState) / sizeof(unsigned int));
        const unsigned int* src = reinterpret_cast<const unsigned int*>(ds);
        for (int i = tid; i < WORDS; i += ICPA_THREADS) host_state[i] = src[i];
        __syncthreads();
#ifdef TFB_ICP_PROFILE
        if (tid == 0) g_icp_prof[61 * 8 + 2] = clock64();
#endif
        if (tid == 0) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned int*>(host_state + WORDS) = host_seq;
        }
#ifdef TFB_ICP_PROFILE
        if (tid == 0) g_icp_prof[61 * 8 + 3] = clock64();
#endif
    }
    // the other half of k_set_type3: the allocation counters of the frame start from zero (after the snapshot above, which
    // still reports the previous frame's)
    if (vis != nullptr && ok && tid == 0) { ds->n_claimed = 0; ds->n_new_frame = 0; }
#ifdef TFB_ICP_PROFILE
    if (tid == 0) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_icp_prof[60 * 8 + 3] = gt; }
#endif
}

#ifdef TFB_ICP_PROFILE
extern "C" __attribute__((visibility("default"))) int tfb_debug_icp_pix(int* out6144) {
    return cudaMemcpyFromSymbol(out6144, g_icp_pix, sizeof(int) * 256 * 24) == cudaSuccess ? 0 : -2;
}
extern "C" __attribute__((visibility("default"))) int tfb_debug_icp_cta(long long* out1024) {
    return cudaMemcpyFromSymbol(out1024, g_icp_cta, sizeof(long long) * 1024) == cudaSuccess ? 0 : -2;
}
extern "C" __attribute__((visibility("default"))) int tfb_debug_icp_profile(long long* out512) {
    return cudaMemcpyFromSymbol(out512, g_icp_prof, sizeof(long long) * 64 * 8) == cudaSuccess ? 0 : -2;
}
#endif

static int launch_icp_args(tfb_ctx* c, IcpAllArgs& a, int total);

// ProjectiveICP::estimateTransform on caller-owned pyramids (projective_icp.cpp:169-212), for the C++ ProjectiveICP class
int launch_icp_all_ext(tfb_ctx* c, int levels, const float* const* vcurr, const float* const* ncurr, const float* const* vprev,
                       const float* const* nprev, int cols, int rows, const int* iters, float dist_thres, float angle_thres) {
    if (levels < 1 || levels > MAX_LEVELS) return set_err(c, TFB_ERR_ARG, "icp: 1..4 pyramid levels");
    IcpAllArgs a;
    memset(&a, 0, sizeof(a));
    a.levels = levels;
    a.min_cosine = cosf(angle_thres);
    a.dist2_thres = dist_thres * dist_thres;
    a.update_pose = 0;
    int total = 0, w = cols, h = rows;
    for (int l = 0; l < levels; ++l) {
        const int div = 1 << l;
        a.lv[l].vcurr = (const float4*)vcurr[l]; a.lv[l].ncurr = (const float4*)ncurr[l];
        a.lv[l].vprev = (const float4*)vprev[l]; a.lv[l].nprev = (const float4*)nprev[l];
        a.lv[l].w = w; a.lv[l].h = h; a.lv[l].iters = iters[l];
        a.lv[l].fx = c->p.fx / div; a.lv[l].fy = c->p.fy / div; a.lv[l].cx = c->p.cx / div; a.lv[l].cy = c->p.cy / div;
        total += iters[l];
        w /= 2; h /= 2;
    }
    return launch_icp_args(c, a, total);
}

int launch_icp_all(tfb_ctx* c, bool update_pose) {
    const tfb_params& p = c->p;
    IcpAllArgs a;
    memset(&a, 0, sizeof(a));
    a.levels = c->levels;
    a.min_cosine = cosf(p.icp_angle_thres);              // ComputeIcpHelper ctor, projective_icp.cpp:11-15
    a.dist2_thres = p.icp_dist_thres * p.icp_dist_thres;
    a.update_pose = update_pose ? 1 : 0;
    int total = 0;
    for (int l = 0; l < c->levels; ++l) {
        const int div = 1 << l;                           // setLevelIntr, projective_icp.cpp:17-23
        a.lv[l].vcurr = c->lv[l].vcurr; a.lv[l].ncurr = c->lv[l].ncurr;
        a.lv[l].vprev = c->lv[l].vprev; a.lv[l].nprev = c->lv[l].nprev;
        a.lv[l].w = c->lv[l].w; a.lv[l].h = c->lv[l].h; a.lv[l].iters = p.icp_iters[l];
        a.lv[l].fx = p.fx / div; a.lv[l].fy = p.fy / div; a.lv[l].cx = p.cx / div; a.lv[l].cy = p.cy / div;
        total += p.icp_iters[l];
    }
    if (c->vlist_ready) { a.lv[0].vlist = c->icp_vlist; a.lv[0].vlist_n = c->icp_vscan; }
    return launch_icp_args(c, a, total);
}

static int launch_icp_args(tfb_ctx* c, IcpAllArgs& a, int total) {
    if (total == 0) return TFB_OK;
    if (total > 63) return set_err(c, TFB_ERR_ARG, "icp: more than 63 iterations in one coarse-to-fine loop (64 row epochs per launch)");
    if (c->icp_grid == 0) {
        int per_sm = 0, sms = 0;
        TFB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_icp_all, ICPA_THREADS, 0));
        TFB_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        if (per_sm < 1) return set_err(c, TFB_ERR_CUDA, "icp: kernel does not fit on an SM");
        c->icp_grid = sms;   // one CTA per SM: the partial rows every CTA folds grow with the grid
        // the partial buffer holds max(icp_max_blocks, 1024) rows whatever the context's own image size is: a small context
        // (the C++ mirror's utility context is 8x8) must not squeeze tfb_icp_estimate_ext's 640x480 pyramids into one CTA
        const int cap = c->icp_max_blocks > 1024 ? c->icp_max_blocks : 1024;
        if (c->icp_grid > cap) c->icp_grid = cap;
    }
    a.epoch_base = (unsigned int)(++c->icp_launches) * 64u;   // rows of earlier launches can never match
    DevState* ds = c->ds;
    float* partial = c->icp_partial;
    // the frame path asks for the zero-copy publish (c->publish_seq != 0); stage-level calls read the state back themselves
    unsigned int* host_state = c->publish_seq ? reinterpret_cast<unsigned int*>(c->hs) : nullptr;
    unsigned int host_seq = c->publish_seq;
    // the frame path lets the kernel's idle CTAs do setToType3 for the allocation stage that follows (launch_allocate skips its own)
    int* vis = c->icp_fuse_type3 ? c->vis_type : nullptr;
    const int *l0 = c->vis_list[0], *l1 = c->vis_list[1];
    if (vis) c->type3_done = true;
    void* args[] = {&a, &ds, &partial, &host_state, &host_seq, &vis, &l0, &l1};
    TFB_KT(c, K_ICP_ALL);
    TFB_CUDA(c, cudaLaunchCooperativeKernel((const void*)k_icp_all, dim3(c->icp_grid), dim3(ICPA_THREADS), args, 0, c->stream));
    TFB_LAUNCH_CHECK(c);
    return TFB_OK;
}

// only the stage-level test entry (tfb_icp_reduce) still needs an explicit state reset; the frame path folds it
// into the first iteration (first_iter)
__global__ void k_icp_begin(DevState* ds) {
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
                         int w, int h, float fx, float fy, float cx, float cy, bool solve, float* out27_dev, bool first_iter,
                         bool last_iter) {
    IcpArgs a;
    a.first_iter = first_iter ? 1 : 0;
    a.last_iter = last_iter ? 1 : 0;
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
