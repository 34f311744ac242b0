// REFERENCE-ARM INFRASTRUCTURE (not product code; nothing under topfusion_b200/, src/ or include/
// links or loads this).
//
// C-ABI handle around the reference library itself — /root/reference/tfusion, patched only as far
// as patch_ref.py lists so that it compiles with CUDA 12.9 for sm_100a — so that Python (ctypes)
// can, on the B200,
//   (a) run the reference's OWN device kernels on given inputs and read their outputs back:
//       imgproc.cu (compute_dists, bilateral, truncate, pyramid, points/normals, resize),
//       proj_icp.cu (icp_helper_kernel + icp_final_reduce_kernel at a fixed transform, and the whole
//       ProjectiveICP::estimateTransform loop) -> golden fixtures for SURVEY §8 rows a1-a8, a17;
//   (b) run the reference's whole TopFu::operator() on a frame sequence, read poses and scene state
//       back, and time it -> the GPU-vs-GPU anchor of BASELINE.md §3 row 2.
// Every computation below is a call into reference code; this file only moves buffers.
//
// Deviation kept deliberately: the reference never initialises entriesVisibleType /
// visibleEntryIDs (cudaMalloc'ed garbage); refgpu_create zeroes them once, as InfiniTAM's MemoryBlock
// did, otherwise the very first visible list is undefined.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>

#define private public  // TopFu / RenderState_VH keep their frame buffers private; the harness reads them
#define protected public
#include "precomp.hpp"
#include "internal.hpp"
#include <tfusion/topfu.hpp>
#undef private
#undef protected

using namespace tfusion;

namespace {
struct Handle {
    TopFu* fu;
    cuda::Depth depth;
    int rows, cols;
    std::streambuf* saved_cout;
    std::ostringstream sink;
};
inline void ck(cudaError_t e, const char* what) {
    if (e != cudaSuccess) { std::fprintf(stderr, "refgpu: %s: %s\n", what, cudaGetErrorString(e)); }
}
template <class T> void down2d(const cuda::DeviceArray2D<T>& a, void* host) {
    ck(cudaMemcpy2D(host, a.cols() * sizeof(T), a.ptr(), a.step(), a.cols() * sizeof(T), a.rows(), cudaMemcpyDeviceToHost), "download");
}
}  // namespace

extern "C" {

const char* refgpu_describe() {
#ifdef REFGPU_NODEBUG
    return "3d-scan/topfusion reference library, patched to compile (patch_ref.py P1-P5: debug downloads/print/render of topfu.cpp removed)";
#else
    return "3d-scan/topfusion reference library, patched to compile (patch_ref.py P1-P4: per-frame code as shipped, incl. its debug downloads)";
#endif
}

int refgpu_hash_total_entries() { return VoxelBlockHash::noTotalEntries; }
int refgpu_num_blocks() { return SDF_LOCAL_BLOCK_NUM; }

// TopFuParams::default_params() with the fields the hot path reads overridden (topfu.cpp:12-53).
void* refgpu_create(int rows, int cols, const float intr[4], float voxel, float mu, int maxW, float vf_min, float vf_max,
                    const int iters[4], float icp_trunc, int quiet) {
    Handle* h = new Handle;
    h->rows = rows; h->cols = cols; h->saved_cout = nullptr;
    if (quiet) h->saved_cout = std::cout.rdbuf(h->sink.rdbuf());  // the per-frame pose print goes to a string
    TopFuParams p = TopFuParams::default_params();
    p.rows = rows; p.cols = cols;
    p.intr = Intr(intr[0], intr[1], intr[2], intr[3]);
    p.icp_truncate_depth_dist = icp_trunc;
    p.icp_iter_num.assign(iters, iters + 4);
    p.sceneParams = new SceneParams(mu, maxW, voxel, vf_min, vf_max, false);
    h->fu = new TopFu(p);
    RenderState_VH* rs = (RenderState_VH*)h->fu->renderState;
    ck(cudaMemset(rs->GetEntriesVisibleType(), 0, VoxelBlockHash::noTotalEntries), "zero entriesVisibleType (F7)");
    ck(cudaMemset(rs->GetVisibleEntryIDs(), 0, SDF_LOCAL_BLOCK_NUM * sizeof(int)), "zero visibleEntryIDs (F7)");
    h->depth.create(rows, cols);
    ck(cudaDeviceSynchronize(), "create");
    return h;
}

void refgpu_destroy(void* hv) {
    Handle* h = (Handle*)hv;
    if (h->saved_cout) std::cout.rdbuf(h->saved_cout);
    // the reference's TopFu has no destructor and leaks its engines (topfu.hpp:106-109); the process ends soon
    delete h;
}

void refgpu_upload(void* hv, const uint16_t* depth_host) {
    Handle* h = (Handle*)hv;
    h->depth.upload(depth_host, h->cols * sizeof(uint16_t), h->rows, h->cols);  // demo.cpp:100
}

// demo.cpp:104 — returns operator()'s bool.  The string sink is emptied so a long run does not grow it.
int refgpu_step(void* hv) {
    Handle* h = (Handle*)hv;
    bool ok = (*h->fu)(h->depth);
    if (h->saved_cout) { h->sink.str(std::string()); }
    return ok ? 1 : 0;
}

int refgpu_frame(void* hv, const uint16_t* depth_host) { refgpu_upload(hv, depth_host); return refgpu_step(hv); }

void refgpu_sync() { ck(cudaDeviceSynchronize(), "sync"); }

int refgpu_num_poses(void* hv) { return (int)((Handle*)hv)->fu->poses_.size(); }
int refgpu_frame_counter(void* hv) { return ((Handle*)hv)->fu->frame_counter_; }
void refgpu_pose(void* hv, int idx, float out16[16]) {
    Affine3f p = ((Handle*)hv)->fu->getCameraPose(idx);
    memcpy(out16, p.matrix.val, 16 * sizeof(float));
}

// which: 0 = curr_, 1 = prev_ (the model maps the next ICP will use).  Dense float4 rows x cols of the level.
void refgpu_get_maps(void* hv, int which, int level, float* points, float* normals) {
    TopFu* f = ((Handle*)hv)->fu;
    cuda::Frame& fr = which ? f->prev_ : f->curr_;
    if (points) down2d(fr.points_pyr[level], points);
    if (normals) down2d(fr.normals_pyr[level], normals);
}
void refgpu_get_depth_pyr(void* hv, int level, uint16_t* out) { down2d(((Handle*)hv)->fu->curr_.depth_pyr[level], out); }
void refgpu_get_dists(void* hv, float* out) { down2d(((Handle*)hv)->fu->dists_, out); }
void refgpu_get_raycast(void* hv, float* out) { down2d(((Handle*)hv)->fu->renderState->raycastResult, out); }
void refgpu_get_range_image(void* hv, float* out) { down2d(((Handle*)hv)->fu->renderState->renderingRangeImage, out); }

// hash table as int32[total][4]: {pos.x | pos.y << 16, pos.z (low 16), offset, ptr} = the 16-byte HashEntry verbatim
void refgpu_get_hash(void* hv, void* out) {
    TopFu* f = ((Handle*)hv)->fu;
    ck(cudaMemcpy(out, f->scene->index.GetEntries(), (size_t)VoxelBlockHash::noTotalEntries * sizeof(HashEntry), cudaMemcpyDeviceToHost), "hash");
}
// voxel pool: num_blocks x 512 x {short sdf; uchar w; pad}
void refgpu_get_voxels(void* hv, void* out) {
    TopFu* f = ((Handle*)hv)->fu;
    ck(cudaMemcpy(out, f->scene->localVBA.GetVoxelBlocks(), (size_t)SDF_LOCAL_BLOCK_NUM * SDF_BLOCK_SIZE3 * sizeof(Voxel_s), cudaMemcpyDeviceToHost), "voxels");
}
int refgpu_get_visible(void* hv, int* ids, unsigned char* types) {
    TopFu* f = ((Handle*)hv)->fu;
    RenderState_VH* rs = (RenderState_VH*)f->renderState;
    int n = rs->noVisibleEntries;
    if (ids) ck(cudaMemcpy(ids, rs->GetVisibleEntryIDs(), (size_t)n * sizeof(int), cudaMemcpyDeviceToHost), "visible ids");
    if (types) ck(cudaMemcpy(types, rs->GetEntriesVisibleType(), VoxelBlockHash::noTotalEntries, cudaMemcpyDeviceToHost), "visible types");
    return n;
}
void refgpu_get_counters(void* hv, int out[3]) {
    TopFu* f = ((Handle*)hv)->fu;
    out[0] = f->scene->localVBA.lastFreeBlockId;
    out[1] = f->scene->index.GetLastFreeExcessListId();
    out[2] = ((RenderState_VH*)f->renderState)->noVisibleEntries;
}
// viewer path (demo.cpp:52): TopFu::renderImage -> rows x cols x uchar4
void refgpu_render(void* hv, unsigned char* out) {
    TopFu* f = ((Handle*)hv)->fu;
    cuda::image4u img;
    f->renderImage(img);
    down2d(img, out);
}

// ---------------------------------------------------------------------------------------------
// Stage level: the reference's imgproc.cu kernels through its own public wrappers
// (include/tfusion/cuda/imgproc.hpp; call order of topfu.cpp:166-197).  Outputs (any may be null):
//   dists f32, bilateral u16 (before truncation), depth pyramid u16 per level, points/normals float4 per level.
// ---------------------------------------------------------------------------------------------
void refgpu_stage_imgproc(const uint16_t* depth_host, int rows, int cols, const float intr[4], int ksz, float sigma_s, float sigma_d,
                          float trunc, int levels, float* dists, uint16_t* bilateral, uint16_t** depth_pyr, float** points, float** normals) {
    cuda::Depth depth;
    depth.create(rows, cols);
    depth.upload(depth_host, cols * sizeof(uint16_t), rows, cols);
    Intr in(intr[0], intr[1], intr[2], intr[3]);
    cuda::Dists d;
    cuda::computeDists(depth, d, in);
    if (dists) down2d(d, dists);
    std::vector<cuda::Depth> pyr(levels);
    std::vector<cuda::Cloud> pts(levels);
    std::vector<cuda::Normals> nrm(levels);
    cuda::depthBilateralFilter(depth, pyr[0], ksz, sigma_s, sigma_d);
    if (bilateral) down2d(pyr[0], bilateral);
    if (trunc > 0) cuda::depthTruncation(pyr[0], trunc);
    for (int i = 1; i < levels; ++i) cuda::depthBuildPyramid(pyr[i - 1], pyr[i], sigma_d);
    for (int i = 0; i < levels; ++i) cuda::computePointNormals(in(i), pyr[i], pts[i], nrm[i]);
    cuda::waitAllDefaultStream();
    for (int i = 0; i < levels; ++i) {
        if (depth_pyr && depth_pyr[i]) down2d(pyr[i], depth_pyr[i]);
        if (points && points[i]) down2d(pts[i], points[i]);
        if (normals && normals[i]) down2d(nrm[i], normals[i]);
    }
}

// resize_points_normals_kernel (imgproc.cu:355-401) through cuda::resizePointsNormals
void refgpu_stage_resize(const float* points_host, const float* normals_host, int rows, int cols, float* points_out, float* normals_out) {
    cuda::Cloud p, po; cuda::Normals n, no;
    p.create(rows, cols); n.create(rows, cols);
    p.upload(points_host, cols * 16, rows, cols);
    n.upload(normals_host, cols * 16, rows, cols);
    cuda::resizePointsNormals(p, n, po, no);
    cuda::waitAllDefaultStream();
    down2d(po, points_out); down2d(no, normals_out);
}

// One ICP reduction at a fixed transform: icp_helper_kernel + icp_final_reduce_kernel + the 27-float copy
// (proj_icp.cu:359-455), exactly as ProjectiveICP::estimateTransform issues it (projective_icp.cpp:169-196).
// R9 row-major, t3; `level` only scales the intrinsics (setLevelIntr).  partials (optional): TOTAL x n_cta floats.
void refgpu_stage_icp_sums(const float* vcurr, const float* ncurr, const float* vprev, const float* nprev, int rows, int cols, int level,
                           const float intr[4], const float R9[9], const float t3[3], float dist_thres, float angle_thres, float out27[27],
                           float* partials, int* n_partials) {
    cuda::Cloud vc, vp; cuda::Normals nc, np;
    vc.create(rows, cols); vp.create(rows, cols); nc.create(rows, cols); np.create(rows, cols);
    vc.upload(vcurr, cols * 16, rows, cols); vp.upload(vprev, cols * 16, rows, cols);
    nc.upload(ncurr, cols * 16, rows, cols); np.upload(nprev, cols * 16, rows, cols);
    device::ComputeIcpHelper helper(dist_thres, angle_thres);
    helper.rows = (float)rows; helper.cols = (float)cols;
    helper.setLevelIntr(level, intr[0], intr[1], intr[2], intr[3]);
    helper.vcurr = vc; helper.ncurr = nc;
    device::Aff3f a;
    for (int i = 0; i < 3; ++i) a.R.data[i] = make_float3(R9[3 * i], R9[3 * i + 1], R9[3 * i + 2]);
    a.t = make_float3(t3[0], t3[1], t3[2]);
    helper.aff = a;
    cuda::DeviceArray2D<float> buffer;
    device::ComputeIcpHelper::allocate_buffer(buffer);
    device::ComputeIcpHelper::PageLockHelper locked;
    cudaStream_t s; ck(cudaStreamCreate(&s), "stream");
    const device::Points& v = (const device::Points&)vp;
    const device::Normals& n = (const device::Normals&)np;
    helper(v, n, buffer, locked.data, s);
    ck(cudaStreamSynchronize(s), "icp sums");
    memcpy(out27, locked.data, 27 * sizeof(float));
    int ncta = ((cols + 31) / 32) * ((rows + 7) / 8);
    if (n_partials) *n_partials = ncta;
    if (partials) ck(cudaMemcpy2D(partials, ncta * sizeof(float), buffer.ptr(), buffer.step(), ncta * sizeof(float), 27, cudaMemcpyDeviceToHost), "partials");
    ck(cudaStreamDestroy(s), "stream");
}

// The whole coarse-to-fine loop: ProjectiveICP::estimateTransform(points variant), projective_icp.cpp:169-212.
// Maps per level l (l < levels) are dense float4 (rows >> l) x (cols >> l).  Returns its bool; affine row-major 4x4.
int refgpu_stage_estimate(float** vcurr, float** ncurr, float** vprev, float** nprev, int rows, int cols, int levels, const float intr[4],
                          const int iters[4], float dist_thres, float angle_thres, float affine16[16]) {
    cuda::ProjectiveICP icp;
    icp.setDistThreshold(dist_thres);
    icp.setAngleThreshold(angle_thres);
    icp.setIterationsNum(std::vector<int>(iters, iters + 4));
    std::vector<cuda::Cloud> vc(4), vp(4);
    std::vector<cuda::Normals> nc(4), np(4);
    for (int l = 0; l < levels; ++l) {
        int r = rows >> l, c = cols >> l;
        vc[l].create(r, c); vp[l].create(r, c); nc[l].create(r, c); np[l].create(r, c);
        vc[l].upload(vcurr[l], c * 16, r, c); vp[l].upload(vprev[l], c * 16, r, c);
        nc[l].upload(ncurr[l], c * 16, r, c); np[l].upload(nprev[l], c * 16, r, c);
    }
    Affine3f aff;
    bool ok = icp.estimateTransform(aff, Intr(intr[0], intr[1], intr[2], intr[3]), vc, nc, vp, np);
    memcpy(affine16, aff.matrix.val, 16 * sizeof(float));
    return ok ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Scene / visualisation engines with an INJECTED pose (decoupled from ICP), on the handle's own scene:
//   refgpu_scene_integrate: cuda::computeDists + AllocateSceneFromDepth + IntegrateIntoScene  (topfu.cpp:166,281-282)
//   refgpu_scene_raycast:   CreateExpectedDepths + CreateICPMaps + resizePointsNormals          (topfu.cpp:306-309)
// pose_c2w: row-major 4x4 camera->world, as poses_ holds them; the w2c the engines want is pose.inv(), as in topfu.cpp.
// ---------------------------------------------------------------------------------------------
static Affine3f affine_from(const float m16[16]) { cv::Matx44f m; memcpy(m.val, m16, 64); return Affine3f(m); }

void refgpu_scene_integrate(void* hv, const uint16_t* depth_host, const float pose_c2w[16]) {
    Handle* h = (Handle*)hv;
    TopFu* f = h->fu;
    const TopFuParams& p = f->params_;
    refgpu_upload(hv, depth_host);
    cuda::computeDists(h->depth, f->dists_, p.intr);
    Affine3f pose = affine_from(pose_c2w);
    f->sceneEngine->AllocateSceneFromDepth(f->scene, p.intr, pose.inv(), f->dists_, f->renderState);
    f->sceneEngine->IntegrateIntoScene(f->scene, p.intr, pose.inv(), f->dists_, f->renderState);
    ck(cudaDeviceSynchronize(), "scene_integrate");
}

void refgpu_scene_raycast(void* hv, const float pose_c2w[16]) {
    Handle* h = (Handle*)hv;
    TopFu* f = h->fu;
    const TopFuParams& p = f->params_;
    Affine3f pose = affine_from(pose_c2w);
    Matrix4f M_d(pose.matrix(0,0),pose.matrix(1,0),pose.matrix(2,0),pose.matrix(3,0),
                 pose.matrix(0,1),pose.matrix(1,1),pose.matrix(2,1),pose.matrix(3,1),
                 pose.matrix(0,2),pose.matrix(1,2),pose.matrix(2,2),pose.matrix(3,2),
                 pose.matrix(0,3),pose.matrix(1,3),pose.matrix(2,3),pose.matrix(3,3));
    f->visualisationEngine->CreateExpectedDepths(f->scene, pose.inv(), p.intr, f->renderState);
    f->visualisationEngine->CreateICPMaps(f->scene, M_d, p.intr, f->prev_.points_pyr[0], f->prev_.normals_pyr[0], f->renderState);
    const int LEVELS = f->icp_->getUsedLevelsNum();
    for (int i = 1; i < LEVELS; ++i)
        cuda::resizePointsNormals(f->prev_.points_pyr[i - 1], f->prev_.normals_pyr[i - 1], f->prev_.points_pyr[i], f->prev_.normals_pyr[i]);
    ck(cudaDeviceSynchronize(), "scene_raycast");
}

// viewer render at an injected pose: poses_.back() is what TopFu::renderImage reads (topfu.cpp:342)
void refgpu_render_at(void* hv, const float pose_c2w[16], unsigned char* out) {
    TopFu* f = ((Handle*)hv)->fu;
    Affine3f saved = f->poses_.back();
    f->poses_.back() = affine_from(pose_c2w);
    refgpu_render(hv, out);
    f->poses_.back() = saved;
}

// the shim's OpenCV stand-ins, exported so a CPU test can pin them against cv2 (tests/test_refgpu_shim.py)
double refgpu_cv_determinant6(const float A[36]) { cv::Matx66f m; memcpy(m.val, A, sizeof(m.val)); return cv::determinant(m); }
void refgpu_cv_solve6(const float A[36], const float b[6], float x[6]) {
    cv::Matx66f m; cv::Vec6f bv, xv; memcpy(m.val, A, sizeof(m.val)); memcpy(bv.val, b, sizeof(bv.val));
    cv::solve(m, bv, xv, cv::DECOMP_SVD); memcpy(x, xv.val, sizeof(xv.val));
}
void refgpu_cv_affine(const float rvec[3], const float t[3], float out16[16]) {
    Vec3f rv(rvec), tv(t); Affine3f a(rv, tv); memcpy(out16, a.matrix.val, sizeof(a.matrix.val));
}
void refgpu_cv_affine_mul(const float a16[16], const float b16[16], float out16[16]) {
    cv::Matx44f ma, mb; memcpy(ma.val, a16, 64); memcpy(mb.val, b16, 64);
    Affine3f r = Affine3f(ma) * Affine3f(mb); memcpy(out16, r.matrix.val, 64);
}
void refgpu_cv_affine_inv(const float a16[16], float out16[16]) {
    cv::Matx44f ma; memcpy(ma.val, a16, 64); Affine3f r = Affine3f(ma).inv(); memcpy(out16, r.matrix.val, 64);
}

}  // extern "C"
