// tfusion::TopFu over the C ABI (reference: src/topfu.cpp).  operator() is one call into the library; the stages, the
// pose bookkeeping on the device and the single host wait live behind tfb_process_frame_device.
#include <cstring>
#include <iostream>
#include <tfusion/topfu.hpp>
#include "detail.hpp"

namespace tfusion {

Intr::Intr() {}
Intr::Intr(float fx_, float fy_, float cx_, float cy_) : fx(fx_), fy(fy_), cx(cx_), cy(cy_) {}
Intr Intr::operator()(int level_index) const {
    int div = 1 << level_index;
    return Intr(fx / div, fy / div, cx / div, cy / div);
}
std::ostream& operator<<(std::ostream& os, const Intr& intr) {
    return os << "([f = " << intr.fx << ", " << intr.fy << "] [cp = " << intr.cx << ", " << intr.cy << "])";
}

// values: reference src/topfu.cpp:12-53
TopFuParams TopFuParams::default_params() {
    const int iters[] = {10, 5, 4, 0};
    TopFuParams p;
    p.cols = 640;
    p.rows = 480;
    p.intr = Intr(504.261f, 503.905f, 352.457f, 272.202f);
    p.volume_dims = Vec3i::all(512);
    p.volume_size = Vec3f::all(3.f);
    p.volume_pose = Affine3f().translate(Vec3f(-p.volume_size[0] / 2, -p.volume_size[1] / 2, 0.5f));
    p.bilateral_sigma_depth = 0.04f;
    p.bilateral_sigma_spatial = 4.5f;
    p.bilateral_kernel_size = 7;
    p.icp_truncate_depth_dist = 2.0f;
    p.icp_dist_thres = 0.1f;
    p.icp_angle_thres = deg2rad(30.f);
    p.icp_iter_num.assign(iters, iters + 4);
    p.tsdf_min_camera_movement = 0.f;
    p.tsdf_trunc_dist = 0.04f;
    p.tsdf_max_weight = 64;
    p.raycast_step_factor = 0.75f;
    p.gradient_delta_factor = 0.5f;
    p.light_pose = Vec3f::all(0.f);
    p.sceneParams = new SceneParams(0.02f, 100, 0.005f, 0.2f, 3.0f, false);
    return p;
}

TopFu::TopFu(const TopFuParams& params) : frame_counter_(0), params_(params), ctx_(0) { create(TopFuSceneConfig()); }
TopFu::TopFu(const TopFuParams& params, const TopFuSceneConfig& sc) : frame_counter_(0), params_(params), ctx_(0) { create(sc); }

void TopFu::create(const TopFuSceneConfig& sc) {
    scene_config_ = sc;
    icp_ = cv::Ptr<cuda::ProjectiveICP>(new cuda::ProjectiveICP());
    icp_->setDistThreshold(params_.icp_dist_thres);
    icp_->setAngleThreshold(params_.icp_angle_thres);
    icp_->setIterationsNum(params_.icp_iter_num);

    tfb_params p;
    tfb_default_params(&p);
    p.cols = params_.cols; p.rows = params_.rows;
    p.fx = params_.intr.fx; p.fy = params_.intr.fy; p.cx = params_.intr.cx; p.cy = params_.intr.cy;
    p.bilateral_sigma_depth = params_.bilateral_sigma_depth;
    p.bilateral_sigma_spatial = params_.bilateral_sigma_spatial;
    p.bilateral_kernel_size = params_.bilateral_kernel_size;
    p.icp_truncate_depth_dist = params_.icp_truncate_depth_dist;
    p.icp_dist_thres = params_.icp_dist_thres;
    p.icp_angle_thres = params_.icp_angle_thres;
    for (int i = 0; i < 4; ++i) p.icp_iters[i] = icp_->iterations()[i];
    if (params_.sceneParams) {
        const SceneParams& s = *params_.sceneParams;
        p.mu = s.mu; p.max_w = s.maxW; p.voxel_size = s.voxelSize;
        p.view_frustum_min = s.viewFrustum_min; p.view_frustum_max = s.viewFrustum_max;
        p.stop_integrating_at_max_w = s.stopIntegratingAtMaxW ? 1 : 0;
    }
    p.num_blocks = sc.num_blocks; p.num_buckets = sc.num_buckets; p.excess_size = sc.excess_size;
    p.depth_cutoff_mm = sc.depth_cutoff_mm; p.corrected_mode = sc.corrected_mode ? 1 : 0;
    p.shard_rank = sc.shard_rank; p.shard_count = sc.shard_count;
    p.defer_tail = sc.defer_tail ? (sc.eager_tail ? 2 : 1) : 0;
    p.ieee_arith = sc.ieee_arith ? 1 : 0;
    int rc = tfb_create(&p, 0, &ctx_);
    if (rc != TFB_OK) cuda::error("tfb_create failed (no CUDA device, or out of memory)", __FILE__, __LINE__, "TopFu::TopFu");
}

TopFu::~TopFu() {
    if (ctx_) tfb_destroy(ctx_);
}

const TopFuParams& TopFu::params() const { return params_; }
TopFuParams& TopFu::params() { return params_; }
const cuda::ProjectiveICP& TopFu::icp() const { return *icp_; }
cuda::ProjectiveICP& TopFu::icp() { return *icp_; }

void TopFu::reset() {
    if (frame_counter_) std::cout << "Reset" << std::endl;
    frame_counter_ = 0;
    TF_CHECK(tfb_reset(ctx_));
}

Affine3f TopFu::getCameraPose(int time) const {
    float m[16];
    TF_CHECK(tfb_get_pose(ctx_, time, m));
    Affine3f a;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) a.matrix(r, c) = m[r * 4 + c];
    return a;
}

long long TopFu::voxelUpdatesLastFrame() const { return tfb_voxel_updates_last(ctx_); }

void TopFu::extractPoints(cuda::DeviceArray<float>& points4, int& count) {
    int n = 0;
    TF_CHECK(tfb_extract_points(ctx_, 0, 0, &n));
    points4.create((size_t)(n > 0 ? n : 1) * 4);
    TF_CHECK(tfb_extract_points(ctx_, points4.ptr(), n, &count));
}

void TopFu::renderPointCloud(cuda::DeviceArray<float>& points4, int& count, bool skipPoints) {
    points4.create((size_t)params_.rows * params_.cols * 4);
    TF_CHECK(tfb_render_point_cloud(ctx_, 0, skipPoints ? 1 : 0, points4.ptr(), params_.rows * params_.cols, &count));
}

void TopFu::saveScene(const std::string& path) { TF_CHECK(tfb_scene_save(ctx_, path.c_str())); }

int TopFu::streamOut(int maxBlocks) {
    int n = 0;
    TF_CHECK(tfb_stream_out(ctx_, maxBlocks, &n));
    return n;
}

int TopFu::streamIn(bool everything) {
    int n = 0;
    TF_CHECK(tfb_stream_in(ctx_, 0, everything ? 1 : 0, &n, 0));
    return n;
}

void TopFu::loadScene(const std::string& path) {
    TF_CHECK(tfb_scene_load(ctx_, path.c_str()));
    frame_counter_ = tfb_num_poses(ctx_);
}

bool TopFu::operator()(const cuda::Depth& depth, const cuda::Image&) {
    if (depth.rows() != params_.rows || depth.cols() != params_.cols)
        cuda::error("depth frame size differs from TopFuParams", __FILE__, __LINE__, "TopFu::operator()");
    // ICP settings may have been changed through icp() since the last frame
    TF_CHECK(tfb_set_icp_params(ctx_, icp_->getDistThreshold(), icp_->getAngleThreshold(), &icp_->iterations()[0]));
    const unsigned short* src = depth.ptr();
    if (depth.step() != (size_t)depth.colsBytes()) {  // caller wrapped pitched memory: densify once
        dense_depth_.create(depth.rows(), depth.cols());
        TF_CHECK(tfb_memcpy_2d(ctx_, dense_depth_.ptr(), dense_depth_.step(), depth.ptr(), depth.step(), (size_t)depth.colsBytes(),
                               depth.rows(), 2));
        src = dense_depth_.ptr();
    }
    int ok = 0;
    TF_CHECK(tfb_process_frame_device(ctx_, src, &ok));
    return finish_frame(ok);
}

bool TopFu::operator()(const io::HostFrame& depth) {
    if (!depth.data || depth.rows != params_.rows || depth.cols != params_.cols)
        cuda::error("depth frame size differs from TopFuParams", __FILE__, __LINE__, "TopFu::operator()");
    TF_CHECK(tfb_set_icp_params(ctx_, icp_->getDistThreshold(), icp_->getAngleThreshold(), &icp_->iterations()[0]));
    int ok = 0;
    TF_CHECK(tfb_process_frame(ctx_, depth.data, depth.step, &ok));
    return finish_frame(ok);
}

// what operator() does once the library has the verdict (topfu.cpp:252, 263-264)
bool TopFu::finish_frame(int ok) {
    if (scene_config_.print_pose && ok) {
        Affine3f p = getCameraPose();
        std::cout << "pose:" << std::endl;
        for (int r = 0; r < 4; ++r) std::cout << p.matrix(r, 0) << ", " << p.matrix(r, 1) << ", " << p.matrix(r, 2) << ", " << p.matrix(r, 3) << "\n";
    }
    if (!ok) {  // the library has already reset the scene and the pose list (topfu.cpp:263-264)
        if (frame_counter_) std::cout << "Reset" << std::endl;
        frame_counter_ = 0;
        return false;
    }
    ++frame_counter_;
    return true;
}

void TopFu::renderImage(cuda::image4u& image) {
    image.create(params_.rows, params_.cols);
    TF_CHECK(tfb_render_image(ctx_, 0, (uint8_t*)image.ptr()));
    TF_CHECK(tfb_sync(ctx_));
}

}  // namespace tfusion
