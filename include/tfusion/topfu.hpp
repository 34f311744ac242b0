// tfusion::TopFu — the public entry point of the library (reference: include/tfusion/topfu.hpp:16-110).
// Same class, same method names and semantics, same TopFuParams field order, so apps/demo.cpp compiles against it
// unchanged.  Everything behind operator() runs in the sm_100a library through the C ABI of tfusion_b200.h.
#pragma once
#include <string>
#include <vector>
#include <tfusion/types.hpp>
#include <tfusion/SceneParams.hpp>
#include <tfusion/cuda/projective_icp.hpp>
#include <tfusion/cuda/imgproc.hpp>
#include <io/frame_ring.hpp>

struct tfb_ctx;

namespace tfusion {
namespace cuda {
KF_EXPORTS int getCudaEnabledDeviceCount();
KF_EXPORTS void setDevice(int device);
KF_EXPORTS std::string getDeviceName(int device);
KF_EXPORTS bool checkIfPreFermiGPU(int device);
KF_EXPORTS void printCudaDeviceInfo(int device);
KF_EXPORTS void printShortCudaDeviceInfo(int device);
}  // namespace cuda

struct KF_EXPORTS TopFuParams {
    static TopFuParams default_params();

    int cols;  // pixels
    int rows;  // pixels
    Intr intr;

    Vec3i volume_dims;     // unused by the voxel-hash scene (kept for layout)
    Vec3f volume_size;     // metres
    Affine3f volume_pose;  // metres

    float bilateral_sigma_depth;    // metres
    float bilateral_sigma_spatial;  // pixels
    int bilateral_kernel_size;      // pixels

    float icp_truncate_depth_dist;  // metres
    float icp_dist_thres;           // metres
    float icp_angle_thres;          // radians
    std::vector<int> icp_iter_num;  // iterations for level index 0,1,..,3

    float tsdf_min_camera_movement;  // unused
    float tsdf_trunc_dist;           // unused
    int tsdf_max_weight;             // unused

    float raycast_step_factor;    // unused
    float gradient_delta_factor;  // unused

    Vec3f light_pose;  // unused

    SceneParams* sceneParams;
};

// Scene geometry the reference fixes with #defines (VoxelBlockHash.hpp:10-18) and behaviour switches that have no slot
// in TopFuParams.  Optional second constructor argument; the defaults reproduce the reference.
struct KF_EXPORTS TopFuSceneConfig {
    int num_blocks = 0x10000;
    int num_buckets = 0x100000;
    int excess_size = 0x20000;
    int depth_cutoff_mm = 2047;
    bool corrected_mode = false;  // true: model maps are moved into the camera frame (fixes SURVEY.md F1)
    int shard_rank = 0, shard_count = 1;
    bool print_pose = false;      // the reference prints the pose every frame (topfu.cpp:252)
    bool defer_tail = true;       // operator() returns once the pose is known; integration / raycast of that frame run beside the
                                  // next frame's preprocessing (or before anything looks at the scene).  Results are identical.
    bool eager_tail = false;      // with defer_tail: those stages are enqueued behind the frame's ICP in the same call (which still
                                  // returns when the pose is known) — the GPU runs them while the host fetches the next frame.
                                  // For loops that hand over host frames and do not look at the scene in between (io::FrameRing).
    bool ieee_arith = false;      // TSDF integration: false = the arithmetic of the reference's GPU build (bit-identical voxels to the
                                  // reference on the same GPU); true = IEEE, bit-identical to a host compile of the same function
};

class KF_EXPORTS TopFu {
public:
    typedef cv::Ptr<TopFu> Ptr;

    TopFu(const TopFuParams& params);
    TopFu(const TopFuParams& params, const TopFuSceneConfig& scene_config);
    ~TopFu();

    const TopFuParams& params() const;
    TopFuParams& params();

    const cuda::ProjectiveICP& icp() const;
    cuda::ProjectiveICP& icp();

    void reset();

    bool operator()(const cuda::Depth& dpeth, const cuda::Image& image = cuda::Image());
    // addition: the frame still in (preferably page-locked) host memory, e.g. a slot of io::FrameRing.  The upload is an
    // asynchronous copy on the library's second stream, beside the previous frame's integration and raycast, instead of
    // the synchronous Depth::upload in front of the call; the buffer may be reused as soon as the call returns.
    bool operator()(const io::HostFrame& depth);

    void renderImage(cuda::image4u& image);

    Affine3f getCameraPose(int time = -1) const;

    // additions (not in the reference): the C context, for callers that want stage-level access or timings
    tfb_ctx* context() const { return ctx_; }
    long long voxelUpdatesLastFrame() const;
    // the reconstruction as surface points (what apps/demo.cpp's take_cloud stub would fetch): x, y, z, 1 per point, world metres
    void extractPoints(cuda::DeviceArray<float>& points4, int& count);
    // the cloud of the current view (renderPointCloud_device, cuda/VisualisationHelper.hpp:150-198 of the reference)
    void renderPointCloud(cuda::DeviceArray<float>& points4, int& count, bool skipPoints = false);
    void saveScene(const std::string& path);
    void loadScene(const std::string& path);
    // block streaming (the reference's dormant swapping, GlobalCache.hpp / Scene(..., useSwapping)): blocks that left the enlarged
    // frustum go to host memory, blocks the (predicted) pose looks at come back; returns the number of blocks moved
    int streamOut(int maxBlocks = 0);
    int streamIn(bool everything = false);

private:
    TopFu(const TopFu&);
    TopFu& operator=(const TopFu&);
    void create(const TopFuSceneConfig& sc);

    int frame_counter_;
    TopFuParams params_;
    TopFuSceneConfig scene_config_;
    cv::Ptr<cuda::ProjectiveICP> icp_;
    tfb_ctx* ctx_;
    cuda::Depth dense_depth_;  // staging when the caller's frame is pitched
    bool finish_frame(int ok);
};

}  // namespace tfusion
