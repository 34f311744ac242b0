// Headless twin of the reference's apps/demo.cpp: the same sequence of public-API calls (setDevice,
// printShortCudaDeviceInfo, default_params, TopFu::Ptr, Depth::upload, operator() inside SampledScopeTime,
// renderImage + download, getCameraPose) without highgui / viz / OpenNI.  Frames are 16-bit PGMs named %04d.pgm
// (what demo.cpp reads), e.g. written by `python -m topfusion_b200.synth_cli`.
//
//   demo_synth <frame_dir> [n_frames] [--corrected] [--out view.pgm] [--ring]
// --ring: the input side as a decode-ahead ring of page-locked frames (io::FrameRing) handed to operator() in host memory,
// instead of demo.cpp's synchronous imread + upload per frame; same poses, same view.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>
#include <io/capture.hpp>
#include <tfusion/topfu.hpp>

using namespace tfusion;

static bool read_pgm16(const std::string& path, std::vector<unsigned short>& px, int& w, int& h) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    int maxv = 0;
    if (std::fscanf(f, "P5 %d %d %d", &w, &h, &maxv) != 3) { std::fclose(f); return false; }
    std::fgetc(f);
    px.resize((size_t)w * h);
    std::vector<unsigned char> raw((size_t)w * h * 2);
    size_t n = std::fread(raw.data(), 1, raw.size(), f);
    std::fclose(f);
    if (n != raw.size()) return false;
    for (size_t i = 0; i < px.size(); ++i) px[i] = (unsigned short)((raw[2 * i] << 8) | raw[2 * i + 1]);
    return true;
}

struct TopFuApp {
    TopFuApp(bool corrected, bool eager_tail) {
        TopFuParams params = TopFuParams::default_params();
        TopFuSceneConfig sc;
        sc.corrected_mode = corrected;
        sc.eager_tail = eager_tail;   // --ring: host frames, nothing between two calls waits for the device
        topfu_ = TopFu::Ptr(new TopFu(params, sc));
    }

    void show_raycasted(TopFu& topfu) {
        topfu.renderImage(view_device_);
        view_host_.resize((size_t)view_device_.rows() * view_device_.cols());
        view_device_.download(view_host_.data(), view_device_.cols() * sizeof(Vector4u));
    }

    bool execute(const std::string& dir, int n_frames) {
        TopFu& topfu = *topfu_;
        double time_ms = 0;
        std::vector<unsigned short> depth;
        for (int i = 0; i < n_frames; ++i) {
            char name[64];
            std::snprintf(name, sizeof(name), "/%04d.pgm", i);
            int w = 0, h = 0;
            if (!read_pgm16(dir + name, depth, w, h)) return std::cout << "Can't grab " << dir << name << std::endl, false;
            depth_device_.upload(depth.data(), w * sizeof(unsigned short), h, w);
            bool has_image;
            {
                SampledScopeTime fps(time_ms); (void)fps;
                has_image = topfu(depth_device_);
            }
            if (has_image) show_raycasted(topfu);
            Affine3f pose = topfu.getCameraPose();
            std::printf("frame %3d ok=%d t=(% .5f % .5f % .5f) voxel-updates=%lld\n", i, (int)has_image, pose.matrix(0, 3),
                        pose.matrix(1, 3), pose.matrix(2, 3), topfu.voxelUpdatesLastFrame());
        }
        return true;
    }

    // the same loop with the frames decoded ahead into page-locked memory and uploaded asynchronously inside operator()
    bool execute_ring(const std::string& dir, int n_frames) {
        TopFu& topfu = *topfu_;
        io::FrameRing ring(dir, 4, 0, n_frames);
        const auto t0 = std::chrono::steady_clock::now();
        int i = 0;
        while (const io::HostFrame* f = ring.next()) {
            const bool has_image = topfu(*f);
            const int index = f->index;
            ring.release(f);   // the slot may be refilled from here on
            if (has_image) show_raycasted(topfu);
            Affine3f pose = topfu.getCameraPose();
            std::printf("frame %3d ok=%d t=(% .5f % .5f % .5f) voxel-updates=%lld\n", index, (int)has_image, pose.matrix(0, 3),
                        pose.matrix(1, 3), pose.matrix(2, 3), topfu.voxelUpdatesLastFrame());
            ++i;
        }
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (!ring.error().empty()) return std::cout << "Can't grab: " << ring.error() << std::endl, false;
        std::printf("ring: %d frames in %.1f ms (decode + upload + operator() + render), consumer waited %.1f ms for the decoder, %s memory\n",
                    i, ms, ring.consumerWaitMs(), ring.pinned() ? "page-locked" : "pageable");
        return i == n_frames;
    }

    // what demo.cpp's take_cloud stub (apps/demo.cpp:70-77) is there for: the reconstruction as a point cloud
    void take_cloud() {
        cuda::DeviceArray<float> cloud;
        int n = 0;
        topfu_->extractPoints(cloud, n);
        std::printf("cloud: %d surface points\n", n);
    }

    void save_view(const std::string& path) {
        if (view_host_.empty()) return;
        FILE* f = std::fopen(path.c_str(), "wb");
        if (!f) return;
        std::fprintf(f, "P5\n%d %d\n255\n", view_device_.cols(), view_device_.rows());
        for (const Vector4u& p : view_host_) std::fputc(p.x, f);
        std::fclose(f);
    }

    TopFu::Ptr topfu_;
    std::vector<Vector4u> view_host_;
    cuda::image4u view_device_;
    cuda::Depth depth_device_;
};

int main(int argc, char* argv[]) {
    if (argc < 2) return std::cout << "usage: demo_synth <frame_dir> [n_frames] [--corrected] [--out view.pgm] [--ring]" << std::endl, 2;
    int n = 20;
    bool corrected = false, use_ring = false;
    std::string out;
    for (int i = 2; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--corrected")) corrected = true;
        else if (!std::strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (!std::strcmp(argv[i], "--ring")) use_ring = true;
        else n = std::atoi(argv[i]);
    }
    int device = 0;
    cuda::setDevice(device);
    cuda::printShortCudaDeviceInfo(device);
    if (cuda::checkIfPreFermiGPU(device)) return std::cout << "pre-Fermi GPUs are not supported" << std::endl, 1;

    OpenNISource capture;  // kept for call-sequence parity with demo.cpp; frames come from files
    (void)capture;
    TopFuApp app(corrected, use_ring);
    bool ok = use_ring ? app.execute_ring(argv[1], n) : app.execute(argv[1], n);
    if (!out.empty()) app.save_view(out);
    if (ok) app.take_cloud();
    return ok ? 0 : 1;
}
