// TEST INFRASTRUCTURE.  Headless stand-in for the cv::viz names apps/demo.cpp uses (demo.cpp:13-38,86,111-116,130).
//   TFUSION_STUB_FRAMES    the window "is closed" after this many frames (wasStopped), default 5
//   TFUSION_STUB_POSE_LOG  every pose handed to setViewerPose is appended there (16 floats per line)
#pragma once
#include <cstdio>
#include <cstdlib>
#include <string>
#include <opencv2/highgui/highgui.hpp>

namespace cv { namespace viz {
struct KeyboardEvent { enum Action { KEY_UP = 0, KEY_DOWN = 1 }; Action action; unsigned char code; };
struct Color { static Color apricot() { return Color(); } };
struct Widget {};
struct WCube : Widget { WCube(const Vec3d&, const Vec3d&, bool, const Color&) {} };
struct WCoordinateSystem : Widget { explicit WCoordinateSystem(double) {} };
inline bool isNan(double x) { return x != x; }
class Viz3d {
public:
    typedef void (*KeyboardCallback)(const KeyboardEvent&, void*);
    Viz3d() : polls_(0), limit_(5) { if (const char* e = std::getenv("TFUSION_STUB_FRAMES")) limit_ = std::atoi(e); }
    void showWidget(const std::string&, const Widget&, const Affine3d& = Affine3d()) {}
    void registerKeyboardCallback(KeyboardCallback, void*) {}
    bool wasStopped() { return polls_++ >= limit_; }
    void spinOnce(int = 1, bool = false) {}
    void setViewerPose(const Affine3d& pose) {
        const char* log = std::getenv("TFUSION_STUB_POSE_LOG");
        if (!log) return;
        if (FILE* f = std::fopen(log, "a")) {
            for (int i = 0; i < 16; ++i) std::fprintf(f, "%.9g%c", pose.matrix.val[i], i == 15 ? '\n' : ' ');
            std::fclose(f);
        }
    }
private:
    int polls_, limit_;
};
} }  // namespace cv::viz
