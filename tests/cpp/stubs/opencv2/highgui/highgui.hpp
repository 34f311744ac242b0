// TEST INFRASTRUCTURE.  Headless stand-in for the parts of OpenCV highgui/core that the reference's viewer
// (/root/reference/apps/demo.cpp) uses, so that the UNMODIFIED demo.cpp compiles, links against this repo's libtfusion.so and
// runs on a machine without OpenCV or a display (tests/test_demo_dropin.py):
//   cv::Mat (storage only), cv::imread (binary 16-bit PGM, the format demo.cpp:93-96 reads), cv::imshow (keeps the last
//   "Scene" view for the test to read), cv::waitKey.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <tfusion/cv_compat.hpp>

#define CV_8U 0
#define CV_16U 2
#define CV_8UC4 24

namespace cv {
typedef Vec<double, 3> Vec3d;
typedef Affine3<double> Affine3d;

struct Mat {
    int rows, cols, type_;
    size_t step;
    unsigned char* data;
    std::shared_ptr<std::vector<unsigned char>> store;
    Mat() : rows(0), cols(0), type_(0), step(0), data(nullptr) {}
    static size_t elem(int type) { static const int d[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return (size_t)d[type & 7] * ((type >> 3) + 1); }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type; step = elem(type) * (size_t)c;
        store = std::make_shared<std::vector<unsigned char>>(step * (size_t)r);
        data = store->data();
    }
    template <typename T> T* ptr(int row = 0) { return reinterpret_cast<T*>(data + step * (size_t)row); }
    void convertTo(Mat& dst, int type, double alpha = 1.0) const {
        dst.create(rows, cols, type);
        if (type_ == CV_16U && type == CV_8U)
            for (int y = 0; y < rows; ++y)
                for (int x = 0; x < cols; ++x) {
                    double v = alpha * reinterpret_cast<const unsigned short*>(data + step * y)[x];
                    dst.data[dst.step * y + x] = (unsigned char)(v > 255 ? 255 : (v < 0 ? 0 : v + 0.5));
                }
    }
};

// binary PGM (P5), maxval > 255 => two bytes per sample, most significant first; anything else => empty Mat, like cv::imread
inline Mat imread(const std::string& path, int flags = 1) {
    (void)flags;
    Mat m;
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return m;
    char magic[3] = {0, 0, 0};
    int w = 0, h = 0, maxv = 0;
    if (std::fscanf(f, "%2s %d %d %d", magic, &w, &h, &maxv) == 4 && !std::strcmp(magic, "P5") && maxv > 255 && w > 0 && h > 0) {
        std::fgetc(f);
        m.create(h, w, CV_16U);
        std::vector<unsigned char> raw((size_t)w * h * 2);
        if (std::fread(raw.data(), 1, raw.size(), f) == raw.size())
            for (size_t i = 0; i < (size_t)w * h; ++i) reinterpret_cast<unsigned short*>(m.data)[i] = (unsigned short)((raw[2 * i] << 8) | raw[2 * i + 1]);
        else m = Mat();
    }
    std::fclose(f);
    return m;
}

// TFUSION_STUB_VIEW_OUT: the last image shown in the window "Scene" is written there (raw bytes), for the test to compare
inline void imshow(const std::string& name, const Mat& m) {
    const char* out = std::getenv("TFUSION_STUB_VIEW_OUT");
    if (!out || name != "Scene" || !m.data) return;
    if (FILE* f = std::fopen(out, "wb")) { std::fwrite(m.data, 1, m.step * (size_t)m.rows, f); std::fclose(f); }
}
inline int waitKey(int = 0) { return -1; }
}  // namespace cv
