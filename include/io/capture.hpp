// Depth-sensor source (reference: include/io/capture.hpp, src/capture.cpp — OpenNI 1.x).  OpenNI is not available;
// this keeps the type so apps/demo.cpp links, and open() fails cleanly.  Frames normally come from files.
#pragma once
#include <string>
#include <tfusion/exports.hpp>

namespace tfusion {
struct KF_EXPORTS OpenNISource {
    OpenNISource();
    OpenNISource(int device);
    OpenNISource(const std::string& oni_filename);
    void open(int device);
    void open(const std::string& oni_filename);
    void release();
    ~OpenNISource();
    bool setRegistration(bool value = false);
    bool isOpen() const { return false; }
    int shadow_value, no_sample_value;
    float depth_focal_length_VGA;
    float baseline;  // mm
    double pixelSize;  // mm
    unsigned short max_depth;  // mm
};
}  // namespace tfusion
