// OpenNISource stub (reference: src/capture.cpp needs OpenNI 1.x, which this platform does not have).
#include <iostream>
#include <io/capture.hpp>

namespace tfusion {
OpenNISource::OpenNISource() : shadow_value(0), no_sample_value(0), depth_focal_length_VGA(0.f), baseline(0.f), pixelSize(0.0), max_depth(0) {}
OpenNISource::OpenNISource(int device) : OpenNISource() { open(device); }
OpenNISource::OpenNISource(const std::string& f) : OpenNISource() { open(f); }
OpenNISource::~OpenNISource() { release(); }
void OpenNISource::open(int) { std::cerr << "OpenNISource: built without OpenNI; read depth frames from files instead" << std::endl; }
void OpenNISource::open(const std::string&) { open(0); }
void OpenNISource::release() {}
bool OpenNISource::setRegistration(bool) { return false; }
}  // namespace tfusion
