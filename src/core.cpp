// Device queries and wall-clock scope timers (reference: src/core.cpp, include/tfusion/types.hpp:83-104).
#include <chrono>
#include <cstdio>
#include <iostream>
#include <tfusion/topfu.hpp>
#include "detail.hpp"

namespace tfusion {

int cuda::getCudaEnabledDeviceCount() { return tfb_device_count(); }
void cuda::setDevice(int device) { TF_CHECK(tfb_set_device(device)); }

std::string cuda::getDeviceName(int device) {
    char name[256] = {0};
    TF_CHECK(tfb_device_info(device, name, sizeof(name), 0, 0, 0, 0));
    return name;
}

bool cuda::checkIfPreFermiGPU(int device) {
    int major = 0;
    TF_CHECK(tfb_device_info(device, 0, 0, &major, 0, 0, 0));
    return major < 2;
}

void cuda::printShortCudaDeviceInfo(int device) {
    char name[256] = {0};
    int major = 0, minor = 0, sms = 0;
    size_t mem = 0;
    TF_CHECK(tfb_device_info(device, name, sizeof(name), &major, &minor, &sms, &mem));
    std::printf("[%s] Device %d: \"%s\"  %.0fMb, sm_%d%d, %d SMs\n", tfb_version(), device, name, mem / 1048576.0, major, minor, sms);
}
void cuda::printCudaDeviceInfo(int device) { printShortCudaDeviceInfo(device); }

static double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

ScopeTime::ScopeTime(const char* name_) : name(name_), start(now_ms()) {}
ScopeTime::~ScopeTime() { std::cout << "Time(" << name << ") = " << (now_ms() - start) << "ms" << std::endl; }

SampledScopeTime::SampledScopeTime(double& time_ms) : time_ms_(time_ms), start(now_ms()) {}
double SampledScopeTime::getTime() { return now_ms() - start; }
SampledScopeTime::~SampledScopeTime() {
    static int i_ = 0;
    time_ms_ += getTime();
    if (i_ % EACH == 0 && i_) {
        std::cout << "Average frame time = " << time_ms_ / EACH << "ms ( " << 1000.f * EACH / time_ms_ << "fps )" << std::endl;
        time_ms_ = 0.0;
    }
    ++i_;
}

}  // namespace tfusion
