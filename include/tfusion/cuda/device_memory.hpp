// Reference-counted device buffers (reference: include/tfusion/cuda/device_memory.hpp, src/device_memory.cpp).
// Copying shares the allocation; the last owner frees it.  2-D buffers are DENSE here (step == colsBytes): the
// reference allocates with cudaMallocPitch but then indexes its images densely.
#pragma once
#include <tfusion/exports.hpp>
#include <tfusion/cuda/kernel_containers.hpp>

namespace tfusion {
namespace cuda {

KF_EXPORTS void error(const char* error_string, const char* file, const int line, const char* func = "");

class KF_EXPORTS DeviceMemory {
public:
    DeviceMemory();
    ~DeviceMemory();
    DeviceMemory(size_t sizeBytes_arg);
    DeviceMemory(void* ptr_arg, size_t sizeBytes_arg);  // user memory, never freed here
    DeviceMemory(const DeviceMemory& other_arg);
    DeviceMemory& operator=(const DeviceMemory& other_arg);

    void create(size_t sizeBytes_arg);
    void release();
    void copyTo(DeviceMemory& other) const;
    void upload(const void* host_ptr_arg, size_t sizeBytes_arg);
    void download(void* host_ptr_arg) const;
    void swap(DeviceMemory& other_arg);
    template <class T> T* ptr() { return (T*)data_; }
    template <class T> const T* ptr() const { return (const T*)data_; }
    template <class U> operator PtrSz<U>() const { return PtrSz<U>((U*)data_, sizeBytes_ / sizeof(U)); }
    bool empty() const;
    size_t sizeBytes() const;

private:
    void* data_;
    size_t sizeBytes_;
    int* refcount_;
};

class KF_EXPORTS DeviceMemory2D {
public:
    DeviceMemory2D();
    ~DeviceMemory2D();
    DeviceMemory2D(int rows_arg, int colsBytes_arg);
    DeviceMemory2D(int rows_arg, int colsBytes_arg, void* data_arg, size_t step_arg);  // user memory
    DeviceMemory2D(const DeviceMemory2D& other_arg);
    DeviceMemory2D& operator=(const DeviceMemory2D& other_arg);

    void create(int rows_arg, int colsBytes_arg);
    void release();
    void copyTo(DeviceMemory2D& other) const;
    void upload(const void* host_ptr_arg, size_t host_step_arg, int rows_arg, int colsBytes_arg);
    void download(void* host_ptr_arg, size_t host_step_arg) const;
    void swap(DeviceMemory2D& other_arg);
    template <class T> T* ptr(int y_arg = 0) { return (T*)((char*)data_ + y_arg * step_); }
    template <class T> const T* ptr(int y_arg = 0) const { return (const T*)((const char*)data_ + y_arg * step_); }
    template <class U> operator PtrStep<U>() const { return PtrStep<U>((U*)data_, step_); }
    template <class U> operator PtrStepSz<U>() const { return PtrStepSz<U>(rows_, colsBytes_ / sizeof(U), (U*)data_, step_); }
    bool empty() const;
    int colsBytes() const;
    int rows() const;
    size_t step() const;

private:
    void* data_;
    size_t step_;
    int colsBytes_;
    int rows_;
    int* refcount_;
};

}  // namespace cuda
namespace device {
using tfusion::cuda::DeviceMemory;
using tfusion::cuda::DeviceMemory2D;
}
}  // namespace tfusion
