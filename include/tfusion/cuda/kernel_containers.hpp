// Plain views handed to kernels: pointer, pointer+size, pointer+row stride (+ dims).  Same names and members as the
// reference's include/tfusion/cuda/kernel_containers.hpp so client code keeps compiling.
#pragma once
#include <cstddef>

#if defined(__CUDACC__)
#define __tf_hdevice__ __host__ __device__ __forceinline__
#define __tf_device__ __device__ __forceinline__
#else
#define __tf_hdevice__
#define __tf_device__
#endif

namespace tfusion {
namespace cuda {

template <typename T> struct DevPtr {
    typedef T elem_type;
    static const size_t elem_size = sizeof(T);
    T* data;
    __tf_hdevice__ DevPtr() : data(0) {}
    __tf_hdevice__ DevPtr(T* p) : data(p) {}
    __tf_hdevice__ size_t elemSize() const { return elem_size; }
    __tf_hdevice__ operator T*() { return data; }
    __tf_hdevice__ operator const T*() const { return data; }
};

template <typename T> struct PtrSz : public DevPtr<T> {
    size_t size;
    __tf_hdevice__ PtrSz() : size(0) {}
    __tf_hdevice__ PtrSz(T* p, size_t n) : DevPtr<T>(p), size(n) {}
};

template <typename T> struct PtrStep : public DevPtr<T> {
    size_t step;  // bytes between consecutive rows
    __tf_hdevice__ PtrStep() : step(0) {}
    __tf_hdevice__ PtrStep(T* p, size_t s) : DevPtr<T>(p), step(s) {}
    __tf_hdevice__ T* ptr(int y = 0) { return (T*)((char*)DevPtr<T>::data + y * step); }
    __tf_hdevice__ const T* ptr(int y = 0) const { return (const T*)((const char*)DevPtr<T>::data + y * step); }
    __tf_hdevice__ T& operator()(int y, int x) { return ptr(y)[x]; }
    __tf_hdevice__ const T& operator()(int y, int x) const { return ptr(y)[x]; }
};

template <typename T> struct PtrStepSz : public PtrStep<T> {
    int cols, rows;
    __tf_hdevice__ PtrStepSz() : cols(0), rows(0) {}
    __tf_hdevice__ PtrStepSz(int r, int c, T* p, size_t s) : PtrStep<T>(p, s), cols(c), rows(r) {}
};

}  // namespace cuda
namespace device {
using tfusion::cuda::PtrSz;
using tfusion::cuda::PtrStep;
using tfusion::cuda::PtrStepSz;
}
}  // namespace tfusion
namespace tf = tfusion;
