// REFERENCE-ARM INFRASTRUCTURE.  Replaces /root/reference/tfusion/src/cuda/texture_binder.hpp
// (legacy cudaBindTexture2D / cudaUnbindTexture, gone since CUDA 12) in the patched build.
// Same class name and constructor shapes so proj_icp.cu:411-412,436-437 compile unchanged.
// A bind = look the (pointer, extent, pitch) up in a small cache of texture objects (created
// once, never destroyed while the process lives: the kernel that samples it may still be in
// flight when the binder goes out of scope) and, if the symbol does not hold it already, one
// cudaMemcpyToSymbolAsync of 8 bytes on the legacy stream (ordered before the ICP stream, which
// is a blocking stream: projective_icp.cpp:37).  Cheaper than the bind/unbind pair it replaces.
#pragma once
#include <tfusion/cuda/device_array.hpp>
#include <safe_call.hpp>
#include <map>
#include <tuple>

namespace tfusion { namespace cuda {
class TextureBinder {
    typedef std::tuple<const void*, int, int, size_t, int> Key;
    template <class T> static cudaTextureObject_t lookup(const void* ptr, int cols, int rows, size_t step) {
        static std::map<Key, cudaTextureObject_t> cache;
        Key k(ptr, cols, rows, step, (int)sizeof(T));
        auto it = cache.find(k);
        if (it != cache.end()) return it->second;
        cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypePitch2D;
        rd.res.pitch2D.devPtr = const_cast<void*>(ptr);
        rd.res.pitch2D.desc = cudaCreateChannelDesc<T>();
        rd.res.pitch2D.width = cols; rd.res.pitch2D.height = rows; rd.res.pitch2D.pitchInBytes = step;
        cudaTextureDesc td; memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        cudaTextureObject_t obj = 0;
        cudaSafeCall(cudaCreateTextureObject(&obj, &rd, &td, 0));
        cache[k] = obj;
        return obj;
    }
    template <class T> void bind(const void* ptr, int cols, int rows, size_t step, tfcompat::TexRef<T>& sym) {
        static std::map<const void*, cudaTextureObject_t> bound;
        cudaTextureObject_t obj = lookup<T>(ptr, cols, rows, step);
        cudaTextureObject_t& cur = bound[(const void*)&sym];
        if (cur != obj) {
            tfcompat::TexRef<T> h; h.obj = obj;
            cudaSafeCall(cudaMemcpyToSymbolAsync(sym, &h, sizeof(h), 0, cudaMemcpyHostToDevice, 0));
            cur = obj;
        }
    }
public:
    template <class T> TextureBinder(const DeviceArray2D<T>& arr, tfcompat::TexRef<T>& tex) { bind<T>(arr.ptr(), arr.cols(), arr.rows(), arr.step(), tex); }
    template <class T> TextureBinder(const PtrStepSz<T>& arr, tfcompat::TexRef<T>& tex) { bind<T>(arr.data, arr.cols, arr.rows, arr.step, tex); }
    ~TextureBinder() {}
};
} namespace device { using tfusion::cuda::TextureBinder; } }
