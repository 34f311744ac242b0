// REFERENCE-ARM INFRASTRUCTURE.  Force-included (-include) in front of every patched
// reference translation unit.  Supplies what the 2014-era sources assume:
//  * <limits>, <cuda_fp16.h> (__float2half_rn / __half2float as used by device.hpp:52-60),
//  * the pre-Volta warp votes without a mask (temp_utils.hpp:585,597),
//  * the legacy texture-REFERENCE API (removed in CUDA 12) re-expressed with texture OBJECTS:
//    `texture<T,2> name;` becomes `__device__ tfcompat::TexRef<T> name;` (done by patch_ref.py),
//    `tex2D(name, x, y)` keeps its spelling and its semantics (unnormalised coordinates, point
//    sampling, clamp addressing — the legacy defaults proj_icp.cu relies on).
#pragma once
#include <limits>
#include <cstddef>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#ifdef __CUDACC__
#define __ballot(p) __ballot_sync(0xffffffffu, (p))
#define __all(p) __all_sync(0xffffffffu, (p))
#define __any(p) __any_sync(0xffffffffu, (p))
namespace tfcompat {
template <class T> struct TexRef { cudaTextureObject_t obj; };
}
template <class T> __device__ __forceinline__ T tex2D(const tfcompat::TexRef<T>& r, float x, float y) { return ::tex2D<T>(r.obj, x, y); }
#endif
