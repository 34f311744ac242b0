// TEST INFRASTRUCTURE (see tfo_types.h).  Adaptors that route the oracle's element-function
// interface onto the reference's OWN _CPU_AND_GPU_CODE_ functions, included where they lie
// under /root/reference (never copied).  Built by oracle/Makefile into oracle/_ref/ with
// nvcc acting as a host compiler (the headers declare __global__ helpers that g++ alone
// rejects); nothing here runs on a GPU.  Valid only at the reference's compile-time hash
// geometry (VoxelBlockHash.hpp:10-18).
#include <limits>
#include <cmath>
#include <cuda_fp16.h>

#include "tfusion/cuda/VoxelTypes.hpp"
#include "tfusion/cuda/SceneReconstructionEngine.hpp"
#include "tfusion/cuda/VisualisationEngine_Shared.hpp"

#include "tfo_types.h"

#ifndef __CUDA_ARCH__

static_assert(sizeof(::HashEntry) == sizeof(tfo::HashEntry), "hash entry layout");
static_assert(sizeof(::Voxel_s) == sizeof(tfo::Voxel), "voxel layout");
static_assert(sizeof(::RenderingBlock) == sizeof(tfo::RenderTile), "tile layout");

namespace tfo {
namespace k {

const char* impl_name() { return "reference"; }

static inline Matrix4f M(const float m[16]) { return Matrix4f(m); }
static inline Vector4f V4(const float v[4]) { return Vector4f(v[0], v[1], v[2], v[3]); }
static inline void check_geom(const HashGeom& g) {
    if (g.num_buckets != SDF_BUCKET_NUM || g.excess_size != SDF_EXCESS_LIST_SIZE || g.hash_mask != SDF_HASH_MASK) {
        fprintf(stderr, "tfo_ref: hash geometry differs from the reference's compile-time constants\n");
        abort();
    }
}

bool mat4_inv(const float in[16], float out[16]) {
    Matrix4f a(in), b;
    bool ok = a.inv(b);
    memcpy(out, b.m, sizeof(float) * 16);
    return ok;
}

void mark_pixel(uint8_t* alloc_type, uint8_t* vis_type, int x, int y, int16_t* block_coords, const float* dists,
                const float inv_m[16], const float inv_proj[4], float mu, int w, int h, float one_over_block_m,
                const HashEntry* table, float vf_min, float vf_max, const HashGeom& g) {
    check_geom(g);
    buildHashAllocAndVisibleTypePP(alloc_type, vis_type, x, y, (Vector4s*)block_coords, dists, M(inv_m), V4(inv_proj),
                                   mu, Vector2i(w, h), one_over_block_m, (const ::HashEntry*)table, vf_min, vf_max);
}

bool block_visible(const int16_t pos[3], const float m[16], const float proj[4], float voxel_size, int w, int h) {
    bool vis = false, enl = false;
    Vector3s p(pos[0], pos[1], pos[2]);
    checkBlockVisibility<false>(vis, enl, p, M(m), V4(proj), voxel_size, Vector2i(w, h));
    return vis;
}

void update_voxel(Voxel& v, const float pt[4], const float m[16], const float proj[4], float mu, int max_w,
                  const float* dists, int w, int h) {
    computeUpdatedVoxelDepthInfo<Voxel_s>(*(Voxel_s*)&v, V4(pt), M(m), V4(proj), mu, max_w, dists, Vector2i(w, h));
}

bool project_block(const int16_t pos[3], const float m[16], const float proj[4], int w, int h, float voxel_size,
                   int ul[2], int lr[2], float z[2]) {
    Vector2i a, b; Vector2f zr;
    bool ok = ProjectSingleBlock(Vector3s(pos[0], pos[1], pos[2]), M(m), V4(proj), Vector2i(w, h), voxel_size, a, b, zr);
    ul[0] = a.x; ul[1] = a.y; lr[0] = b.x; lr[1] = b.y; z[0] = zr.x; z[1] = zr.y;
    return ok;
}

int split_tiles(RenderTile* list, int offset, const int ul[2], const int lr[2], const float z[2]) {
    // CreateRenderingBlocks takes the offset by value; recompute the count it appends
    Vector2i a(ul[0], ul[1]), b(lr[0], lr[1]); Vector2f zr(z[0], z[1]);
    CreateRenderingBlocks((RenderingBlock*)list, offset, a, b, zr);
    int ny = (int)ceil((float)(1 + lr[1] - ul[1]) / renderingBlockSizeY);
    int nx = (int)ceil((float)(1 + lr[0] - ul[0]) / renderingBlockSizeX);
    int n = offset + nx * ny;
    return n > MAX_RENDERING_BLOCKS ? MAX_RENDERING_BLOCKS : n;
}

bool cast_ray(float out[4], uint8_t* vis_type, int x, int y, const Voxel* voxels, const HashEntry* table,
              const float inv_m[16], const float inv_proj[4], float one_over_voxel, float mu, const float minmax[2],
              const HashGeom& g) {
    check_geom(g);
    Vector4f o; Vector2f mm(minmax[0], minmax[1]);
    bool hit;
    if (vis_type)
        hit = castRay<Voxel_s, tfusion::VoxelBlockHash, true>(o, vis_type, x, y, (const Voxel_s*)voxels, (const ::HashEntry*)table,
                                                              M(inv_m), V4(inv_proj), one_over_voxel, mu, mm);
    else
        hit = castRay<Voxel_s, tfusion::VoxelBlockHash, false>(o, vis_type, x, y, (const Voxel_s*)voxels, (const ::HashEntry*)table,
                                                               M(inv_m), V4(inv_proj), one_over_voxel, mu, mm);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
    return hit;
}

void icp_map_pixel(float* points, float* normals, const float* ray, int w, int h, int x, int y, float voxel_size,
                   const float light[3]) {
    Vector2i sz(w, h);
    Vector3f l(light[0], light[1], light[2]);
    processPixelICP<false, false>((Vector4f*)points, (Vector4f*)normals, (const Vector4f*)ray, sz, x, y, voxel_size, l);
    // host build of processPixelICP writes 0 where the device build writes NaN
    // (VisualisationEngine_Shared.hpp:383-391); restore the device behaviour.
    int id = x + y * w;
    if (points[4 * id + 3] == 0.0f) {
        float q = std::numeric_limits<float>::quiet_NaN();
        for (int i = 0; i < 4; ++i) points[4 * id + i] = normals[4 * id + i] = q;
    }
}

void shade_pixel_grey(uint8_t out[4], const float ray[4], const Voxel* voxels, const HashEntry* table, const float light[3],
                      const HashGeom& g) {
    check_geom(g);
    Vector4u px;
    Vector3f pt(ray[0], ray[1], ray[2]), l(light[0], light[1], light[2]);
    processPixelGrey<Voxel_s, tfusion::VoxelBlockHash>(px, pt, ray[3] > 0, (const Voxel_s*)voxels, (const ::HashEntry*)table, l);
    out[0] = px.x; out[1] = px.y; out[2] = px.z; out[3] = px.w;
}

}  // namespace k
}  // namespace tfo

#endif  // !__CUDA_ARCH__
