// TEST INFRASTRUCTURE — CPU oracle for the topfusion per-frame hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.  The product (topfusion_b200/) never links or calls it.
//
// Parity status: the reference ships no golden vectors, so the oracle is
// pinned against the reference's own host-compilable per-pixel / per-voxel functions
// (oracle/_ref, built from /root/reference by oracle/Makefile) and against cv2 for the
// OpenCV calls the reference makes; device-only stages (imgproc.cu, proj_icp.cu) are a
// line-by-line restatement and are "parity unpinned" beyond that.
#pragma once
#include <cstdint>
#include <cstddef>

extern "C" {

// Mirrors tfusion::TopFuParams + SceneParams (/root/reference/tfusion/src/topfu.cpp:12-53,
// include/tfusion/SceneParams.hpp:45-53) plus the hash geometry the reference hard-codes
// (include/tfusion/cuda/VoxelBlockHash.hpp:10-18).
typedef struct tfo_params {
    int32_t cols, rows;
    float fx, fy, cx, cy;
    float bilateral_sigma_depth;    // metres
    float bilateral_sigma_spatial;  // pixels
    int32_t bilateral_kernel_size;
    float icp_truncate_depth_dist;  // metres, <=0 disables
    float icp_dist_thres;           // metres
    float icp_angle_thres;          // radians
    int32_t icp_iters[4];           // level 0..3
    float mu;
    int32_t max_w;
    float voxel_size;
    float view_frustum_min, view_frustum_max;
    int32_t stop_integrating_at_max_w;
    int32_t num_blocks;             // SDF_LOCAL_BLOCK_NUM
    int32_t num_buckets;            // SDF_BUCKET_NUM (power of two)
    int32_t excess_size;            // SDF_EXCESS_LIST_SIZE
    int32_t depth_cutoff_mm;        // imgproc.cu:277 hard-codes 2047
    int32_t corrected_mode;         // 0 = reference behaviour, 1 = maps moved to the camera frame
    int32_t shard_rank, shard_count;
} tfo_params;

}  // extern "C"

namespace tfo {

// include/tfusion/cuda/VoxelBlockHash.hpp:32-44
struct HashEntry {
    int16_t pos[3];
    int16_t pad_;
    int32_t offset;
    int32_t ptr;
};
static_assert(sizeof(HashEntry) == 16, "HashEntry layout");

// include/tfusion/cuda/VoxelTypes.hpp:69-92 (Voxel_s)
struct Voxel {
    int16_t sdf;
    uint8_t w_depth;
    uint8_t pad_;
};
static_assert(sizeof(Voxel) == 4, "Voxel layout");

struct HashGeom {
    int num_buckets;
    int hash_mask;
    int excess_size;
    int total_entries() const { return num_buckets + excess_size; }
};

// include/tfusion/cuda/VisualisationEngine_Shared.hpp:11-15
struct RenderTile {
    int16_t ul[2];
    int16_t lr[2];
    float z[2];
};
static_assert(sizeof(RenderTile) == 16, "RenderTile layout");

enum { BLOCK = 8, BLOCK3 = 512, MINMAX_SUBSAMPLE = 8, TILE = 16, MAX_TILES = 65536 * 4 };
static const float FAR_AWAY_F = 999999.9f;
static const float VERY_CLOSE_F = 0.05f;

// Per-pixel / per-voxel / per-block element functions.  Implemented twice:
//   tfo_kernels_port.cpp — restatement (the oracle proper)
//   tfo_kernels_ref.cu   — thin adaptors onto the reference's own _CPU_AND_GPU_CODE_
//                          functions, compiled from /root/reference into oracle/_ref
// All 4x4 matrices here are column-major float[16] like the reference's Matrix4f
// (include/Matrix.hpp:24-34).
namespace k {
const char* impl_name();
bool mat4_inv(const float in[16], float out[16]);
void mark_pixel(uint8_t* alloc_type, uint8_t* vis_type, int x, int y, int16_t* block_coords,
                const float* dists, const float inv_m[16], const float inv_proj[4], float mu,
                int w, int h, float one_over_block_m, const HashEntry* table,
                float vf_min, float vf_max, const HashGeom& g);
bool block_visible(const int16_t pos[3], const float m[16], const float proj[4],
                   float voxel_size, int w, int h);
void update_voxel(Voxel& v, const float pt_model[4], const float m[16], const float proj[4],
                  float mu, int max_w, const float* dists, int w, int h);
bool project_block(const int16_t pos[3], const float m[16], const float proj[4], int w, int h,
                   float voxel_size, int ul[2], int lr[2], float z[2]);
// appends tiles at offset, returns the new offset (CreateRenderingBlocks)
int split_tiles(RenderTile* list, int offset, const int ul[2], const int lr[2], const float z[2]);
bool cast_ray(float out[4], uint8_t* vis_type, int x, int y, const Voxel* voxels,
              const HashEntry* table, const float inv_m[16], const float inv_proj[4],
              float one_over_voxel, float mu, const float minmax[2], const HashGeom& g);
void icp_map_pixel(float* points, float* normals, const float* ray, int w, int h, int x, int y,
                   float voxel_size, const float light[3]);
// processPixelGrey<Voxel_s, VoxelBlockHash>: SDF-gradient normal + Lambert shade of one raycast point (viewer path)
void shade_pixel_grey(uint8_t out[4], const float ray[4], const Voxel* voxels, const HashEntry* table,
                      const float light[3], const HashGeom& g);
}  // namespace k
}  // namespace tfo
