// tfusion_b200 internals shared by the translation units.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tfusion_b200.h"

namespace tfb {

// kernel ids for the per-launch CUDA-event timing (tfb_ktiming_*)
enum KernelId {
    K_BILATERAL = 0, K_DEPTH_PYR, K_POINTS_NORMALS, K_RESIZE_MAPS, K_COMPUTE_DISTS, K_TRUNCATE,
    K_ICP_BEGIN, K_ICP_L0, K_ICP_L1, K_ICP_L2, K_ICP_L3, K_POSE_UPDATE, K_POSE_SET,
    K_SET_TYPE3, K_MARK, K_ALLOC, K_VISIBLE_LIST, K_LIST_FLIP, K_INTEGRATE_BEGIN, K_INTEGRATE,
    K_MINMAX_INIT, K_EXPECTED_DEPTHS, K_RAYCAST, K_ICP_MAPS, K_RESET_SCENE, K_ICP_ALL, K_RENDER_GREY, K_RAYCAST_SHARDED, K_MODEL_MAPS, K_PYR_MAPS, K_SHARD_BARRIER, K_PUSH_FRAME, K_WAIT_FRAME, K_GATHER_FOREIGN, K_COUNT
};
constexpr int KT_MAX_EVENTS = 512;

constexpr int BLOCK = 8;          // SDF_BLOCK_SIZE, include/tfusion/cuda/VoxelBlockHash.hpp:10
constexpr int BLOCK3 = 512;       // SDF_BLOCK_SIZE3
constexpr int MINMAX_SUB = 8;     // minmaximg_subsample, VisualisationEngine_Shared.hpp:7
constexpr int MAX_LEVELS = 4;     // ProjectiveICP::MAX_PYRAMID_LEVELS
constexpr int ICP_TERMS = 27;     // proj_icp.cu:13-28
constexpr int NUM_SMS = 148;      // B200
#define TFB_FAR_AWAY 999999.9f    // VisualisationEngine_Shared.hpp:17-23
#define TFB_VERY_CLOSE 0.05f

// 16-byte hash entry (VoxelBlockHash.hpp:32-44), moved as one 128-bit word
struct __align__(16) HashEntry {
    short pos[3];
    short pad_;
    int offset;
    int ptr;
};
static_assert(sizeof(HashEntry) == 16, "HashEntry");

// Block directory (new; the reference resolves every block through the hash): a dense window of DIR_N^3 blocks centred on the
// world origin — the first camera pose — with one 8-byte cell {slot of the hash entry, ptr into the voxel pool} per block, {-1,-1}
// where no block is allocated.  It is an INDEX of the hash table, written where entries are written (allocation, scene load) and
// cleared with it (reset); the table stays the authority and serves blocks outside the window.  What it buys the raycast: a block
// lookup is ONE load instead of occupancy word -> entry -> chain, and the loads are spatially coherent — a ray's next block and
// the up to eight blocks of a trilinear read sit in the same or a neighbouring 128-byte line (cells are tiled 4x2x2 per line),
// where hashing scatters them over a 19 MB table on purpose.  256^3 cells = 128 MiB of the 180 GB: 10 m at 5 mm voxels, 4 m at 2 mm.
constexpr int DIR_BITS = 8, DIR_N = 1 << DIR_BITS, DIR_HALF = DIR_N / 2;
constexpr size_t DIR_CELLS = (size_t)DIR_N * DIR_N * DIR_N;
__host__ __device__ __forceinline__ bool dir_inside(int bx, int by, int bz) {
    return ((((unsigned)(bx + DIR_HALF)) | ((unsigned)(by + DIR_HALF)) | ((unsigned)(bz + DIR_HALF))) >> DIR_BITS) == 0u;
}
__host__ __device__ __forceinline__ unsigned int dir_index(int bx, int by, int bz) {   // only for blocks inside the window
    const unsigned int ux = (unsigned)(bx + DIR_HALF), uy = (unsigned)(by + DIR_HALF), uz = (unsigned)(bz + DIR_HALF);
    return ((uz >> 1) << (2 * DIR_BITS + 1)) | ((uy >> 1) << (DIR_BITS + 2)) | ((ux >> 2) << 4) | ((uz & 1u) << 3) | ((uy & 1u) << 2) | (ux & 3u);
}

// Voxel_s (VoxelTypes.hpp:69-92): {short sdf; uchar w_depth; pad}
struct __align__(4) Voxel {
    short sdf;
    unsigned char w_depth;
    unsigned char pad_;
};
static_assert(sizeof(Voxel) == 4, "Voxel");

// Everything a frame needs that is decided on the device lives here, so no stage has to wait
// for the host (the reference syncs 22 times per frame).
struct DevState {
    // poses, all float
    float pose_c2w[16];    // row-major, TopFu::poses_.back()
    float pose_w2c[16];    // row-major, pose.inv()
    float M_w2c[16];       // column-major Matrix4f of pose_w2c ("M_d")
    float invM_w2c[16];    // column-major cofactor inverse of M_w2c (SceneReconstructionEngine_host.cu:103-104)
    float M_c2w[16];       // column-major of pose_c2w (raycast "invM")
    float affine[16];      // ICP running estimate, row-major
    // free lists (LocalVBA.lastFreeBlockId, VoxelBlockHash.lastFreeExcessListId)
    int last_free_block;
    int last_free_excess;
    // visible list bookkeeping
    int n_visible;         // entries in the current list
    int n_next;            // entries accumulated in the next list (raycast extras, new marks, survivors)
    int n_claimed;         // slots claimed for allocation this frame
    int n_new_frame;       // blocks allocated this frame
    int n_extras;          // raycast-marked entries carried into the next frame
    int cur_list;          // which of the two list buffers is current (flipped on the device, never by the host)
    // ICP
    int icp_failed;        // sticky for the frame: |det| < 1e-15 or NaN (projective_icp.cpp:197-203)
    unsigned int icp_ticket;
    int icp_corresp;
    // statistics
    unsigned long long voxel_updates;
    unsigned int list_ticket;   // CTAs of k_visible_list that are done; the last one flips the lists
    int int_cursor;             // next visible-list position k_integrate hands out
    int shard_error;            // a cross-GPU barrier timed out
    int n_cached;               // sharded scene: foreign visible blocks copied into the local cache this frame
    int pad2_[2];
};

// payload owner of a block when the scene is sharded (new; the reference is single-GPU).  A different mix than
// hashIndex so a rank does not end up with 1/n of its buckets.  oracle/tfo_oracle.cpp `owns` is the same.
__host__ __device__ __forceinline__ int owner_rank(int bx, int by, int bz, int count) {
    if (count <= 1) return 0;
    unsigned h = ((unsigned)bx * 0x9E3779B1u) ^ ((unsigned)by * 0x85EBCA77u) ^ ((unsigned)bz * 0xC2B2AE3Du);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return (int)(h % (unsigned)count);
}

// what a rank needs of every rank (itself included) to cast rays through a sharded scene: kernel parameter, by value
struct ShardView {
    int rank, count, marks_cap, cache_cap;
    // local copies of the foreign visible blocks of this frame (k_gather_foreign): tag[slot] = epoch << 32 | cache index
    const unsigned int* cache_pool;
    const unsigned long long* cache_tag;
    unsigned int cache_epoch, pad_;
    const int4* table[TFB_MAX_SHARDS];
    const unsigned int* vba[TFB_MAX_SHARDS];
    float4* raycast[TFB_MAX_SHARDS];
    unsigned int* marks[TFB_MAX_SHARDS];
    uint16_t* frame[TFB_MAX_SHARDS];
    unsigned int* flags[TFB_MAX_SHARDS];   // SHARD_FLAG_WORDS words per rank, see below
};
// a rank's flag array: [0, TFB_MAX_SHARDS) the barrier epochs the other ranks have published here, then
constexpr int SHARD_FLAG_FRAME = TFB_MAX_SHARDS;            // sequence number of the last frame rank 0 has pushed here
constexpr int SHARD_FLAG_PUSH_TICKET = TFB_MAX_SHARDS + 1;  // local: CTAs of k_push_frame that are done
constexpr int SHARD_FLAG_RAY_TICKET = TFB_MAX_SHARDS + 2;   // local: CTAs of k_raycast_sharded that are done
constexpr int SHARD_FLAG_ACK = 2 * TFB_MAX_SHARDS;          // [ACK + r]: the last pushed frame rank r has finished reading
constexpr int SHARD_FLAG_WORDS = 4 * TFB_MAX_SHARDS;

struct LevelBuf {
    int w, h;
    uint16_t* depth;   // current filtered depth
    float4* vcurr;
    float4* ncurr;
    float4* vprev;     // model maps
    float4* nprev;
};

}  // namespace tfb

struct tfb_ctx {
    tfb_params p;
    cudaStream_t stream;
    bool own_stream;
    int device;
    char err[256];

    int total_entries;
    int hash_mask;
    int levels;

    // scene
    tfb::HashEntry* table;
    tfb::Voxel* vba;
    int* vba_free;
    int* excess_free;
    // allocation scratch
    unsigned int* claim_key;   // per slot, 0 = unclaimed
    int* claimed;              // compact list of claimed slots
    unsigned int* bucket_bits; // 1 bit per bucket: head entry allocated (empty-space skipping without touching the table)
    bool icp_fuse_type3;       // set around the frame path's ICP launch: k_icp_all's epilogue does setToType3
    bool type3_done;           // ... and the next launch_allocate skips k_set_type3
    struct HostBlockStore* store;   // blocks streamed out to the host (tfb_stream_out / tfb_stream_in), created on first use
    int2* block_dir;           // dense directory of the blocks around the origin, {slot, ptr} per block or {-1, -1}: see BlockDir
    // render state
    int* vis_type;             // per slot (reference: uchar entriesVisibleType)
    int* vis_list[2];          // double-buffered visibleEntryIDs; DevState::cur_list says which is current
    float2* minmax;            // (rows/8) x (cols/8)
    float4* raycast;           // rows x cols
    // frames
    tfb::LevelBuf lv[tfb::MAX_LEVELS];
    float* dists;              // metres image of the frame being tracked (one of dists_buf, alternating per frame)
    float* dists_buf[2];
    uint16_t* depth_in;        // device copy of the raw frame
    // ICP
    float* icp_partial;        // [ICP_TERMS][max_blocks]
    int icp_max_blocks;
    int icp_grid;              // persistent ICP grid (co-resident CTAs), sized on first use
    unsigned int icp_launches; // epoch range of the partial rows, 64 per launch
    int* icp_vlist;            // level-0 pixels that have a vertex, ascending (built inside the preprocessing launches, tfb_imgproc.cu)
    unsigned int* icp_vmask;   // one validity bit per level-0 pixel (k_pyr_maps)
    unsigned int* icp_vscan;   // [0] = length of the list
    bool vlist_built;          // at least one list has been built (tfb_export_icp_valid_list)
    bool vlist_ready;          // the list belongs to the current maps of the frame path (consumed by the next k_icp_all)
    unsigned int publish_seq;  // != 0: k_icp_all writes the state block + this number into the pinned mirror (zero-copy)
    unsigned int seq_counter;
    // state
    tfb::DevState* ds;         // device
    tfb::DevState* hs;         // pinned host mirror
    float* h_pose_stage;       // pinned, 64 floats
    float* h_icp27;            // pinned

    // host bookkeeping
    int frame_counter;
    int resets;
    float* poses;              // host history, 16 floats each
    int n_poses, cap_poses;
    long long launches;
    long long voxel_updates_last;
    long long voxel_updates_total;   // over every integration finished so far

    // timing
    bool timing;
    cudaEvent_t ev[16];
    float stage_ms[9];
    // per-launch timing (off by default): event pairs recorded around every launch, folded after the frame's sync
    bool ktiming;
    cudaEvent_t* kt_ev;          // KT_MAX_EVENTS
    int kt_n;                    // events used this frame
    unsigned char kt_id[tfb::KT_MAX_EVENTS / 2];
    double kt_ms[tfb::K_COUNT];
    long long kt_cnt[tfb::K_COUNT];
    cudaEvent_t mark_ev[8];
    void* l2_scratch;
    int l2_toggle;
    // sharding
    tfb::ShardView shard;
    tfb::ShardView* shard_dev; // device copy (kernels that take it by pointer)
    unsigned int attached;     // bit r set once rank r's buffers are attached
    unsigned int* marks;       // incoming visibility marks: [0] count, [1] pad, then 2 words per mark
    unsigned int* cache_pool;        // sharded scene: this frame's copies of the foreign visible blocks (2 KB each)
    unsigned long long* cache_tag;   // per hash slot: epoch << 32 | index into cache_pool
    unsigned int* sync_flags;  // TFB_MAX_SHARDS words: the barrier epochs the other ranks have published here
    unsigned int sync_epoch;
    unsigned int gather_epoch;
    unsigned int frame_seq;          // collective frames so far (tfb_process_frame_sharded)
    const uint16_t* push_src;        // collective frame, rank with the sensor: the frame to push beside the previous frame's tail
    int frame_stage;           // 0 idle, 1 after tfb_frame_begin, 2 after tfb_frame_raycast
    bool frame_first;
    // software pipeline of the unsharded frame: preprocessing runs on stream_pre beside the deferred tail
    cudaStream_t stream_pre;
    cudaEvent_t ev_fork, ev_join, ev_pre0, ev_pre1, ev_alloc, ev_expect;
    bool tail_pending;         // allocation .. model maps of the last tracked frame are still to be enqueued
    bool tail_inflight;        // defer_tail = 2: they were enqueued behind that frame's ICP and its counters have not been read yet
    const float* tail_dists;
};

namespace tfb {

inline int set_err(tfb_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
    if (c) {
        if (e != cudaSuccess) snprintf(c->err, sizeof(c->err), "%s: %s", what, cudaGetErrorString(e));
        else snprintf(c->err, sizeof(c->err), "%s", what);
    }
    return code;
}

#define TFB_CUDA(c, call)                                                         \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) return tfb::set_err((c), TFB_ERR_CUDA, #call, e__); \
    } while (0)

// TFB_KT(c, id): put in front of a launch; TFB_LAUNCH_CHECK(c) behind it closes the pair
#define TFB_KT(c, id)                                                                    \
    do {                                                                                 \
        if ((c)->ktiming && (c)->kt_n + 2 <= tfb::KT_MAX_EVENTS) {                       \
            (c)->kt_id[(c)->kt_n / 2] = (unsigned char)(id);                             \
            cudaEventRecord((c)->kt_ev[(c)->kt_n], (c)->stream);                         \
            (c)->kt_n |= 0x40000000;                                                     \
        }                                                                                \
    } while (0)

#define TFB_LAUNCH_CHECK(c)                                                              \
    do {                                                                                 \
        (c)->launches++;                                                                 \
        if ((c)->kt_n & 0x40000000) {                                                    \
            (c)->kt_n &= ~0x40000000;                                                    \
            cudaEventRecord((c)->kt_ev[(c)->kt_n + 1], (c)->stream);                     \
            (c)->kt_n += 2;                                                              \
        }                                                                                \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) return tfb::set_err((c), TFB_ERR_CUDA, "kernel launch", e__); \
    } while (0)

inline int div_up(int a, int b) { return (a + b - 1) / b; }

// stage launchers implemented in the .cu files ------------------------------------------------
// imgproc
int launch_compute_dists(tfb_ctx* c, const uint16_t* depth, float* dists, int w, int h);
int launch_bilateral(tfb_ctx* c, const uint16_t* src, uint16_t* dst, int w, int h, int ksz, float ss, float sd_m,
                     float trunc_m, float* dists_or_null);
int launch_truncate(tfb_ctx* c, uint16_t* depth, int w, int h, float max_dist);
int launch_depth_pyr(tfb_ctx* c, const uint16_t* src, uint16_t* dst, int sw, int sh, float sigma_depth_m);
int launch_pyr_maps(tfb_ctx* c, const uint16_t* src, uint16_t* dst, float4* pts, float4* nrm, int sw, int sh, float sigma_depth_m,
                    float fx, float fy, float cx, float cy, unsigned int* vmask = nullptr);
int launch_points_normals(tfb_ctx* c, const uint16_t* depth, float4* pts, float4* nrm, int w, int h, float fx, float fy,
                          float cx, float cy, bool with_valid_list = false);
int launch_resize_points_normals(tfb_ctx* c, const float4* v, const float4* n, float4* vo, float4* no, int sw, int sh);
// icp
int launch_icp_iteration(tfb_ctx* c, int level, const float4* vcurr, const float4* ncurr, const float4* vprev,
                         const float4* nprev, int w, int h, float fx, float fy, float cx, float cy, bool solve,
                         float* out27_dev, bool first_iter, bool last_iter);
int launch_icp_begin(tfb_ctx* c);
int launch_icp_all(tfb_ctx* c, bool update_pose);
int launch_icp_all_ext(tfb_ctx* c, int levels, const float* const* vcurr, const float* const* ncurr, const float* const* vprev,
                       const float* const* nprev, int cols, int rows, const int* iters, float dist_thres, float angle_thres);
int launch_pose_set(tfb_ctx* c, const float* pose_row_major_host, bool is_w2c);
// scene
int launch_reset_scene(tfb_ctx* c);
int launch_dir_rebuild(tfb_ctx* c);
int launch_stream_select(tfb_ctx* c, int mode, const float* pose_w2c_rowmajor, int* list_dev, int cap, int* counter_dev);
int launch_stream_evict(tfb_ctx* c, const int* list_dev, int n, void* xfer_dev);
int launch_stream_restore(tfb_ctx* c, const int* list_dev, int n, const void* xfer_dev, int* restored_dev);
void stream_store_clear(tfb_ctx* c);   // tfb_export.cu: the host side of the block streaming
void stream_store_free(tfb_ctx* c);
int launch_allocate(tfb_ctx* c, const float* dists);
int launch_integrate(tfb_ctx* c, const float* dists);
int launch_rebuild_visible(tfb_ctx* c);
// vis
int launch_expected_depths(tfb_ctx* c, bool reset_image = false);
int launch_icp_maps(tfb_ctx* c, float4* points, float4* normals, bool do_raycast = true);
int launch_render_grey(tfb_ctx* c, uchar4* out);
int launch_point_cloud(tfb_ctx* c, float4* out, int capacity, unsigned int* counter, bool skip_points);
int launch_raycast(tfb_ctx* c, bool update_visible);
// publish_epoch != 0: the last CTA to finish publishes that barrier epoch to every rank (this rank's rows and marks are out)
int launch_raycast_sharded(tfb_ctx* c, bool viewer, unsigned int publish_epoch = 0u);
int launch_shard_barrier(tfb_ctx* c);
// barrier_epoch != 0: the launch is also the cross-GPU barrier in front of it (publishes, then every CTA waits)
int launch_gather_foreign(tfb_ctx* c, unsigned int barrier_epoch = 0u);
int launch_wait_frame(tfb_ctx* c, unsigned int seq);
// Sharded scene: a new generation of the foreign-block cache (after anything that changes voxels, and for every gather).
// 0 is the generation of the zeroed tags and 1 the one the device-resident ShardView carries: neither is ever handed out.
inline void next_cache_epoch(tfb_ctx* c) {
    if (!c->cache_tag) return;
    if (++c->gather_epoch < 2u) {   // wrapped: no tag of the previous cycle may match again
        cudaMemsetAsync(c->cache_tag, 0, (size_t)c->total_entries * sizeof(unsigned long long), c->stream);
        c->gather_epoch = 2u;
    }
    c->shard.cache_epoch = c->gather_epoch;
}
// collective: into landing buffer (seq & 1) of every rank, and the last CTA publishes seq in every rank's frame flag
int launch_shard_push_frame(tfb_ctx* c, const uint16_t* depth_dev, bool collective = false, unsigned int seq = 0u);
// wait_epoch != 0: every CTA first waits until all ranks have published that epoch (their raycast rows have arrived)
int launch_model_maps(tfb_ctx* c, unsigned int wait_epoch = 0u);

}  // namespace tfb
