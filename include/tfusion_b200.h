/* tfusion_b200 — C ABI of the B200-native per-frame dense-reconstruction hot path.
 *
 * This is the boundary a maintainer of 3d-scan/topfusion binds instead of the reference's
 * internal host->device seam.  Every entry point names the reference interface it replaces
 * (paths relative to /root/reference/tfusion).  Plain pointers and sizes only; no torch,
 * OpenCV or C++ types.  All functions return TFB_OK (0) or a negative tfb_status and never
 * call exit() (the reference prints and exits: src/safe_call.hpp:13-27,
 * include/tfusion/cuda/CUDADefines.hpp:22-33).  A context is not thread-safe (neither is
 * the reference).  There is NO CPU fallback: without a CUDA device tfb_create fails.
 *
 * Conventions
 *   - images are dense row-major device buffers (no pitch)
 *   - points / normals are float4 per pixel (x,y,z,w), NaN x marks an invalid pixel
 *   - poses are row-major float[16] (the storage of cv::Affine3f::matrix.val), metres
 *   - "pose_w2c" = world->camera, "pose_c2w" = camera->world (TopFu::poses_ are c2w)
 */
#ifndef TFUSION_B200_H
#define TFUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TFB_API __attribute__((visibility("default")))
#else
#define TFB_API
#endif

typedef enum tfb_status {
    TFB_OK = 0,
    TFB_ERR_ARG = -1,    /* null pointer, bad size, unsupported parameter */
    TFB_ERR_CUDA = -2,   /* a CUDA call failed; see tfb_last_error */
    TFB_ERR_NOMEM = -3,
    TFB_ERR_STATE = -4   /* call not valid in the current state */
} tfb_status;

/* TopFuParams + SceneParams (src/topfu.cpp:12-53, include/tfusion/SceneParams.hpp:45-53) and
 * the hash geometry the reference fixes at compile time (include/tfusion/cuda/VoxelBlockHash.hpp:10-18),
 * here run-time so the 2M-block / 2 mm configurations fit. */
typedef struct tfb_params {
    int32_t cols, rows;             /* rows % 8 == 0 and cols % 8 == 0 */
    float fx, fy, cx, cy;
    float bilateral_sigma_depth;    /* metres */
    float bilateral_sigma_spatial;  /* pixels */
    int32_t bilateral_kernel_size;  /* odd, <= 15 */
    float icp_truncate_depth_dist;  /* metres; <= 0 disables (topfu.cpp:190-191) */
    float icp_dist_thres;           /* metres */
    float icp_angle_thres;          /* radians */
    int32_t icp_iters[4];           /* iterations for pyramid level 0..3 (topfu.cpp:14) */
    float mu;                       /* truncation band, metres */
    int32_t max_w;
    float voxel_size;               /* metres */
    float view_frustum_min, view_frustum_max;
    int32_t stop_integrating_at_max_w;
    int32_t num_blocks;             /* SDF_LOCAL_BLOCK_NUM  (0x10000) */
    int32_t num_buckets;            /* SDF_BUCKET_NUM, power of two (0x100000) */
    int32_t excess_size;            /* SDF_EXCESS_LIST_SIZE (0x20000) */
    int32_t depth_cutoff_mm;        /* src/cuda/imgproc.cu:277 hard-codes 2047 */
    int32_t corrected_mode;         /* 0: reference behaviour incl. SURVEY.md F1; 1: model maps in the camera frame */
    int32_t shard_rank, shard_count;/* voxel payload sharded by block coordinate; index replicated */
    int32_t defer_tail;             /* 1 (default): tfb_process_frame returns as soon as the pose is known; allocation,
                                     * integration, raycast and model maps of that frame are enqueued by the next call (beside
                                     * its preprocessing, on a second stream) or by any call that looks at the scene.  0: the
                                     * whole frame is finished before the call returns.  2: the call still returns as soon as
                                     * the pose is known, but those stages are already enqueued behind the frame's ICP — the GPU
                                     * runs them while the host turns around (for consumers that do not synchronise with the
                                     * device between frames: a decode-ahead ring, a script).  Results are identical. */
    int32_t ieee_arith;             /* TSDF integration arithmetic.  0 (default): the arithmetic of the reference's own GPU
                                     * build (tfusion/CMakeLists.txt:1: --ftz=true --prec-div=false, fused multiply-adds) —
                                     * every voxel bit-identical to what the reference's integrateIntoScene_device writes on
                                     * the same GPU.  1: IEEE evaluation without contraction — every voxel bit-identical to a
                                     * HOST compile of computeUpdatedVoxelDepthInfo (the CPU oracle).  DESIGN.md section 4. */
} tfb_params;

typedef struct tfb_ctx tfb_ctx;

/* ---- life cycle ------------------------------------------------------------------------ */
/* TopFuParams::default_params(), src/topfu.cpp:12-53 */
TFB_API int tfb_default_params(tfb_params* p);
/* TopFu::TopFu, src/topfu.cpp:55-84 (scene, engines, render state, ICP, buffers, reset).
 * stream: a cudaStream_t to run on, or NULL to let the context create its own non-blocking stream. */
TFB_API int tfb_create(const tfb_params* p, void* stream, tfb_ctx** out);
TFB_API int tfb_destroy(tfb_ctx* c);
/* TopFu::reset, src/topfu.cpp:141-152 -> ResetScene, src/cuda/SceneReconstructionEngine_host.cu:52-73 */
TFB_API int tfb_reset(tfb_ctx* c);
TFB_API const char* tfb_last_error(const tfb_ctx* c);
TFB_API const char* tfb_version(void);

/* ---- raw device memory helpers (cuda::DeviceMemory, src/device_memory.cpp:46-135) ------- */
TFB_API int tfb_dev_alloc(void** dptr, size_t bytes);
TFB_API int tfb_dev_free(void* dptr);
TFB_API int tfb_host_alloc_pinned(void** hptr, size_t bytes);
TFB_API int tfb_host_free_pinned(void* hptr);
TFB_API int tfb_h2d(tfb_ctx* c, void* dst_dev, const void* src_host, size_t bytes);  /* async on the context stream */
TFB_API int tfb_d2h(tfb_ctx* c, void* dst_host, const void* src_dev, size_t bytes);  /* waits for the stream */
TFB_API int tfb_sync(tfb_ctx* c);                                                    /* cuda::waitAllDefaultStream */

/* ---- image stages: tfusion::device::* declared in src/internal.hpp:122-132 ------------- */
/* compute_dists, src/cuda/imgproc.cu:263-290 */
TFB_API int tfb_compute_dists(tfb_ctx* c, const uint16_t* depth, float* dists, int cols, int rows);
/* bilateralFilter, imgproc.cu:10-61 */
TFB_API int tfb_bilateral_filter(tfb_ctx* c, const uint16_t* src, uint16_t* dst, int cols, int rows,
                                 int kernel_size, float sigma_spatial, float sigma_depth);
/* truncateDepth, imgproc.cu:70-89 */
TFB_API int tfb_truncate_depth(tfb_ctx* c, uint16_t* depth, int cols, int rows, float max_dist);
/* depthPyr, imgproc.cu:98-140; dst is (src_cols/2, src_rows/2) */
TFB_API int tfb_depth_pyr(tfb_ctx* c, const uint16_t* src, uint16_t* dst, int src_cols, int src_rows, float sigma_depth);
/* computePointNormals, imgproc.cu:214-254 */
TFB_API int tfb_compute_point_normals(tfb_ctx* c, const uint16_t* depth, float* points, float* normals,
                                      int cols, int rows, float fx, float fy, float cx, float cy);
/* resizePointsNormals, imgproc.cu:355-401 */
TFB_API int tfb_resize_points_normals(tfb_ctx* c, const float* points, const float* normals,
                                      float* points_out, float* normals_out, int src_cols, int src_rows);
/* the fused form the frame path uses: dists + bilateral + truncate + pyramid + maps for all used levels,
 * written into the context's current-frame buffers (src/topfu.cpp:166-197) */
TFB_API int tfb_preprocess(tfb_ctx* c, const uint16_t* depth_dev);

/* ---- ICP: ComputeIcpHelper::operator() src/cuda/proj_icp.cu:432-455, StreamHelper::get and
 *      ProjectiveICP::estimateTransform src/projective_icp.cpp:43-62,169-212 ---------------- */
/* one reduction for a fixed transform: 27 sums in the reference's packed order A00..A05,b0,A11..b5 */
TFB_API int tfb_icp_reduce(tfb_ctx* c, int cols, int rows, float fx, float fy, float cx, float cy,
                           const float aff[16], const float* vcurr, const float* ncurr,
                           const float* vprev, const float* nprev, float out27_host[27]);
/* the whole coarse-to-fine loop on the context's current / model pyramids; solve on the device */
TFB_API int tfb_icp_estimate(tfb_ctx* c, float affine_out[16], int* ok);

/* ---- scene engine: SceneReconstructionEngine_CUDA, include/tfusion/cuda/SceneReconstructionEngine_host.hpp:52-57 */
/* AllocateSceneFromDepth, src/cuda/SceneReconstructionEngine_host.cu:76-195 */
TFB_API int tfb_allocate_scene_from_depth(tfb_ctx* c, const float pose_w2c[16], const float* dists_dev);
/* IntegrateIntoScene, SceneReconstructionEngine_host.cu:198-251 */
TFB_API int tfb_integrate_into_scene(tfb_ctx* c, const float pose_w2c[16], const float* dists_dev);

/* ---- visualisation engine: VisualisationEngine_CUDA, include/tfusion/cuda/VisualisationEngine_CUDA.hpp:38-45 */
/* CreateExpectedDepths, src/cuda/VisualisationEngine_CUDA.cu:120-173 */
TFB_API int tfb_create_expected_depths(tfb_ctx* c, const float pose_w2c[16]);
/* CreateICPMaps, VisualisationEngine_CUDA.cu:324-360,474-493: raycast (updates the visible set) + model maps */
TFB_API int tfb_create_icp_maps(tfb_ctx* c, const float pose_c2w[16], float* points_dev, float* normals_dev);

/* TopFu::renderImage -> RenderImage(RENDER_SHADED_GREYSCALE, RENDER_FROM_NEW_RAYCAST), src/topfu.cpp:332-356,
 * VisualisationEngine_CUDA.cu:220-291: raycast (visibility untouched) + SDF-gradient shading.  rgba_dev: rows x cols x 4 bytes.
 * pose_c2w may be NULL to use the context's current camera pose (what the reference does). */
TFB_API int tfb_render_image(tfb_ctx* c, const float* pose_c2w_or_null, uint8_t* rgba_dev);
/* ProjectiveICP::estimateTransform(affine, intr, vcurr, ncurr, vprev, nprev) on caller-owned device pyramids
 * (src/projective_icp.cpp:169-212); level l is (cols >> l) x (rows >> l) float4 maps, intrinsics from the context */
TFB_API int tfb_icp_estimate_ext(tfb_ctx* c, int levels, const float* const* vcurr, const float* const* ncurr,
                                 const float* const* vprev, const float* const* nprev, int cols, int rows,
                                 const int* iters, float dist_thres, float angle_thres, const float intr_or_null[4],
                                 float affine_out[16], int* ok);
/* ProjectiveICP::setDistThreshold / setAngleThreshold / setIterationsNum (include/tfusion/cuda/projective_icp.hpp:21-28);
 * the number of used pyramid levels may not grow beyond what the context was created with */
TFB_API int tfb_set_icp_params(tfb_ctx* c, float dist_thres, float angle_thres, const int iters[4]);
/* cuda::getCudaEnabledDeviceCount / getDeviceName / printShortCudaDeviceInfo, src/core.cpp */
TFB_API int tfb_device_count(void);
TFB_API int tfb_device_info(int device, char* name, int name_len, int* cc_major, int* cc_minor, int* sm_count, size_t* total_mem);
/* DeviceMemory2D::upload / download / copyTo with row strides, src/device_memory.cpp:193-250.
 * kind: 0 host->device (async), 1 device->host (waits), 2 device->device (async) */
TFB_API int tfb_memcpy_2d(tfb_ctx* c, void* dst, size_t dst_step, const void* src, size_t src_step, size_t width_bytes,
                          int rows, int kind);
/* DeviceMemory::copyTo, src/device_memory.cpp */
TFB_API int tfb_memcpy_d2d(tfb_ctx* c, void* dst_dev, const void* src_dev, size_t bytes);

/* ---- the frame: TopFu::operator(), src/topfu.cpp:161-330 ------------------------------- */
/* depth_host: rows x cols u16 millimetres with row stride step_bytes (cv::Mat data/step; pinned memory
 * makes the upload asynchronous).  *ok receives operator()'s return value. */
TFB_API int tfb_process_frame(tfb_ctx* c, const uint16_t* depth_host, size_t step_bytes, int* ok);
/* same with the frame already on the device, dense (cuda::Depth after upload, apps/demo.cpp:100-104) */
TFB_API int tfb_process_frame_device(tfb_ctx* c, const uint16_t* depth_dev, int* ok);
/* TopFu::getCameraPose, src/topfu.cpp:154-159 */
TFB_API int tfb_get_pose(const tfb_ctx* c, int time, float out16[16]);
TFB_API int tfb_num_poses(const tfb_ctx* c);

/* ---- scene sharded across GPUs (new: the reference is single-GPU) ---------------
 * One context per GPU (one process per GPU), created with shard_rank / shard_count.  The hash INDEX is replicated
 * (every rank sees every frame and takes the same allocation decisions); the 2 KB voxel payload of a block lives
 * only in the pool of its owner (a mix of the block coordinate).  A frame runs in three stages with a cross-GPU
 * barrier on the context stream between them (the caller supplies it: an NCCL all-reduce on the same stream):
 *
 *   tfb_frame_begin    preprocess + ICP + allocation (replicated), integration of this rank's blocks (no exchange)
 *   -- barrier --
 *   tfb_frame_raycast  expected depths; this rank casts its 8-row strips (strip % shard_count == shard_rank) and reads
 *                      the voxels of foreign blocks straight out of the owner's table + pool over NVLink peer memory;
 *                      finished rows and visibility marks are stored into EVERY rank's buffers by the same kernel
 *   -- barrier --
 *   tfb_frame_end      incoming visibility marks, model maps, map pyramid, state read-back (operator()'s verdict)
 *
 * The march of every ray is the single-GPU march, so the result is identical to one context holding the whole scene.
 * Peer buffers are attached as plain device pointers valid in the calling process: another context's pointers on the
 * same device (single-process emulation, tests) or pointers opened from CUDA IPC handles (one process per GPU). */
#define TFB_MAX_SHARDS 16
typedef struct tfb_shard_ptrs {
    void* table;     /* total_entries x 16 B hash entries */
    void* vba;       /* num_blocks x 2 KB voxel pool */
    void* raycast;   /* rows x cols float4 raycast result */
    void* marks;     /* incoming visibility marks: u32 count, pad, then 2 x u32 per mark */
    void* frame;     /* 2 x rows x cols u16: landing buffers for the depth frame the sensor's rank pushes to every rank
                      * (tfb_shard_push_frame uses the first; tfb_process_frame_sharded alternates) */
    void* flags;     /* 4 x TFB_MAX_SHARDS x u32, written by the peers: [r] = last barrier epoch rank r has reached,
                      * [TFB_MAX_SHARDS] = sequence number of the last frame pushed here, [2*TFB_MAX_SHARDS + r] = last pushed
                      * frame rank r has finished reading; the words in between are this rank's own tickets */
} tfb_shard_ptrs;
TFB_API int tfb_shard_local_ptrs(tfb_ctx* c, tfb_shard_ptrs* out);
TFB_API int tfb_shard_attach(tfb_ctx* c, int rank, const tfb_shard_ptrs* peer);
TFB_API int tfb_ipc_export(const void* dev_ptr, unsigned char handle64[64]);   /* cudaIpcGetMemHandle */
TFB_API int tfb_ipc_open(const unsigned char handle64[64], void** dev_ptr);    /* cudaIpcOpenMemHandle, peer access enabled */
TFB_API int tfb_ipc_close(void* dev_ptr);
/* Cross-GPU plumbing over the same peer pointers, ON THE CONTEXT STREAM (no host synchronisation, no NCCL call):
 * tfb_shard_push_frame  rank 0 only: one kernel stores its device frame into every rank's frame buffer (NVLink stores);
 * tfb_shard_barrier     every rank: one warp publishes this rank's epoch into every rank's flag array with system-scope
 *                       release stores and waits until every rank has published it.  Real multi-GPU only: the ranks'
 *                       kernels wait on one another, so every rank must own a GPU (never two ranks on one device — the
 *                       single-process emulation orders the stages by stream order instead).  A rank that waits longer
 *                       than ~10 s gives up; the frame then reports TFB_ERR_STATE.
 * tfb_frame_begin(c, NULL) then tracks the frame in the context's frame buffer. */
/* tfb_process_frame_sharded: the whole sharded frame in ONE call per rank, software-pipelined like tfb_process_frame.  The
 * rank that passes a frame (the sensor's; always the same one, normally rank 0 — the others pass NULL) pushes it into a
 * landing buffer of every rank on the second stream, beside the previous frame's tail; the receivers wait for its sequence
 * number there and acknowledge when they have read it.  The tail's two cross-GPU barriers are folded into its kernels
 * (k_gather_foreign publishes "integrated" and waits; k_raycast_sharded's last CTA publishes "rows out", k_model_maps
 * waits), so the sharded frame has no barrier launch.  *ok is operator()'s verdict, identical on every rank.  Real
 * multi-GPU only (kernels of different ranks wait on one another).  Collective: every rank calls it
 * for every frame, in the same order; with defer_tail, so is every call that finishes a pending tail on a sharded context
 * (tfb_sync, tfb_get_counters, the tfb_export_* family, tfb_render_image ...). */
TFB_API int tfb_process_frame_sharded(tfb_ctx* c, const uint16_t* depth_dev_or_null, int* ok);
TFB_API int tfb_shard_push_frame(tfb_ctx* c, const uint16_t* depth_dev);
TFB_API int tfb_shard_barrier(tfb_ctx* c);
TFB_API int tfb_frame_begin(tfb_ctx* c, const uint16_t* depth_dev);
TFB_API int tfb_frame_raycast(tfb_ctx* c);
TFB_API int tfb_frame_end(tfb_ctx* c, int* ok);
TFB_API void* tfb_stream(tfb_ctx* c);   /* the cudaStream_t the context runs on */

/* ---- getting the reconstruction out and back in (the step after the path) -------------------
 * tfb_extract_points: surface points of the whole scene — zero crossings of the TSDF along the three voxel edges
 * (both voxels observed, w > 0), linearly interpolated, world metres, float4 (x, y, z, 1).  The reference has only a stub
 * (take_cloud, apps/demo.cpp:70-77) and a dormant per-pixel renderPointCloud_device (VisualisationHelper.hpp:150-198).
 * *n_out = points found; min(*n_out, capacity) were written (call with capacity 0 to size the buffer).  Order is arbitrary. */
TFB_API int tfb_extract_points(tfb_ctx* c, float* points_dev, int capacity, int* n_out);
/* renderPointCloud_device, include/tfusion/cuda/VisualisationHelper.hpp:150-198 (dormant in the reference: its caller
 * take_cloud is a stub, apps/demo.cpp:70-77): the surface points one view sees — raycast from pose_c2w (NULL: the current
 * pose) without touching visibility, every hit with a light-facing SDF-gradient normal, x y z 1 in world metres.
 * skip_points != 0 keeps pixels with odd x and odd y only.  n_out = points found; min(n_out, capacity) are written. */
TFB_API int tfb_render_point_cloud(tfb_ctx* c, const float* pose_c2w_or_null, int skip_points, float* points_dev, int capacity, int* n_out);
/* allocated blocks + pose history in one file; tfb_scene_load restores them into a context created with the same voxel
 * size, band and hash geometry, rebuilds the visible list and the model maps for the last pose, and tracking goes on. */
TFB_API int tfb_scene_save(tfb_ctx* c, const char* path);
TFB_API int tfb_scene_load(tfb_ctx* c, const char* path);

/* Block streaming between the voxel pool and host memory, for scenes larger than the pool.  Replaces the reference's host swap
 * cache and its kernels, which the reference never switches on: GlobalCache (tfusion/include/tfusion/GlobalCache.hpp:14-135),
 * reAllocateSwappedOutVoxelBlocks_device (tfusion/src/cuda/SceneReconstructionEngine_host.cu:417-432), the enlarged frustum of
 * checkPointVisibility<true> (tfusion/include/tfusion/cuda/SceneReconstructionEngine.hpp:315-322), Scene(..., useSwapping)
 * (tfusion/src/topfu.cpp:67).  Synchronous calls between frames; not for a sharded context.
 * tfb_stream_out: moves up to max_blocks (<= 0: all) resident blocks that are not visible in the current frame and lie outside the
 *   enlarged frustum (image widened by 1/8 per side) of the current pose into the host store; their pool slots are free again and
 *   their hash entries carry ptr = -1, which the integration skips and the raycast reads as empty space, as in the reference.
 * tfb_stream_in: brings back the stored blocks inside the enlarged frustum of pose_w2c (row-major world->camera; NULL: the current
 *   pose), or every stored block when all != 0.  A block that finds no free pool slot stays in the store (*n_left_in_store).
 *   Restoring before the frame that looks at a block makes the scene identical to one that was never streamed.
 * tfb_reset drops the store; tfb_scene_save fails with TFB_ERR_STATE while blocks are out. */
TFB_API int tfb_stream_out(tfb_ctx* c, int max_blocks, int* n_out);
TFB_API int tfb_stream_in(tfb_ctx* c, const float* pose_w2c_or_null, int all, int* n_in, int* n_left_in_store);
TFB_API int tfb_stream_stats(tfb_ctx* c, long long* blocks_in_pool, long long* blocks_in_store);

/* ---- inspection (tests, bench, debug dumps) ---------------------------------------------- */
/* out[8] = n_visible, last_free_block, last_free_excess, n_new_this_frame, frame_counter, resets,
 *          n_raycast_extras, n_allocated */
TFB_API int tfb_get_counters(tfb_ctx* c, long long out[8]);
TFB_API long long tfb_voxel_updates_last(tfb_ctx* c);   /* 512 x visible entries with ptr >= 0; finishes a deferred tail */
TFB_API long long tfb_voxel_updates_total(const tfb_ctx* c);   /* summed over every integration finished so far; does not wait */
TFB_API int tfb_total_entries(const tfb_ctx* c);
TFB_API int tfb_export_table(tfb_ctx* c, void* host_entries);       /* total_entries x 16 B HashEntry */
TFB_API int tfb_export_vis_type(tfb_ctx* c, uint8_t* host);         /* total_entries */
TFB_API int tfb_export_visible_ids(tfb_ctx* c, int32_t* host, int capacity, int* n);
TFB_API int tfb_export_block(tfb_ctx* c, int ptr, void* host512x4); /* one block: 512 x {i16 sdf, u8 w, u8 pad} */
TFB_API int tfb_export_minmax(tfb_ctx* c, float* host);             /* (rows/8) x (cols/8) x 2 */
TFB_API int tfb_export_raycast(tfb_ctx* c, float* host);            /* rows x cols x 4, voxel units */
TFB_API int tfb_export_dists(tfb_ctx* c, float* host);
/* which: 0 current depth (u16), 1 current points, 2 current normals, 3 model points, 4 model normals */
TFB_API int tfb_export_level(tfb_ctx* c, int which, int level, void* host);
TFB_API int tfb_import_level(tfb_ctx* c, int which, int level, const void* host);
/* the level-0 pixels that hold a vertex, ascending (row-major index), as the last tfb_process_frame left them for its ICP:
 * proj_icp.cu:86-88 skips a pixel whose vertex is NaN; here every CTA of the ICP kernel takes an equal share of this list.
 * *n = 0 before the first tracked frame, or when the image width is not a multiple of 32 (the list is then not built). */
TFB_API int tfb_export_icp_valid_list(tfb_ctx* c, int32_t* host, int capacity, int* n);
/* device pointers of the context's own buffers (same `which` codes; 5 = dists) */
TFB_API void* tfb_level_ptr(tfb_ctx* c, int which, int level);

/* ---- timing on the context stream (CUDA events) -------------------------------------------
 * stage ids: 0 upload, 1 preprocess (upload + preprocessing, on the second stream), 2 icp, 3 allocate, 4 integrate,
 *            5 expected depth, 6 raycast + model maps, 7 wait for the preprocessing stream, 8 whole call.
 *            With defer_tail, stages 3..6 of a call are those of the PREVIOUS frame (they run first, beside stage 1). */
TFB_API int tfb_timing_enable(tfb_ctx* c, int on);
TFB_API int tfb_timing_last_ms(tfb_ctx* c, float out9[9]);
/* number of kernels this library launched since the context was created */
TFB_API long long tfb_kernel_launches(const tfb_ctx* c);
/* overwrite a 256 MB scratch buffer on the context stream: evicts the 126 MB L2 between timed steps */
TFB_API int tfb_flush_l2(tfb_ctx* c);
/* cuda::setDevice, src/core.cpp — must precede tfb_create in a multi-GPU process (one rank per GPU) */
TFB_API int tfb_set_device(int device);
/* event markers on the context stream: tfb_mark(c, a) ... tfb_mark(c, b); tfb_elapsed_ms(c, a, b, &ms) */
TFB_API int tfb_mark(tfb_ctx* c, int slot);
TFB_API int tfb_elapsed_ms(tfb_ctx* c, int slot_a, int slot_b, float* ms);
/* per-launch timing: when enabled every kernel launch is bracketed by two events on the context stream and the
 * durations are accumulated per kernel (adds ~2 us per launch; keep it off in throughput runs) */
TFB_API int tfb_ktiming_enable(tfb_ctx* c, int on);
TFB_API int tfb_ktiming_reset(tfb_ctx* c);
TFB_API int tfb_ktiming_count(void);
TFB_API const char* tfb_ktiming_name(int id);
TFB_API int tfb_ktiming_get(tfb_ctx* c, int id, double* total_ms, long long* launches);

#ifdef __cplusplus
}
#endif
#endif /* TFUSION_B200_H */
