// The step after the hot path: getting the reconstruction out of the voxel-block hash and back in.
//
//   tfb_extract_points   surface points of the whole scene: zero crossings of the TSDF along the three voxel edges.  The
//                        reference only has a stub for this (take_cloud, /root/reference/apps/demo.cpp:70-77) and a dormant
//                        per-pixel renderPointCloud_device (include/tfusion/cuda/VisualisationHelper.hpp:150-198); this is
//                        the volumetric extraction of the kinfu lineage, on the hash.
//   tfb_scene_save/load  the allocated blocks + pose history in one file; a loaded context can go on tracking.
//
// Compiled with --fmad=false like the rest: the interpolation is checked bit for bit against a numpy restatement.
#include <stdio.h>
#include <stdlib.h>

#include <new>
#include <vector>

#include "tfb_common.cuh"
#include <unordered_map>

// Host side of the block streaming (tfb_stream_out / tfb_stream_in): GlobalCache restated (GlobalCache.hpp:14-135 keeps one 2 KB
// slot per hash entry, 2.4 GB for the default table; here blocks are stored where they arrive and found through a map).
struct HostBlockStore {
    std::unordered_map<int, size_t> where;      // hash entry -> index of its 2 KB in `data`
    std::vector<unsigned int> data;             // BLOCK3 words per stored block
    std::vector<size_t> free_slots;
    int* list_dev = nullptr;                    // transfer: entry ids, success flags, counter, blocks
    int* flag_dev = nullptr;
    int* counter_dev = nullptr;
    unsigned int* xfer_dev = nullptr;
    unsigned int* xfer_host = nullptr;          // pinned
    int* list_host = nullptr;                   // pinned: ids + flags + counter
    static constexpr int CHUNK = 8192;          // blocks per transfer (16 MB; the reference's SDF_TRANSFER_BLOCK_NUM is 0x1000)
};

namespace tfb {

__device__ __forceinline__ int ex_hash(int bx, int by, int bz, int mask) {
    return (int)((((unsigned)bx * 73856093u) ^ ((unsigned)by * 19349669u) ^ ((unsigned)bz * 83492791u)) & (unsigned)mask);
}

// pool index of block (bx,by,bz), or -1
__device__ int ex_find(const int4* __restrict__ table, int bx, int by, int bz, int num_buckets, int mask) {
    int slot = ex_hash(bx, by, bz, mask);
    const int k0 = (bx & 0xffff) | (by << 16);
    for (;;) {
        const int4 e = __ldg(table + slot);
        if (e.x == k0 && (short)(e.y & 0xffff) == bz && e.w >= 0) return e.w;
        if (e.z < 1) return -1;
        slot = num_buckets + e.z - 1;
    }
}

constexpr int EX_WARPS = 8;

// One warp per allocated block: the block is staged in shared memory, the three +1 neighbour blocks are looked up once,
// every lane examines 16 voxels x 3 edges.  Two passes (count, then write) so each block needs one atomicAdd.
__global__ void __launch_bounds__(EX_WARPS * 32)
    k_extract_points(const int4* __restrict__ table, const unsigned int* __restrict__ vox, int total_entries, int num_buckets, int mask,
                     float voxel_size, float4* __restrict__ out, int capacity, unsigned int* __restrict__ counter) {
    __shared__ unsigned int s_blk[EX_WARPS][BLOCK3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps_total = gridDim.x * EX_WARPS;
    for (int base = (blockIdx.x * EX_WARPS + warp) * 32; base < total_entries; base += warps_total * 32) {
        int4 ev = make_int4(0, 0, 0, -2);
        if (base + lane < total_entries) ev = __ldg(table + base + lane);
        unsigned int todo = __ballot_sync(0xffffffffu, ev.w >= 0);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int ex = __shfl_sync(0xffffffffu, ev.x, src), ey = __shfl_sync(0xffffffffu, ev.y, src);
            const int ptr = __shfl_sync(0xffffffffu, ev.w, src);
            const int bx = (short)(ex & 0xffff), by = ex >> 16, bz = (short)(ey & 0xffff);
            const unsigned int* blk = vox + (size_t)ptr * BLOCK3;
            __syncwarp();
            for (int i = lane; i < BLOCK3; i += 32) s_blk[warp][i] = __ldg(blk + i);
            // lanes 0..2 find the +x, +y, +z neighbour blocks
            int nptr = -1;
            if (lane < 3) nptr = ex_find(table, bx + (lane == 0), by + (lane == 1), bz + (lane == 2), num_buckets, mask);
            const int nb[3] = {__shfl_sync(0xffffffffu, nptr, 0), __shfl_sync(0xffffffffu, nptr, 1), __shfl_sync(0xffffffffu, nptr, 2)};
            __syncwarp();
            unsigned int my_off = 0;
            for (int pass = 0; pass < 2; ++pass) {
                unsigned int cnt = 0;
                for (int k = 0; k < BLOCK3 / 32; ++k) {
                    const int lin = lane + 32 * k;
                    const unsigned int v0 = s_blk[warp][lin];
                    if (((v0 >> 16) & 0xffu) == 0) continue;
                    const float f0 = (float)(short)(v0 & 0xffffu) / 32767.0f;
                    const int lx = lin & 7, ly = (lin >> 3) & 7, lz = lin >> 6;
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
                        const int l = ax == 0 ? lx : (ax == 1 ? ly : lz);
                        const int stride = ax == 0 ? 1 : (ax == 1 ? BLOCK : BLOCK * BLOCK);
                        unsigned int v1;
                        if (l < BLOCK - 1) v1 = s_blk[warp][lin + stride];
                        else if (nb[ax] >= 0) v1 = __ldg(vox + (size_t)nb[ax] * BLOCK3 + lin - (BLOCK - 1) * stride);
                        else continue;
                        if (((v1 >> 16) & 0xffu) == 0) continue;
                        const float f1 = (float)(short)(v1 & 0xffffu) / 32767.0f;
                        if (!((f0 > 0.f && f1 < 0.f) || (f0 < 0.f && f1 > 0.f))) continue;
                        if (pass == 1) {
                            const unsigned int at = my_off + cnt;
                            if (at < (unsigned)capacity) {
                                const float t = f0 / (f0 - f1);
                                float g[3] = {(float)(bx * BLOCK + lx), (float)(by * BLOCK + ly), (float)(bz * BLOCK + lz)};
                                g[ax] = g[ax] + t;
                                out[at] = make_float4(g[0] * voxel_size, g[1] * voxel_size, g[2] * voxel_size, 1.0f);
                            }
                        }
                        ++cnt;
                    }
                }
                if (pass == 0) {
                    // exclusive prefix over the lanes, one reservation per block
                    unsigned int incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += up;
                    }
                    const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
                    unsigned int start = 0;
                    if (lane == 0 && total) start = atomicAdd(counter, total);
                    start = __shfl_sync(0xffffffffu, start, 0);
                    my_off = start + incl - cnt;
                    if (total == 0) break;
                }
            }
        }
    }
}

}  // namespace tfb

using namespace tfb;

extern "C" {

int tfb_extract_points(tfb_ctx* c, float* points_dev, int capacity, int* n_out) {
    if (!c || !n_out || capacity < 0 || (capacity > 0 && !points_dev)) return TFB_ERR_ARG;
    if (c->p.shard_count > 1) return set_err(c, TFB_ERR_STATE, "tfb_extract_points: every rank extracts its own blocks; face neighbours on other ranks are not read yet");
    int r = tfb_sync(c);   // finishes a deferred tail
    if (r) return r;
    unsigned int* counter = c->marks;   // idle outside a sharded frame
    TFB_CUDA(c, cudaMemsetAsync(counter, 0, sizeof(unsigned int), c->stream));
    k_extract_points<<<NUM_SMS * 4, EX_WARPS * 32, 0, c->stream>>>(reinterpret_cast<const int4*>(c->table),
                                                                   reinterpret_cast<const unsigned int*>(c->vba), c->total_entries,
                                                                   c->p.num_buckets, c->hash_mask, c->p.voxel_size,
                                                                   reinterpret_cast<float4*>(points_dev), capacity, counter);
    c->launches++;
    TFB_CUDA(c, cudaGetLastError());
    unsigned int n = 0;
    TFB_CUDA(c, cudaMemcpyAsync(&n, counter, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(c, cudaStreamSynchronize(c->stream));
    TFB_CUDA(c, cudaMemsetAsync(counter, 0, sizeof(unsigned int), c->stream));
    *n_out = (int)n;   // the number of surface points found; only min(n, capacity) were written
    return TFB_OK;
}

// renderPointCloud_device (include/tfusion/cuda/VisualisationHelper.hpp:150-198): the cloud of ONE view — a raycast from pose_c2w
// (poses_.back() when null) that does not touch visibility, then every hit pixel with a light-facing SDF-gradient normal, in
// world metres.  skip_points keeps only pixels with odd x and odd y, as the reference's flag does.  Order is unspecified.
int tfb_render_point_cloud(tfb_ctx* c, const float* pose_c2w_or_null, int skip_points, float* points_dev, int capacity, int* n_out) {
    if (!c || !n_out || capacity < 0 || (capacity > 0 && !points_dev)) return TFB_ERR_ARG;
    if (c->p.shard_count > 1) return set_err(c, TFB_ERR_STATE, "tfb_render_point_cloud: the mark queue doubles as the counter; use an unsharded context");
    int r = tfb_sync(c);   // finishes a deferred tail
    if (r) return r;
    if (pose_c2w_or_null && (r = launch_pose_set(c, pose_c2w_or_null, false))) return r;
    unsigned int* counter = c->marks;   // idle outside a sharded frame
    TFB_CUDA(c, cudaMemsetAsync(counter, 0, sizeof(unsigned int), c->stream));
    if ((r = launch_point_cloud(c, reinterpret_cast<float4*>(points_dev), capacity, counter, skip_points != 0))) return r;
    unsigned int n = 0;
    TFB_CUDA(c, cudaMemcpyAsync(&n, counter, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
    TFB_CUDA(c, cudaStreamSynchronize(c->stream));
    TFB_CUDA(c, cudaMemsetAsync(counter, 0, sizeof(unsigned int), c->stream));
    *n_out = (int)n;   // points found; only min(n, capacity) were written
    return TFB_OK;
}

// ---- scene file ------------------------------------------------------------------------------------------
// header: magic, version, voxel_size, mu, num_buckets, excess_size, n_blocks, n_poses; then n_poses x 16 floats;
// then n_blocks x { short pos[3], pad; 512 x u32 voxels }.  The hash geometry of the loading context must match.
constexpr int MAX_FILE_POSES = 1 << 24;   // 1 GB of poses: anything above is a corrupt header, not a trajectory
struct SceneFileHeader {
    char magic[8];
    int32_t version;
    float voxel_size, mu;
    int32_t num_buckets, excess_size, n_blocks, n_poses;
};

static int scene_save_impl(tfb_ctx* c, const char* path) {
    int r = tfb_sync(c);
    if (r) return r;
    if (c->store && !c->store->where.empty())
        return set_err(c, TFB_ERR_STATE, "tfb_scene_save: blocks are streamed out to the host; call tfb_stream_in(c, NULL, 1, ...) first");
    std::vector<HashEntry> table((size_t)c->total_entries);
    TFB_CUDA(c, cudaMemcpy(table.data(), c->table, table.size() * sizeof(HashEntry), cudaMemcpyDeviceToHost));
    std::vector<int> used;
    for (int i = 0; i < c->total_entries; ++i)
        if (table[i].ptr >= 0) used.push_back(i);
    FILE* f = fopen(path, "wb");
    if (!f) return set_err(c, TFB_ERR_ARG, "tfb_scene_save: cannot open the file");
    SceneFileHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "TFBSCENE", 8);
    h.version = 1; h.voxel_size = c->p.voxel_size; h.mu = c->p.mu;
    h.num_buckets = c->p.num_buckets; h.excess_size = c->p.excess_size;
    h.n_blocks = (int)used.size(); h.n_poses = c->n_poses;
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    ok = ok && fwrite(c->poses, 16 * sizeof(float), (size_t)c->n_poses, f) == (size_t)c->n_poses;
    std::vector<unsigned int> blk(BLOCK3);
    for (size_t k = 0; ok && k < used.size(); ++k) {
        const HashEntry& e = table[used[k]];
        short pos[4] = {e.pos[0], e.pos[1], e.pos[2], 0};
        if (cudaMemcpy(blk.data(), c->vba + (size_t)e.ptr * BLOCK3, BLOCK3 * sizeof(Voxel), cudaMemcpyDeviceToHost) != cudaSuccess) ok = false;
        ok = ok && fwrite(pos, sizeof(pos), 1, f) == 1 && fwrite(blk.data(), sizeof(unsigned int), BLOCK3, f) == (size_t)BLOCK3;
    }
    fclose(f);
    return ok ? TFB_OK : set_err(c, TFB_ERR_CUDA, "tfb_scene_save: write failed");
}

// std::vector may throw; nothing may leave an extern "C" function
int tfb_scene_save(tfb_ctx* c, const char* path) {
    if (!c || !path) return TFB_ERR_ARG;
    try {
        return scene_save_impl(c, path);
    } catch (const std::bad_alloc&) {
        return set_err(c, TFB_ERR_NOMEM, "tfb_scene_save: out of host memory");
    } catch (...) {
        return set_err(c, TFB_ERR_STATE, "tfb_scene_save: unexpected failure");
    }
}

static int scene_load_impl(tfb_ctx* c, const char* path) {
    if (c->p.shard_count > 1) return set_err(c, TFB_ERR_STATE, "tfb_scene_load: not for a sharded context");
    FILE* f = fopen(path, "rb");
    if (!f) return set_err(c, TFB_ERR_ARG, "tfb_scene_load: cannot open the file");
    SceneFileHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "TFBSCENE", 8) != 0 || h.version != 1) {
        fclose(f);
        return set_err(c, TFB_ERR_ARG, "tfb_scene_load: not a scene file");
    }
    if (h.num_buckets != c->p.num_buckets || h.excess_size != c->p.excess_size || h.voxel_size != c->p.voxel_size || h.mu != c->p.mu ||
        h.n_blocks < 0 || h.n_blocks > c->p.num_blocks || h.n_poses < 1 || h.n_poses > MAX_FILE_POSES) {
        fclose(f);
        return set_err(c, TFB_ERR_ARG, "tfb_scene_load: the file was written with a different voxel size / band / hash geometry, or its header is corrupt");
    }
    {   // the header must describe exactly the bytes that follow: sizes below are taken from it
        const long long want = (long long)sizeof(h) + (long long)h.n_poses * 64 + (long long)h.n_blocks * (8 + (long long)BLOCK3 * 4);
        long long have = -1;
        const long here = ftell(f);
        if (here >= 0 && fseek(f, 0, SEEK_END) == 0) { have = ftell(f); fseek(f, here, SEEK_SET); }
        if (have != want) {
            fclose(f);
            return set_err(c, TFB_ERR_ARG, "tfb_scene_load: file size does not match its header");
        }
    }
    std::vector<float> poses((size_t)h.n_poses * 16);
    bool ok = fread(poses.data(), 16 * sizeof(float), (size_t)h.n_poses, f) == (size_t)h.n_poses;
    // rebuild the table on the host with the allocator's own rules: bucket head if free, else a child taken from the top
    // of the excess free list and linked from the chain tail (allocateVoxelBlocksList_device); pool slots from the top too
    const int nbk = c->p.num_buckets, nex = c->p.excess_size, nbl = c->p.num_blocks, mask = c->hash_mask;
    std::vector<HashEntry> table((size_t)c->total_entries);
    for (auto& e : table) { e.pos[0] = e.pos[1] = e.pos[2] = 0; e.pad_ = 0; e.offset = 0; e.ptr = -2; }
    std::vector<unsigned int> bits((size_t)(nbk + 31) / 32, 0u);
    std::vector<unsigned int> pool((size_t)h.n_blocks * BLOCK3);
    std::vector<int> list;
    list.reserve((size_t)h.n_blocks);
    int last_free_block = nbl - 1, last_free_excess = nex - 1;
    for (int k = 0; ok && k < h.n_blocks; ++k) {
        short pos[4];
        ok = fread(pos, sizeof(pos), 1, f) == 1 && fread(&pool[(size_t)k * BLOCK3], sizeof(unsigned int), BLOCK3, f) == (size_t)BLOCK3;
        if (!ok) break;
        const int ptr = last_free_block--;   // vba_free is the identity permutation after a reset
        int slot = (int)((((unsigned)(int)pos[0] * 73856093u) ^ ((unsigned)(int)pos[1] * 19349669u) ^ ((unsigned)(int)pos[2] * 83492791u)) & (unsigned)mask);
        auto same = [&](int s_) { return table[s_].pos[0] == pos[0] && table[s_].pos[1] == pos[1] && table[s_].pos[2] == pos[2]; };
        if (table[slot].ptr < -1) {
            bits[slot >> 5] |= 1u << (slot & 31);
        } else {
            bool dup = same(slot);
            while (!dup && table[slot].offset >= 1) { slot = nbk + table[slot].offset - 1; dup = same(slot); }
            if (dup) { fclose(f); return set_err(c, TFB_ERR_ARG, "tfb_scene_load: a block position occurs twice in the file"); }
            if (last_free_excess < 0) { ok = false; break; }
            const int off = last_free_excess--;   // excess_free is the identity permutation after a reset
            table[slot].offset = off + 1;
            slot = nbk + off;
        }
        table[slot].pos[0] = pos[0]; table[slot].pos[1] = pos[1]; table[slot].pos[2] = pos[2];
        table[slot].offset = 0; table[slot].ptr = ptr;
        list.push_back(slot);
    }
    fclose(f);
    if (!ok) return set_err(c, TFB_ERR_ARG, "tfb_scene_load: truncated or inconsistent file");

    int r = tfb_reset(c);   // voxels, free lists, claim keys, bitmap; pose history back to identity
    if (r) return r;
    TFB_CUDA(c, cudaMemcpy(c->table, table.data(), table.size() * sizeof(HashEntry), cudaMemcpyHostToDevice));
    TFB_CUDA(c, cudaMemcpy(c->bucket_bits, bits.data(), bits.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
    if ((r = launch_dir_rebuild(c))) return r;
    // block k went to pool slot nbl-1-k: upload the pool tail in one piece, reversed block order
    {
        std::vector<unsigned int> rev(pool.size());
        for (int k = 0; k < h.n_blocks; ++k)
            memcpy(&rev[(size_t)(h.n_blocks - 1 - k) * BLOCK3], &pool[(size_t)k * BLOCK3], BLOCK3 * sizeof(unsigned int));
        if (h.n_blocks)
            TFB_CUDA(c, cudaMemcpy(c->vba + (size_t)(nbl - h.n_blocks) * BLOCK3, rev.data(), rev.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
    }
    // every block is a candidate for the visible list; the frustum test of the restored pose decides (type 3)
    TFB_CUDA(c, cudaMemset(c->vis_type, 0, (size_t)c->total_entries * sizeof(int)));
    {
        std::vector<int> three(list.size(), 3);
        (void)three;
        std::vector<int> vt((size_t)c->total_entries, 0);
        for (int s2 : list) vt[s2] = 3;
        TFB_CUDA(c, cudaMemcpy(c->vis_type, vt.data(), vt.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (!list.empty()) TFB_CUDA(c, cudaMemcpy(c->vis_list[0], list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    DevState st;
    TFB_CUDA(c, cudaMemcpy(&st, c->ds, sizeof(st), cudaMemcpyDeviceToHost));
    st.last_free_block = last_free_block; st.last_free_excess = last_free_excess;
    st.n_visible = (int)list.size(); st.n_next = 0; st.cur_list = 0; st.n_claimed = 0; st.n_new_frame = 0; st.n_extras = 0;
    st.icp_failed = 0; st.list_ticket = 0; st.int_cursor = 0;
    TFB_CUDA(c, cudaMemcpy(c->ds, &st, sizeof(st), cudaMemcpyHostToDevice));
    // pose history and the model maps the next frame's ICP needs
    if (h.n_poses > c->cap_poses) {
        float* np = (float*)realloc(c->poses, (size_t)h.n_poses * 2 * 16 * sizeof(float));
        if (!np) return set_err(c, TFB_ERR_NOMEM, "pose history");
        c->poses = np; c->cap_poses = h.n_poses * 2;
    }
    memcpy(c->poses, poses.data(), poses.size() * sizeof(float));
    c->n_poses = h.n_poses;
    c->frame_counter = h.n_poses;   // > 0: the next frame is tracked against the restored model
    if ((r = launch_pose_set(c, &poses[(size_t)(h.n_poses - 1) * 16], false))) return r;
    if ((r = launch_rebuild_visible(c))) return r;
    if ((r = launch_expected_depths(c))) return r;
    if ((r = launch_raycast(c, true))) return r;
    if ((r = launch_model_maps(c))) return r;
    TFB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TFB_OK;
}

int tfb_scene_load(tfb_ctx* c, const char* path) {
    if (!c || !path) return TFB_ERR_ARG;
    try {
        return scene_load_impl(c, path);
    } catch (const std::bad_alloc&) {
        return set_err(c, TFB_ERR_NOMEM, "tfb_scene_load: out of host memory");
    } catch (...) {
        return set_err(c, TFB_ERR_STATE, "tfb_scene_load: unexpected failure");
    }
}

// ---- block streaming (device side and reference citations in tfb_scene.cu) ----
static int stream_prepare(tfb_ctx* c) {
    if (c->p.shard_count > 1) return set_err(c, TFB_ERR_STATE, "block streaming: not for a sharded context (ptr = -1 means \"held by another rank\" there)");
    if (c->store) return TFB_OK;
    HostBlockStore* st = new (std::nothrow) HostBlockStore();
    if (!st) return set_err(c, TFB_ERR_NOMEM, "block streaming: host store");
    const int n = HostBlockStore::CHUNK;
    bool ok = cudaMalloc((void**)&st->list_dev, n * sizeof(int)) == cudaSuccess && cudaMalloc((void**)&st->flag_dev, n * sizeof(int)) == cudaSuccess &&
              cudaMalloc((void**)&st->counter_dev, sizeof(int)) == cudaSuccess &&
              cudaMalloc((void**)&st->xfer_dev, (size_t)n * BLOCK3 * sizeof(unsigned int)) == cudaSuccess &&
              cudaMallocHost((void**)&st->xfer_host, (size_t)n * BLOCK3 * sizeof(unsigned int)) == cudaSuccess &&
              cudaMallocHost((void**)&st->list_host, (2 * n + 1) * sizeof(int)) == cudaSuccess;
    c->store = st;
    if (!ok) { stream_store_free(c); return set_err(c, TFB_ERR_NOMEM, "block streaming: transfer buffers"); }
    return TFB_OK;
}

// the pose the scene was last updated with, as the device holds it (the host mirror is only refreshed by the frame path)
static int current_pose_w2c(tfb_ctx* c, float out[16]) {
    DevState ds;
    TFB_CUDA(c, cudaMemcpy(&ds, c->ds, sizeof(ds), cudaMemcpyDeviceToHost));
    memcpy(out, ds.pose_w2c, 16 * sizeof(float));
    return TFB_OK;
}

static int stream_out_impl(tfb_ctx* c, int max_blocks, int* n_out) {
    int r = tfb_sync(c);
    if (r || (r = stream_prepare(c))) return r;
    HostBlockStore* st = c->store;
    const int N = HostBlockStore::CHUNK;
    int done = 0;
    float pose[16];
    if ((r = current_pose_w2c(c, pose))) return r;
    for (;;) {
        const int want = max_blocks > 0 ? (max_blocks - done < N ? max_blocks - done : N) : N;
        if (want <= 0) break;
        if ((r = launch_stream_select(c, 0, pose, st->list_dev, want, st->counter_dev))) return r;
        TFB_CUDA(c, cudaMemcpyAsync(st->list_host + 2 * N, st->counter_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(c, cudaStreamSynchronize(c->stream));
        const int found = st->list_host[2 * N];
        const int n = found < want ? found : want;
        if (n == 0) break;
        if ((r = launch_stream_evict(c, st->list_dev, n, st->xfer_dev))) return r;
        TFB_CUDA(c, cudaMemcpyAsync(st->list_host, st->list_dev, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(c, cudaMemcpyAsync(st->xfer_host, st->xfer_dev, (size_t)n * BLOCK3 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int i = 0; i < n; ++i) {
            size_t at;
            if (!st->free_slots.empty()) { at = st->free_slots.back(); st->free_slots.pop_back(); }
            else { at = st->data.size() / BLOCK3; st->data.resize(st->data.size() + BLOCK3); }
            memcpy(&st->data[at * BLOCK3], st->xfer_host + (size_t)i * BLOCK3, BLOCK3 * sizeof(unsigned int));
            st->where[st->list_host[i]] = at;
        }
        done += n;
        if (found <= want) break;   // the sweep found no more than fitted
    }
    if (n_out) *n_out = done;
    return TFB_OK;
}

static int stream_in_impl(tfb_ctx* c, const float* pose_w2c, int all, int* n_in, int* n_left) {
    int r = tfb_sync(c);
    if (r || (r = stream_prepare(c))) return r;
    HostBlockStore* st = c->store;
    const int N = HostBlockStore::CHUNK;
    int done = 0;
    bool pool_full = false;
    float pose[16];
    if (pose_w2c) memcpy(pose, pose_w2c, sizeof(pose));
    else if ((r = current_pose_w2c(c, pose))) return r;
    while (!st->where.empty() && !pool_full) {
        if ((r = launch_stream_select(c, all ? 2 : 1, pose, st->list_dev, N, st->counter_dev))) return r;
        TFB_CUDA(c, cudaMemcpyAsync(st->list_host + 2 * N, st->counter_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(c, cudaMemcpyAsync(st->list_host, st->list_dev, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(c, cudaStreamSynchronize(c->stream));
        const int found = st->list_host[2 * N];
        const int n = found < N ? found : N;
        if (n == 0) break;
        for (int i = 0; i < n; ++i) {
            auto it = st->where.find(st->list_host[i]);
            if (it == st->where.end()) return set_err(c, TFB_ERR_STATE, "block streaming: an entry is marked swapped out but the host store does not hold it");
            memcpy(st->xfer_host + (size_t)i * BLOCK3, &st->data[it->second * BLOCK3], BLOCK3 * sizeof(unsigned int));
        }
        TFB_CUDA(c, cudaMemcpyAsync(st->xfer_dev, st->xfer_host, (size_t)n * BLOCK3 * sizeof(unsigned int), cudaMemcpyHostToDevice, c->stream));
        if ((r = launch_stream_restore(c, st->list_dev, n, st->xfer_dev, st->flag_dev))) return r;
        TFB_CUDA(c, cudaMemcpyAsync(st->list_host + N, st->flag_dev, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        TFB_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int i = 0; i < n; ++i) {
            if (!st->list_host[N + i]) { pool_full = true; continue; }   // no pool slot: stays in the store
            auto it = st->where.find(st->list_host[i]);
            st->free_slots.push_back(it->second);
            st->where.erase(it);
            ++done;
        }
        if (found <= N) break;
    }
    if (n_in) *n_in = done;
    if (n_left) *n_left = (int)st->where.size();
    return TFB_OK;
}

int tfb_stream_out(tfb_ctx* c, int max_blocks, int* n_out) {
    if (!c) return TFB_ERR_ARG;
    try {
        return stream_out_impl(c, max_blocks, n_out);
    } catch (const std::bad_alloc&) {
        return set_err(c, TFB_ERR_NOMEM, "tfb_stream_out: out of host memory");
    } catch (...) {
        return set_err(c, TFB_ERR_STATE, "tfb_stream_out: unexpected failure");
    }
}
int tfb_stream_in(tfb_ctx* c, const float* pose_w2c_or_null, int all, int* n_in, int* n_left_in_store) {
    if (!c) return TFB_ERR_ARG;
    try {
        return stream_in_impl(c, pose_w2c_or_null, all, n_in, n_left_in_store);
    } catch (const std::bad_alloc&) {
        return set_err(c, TFB_ERR_NOMEM, "tfb_stream_in: out of host memory");
    } catch (...) {
        return set_err(c, TFB_ERR_STATE, "tfb_stream_in: unexpected failure");
    }
}
int tfb_stream_stats(tfb_ctx* c, long long* blocks_in_pool, long long* blocks_in_store) {
    if (!c) return TFB_ERR_ARG;
    int r = tfb_sync(c);
    if (r) return r;
    DevState st;
    TFB_CUDA(c, cudaMemcpy(&st, c->ds, sizeof(st), cudaMemcpyDeviceToHost));
    if (blocks_in_pool) *blocks_in_pool = (long long)c->p.num_blocks - 1 - st.last_free_block;
    if (blocks_in_store) *blocks_in_store = c->store ? (long long)c->store->where.size() : 0;
    return TFB_OK;
}

}  // extern "C"

void tfb::stream_store_clear(tfb_ctx* c) {
    if (!c->store) return;
    c->store->where.clear(); c->store->data.clear(); c->store->free_slots.clear();
}
void tfb::stream_store_free(tfb_ctx* c) {
    if (!c->store) return;
    HostBlockStore* st = c->store;
    cudaFree(st->list_dev); cudaFree(st->flag_dev); cudaFree(st->counter_dev); cudaFree(st->xfer_dev);
    if (st->xfer_host) cudaFreeHost(st->xfer_host);
    if (st->list_host) cudaFreeHost(st->list_host);
    delete st;
    c->store = nullptr;
}

