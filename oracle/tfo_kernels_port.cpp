// TEST INFRASTRUCTURE (see tfo_types.h).  CPU restatement of the reference's per-pixel /
// per-voxel / per-block element functions.  Compiled with -ffp-contract=off so every float
// operation rounds exactly once, in the order the reference's host build would perform it.
// The sibling tfo_kernels_ref.cu maps the same interface onto the reference's own code;
// tests/test_oracle_vs_ref.py demands bit-equal outputs from the two.
#include "tfo_types.h"

#include <cmath>
#include <cstring>
#include <limits>

namespace tfo {
namespace k {

const char* impl_name() { return "port"; }

namespace {

struct V3 { float x, y, z; };

// column-major 4x4 times (x,y,z,w): include/Matrix.hpp:128-135 (left-to-right sums)
inline void mul4(const float m[16], float x, float y, float z, float w, float r[4]) {
    r[0] = m[0] * x + m[4] * y + m[8] * z + m[12] * w;
    r[1] = m[1] * x + m[5] * y + m[9] * z + m[13] * w;
    r[2] = m[2] * x + m[6] * y + m[10] * z + m[14] * w;
    r[3] = m[3] * x + m[7] * y + m[11] * z + m[15] * w;
}

// include/tfusion/cuda/RepresentationAccess.hpp:5-7
inline int hash_of(int bx, int by, int bz, const HashGeom& g) {
    return (int)((((unsigned)bx * 73856093u) ^ ((unsigned)by * 19349669u) ^ ((unsigned)bz * 83492791u)) &
                 (unsigned)g.hash_mask);
}

inline bool same_pos(const HashEntry& e, int bx, int by, int bz) {
    return e.pos[0] == bx && e.pos[1] == by && e.pos[2] == bz;
}

// MathUtils.hpp:19-21 ROUND then (int) truncation
inline int round_half_away(float v) { return (int)((v < 0) ? (v - 0.5f) : (v + 0.5f)); }

// RepresentationAccess.hpp:9-17: block coordinate by floor division, linear index inside it
inline int voxel_block_of(int px, int py, int pz, int& bx, int& by, int& bz) {
    bx = ((px < 0) ? px - BLOCK + 1 : px) / BLOCK;
    by = ((py < 0) ? py - BLOCK + 1 : py) / BLOCK;
    bz = ((pz < 0) ? pz - BLOCK + 1 : pz) / BLOCK;
    return (px - bx * BLOCK) + (py - by * BLOCK) * BLOCK + (pz - bz * BLOCK) * BLOCK * BLOCK;
}

struct BlockCache {  // VoxelBlockHash.hpp:58-62
    int bx = 0x7fffffff, by = 0x7fffffff, bz = 0x7fffffff;
    int base = -1;
};

// RepresentationAccess.hpp:67-100.  found: 0 = missing, 1 = cache hit ("true"), slot+1 = table hit.
inline Voxel read_voxel(const Voxel* voxels, const HashEntry* table, int px, int py, int pz, int& found,
                        BlockCache& c, const HashGeom& g) {
    int bx, by, bz;
    int lin = voxel_block_of(px, py, pz, bx, by, bz);
    if (bx == c.bx && by == c.by && bz == c.bz) {
        found = 1;
        return voxels[c.base + lin];
    }
    int slot = hash_of(bx, by, bz, g);
    for (;;) {
        const HashEntry e = table[slot];
        if (same_pos(e, bx, by, bz) && e.ptr >= 0) {
            c.bx = bx; c.by = by; c.bz = bz;
            c.base = e.ptr * BLOCK3;
            found = slot + 1;
            return voxels[c.base + lin];
        }
        if (e.offset < 1) break;
        slot = g.num_buckets + e.offset - 1;
    }
    found = 0;
    Voxel empty;
    empty.sdf = 32767; empty.w_depth = 0; empty.pad_ = 0;
    return empty;
}

inline float sdf_to_float(float raw) { return raw / 32767.0f; }

// RepresentationAccess.hpp:129-134 — nearest voxel
inline float read_nearest(const Voxel* voxels, const HashEntry* table, V3 p, int& found, BlockCache& c,
                          const HashGeom& g) {
    Voxel v = read_voxel(voxels, table, round_half_away(p.x), round_half_away(p.y), round_half_away(p.z), found, c, g);
    return sdf_to_float((float)v.sdf);
}

// RepresentationAccess.hpp:137-162 (+ :165-199 for the confidence variant): trilinear blend of the raw
// int16 values, x pairs first, then y, then z; `found` is forced to 1 afterwards.
inline float read_trilinear(const Voxel* voxels, const HashEntry* table, V3 p, int& found, BlockCache& c,
                            const HashGeom& g, float* conf) {
    float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    float cx = p.x - fx, cy = p.y - fy, cz = p.z - fz;
    int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    float s[2], w[2];
    for (int dz = 0; dz < 2; ++dz) {
        Voxel a = read_voxel(voxels, table, ix, iy, iz + dz, found, c, g);
        Voxel b = read_voxel(voxels, table, ix + 1, iy, iz + dz, found, c, g);
        float rs = (1.0f - cx) * (float)a.sdf + cx * (float)b.sdf;
        float rw = (1.0f - cx) * (float)a.w_depth + cx * (float)b.w_depth;
        a = read_voxel(voxels, table, ix, iy + 1, iz + dz, found, c, g);
        b = read_voxel(voxels, table, ix + 1, iy + 1, iz + dz, found, c, g);
        rs = (1.0f - cy) * rs + cy * ((1.0f - cx) * (float)a.sdf + cx * (float)b.sdf);
        rw = (1.0f - cy) * rw + cy * ((1.0f - cx) * (float)a.w_depth + cx * (float)b.w_depth);
        s[dz] = rs; w[dz] = rw;
    }
    found = 1;
    if (conf) *conf = (1.0f - cz) * w[0] + cz * w[1];
    return sdf_to_float((1.0f - cz) * s[0] + cz * s[1]);
}

}  // namespace

// include/Matrix.hpp:173-234 — cofactor inverse; operand order kept so rounding matches.
bool mat4_inv(const float in[16], float out[16]) {
    float s[16], t[12];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) s[i + 4 * j] = in[i * 4 + j];
    // products of the lower two source rows
    static const unsigned char P1[12][2] = {{10, 15}, {11, 14}, {9, 15}, {11, 13}, {9, 14}, {10, 13},
                                            {8, 15},  {11, 12}, {8, 14}, {10, 12}, {8, 13}, {9, 12}};
    static const unsigned char P2[12][2] = {{2, 7}, {3, 6}, {1, 7}, {3, 5}, {1, 6}, {2, 5},
                                            {0, 7}, {3, 4}, {0, 6}, {2, 4}, {0, 5}, {1, 4}};
    // dst[i] = (t[a]*s[b] + t[c]*s[d] + t[e]*s[f]) - (t[g]*s[h] + t[i]*s[j] + t[k]*s[l])
    static const unsigned char C[16][12] = {
        {0, 5, 3, 6, 4, 7, 1, 5, 2, 6, 5, 7},        {1, 4, 6, 6, 9, 7, 0, 4, 7, 6, 8, 7},
        {2, 4, 7, 5, 10, 7, 3, 4, 6, 5, 11, 7},      {5, 4, 8, 5, 11, 6, 4, 4, 9, 5, 10, 6},
        {1, 1, 2, 2, 5, 3, 0, 1, 3, 2, 4, 3},        {0, 0, 7, 2, 8, 3, 1, 0, 6, 2, 9, 3},
        {3, 0, 6, 1, 11, 3, 2, 0, 7, 1, 10, 3},      {4, 0, 9, 1, 10, 2, 5, 0, 8, 1, 11, 2},
        {0, 13, 3, 14, 4, 15, 1, 13, 2, 14, 5, 15},  {1, 12, 6, 14, 9, 15, 0, 12, 7, 14, 8, 15},
        {2, 12, 7, 13, 10, 15, 3, 12, 6, 13, 11, 15}, {5, 12, 8, 13, 11, 14, 4, 12, 9, 13, 10, 14},
        {2, 10, 5, 11, 1, 9, 4, 11, 0, 9, 3, 10},    {8, 11, 0, 8, 7, 10, 6, 10, 9, 11, 1, 8},
        {6, 9, 11, 11, 3, 8, 10, 11, 2, 8, 7, 9},    {10, 10, 4, 8, 9, 9, 8, 9, 11, 10, 5, 8}};
    auto cof = [&](int i) {
        const unsigned char* c = C[i];
        return (t[c[0]] * s[c[1]] + t[c[2]] * s[c[3]] + t[c[4]] * s[c[5]]) -
               (t[c[6]] * s[c[7]] + t[c[8]] * s[c[9]] + t[c[10]] * s[c[11]]);
    };
    for (int i = 0; i < 12; ++i) t[i] = s[P1[i][0]] * s[P1[i][1]];
    for (int i = 0; i < 4; ++i) out[i] = cof(i);
    float det = s[0] * out[0] + s[1] * out[1] + s[2] * out[2] + s[3] * out[3];
    if (det == 0.0f) return false;
    for (int i = 4; i < 8; ++i) out[i] = cof(i);
    for (int i = 0; i < 12; ++i) t[i] = s[P2[i][0]] * s[P2[i][1]];
    for (int i = 8; i < 16; ++i) out[i] = cof(i);
    float r = 1 / det;
    for (int i = 0; i < 16; ++i) out[i] *= r;
    return true;
}

// include/tfusion/cuda/SceneReconstructionEngine.hpp:206-298
void mark_pixel(uint8_t* alloc_type, uint8_t* vis_type, int x, int y, int16_t* block_coords,
                const float* dists, const float inv_m[16], const float inv_proj[4], float mu, int w, int h,
                float one_over_block_m, const HashEntry* table, float vf_min, float vf_max, const HashGeom& g) {
    (void)h;
    float d = dists[x + y * w];
    if (d <= 0 || (d - mu) < 0 || (d - mu) < vf_min || (d + mu) > vf_max) return;

    float cz = d;
    float cx = cz * (((float)x - inv_proj[2]) * inv_proj[0]);
    float cy = cz * (((float)y - inv_proj[3]) * inv_proj[1]);
    float len = sqrtf(cx * cx + cy * cy + cz * cz);

    float r[4];
    float sc = 1.0f - mu / len;
    mul4(inv_m, cx * sc, cy * sc, cz * sc, 1.0f, r);
    V3 p = {r[0] * one_over_block_m, r[1] * one_over_block_m, r[2] * one_over_block_m};
    sc = 1.0f + mu / len;
    mul4(inv_m, cx * sc, cy * sc, cz * sc, 1.0f, r);
    V3 e = {r[0] * one_over_block_m, r[1] * one_over_block_m, r[2] * one_over_block_m};

    V3 dir = {e.x - p.x, e.y - p.y, e.z - p.z};
    len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    int steps = (int)ceilf(2.0f * len);
    float den = (float)(steps - 1);
    dir.x /= den; dir.y /= den; dir.z /= den;

    for (int i = 0; i < steps; ++i) {
        int bx = (int16_t)floorf(p.x), by = (int16_t)floorf(p.y), bz = (int16_t)floorf(p.z);
        int slot = hash_of(bx, by, bz, g);
        HashEntry ent = table[slot];
        bool found = false;
        if (same_pos(ent, bx, by, bz) && ent.ptr >= -1) {
            vis_type[slot] = (ent.ptr == -1) ? 2 : 1;
            found = true;
        }
        if (!found) {
            bool in_chain = false;
            if (ent.ptr >= -1) {  // bucket head taken: walk the excess chain to its tail
                while (ent.offset >= 1) {
                    slot = g.num_buckets + ent.offset - 1;
                    ent = table[slot];
                    if (same_pos(ent, bx, by, bz) && ent.ptr >= -1) {
                        vis_type[slot] = (ent.ptr == -1) ? 2 : 1;
                        found = true;
                        break;
                    }
                }
                in_chain = true;
            }
            if (!found) {
                alloc_type[slot] = in_chain ? 2 : 1;
                if (!in_chain) vis_type[slot] = 1;
                int16_t* bc = block_coords + 4 * slot;
                bc[0] = (int16_t)bx; bc[1] = (int16_t)by; bc[2] = (int16_t)bz; bc[3] = 1;
            }
        }
        p.x += dir.x; p.y += dir.y; p.z += dir.z;
    }
}

// SceneReconstructionEngine.hpp:300-375 (useSwapping = false)
bool block_visible(const int16_t pos[3], const float m[16], const float proj[4], float voxel_size, int w, int h) {
    float f = (float)BLOCK * voxel_size;
    float p[3] = {(float)pos[0] * f, (float)pos[1] * f, (float)pos[2] * f};
    // corner walk 000 001 011 111 110 100 010 101, accumulated in place
    static const signed char step[8][3] = {{0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {1, 0, 0},
                                           {0, 0, -1}, {0, -1, 0}, {-1, 1, 0}, {1, -1, 1}};
    for (int c = 0; c < 8; ++c) {
        // multi-axis moves apply x, then y, then z (each a separate rounded add)
        if (step[c][0] > 0) p[0] += f; else if (step[c][0] < 0) p[0] -= f;
        if (step[c][1] > 0) p[1] += f; else if (step[c][1] < 0) p[1] -= f;
        if (step[c][2] > 0) p[2] += f; else if (step[c][2] < 0) p[2] -= f;
        float r[4];
        mul4(m, p[0], p[1], p[2], 1.0f, r);
        if (r[2] < 1e-10f) continue;
        float u = proj[0] * r[0] / r[2] + proj[2];
        float v = proj[1] * r[1] / r[2] + proj[3];
        if (u >= 0 && u < w && v >= 0 && v < h) return true;
    }
    return false;
}

// SceneReconstructionEngine.hpp:23-71
void update_voxel(Voxel& v, const float pt[4], const float m[16], const float proj[4], float mu, int max_w,
                  const float* dists, int w, int h) {
    float r[4];
    mul4(m, pt[0], pt[1], pt[2], pt[3], r);
    if (r[2] <= 0) return;
    float u = proj[0] * r[0] / r[2] + proj[2];
    float vv = proj[1] * r[1] / r[2] + proj[3];
    if ((u < 1) || (u > w - 2) || (vv < 1) || (vv > h - 2)) return;
    float dm = dists[(int)(u + 0.5f) + (int)(vv + 0.5f) * w];
    if (dm <= 0.0f) return;
    float eta = dm - r[2];
    if (eta < -mu) return;
    float old_f = (float)v.sdf / 32767.0f;
    int old_w = v.w_depth;
    float new_f = eta / mu;
    new_f = (1.0f < new_f) ? 1.0f : new_f;
    int new_w = 1;
    new_f = old_w * old_f + new_w * new_f;
    new_w = old_w + new_w;
    new_f /= new_w;
    new_w = (new_w < max_w) ? new_w : max_w;
    v.sdf = (int16_t)(new_f * 32767.0f);
    v.w_depth = (uint8_t)new_w;
}

// VisualisationEngine_Shared.hpp:33-75
bool project_block(const int16_t pos[3], const float m[16], const float proj[4], int w, int h, float voxel_size,
                   int ul[2], int lr[2], float z[2]) {
    ul[0] = w / MINMAX_SUBSAMPLE; ul[1] = h / MINMAX_SUBSAMPLE;
    lr[0] = -1; lr[1] = -1;
    z[0] = FAR_AWAY_F; z[1] = VERY_CLOSE_F;
    for (int c = 0; c < 8; ++c) {
        int16_t q[3] = {pos[0], pos[1], pos[2]};
        q[0] += (c & 1) ? 1 : 0;
        q[1] += (c & 2) ? 1 : 0;
        q[2] += (c & 4) ? 1 : 0;
        float r[4];
        mul4(m, (float)q[0] * (float)BLOCK * voxel_size, (float)q[1] * (float)BLOCK * voxel_size,
             (float)q[2] * (float)BLOCK * voxel_size, 1.0f, r);
        if (r[2] < 1e-6) continue;
        float px = (proj[0] * r[0] / r[2] + proj[2]) / MINMAX_SUBSAMPLE;
        float py = (proj[1] * r[1] / r[2] + proj[3]) / MINMAX_SUBSAMPLE;
        if (ul[0] > floorf(px)) ul[0] = (int)floorf(px);
        if (lr[0] < ceilf(px)) lr[0] = (int)ceilf(px);
        if (ul[1] > floorf(py)) ul[1] = (int)floorf(py);
        if (lr[1] < ceilf(py)) lr[1] = (int)ceilf(py);
        if (z[0] > r[2]) z[0] = r[2];
        if (z[1] < r[2]) z[1] = r[2];
    }
    if (ul[0] < 0) ul[0] = 0;
    if (ul[1] < 0) ul[1] = 0;
    if (lr[0] >= w) lr[0] = w - 1;
    if (lr[1] >= h) lr[1] = h - 1;
    if (ul[0] > lr[0]) return false;
    if (ul[1] > lr[1]) return false;
    if (z[0] < VERY_CLOSE_F) z[0] = VERY_CLOSE_F;
    if (z[1] < VERY_CLOSE_F) return false;
    return true;
}

// VisualisationEngine_Shared.hpp:77-95
int split_tiles(RenderTile* list, int offset, const int ul[2], const int lr[2], const float z[2]) {
    int ny = (int)ceilf((float)(1 + lr[1] - ul[1]) / TILE);
    int nx = (int)ceilf((float)(1 + lr[0] - ul[0]) / TILE);
    for (int by = 0; by < ny; ++by)
        for (int bx = 0; bx < nx; ++bx) {
            if (offset >= MAX_TILES) return offset;
            RenderTile& t = list[offset++];
            t.ul[0] = (int16_t)(ul[0] + bx * TILE);
            t.ul[1] = (int16_t)(ul[1] + by * TILE);
            t.lr[0] = (int16_t)(ul[0] + (bx + 1) * TILE - 1);
            t.lr[1] = (int16_t)(ul[1] + (by + 1) * TILE - 1);
            if (t.lr[0] > lr[0]) t.lr[0] = (int16_t)lr[0];
            if (t.lr[1] > lr[1]) t.lr[1] = (int16_t)lr[1];
            t.z[0] = z[0]; t.z[1] = z[1];
        }
    return offset;
}

// debug only (tools/ray_stats.py): per-ray step / missing-block step / trilinear-read counts of the march
static int* g_ray_stats = nullptr;
static int g_ray_stats_cols = 0;
extern "C" __attribute__((visibility("default"))) void tfo_debug_ray_stats(int* buf, int cols) { g_ray_stats = buf; g_ray_stats_cols = cols; }

// VisualisationEngine_Shared.hpp:99-172 (modifyVisibleEntries = vis_type != nullptr)
bool cast_ray(float out[4], uint8_t* vis_type, int x, int y, const Voxel* voxels, const HashEntry* table,
              const float inv_m[16], const float inv_proj[4], float one_over_voxel, float mu,
              const float minmax[2], const HashGeom& g) {
    float step_scale = mu * one_over_voxel;
    float r[4];

    float cz = minmax[0];
    float cx = cz * (((float)x + inv_proj[2]) * inv_proj[0]);
    float cy = cz * (((float)y + inv_proj[3]) * inv_proj[1]);
    float total = sqrtf(cx * cx + cy * cy + cz * cz) * one_over_voxel;
    mul4(inv_m, cx, cy, cz, 1.0f, r);
    V3 s = {r[0] * one_over_voxel, r[1] * one_over_voxel, r[2] * one_over_voxel};

    cz = minmax[1];
    cx = cz * (((float)x + inv_proj[2]) * inv_proj[0]);
    cy = cz * (((float)y + inv_proj[3]) * inv_proj[1]);
    float total_max = sqrtf(cx * cx + cy * cy + cz * cz) * one_over_voxel;
    mul4(inv_m, cx, cy, cz, 1.0f, r);
    V3 e = {r[0] * one_over_voxel, r[1] * one_over_voxel, r[2] * one_over_voxel};

    V3 dir = {e.x - s.x, e.y - s.y, e.z - s.z};
    float inv_len = 1.0f / sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    dir.x *= inv_len; dir.y *= inv_len; dir.z *= inv_len;

    V3 p = s;
    BlockCache cache;
    float sdf = 1.0f, conf = 0.0f, step;
    int found;
    int n_steps = 0, n_missing = 0, n_tri = 0;
    while (total < total_max) {
        sdf = read_nearest(voxels, table, p, found, cache, g);
        ++n_steps;
        if (vis_type && found) vis_type[found - 1] = 1;
        if (!found) {
            step = BLOCK;
            ++n_missing;
        } else {
            if ((sdf <= 0.1f) && (sdf >= -0.5f)) { sdf = read_trilinear(voxels, table, p, found, cache, g, nullptr); ++n_tri; }
            if (sdf <= 0.0f) break;
            step = sdf * step_scale;
            step = (step < 1.0f) ? 1.0f : step;
        }
        p.x += step * dir.x; p.y += step * dir.y; p.z += step * dir.z;
        total += step;
    }
    bool hit;
    if (sdf <= 0.0f) {
        step = sdf * step_scale;
        p.x += step * dir.x; p.y += step * dir.y; p.z += step * dir.z;
        sdf = read_trilinear(voxels, table, p, found, cache, g, &conf);
        step = sdf * step_scale;
        p.x += step * dir.x; p.y += step * dir.y; p.z += step * dir.z;
        hit = true;
    } else {
        hit = false;
    }
    if (g_ray_stats) {
        int* st = g_ray_stats + 3 * (x + y * g_ray_stats_cols);
        st[0] = n_steps; st[1] = n_missing; st[2] = n_tri;
    }
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
    out[3] = hit ? conf + 1.0f : 0.0f;
    return hit;
}

// VisualisationEngine_Shared.hpp:205-270 (useSmoothing=false, flipNormals=false) + :355-397.
// The reference writes NaN on the device and 0 on the host for invalid pixels; the oracle
// follows the device behaviour (SURVEY.md §8c caveat).
void icp_map_pixel(float* points, float* normals, const float* ray, int w, int h, int x, int y, float voxel_size,
                   const float light[3]) {
    int id = x + y * w;
    const float* p = ray + 4 * id;
    bool ok = p[3] > 0.0f;
    float n[3] = {0, 0, 0};
    if (ok) {
        if (y <= 1 || y >= h - 2 || x <= 1 || x >= w - 2) ok = false;
    }
    if (ok) {
        const float* xp = ray + 4 * ((x + 1) + y * w);
        const float* yp = ray + 4 * (x + (y + 1) * w);
        const float* xm = ray + 4 * ((x - 1) + y * w);
        const float* ym = ray + 4 * (x + (y - 1) * w);
        if (xp[3] <= 0 || yp[3] <= 0 || xm[3] <= 0 || ym[3] <= 0) {
            ok = false;
        } else {
            float dx[3] = {xp[0] - xm[0], xp[1] - xm[1], xp[2] - xm[2]};
            float dy[3] = {yp[0] - ym[0], yp[1] - ym[1], yp[2] - ym[2]};
            // the "too far apart" retry re-tests the same neighbours when smoothing is off: a no-op
            n[0] = -(dx[1] * dy[2] - dx[2] * dy[1]);
            n[1] = -(dx[2] * dy[0] - dx[0] * dy[2]);
            n[2] = -(dx[0] * dy[1] - dx[1] * dy[0]);
            float sc = 1.0f / sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            n[0] *= sc; n[1] *= sc; n[2] *= sc;
            float ang = n[0] * light[0] + n[1] * light[1] + n[2] * light[2];
            if (!(ang > 0.0)) ok = false;
        }
    }
    float* op = points + 4 * id;
    float* on = normals + 4 * id;
    if (ok) {
        op[0] = p[0] * voxel_size; op[1] = p[1] * voxel_size; op[2] = p[2] * voxel_size; op[3] = 1.0f;
        on[0] = n[0]; on[1] = n[1]; on[2] = n[2]; on[3] = 1.0f;
    } else {
        float q = std::numeric_limits<float>::quiet_NaN();
        for (int i = 0; i < 4; ++i) op[i] = on[i] = q;
    }
}

// computeSingleNormalFromSDF (RepresentationAccess.hpp:340-453): per axis A (other axes B, C) the difference of the
// trilinear SDF one voxel ahead and behind, each face blended in the order (b,c) = (0,0),(1,0),(0,1),(1,1).
static inline float face(const Voxel* voxels, const HashEntry* table, const HashGeom& g, int ix, int iy, int iz, int axis, int off,
                         float cB, float cC) {
    float v[4];
    for (int i = 0; i < 4; ++i) {
        int b = i & 1, c = i >> 1;
        int d[3];
        if (axis == 0) { d[0] = off; d[1] = b; d[2] = c; }
        else if (axis == 1) { d[0] = b; d[1] = off; d[2] = c; }
        else { d[0] = b; d[1] = c; d[2] = off; }
        int found; BlockCache fresh;
        v[i] = (float)read_voxel(voxels, table, ix + d[0], iy + d[1], iz + d[2], found, fresh, g).sdf;
    }
    float nB = 1.0f - cB, nC = 1.0f - cC;
    return v[0] * nB * nC + v[1] * cB * nC + v[2] * nB * cC + v[3] * cB * cC;
}

// processPixelGrey + computeNormalAndAngle<TVoxel,TIndex> + drawPixelGrey (VisualisationEngine_Shared.hpp:189-203,272-276,450-462)
void shade_pixel_grey(uint8_t out[4], const float ray[4], const Voxel* voxels, const HashEntry* table, const float light[3],
                      const HashGeom& g) {
    bool ok = ray[3] > 0;
    float ang = 0.f;
    if (ok) {
        float f[3] = {floorf(ray[0]), floorf(ray[1]), floorf(ray[2])};
        float c[3] = {ray[0] - f[0], ray[1] - f[1], ray[2] - f[2]};
        int ix = (int)f[0], iy = (int)f[1], iz = (int)f[2];
        float n[3];
        for (int axis = 0; axis < 3; ++axis) {
            float cA = c[axis], nA = 1.0f - cA;
            float cB = (axis == 0) ? c[1] : c[0];
            float cC = (axis == 2) ? c[1] : c[2];
            float p1 = face(voxels, table, g, ix, iy, iz, axis, 0, cB, cC);
            float p2 = face(voxels, table, g, ix, iy, iz, axis, -1, cB, cC);
            float v1 = p1 * cA + p2 * nA;
            p1 = face(voxels, table, g, ix, iy, iz, axis, 1, cB, cC);
            p2 = face(voxels, table, g, ix, iy, iz, axis, 2, cB, cC);
            n[axis] = (p1 * nA + p2 * cA - v1) / 32767.0f;
        }
        float sc = 1.0f / sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        n[0] *= sc; n[1] *= sc; n[2] *= sc;
        ang = n[0] * light[0] + n[1] * light[1] + n[2] * light[2];
        if (!(ang > 0.0)) ok = false;
    }
    uint8_t v = 0;
    if (ok) v = (uint8_t)((0.8f * ang + 0.2f) * 255.0f);
    out[0] = out[1] = out[2] = out[3] = v;
}

}  // namespace k
}  // namespace tfo
