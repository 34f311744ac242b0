// Depth pre-processing: raw depth -> metres, bilateral filter (+ fused ICP truncation), depth
// pyramid, vertex/normal maps, model-map pyramid.  Replaces tfusion::device::* of
// /root/reference/tfusion/src/cuda/imgproc.cu (declared in src/internal.hpp:122-132).
//
// All of these are streaming kernels bounded by HBM/L2 bandwidth (50.6 B per
// level-0 pixel).  Tiles are staged in shared memory so every input pixel is fetched from
// global memory once per CTA instead of 49 / 25 times.
//
// Compiled with --fmad=false and with accurate expf / sqrtf where the reference uses the approximate
// __expf / rsqrt: the stages are memory bound, and this keeps the pyramid (and therefore the ICP input)
// equal to the oracle's up to the last-bit difference between CUDA's and glibc's expf.
#include "tfb_common.cuh"

namespace tfb {

// ---------------------------------------------------------------------------------------------
// compute_dists (imgproc.cu:263-280) — stand-alone form; the frame path fuses it into k_bilateral
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_compute_dists(const uint16_t* __restrict__ depth, float* __restrict__ dists,
                                                       int n, int cutoff_mm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d = depth[i];
    dists[i] = (d >= cutoff_mm || d <= 0) ? -1.0f : d * 0.001f;
}

int launch_compute_dists(tfb_ctx* c, const uint16_t* depth, float* dists, int w, int h) {
    int n = w * h;
    TFB_KT(c, K_COMPUTE_DISTS);
    k_compute_dists<<<div_up(n, 256), 256, 0, c->stream>>>(depth, dists, n, c->p.depth_cutoff_mm);
    TFB_LAUNCH_CHECK(c);
    return TFB_OK;
}

// ---------------------------------------------------------------------------------------------
// bilateral filter (imgproc.cu:10-61) with the raw->metres conversion (imgproc.cu:263-280) and the
// ICP depth truncation (imgproc.cu:70-89) fused in: one read of the raw frame feeds all three.
// Window [x-R, min(x-R+k, cols-1)) x [y-R, min(y-R+k, rows-1)) — right/bottom edge exclusive, raw
// zeros take part, as in the reference.
// ---------------------------------------------------------------------------------------------
constexpr int BF_TX = 32, BF_TY = 8;

__global__ void __launch_bounds__(BF_TX* BF_TY)
    k_bilateral(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, float* __restrict__ dists, int w, int h, int ksz,
                float ss, float sd, int trunc_mm, int cutoff_mm) {
    extern __shared__ uint16_t tile[];
    const int R = ksz / 2;
    const int tw = BF_TX + 2 * R, th = BF_TY + 2 * R;
    const int x0 = blockIdx.x * BF_TX - R, y0 = blockIdx.y * BF_TY - R;
    const int tid = threadIdx.y * BF_TX + threadIdx.x;
    for (int i = tid; i < tw * th; i += BF_TX * BF_TY) {
        int tx = i % tw, ty = i / tw;
        int gx = x0 + tx, gy = y0 + ty;
        tile[i] = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? src[gy * w + gx] : (uint16_t)0;
    }
    __syncthreads();
    const int x = blockIdx.x * BF_TX + threadIdx.x, y = blockIdx.y * BF_TY + threadIdx.y;
    if (x >= w || y >= h) return;

    const int value = tile[(threadIdx.y + R) * tw + threadIdx.x + R];
    if (dists) dists[y * w + x] = (value >= cutoff_mm || value <= 0) ? -1.0f : value * 0.001f;

    const int txe = min(x - R + ksz, w - 1), tye = min(y - R + ksz, h - 1);
    float sum1 = 0.f, sum2 = 0.f;
    for (int cy = max(y - R, 0); cy < tye; ++cy) {
        const uint16_t* row = tile + (cy - y0) * tw - x0;
        const int dy2 = (y - cy) * (y - cy);
        for (int cx = max(x - R, 0); cx < txe; ++cx) {
            int depth = row[cx];
            float space2 = (float)((x - cx) * (x - cx) + dy2);
            float color2 = (float)((value - depth) * (value - depth));
            float weight = expf(-(space2 * ss + color2 * sd));  // reference: __expf (2^-21 abs. error); accurate here
            sum1 += depth * weight;
            sum2 += weight;
        }
    }
    unsigned short o = (unsigned short)__float2int_rn(sum1 / sum2);
    if (trunc_mm > 0 && o > trunc_mm) o = 0;
    dst[y * w + x] = o;
}

int launch_bilateral(tfb_ctx* c, const uint16_t* src, uint16_t* dst, int w, int h, int ksz, float sigma_spatial, float sigma_depth_m,
                     float trunc_m, float* dists_or_null) {
    if (ksz < 1 || ksz > 15 || !(ksz & 1)) return set_err(c, TFB_ERR_ARG, "bilateral kernel size must be odd and <= 15");
    float sigma_depth = sigma_depth_m * 1000;  // metres -> mm, imgproc.cu:53
    float ss = 0.5f / (sigma_spatial * sigma_spatial);
    float sd = 0.5f / (sigma_depth * sigma_depth);
    int trunc_mm = trunc_m > 0 ? (int)(unsigned short)(trunc_m * 1000.f) : 0;  // imgproc.cu:87
    int R = ksz / 2;
    size_t smem = (size_t)(BF_TX + 2 * R) * (BF_TY + 2 * R) * sizeof(uint16_t);
    dim3 block(BF_TX, BF_TY), grid(div_up(w, BF_TX), div_up(h, BF_TY));
    TFB_KT(c, K_BILATERAL);
    k_bilateral<<<grid, block, smem, c->stream>>>(src, dst, dists_or_null, w, h, ksz, ss, sd, trunc_mm, c->p.depth_cutoff_mm);
    TFB_LAUNCH_CHECK(c);
    return TFB_OK;
}

__global__ void __launch_bounds__(256) k_truncate(uint16_t* depth, int n, int max_mm) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && depth[i] > max_mm) depth[i] = 0;
}

int launch_truncate(tfb_ctx* c, uint16_t* depth, int w, int h, float max_dist) {
    int n = w * h;
    TFB_KT(c, K_TRUNCATE);
    k_truncate<<<div_up(n, 256), 256, 0, c->stream>>>(depth, n, (int)(unsigned short)(max_dist * 1000.f));
    TFB_LAUNCH_CHECK(c);
    return TFB_OK;
}

// ---------------------------------------------------------------------------------------------
// depth pyramid (imgproc.cu:98-140): integer mean of the 5x5 (edge-exclusive) window samples within
// 3 sigma of the centre, integer division.
// ---------------------------------------------------------------------------------------------
constexpr int PY_TX = 32, PY_TY = 8;

__global__ void __launch_bounds__(PY_TX* PY_TY)
    k_depth_pyr(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int sw, int sh, int dw, int dh, float thr) {
    // source tile: (2*PY_TX + 3) x (2*PY_TY + 3) starting at (2*bx*PY_TX - 2, 2*by*PY_TY - 2)
    constexpr int TW = 2 * PY_TX + 4, TH = 2 * PY_TY + 4;
    __shared__ uint16_t tile[TH][TW];
    const int x0 = 2 * blockIdx.x * PY_TX - 2, y0 = 2 * blockIdx.y * PY_TY - 2;
    const int tid = threadIdx.y * PY_TX + threadIdx.x;
    for (int i = tid; i < TW * TH; i += PY_TX * PY_TY) {
        int tx = i % TW, ty = i / TW;
        int gx = x0 + tx, gy = y0 + ty;
        tile[ty][tx] = (gx >= 0 && gx < sw && gy >= 0 && gy < sh) ? src[gy * sw + gx] : (uint16_t)0;
    }
    __syncthreads();
    const int x = blockIdx.x * PY_TX + threadIdx.x, y = blockIdx.y * PY_TY + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const int D = 5;
    const int center = tile[2 * y - y0][2 * x - x0];
    const int txe = min(2 * x - D / 2 + D, sw - 1), tye = min(2 * y - D / 2 + D, sh - 1);
    int sum = 0, count = 0;
    for (int cy = max(0, 2 * y - D / 2); cy < tye; ++cy)
        for (int cx = max(0, 2 * x - D / 2); cx < txe; ++cx) {
            int val = tile[cy - y0][cx - x0];
            if (abs(val - center) < thr) { sum += val; ++count; }
        }
    dst[y * dw + x] = (uint16_t)((count == 0) ? 0 : sum / count);
}

int launch_depth_pyr(tfb_ctx* c, const uint16_t* src, uint16_t* dst, int sw, int sh, float sigma_depth_m) {
    float thr = sigma_depth_m * 1000 * 3;  // imgproc.cu:132,138
    int dw = sw / 2, dh = sh / 2;
    dim3 block(PY_TX, PY_TY), grid(div_up(dw, PY_TX), div_up(dh, PY_TY));
    TFB_KT(c, K_DEPTH_PYR);
    k_depth_pyr<<<grid, block, 0, c->stream>>>(src, dst, sw, sh, dw, dh, thr);
    TFB_LAUNCH_CHECK(c);
    return TFB_OK;
}

// ---------------------------------------------------------------------------------------------
// vertex + normal maps (imgproc.cu:214-254): v = z * ((u-cx)/fx, (v-cy)/fy, 1),
// n = -normalize((v01 - v00) x (v10 - v00)); NaN x4 when any of the three depths is 0 or on the
// last row / column.
// ---------------------------------------------------------------------------------------------
// The list of level-0 pixels that hold a vertex, ascending, for k_icp_all (tfb_icp.cu: every CTA takes an equal share).  The
// level-0 launch of k_pyr_maps leaves one validity bit per pixel (a word per 32 pixels: 38 KB at 640x480); a tile of 64 words
// is compacted by one CTA, which finds its place in the list by counting the bits of all the words in front of its own —
// no ticket, no look-back, no waiting between CTAs, a fixed order.  The tiles ride in the rows of k_points_normals' grid past
// the image (the last launch of the preprocessing), so the list costs no launch and no event of its own.
struct ValidListArgs {
    const unsigned int* mask;   // null: no list
    int n_words;
    int* list;
    unsigned int* n_total;
    int image_rows;             // grid rows of the launch that belong to the image
};
constexpr int VL_WORDS = 64;    // words of 32 pixels per tile

__device__ __forceinline__ void valid_list_tile(const ValidListArgs& a, int tile, int tid) {
    __shared__ int s_w[8];
    __shared__ int s_off[VL_WORDS + 1];
    const int lane = tid & 31, warp = tid >> 5;
    const int w0 = tile * VL_WORDS;
    int before = 0;
    const uint4* m4 = reinterpret_cast<const uint4*>(a.mask);   // w0 is a multiple of 64 words: whole 16-byte groups
    for (int i = tid; i < w0 / 4; i += 256) {
        const uint4 q = __ldg(m4 + i);
        before += __popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0) s_w[warp] = before;
    unsigned int mine = 0;
    if (tid < VL_WORDS) {
        mine = (w0 + tid < a.n_words) ? __ldg(a.mask + w0 + tid) : 0u;
        s_off[tid] = __popc(mine);
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) run += s_w[w];
        for (int i = 0; i < VL_WORDS; ++i) { const int c = s_off[i]; s_off[i] = run; run += c; }
        s_off[VL_WORDS] = run;
        if (w0 + VL_WORDS >= a.n_words) *a.n_total = (unsigned int)run;   // the last tile knows the length
    }
    __syncthreads();
    for (int k = warp; k < VL_WORDS; k += 8) {
        if (w0 + k >= a.n_words) break;
        const unsigned int m = __ldg(a.mask + w0 + k);
        if ((m >> lane) & 1u) a.list[s_off[k] + __popc(m & ((1u << lane) - 1u))] = (w0 + k) * 32 + lane;
    }
}

__global__ void __launch_bounds__(256) k_points_normals(const uint16_t* __restrict__ depth, float4* __restrict__ points,
                                                        float4* __restrict__ normals, int w, int h, float finvx, float finvy,
                                                        float cx, float cy, ValidListArgs vl) {
    if (vl.mask != nullptr && (int)blockIdx.y >= vl.image_rows) {
        const int tile = ((int)blockIdx.y - vl.image_rows) * (int)gridDim.x + (int)blockIdx.x;
        if (tile * VL_WORDS < vl.n_words) valid_list_tile(vl, tile, threadIdx.y * 32 + threadIdx.x);
        return;
    }
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const float qnan = __int_as_float(0x7fffffff);
    float4 p = make_float4(qnan, qnan, qnan, qnan), n = p;
    if (x < w - 1 && y < h - 1) {
        float z00 = depth[y * w + x] * 0.001f;
        float z01 = depth[y * w + x + 1] * 0.001f;
        float z10 = depth[(y + 1) * w + x] * 0.001f;
        if (z00 * z01 * z10 != 0) {
            float3 v00 = make_float3(z00 * (x - cx) * finvx, z00 * (y - cy) * finvy, z00);
            float3 v01 = make_float3(z01 * (x + 1 - cx) * finvx, z01 * (y - cy) * finvy, z01);
            float3 v10 = make_float3(z10 * (x - cx) * finvx, z10 * (y + 1 - cy) * finvy, z10);
            float3 a = make_float3(v01.x - v00.x, v01.y - v00.y, v01.z - v00.z);
            float3 b = make_float3(v10.x - v00.x, v10.y - v00.y, v10.z - v00.z);
            float3 cr = make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
            float r = 1.0f / sqrtf(__fmaf_rn(cr.x, cr.x, __fmaf_rn(cr.y, cr.y, cr.z * cr.z)));  // reference: rsqrt (approximate)
            n = make_float4(-(cr.x * r), -(cr.y * r), -(cr.z * r), 1.0f);
            p = make_float4(v00.x, v00.y, v00.z, 1.0f);
        }
    }
    points[y * w + x] = p;
    normals[y * w + x] = n;
}

// One launch per pyramid level in the frame path: the CTA stages a (64+4) x (16+4) tile of level l once and produces from it
// both the 32 x 8 tile of level l+1 (pyramid_kernel, imgproc.cu:98-127) and the vertex / normal maps of its 64 x 16 pixels of
// level l (points_normals_kernel, imgproc.cu:214-243) — the reference reads level l twice, in two kernels.
__global__ void __launch_bounds__(PY_TX* PY_TY)
    k_pyr_maps(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, float4* __restrict__ points, float4* __restrict__ normals,
               int sw, int sh, int dw, int dh, float thr, float finvx, float finvy, float cx, float cy, unsigned int* __restrict__ vmask) {
    constexpr int TW = 2 * PY_TX + 4, TH = 2 * PY_TY + 4;
    __shared__ uint16_t tile[TH][TW];
    const int x0 = 2 * blockIdx.x * PY_TX - 2, y0 = 2 * blockIdx.y * PY_TY - 2;
    const int tid = threadIdx.y * PY_TX + threadIdx.x;
    for (int i = tid; i < TW * TH; i += PY_TX * PY_TY) {
        int tx = i % TW, ty = i / TW;
        int gx = x0 + tx, gy = y0 + ty;
        tile[ty][tx] = (gx >= 0 && gx < sw && gy >= 0 && gy < sh) ? src[gy * sw + gx] : (uint16_t)0;
    }
    __syncthreads();
    // level l + 1
    {
        const int x = blockIdx.x * PY_TX + threadIdx.x, y = blockIdx.y * PY_TY + threadIdx.y;
        if (x < dw && y < dh) {
            const int D = 5;
            const int center = tile[2 * y - y0][2 * x - x0];
            const int txe = min(2 * x - D / 2 + D, sw - 1), tye = min(2 * y - D / 2 + D, sh - 1);
            int sum = 0, count = 0;
            for (int cy2 = max(0, 2 * y - D / 2); cy2 < tye; ++cy2)
                for (int cx2 = max(0, 2 * x - D / 2); cx2 < txe; ++cx2) {
                    int val = tile[cy2 - y0][cx2 - x0];
                    if (abs(val - center) < thr) { sum += val; ++count; }
                }
            dst[y * dw + x] = (uint16_t)((count == 0) ? 0 : sum / count);
        }
    }
    // maps of level l: 64 x 16 pixels, four per thread, a row of 64 per two warps (coalesced 16 B stores)
    const float qnan = __int_as_float(0x7fffffff);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = tid + k * PY_TX * PY_TY;
        const int lx = i & (2 * PY_TX - 1), ly = i / (2 * PY_TX);
        const int x = 2 * blockIdx.x * PY_TX + lx, y = 2 * blockIdx.y * PY_TY + ly;
        const bool in = x < sw && y < sh;
        float4 p = make_float4(qnan, qnan, qnan, qnan), n = p;
        bool has_vertex = false;
        if (in && x < sw - 1 && y < sh - 1) {
            float z00 = tile[ly + 2][lx + 2] * 0.001f;
            float z01 = tile[ly + 2][lx + 3] * 0.001f;
            float z10 = tile[ly + 3][lx + 2] * 0.001f;
            if (z00 * z01 * z10 != 0) {
                float3 v00 = make_float3(z00 * (x - cx) * finvx, z00 * (y - cy) * finvy, z00);
                float3 v01 = make_float3(z01 * (x + 1 - cx) * finvx, z01 * (y - cy) * finvy, z01);
                float3 v10 = make_float3(z10 * (x - cx) * finvx, z10 * (y + 1 - cy) * finvy, z10);
                float3 a = make_float3(v01.x - v00.x, v01.y - v00.y, v01.z - v00.z);
                float3 b = make_float3(v10.x - v00.x, v10.y - v00.y, v10.z - v00.z);
                float3 cr = make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
                float r = 1.0f / sqrtf(__fmaf_rn(cr.x, cr.x, __fmaf_rn(cr.y, cr.y, cr.z * cr.z)));  // reference: rsqrt (approximate)
                n = make_float4(-(cr.x * r), -(cr.y * r), -(cr.z * r), 1.0f);
                p = make_float4(v00.x, v00.y, v00.z, 1.0f);
                has_vertex = true;
            }
        }
        if (in) {
            points[y * sw + x] = p;
            normals[y * sw + x] = n;
        }
        if (vmask != nullptr) {   // a warp holds 32 consecutive pixels of one row, starting at a multiple of 32 (sw % 32 == 0)
            const unsigned int m = __ballot_sync(0xffffffffu, has_vertex);
            if ((threadIdx.x & 31) == 0 && in) vmask[(y * sw + x) >> 5] = m;
        }
    }
}

int launch_pyr_maps(tfb_ctx* c, const uint16_t* src, uint16_t* dst, float4* pts, float4* nrm, int sw, int sh, float sigma_depth_m,
                    float fx, float fy, float cx, float cy, unsigned int* vmask) {
    float thr = sigma_depth_m * 1000 * 3;  // imgproc.cu:132,138
    int dw = sw / 2, dh = sh / 2;
    dim3 block(PY_TX, PY_TY), grid(div_up(sw, 2 * PY_TX), div_up(sh, 2 * PY_TY));
    TFB_KT(c, K_PYR_MAPS);
    k_pyr_maps<<<grid, block, 0, c->stream>>>(src, dst, pts, nrm, sw, sh, dw, dh, thr, 1.f / fx, 1.f / fy, cx, cy, vmask);
    TFB_LAUNCH_CHECK(c);
    return TFB_OK;
}

int launch_points_normals(tfb_ctx* c, const uint16_t* depth, float4* pts, float4* nrm, int w, int h, float fx, float fy, float cx,
                          float cy, bool with_valid_list) {
    dim3 block(32, 8), grid(div_up(w, 32), div_up(h, 8));
    ValidListArgs vl;
    vl.mask = nullptr; vl.n_words = 0; vl.list = nullptr; vl.n_total = nullptr; vl.image_rows = (int)grid.y;
    if (with_valid_list) {   // the level-0 mask was written by the first k_pyr_maps of this preprocessing (same stream)
        vl.mask = c->icp_vmask; vl.n_words = c->lv[0].w * c->lv[0].h / 32; vl.list = c->icp_vlist; vl.n_total = c->icp_vscan;
        grid.y += div_up(div_up(vl.n_words, VL_WORDS), (int)grid.x);
    }
    TFB_KT(c, K_POINTS_NORMALS);
    k_points_normals<<<grid, block, 0, c->stream>>>(depth, pts, nrm, w, h, 1.f / fx, 1.f / fy, cx, cy, vl);
    TFB_LAUNCH_CHECK(c);
    return TFB_OK;
}

// ---------------------------------------------------------------------------------------------
// model-map pyramid (imgproc.cu:355-401): 2x2 mean of points (w = 1) and normals (w = 0, not
// renormalised); NaN when any of the four source points is NaN.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resize_points_normals(const float4* __restrict__ vsrc, const float4* __restrict__ nsrc,
                                                               float4* __restrict__ vdst, float4* __restrict__ ndst, int sw, int dw,
                                                               int dh) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const float qnan = __int_as_float(0x7fffffff);
    float4 vo = make_float4(qnan, qnan, qnan, 0.f), no = vo;
    const int xs = 2 * x, ys = 2 * y;
    float4 d00 = vsrc[ys * sw + xs], d01 = vsrc[ys * sw + xs + 1];
    float4 d10 = vsrc[(ys + 1) * sw + xs], d11 = vsrc[(ys + 1) * sw + xs + 1];
    if (!isnan(d00.x * d01.x * d10.x * d11.x)) {
        vo = make_float4((d00.x + d01.x + d10.x + d11.x) * 0.25f, (d00.y + d01.y + d10.y + d11.y) * 0.25f,
                         (d00.z + d01.z + d10.z + d11.z) * 0.25f, 1.0f);
        float4 n00 = nsrc[ys * sw + xs], n01 = nsrc[ys * sw + xs + 1];
        float4 n10 = nsrc[(ys + 1) * sw + xs], n11 = nsrc[(ys + 1) * sw + xs + 1];
        no = make_float4((n00.x + n01.x + n10.x + n11.x) * 0.25f, (n00.y + n01.y + n10.y + n11.y) * 0.25f,
                         (n00.z + n01.z + n10.z + n11.z) * 0.25f, 0.f);
    }
    vdst[y * dw + x] = vo;
    ndst[y * dw + x] = no;
}

int launch_resize_points_normals(tfb_ctx* c, const float4* v, const float4* n, float4* vo, float4* no, int sw, int sh) {
    int dw = sw / 2, dh = sh / 2;
    dim3 block(32, 8), grid(div_up(dw, 32), div_up(dh, 8));
    TFB_KT(c, K_RESIZE_MAPS);
    k_resize_points_normals<<<grid, block, 0, c->stream>>>(v, n, vo, no, sw, dw, dh);
    TFB_LAUNCH_CHECK(c);
    return TFB_OK;
}

}  // namespace tfb
