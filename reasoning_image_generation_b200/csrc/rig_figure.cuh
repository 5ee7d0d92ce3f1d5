// rig_figure.cuh -- k_figure: the sm_100a kernel of the figure (matplotlib/Agg) path.
//
// Persistent warps, one band of FIG_BAND_ROWS canvas rows per warp at a time, exactly like k_raster:
//   * bands no layer reaches are written by the TMA engine straight from a constant white template (one instruction);
//   * other bands are built in shared memory: for every layer (one rasteriser pass of one artist) that reaches the band,
//       scatter   lanes take the layer's edges (16-byte coalesced loads; only the edges the slab directory says can reach
//                 the band), walk Agg's row / cell split of each edge and add (cover, area) to per-cell accumulators in
//                 shared memory with atomics -- the sums are commutative, so lane order does not matter;
//       resolve   row by row, the 32-cell groups that were touched are visited left to right; a warp scan turns the
//                 cells' covers into the running cover, every lane computes its pixel's 8-bit coverage and blends the
//                 layer colour with matplotlib's blender; runs between groups with a non-zero running cover (polygon
//                 interiors) are filled with their constant coverage; the accumulators are left zeroed;
//     then fence.proxy.async + one cp.async.bulk shared -> global store of the whole band.
// Pixels leave the SM once, as full lines; the roofline that bounds the kernel is HBM write bandwidth (3*W*H bytes/image).
#pragma once
#include <cuda_runtime.h>

#include "rig_figure_core.h"
#include "rig_tma.cuh"

namespace rig {

#ifndef FIG_BAND_ROWS
#define FIG_BAND_ROWS 4
#endif
constexpr int FIG_ROWS = FIG_BAND_ROWS;
constexpr int FIG_MAX_WARPS = 32;

struct FigureParams {
    const uint32_t *image_layer_begin;   // [n_images + 1]
    const FLayer *layers;
    const FEdge *edges;
    const uint32_t *dir;
    const uint32_t *order;               // [n_images] image ids, costliest first (bands are claimed in this order)
    uint8_t *out;
    uint32_t *work;                      // band claim counter
    uint32_t n_images;
    int W, H;
    uint32_t bg;                         // c0 | c1 << 8 | c2 << 16
    uint32_t max_claim;
};

// bytes of shared memory one warp needs: band pixels, cover + area accumulators, touched-group masks
__host__ __device__ inline size_t figure_warp_bytes(int W)
{
    return (size_t)FIG_ROWS * ((size_t)3 * W + (size_t)8 * W) + (size_t)FIG_ROWS * 8;
}
inline size_t figure_smem_bytes(int W, int nwarps) { return (size_t)3 * W * FIG_ROWS + (size_t)nwarps * figure_warp_bytes(W); }

struct FigAcc {
    int *cover, *area;
    unsigned *gmask;     // 2 words per band row: bit g = cells [32 g, 32 g + 32) of the row were touched
    int W, r0, off;
    unsigned *gm;
    __device__ __forceinline__ void row(int ey)
    {
        off = (ey - r0) * W;
        gm = gmask + (ey - r0) * 2;
    }
    __device__ __forceinline__ void operator()(int ex, int dc, int da)
    {
        if ((unsigned)ex >= (unsigned)W || (dc | da) == 0) return;
        atomicAdd(cover + off + ex, dc);
        atomicAdd(area + off + ex, da);
        atomicOr(gm + (ex >> 10), 1u << ((ex >> 5) & 31));
    }
};

__device__ __forceinline__ void fig_put3(uint8_t *p, uint32_t c)
{
    p[0] = (uint8_t)c; p[1] = (uint8_t)(c >> 8); p[2] = (uint8_t)(c >> 16);
}
__device__ __forceinline__ uint32_t fig_get3(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16); }

__global__ void __launch_bounds__(FIG_MAX_WARPS * 32, 1) k_figure(const FigureParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = P.W;
    const uint32_t band_bytes = (uint32_t)(3 * W * FIG_ROWS);
    uint8_t *tmpl = smem;
    uint8_t *wbase = smem + band_bytes + (size_t)warp * figure_warp_bytes(W);
    uint8_t *mine = wbase;
    int *cover = (int *)(wbase + band_bytes);
    int *area = cover + FIG_ROWS * W;
    unsigned *gmask = (unsigned *)(area + FIG_ROWS * W);

    // ---- CTA prologue: white template; every warp zeroes its accumulators once (resolve leaves them zeroed)
    {
        const uint32_t c0 = P.bg & 0xff, c1 = (P.bg >> 8) & 0xff, c2 = (P.bg >> 16) & 0xff;
        const uint32_t w0 = c0 | (c1 << 8) | (c2 << 16) | (c0 << 24);
        const uint32_t w1 = c1 | (c2 << 8) | (c0 << 16) | (c1 << 24);
        const uint32_t w2 = c2 | (c0 << 8) | (c1 << 16) | (c2 << 24);
        uint32_t *t0 = (uint32_t *)tmpl;
        for (uint32_t i = threadIdx.x; i < band_bytes / 4; i += blockDim.x) {
            const uint32_t m = i % 3;
            t0[i] = m == 0 ? w0 : (m == 1 ? w1 : w2);
        }
        for (int i = lane; i < 2 * FIG_ROWS * W; i += 32) cover[i] = 0;     // cover and area are contiguous
        if (lane < 2 * FIG_ROWS) gmask[lane] = 0u;
        fence_async_smem();
        __syncthreads();
    }
    const int bands_per_image = (P.H + FIG_ROWS - 1) / FIG_ROWS;
    const uint32_t total_bands = P.n_images * (uint32_t)bands_per_image;
    bool dirty = false;
    uint32_t cur_slot = 0xffffffffu, image = 0, lb = 0, le = 0;
    int my_ymin = 1, my_ymax = 0;      // row range of the image's lane-th layer

    for (;;) {
        uint32_t first = 0, take = 1;
        if (lane == 0) {
            const uint32_t cur = *(volatile uint32_t *)P.work;
            const uint32_t left = cur < total_bands ? total_bands - cur : 0;
            take = left / (4u * gridDim.x * (uint32_t)nwarps);
            take = take < 1u ? 1u : (take > P.max_claim ? P.max_claim : take);
            first = atomicAdd(P.work, take);
        }
        first = __shfl_sync(0xffffffffu, first, 0);
        take = __shfl_sync(0xffffffffu, take, 0);
        if (first >= total_bands) break;
        for (uint32_t bi = first; bi < first + take && bi < total_bands; bi++) {
            const uint32_t slot = bi / bands_per_image;
            if (slot != cur_slot) {
                cur_slot = slot;
                image = P.order[slot];
                lb = P.image_layer_begin[image];
                le = P.image_layer_begin[image + 1];
                my_ymin = 1; my_ymax = 0;
                if (lb + lane < le) {
                    const FLayer *L = P.layers + lb + lane;
                    my_ymin = L->ymin; my_ymax = L->ymax;
                }
            }
            const int band = (int)(bi - slot * bands_per_image);
            const int r0 = band * FIG_ROWS;
            int r1 = r0 + FIG_ROWS - 1;
            if (r1 > P.H - 1) r1 = P.H - 1;
            const uint32_t out_bytes = (uint32_t)(3 * W * (r1 - r0 + 1));
            uint8_t *gdst = P.out + ((size_t)image * P.H + r0) * (size_t)W * 3;
            bool more = false;                                    // images with more than 32 layers: the rest from memory
            for (uint32_t l = lb + 32 + lane; l < le; l += 32) more |= P.layers[l].ymin <= r1 && P.layers[l].ymax >= r0;
            const uint32_t lmask = __ballot_sync(0xffffffffu, my_ymin <= r1 && my_ymax >= r0);
            if (lmask == 0u && !__any_sync(0xffffffffu, more)) {
                if (lane == 0) bulk_store(gdst, tmpl, out_bytes);
                continue;
            }
            if (dirty) {
                if (lane == 0) bulk_wait_read();
                __syncwarp();
            }
            {
                const uint4 *src = (const uint4 *)tmpl;
                uint4 *dst = (uint4 *)mine;
                for (uint32_t i = lane; i < band_bytes / 16; i += 32) dst[i] = src[i];
            }
            __syncwarp();
            for (uint32_t l = lb; l < le; l++) {
                if (l - lb < 32u) {
                    if (!((lmask >> (l - lb)) & 1u)) continue;
                }
                const FLayer L = P.layers[l];
                if (L.ymin > r1 || L.ymax < r0) continue;
                // ---- scatter: the edges that can reach rows [r0, r1]
                {
                    int lo_row = r0 - (int)L.maxspan;
                    if (lo_row < L.ymin) lo_row = L.ymin;
                    const int s_lo = (lo_row - L.ymin) / FIG_SLAB_ROWS, s_hi = (r1 - L.ymin) / FIG_SLAB_ROWS + 1;
                    const uint32_t e_lo = P.dir[L.dir_begin + s_lo];
                    const uint32_t e_hi = s_hi >= (int)L.ndir ? L.n_edges : P.dir[L.dir_begin + s_hi];
                    FigAcc acc;
                    acc.cover = cover; acc.area = area; acc.gmask = gmask; acc.W = W; acc.r0 = r0; acc.off = 0; acc.gm = gmask;
                    const FEdge *ep = P.edges + L.edge_begin;
                    for (uint32_t e = e_lo + lane; e < e_hi; e += 32) {
                        const int4 v = __ldg((const int4 *)(ep + e));
                        FEdge ed;
                        ed.x1 = v.x; ed.y1 = v.y; ed.x2 = v.z; ed.y2 = v.w;
                        fig_edge_rows(ed, r0, r1, acc);
                    }
                }
                __syncwarp();
                // ---- resolve
                const int ya = L.ymin > r0 ? L.ymin : r0, yb = L.ymax < r1 ? L.ymax : r1;
                for (int y = ya; y <= yb; y++) {
                    const int ro = (y - r0) * W;
                    unsigned long long gm = (unsigned long long)gmask[(y - r0) * 2] | ((unsigned long long)gmask[(y - r0) * 2 + 1] << 32);
                    if (gm == 0ull) continue;
                    __syncwarp();
                    if (lane < 2) gmask[(y - r0) * 2 + lane] = 0u;
                    uint8_t *rowpx = mine + (size_t)ro * 3;
                    int carry = 0, prev_end = 0;
                    while (gm) {
                        const int g = __ffsll((long long)gm) - 1;
                        gm &= gm - 1;
                        const int xs = g * 32;
                        if (carry != 0 && xs > prev_end) {          // interior run at constant coverage
                            const uint32_t al = fig_alpha(carry, 0);
                            for (int x = prev_end + lane; x < xs; x += 32) fig_put3(rowpx + x * 3, fig_blend_px(fig_get3(rowpx + x * 3), L.rgba, al));
                        }
                        const int x = xs + lane;
                        const bool in = x < W;
                        const int c = in ? cover[ro + x] : 0, a = in ? area[ro + x] : 0;
                        if (in && (c | a)) { cover[ro + x] = 0; area[ro + x] = 0; }
                        int ci = c;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int t = __shfl_up_sync(0xffffffffu, ci, o);
                            if (lane >= o) ci += t;
                        }
                        const uint32_t al = fig_alpha(carry + ci, a);
                        if (in && al) fig_put3(rowpx + x * 3, fig_blend_px(fig_get3(rowpx + x * 3), L.rgba, al));
                        carry += __shfl_sync(0xffffffffu, ci, 31);
                        prev_end = xs + 32;
                    }
                    if (carry != 0 && prev_end < W) {               // polygon clipped by the right canvas edge
                        const uint32_t al = fig_alpha(carry, 0);
                        for (int x = prev_end + lane; x < W; x += 32) fig_put3(rowpx + x * 3, fig_blend_px(fig_get3(rowpx + x * 3), L.rgba, al));
                    }
                }
                __syncwarp();
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bulk_store(gdst, mine, out_bytes);
            dirty = true;
        }
    }
    if (lane == 0) bulk_wait_read();
}

}  // namespace rig
