// rig_figure.cuh -- k_figure: the sm_100a kernel of the figure (matplotlib/Agg) path.
//
// Persistent warps, one band of FIG_BAND_ROWS canvas rows per warp at a time, exactly like k_raster:
//   * bands no layer reaches are written by the TMA engine straight from a constant white template (one instruction);
//   * other bands are built in shared memory: for every layer (one rasteriser pass of one artist) that reaches the band,
//       scatter   lanes take the layer's edges (16-byte coalesced loads; only the edges the slab directory says can reach
//                 the band), walk Agg's row / cell split of each edge and add (cover, area) to per-cell accumulators in
//                 shared memory with atomics -- the sums are commutative, so lane order does not matter;
//       resolve   two rows at a time (one per half-warp), the 16-cell groups that were touched are visited left to right; a scan
//                 over the half-warp turns the cells' covers into the running cover, every lane computes its pixel's 8-bit coverage and blends the
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
#define FIG_BAND_ROWS 2
#endif
#ifndef FIG_ACC_CELLS
#define FIG_ACC_CELLS 256
#endif
constexpr int FIG_ROWS = FIG_BAND_ROWS;
constexpr int FIG_ACC = FIG_ACC_CELLS;     // accumulator columns per row; wider layers are resolved in column chunks (multiple of 32, <= 512)
static_assert(FIG_ACC % 32 == 0 && FIG_ACC <= 512, "one touched-group bit per 16 cells must fit a 32-bit mask");
constexpr int FIG_MAX_WARPS = 32;
constexpr uint32_t FIG_NO_ROWS = 0x80007fffu;   // ymin = 32767, ymax = -32768: overlaps no band (a "1 > 0" range would overlap rows 0..1)

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
    uint8_t *blank;                      // [n_images * bands per image] or NULL: 1 = the band holds only background pixels
};

// bytes of shared memory one warp needs: band pixels, cover + area accumulators, per-row touched-group mask and carry-in
__host__ __device__ inline size_t figure_warp_bytes(int W)
{
    return (size_t)FIG_ROWS * ((size_t)3 * W + (size_t)8 * FIG_ACC) + (size_t)FIG_ROWS * 8;
}
inline size_t figure_smem_bytes(int W, int nwarps) { return (size_t)3 * W * FIG_ROWS + 1024 + (size_t)nwarps * figure_warp_bytes(W); }

struct FigAcc {
    int *cover, *area;
    unsigned *gmask;     // per band row: bit g = cells [x0 + 16 g, x0 + 16 g + 16) were touched
    int *carry_in;       // per band row: cover of the cells left of x0
    int x0, r0, off, row_i;
    __device__ __forceinline__ void row(int ey)
    {
        row_i = ey - r0;
        off = row_i * FIG_ACC;
    }
    __device__ __forceinline__ void operator()(int ex, int dc, int da)
    {
        if ((dc | da) == 0) return;
        const int cx = ex - x0;
        if (cx < 0) { atomicAdd(carry_in + row_i, dc); return; }
        if (cx >= FIG_ACC) return;
        atomicAdd(cover + off + cx, dc);
        atomicAdd(area + off + cx, da);
        atomicOr(gmask + row_i, 1u << (cx >> 4));
    }
};

__device__ __forceinline__ void fig_put3(uint8_t *p, uint32_t c)
{
    p[0] = (uint8_t)c; p[1] = (uint8_t)(c >> 8); p[2] = (uint8_t)(c >> 16);
}
__device__ __forceinline__ uint32_t fig_get3(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16); }

// the 7 used words of a layer header, kept one layer per lane and handed round by shuffles
struct FigHdr {
    uint32_t edge_begin, n_edges, rgba, yy, xx, dir_begin, sd;   // yy = ymin | ymax << 16, xx = xmin | xmax << 16, sd = maxspan | ndir << 16
};
__device__ __forceinline__ FigHdr fig_load_hdr(const FLayer *L)
{
    const uint4 a = __ldg((const uint4 *)L), b = __ldg((const uint4 *)L + 1);
    FigHdr h;
    h.edge_begin = a.x; h.n_edges = a.y; h.rgba = a.z; h.yy = a.w; h.xx = b.x; h.dir_begin = b.y; h.sd = b.z;
    return h;
}
__device__ __forceinline__ FigHdr fig_shfl_hdr(const FigHdr &m, int src)
{
    FigHdr h;
    h.edge_begin = __shfl_sync(0xffffffffu, m.edge_begin, src); h.n_edges = __shfl_sync(0xffffffffu, m.n_edges, src);
    h.rgba = __shfl_sync(0xffffffffu, m.rgba, src); h.yy = __shfl_sync(0xffffffffu, m.yy, src);
    h.xx = __shfl_sync(0xffffffffu, m.xx, src); h.dir_begin = __shfl_sync(0xffffffffu, m.dir_begin, src);
    h.sd = __shfl_sync(0xffffffffu, m.sd, src);
    return h;
}

__global__ void __launch_bounds__(FIG_MAX_WARPS * 32, 1) k_figure(const FigureParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, hl = lane & 15;
    const int W = P.W;
    const uint32_t band_bytes = (uint32_t)(3 * W * FIG_ROWS);
    uint8_t *tmpl = smem;
    const uint32_t *magic = (const uint32_t *)(smem + band_bytes);       // 256 division multipliers (fig_div_magic)
    uint8_t *wbase = smem + band_bytes + 1024 + (size_t)warp * figure_warp_bytes(W);
    uint8_t *mine = wbase;
    int *cover = (int *)(wbase + band_bytes);
    int *area = cover + FIG_ROWS * FIG_ACC;
    unsigned *gmask = (unsigned *)(area + FIG_ROWS * FIG_ACC);
    int *carry_in = (int *)(gmask + FIG_ROWS);

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
        for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) ((uint32_t *)(smem + band_bytes))[i] = fig_div_magic(i);
        for (int i = lane; i < 2 * FIG_ROWS * FIG_ACC; i += 32) cover[i] = 0;     // cover and area are contiguous
        if (lane < 2 * FIG_ROWS) gmask[lane] = 0u;                                 // masks and carry-ins
        fence_async_smem();
        __syncthreads();
    }
    const int bands_per_image = (P.H + FIG_ROWS - 1) / FIG_ROWS;
    const uint32_t total_bands = P.n_images * (uint32_t)bands_per_image;
    bool dirty = false;
    uint32_t cur_slot = 0xffffffffu, image = 0, lb = 0, le = 0;
    FigHdr myh;                         // header of the image's lane-th layer
    myh.edge_begin = myh.n_edges = myh.rgba = myh.xx = myh.dir_begin = myh.sd = 0u;
    myh.yy = FIG_NO_ROWS;

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
        uint32_t slot = first / bands_per_image;
        int band = (int)(first - slot * bands_per_image) - 1;
        for (uint32_t bi = first; bi < first + take && bi < total_bands; bi++) {
            if (++band == bands_per_image) { band = 0; slot++; }
            if (slot != cur_slot) {
                cur_slot = slot;
                image = P.order[slot];
                lb = P.image_layer_begin[image];
                le = P.image_layer_begin[image + 1];
                myh.yy = FIG_NO_ROWS;
                myh.n_edges = 0u;
                if (lb + lane < le) myh = fig_load_hdr(P.layers + lb + lane);
            }
            const int r0 = band * FIG_ROWS;
            int r1 = r0 + FIG_ROWS - 1;
            if (r1 > P.H - 1) r1 = P.H - 1;
            const uint32_t out_bytes = (uint32_t)(3 * W * (r1 - r0 + 1));
            uint8_t *gdst = P.out + ((size_t)image * P.H + r0) * (size_t)W * 3;
            bool more = false;                                    // images with more than 32 layers: the rest from memory
            for (uint32_t l = lb + 32 + lane; l < le; l += 32) more |= P.layers[l].ymin <= r1 && P.layers[l].ymax >= r0;
            const uint32_t lmask = __ballot_sync(0xffffffffu, (int)(int16_t)(myh.yy & 0xffffu) <= r1 && (int)(int16_t)(myh.yy >> 16) >= r0);
            const bool any_more = __any_sync(0xffffffffu, more);
            if (lmask == 0u && !any_more) {
                if (lane == 0) {
                    bulk_store(gdst, tmpl, out_bytes);
                    if (P.blank) P.blank[(size_t)image * bands_per_image + band] = 1;
                }
                continue;
            }
            if (P.blank && lane == 0) P.blank[(size_t)image * bands_per_image + band] = 0;
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
            uint32_t todo = lmask;
            uint32_t lnext = lb + 32;          // next layer beyond the 32 kept in registers
            for (;;) {
                FigHdr L;
                if (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    L = fig_shfl_hdr(myh, src);
                } else {
                    if (!any_more || lnext >= le) break;
                    L = fig_load_hdr(P.layers + lnext);
                    lnext++;
                    if ((int)(int16_t)(L.yy & 0xffffu) > r1 || (int)(int16_t)(L.yy >> 16) < r0) continue;
                }
                const int Lymin = (int16_t)(L.yy & 0xffffu), Lymax = (int16_t)(L.yy >> 16);
                const int Lxmin = (int16_t)(L.xx & 0xffffu), Lxmax = (int16_t)(L.xx >> 16);
                // the edges that can reach rows [r0, r1]
                uint32_t e_lo = 0, e_hi = L.n_edges;
                if (L.n_edges >= FIG_DIR_MIN_EDGES) {
                    int lo_row = r0 - (int)(L.sd & 0xffffu);
                    if (lo_row < Lymin) lo_row = Lymin;
                    const int s_lo = (lo_row - Lymin) / FIG_SLAB_ROWS, s_hi = (r1 - Lymin) / FIG_SLAB_ROWS + 1;
                    e_lo = __ldg(P.dir + L.dir_begin + s_lo);
                    if (s_hi < (int)(L.sd >> 16)) e_hi = __ldg(P.dir + L.dir_begin + s_hi);
                }
                const FEdge *ep = P.edges + L.edge_begin;
                const int ya = Lymin > r0 ? Lymin : r0, yb = Lymax < r1 ? Lymax : r1;
                // column chunks of FIG_ACC cells (one chunk unless the layer is wider than the accumulators)
                for (int x0 = Lxmin & ~31; x0 <= Lxmax; x0 += FIG_ACC) {
                    // ---- scatter
                    {
                        FigAcc acc;
                        acc.cover = cover; acc.area = area; acc.gmask = gmask; acc.carry_in = carry_in;
                        acc.x0 = x0; acc.r0 = r0; acc.off = 0; acc.row_i = 0;
                        for (uint32_t e = e_lo + lane; e < e_hi; e += 32) {
                            const int4 v = __ldg((const int4 *)(ep + e));
                            FEdge ed;
                            ed.x1 = v.x; ed.y1 = v.y; ed.x2 = v.z; ed.y2 = v.w;
                            fig_edge_rows(ed, r0, r1, acc);
                        }
                    }
                    __syncwarp();
                    // ---- resolve
                    const bool last_chunk = x0 + FIG_ACC > Lxmax;
                    const int fill_end = last_chunk ? W : x0 + FIG_ACC;
                    // two rows at a time, one per half-warp, 16-cell groups: a band is two rows, and the cells an outline touches in a
                    // row rarely fill a 32-cell group, so the halves resolve both rows in the passes one row used to take
                    for (int y0 = ya; y0 <= yb; y0 += 2) {
                        const int y = y0 + half;
                        const bool rowok = y <= yb;
                        const int ri = rowok ? y - r0 : 0, ro = ri * FIG_ACC;
                        uint32_t gm = rowok ? gmask[ri] : 0u;
                        int carry = rowok ? carry_in[ri] : 0;
                        if (!__any_sync(0xffffffffu, gm != 0u || carry != 0)) continue;
                        __syncwarp();
                        if (hl == 0 && rowok) { gmask[ri] = 0u; carry_in[ri] = 0; }
                        uint8_t *rowpx = mine + (size_t)ri * W * 3;
                        int prev_end = x0;
                        while (__any_sync(0xffffffffu, gm != 0u)) {
                            const bool on = gm != 0u;
                            const int g = on ? __ffs(gm) - 1 : 0;
                            if (on) gm &= gm - 1;
                            const int xs = x0 + g * 16;
                            if (on && carry != 0 && xs > prev_end) {    // interior run at constant coverage
                                const uint32_t al = fig_alpha(carry, 0);
                                for (int x = prev_end + hl; x < xs; x += 16) fig_put3(rowpx + x * 3, fig_blend_px_m(fig_get3(rowpx + x * 3), L.rgba, al, magic));
                            }
                            const int x = xs + hl;
                            int c = 0, a = 0;
                            if (on) {
                                c = cover[ro + g * 16 + hl]; a = area[ro + g * 16 + hl];
                                if (c | a) { cover[ro + g * 16 + hl] = 0; area[ro + g * 16 + hl] = 0; }
                            }
                            int ci = c;
#pragma unroll
                            for (int o = 1; o < 16; o <<= 1) {
                                const int t = __shfl_up_sync(0xffffffffu, ci, o, 16);
                                if (hl >= o) ci += t;
                            }
                            const uint32_t al = fig_alpha(carry + ci, a);
                            if (on && x < W && al) fig_put3(rowpx + x * 3, fig_blend_px_m(fig_get3(rowpx + x * 3), L.rgba, al, magic));
                            const int tot = __shfl_sync(0xffffffffu, ci, 15, 16);
                            if (on) { carry += tot; prev_end = xs + 16; }
                        }
                        if (rowok && carry != 0 && prev_end < fill_end) {   // interior continues to the next chunk / the right canvas edge
                            const uint32_t al = fig_alpha(carry, 0);
                            for (int x = prev_end + hl; x < fill_end; x += 16) fig_put3(rowpx + x * 3, fig_blend_px_m(fig_get3(rowpx + x * 3), L.rgba, al, magic));
                        }
                    }
                    __syncwarp();
                }
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
