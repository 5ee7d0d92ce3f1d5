// rig_kernels.cuh -- the two sm_100a kernels of the render path.
//
//   k_expand  one warp per element: float64 vertex prologue (rotate / translate / wrap copies fused
//             in), then one lane per polygon edge / stroke segment writes 32-byte draw records.
//   k_raster  persistent warps; each warp claims bands of RIG_BAND_ROWS (5) canvas rows, keeps the
//             band in shared memory (3*W*5 bytes, packed RGB; 27 warps per CTA at W = 512), replays
//             the panel's records in draw order with lanes spread over the iterations that fall
//             inside the band, and hands the finished band to the TMA engine (cp.async.bulk shared
//             -> global, one contiguous 7.5 KB write at W = 512).  Bands no primitive touches are
//             never built: the TMA engine writes them straight from a constant template held once
//             per CTA, so background pixels cost one instruction per 7.5 KB.
//
// The roofline that bounds k_raster is HBM write bandwidth: 3*W*H bytes per panel leave the SM
// exactly once, as full 128-byte lines, and nothing is read back.
#pragma once
#include <cuda_runtime.h>

#include "rig_expand.h"
#include "rig_raster.h"
#include "rig_tma.cuh"

namespace rig {

#ifndef RIG_BAND_ROWS
#define RIG_BAND_ROWS 5
#endif
constexpr int BAND_ROWS = RIG_BAND_ROWS;
// warps per CTA: as many band buffers (+2 templates) as fit 227 KB at W = 512, and registers to match
constexpr int RASTER_MAX_WARPS = BAND_ROWS >= 8 ? 16 : BAND_ROWS == 7 ? 19 : BAND_ROWS == 6 ? 22 : BAND_ROWS == 5 ? 27 : 32;
#ifndef RIG_EXPAND_WARPS
#define RIG_EXPAND_WARPS 1
#endif
constexpr int EXPAND_WARPS = RIG_EXPAND_WARPS;   // 1: a CTA never waits for its slowest element

enum Counter { CTR_REC = 0, CTR_AUX = 1, CTR_OVERFLOW = 2, CTR_WORK = 3, CTR_COUNT = 8 };

struct WarpSync {
    __device__ __forceinline__ void operator()() const { __syncwarp(); }
};

#ifndef RIG_EXPAND_MIN_CTAS
#define RIG_EXPAND_MIN_CTAS 20
#endif
__global__ void __launch_bounds__(EXPAND_WARPS * 32, RIG_EXPAND_MIN_CTAS)
k_expand(const DElem *__restrict__ elems, const uint32_t *__restrict__ eorder, uint32_t e0, uint32_t e1,
         const ShapeTables *__restrict__ T, int W, int H,
         ElemHdr *__restrict__ hdrs, Rec *__restrict__ recs, uint32_t rec_cap, Q16 *__restrict__ aux, uint32_t aux_cap,
         uint32_t *__restrict__ counters)
{
    __shared__ ExpandScratch scratch[EXPAND_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ei = e0 + blockIdx.x * EXPAND_WARPS + warp;
    if (ei >= e1) return;
    const uint32_t e = eorder[ei];      // costliest elements of the chunk first
    const DElem el = elems[e];
    ExpandScratch &S = scratch[warp];
    ElemPlan plan = expand_plan<32>(el, *T, W, H, lane, S, WarpSync());
    uint32_t r0 = 0, a0 = 0;
    const uint32_t nsum = summary_slots(plan.nrec);
    if (lane == 0) {
        r0 = atomicAdd(&counters[CTR_REC], plan.nrec + nsum);
        a0 = atomicAdd(&counters[CTR_AUX], plan.naux);
    }
    r0 = __shfl_sync(0xffffffffu, r0, 0);
    a0 = __shfl_sync(0xffffffffu, a0, 0);
    const bool fits = (uint64_t)r0 + plan.nrec + nsum <= rec_cap && (uint64_t)a0 + plan.naux <= aux_cap;
    if (lane == 0) {
        ElemHdr h;
        h.rec_begin = r0;
        h.nrec = fits ? plan.nrec : 0;
        h.ymin = (int16_t)((fits && plan.ymin <= plan.ymax) ? plan.ymin : 32767);   // empty: a range no band overlaps
        h.ymax = (int16_t)((fits && plan.ymin <= plan.ymax) ? plan.ymax : -32768);  // (1 > 0 would overlap the band of rows 0..4)
        h.nsum = nsum;
        hdrs[e] = h;
        if (!fits) atomicExch(&counters[CTR_OVERFLOW], 1u);
    }
    if (!fits) return;
    expand_emit<32>(el, plan, *T, W, H, lane, S, recs, r0 + nsum, aux, a0, WarpSync());
    __syncwarp();
    write_summaries<32>(recs, r0, r0 + nsum, plan.nrec, lane);
}

// folds one chunk's counters into the per-render totals: [0] overflow flag, [1] records, [2] aux units
__global__ void k_collect_overflow(const uint32_t *__restrict__ counters, uint32_t *__restrict__ totals)
{
    if (threadIdx.x == 0) {
        totals[0] |= counters[CTR_OVERFLOW];
        totals[1] += counters[CTR_REC];
        totals[2] += counters[CTR_AUX];
    }
}

struct RasterOps {
    Rec mine;            // the slot this lane loaded in the current 32-slot window
    int j;               // lane that holds the record being executed
    const Rec *base;     // global address of slot 0 of the window (fallback when a record leaves the window)
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ void append(RasterScratch &S, int32_t v) const
    {
        int i = atomicAdd(&S.count, 1);
        if (i < 64) S.cross[i] = v;
    }
    __device__ __forceinline__ int bcast(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ int reduce_add(int v) const { return __reduce_add_sync(0xffffffffu, v); }
    __device__ __forceinline__ int reduce_min(int v) const { return __reduce_min_sync(0xffffffffu, v); }
    __device__ __forceinline__ int reduce_max(int v) const { return __reduce_max_sync(0xffffffffu, v); }
    __device__ __forceinline__ Rec slot_part(int k, int lo, int hi) const
    {
        Rec r;
        const int src = j + k;
        if (src < 32) {
#pragma unroll
            for (int i = 0; i < 8; i++) r.w[i] = (i >= lo && i <= hi) ? __shfl_sync(0xffffffffu, mine.w[i], src) : 0u;
        } else {
            r = slot(k);
        }
        return r;
    }
    __device__ __forceinline__ Rec slot(int k) const
    {
        Rec r;
        const int src = j + k;
        if (src < 32) {
#pragma unroll
            for (int i = 0; i < 8; i++) r.w[i] = __shfl_sync(0xffffffffu, mine.w[i], src);
        } else {
            const uint4 *rp = (const uint4 *)(base + src);
            uint4 lo = __ldg(rp), hi = __ldg(rp + 1);
            r.w[0] = lo.x; r.w[1] = lo.y; r.w[2] = lo.z; r.w[3] = lo.w;
            r.w[4] = hi.x; r.w[5] = hi.y; r.w[6] = hi.z; r.w[7] = hi.w;
        }
        return r;
    }
};

struct RasterParams {
    const uint32_t *panel_begin;   // [n_panels + 1] element offsets (absolute)
    const uint8_t *panel_flags;    // [n_panels]
    const uint32_t *order;         // [n_panels] panel ids, heaviest first inside each chunk (work is claimed in this order)
    const ElemHdr *hdrs;           // indexed by absolute element id
    const Rec *recs;
    const Q16 *aux;
    const ShapeTables *tables;
    uint8_t *out;                  // image of panel p0 starts at out
    uint32_t *counters;
    uint32_t p0, n_panels;         // panels [p0, p0 + n_panels) of the resident scene
    int W, H, grid_size;
    int n_grid;                    // grid lines per direction (grid_size - 1, at most 15)
    int16_t grid_x[16], grid_y[16]; // int(round(i * W / grid_size)) etc. (generator.py:117,121), computed on the host
    uint32_t bg;                   // c0 | c1<<8 | c2<<16
    uint32_t *sums;                // HASH variant: [n_panels_total][256] block-sum deltas vs the all-template panel (else unused)
    uint32_t coef;                 // HASH variant: luma weights in output channel order (dp4a operand)
    uint32_t skip_mask;            // diagnostics (RIG_SKIP_MASK): bit t set = records of type t are not executed
    uint32_t max_claim, heavy_bands; // scheduling: largest claim; the first heavy_bands bands (costliest panels) are claimed singly
};

// luma sum of 4 consecutive packed pixels (3 words): Y = (77 R + 150 G + 29 B) >> 8 per pixel
__device__ __forceinline__ uint32_t luma_quad(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t coef)
{
    return (__dp4a(w0 & 0x00ffffffu, coef, 0u) >> 8) + (__dp4a(__byte_perm(w0, w1, 0x4543) & 0x00ffffffu, coef, 0u) >> 8) +
           (__dp4a(__byte_perm(w1, w2, 0x4432) & 0x00ffffffu, coef, 0u) >> 8) + (__dp4a(w2 >> 8, coef, 0u) >> 8);
}
__device__ __forceinline__ uint32_t luma_group16(const uint8_t *p, uint32_t coef)   // 16 pixels = 48 bytes, 16-byte aligned
{
    const uint4 a = ((const uint4 *)p)[0], b = ((const uint4 *)p)[1], c = ((const uint4 *)p)[2];
    return luma_quad(a.x, a.y, a.z, coef) + luma_quad(a.w, b.x, b.y, coef) + luma_quad(b.z, b.w, c.x, coef) + luma_quad(c.y, c.z, c.w, coef);
}

// dynamic shared memory layout: [template plain][template grid][warp bands ...][RasterScratch per warp][filter 64][template lumas]
// HASH: the perceptual-hash block sums are accumulated from the finished band while it is still in shared memory
// (as deltas against the all-template panel), so hashing never re-reads the 786 KB image from HBM.
template <bool HASH>
__global__ void __launch_bounds__(RASTER_MAX_WARPS * 32, 1) k_raster(const RasterParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t band_bytes = (uint32_t)(3 * P.W * BAND_ROWS);
    uint8_t *tmpl0 = smem, *tmpl1 = smem + band_bytes;
    uint8_t *mine = smem + (size_t)(2 + warp) * band_bytes;
    RasterScratch *scratch = (RasterScratch *)(smem + (size_t)(2 + nwarps) * band_bytes);
    uint8_t *filter = (uint8_t *)(scratch + nwarps);
    uint32_t *tluma = (uint32_t *)(filter + 64);      // HASH: luma of each 16-pixel group of one template row, [2][W/16]
    RasterScratch &S = scratch[warp];

    // ---- CTA prologue: constant templates + filter table
    {
        const uint32_t c0 = P.bg & 0xff, c1 = (P.bg >> 8) & 0xff, c2 = (P.bg >> 16) & 0xff;
        // 12 bytes = 4 pixels; three 32-bit words repeat
        const uint32_t w0 = c0 | (c1 << 8) | (c2 << 16) | (c0 << 24);
        const uint32_t w1 = c1 | (c2 << 8) | (c0 << 16) | (c1 << 24);
        const uint32_t w2 = c2 | (c0 << 8) | (c1 << 16) | (c2 << 24);
        uint32_t *t0 = (uint32_t *)tmpl0, *t1 = (uint32_t *)tmpl1;
        for (uint32_t i = threadIdx.x; i < band_bytes / 4; i += blockDim.x) {
            uint32_t m = i % 3, v = m == 0 ? w0 : (m == 1 ? w1 : w2);
            t0[i] = v;
            t1[i] = v;
        }
        if (threadIdx.x < 64) filter[threadIdx.x] = P.tables->filter[threadIdx.x];
        __syncthreads();
        for (int g = 0; g < P.n_grid; g++) {
            const int x = P.grid_x[g];
            if ((unsigned)x < (unsigned)P.W && threadIdx.x < BAND_ROWS) put3(tmpl1 + (threadIdx.x * P.W + x) * 3, 0u);
        }
        fence_async_smem();
        __syncthreads();
        if (HASH) {
            const int groups = P.W / 16;
            for (int g = threadIdx.x; g < 2 * groups; g += blockDim.x)
                tluma[g] = luma_group16((g < groups ? tmpl0 : tmpl1) + (g % groups) * 48, P.coef);
            __syncthreads();
        }
    }
    // rows of the horizontal grid lines
    const int bands_per_panel = (P.H + BAND_ROWS - 1) / BAND_ROWS;   // the last band of a panel may be shorter
    const uint32_t total_bands = P.n_panels * (uint32_t)bands_per_panel;
    bool dirty = false;   // my band buffer has an outstanding bulk read
    uint32_t cur_slot = 0xffffffffu, panel = 0, eb = 0, ee = 0;
    bool use_grid = false;
    ElemHdr myh;
    myh.rec_begin = 0; myh.nrec = 0; myh.ymin = 32767; myh.ymax = -32768; myh.nsum = 0;
    RasterOps ops;
    ops.mine = rec_nop(); ops.j = 0; ops.base = P.recs;

    for (;;) {
        // guided self-scheduling: large claims while much work is left, single bands at the end
        uint32_t first = 0, take = 1;
        if (lane == 0) {
            const uint32_t cur = *(volatile uint32_t *)&P.counters[CTR_WORK];
            const uint32_t left = cur < total_bands ? total_bands - cur : 0;
            take = left / (4u * gridDim.x * (uint32_t)nwarps);
            take = take < 1u ? 1u : (take > P.max_claim ? P.max_claim : take);
            if (cur < P.heavy_bands) take = 1u;      // panels are sorted heaviest first: spread those band by band
            first = atomicAdd(&P.counters[CTR_WORK], take);
        }
        first = __shfl_sync(0xffffffffu, first, 0);
        take = __shfl_sync(0xffffffffu, take, 0);
        if (first >= total_bands) break;
        for (uint32_t bi = first; bi < first + take && bi < total_bands; bi++) {
            // panel metadata is fetched once per panel, not once per band: consecutive claims mostly stay in one panel
            const uint32_t slot = bi / bands_per_panel;
            if (slot != cur_slot) {
                cur_slot = slot;
                panel = P.order[P.p0 + slot];
                eb = P.panel_begin[panel];
                ee = P.panel_begin[panel + 1];
                use_grid = P.panel_flags[panel] & 1;
                myh.rec_begin = 0; myh.nrec = 0; myh.ymin = 32767; myh.ymax = -32768; myh.nsum = 0;
                if (eb + lane < ee) myh = P.hdrs[eb + lane];     // lane e keeps the header of the panel's e-th element
            }
            const uint32_t pl = panel - P.p0;
            const int band = (int)(bi - slot * bands_per_panel);
            Band B;
            B.r0 = band * BAND_ROWS; B.r1 = B.r0 + BAND_ROWS - 1; B.W = P.W; B.H = P.H; B.buf = mine;
            if (B.r1 > P.H - 1) B.r1 = P.H - 1;
            const uint32_t out_bytes = (uint32_t)(3 * P.W * (B.r1 - B.r0 + 1));
            uint8_t *gdst = P.out + ((size_t)pl * P.H + B.r0) * (size_t)P.W * 3;
            // ---- is anything drawn into this band?
            bool mine_hits = myh.ymin <= B.r1 && myh.ymax >= B.r0;
            bool more = false;                                     // panels with more than 32 elements: the rest from memory
            for (uint32_t e = eb + 32 + lane; e < ee; e += 32) {
                const ElemHdr h = P.hdrs[e];
                more |= h.ymin <= B.r1 && h.ymax >= B.r0;
            }
            bool gline = false;
            if (use_grid && lane < P.n_grid) gline = P.grid_y[lane] >= B.r0 && P.grid_y[lane] <= B.r1;
            const uint32_t emask = __ballot_sync(0xffffffffu, mine_hits);
            const bool any = emask != 0u || __any_sync(0xffffffffu, more || gline);
            if (!any) {
                if (lane == 0) bulk_store(gdst, use_grid ? tmpl1 : tmpl0, out_bytes);
                continue;
            }
            // ---- build the band in shared memory
            if (dirty) {
                if (lane == 0) bulk_wait_read();
                __syncwarp();
            }
            {
                const uint4 *src = (const uint4 *)tmpl0;
                uint4 *dst = (uint4 *)mine;
                for (uint32_t i = lane; i < band_bytes / 16; i += 32) dst[i] = src[i];
            }
            __syncwarp();
            for (uint32_t e = eb; e < ee; e++) {
                ElemHdr h;
                if (e - eb < 32u) {
                    if (!((emask >> (e - eb)) & 1u)) continue;
                    const int src = (int)(e - eb);
                    h.rec_begin = __shfl_sync(0xffffffffu, myh.rec_begin, src);
                    h.nrec = __shfl_sync(0xffffffffu, myh.nrec, src);
                    h.nsum = __shfl_sync(0xffffffffu, myh.nsum, src);
                    h.ymin = 0; h.ymax = 0;
                } else {
                    h = P.hdrs[e];
                    if (h.ymin > B.r1 || h.ymax < B.r0) continue;
                }
                const uint32_t *sums = (const uint32_t *)(P.recs + h.rec_begin);
                const Rec *payload = P.recs + h.rec_begin + h.nsum;
                const uint32_t nwin = (h.nrec + 31u) / 32u;
                for (uint32_t wb = 0; wb < nwin; wb += 32) {
                    // one 4-byte row range per 32-slot window: skip the windows that hold nothing for this band
                    bool wh = false;
                    if (wb + lane < nwin) {
                        const uint32_t sv = __ldg(sums + wb + lane);
                        wh = (int16_t)(sv & 0xffff) <= B.r1 && (int16_t)(sv >> 16) >= B.r0;
                    }
                    uint32_t wm = __ballot_sync(0xffffffffu, wh);
                    while (wm) {
                        const uint32_t base = (wb + (uint32_t)(__ffs(wm) - 1)) * 32u;
                        wm &= wm - 1;
                        // every lane pulls one whole 32-byte slot (1 KB per warp, coalesced); records are
                        // then handed round by shuffles, so a record costs no further global latency
                        bool hit = false;
                        ops.base = payload + base;
                        if (base + lane < h.nrec) {
                            const uint4 *rp = (const uint4 *)(ops.base + lane);
                            const uint4 lo = __ldg(rp), hi = __ldg(rp + 1);
                            ops.mine.w[0] = lo.x; ops.mine.w[1] = lo.y; ops.mine.w[2] = lo.z; ops.mine.w[3] = lo.w;
                            ops.mine.w[4] = hi.x; ops.mine.w[5] = hi.y; ops.mine.w[6] = hi.z; ops.mine.w[7] = hi.w;
                            hit = rec_ymin(ops.mine) <= B.r1 && rec_ymax(ops.mine) >= B.r0;
                        } else {
                            ops.mine = rec_nop();
                        }
                        uint32_t m = __ballot_sync(0xffffffffu, hit);
                        while (m) {
                            ops.j = __ffs(m) - 1;
                            m &= m - 1;
                            const Rec r = ops.slot(0);
                            if ((P.skip_mask >> rec_type(r)) & 1u) continue;
                            exec_record<32>(r, P.aux, filter, B, lane, S, ops);
                            __syncwarp();
                        }
                    }
                }
            }
            if (use_grid) draw_grid<32>(B, P.n_grid, P.grid_x, P.grid_y, lane);
            if (HASH) {
                __syncwarp();
                const int groups = P.W / 16, bh = P.H / 16, lpb = P.W / 256;      // lanes (16-px groups) per hash block
                const uint32_t *tl = tluma + (use_grid ? groups : 0);
                uint32_t *psum = P.sums + (size_t)panel * 256u;
                for (int g0 = 0; g0 < groups; g0 += 32) {
                    const int g = g0 + lane;
                    int cur_by = B.r0 / bh;
                    uint32_t acc = 0;
                    for (int y = B.r0; y <= B.r1 + 1; y++) {
                        const int by = y <= B.r1 ? y / bh : -1;
                        if (by != cur_by) {                                      // flush the block row: segmented sum, one RED per block
                            uint32_t s = acc;
                            for (int o = 1; o < lpb && o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                            const int grp = lpb < 32 ? lpb : 32;
                            if ((lane & (grp - 1)) == 0 && g < groups && s != 0u) atomicAdd(psum + cur_by * 16 + (g * 16) / (P.W / 16), s);
                            acc = 0; cur_by = by;
                        }
                        if (y <= B.r1 && g < groups) acc += luma_group16(mine + ((y - B.r0) * P.W + g * 16) * 3, P.coef) - tl[g];
                    }
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

inline size_t raster_smem_bytes(int W, int nwarps)
{
    return (size_t)(2 + nwarps) * 3 * W * BAND_ROWS + (size_t)nwarps * sizeof(RasterScratch) + 64 + 2 * (size_t)(W / 16) * 4;
}

}  // namespace rig
