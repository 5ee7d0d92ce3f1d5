// rig_grid.cuh -- k_grid_overlay + k_grid_cells: the overview-image compositor (/root/reference/src/layout.py:17-31,49-192).
//
// Overlay pixels (background, borders, labels, the '?' cell) are copied;
// a pixel inside a panel cell is cv2.resize(INTER_AREA) of that panel, evaluated the way OpenCV's ResizeArea_ does it
// (resize.cpp: computeResizeAreaTab + ResizeArea_Invoker, float32 work type): for every contributing source row, in order,
//     buf = 0;  for every contributing source column, in order:  buf = buf + S * alpha      (separate multiply and add)
//     sum = first row ? beta * buf : sum + beta * buf
// and the result is saturate_cast<uchar>(sum) = round half to even.  The (index, weight) tables depend on the sizes only and
// are built on the host with OpenCV's formulas (doubles, then cast to float), so device and CPU use the same weights.
// Bound: HBM read of the panels (786,432 B each, read once) -- the overview itself is small.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rig {

constexpr int GRID_MAX_TAPS = 12;     // source pixels one destination pixel draws from, per axis (scale < 10)
constexpr int GRID_MAX_CELLS = 64;

struct GridTap {
    int32_t first;                    // first source index
    int32_t count;                    // taps
    float w[GRID_MAX_TAPS];           // weights, in source order
};

struct GridParams {
    const uint8_t *panels;
    const int32_t *panel_index;       // [n_samples * n_cells], device
    const uint8_t *overlay;
    const int8_t *cell_of_pixel;
    const GridTap *taps;              // [cell_size] (square panels and cells: the same table for x and y)
    uint8_t *out;
    int n_samples, n_cells, panel_w, panel_h, grid_w, grid_h, cell_size;
    int cell_x[GRID_MAX_CELLS], cell_y[GRID_MAX_CELLS];
};

__device__ __forceinline__ uint32_t grid_round_u8(float v)
{
    int r = __float2int_rn(v);        // round half to even, as cvRound
    return (uint32_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

// overlay pixels of every sample: a straight 16-byte copy of the constant canvas (cell pixels are overwritten by k_grid_cells)
__global__ void __launch_bounds__(256) k_grid_overlay(const GridParams P)
{
    const size_t bytes = (size_t)P.grid_w * P.grid_h * 3;
    const int sample = blockIdx.y;
    uint8_t *o = P.out + (size_t)sample * bytes;
    const size_t vec = (((uintptr_t)o | (uintptr_t)P.overlay) & 15u) == 0 ? bytes / 16 : 0;
    for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < vec; i += (size_t)gridDim.x * blockDim.x)
        ((uint4 *)o)[i] = __ldg((const uint4 *)P.overlay + i);
    for (size_t i = vec * 16 + blockIdx.x * blockDim.x + threadIdx.x; i < bytes; i += (size_t)gridDim.x * blockDim.x) o[i] = P.overlay[i];
}

// One CTA per (sample, cell, destination row): the source rows that row draws from are staged in shared memory with
// 16-byte coalesced loads (each panel byte leaves HBM once, plus the row shared by two neighbouring destination rows),
// then thread x evaluates destination pixel x in OpenCV's order.
__global__ void __launch_bounds__(128) k_grid_cells(const GridParams P)
{
    extern __shared__ __align__(16) uint8_t rows[];
    const int sample = blockIdx.y;
    const int cell = blockIdx.x / P.cell_size, yc = blockIdx.x - cell * P.cell_size;
    const int pi = P.panel_index[sample * P.n_cells + cell];
    const int gy = P.cell_y[cell] + yc, gx0 = P.cell_x[cell];
    const size_t npix = (size_t)P.grid_w * P.grid_h;
    uint8_t *orow = P.out + ((size_t)sample * npix + (size_t)gy * P.grid_w + gx0) * 3;
    const int8_t *cmap = P.cell_of_pixel + (size_t)gy * P.grid_w + gx0;
    if (pi < 0) {                                            // blank cell: white
        for (int x = threadIdx.x; x < P.cell_size; x += blockDim.x)
            if (cmap[x] == cell) { orow[3 * x] = 255; orow[3 * x + 1] = 255; orow[3 * x + 2] = 255; }
        return;
    }
    const GridTap *ty = P.taps + yc;                          // (read in place: a by-value copy indexed by k would live in local memory)
    const int ty_first = ty->first, ty_count = ty->count;
    const int row_bytes = P.panel_w * 3;
    {
        const uint4 *src = (const uint4 *)(P.panels + ((size_t)pi * P.panel_h + ty_first) * row_bytes);
        uint4 *dst = (uint4 *)rows;
        const int n16 = ty_count * row_bytes / 16;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < P.cell_size; x += blockDim.x) {
        if (cmap[x] != cell) continue;                       // border ring / overlay pixels inside the cell square
        const GridTap *tx = P.taps + x;
        const int tx_first = tx->first, tx_count = tx->count;
        float wx[GRID_MAX_TAPS];
#pragma unroll
        for (int k = 0; k < GRID_MAX_TAPS; k++) wx[k] = __ldg(&tx->w[k]);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        const int byte0 = tx_first * 3, base16 = byte0 & ~15, sw = (byte0 & 15) >> 2;
        const uint32_t sh = 8u * (uint32_t)(byte0 & 3);
        for (int j = 0; j < ty_count; j++) {
            float b0 = 0.f, b1 = 0.f, b2 = 0.f;
            if (tx_count <= 8) {
                // the span's (at most 24) bytes with three conflict-free 16-byte loads, re-aligned in registers: byte loads at a
                // 16-byte stride between neighbouring threads would serialise four ways in shared memory
                const uint4 *rp = (const uint4 *)(rows + (size_t)j * row_bytes + base16);
                const uint4 A = rp[0], B = rp[1], C = rp[2];
                uint32_t v[6];
                switch (sw) {
                case 0:
                    v[0] = __funnelshift_r(A.x, A.y, sh); v[1] = __funnelshift_r(A.y, A.z, sh); v[2] = __funnelshift_r(A.z, A.w, sh);
                    v[3] = __funnelshift_r(A.w, B.x, sh); v[4] = __funnelshift_r(B.x, B.y, sh); v[5] = __funnelshift_r(B.y, B.z, sh);
                    break;
                case 1:
                    v[0] = __funnelshift_r(A.y, A.z, sh); v[1] = __funnelshift_r(A.z, A.w, sh); v[2] = __funnelshift_r(A.w, B.x, sh);
                    v[3] = __funnelshift_r(B.x, B.y, sh); v[4] = __funnelshift_r(B.y, B.z, sh); v[5] = __funnelshift_r(B.z, B.w, sh);
                    break;
                case 2:
                    v[0] = __funnelshift_r(A.z, A.w, sh); v[1] = __funnelshift_r(A.w, B.x, sh); v[2] = __funnelshift_r(B.x, B.y, sh);
                    v[3] = __funnelshift_r(B.y, B.z, sh); v[4] = __funnelshift_r(B.z, B.w, sh); v[5] = __funnelshift_r(B.w, C.x, sh);
                    break;
                default:
                    v[0] = __funnelshift_r(A.w, B.x, sh); v[1] = __funnelshift_r(B.x, B.y, sh); v[2] = __funnelshift_r(B.y, B.z, sh);
                    v[3] = __funnelshift_r(B.z, B.w, sh); v[4] = __funnelshift_r(B.w, C.x, sh); v[5] = __funnelshift_r(C.x, C.y, sh);
                    break;
                }
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (k >= tx_count) break;
                    const float a = wx[k];
                    b0 = __fadd_rn(b0, __fmul_rn((float)((v[(3 * k) >> 2] >> (8 * ((3 * k) & 3))) & 0xffu), a));
                    b1 = __fadd_rn(b1, __fmul_rn((float)((v[(3 * k + 1) >> 2] >> (8 * ((3 * k + 1) & 3))) & 0xffu), a));
                    b2 = __fadd_rn(b2, __fmul_rn((float)((v[(3 * k + 2) >> 2] >> (8 * ((3 * k + 2) & 3))) & 0xffu), a));
                }
            } else {
                const uint8_t *row = rows + (size_t)j * row_bytes + byte0;
#pragma unroll
                for (int k = 0; k < GRID_MAX_TAPS; k++) {
                    if (k >= tx_count) break;
                    const float a = wx[k];
                    b0 = __fadd_rn(b0, __fmul_rn((float)row[3 * k], a));
                    b1 = __fadd_rn(b1, __fmul_rn((float)row[3 * k + 1], a));
                    b2 = __fadd_rn(b2, __fmul_rn((float)row[3 * k + 2], a));
                }
            }
            const float beta = __ldg(&ty->w[j]);
            if (j == 0) { s0 = __fmul_rn(beta, b0); s1 = __fmul_rn(beta, b1); s2 = __fmul_rn(beta, b2); }
            else {
                s0 = __fadd_rn(s0, __fmul_rn(beta, b0));
                s1 = __fadd_rn(s1, __fmul_rn(beta, b1));
                s2 = __fadd_rn(s2, __fmul_rn(beta, b2));
            }
        }
        orow[3 * x] = (uint8_t)grid_round_u8(s0); orow[3 * x + 1] = (uint8_t)grid_round_u8(s1); orow[3 * x + 2] = (uint8_t)grid_round_u8(s2);
    }
}

}  // namespace rig
