// rig_grid.cuh -- k_grid: the overview-image compositor (/root/reference/src/layout.py:17-31,49-192).
//
// One thread per pixel of the overview canvas.  Overlay pixels (background, borders, labels, the '?' cell) are copied;
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

__global__ void __launch_bounds__(256) k_grid(const GridParams P)
{
    const int npix = P.grid_w * P.grid_h;
    const int sample = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        uint8_t *o = P.out + ((size_t)sample * npix + i) * 3;
        const int cell = P.cell_of_pixel[i];
        if (cell < 0) {
            o[0] = P.overlay[3 * i]; o[1] = P.overlay[3 * i + 1]; o[2] = P.overlay[3 * i + 2];
            continue;
        }
        const int pi = P.panel_index[sample * P.n_cells + cell];
        if (pi < 0) { o[0] = o[1] = o[2] = 255; continue; }
        const int y = i / P.grid_w, x = i - y * P.grid_w;
        const GridTap ty = P.taps[y - P.cell_y[cell]], tx = P.taps[x - P.cell_x[cell]];
        const uint8_t *S = P.panels + (size_t)pi * P.panel_h * P.panel_w * 3;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        for (int j = 0; j < ty.count; j++) {
            const uint8_t *row = S + ((size_t)(ty.first + j) * P.panel_w + tx.first) * 3;
            float b0 = 0.f, b1 = 0.f, b2 = 0.f;
            for (int k = 0; k < tx.count; k++) {
                const float a = tx.w[k];
                b0 = __fadd_rn(b0, __fmul_rn((float)row[3 * k], a));
                b1 = __fadd_rn(b1, __fmul_rn((float)row[3 * k + 1], a));
                b2 = __fadd_rn(b2, __fmul_rn((float)row[3 * k + 2], a));
            }
            const float beta = ty.w[j];
            if (j == 0) { s0 = __fmul_rn(beta, b0); s1 = __fmul_rn(beta, b1); s2 = __fmul_rn(beta, b2); }
            else {
                s0 = __fadd_rn(s0, __fmul_rn(beta, b0));
                s1 = __fadd_rn(s1, __fmul_rn(beta, b1));
                s2 = __fadd_rn(s2, __fmul_rn(beta, b2));
            }
        }
        o[0] = (uint8_t)grid_round_u8(s0); o[1] = (uint8_t)grid_round_u8(s1); o[2] = (uint8_t)grid_round_u8(s2);
    }
}

}  // namespace rig
