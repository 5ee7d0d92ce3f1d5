// tests/cpu_emul/figure_emul.cpp -- TEST INFRASTRUCTURE.  Compiles the per-lane arithmetic of the figure kernel
// (csrc/rig_figure_core.h: Agg's row / cell split in closed form, coverage, blender) for the CPU and drives it the way
// k_figure does -- band by band, layer by layer, scatter then a left-to-right resolve -- so the arithmetic and the band
// clipping can be checked against oracle/aggdraw.c without a GPU.  band_rows is a run-time argument here so the tests can
// start the closed-form DDA at every possible row offset.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../reasoning_image_generation_b200/csrc/rig_figure_core.h"

using namespace rig;

namespace {
struct HostAcc {
    std::vector<int> &cover, &area;
    int W, r0, off = 0;
    HostAcc(std::vector<int> &c, std::vector<int> &a, int W_, int r0_) : cover(c), area(a), W(W_), r0(r0_) {}
    void row(int ey) { off = (ey - r0) * W; }
    void operator()(int ex, int dc, int da)
    {
        if ((unsigned)ex >= (unsigned)W) return;
        cover[off + ex] += dc;
        area[off + ex] += da;
    }
};
}  // namespace

extern "C" void figure_emul_render(int W, int H, int n_layers, const uint32_t *layers, const int32_t *edges, int band_rows, uint8_t *rgb)
{
    memset(rgb, 255, (size_t)W * H * 3);
    std::vector<int> cover((size_t)band_rows * W), area((size_t)band_rows * W);
    for (int r0 = 0; r0 < H; r0 += band_rows) {
        const int r1 = r0 + band_rows - 1 < H - 1 ? r0 + band_rows - 1 : H - 1;
        for (int l = 0; l < n_layers; l++) {
            const uint32_t *L = layers + 8 * l;
            const int ymin = (int32_t)L[6], ymax = (int32_t)L[7];
            if (ymin > r1 || ymax < r0) continue;
            std::fill(cover.begin(), cover.end(), 0);
            std::fill(area.begin(), area.end(), 0);
            HostAcc acc(cover, area, W, r0);
            for (uint32_t e = 0; e < L[2]; e++) {
                const int32_t *p = edges + 4 * (size_t)(L[1] + e);
                FEdge ed{p[0], p[1], p[2], p[3]};
                fig_edge_rows(ed, r0, r1, acc);
            }
            for (int y = r0; y <= r1; y++) {
                int run = 0;
                for (int x = 0; x < W; x++) {
                    run += cover[(size_t)(y - r0) * W + x];
                    const uint32_t al = fig_alpha(run, area[(size_t)(y - r0) * W + x]);
                    if (!al) continue;
                    uint8_t *px = rgb + ((size_t)y * W + x) * 3;
                    const uint32_t v = fig_blend_px((uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16), L[3], al);
                    px[0] = (uint8_t)v; px[1] = (uint8_t)(v >> 8); px[2] = (uint8_t)(v >> 16);
                }
            }
        }
    }
}
