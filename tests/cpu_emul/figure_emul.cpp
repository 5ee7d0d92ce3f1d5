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
    std::vector<int> &cover, &area, &carry_in;
    int acc_cells, x0, r0, off = 0, row_i = 0;
    HostAcc(std::vector<int> &c, std::vector<int> &a, std::vector<int> &ci, int acc_cells_, int x0_, int r0_)
        : cover(c), area(a), carry_in(ci), acc_cells(acc_cells_), x0(x0_), r0(r0_) {}
    void row(int ey) { row_i = ey - r0; off = row_i * acc_cells; }
    void operator()(int ex, int dc, int da)
    {
        if ((dc | da) == 0) return;
        const int cx = ex - x0;
        if (cx < 0) { carry_in[row_i] += dc; return; }
        if (cx >= acc_cells) return;
        cover[off + cx] += dc;
        area[off + cx] += da;
    }
};
void blend(uint8_t *px, uint32_t rgba, uint32_t al)
{
    const uint32_t v = fig_blend_px((uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16), rgba, al);
    px[0] = (uint8_t)v; px[1] = (uint8_t)(v >> 8); px[2] = (uint8_t)(v >> 16);
}
}  // namespace

// acc_cells: accumulator columns per row (multiple of 32); layers wider than that are resolved in column chunks with a
// carry-in, as k_figure does with FIG_ACC_CELLS
static void emul_render(int W, int H, int n_layers, const uint32_t *layers, const int32_t *edges, int band_rows, int acc_cells, int group,
                        uint8_t *rgb);

extern "C" void figure_emul_render(int W, int H, int n_layers, const uint32_t *layers, const int32_t *edges, int band_rows, int acc_cells,
                                   uint8_t *rgb)
{
    emul_render(W, H, n_layers, layers, edges, band_rows, acc_cells, 0, rgb);
}

// The resolve step organised as k_figure organises it: only the `group`-cell groups (16 in the kernel) that the scatter touched are
// visited, left to right; between them the running cover is constant, so a non-zero cover there is an interior run filled at
// constant coverage, and so is what follows the last group up to the chunk's end / the canvas edge.  (The kernel resolves the band's
// two rows at once, one per half-warp; rows are independent, so a serial emulation visits them in turn.)
extern "C" void figure_emul_render_grouped(int W, int H, int n_layers, const uint32_t *layers, const int32_t *edges, int band_rows,
                                           int acc_cells, int group, uint8_t *rgb)
{
    emul_render(W, H, n_layers, layers, edges, band_rows, acc_cells, group, rgb);
}

static void emul_render(int W, int H, int n_layers, const uint32_t *layers, const int32_t *edges, int band_rows, int acc_cells, int group,
                        uint8_t *rgb)
{
    memset(rgb, 255, (size_t)W * H * 3);
    std::vector<int> cover((size_t)band_rows * acc_cells), area((size_t)band_rows * acc_cells), carry_in(band_rows);
    for (int r0 = 0; r0 < H; r0 += band_rows) {
        const int r1 = r0 + band_rows - 1 < H - 1 ? r0 + band_rows - 1 : H - 1;
        for (int l = 0; l < n_layers; l++) {
            const uint32_t *L = layers + 8 * l;
            const int xmin = (int32_t)L[4], xmax = (int32_t)L[5], ymin = (int32_t)L[6], ymax = (int32_t)L[7];
            if (ymin > r1 || ymax < r0) continue;
            for (int x0 = xmin & ~31; x0 <= xmax; x0 += acc_cells) {
                std::fill(cover.begin(), cover.end(), 0);
                std::fill(area.begin(), area.end(), 0);
                std::fill(carry_in.begin(), carry_in.end(), 0);
                HostAcc acc(cover, area, carry_in, acc_cells, x0, r0);
                for (uint32_t e = 0; e < L[2]; e++) {
                    const int32_t *p = edges + 4 * (size_t)(L[1] + e);
                    FEdge ed{p[0], p[1], p[2], p[3]};
                    fig_edge_rows(ed, r0, r1, acc);
                }
                const int fill_end = x0 + acc_cells > xmax ? W : x0 + acc_cells;
                for (int y = (ymin > r0 ? ymin : r0); group > 0 && y <= (ymax < r1 ? ymax : r1); y++) {
                    const size_t ro = (size_t)(y - r0) * acc_cells;
                    int carry = carry_in[y - r0], prev_end = x0;
                    for (int g = 0; g < acc_cells / group; g++) {
                        bool touched = false;                   // (the kernel's scatter sets a bit per group it adds to)
                        for (int k = 0; k < group; k++) touched |= (cover[ro + g * group + k] | area[ro + g * group + k]) != 0;
                        if (!touched) continue;
                        const int xs = x0 + g * group;
                        if (carry != 0)
                            for (int x = prev_end; x < xs; x++) blend(rgb + ((size_t)y * W + x) * 3, L[3], fig_alpha(carry, 0));
                        int ci = 0;
                        for (int k = 0; k < group; k++) {
                            ci += cover[ro + g * group + k];
                            const uint32_t al = fig_alpha(carry + ci, area[ro + g * group + k]);
                            if (al && xs + k < W) blend(rgb + ((size_t)y * W + xs + k) * 3, L[3], al);
                        }
                        carry += ci;
                        prev_end = xs + group;
                    }
                    if (carry != 0)
                        for (int x = prev_end; x < fill_end; x++) blend(rgb + ((size_t)y * W + x) * 3, L[3], fig_alpha(carry, 0));
                }
                for (int y = (ymin > r0 ? ymin : r0); group == 0 && y <= (ymax < r1 ? ymax : r1); y++) {
                    int run = carry_in[y - r0];
                    int last = x0;          // first pixel not yet decided
                    for (int cx = 0; cx < acc_cells; cx++) {
                        const int c = cover[(size_t)(y - r0) * acc_cells + cx], a = area[(size_t)(y - r0) * acc_cells + cx];
                        run += c;
                        const uint32_t al = fig_alpha(run, a);
                        const int x = x0 + cx;
                        if (al && x < W) blend(rgb + ((size_t)y * W + x) * 3, L[3], al);
                        last = x + 1;
                    }
                    if (run != 0)
                        for (int x = last; x < fill_end; x++) blend(rgb + ((size_t)y * W + x) * 3, L[3], fig_alpha(run, 0));
                }
            }
        }
    }
}

// exhaustive check of the division-free blend used by the kernel against the plain one: every destination value, colour
// value and alpha; returns the number of mismatches
extern "C" long figure_emul_check_magic()
{
    uint32_t magic[256];
    for (uint32_t a = 0; a < 256; a++) magic[a] = fig_div_magic(a);
    long bad = 0;
    for (uint32_t alpha = 1; alpha < 256; alpha++)
        for (uint32_t p = 0; p < 256; p++)
            for (uint32_t c = 0; c < 256; c++)
                bad += fig_blend1(p, c, alpha) != fig_blend1_m(p, c, alpha, magic[alpha]);
    for (uint32_t ca = 0; ca < 256; ca += 5)
        for (uint32_t cover = 1; cover < 256; cover += 3) {
            const uint32_t px = (ca * 2654435761u + cover * 40503u) & 0xffffffu, rgba = ((cover * 97u + ca * 13u) & 0xffffffu) | (ca << 24);
            bad += fig_blend_px(px, rgba, cover) != fig_blend_px_m(px, rgba, cover, magic);
        }
    return bad;
}
