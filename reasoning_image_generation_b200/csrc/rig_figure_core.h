// rig_figure_core.h -- per-lane arithmetic of the figure (matplotlib/Agg) rasteriser, host + device.
//
// The reference's multigraph path rasterises with Anti-Grain Geometry through matplotlib's Agg backend
// (multigraph_generation/generator.py:250-259 -> fig.savefig).  Agg's scanline rasteriser is integer work on 24.8
// fixed-point edges: every edge adds a signed (cover, area) pair to each pixel cell it crosses
// (agg_rasterizer_cells_aa.h: line() / render_hline()), a scanline sweep turns the running cover and the cell's area
// into an 8-bit coverage, and matplotlib's fixed_blender_rgba_plain blends the colour with it.  The cell sums are
// commutative, so the GPU accumulates them with shared-memory atomics in any order; what must be reproduced exactly
// is each edge's split into rows and cells.  Agg walks an edge row by row with a DDA; here the state of that DDA at
// the first row of a band is computed in closed form (floor((p + j q) / dy) decomposed so that everything stays in
// 32 bits for canvases up to 3072 x 3072), and from there the walk is Agg's own.
//
// The same code is compiled for the CPU by tests/cpu_emul/figure_emul.cpp and checked against oracle/aggdraw.c.
#pragma once
#include <stdint.h>

#include "rig_core.h"

namespace rig {

// one polygon edge after stroking + clipping, 24.8 fixed point, y down (the arguments of agg's cells.line())
struct alignas(16) FEdge {
    int32_t x1, y1, x2, y2;
};

// One rasteriser pass (the face or the stroke of one artist): a closed set of edges and a colour.
// Edges are sorted by their first row; dir[] (ndir entries, 16-row slabs from ymin) holds, per slab, the index of the
// first edge that starts in or after it, so a band only looks at edges that can reach it.
struct alignas(16) FLayer {
    uint32_t edge_begin, n_edges;
    uint32_t rgba;                   // r | g << 8 | b << 16 | a << 24, non-premultiplied, in output channel order
    int16_t ymin, ymax, xmin, xmax;  // pixel cells the edges touch (inclusive, clipped to the canvas; empty: ymin > ymax)
    uint32_t dir_begin;
    uint16_t maxspan, ndir;          // most rows any one edge spans - 1; entries of dir[]
    uint32_t pad;
};
constexpr int FIG_SLAB_ROWS = 16;
constexpr uint32_t FIG_DIR_MIN_EDGES = 96;   // k_figure scans layers with fewer edges whole (no slab directory look-up)

// ---- render_hline: one edge inside one row ------------------------------------------------------
// acc(ex, dcover, darea) is called for every cell with a non-zero contribution.
template <class Acc>
RIG_HD void fig_hline(int x1, int y1, int x2, int y2, Acc &acc)
{
    int ex1 = x1 >> 8;
    const int ex2 = x2 >> 8, fx1 = x1 & 255, fx2 = x2 & 255;
    if (y1 == y2) return;
    if (ex1 == ex2) {
        const int d = y2 - y1;
        acc(ex1, d, (fx1 + fx2) * d);
        return;
    }
    int p = (256 - fx1) * (y2 - y1), first = 256, incr = 1, dx = x2 - x1;
    if (dx < 0) {
        p = fx1 * (y2 - y1);
        first = 0;
        incr = -1;
        dx = -dx;
    }
    int delta = p / dx, mod = p % dx;
    if (mod < 0) { delta--; mod += dx; }
    acc(ex1, delta, (fx1 + first) * delta);
    ex1 += incr;
    y1 += delta;
    if (ex1 != ex2) {
        p = 256 * (y2 - y1 + delta);
        int lift = p / dx, rem = p % dx;
        if (rem < 0) { lift--; rem += dx; }
        mod -= dx;
        while (ex1 != ex2) {
            delta = lift;
            mod += rem;
            if (mod >= 0) { mod -= dx; delta++; }
            acc(ex1, delta, 256 * delta);
            y1 += delta;
            ex1 += incr;
        }
    }
    delta = y2 - y1;
    acc(ex1, delta, (fx2 + 256 - first) * delta);
}

RIG_HD int fig_floordiv(int a, int b)   // b > 0
{
    int q = a / b;
    return (a % b < 0) ? q - 1 : q;
}

// ---- line(): the rows [r0, r1] of one edge --------------------------------------------------------
// acc.row(ey) is called before the cells of row ey; acc(ex, dcover, darea) per cell.
template <class Acc>
RIG_HD void fig_edge_rows(const FEdge &e, int r0, int r1, Acc &acc)
{
    const int x1 = e.x1, y1 = e.y1, x2 = e.x2, y2 = e.y2;
    const int ey1 = y1 >> 8, ey2 = y2 >> 8, fy1 = y1 & 255, fy2 = y2 & 255;
    if (ey1 == ey2) {
        if (ey1 < r0 || ey1 > r1) return;
        acc.row(ey1);
        fig_hline(x1, fy1, x2, fy2, acc);
        return;
    }
    const int dx = x2 - x1, dy = y2 - y1;
    const int incr = dy > 0 ? 1 : -1, first = dy > 0 ? 256 : 0;
    const int n = dy > 0 ? ey2 - ey1 : ey1 - ey2;     // rows j = 0 .. n, row(j) = ey1 + incr * j
    int ja, jb;
    if (incr > 0) { ja = r0 - ey1; jb = r1 - ey1; }
    else          { ja = ey1 - r1; jb = ey1 - r0; }
    if (ja < 0) ja = 0;
    if (jb > n) jb = n;
    if (ja > jb) return;
    if (dx == 0) {
        const int ex = x1 >> 8, two_fx = (x1 & 255) << 1;
        for (int j = ja; j <= jb; j++) {
            const int d = j == 0 ? first - fy1 : (j == n ? fy2 - 256 + first : first + first - 256);
            acc.row(ey1 + incr * j);
            if (d) acc(ex, d, two_fx * d);
        }
        return;
    }
    const int ady = dy > 0 ? dy : -dy;
    // exit x of row j (j < n):  X_j = x1 + floor((p + j q) / |dy|),  p = (rows' first partial height) * dx, q = 256 dx
    //                               = x1 + delta0 + j lift + floor((mod0 + j rem) / |dy|)
    const int p = (dy > 0 ? 256 - fy1 : fy1) * dx;
    const int delta0 = fig_floordiv(p, ady), mod0 = p - delta0 * ady;
    const int q = 256 * dx;
    const int lift = fig_floordiv(q, ady), rem = q - lift * ady;
    int xfrom = x1, yfrom = fy1, m = mod0;
    if (ja > 0) {
        // mod0, rem < |dy| <= 256 * rows and ja <= rows: for canvases of up to 3072 rows the sum stays below 3072 * 786 432 + 786 432
        // = 2.42e9 < 2^32 (unsigned arithmetic)
        const uint32_t t = (uint32_t)mod0 + (uint32_t)(ja - 1) * (uint32_t)rem;
        const uint32_t qd = t / (uint32_t)ady;
        m = (int)(t - qd * (uint32_t)ady);
        xfrom = x1 + delta0 + (ja - 1) * lift + (int)qd;
        yfrom = 256 - first;
    }
    for (int j = ja; j <= jb; j++) {
        int xto, yto;
        if (j == n) { xto = x2; yto = fy2; }
        else if (j == 0) { xto = x1 + delta0; yto = first; }
        else {
            int d = lift;
            m += rem;
            if (m >= ady) { m -= ady; d++; }
            xto = xfrom + d;
            yto = first;
        }
        acc.row(ey1 + incr * j);
        fig_hline(xfrom, yfrom, xto, yto, acc);
        xfrom = xto;
        yfrom = 256 - first;
    }
}

// ---- sweep_scanline + fixed_blender_rgba_plain ----------------------------------------------------
// coverage of a pixel from the running cover (cells up to and including this one) and the cell's own area
RIG_HD uint32_t fig_alpha(int cover_incl, int area)
{
    int c = ((cover_incl << 9) - area) >> 9;
    if (c < 0) c = -c;
    return (uint32_t)(c > 255 ? 255 : c);
}
RIG_HD uint32_t fig_mult_cover(uint32_t a, uint32_t b)   // rgba8::mult_cover: round(a * b / 255)
{
    const uint32_t t = a * b + 128u;
    return ((t >> 8) + t) >> 8;
}
// one channel over an opaque destination:  (((c << 8) - p * 255) * alpha + ((p * 255) << 8)) / (((alpha + 255) << 8) - alpha * 255)
RIG_HD uint32_t fig_blend1(uint32_t p, uint32_t c, uint32_t alpha)
{
    const uint32_t r = p * 255u;
    return ((((c << 8) - r) * alpha + (r << 8))) / (65280u + alpha);
}
// packed 3-channel pixel (bits 0..23), colour rgba as FLayer::rgba, cover 1..255
RIG_HD uint32_t fig_blend_px(uint32_t px, uint32_t rgba, uint32_t cover)
{
    const uint32_t ca = rgba >> 24;
    if (ca == 255u && cover == 255u) return rgba & 0x00ffffffu;
    const uint32_t alpha = fig_mult_cover(ca, cover);
    if (alpha == 0u) return px;
    const uint32_t v0 = fig_blend1(px & 0xffu, rgba & 0xffu, alpha);
    const uint32_t v1 = fig_blend1((px >> 8) & 0xffu, (rgba >> 8) & 0xffu, alpha);
    const uint32_t v2 = fig_blend1((px >> 16) & 0xffu, (rgba >> 16) & 0xffu, alpha);
    return (v0 & 0xffu) | ((v1 & 0xffu) << 8) | ((v2 & 0xffu) << 16);
}


// ---- the same blend without divisions ---------------------------------------------------------------
// Every numerator is below 2^24 and the divisor 65280 + alpha is below 2^16, so floor(n / d) == (n * ceil(2^40 / d)) >> 40
// exactly (the error term n * (m d - 2^40) < 2^24 * 2^16).  The kernel keeps the 256 multipliers in shared memory.
RIG_HD uint32_t fig_div_magic(uint32_t alpha)
{
    const unsigned long long d = 65280ull + alpha;
    return (uint32_t)(((1ull << 40) + d - 1ull) / d);
}
RIG_HD uint32_t fig_mulhi(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((unsigned long long)a * b) >> 32);
#endif
}
RIG_HD uint32_t fig_blend1_m(uint32_t p, uint32_t c, uint32_t alpha, uint32_t magic)
{
    const uint32_t r = p * 255u;
    return fig_mulhi((((c << 8) - r) * alpha + (r << 8)), magic) >> 8;
}
// magic: table of fig_div_magic(alpha), alpha = 0 .. 255
RIG_HD uint32_t fig_blend_px_m(uint32_t px, uint32_t rgba, uint32_t cover, const uint32_t *magic)
{
    const uint32_t ca = rgba >> 24;
    if (ca == 255u && cover == 255u) return rgba & 0x00ffffffu;
    const uint32_t alpha = fig_mult_cover(ca, cover);
    if (alpha == 0u) return px;
    const uint32_t m = magic[alpha];
    const uint32_t v0 = fig_blend1_m(px & 0xffu, rgba & 0xffu, alpha, m);
    const uint32_t v1 = fig_blend1_m((px >> 8) & 0xffu, (rgba >> 8) & 0xffu, alpha, m);
    const uint32_t v2 = fig_blend1_m((px >> 16) & 0xffu, (rgba >> 16) & 0xffu, alpha, m);
    return v0 | (v1 << 8) | (v2 << 16);
}

}  // namespace rig
