// rig_shapes.h -- element -> integer vertices ("transform prologue"), shared by device and host.
//
// Restates the float64 vertex math of Shape.draw (/root/reference/src/shapes.py:238-258,278-280,
// 358-375,390-544,596-656): centre-relative outline, rotation by -angle (cos/sin supplied by the
// host so they are the libm values Python sees), optional flip, then round-half-even to int32.
// Every product and sum is rounded separately (no FMA) in the reference's operation order.
#pragma once
#include "rig_core.h"

namespace rig {

enum Kind : int {
    K_SQUARE = 0, K_CIRCLE, K_TRIANGLE, K_DIAMOND, K_STAR, K_PENTAGON, K_HEXAGON, K_PLUS, K_HEART,
    K_CRESCENT, K_ROUNDED_SQUARE, K_COUNT
};

constexpr int MAX_RING_VERTS = 74;   // 5-degree circle polygon: 73 points
constexpr int MAX_STROKE = 4;        // caps stay 90-degree diamonds up to thickness 4

// Unit tables, filled on the host with the same libm calls CPython makes (rig_tables.cpp)
struct ShapeTables {
    double reg_cs[3][6][2];   // [star, pentagon, hexagon][i] -> cos, sin of radians(i*step - 90)
    double heart[60][2];      // 16 sin^3 t ; 13 cos t - 5 cos 2t - 2 cos 3t - cos 4t
    double arc_cs[48][2];     // rounded_square corner arcs: cos, sin of the 4 x 12 linspace angles
    float sin_tab[451];       // OpenCV SinTable (degrees)
    uint8_t slope_corr[32];   // OpenCV icvSlopeCorrTable
    uint8_t filter[64];       // OpenCV FilterTable
};

// element as uploaded to the device (40 bytes)
struct DElem {
    double ca, sa;            // cos / sin of radians(-angle)
    int32_t cx, cy;           // centre after int() truncation and grid snap
    int32_t size;
    int16_t stroke;           // max(1, stroke_width)
    uint8_t kind, flags;      // flags bit0 fill, bit1 flip-x, bit2 flip-y
    uint8_t c0, c1, c2, pad;  // fill colour in output channel order
    uint32_t pad2;
};
static_assert(sizeof(DElem) == 40, "DElem layout");

// what the expand kernel leaves for the rasteriser, per element
struct ElemHdr {
    uint32_t rec_begin, nrec; // arena range: nsum summary slots at rec_begin, then nrec payload slots
    int16_t ymin, ymax;       // union of the records' row ranges (ymin > ymax: nothing visible)
    uint32_t nsum;
};

RIG_HD int kind_loops(int kind) { return kind == K_PLUS ? 2 : 1; }
RIG_HD int kind_nverts(int kind)
{
    switch (kind) {
    case K_SQUARE: case K_DIAMOND: case K_PLUS: return 4;
    case K_TRIANGLE: return 3;
    case K_STAR: case K_PENTAGON: return 5;
    case K_HEXAGON: return 6;
    case K_HEART: return 60;
    case K_ROUNDED_SQUARE: return 52;
    default: return 0;
    }
}

// centre-relative, unrotated vertex i of loop `loop`
RIG_FN void unit_vertex(const ShapeTables &T, int kind, int loop, int i, int s, double &x, double &y)
{
    const double half = ddiv((double)s, 2.0);
    switch (kind) {
    case K_SQUARE: {
        x = (i == 0 || i == 3) ? -half : half;
        y = (i < 2) ? -half : half;
        break;
    }
    case K_TRIANGLE: {
        x = i == 0 ? -half : (i == 1 ? 0.0 : half);
        y = i == 1 ? -half : half;
        break;
    }
    case K_DIAMOND: {
        x = i == 1 ? half : (i == 3 ? -half : 0.0);
        y = i == 0 ? -half : (i == 2 ? half : 0.0);
        break;
    }
    case K_STAR: case K_PENTAGON: case K_HEXAGON: {
        int t = kind == K_STAR ? 0 : (kind == K_PENTAGON ? 1 : 2);
        x = dmul(half, T.reg_cs[t][i][0]);
        y = dmul(half, T.reg_cs[t][i][1]);
        break;
    }
    case K_PLUS: {
        int arm = dround(dmul((double)s, 0.25));
        int len = dround(dmul((double)s, 0.9));
        double a2 = ddiv((double)arm, 2.0), l2 = ddiv((double)len, 2.0);
        double hx = loop == 0 ? a2 : l2, hy = loop == 0 ? l2 : a2;
        x = (i == 0 || i == 3) ? -hx : hx;
        y = (i < 2) ? -hy : hy;
        break;
    }
    case K_HEART: {
        double r = dmul(half, 0.8);
        double scale = ddiv(r, 16.0);
        x = dmul(T.heart[i][0], scale);
        y = dmul(-T.heart[i][1], scale);
        break;
    }
    case K_ROUNDED_SQUARE: {
        int r = dround(dmul(half, 0.4));
        int r2 = dround(dsub(half, 1.0));
        if (r2 < r) r = r2;
        double lo = dadd(-half, (double)r), hi = dsub(half, (double)r);
        int corner = i / 13, j = i % 13;   // 12 arc points then one straight-edge point
        double ccx = (corner == 0 || corner == 3) ? lo : hi;
        double ccy = (corner < 2) ? lo : hi;
        if (j < 12) {
            x = dadd(ccx, dmul((double)r, T.arc_cs[corner * 12 + j][0]));
            y = dadd(ccy, dmul((double)r, T.arc_cs[corner * 12 + j][1]));
        } else if (corner == 0) { x = hi; y = dsub(lo, (double)r); }
        else if (corner == 1)   { x = dadd(hi, (double)r); y = hi; }
        else if (corner == 2)   { x = lo; y = dadd(hi, (double)r); }
        else                    { x = dsub(lo, (double)r); y = lo; }
        break;
    }
    default: x = y = 0.0;
    }
}

RIG_HD void place_vertex(double x, double y, const DElem &e, int &ox, int &oy)
{
    double xr = dsub(dmul(x, e.ca), dmul(y, e.sa));
    double yr = dadd(dmul(x, e.sa), dmul(y, e.ca));
    if (e.flags & 2) xr = -xr;
    if (e.flags & 4) yr = -yr;
    ox = dround(dadd((double)e.cx, xr));
    oy = dround(dadd((double)e.cy, yr));
}

// EllipseEx polygon of a circle outline: centre (cx,cy) px, radius r px -> 16.16 vertices.
// Returns the vertex count after OpenCV's consecutive-duplicate removal (>= 2).
RIG_HD int ring_delta(long long rad_fx)
{
    int d = (int)((rad_fx + (XY_ONE >> 1)) >> XY_SHIFT);
    return d < 3 ? 90 : d < 10 ? 30 : d < 15 ? 18 : 5;
}
RIG_FN P2L ring_vertex(const float *sin_tab, long long cx_fx, long long cy_fx, long long rad_fx, int angle_deg)
{
    double x = dmul((double)rad_fx, (double)sin_tab[450 - angle_deg]);
    double y = dmul((double)rad_fx, (double)sin_tab[angle_deg]);
    double px = dsub(dadd((double)cx_fx, dmul(x, 1.0)), dmul(y, 0.0));
    double py = dadd(dadd((double)cy_fx, dmul(x, 0.0)), dmul(y, 1.0));
    P2L pt;
    pt.x = (long long)dround(dscale_down16(px)) * XY_ONE;
    pt.y = (long long)dround(dscale_down16(py)) * XY_ONE;
    pt.x += dround(dsub(px, (double)pt.x));
    pt.y += dround(dsub(py, (double)pt.y));
    return pt;
}

// wrap-around copies (shapes.py:260-266): bit (3*ix + iy) set when copy (ox,oy) = ((ix-1)W,(iy-1)H) is drawn
RIG_HD uint32_t copy_mask(int minx, int maxx, int miny, int maxy, int W, int H, bool origin_always)
{
    uint32_t m = 0;
    for (int ix = 0; ix < 3; ix++)
        for (int iy = 0; iy < 3; iy++) {
            int ox = (ix - 1) * W, oy = (iy - 1) * H;
            bool hit = !(maxx + ox < 0 || minx + ox >= W) && !(maxy + oy < 0 || miny + oy >= H);
            if (hit || (origin_always && ix == 1 && iy == 1)) m |= 1u << (3 * ix + iy);
        }
    return m;
}

// record slots of one stroked segment
RIG_HD int stroke_recs(int t) { return t <= 1 ? 1 : QUAD_SLOTS + CAP_SLOTS; }

}  // namespace rig
