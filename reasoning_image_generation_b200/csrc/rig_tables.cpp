// rig_tables.cpp -- host-side constants of the rasteriser.
//
// The unit shape tables are evaluated with the same libm entry points, in the same operation
// order, as the CPython expressions of the reference (/root/reference/src/shapes.py:431-435,
// 456-460,518-527,613-637), so that the device prologue reproduces Python's float64 vertices bit
// for bit.  The three OpenCV tables are the published constants of drawing.cpp's LineAA /
// ellipse2Poly (icvSlopeCorrTable, FilterTable, SinTable); SinTable is round(sin(deg), 7) parsed
// as a float literal.
#include "rig_host.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace rig {

static const double kPi = 3.141592653589793;  // math.pi

static double py_radians(double deg) { return deg * (kPi / 180.0); }

// numpy.linspace(start, stop, num)[i] with endpoint=True
static double np_linspace(double start, double stop, int num, int i)
{
    if (i == num - 1) return stop;
    double step = (stop - start) / (double)(num - 1);
    volatile double v = (double)i * step;
    return v + start;
}

void fill_shape_tables(ShapeTables &T)
{
    // 'star': radians(i*72 - 90) with an int argument; pentagon/hexagon: radians(i*(360.0/n) - 90)
    for (int i = 0; i < 6; i++) {
        double a_star = py_radians((double)(i * 72 - 90));
        volatile double s5 = 360.0 / 5, s6 = 360.0 / 6;
        volatile double m5 = (double)i * s5, m6 = (double)i * s6;
        double a_pent = py_radians(m5 - 90);
        double a_hex = py_radians(m6 - 90);
        T.reg_cs[0][i][0] = cos(a_star); T.reg_cs[0][i][1] = sin(a_star);
        T.reg_cs[1][i][0] = cos(a_pent); T.reg_cs[1][i][1] = sin(a_pent);
        T.reg_cs[2][i][0] = cos(a_hex);  T.reg_cs[2][i][1] = sin(a_hex);
    }
    volatile double two_pi = 2 * kPi;
    for (int i = 0; i < 60; i++) {
        double t = np_linspace(0.0, two_pi, 60, i);
        volatile double s3 = pow(sin(t), 3.0);
        volatile double x = 16 * s3;
        volatile double t2 = 2 * t, t3 = 3 * t, t4 = 4 * t;
        volatile double a = 13 * cos(t), b = 5 * cos(t2), c = 2 * cos(t3), d = cos(t4);
        volatile double y = a - b;
        y = y - c;
        y = y - d;
        T.heart[i][0] = x;
        T.heart[i][1] = y;
    }
    volatile double three_pi = 3 * kPi;
    const double q[5] = {kPi, three_pi / 2, two_pi, kPi / 2, kPi};
    // corner arcs in outline order: pi..3pi/2, 3pi/2..2pi, 0..pi/2, pi/2..pi
    const double starts[4] = {q[0], q[1], 0.0, q[3]};
    const double stops[4] = {q[1], q[2], q[3], q[4]};
    for (int c = 0; c < 4; c++)
        for (int j = 0; j < 12; j++) {
            double th = np_linspace(starts[c], stops[c], 12, j);
            T.arc_cs[c * 12 + j][0] = cos(th);
            T.arc_cs[c * 12 + j][1] = sin(th);
        }
    for (int d = 0; d <= 450; d++) {
        char buf[32];
        snprintf(buf, sizeof buf, "%.7f", sin(py_radians((double)d)));
        T.sin_tab[d] = strtof(buf, nullptr);
    }
    static const uint8_t slope_corr[32] = {181, 181, 181, 182, 182, 183, 184, 185, 187, 188, 190, 192, 194, 196, 198, 201,
                                           203, 206, 209, 211, 214, 218, 221, 224, 227, 231, 235, 238, 242, 246, 250, 254};
    static const uint8_t filter[64] = {168, 177, 185, 194, 202, 210, 218, 224, 231, 236, 241, 246, 249, 252, 254, 254,
                                       254, 254, 252, 249, 246, 241, 236, 231, 224, 218, 210, 202, 194, 185, 177, 168,
                                       158, 149, 140, 131, 122, 114, 105, 97,  89,  82,  75,  68,  62,  56,  50,  45,
                                       40,  36,  32,  28,  25,  22,  19,  16,  14,  12,  11,  9,   8,   7,   5,   5};
    for (int i = 0; i < 32; i++) T.slope_corr[i] = slope_corr[i];
    for (int i = 0; i < 64; i++) T.filter[i] = filter[i];
}

// CPython float floor division (Objects/floatobject.c, float_floor_div)
static double py_floordiv(double vx, double wx)
{
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    if (mod) {
        if ((wx < 0) != (mod < 0)) div -= 1.0;
    }
    double floordiv;
    if (div) {
        floordiv = floor(div);
        if (div - floordiv > 0.5) floordiv += 1.0;
    } else
        floordiv = copysign(0.0, vx / wx);
    return floordiv;
}

int pack_elements(const rig_scene &s, int W, int H, int layout_bgr, DElem *out, char *err, size_t errlen)
{
    const int gs = s.grid_size > 0 ? s.grid_size : 3;
    volatile double cell_w = (double)W / (double)gs, cell_h = (double)H / (double)gs;
    for (int p = 0; p < s.n_panels; p++) {
        const bool use_grid = s.panel_flags && (s.panel_flags[p] & 1);
        for (uint32_t i = s.panel_elem_begin[p]; i < s.panel_elem_begin[p + 1]; i++) {
            DElem &e = out[i];
            int kind = s.kind[i];
            if (kind < 0 || kind >= K_COUNT) {
                snprintf(err, errlen, "element %u: unsupported shape kind %d", i, kind);
                return RIG_E_INVALID;
            }
            int stroke = s.stroke[i] < 1 ? 1 : s.stroke[i];
            if (stroke > MAX_STROKE) {
                snprintf(err, errlen, "element %u: stroke_width %d > %d is not supported", i, stroke, MAX_STROKE);
                return RIG_E_UNSUPPORTED;
            }
            double cx = s.cx[i], cy = s.cy[i];
            int icx, icy;
            if (use_grid) {
                // generator.py:97-100
                double colf = py_floordiv(cx, cell_w), rowf = py_floordiv(cy, cell_h);
                int col = (int)colf, row = (int)rowf;
                col = col < 0 ? 0 : col; col = col > gs - 1 ? gs - 1 : col;
                row = row < 0 ? 0 : row; row = row > gs - 1 ? gs - 1 : row;
                volatile double fx = ((double)col + 0.5) * cell_w, fy = ((double)row + 0.5) * cell_h;
                icx = (int)fx; icy = (int)fy;
            } else {
                icx = (int)cx; icy = (int)cy;   // shapes.py:231 int() truncates toward zero
            }
            if (s.size[i] < 0 || s.size[i] > 4096 || icx < -8192 || icx > 8192 || icy < -8192 || icy > 8192) {
                snprintf(err, errlen, "element %u: size/centre out of the supported range", i);
                return RIG_E_UNSUPPORTED;
            }
            double a = py_radians((double)(-(long long)s.angle[i]));   // math.radians(-angle)
            e.ca = cos(a);
            e.sa = sin(a);
            e.cx = icx; e.cy = icy;
            e.size = s.size[i];
            e.stroke = (int16_t)stroke;
            e.kind = (uint8_t)kind;
            e.flags = (uint8_t)((s.fill[i] ? 1 : 0) | ((s.flip ? s.flip[i] : 0) & 3) << 1);
            uint8_t r = s.rgb[3 * i], g = s.rgb[3 * i + 1], b = s.rgb[3 * i + 2];
            if (layout_bgr) { e.c0 = b; e.c1 = g; e.c2 = r; }
            else            { e.c0 = r; e.c1 = g; e.c2 = b; }
            e.pad = 0; e.pad2 = 0;
        }
    }
    return RIG_OK;
}

}  // namespace rig
