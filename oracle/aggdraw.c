/* oracle/aggdraw.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of what the reference's multigraph path does to pixels:
 *     GeometryGenerator._save_figure -> fig.savefig(dpi)      /root/reference/multigraph_generation/generator.py:250-259
 * i.e. matplotlib's Agg backend drawing ax.patches / ax.lines (style.py:29-66,121-139; single_variants.py:285-288,
 * 313-333,352-374,484-508,552-633).  matplotlib and its bundled Anti-Grain Geometry (extern/agg24-svn) are third-party
 * dependencies that are ABSENT from /root/reference, unpinned in its requirements.txt and not installable here (no wheel,
 * no network), so this file restates their PUBLISHED algorithm from the upstream sources' structure:
 *
 *   RendererAgg::draw_path (src/_backend_agg.h)     trans *= scale(1,-1); trans *= translate(0,H); the converter chain
 *                                                   transform -> [nan remover, clipper: identity for finite in-canvas
 *                                                   geometry] -> PathSnapper -> [simplifier: identity below 128
 *                                                   vertices] -> conv_curve -> (conv_stroke) -> rasterizer
 *   PathSnapper (src/path_converters.h)             rectilinear paths snap to pixel centres / edges
 *   agg::curve3_div / curve4_div (agg_curves.cpp)   recursive subdivision, approximation_scale 1, angle_tolerance 0
 *   agg::vcgen_stroke + math_stroke (agg_math_stroke.h)   joins (miter_revert / round / bevel), caps, inner_miter,
 *                                                   miter_limit = line width in pixels (as _draw_path sets it)
 *   agg::rasterizer_sl_clip<ras_conv_dbl>           clip box in doubles, iround(v * 256) afterwards
 *   agg::rasterizer_cells_aa (agg_rasterizer_cells_aa.h)   24.8 cells with cover / area, line() and render_hline()
 *   agg::rasterizer_scanline_aa::sweep_scanline     non-zero winding, alpha = min(|area| >> 9, 255), linear gamma
 *   fixed_blender_rgba_plain (src/_backend_agg_basic_types.h) + rgba8::mult_cover   non-premultiplied blend
 *
 * PARITY STATUS: UNPINNED.  No fixture rendered by a real matplotlib exists and none can be produced in this container.
 * The known-answer tests in tests/test_figure_oracle.py pin the integer arithmetic against values that follow from the
 * definitions (exact box coverage, symmetry, closed-form areas) only.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846

enum { CMD_STOP = 0, CMD_MOVE = 1, CMD_LINE = 2, CMD_CURVE3 = 3, CMD_CURVE4 = 4, CMD_CLOSE = 0x4F };
enum { CAP_BUTT = 0, CAP_SQUARE = 1, CAP_ROUND = 2 };
enum { JOIN_MITER = 0, JOIN_ROUND = 1, JOIN_BEVEL = 2 };
enum { SNAP_FALSE = 0, SNAP_TRUE = 1, SNAP_AUTO = 2 };

typedef struct { double x, y; int cmd; } Vtx;
typedef struct { Vtx *v; int n, cap; } VBuf;

static void vb_push(VBuf *b, double x, double y, int cmd)
{
    if (b->n == b->cap) {
        b->cap = b->cap ? b->cap * 2 : 64;
        b->v = (Vtx *)realloc(b->v, sizeof(Vtx) * (size_t)b->cap);
    }
    b->v[b->n].x = x; b->v[b->n].y = y; b->v[b->n].cmd = cmd;
    b->n++;
}
static int is_vertex(int c) { return c >= CMD_MOVE && c < 0x0F; }

/* ---- agg::trans_affine ------------------------------------------------------------------------ */
typedef struct { double sx, shy, shx, sy, tx, ty; } Affine;
static void affine_multiply(Affine *a, const Affine *m)
{
    double t0 = a->sx * m->sx + a->shy * m->shx;
    double t2 = a->shx * m->sx + a->sy * m->shx;
    double t4 = a->tx * m->sx + a->ty * m->shx + m->tx;
    a->shy = a->sx * m->shy + a->shy * m->sy;
    a->sy = a->shx * m->shy + a->sy * m->sy;
    a->ty = a->tx * m->shy + a->ty * m->sy + m->ty;
    a->sx = t0; a->shx = t2; a->tx = t4;
}
static void affine_apply(const Affine *a, double *x, double *y)
{
    double tmp = *x;
    *x = tmp * a->sx + *y * a->shx + a->tx;
    *y = tmp * a->shy + *y * a->sy + a->ty;
}

/* ---- PathSnapper ------------------------------------------------------------------------------ */
static int should_snap(const VBuf *p, int mode, int total_vertices)
{
    if (mode == SNAP_FALSE) return 0;
    if (mode == SNAP_TRUE) return 1;
    if (total_vertices > 1024 || p->n == 0) return 0;
    double x0 = p->v[0].x, y0 = p->v[0].y;
    for (int i = 1; i < p->n; i++) {
        int c = p->v[i].cmd;
        double x1 = p->v[i].x, y1 = p->v[i].y;
        if (c == CMD_CURVE3 || c == CMD_CURVE4) return 0;
        if (c == CMD_LINE && !(fabs(x0 - x1) < 1e-4 || fabs(y0 - y1) < 1e-4)) return 0;
        x0 = x1; y0 = y1;
    }
    return 1;
}

/* ---- agg::curve3_div / curve4_div ------------------------------------------------------------- */
#define CURVE_COLLINEARITY_EPS 1e-30
#define CURVE_RECURSION_LIMIT 32
static double sq_dist(double x1, double y1, double x2, double y2)
{
    double dx = x2 - x1, dy = y2 - y1;
    return dx * dx + dy * dy;
}
static void curve3_rec(VBuf *o, double tol2, double x1, double y1, double x2, double y2, double x3, double y3, int level)
{
    if (level > CURVE_RECURSION_LIMIT) return;
    double x12 = (x1 + x2) / 2, y12 = (y1 + y2) / 2, x23 = (x2 + x3) / 2, y23 = (y2 + y3) / 2;
    double x123 = (x12 + x23) / 2, y123 = (y12 + y23) / 2;
    double dx = x3 - x1, dy = y3 - y1;
    double d = fabs(((x2 - x3) * dy - (y2 - y3) * dx));
    if (d > CURVE_COLLINEARITY_EPS) {
        if (d * d <= tol2 * (dx * dx + dy * dy)) { vb_push(o, x123, y123, CMD_LINE); return; }
    } else {
        double da = dx * dx + dy * dy;
        if (da == 0) d = sq_dist(x1, y1, x2, y2);
        else {
            d = ((x2 - x1) * dx + (y2 - y1) * dy) / da;
            if (d > 0 && d < 1) return;
            if (d <= 0) d = sq_dist(x2, y2, x1, y1);
            else if (d >= 1) d = sq_dist(x2, y2, x3, y3);
            else d = sq_dist(x2, y2, x1 + d * dx, y1 + d * dy);
        }
        if (d < tol2) { vb_push(o, x2, y2, CMD_LINE); return; }
    }
    curve3_rec(o, tol2, x1, y1, x12, y12, x123, y123, level + 1);
    curve3_rec(o, tol2, x123, y123, x23, y23, x3, y3, level + 1);
}
static void curve4_rec(VBuf *o, double tol2, double x1, double y1, double x2, double y2, double x3, double y3, double x4,
                       double y4, int level)
{
    if (level > CURVE_RECURSION_LIMIT) return;
    double x12 = (x1 + x2) / 2, y12 = (y1 + y2) / 2, x23 = (x2 + x3) / 2, y23 = (y2 + y3) / 2;
    double x34 = (x3 + x4) / 2, y34 = (y3 + y4) / 2;
    double x123 = (x12 + x23) / 2, y123 = (y12 + y23) / 2, x234 = (x23 + x34) / 2, y234 = (y23 + y34) / 2;
    double x1234 = (x123 + x234) / 2, y1234 = (y123 + y234) / 2;
    double dx = x4 - x1, dy = y4 - y1;
    double d2 = fabs(((x2 - x4) * dy - (y2 - y4) * dx));
    double d3 = fabs(((x3 - x4) * dy - (y3 - y4) * dx));
    double da1, da2, k;
    switch (((d2 > CURVE_COLLINEARITY_EPS) << 1) + (d3 > CURVE_COLLINEARITY_EPS)) {
    case 0:
        k = dx * dx + dy * dy;
        if (k == 0) {
            d2 = sq_dist(x1, y1, x2, y2);
            d3 = sq_dist(x4, y4, x3, y3);
        } else {
            k = 1 / k;
            da1 = x2 - x1; da2 = y2 - y1;
            d2 = k * (da1 * dx + da2 * dy);
            da1 = x3 - x1; da2 = y3 - y1;
            d3 = k * (da1 * dx + da2 * dy);
            if (d2 > 0 && d2 < 1 && d3 > 0 && d3 < 1) return;
            if (d2 <= 0) d2 = sq_dist(x2, y2, x1, y1);
            else if (d2 >= 1) d2 = sq_dist(x2, y2, x4, y4);
            else d2 = sq_dist(x2, y2, x1 + d2 * dx, y1 + d2 * dy);
            if (d3 <= 0) d3 = sq_dist(x3, y3, x1, y1);
            else if (d3 >= 1) d3 = sq_dist(x3, y3, x4, y4);
            else d3 = sq_dist(x3, y3, x1 + d3 * dx, y1 + d3 * dy);
        }
        if (d2 > d3) {
            if (d2 < tol2) { vb_push(o, x2, y2, CMD_LINE); return; }
        } else {
            if (d3 < tol2) { vb_push(o, x3, y3, CMD_LINE); return; }
        }
        break;
    case 1:
        if (d3 * d3 <= tol2 * (dx * dx + dy * dy)) { vb_push(o, x23, y23, CMD_LINE); return; }
        break;
    case 2:
        if (d2 * d2 <= tol2 * (dx * dx + dy * dy)) { vb_push(o, x23, y23, CMD_LINE); return; }
        break;
    case 3:
        if ((d2 + d3) * (d2 + d3) <= tol2 * (dx * dx + dy * dy)) { vb_push(o, x23, y23, CMD_LINE); return; }
        break;
    }
    curve4_rec(o, tol2, x1, y1, x12, y12, x123, y123, x1234, y1234, level + 1);
    curve4_rec(o, tol2, x1234, y1234, x234, y234, x34, y34, x4, y4, level + 1);
}
/* agg::conv_curve: the curve's own first point is skipped, every later point is a line_to */
static void conv_curve(const VBuf *in, VBuf *out)
{
    const double tol2 = 0.5 * 0.5;   /* (0.5 / approximation_scale)^2, approximation_scale = 1 */
    double lx = 0, ly = 0;
    for (int i = 0; i < in->n; i++) {
        const Vtx v = in->v[i];
        if (v.cmd == CMD_CURVE3 && i + 1 < in->n) {
            const Vtx e = in->v[++i];
            curve3_rec(out, tol2, lx, ly, v.x, v.y, e.x, e.y, 0);
            vb_push(out, e.x, e.y, CMD_LINE);
            lx = e.x; ly = e.y;
        } else if (v.cmd == CMD_CURVE4 && i + 2 < in->n) {
            const Vtx c2 = in->v[++i];
            const Vtx e = in->v[++i];
            curve4_rec(out, tol2, lx, ly, v.x, v.y, c2.x, c2.y, e.x, e.y, 0);
            vb_push(out, e.x, e.y, CMD_LINE);
            lx = e.x; ly = e.y;
        } else {
            vb_push(out, v.x, v.y, v.cmd);
            if (is_vertex(v.cmd)) { lx = v.x; ly = v.y; }
        }
    }
}

/* ---- agg::math_stroke ------------------------------------------------------------------------- */
#define VERTEX_DIST_EPS 1e-14
#define INTERSECTION_EPS 1.0e-30
typedef struct { double x, y, dist; } VD;
typedef struct {
    double width, width_abs, width_eps, miter_limit, inner_miter_limit, approx_scale;
    int width_sign, cap, join;
} Stroker;
static double cross_product(double x1, double y1, double x2, double y2, double x, double y)
{
    return (x - x2) * (y2 - y1) - (y - y2) * (x2 - x1);
}
static int calc_intersection(double ax, double ay, double bx, double by, double cx, double cy, double dx, double dy,
                             double *x, double *y)
{
    double num = (ay - cy) * (dx - cx) - (ax - cx) * (dy - cy);
    double den = (bx - ax) * (dy - cy) - (by - ay) * (dx - cx);
    if (fabs(den) < INTERSECTION_EPS) return 0;
    double r = num / den;
    *x = ax + r * (bx - ax);
    *y = ay + r * (by - ay);
    return 1;
}
static void calc_arc(const Stroker *s, VBuf *vc, double x, double y, double dx1, double dy1, double dx2, double dy2)
{
    double a1 = atan2(dy1 * s->width_sign, dx1 * s->width_sign);
    double a2 = atan2(dy2 * s->width_sign, dx2 * s->width_sign);
    double da = a1 - a2;
    int i, n;
    da = acos(s->width_abs / (s->width_abs + 0.125 / s->approx_scale)) * 2;
    vb_push(vc, x + dx1, y + dy1, CMD_LINE);
    if (s->width_sign > 0) {
        if (a1 > a2) a2 += 2 * PI;
        n = (int)((a2 - a1) / da);
        da = (a2 - a1) / (n + 1);
        a1 += da;
        for (i = 0; i < n; i++) {
            vb_push(vc, x + cos(a1) * s->width, y + sin(a1) * s->width, CMD_LINE);
            a1 += da;
        }
    } else {
        if (a1 < a2) a2 -= 2 * PI;
        n = (int)((a1 - a2) / da);
        da = (a1 - a2) / (n + 1);
        a1 -= da;
        for (i = 0; i < n; i++) {
            vb_push(vc, x + cos(a1) * s->width, y + sin(a1) * s->width, CMD_LINE);
            a1 -= da;
        }
    }
    vb_push(vc, x + dx2, y + dy2, CMD_LINE);
}
enum { LJ_MITER = 0, LJ_MITER_REVERT = 1, LJ_ROUND = 2, LJ_BEVEL = 3 };
static void calc_miter(const Stroker *s, VBuf *vc, const VD *v0, const VD *v1, const VD *v2, double dx1, double dy1,
                       double dx2, double dy2, int lj, double mlimit, double dbevel)
{
    double xi = v1->x, yi = v1->y, di = 1;
    double lim = s->width_abs * mlimit;
    int miter_limit_exceeded = 1, intersection_failed = 1;
    if (calc_intersection(v0->x + dx1, v0->y - dy1, v1->x + dx1, v1->y - dy1, v1->x + dx2, v1->y - dy2, v2->x + dx2,
                          v2->y - dy2, &xi, &yi)) {
        di = sqrt(sq_dist(v1->x, v1->y, xi, yi));
        if (di <= lim) {
            vb_push(vc, xi, yi, CMD_LINE);
            miter_limit_exceeded = 0;
        }
        intersection_failed = 0;
    } else {
        double x2 = v1->x + dx1, y2 = v1->y - dy1;
        if ((cross_product(v0->x, v0->y, v1->x, v1->y, x2, y2) < 0.0) ==
            (cross_product(v1->x, v1->y, v2->x, v2->y, x2, y2) < 0.0)) {
            vb_push(vc, v1->x + dx1, v1->y - dy1, CMD_LINE);
            miter_limit_exceeded = 0;
        }
    }
    if (miter_limit_exceeded) {
        if (lj == LJ_MITER_REVERT) {
            vb_push(vc, v1->x + dx1, v1->y - dy1, CMD_LINE);
            vb_push(vc, v1->x + dx2, v1->y - dy2, CMD_LINE);
        } else if (intersection_failed) {
            mlimit *= s->width_sign;
            vb_push(vc, v1->x + dx1 + dy1 * mlimit, v1->y - dy1 + dx1 * mlimit, CMD_LINE);
            vb_push(vc, v1->x + dx2 - dy2 * mlimit, v1->y - dy2 - dx2 * mlimit, CMD_LINE);
        } else {
            double x1 = v1->x + dx1, y1 = v1->y - dy1, x2 = v1->x + dx2, y2 = v1->y - dy2;
            di = (lim - dbevel) / (di - dbevel);
            vb_push(vc, x1 + (xi - x1) * di, y1 + (yi - y1) * di, CMD_LINE);
            vb_push(vc, x2 + (xi - x2) * di, y2 + (yi - y2) * di, CMD_LINE);
        }
    }
}
static void calc_cap(const Stroker *s, VBuf *vc, const VD *v0, const VD *v1, double len)
{
    vc->n = 0;
    double dx1 = (v1->y - v0->y) / len, dy1 = (v1->x - v0->x) / len, dx2 = 0, dy2 = 0;
    dx1 *= s->width;
    dy1 *= s->width;
    if (s->cap != CAP_ROUND) {
        if (s->cap == CAP_SQUARE) {
            dx2 = dy1 * s->width_sign;
            dy2 = dx1 * s->width_sign;
        }
        vb_push(vc, v0->x - dx1 - dx2, v0->y + dy1 - dy2, CMD_LINE);
        vb_push(vc, v0->x + dx1 - dx2, v0->y - dy1 - dy2, CMD_LINE);
    } else {
        double da = acos(s->width_abs / (s->width_abs + 0.125 / s->approx_scale)) * 2;
        double a1;
        int i, n = (int)(PI / da);
        da = PI / (n + 1);
        vb_push(vc, v0->x - dx1, v0->y + dy1, CMD_LINE);
        if (s->width_sign > 0) {
            a1 = atan2(dy1, -dx1);
            a1 += da;
            for (i = 0; i < n; i++) {
                vb_push(vc, v0->x + cos(a1) * s->width, v0->y + sin(a1) * s->width, CMD_LINE);
                a1 += da;
            }
        } else {
            a1 = atan2(-dy1, dx1);
            a1 -= da;
            for (i = 0; i < n; i++) {
                vb_push(vc, v0->x + cos(a1) * s->width, v0->y + sin(a1) * s->width, CMD_LINE);
                a1 -= da;
            }
        }
        vb_push(vc, v0->x + dx1, v0->y - dy1, CMD_LINE);
    }
}
static void calc_join(const Stroker *s, VBuf *vc, const VD *v0, const VD *v1, const VD *v2, double len1, double len2)
{
    double dx1 = s->width * (v1->y - v0->y) / len1;
    double dy1 = s->width * (v1->x - v0->x) / len1;
    double dx2 = s->width * (v2->y - v1->y) / len2;
    double dy2 = s->width * (v2->x - v1->x) / len2;
    vc->n = 0;
    double cp = cross_product(v0->x, v0->y, v1->x, v1->y, v2->x, v2->y);
    if ((cp > VERTEX_DIST_EPS && s->width > 0) || (cp < -VERTEX_DIST_EPS && s->width < 0)) {
        /* inner join: agg's default inner_miter */
        double limit = ((len1 < len2) ? len1 : len2) / s->width_abs;
        if (limit < s->inner_miter_limit) limit = s->inner_miter_limit;
        calc_miter(s, vc, v0, v1, v2, dx1, dy1, dx2, dy2, LJ_MITER_REVERT, limit, 0);
    } else {
        double dx = (dx1 + dx2) / 2, dy = (dy1 + dy2) / 2;
        double dbevel = sqrt(dx * dx + dy * dy);
        if (s->join == LJ_ROUND || s->join == LJ_BEVEL) {
            if (s->approx_scale * (s->width_abs - dbevel) < s->width_eps) {
                if (calc_intersection(v0->x + dx1, v0->y - dy1, v1->x + dx1, v1->y - dy1, v1->x + dx2, v1->y - dy2,
                                      v2->x + dx2, v2->y - dy2, &dx, &dy))
                    vb_push(vc, dx, dy, CMD_LINE);
                else
                    vb_push(vc, v1->x + dx1, v1->y - dy1, CMD_LINE);
                return;
            }
        }
        switch (s->join) {
        case LJ_MITER:
        case LJ_MITER_REVERT:
            calc_miter(s, vc, v0, v1, v2, dx1, dy1, dx2, dy2, s->join, s->miter_limit, dbevel);
            break;
        case LJ_ROUND:
            calc_arc(s, vc, v1->x, v1->y, dx1, -dy1, dx2, -dy2);
            break;
        default:
            vb_push(vc, v1->x + dx1, v1->y - dy1, CMD_LINE);
            vb_push(vc, v1->x + dx2, v1->y - dy2, CMD_LINE);
            break;
        }
    }
}

/* ---- agg::vertex_sequence<vertex_dist> + vcgen_stroke ----------------------------------------- */
typedef struct { VD *v; int n, cap; } VSeq;
static int vd_measure(VD *a, const VD *b)
{
    a->dist = sqrt(sq_dist(a->x, a->y, b->x, b->y));
    int ret = a->dist > VERTEX_DIST_EPS;
    if (!ret) a->dist = 1.0 / VERTEX_DIST_EPS;
    return ret;
}
static void vs_add(VSeq *s, double x, double y)
{
    if (s->n > 1 && !vd_measure(&s->v[s->n - 2], &s->v[s->n - 1])) s->n--;
    if (s->n == s->cap) {
        s->cap = s->cap ? s->cap * 2 : 64;
        s->v = (VD *)realloc(s->v, sizeof(VD) * (size_t)s->cap);
    }
    s->v[s->n].x = x; s->v[s->n].y = y; s->v[s->n].dist = 0;
    s->n++;
}
static void vs_modify_last(VSeq *s, double x, double y)
{
    if (s->n) s->n--;
    vs_add(s, x, y);
}
static void vs_close(VSeq *s, int closed)
{
    while (s->n > 1) {
        if (vd_measure(&s->v[s->n - 2], &s->v[s->n - 1])) break;
        VD t = s->v[s->n - 1];
        s->n--;
        vs_modify_last(s, t.x, t.y);
    }
    if (closed) {
        while (s->n > 1) {
            if (vd_measure(&s->v[s->n - 1], &s->v[0])) break;
            s->n--;
        }
    }
}
static void emit_contour(VBuf *out, const VBuf *vc, int *first)
{
    for (int i = 0; i < vc->n; i++) {
        vb_push(out, vc->v[i].x, vc->v[i].y, *first ? CMD_MOVE : CMD_LINE);
        *first = 0;
    }
}
static void stroke_subpath(const Stroker *st, VSeq *s, int closed, VBuf *out)
{
    vs_close(s, closed);
    if (s->n < 3) closed = 0;
    if (s->n < 2 + (closed ? 1 : 0)) return;
    VBuf vc = {0, 0, 0};
    const int n = s->n;
    int first = 1;
    if (closed) {
        for (int i = 0; i < n; i++) {
            const VD *p = &s->v[(i + n - 1) % n], *c = &s->v[i], *nx = &s->v[(i + 1) % n];
            calc_join(st, &vc, p, c, nx, p->dist, c->dist);
            emit_contour(out, &vc, &first);
        }
        vb_push(out, 0, 0, CMD_CLOSE);
        first = 1;
        for (int i = n - 1; i >= 0; i--) {
            const VD *p = &s->v[(i + n - 1) % n], *c = &s->v[i], *nx = &s->v[(i + 1) % n];
            calc_join(st, &vc, nx, c, p, c->dist, p->dist);
            emit_contour(out, &vc, &first);
        }
        vb_push(out, 0, 0, CMD_CLOSE);
    } else {
        calc_cap(st, &vc, &s->v[0], &s->v[1], s->v[0].dist);
        emit_contour(out, &vc, &first);
        for (int i = 1; i < n - 1; i++) {
            calc_join(st, &vc, &s->v[i - 1], &s->v[i], &s->v[i + 1], s->v[i - 1].dist, s->v[i].dist);
            emit_contour(out, &vc, &first);
        }
        calc_cap(st, &vc, &s->v[n - 1], &s->v[n - 2], s->v[n - 2].dist);
        emit_contour(out, &vc, &first);
        for (int i = n - 2; i >= 1; i--) {
            calc_join(st, &vc, &s->v[i + 1], &s->v[i], &s->v[i - 1], s->v[i].dist, s->v[i - 1].dist);
            emit_contour(out, &vc, &first);
        }
        vb_push(out, 0, 0, CMD_CLOSE);
    }
    free(vc.v);
}
/* agg::conv_stroke = conv_adaptor_vcgen<vcgen_stroke>: one generator run per sub-path */
static void conv_stroke(const VBuf *in, const Stroker *st, VBuf *out)
{
    VSeq s = {0, 0, 0};
    double sx = 0, sy = 0;
    int i = 0;
    /* skip to the first vertex command */
    while (i < in->n && !is_vertex(in->v[i].cmd)) i++;
    if (i >= in->n) return;
    sx = in->v[i].x; sy = in->v[i].y;
    i++;
    int last_stop = 0;
    while (!last_stop) {
        s.n = 0;
        int closed = 0;
        vs_modify_last(&s, sx, sy);   /* add_vertex(start, move_to) */
        for (;;) {
            if (i >= in->n) { last_stop = 1; break; }
            const Vtx v = in->v[i++];
            if (is_vertex(v.cmd)) {
                if (v.cmd == CMD_MOVE) { sx = v.x; sy = v.y; break; }
                vs_add(&s, v.x, v.y);
            } else if (v.cmd == CMD_STOP) {
                last_stop = 1;
                break;
            } else {   /* end_poly */
                closed = (v.cmd & 0x40) != 0;
                break;
            }
        }
        stroke_subpath(st, &s, closed, out);
    }
    free(s.v);
}

/* ---- agg::rasterizer_cells_aa ----------------------------------------------------------------- */
typedef struct { int x, y, cover, area; } Cell;
typedef struct {
    Cell *c; int n, cap;
    Cell cur;
} Cells;
static void cells_flush(Cells *cs)
{
    if (cs->cur.area | cs->cur.cover) {
        if (cs->n == cs->cap) {
            cs->cap = cs->cap ? cs->cap * 2 : 1024;
            cs->c = (Cell *)realloc(cs->c, sizeof(Cell) * (size_t)cs->cap);
        }
        cs->c[cs->n++] = cs->cur;
    }
}
static void set_curr_cell(Cells *cs, int x, int y)
{
    if (cs->cur.x != x || cs->cur.y != y) {
        cells_flush(cs);
        cs->cur.x = x; cs->cur.y = y; cs->cur.cover = 0; cs->cur.area = 0;
    }
}
static void render_hline(Cells *cs, int ey, int x1, int y1, int x2, int y2)
{
    int ex1 = x1 >> 8, ex2 = x2 >> 8, fx1 = x1 & 255, fx2 = x2 & 255;
    int delta, p, first, dx, incr, lift, mod, rem;
    if (y1 == y2) { set_curr_cell(cs, ex2, ey); return; }
    if (ex1 == ex2) {
        delta = y2 - y1;
        cs->cur.cover += delta;
        cs->cur.area += (fx1 + fx2) * delta;
        return;
    }
    p = (256 - fx1) * (y2 - y1);
    first = 256;
    incr = 1;
    dx = x2 - x1;
    if (dx < 0) {
        p = fx1 * (y2 - y1);
        first = 0;
        incr = -1;
        dx = -dx;
    }
    delta = p / dx;
    mod = p % dx;
    if (mod < 0) { delta--; mod += dx; }
    cs->cur.cover += delta;
    cs->cur.area += (fx1 + first) * delta;
    ex1 += incr;
    set_curr_cell(cs, ex1, ey);
    y1 += delta;
    if (ex1 != ex2) {
        p = 256 * (y2 - y1 + delta);
        lift = p / dx;
        rem = p % dx;
        if (rem < 0) { lift--; rem += dx; }
        mod -= dx;
        while (ex1 != ex2) {
            delta = lift;
            mod += rem;
            if (mod >= 0) { mod -= dx; delta++; }
            cs->cur.cover += delta;
            cs->cur.area += 256 * delta;
            y1 += delta;
            ex1 += incr;
            set_curr_cell(cs, ex1, ey);
        }
    }
    delta = y2 - y1;
    cs->cur.cover += delta;
    cs->cur.area += (fx2 + 256 - first) * delta;
}
static void cells_line(Cells *cs, int x1, int y1, int x2, int y2)
{
    const long dx_limit = 16384L << 8;
    long dx = (long)x2 - (long)x1;
    if (dx >= dx_limit || dx <= -dx_limit) {
        int cx = (int)(((long)x1 + x2) >> 1), cy = (int)(((long)y1 + y2) >> 1);
        cells_line(cs, x1, y1, cx, cy);
        cells_line(cs, cx, cy, x2, y2);
        return;
    }
    long dy = (long)y2 - (long)y1;
    int ex1 = x1 >> 8, ey1 = y1 >> 8, ey2 = y2 >> 8, fy1 = y1 & 255, fy2 = y2 & 255;
    int x_from, x_to, rem, mod, lift, delta, first, incr;
    long p;
    set_curr_cell(cs, ex1, ey1);
    if (ey1 == ey2) { render_hline(cs, ey1, x1, fy1, x2, fy2); return; }
    incr = 1;
    if (dx == 0) {
        int ex = x1 >> 8, two_fx = (x1 - (ex << 8)) << 1, area;
        first = 256;
        if (dy < 0) { first = 0; incr = -1; }
        delta = first - fy1;
        cs->cur.cover += delta;
        cs->cur.area += two_fx * delta;
        ey1 += incr;
        set_curr_cell(cs, ex, ey1);
        delta = first + first - 256;
        area = two_fx * delta;
        while (ey1 != ey2) {
            cs->cur.cover = delta;
            cs->cur.area = area;
            ey1 += incr;
            set_curr_cell(cs, ex, ey1);
        }
        delta = fy2 - 256 + first;
        cs->cur.cover += delta;
        cs->cur.area += two_fx * delta;
        return;
    }
    p = (256 - fy1) * dx;
    first = 256;
    if (dy < 0) {
        p = fy1 * dx;
        first = 0;
        incr = -1;
        dy = -dy;
    }
    delta = (int)(p / dy);
    mod = (int)(p % dy);
    if (mod < 0) { delta--; mod += (int)dy; }
    x_from = x1 + delta;
    render_hline(cs, ey1, x1, fy1, x_from, first);
    ey1 += incr;
    set_curr_cell(cs, x_from >> 8, ey1);
    if (ey1 != ey2) {
        p = 256 * dx;
        lift = (int)(p / dy);
        rem = (int)(p % dy);
        if (rem < 0) { lift--; rem += (int)dy; }
        mod -= (int)dy;
        while (ey1 != ey2) {
            delta = lift;
            mod += rem;
            if (mod >= 0) { mod -= (int)dy; delta++; }
            x_to = x_from + delta;
            render_hline(cs, ey1, x_from, 256 - first, x_to, first);
            x_from = x_to;
            ey1 += incr;
            set_curr_cell(cs, x_from >> 8, ey1);
        }
    }
    render_hline(cs, ey1, x_from, 256 - first, x2, fy2);
}

/* ---- agg::rasterizer_sl_clip<ras_conv_dbl> ---------------------------------------------------- */
typedef struct {
    Cells *cs;
    double cx1, cy1, cx2, cy2;   /* clip box */
    double x1, y1, start_x, start_y;
    unsigned f1;
    int status;                  /* 0 initial, 1 move_to, 2 line_to, 3 closed */
} Ras;
static int iround(double v) { return (int)((v < 0.0) ? v - 0.5 : v + 0.5); }
static unsigned clip_flags(const Ras *r, double x, double y)
{
    return (unsigned)(x > r->cx2) | ((unsigned)(y > r->cy2) << 1) | ((unsigned)(x < r->cx1) << 2) | ((unsigned)(y < r->cy1) << 3);
}
static unsigned clip_flags_y(const Ras *r, double y) { return ((unsigned)(y > r->cy2) << 1) | ((unsigned)(y < r->cy1) << 3); }
static void line_clip_y(Ras *r, double x1, double y1, double x2, double y2, unsigned f1, unsigned f2)
{
    f1 &= 10;
    f2 &= 10;
    if ((f1 | f2) == 0) {
        cells_line(r->cs, iround(x1 * 256), iround(y1 * 256), iround(x2 * 256), iround(y2 * 256));
    } else {
        if (f1 == f2) return;
        double tx1 = x1, ty1 = y1, tx2 = x2, ty2 = y2;
        if (f1 & 8) { tx1 = x1 + (r->cy1 - y1) * (x2 - x1) / (y2 - y1); ty1 = r->cy1; }
        if (f1 & 2) { tx1 = x1 + (r->cy2 - y1) * (x2 - x1) / (y2 - y1); ty1 = r->cy2; }
        if (f2 & 8) { tx2 = x1 + (r->cy1 - y1) * (x2 - x1) / (y2 - y1); ty2 = r->cy1; }
        if (f2 & 2) { tx2 = x1 + (r->cy2 - y1) * (x2 - x1) / (y2 - y1); ty2 = r->cy2; }
        cells_line(r->cs, iround(tx1 * 256), iround(ty1 * 256), iround(tx2 * 256), iround(ty2 * 256));
    }
}
static void ras_move_to(Ras *r, double x, double y)
{
    r->x1 = x; r->y1 = y;
    r->f1 = clip_flags(r, x, y);
}
static void ras_line_to(Ras *r, double x2, double y2)
{
    unsigned f2 = clip_flags(r, x2, y2);
    if ((r->f1 & 10) == (f2 & 10) && (r->f1 & 10) != 0) {
        r->x1 = x2; r->y1 = y2; r->f1 = f2;
        return;
    }
    double x1 = r->x1, y1 = r->y1, y3, y4;
    unsigned f1 = r->f1, f3, f4;
    switch (((f1 & 5) << 1) | (f2 & 5)) {
    case 0: line_clip_y(r, x1, y1, x2, y2, f1, f2); break;
    case 1:
        y3 = y1 + (r->cx2 - x1) * (y2 - y1) / (x2 - x1); f3 = clip_flags_y(r, y3);
        line_clip_y(r, x1, y1, r->cx2, y3, f1, f3);
        line_clip_y(r, r->cx2, y3, r->cx2, y2, f3, f2);
        break;
    case 2:
        y3 = y1 + (r->cx2 - x1) * (y2 - y1) / (x2 - x1); f3 = clip_flags_y(r, y3);
        line_clip_y(r, r->cx2, y1, r->cx2, y3, f1, f3);
        line_clip_y(r, r->cx2, y3, x2, y2, f3, f2);
        break;
    case 3: line_clip_y(r, r->cx2, y1, r->cx2, y2, f1, f2); break;
    case 4:
        y3 = y1 + (r->cx1 - x1) * (y2 - y1) / (x2 - x1); f3 = clip_flags_y(r, y3);
        line_clip_y(r, x1, y1, r->cx1, y3, f1, f3);
        line_clip_y(r, r->cx1, y3, r->cx1, y2, f3, f2);
        break;
    case 6:
        y3 = y1 + (r->cx2 - x1) * (y2 - y1) / (x2 - x1);
        y4 = y1 + (r->cx1 - x1) * (y2 - y1) / (x2 - x1);
        f3 = clip_flags_y(r, y3); f4 = clip_flags_y(r, y4);
        line_clip_y(r, r->cx2, y1, r->cx2, y3, f1, f3);
        line_clip_y(r, r->cx2, y3, r->cx1, y4, f3, f4);
        line_clip_y(r, r->cx1, y4, r->cx1, y2, f4, f2);
        break;
    case 8:
        y3 = y1 + (r->cx1 - x1) * (y2 - y1) / (x2 - x1); f3 = clip_flags_y(r, y3);
        line_clip_y(r, r->cx1, y1, r->cx1, y3, f1, f3);
        line_clip_y(r, r->cx1, y3, x2, y2, f3, f2);
        break;
    case 9:
        y3 = y1 + (r->cx1 - x1) * (y2 - y1) / (x2 - x1);
        y4 = y1 + (r->cx2 - x1) * (y2 - y1) / (x2 - x1);
        f3 = clip_flags_y(r, y3); f4 = clip_flags_y(r, y4);
        line_clip_y(r, r->cx1, y1, r->cx1, y3, f1, f3);
        line_clip_y(r, r->cx1, y3, r->cx2, y4, f3, f4);
        line_clip_y(r, r->cx2, y4, r->cx2, y2, f4, f2);
        break;
    case 12: line_clip_y(r, r->cx1, y1, r->cx1, y2, f1, f2); break;
    }
    r->f1 = f2;
    r->x1 = x2; r->y1 = y2;
}
static void ras_close(Ras *r)
{
    if (r->status == 2) {
        ras_line_to(r, r->start_x, r->start_y);
        r->status = 3;
    }
}
/* rasterizer_scanline_aa::add_path (auto_close = true) */
static void ras_add_path(Ras *r, const VBuf *p)
{
    for (int i = 0; i < p->n; i++) {
        const Vtx v = p->v[i];
        if (v.cmd == CMD_STOP) break;
        if (v.cmd == CMD_MOVE) {
            ras_close(r);
            r->start_x = v.x; r->start_y = v.y;
            ras_move_to(r, v.x, v.y);
            r->status = 1;
        } else if (is_vertex(v.cmd)) {
            ras_line_to(r, v.x, v.y);
            r->status = 2;
        } else if (v.cmd & 0x40) {
            ras_close(r);
        }
    }
    ras_close(r);   /* sort(): if(m_auto_close) close_polygon() */
    cells_flush(r->cs);
    r->cs->cur.x = r->cs->cur.y = 0x7fffffff;
    r->cs->cur.cover = r->cs->cur.area = 0;
}

/* ---- sweep + fixed_blender_rgba_plain --------------------------------------------------------- */
static int cell_cmp(const void *a, const void *b)
{
    const Cell *p = (const Cell *)a, *q = (const Cell *)b;
    if (p->y != q->y) return p->y < q->y ? -1 : 1;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    return 0;
}
static unsigned calc_alpha(int area)
{
    int cover = area >> 9;
    if (cover < 0) cover = -cover;
    if (cover > 255) cover = 255;
    return (unsigned)cover;
}
static unsigned mult_cover(unsigned a, unsigned b)
{
    unsigned t = a * b + 128;
    return ((t >> 8) + t) >> 8;
}
static void blend_px(uint8_t *p, const uint8_t c[4], unsigned cover)
{
    if (c[3] == 0) return;
    if (c[3] == 255 && cover == 255) { p[0] = c[0]; p[1] = c[1]; p[2] = c[2]; p[3] = 255; return; }
    unsigned alpha = mult_cover(c[3], cover);
    if (alpha == 0) return;
    unsigned a = p[3];
    unsigned r = p[0] * a, g = p[1] * a, b = p[2] * a;
    a = ((alpha + a) << 8) - alpha * a;
    p[3] = (uint8_t)(a >> 8);
    p[0] = (uint8_t)(((((unsigned)c[0] << 8) - r) * alpha + (r << 8)) / a);
    p[1] = (uint8_t)(((((unsigned)c[1] << 8) - g) * alpha + (g << 8)) / a);
    p[2] = (uint8_t)(((((unsigned)c[2] << 8) - b) * alpha + (b << 8)) / a);
}
static void sweep_and_blend(Cells *cs, int W, int H, uint8_t *rgba, const uint8_t color[4])
{
    qsort(cs->c, (size_t)cs->n, sizeof(Cell), cell_cmp);
    int i = 0;
    while (i < cs->n) {
        const int y = cs->c[i].y;
        int cover = 0;
        while (i < cs->n && cs->c[i].y == y) {
            const int x = cs->c[i].x;
            int area = 0;
            while (i < cs->n && cs->c[i].y == y && cs->c[i].x == x) {
                area += cs->c[i].area;
                cover += cs->c[i].cover;
                i++;
            }
            int xs = x;
            if (area) {
                unsigned al = calc_alpha((cover << 9) - area);
                if (al && y >= 0 && y < H && x >= 0 && x < W) blend_px(rgba + ((size_t)y * W + x) * 4, color, al);
                xs = x + 1;
            }
            if (i < cs->n && cs->c[i].y == y && cs->c[i].x > xs) {
                unsigned al = calc_alpha(cover << 9);
                if (al && y >= 0 && y < H)
                    for (int xx = xs < 0 ? 0 : xs; xx < cs->c[i].x && xx < W; xx++) blend_px(rgba + ((size_t)y * W + xx) * 4, color, al);
            }
        }
    }
}

/* ---- public entry points ---------------------------------------------------------------------- */
typedef struct {
    int32_t n_artists;
    const uint32_t *vertex_begin;   /* [n_artists + 1] */
    const double *vertices;         /* x, y pairs */
    const uint8_t *codes;           /* matplotlib Path codes */
    const double *affine;           /* 6 per artist: a, b, c, d, e, f of Affine2D.to_values() */
    const double *linewidth;        /* pixels */
    const uint8_t *cap, *join, *snap;
    const uint8_t *edge_rgba, *face_rgba, *has_face;
    const double *clip;             /* 4 per artist, display coordinates (y up), or NULL */
    const uint8_t *has_clip;        /* or NULL */
} AggFigure;

/* the converter chain of RendererAgg::draw_path up to (and including) conv_curve, in Agg's y-down pixel space */
static void prepare_path(const AggFigure *f, int a, int H, VBuf *curve, double *lw_out)
{
    VBuf t = {0, 0, 0};
    Affine tr = {f->affine[6 * a + 0], f->affine[6 * a + 1], f->affine[6 * a + 2], f->affine[6 * a + 3], f->affine[6 * a + 4],
                 f->affine[6 * a + 5]};
    const Affine flip = {1.0, 0.0, 0.0, -1.0, 0.0, 0.0}, shift = {1.0, 0.0, 0.0, 1.0, 0.0, (double)H};
    affine_multiply(&tr, &flip);
    affine_multiply(&tr, &shift);
    const uint32_t b = f->vertex_begin[a], e = f->vertex_begin[a + 1];
    for (uint32_t i = b; i < e; i++) {
        double x = f->vertices[2 * i], y = f->vertices[2 * i + 1];
        const int cmd = f->codes[i];
        if (is_vertex(cmd)) affine_apply(&tr, &x, &y);
        vb_push(&t, x, y, cmd);
    }
    double lw = f->linewidth[a];
    double snapping_lw = f->edge_rgba[4 * a + 3] == 0 ? 0.0 : lw;
    if (should_snap(&t, f->snap[a], (int)(e - b))) {
        const int r = (int)(snapping_lw + ((snapping_lw >= 0.0) ? 0.5 : -0.5));
        const double sv = (r % 2) ? 0.5 : 0.0;
        for (int i = 0; i < t.n; i++)
            if (is_vertex(t.v[i].cmd)) {
                t.v[i].x = floor(t.v[i].x + 0.5) + sv;
                t.v[i].y = floor(t.v[i].y + 0.5) + sv;
            }
    }
    conv_curve(&t, curve);
    free(t.v);
    *lw_out = lw;
}
static void set_clip(const AggFigure *f, int a, int W, int H, Ras *r)
{
    r->cx1 = 0; r->cy1 = 0; r->cx2 = W; r->cy2 = H;
    if (f->has_clip && f->has_clip[a] && f->clip) {
        const double *c = f->clip + 4 * a;
        int x1 = (int)floor(c[0] + 0.5), y1 = (int)floor((double)H - c[1] + 0.5);
        int x2 = (int)floor(c[2] + 0.5), y2 = (int)floor((double)H - c[3] + 0.5);
        /* set_clipbox: clip_box(max(x1,0), max(H-y1,0), min(x2,W), min(H-y2,H)); agg normalises the box */
        int bx1 = x1 > 0 ? x1 : 0, by1 = y1 > 0 ? y1 : 0, bx2 = x2 < W ? x2 : W, by2 = y2 < H ? y2 : H;
        if (bx1 > bx2) { int t = bx1; bx1 = bx2; bx2 = t; }
        if (by1 > by2) { int t = by1; by1 = by2; by2 = t; }
        r->cx1 = bx1; r->cy1 = by1; r->cx2 = bx2; r->cy2 = by2;
    }
}
static Stroker make_stroker(double lw, int cap, int join)
{
    Stroker s;
    s.width = lw * 0.5;
    if (s.width < 0) { s.width_abs = -s.width; s.width_sign = -1; }
    else { s.width_abs = s.width; s.width_sign = 1; }
    s.width_eps = s.width / 1024.0;
    s.miter_limit = lw;             /* stroke.miter_limit(points_to_pixels(gc.linewidth)) */
    s.inner_miter_limit = 1.01;
    s.approx_scale = 1.0;
    s.cap = cap;
    s.join = join == JOIN_MITER ? LJ_MITER_REVERT : (join == JOIN_ROUND ? LJ_ROUND : LJ_BEVEL);
    return s;
}

/* Renders the artists of one figure, in order, over `rgba` (H * W * 4, caller-initialised: savefig paints the figure
 * patch first, i.e. opaque white). */
void agg_render_figure(const AggFigure *f, int W, int H, uint8_t *rgba)
{
    Cells cs = {0, 0, 0, {0x7fffffff, 0x7fffffff, 0, 0}};
    for (int a = 0; a < f->n_artists; a++) {
        VBuf curve = {0, 0, 0};
        double lw;
        prepare_path(f, a, H, &curve, &lw);
        Ras r;
        memset(&r, 0, sizeof r);
        r.cs = &cs;
        set_clip(f, a, W, H, &r);
        if (f->has_face[a]) {
            cs.n = 0;
            r.status = 0;
            ras_add_path(&r, &curve);
            sweep_and_blend(&cs, W, H, rgba, f->face_rgba + 4 * a);
        }
        if (lw != 0.0) {
            VBuf outline = {0, 0, 0};
            Stroker st = make_stroker(lw, f->cap[a], f->join[a]);
            conv_stroke(&curve, &st, &outline);
            cs.n = 0;
            r.status = 0;
            ras_add_path(&r, &outline);
            sweep_and_blend(&cs, W, H, rgba, f->edge_rgba + 4 * a);
            free(outline.v);
        }
        free(curve.v);
    }
    free(cs.c);
}

/* Debug / test hooks: the stroked outline (or, with stroke == 0, the flattened path) of artist a as (x, y, cmd) triples
 * in Agg pixel space.  Returns the number of vertices; writes at most `cap` of them. */
int agg_debug_outline(const AggFigure *f, int a, int H, int stroke, double *xyc, int cap)
{
    VBuf curve = {0, 0, 0}, outline = {0, 0, 0};
    double lw;
    prepare_path(f, a, H, &curve, &lw);
    const VBuf *src = &curve;
    if (stroke) {
        Stroker st = make_stroker(lw, f->cap[a], f->join[a]);
        conv_stroke(&curve, &st, &outline);
        src = &outline;
    }
    for (int i = 0; i < src->n && i < cap; i++) {
        xyc[3 * i] = src->v[i].x; xyc[3 * i + 1] = src->v[i].y; xyc[3 * i + 2] = src->v[i].cmd;
    }
    int n = src->n;
    free(curve.v);
    free(outline.v);
    return n;
}
