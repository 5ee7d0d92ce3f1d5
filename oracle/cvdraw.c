/*
 * oracle/cvdraw.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar C restatement of the OpenCV 4.x 2-D drawing arithmetic that the reference's sequence
 * renderer reaches through cv2 (reference call sites: src/shapes.py:289-302,564-594 and
 * src/generator.py:118-122).  OpenCV itself is a third-party dependency that is NOT vendored under
 * /root/reference and is not even listed in its requirements.txt (unpinned); the version this
 * restatement is pinned against is the opencv-python-headless 4.13.0.92 wheel installed in this
 * image.  The algorithm restated is the published one in modules/imgproc/src/drawing.cpp
 * (clipLine, LineIterator/Line, LineAA, FillConvexPoly, CollectPolyEdges/FillEdgeCollection,
 * ThickLine, PolyLine, ellipse2Poly/EllipseEx, Circle) -- written from knowledge of that algorithm
 * and pinned by differential tests against the installed wheel (tests/test_oracle_vs_cv2.py) and
 * by the known-answer vectors of SURVEY.md section 8(c) (tests/test_oracle_known_answers.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this file.  The product path (reasoning_image_generation_b200/csrc) never does.
 *
 * Canvas: uint8, 3 interleaved channels, row stride = 3*width, no padding.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <limits.h>

#include "cvdraw_tables.h"

#define XY_SHIFT 16
#define XY_ONE (1 << XY_SHIFT)
#define LT_8 8
#define LT_AA 16

typedef struct { int64_t x, y; } P2L;

typedef struct {
    uint8_t *data;
    int w, h;
} Canvas;

/* round-half-to-even on doubles, the cvRound/lrint contract */
static inline int cv_round(double v) { return (int)lrint(v); }

static inline void put_px(const Canvas *c, int x, int y, const uint8_t *col)
{
    uint8_t *p = c->data + ((size_t)y * c->w + x) * 3;
    p[0] = col[0]; p[1] = col[1]; p[2] = col[2];
}

static inline void hline(const Canvas *c, int y, int x1, int x2, const uint8_t *col)
{
    for (int x = x1; x <= x2; x++) put_px(c, x, y, col);
}

/* ---- clipLine on 64-bit points (used both in pixel units and in 16.16 units) ------------- */
static int clip_line(int64_t width, int64_t height, P2L *p1, P2L *p2)
{
    int64_t right = width - 1, bottom = height - 1;
    if (width <= 0 || height <= 0) return 0;
    int64_t x1 = p1->x, y1 = p1->y, x2 = p2->x, y2 = p2->y;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;

    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        int64_t a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (int64_t)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (int64_t)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (int64_t)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (int64_t)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
        p1->x = x1; p1->y = y1; p2->x = x2; p2->y = y2;
    }
    return (c1 | c2) == 0;
}

/* ---- Line(): 8-connected Bresenham walk, always started from the left end ---------------- */
static void line8(const Canvas *c, P2L a, P2L b, const uint8_t *col)
{
    /* LineIterator works on 32-bit points */
    P2L p1 = { (int)a.x, (int)a.y }, p2 = { (int)b.x, (int)b.y };
    if ((uint64_t)p1.x >= (uint64_t)c->w || (uint64_t)p2.x >= (uint64_t)c->w ||
        (uint64_t)p1.y >= (uint64_t)c->h || (uint64_t)p2.y >= (uint64_t)c->h) {
        if (!clip_line(c->w, c->h, &p1, &p2)) return;
    }
    int dx = (int)(p2.x - p1.x), dy = (int)(p2.y - p1.y);
    int sx = 1, sy = 1;
    int x = (int)p1.x, y = (int)p1.y;
    if (dx < 0) { dx = -dx; dy = -dy; x = (int)p2.x; y = (int)p2.y; }
    if (dy < 0) { dy = -dy; sy = -1; }
    int vert = dy > dx;
    if (vert) { int t = dx; dx = dy; dy = t; }
    int err = dx - (dy + dy);
    int plus = dx + dx, minus = -(dy + dy);
    int count = dx + 1;
    for (int i = 0; i < count; i++) {
        put_px(c, x, y, col);
        int m = err < 0;
        err += minus + (m ? plus : 0);
        if (vert) { y += sy; if (m) x += sx; }
        else      { x += sx; if (m) y += sy; }
    }
}

/* ---- LineAA(): 3-tap filtered line on 16.16 end points ------------------------------------ */
static inline void blend_px(const Canvas *c, int x, int y, const uint8_t *col, int a)
{
    uint8_t *p = c->data + ((size_t)y * c->w + x) * 3;
    for (int k = 0; k < 3; k++) {
        int v = p[k];
        v += ((col[k] - v) * a + 127) >> 8;
        v += ((col[k] - v) * a + 127) >> 8;
        p[k] = (uint8_t)v;
    }
}

static void line_aa(const Canvas *c, P2L pt1, P2L pt2, const uint8_t *col)
{
    int64_t dx, dy, ax, ay, i, j, x_step, y_step;
    int ecount, scount = 0, slope;
    int ep_table[9];

    if (!clip_line((int64_t)c->w << XY_SHIFT, (int64_t)c->h << XY_SHIFT, &pt1, &pt2)) return;

    dx = pt2.x - pt1.x;
    dy = pt2.y - pt1.y;
    j = dx < 0 ? -1 : 0;
    ax = (dx ^ j) - j;
    i = dy < 0 ? -1 : 0;
    ay = (dy ^ i) - i;

    if (ax > ay) {
        dy = (dy ^ j) - j;
        if (j) { P2L t = pt1; pt1 = pt2; pt2 = t; }
        x_step = XY_ONE;
        y_step = (int64_t)((uint64_t)dy << XY_SHIFT) / (ax | 1);
        pt2.x += XY_ONE;
        ecount = (int)((pt2.x >> XY_SHIFT) - (pt1.x >> XY_SHIFT));
        j = -(pt1.x & (XY_ONE - 1));
        pt1.y += ((y_step * j) >> XY_SHIFT) + (XY_ONE >> 1);
        slope = (int)((y_step >> (XY_SHIFT - 5)) & 0x3f);
        slope ^= (y_step < 0 ? 0x3f : 0);
        i = (pt1.x >> (XY_SHIFT - 7)) & 0x78;
        j = (pt2.x >> (XY_SHIFT - 7)) & 0x78;
    } else {
        dx = (dx ^ i) - i;
        if (i) { P2L t = pt1; pt1 = pt2; pt2 = t; }
        x_step = (int64_t)((uint64_t)dx << XY_SHIFT) / (ay | 1);
        y_step = XY_ONE;
        pt2.y += XY_ONE;
        ecount = (int)((pt2.y >> XY_SHIFT) - (pt1.y >> XY_SHIFT));
        j = -(pt1.y & (XY_ONE - 1));
        pt1.x += ((x_step * j) >> XY_SHIFT) + (XY_ONE >> 1);
        slope = (int)((x_step >> (XY_SHIFT - 5)) & 0x3f);
        slope ^= (x_step < 0 ? 0x3f : 0);
        i = (pt1.y >> (XY_SHIFT - 7)) & 0x78;
        j = (pt2.y >> (XY_SHIFT - 7)) & 0x78;
    }

    slope = (slope & 0x20) ? 0x100 : CVD_SLOPE_CORR[slope];

    {
        int t0 = slope << 7;
        int t1 = ((0x78 - (int)i) | 4) * slope;
        int t2 = ((int)j | 4) * slope;
        ep_table[0] = 0;
        ep_table[8] = slope;
        ep_table[1] = ep_table[3] = (((((int)j - (int)i) & 0x78) | 4) * slope >> 8) & 0x1ff;
        ep_table[2] = (t1 >> 8) & 0x1ff;
        ep_table[4] = (((((int)j - (int)i) + 0x80) | 4) * slope >> 8) & 0x1ff;
        ep_table[5] = ((t1 + t0) >> 8) & 0x1ff;
        ep_table[6] = (t2 >> 8) & 0x1ff;
        ep_table[7] = ((t2 + t0) >> 8) & 0x1ff;
    }

    if (ax > ay) {
        int x = (int)(pt1.x >> XY_SHIFT);
        for (; ecount >= 0; x++, pt1.y += y_step, scount++, ecount--) {
            if ((unsigned)x >= (unsigned)c->w) continue;
            int y = (int)((pt1.y >> XY_SHIFT) - 1);
            int ep_corr = ep_table[(((scount >= 2) + 1) & (scount | 2)) * 3 +
                                   (((ecount >= 2) + 1) & (ecount | 2))];
            int a, dist = (int)((pt1.y >> (XY_SHIFT - 5)) & 31);
            a = (ep_corr * CVD_FILTER[dist + 32] >> 8) & 0xff;
            if ((unsigned)y < (unsigned)c->h) blend_px(c, x, y, col, a);
            a = (ep_corr * CVD_FILTER[dist] >> 8) & 0xff;
            if ((unsigned)(y + 1) < (unsigned)c->h) blend_px(c, x, y + 1, col, a);
            a = (ep_corr * CVD_FILTER[63 - dist] >> 8) & 0xff;
            if ((unsigned)(y + 2) < (unsigned)c->h) blend_px(c, x, y + 2, col, a);
        }
    } else {
        int y = (int)(pt1.y >> XY_SHIFT);
        for (; ecount >= 0; y++, pt1.x += x_step, scount++, ecount--) {
            if ((unsigned)y >= (unsigned)c->h) continue;
            int x = (int)((pt1.x >> XY_SHIFT) - 1);
            int ep_corr = ep_table[(((scount >= 2) + 1) & (scount | 2)) * 3 +
                                   (((ecount >= 2) + 1) & (ecount | 2))];
            int a, dist = (int)((pt1.x >> (XY_SHIFT - 5)) & 31);
            a = (ep_corr * CVD_FILTER[dist + 32] >> 8) & 0xff;
            if ((unsigned)x < (unsigned)c->w) blend_px(c, x, y, col, a);
            a = (ep_corr * CVD_FILTER[dist] >> 8) & 0xff;
            if ((unsigned)(x + 1) < (unsigned)c->w) blend_px(c, x + 1, y, col, a);
            a = (ep_corr * CVD_FILTER[63 - dist] >> 8) & 0xff;
            if ((unsigned)(x + 2) < (unsigned)c->w) blend_px(c, x + 2, y, col, a);
        }
    }
}

/* ---- FillConvexPoly(): two-edge walk; AA mode draws the outline with LineAA first -------- */
static void fill_convex_poly(const Canvas *c, const P2L *v, int npts, const uint8_t *col,
                             int line_type, int shift)
{
    struct { int idx, di; int64_t x, dx; int ye; } edge[2];
    int delta = 1 << shift >> 1;
    int i, y, imin = 0;
    int edges = npts;
    int64_t xmin, xmax, ymin, ymax;
    int delta1, delta2;
    P2L p0;

    if (line_type < LT_AA) delta1 = delta2 = XY_ONE >> 1;
    else { delta1 = XY_ONE - 1; delta2 = 0; }

    p0 = v[npts - 1];
    p0.x <<= XY_SHIFT - shift;
    p0.y <<= XY_SHIFT - shift;

    xmin = xmax = v[0].x;
    ymin = ymax = v[0].y;

    for (i = 0; i < npts; i++) {
        P2L p = v[i];
        if (p.y < ymin) { ymin = p.y; imin = i; }
        if (p.y > ymax) ymax = p.y;
        if (p.x > xmax) xmax = p.x;
        if (p.x < xmin) xmin = p.x;
        p.x <<= XY_SHIFT - shift;
        p.y <<= XY_SHIFT - shift;
        if (line_type <= 8) {
            if (shift == 0) {
                P2L a = { p0.x >> XY_SHIFT, p0.y >> XY_SHIFT }, b = { p.x >> XY_SHIFT, p.y >> XY_SHIFT };
                line8(c, a, b, col);
            } else {
                /* Line2 (sub-pixel LINE_8) is not on the reference path */
                abort();
            }
        } else
            line_aa(c, p0, p, col);
        p0 = p;
    }

    xmin = (xmin + delta) >> shift;
    xmax = (xmax + delta) >> shift;
    ymin = (ymin + delta) >> shift;
    ymax = (ymax + delta) >> shift;

    if (npts < 3 || (int)xmax < 0 || (int)ymax < 0 || (int)xmin >= c->w || (int)ymin >= c->h)
        return;

    if (ymax > c->h - 1) ymax = c->h - 1;
    edge[0].idx = edge[1].idx = imin;
    edge[0].ye = edge[1].ye = y = (int)ymin;
    edge[0].di = 1;
    edge[1].di = npts - 1;
    edge[0].x = edge[1].x = -XY_ONE;
    edge[0].dx = edge[1].dx = 0;

    do {
        if (line_type < LT_AA || y < (int)ymax || y == (int)ymin) {
            for (i = 0; i < 2; i++) {
                if (y >= edge[i].ye) {
                    int idx0 = edge[i].idx, di = edge[i].di;
                    int idx = idx0 + di;
                    if (idx >= npts) idx -= npts;
                    int ty = 0;
                    for (; edges-- > 0;) {
                        ty = (int)((v[idx].y + delta) >> shift);
                        if (ty > y) {
                            int64_t xs = v[idx0].x, xe = v[idx].x;
                            if (shift != XY_SHIFT) {
                                xs <<= XY_SHIFT - shift;
                                xe <<= XY_SHIFT - shift;
                            }
                            edge[i].ye = ty;
                            edge[i].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
                            edge[i].x = xs;
                            edge[i].idx = idx;
                            break;
                        }
                        idx0 = idx;
                        idx += di;
                        if (idx >= npts) idx -= npts;
                    }
                }
            }
        }

        if (edges < 0) break;

        if (y >= 0) {
            int left = 0, right = 1;
            if (edge[0].x > edge[1].x) { left = 1; right = 0; }
            int xx1 = (int)((edge[left].x + delta1) >> XY_SHIFT);
            int xx2 = (int)((edge[right].x + delta2) >> XY_SHIFT);
            if (xx2 >= 0 && xx1 < c->w) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= c->w) xx2 = c->w - 1;
                hline(c, y, xx1, xx2, col);
            }
        }
        edge[0].x += edge[0].dx;
        edge[1].x += edge[1].dx;
    } while (++y <= (int)ymax);
}

/* ---- fillPoly(): edge collection + even-odd scanline fill -------------------------------- */
typedef struct PolyEdge {
    int y0, y1;
    int64_t x, dx;
    struct PolyEdge *next;
} PolyEdge;

static int collect_poly_edges(const Canvas *c, const P2L *v, int count, PolyEdge *edges, int ne,
                              const uint8_t *col, int line_type, int shift)
{
    int i, delta = ((1 << shift) >> 1);
    P2L pt0 = v[count - 1], pt1;
    pt0.x = pt0.x << (XY_SHIFT - shift);
    pt0.y = (pt0.y + delta) >> shift;

    for (i = 0; i < count; i++, pt0 = pt1) {
        P2L t0, t1;
        PolyEdge edge;
        pt1 = v[i];
        pt1.x = pt1.x << (XY_SHIFT - shift);
        pt1.y = (pt1.y + delta) >> shift;
        P2L pt0c = pt0, pt1c = pt1;

        if (line_type < LT_AA) {
            t0.y = pt0.y; t1.y = pt1.y;
            t0.x = (pt0.x + (XY_ONE >> 1)) >> XY_SHIFT;
            t1.x = (pt1.x + (XY_ONE >> 1)) >> XY_SHIFT;
            line8(c, t0, t1, col);
            /* clipped end points make the scan edge agree with the drawn boundary */
            if ((uint64_t)t0.x >= (uint64_t)c->w || (uint64_t)t1.x >= (uint64_t)c->w ||
                (uint64_t)t0.y >= (uint64_t)c->h || (uint64_t)t1.y >= (uint64_t)c->h) {
                P2L a = { (int)t0.x, (int)t0.y }, b = { (int)t1.x, (int)t1.y };
                clip_line(c->w, c->h, &a, &b);
                /* x is always taken from the clipped segment; y only when it still spans rows
                 * (probed on the 4.13.0 wheel: a segment that clips to one point becomes a
                 * vertical edge on the border column over the original row range) */
                if (a.y != b.y) { pt0c.y = a.y; pt1c.y = b.y; }
                pt0c.x = a.x << XY_SHIFT;
                pt1c.x = b.x << XY_SHIFT;
            }
        } else {
            t0.x = pt0.x; t1.x = pt1.x;
            t0.y = pt0.y << XY_SHIFT;
            t1.y = pt1.y << XY_SHIFT;
            line_aa(c, t0, t1, col);
        }

        if (pt0.y == pt1.y) continue;

        edge.dx = (pt1c.x - pt0c.x) / (pt1c.y - pt0c.y);
        if (pt0.y < pt1.y) {
            edge.y0 = (int)pt0.y; edge.y1 = (int)pt1.y;
            edge.x = pt0c.x + (pt0.y - pt0c.y) * edge.dx; /* back-extrapolate a clipped start */
        } else {
            edge.y0 = (int)pt1.y; edge.y1 = (int)pt0.y;
            edge.x = pt1c.x + (pt1.y - pt1c.y) * edge.dx;
        }
        edge.next = 0;
        edges[ne++] = edge;
    }
    return ne;
}

static int cmp_edges(const void *pa, const void *pb)
{
    const PolyEdge *e1 = pa, *e2 = pb;
    if (e1->y0 != e2->y0) return e1->y0 < e2->y0 ? -1 : 1;
    if (e1->x != e2->x) return e1->x < e2->x ? -1 : 1;
    if (e1->dx != e2->dx) return e1->dx < e2->dx ? -1 : 1;
    return 0;
}

static void fill_edge_collection(const Canvas *c, PolyEdge *edges, int total, const uint8_t *col,
                                 int line_type)
{
    PolyEdge tmp;
    int i, y;
    PolyEdge *e;
    int y_max = INT_MIN, y_min = INT_MAX;
    int64_t x_max = -1, x_min = INT64_MAX;
    int delta = XY_ONE - 1; /* left end rounds up, right end rounds down, for every line type */

    if (total < 2) return;

    for (i = 0; i < total; i++) {
        PolyEdge *e1 = &edges[i];
        int64_t x1 = e1->x + (int64_t)(e1->y1 - e1->y0) * e1->dx;
        if (e1->y0 < y_min) y_min = e1->y0;
        if (e1->y1 > y_max) y_max = e1->y1;
        if (e1->x < x_min) x_min = e1->x;
        if (e1->x > x_max) x_max = e1->x;
        if (x1 < x_min) x_min = x1;
        if (x1 > x_max) x_max = x1;
    }
    if (y_max < 0 || y_min >= c->h || x_max < 0 || x_min >= ((int64_t)c->w << XY_SHIFT)) return;

    qsort(edges, total, sizeof(PolyEdge), cmp_edges);

    tmp.y0 = INT_MAX;
    edges[total] = tmp; /* sentinel; caller reserved the slot */
    i = 0;
    tmp.next = 0;
    e = &edges[i];
    if (y_max > c->h) y_max = c->h;

    for (y = e->y0; y < y_max; y++) {
        PolyEdge *last, *prelast, *keep_prelast;
        int draw = 0;
        int clipline = y < 0;

        prelast = &tmp;
        last = tmp.next;
        while (last || e->y0 == y) {
            if (last && last->y1 == y) {
                prelast->next = last->next;
                last = last->next;
                continue;
            }
            keep_prelast = prelast;
            if (last && (e->y0 > y || last->x < e->x)) {
                prelast = last;
                last = last->next;
            } else if (i < total) {
                prelast->next = e;
                e->next = last;
                prelast = e;
                e = &edges[++i];
            } else
                break;

            if (draw) {
                if (!clipline) {
                    int x1, x2;
                    if (keep_prelast->x > prelast->x) {
                        x1 = (int)((prelast->x + delta) >> XY_SHIFT);
                        x2 = (int)(keep_prelast->x >> XY_SHIFT);
                    } else {
                        x1 = (int)((keep_prelast->x + delta) >> XY_SHIFT);
                        x2 = (int)(prelast->x >> XY_SHIFT);
                    }
                    if (x1 < c->w && x2 >= 0) {
                        if (x1 < 0) x1 = 0;
                        if (x2 >= c->w) x2 = c->w - 1;
                        hline(c, y, x1, x2, col);
                    }
                }
                keep_prelast->x += keep_prelast->dx;
                prelast->x += prelast->dx;
            }
            draw ^= 1;
        }

        /* bubble-sort the active list by x for the next scanline */
        keep_prelast = 0;
        do {
            prelast = &tmp;
            last = tmp.next;
            PolyEdge *last_exchange = 0;
            while (last != keep_prelast && last->next != 0) {
                PolyEdge *te = last->next;
                if (last->x > te->x) {
                    prelast->next = te;
                    last->next = te->next;
                    te->next = last;
                    prelast = te;
                    last_exchange = prelast;
                } else {
                    prelast = last;
                    last = te;
                }
            }
            if (last_exchange == NULL) break;
            keep_prelast = last_exchange;
        } while (keep_prelast != tmp.next && keep_prelast != &tmp);
    }
}

/* ---- ellipse2Poly / EllipseEx / Circle / ThickLine / PolyLine ---------------------------- */
static int ellipse2poly(double cx, double cy, double aw, double ah, int angle, int arc_start,
                        int arc_end, int delta, double *ox, double *oy)
{
    float alpha, beta;
    int i, n = 0;
    while (angle < 0) angle += 360;
    while (angle > 360) angle -= 360;
    if (arc_start > arc_end) { i = arc_start; arc_start = arc_end; arc_end = i; }
    while (arc_start < 0) { arc_start += 360; arc_end += 360; }
    while (arc_end > 360) { arc_end -= 360; arc_start -= 360; }
    if (arc_end - arc_start > 360) { arc_start = 0; arc_end = 360; }
    beta = CVD_SIN[angle];
    alpha = CVD_SIN[450 - angle];
    for (i = arc_start; i < arc_end + delta; i += delta) {
        double x, y;
        int a = i;
        if (a > arc_end) a = arc_end;
        if (a < 0) a += 360;
        x = aw * CVD_SIN[450 - a];
        y = ah * CVD_SIN[a];
        ox[n] = cx + x * alpha - y * beta;
        oy[n] = cy + x * beta + y * alpha;
        n++;
    }
    if (n == 1) { ox[1] = ox[0] = cx; oy[1] = oy[0] = cy; n = 2; }
    return n;
}

static void thick_line(const Canvas *c, P2L p0, P2L p1, const uint8_t *col, int thickness,
                       int line_type, int flags, int shift);

static void poly_line(const Canvas *c, const P2L *v, int count, int is_closed, const uint8_t *col,
                      int thickness, int line_type, int shift)
{
    if (!v || count <= 0) return;
    int i = is_closed ? count - 1 : 0;
    int flags = 2 + !is_closed;
    P2L p0 = v[i];
    for (i = !is_closed; i < count; i++) {
        P2L p = v[i];
        thick_line(c, p0, p, col, thickness, line_type, flags, shift);
        p0 = p;
        flags = 2;
    }
}

static void circle_fill8(const Canvas *c, int cx, int cy, int radius, const uint8_t *col);

static void ellipse_ex(const Canvas *c, P2L center, int64_t aw, int64_t ah, int angle, int arc_start,
                       int arc_end, const uint8_t *col, int thickness, int line_type)
{
    if (aw < 0) aw = -aw;
    if (ah < 0) ah = -ah;
    int delta = (int)(((aw > ah ? aw : ah) + (XY_ONE >> 1)) >> XY_SHIFT);
    delta = delta < 3 ? 90 : delta < 10 ? 30 : delta < 15 ? 18 : 5;

    double px[400], py[400];
    int n = ellipse2poly((double)center.x, (double)center.y, (double)aw, (double)ah, angle,
                         arc_start, arc_end, delta, px, py);
    P2L v[402];
    int nv = 0;
    P2L prev = { -1, -1 }; /* 0xFFFF...: matches the all-ones initial value */
    for (int i = 0; i < n; i++) {
        P2L pt;
        pt.x = (int64_t)cv_round(px[i] / (double)XY_ONE) << XY_SHIFT;
        pt.y = (int64_t)cv_round(py[i] / (double)XY_ONE) << XY_SHIFT;
        pt.x += cv_round(px[i] - (double)pt.x);
        pt.y += cv_round(py[i] - (double)pt.y);
        if (pt.x != prev.x || pt.y != prev.y) {
            v[nv++] = pt;
            prev = pt;
        }
    }
    if (nv <= 1) { v[0] = v[1] = center; nv = 2; }

    if (thickness >= 0)
        poly_line(c, v, nv, 0, col, thickness, line_type, XY_SHIFT);
    else if (arc_end - arc_start >= 360)
        fill_convex_poly(c, v, nv, col, line_type, XY_SHIFT);
    else
        abort(); /* partial filled arcs are not on the reference path */
}

static void thick_line(const Canvas *c, P2L p0, P2L p1, const uint8_t *col, int thickness,
                       int line_type, int flags, int shift)
{
    static const double INV_XY_ONE = 1. / XY_ONE;
    if (shift == 0 && thickness > 1) {
        /* probed on the 4.13.0 wheel (not in older drawing.cpp): integer end points of a thick
         * line are first clipped to the canvas grown by `thickness` px on every side; a segment
         * entirely outside that rectangle draws nothing, caps included */
        P2L a = { p0.x + thickness, p0.y + thickness }, b = { p1.x + thickness, p1.y + thickness };
        if (!clip_line((int64_t)c->w + 2 * thickness, (int64_t)c->h + 2 * thickness, &a, &b)) return;
        p0.x = a.x - thickness; p0.y = a.y - thickness;
        p1.x = b.x - thickness; p1.y = b.y - thickness;
    }
    p0.x <<= XY_SHIFT - shift;
    p0.y <<= XY_SHIFT - shift;
    p1.x <<= XY_SHIFT - shift;
    p1.y <<= XY_SHIFT - shift;

    if (thickness <= 1) {
        if (line_type < LT_AA) {
            if (line_type == 1 || line_type == 4 || shift == 0) {
                p0.x = (p0.x + (XY_ONE >> 1)) >> XY_SHIFT;
                p0.y = (p0.y + (XY_ONE >> 1)) >> XY_SHIFT;
                p1.x = (p1.x + (XY_ONE >> 1)) >> XY_SHIFT;
                p1.y = (p1.y + (XY_ONE >> 1)) >> XY_SHIFT;
                line8(c, p0, p1, col);
            } else
                abort(); /* Line2 not on the reference path */
        } else
            line_aa(c, p0, p1, col);
    } else {
        P2L pt[4], dp = { 0, 0 };
        double dx = (p0.x - p1.x) * INV_XY_ONE, dy = (p1.y - p0.y) * INV_XY_ONE;
        double r = dx * dx + dy * dy;
        int i, odd = thickness & 1;
        thickness <<= XY_SHIFT - 1;

        if (fabs(r) > DBL_EPSILON) {
            r = (thickness + odd * XY_ONE * 0.5) / sqrt(r);
            dp.x = cv_round(dy * r);
            dp.y = cv_round(dx * r);
            pt[0].x = p0.x + dp.x; pt[0].y = p0.y + dp.y;
            pt[1].x = p0.x - dp.x; pt[1].y = p0.y - dp.y;
            pt[2].x = p1.x - dp.x; pt[2].y = p1.y - dp.y;
            pt[3].x = p1.x + dp.x; pt[3].y = p1.y + dp.y;
            fill_convex_poly(c, pt, 4, col, line_type, XY_SHIFT);
        }

        for (i = 0; i < 2; i++) {
            if (flags & (i + 1)) {
                if (line_type < LT_AA) {
                    int ccx = (int)((p0.x + (XY_ONE >> 1)) >> XY_SHIFT);
                    int ccy = (int)((p0.y + (XY_ONE >> 1)) >> XY_SHIFT);
                    circle_fill8(c, ccx, ccy, (thickness + (XY_ONE >> 1)) >> XY_SHIFT, col);
                } else
                    ellipse_ex(c, p0, thickness, thickness, 0, 0, 360, col, -1, line_type);
            }
            p0 = p1;
        }
    }
}

/* Circle(..., fill=1): midpoint circle, horizontal spans */
static void circle_fill8(const Canvas *c, int cx, int cy, int radius, const uint8_t *col)
{
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    int W = c->w, H = c->h;
    while (dx >= dy) {
        int mask;
        int y11 = cy - dy, y12 = cy + dy, y21 = cy - dx, y22 = cy + dx;
        int x11 = cx - dx, x12 = cx + dx, x21 = cx - dy, x22 = cx + dy;
        if (x11 < W && x12 >= 0 && y21 < H && y22 >= 0) {
            int a1 = x11 < 0 ? 0 : x11, a2 = x12 > W - 1 ? W - 1 : x12;
            if ((unsigned)y11 < (unsigned)H) hline(c, y11, a1, a2, col);
            if ((unsigned)y12 < (unsigned)H) hline(c, y12, a1, a2, col);
            if (x21 < W && x22 >= 0) {
                int b1 = x21 < 0 ? 0 : x21, b2 = x22 > W - 1 ? W - 1 : x22;
                if ((unsigned)y21 < (unsigned)H) hline(c, y21, b1, b2, col);
                if ((unsigned)y22 < (unsigned)H) hline(c, y22, b1, b2, col);
            }
        }
        dy++;
        err += plus;
        plus += 2;
        mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

/* =========================== exported, cv2-shaped entry points ============================ */

/* cv2.fillPoly(img, [pts], color)  -- lineType=LINE_8, shift=0 */
void cvd_fill_poly(uint8_t *img, int w, int h, const int32_t *pts, int n, const uint8_t *bgr)
{
    if (n <= 0) return;
    Canvas c = { img, w, h };
    P2L *v = malloc(sizeof(P2L) * n);
    PolyEdge *edges = malloc(sizeof(PolyEdge) * (n + 2));
    for (int i = 0; i < n; i++) { v[i].x = pts[2 * i]; v[i].y = pts[2 * i + 1]; }
    int ne = collect_poly_edges(&c, v, n, edges, 0, bgr, LT_8, 0);
    fill_edge_collection(&c, edges, ne, bgr, LT_8);
    free(edges);
    free(v);
}

/* cv2.polylines(img, [pts], isClosed, color, thickness, lineType) -- shift=0 */
void cvd_polylines(uint8_t *img, int w, int h, const int32_t *pts, int n, int closed,
                   const uint8_t *bgr, int thickness, int line_type)
{
    if (n <= 0) return;
    Canvas c = { img, w, h };
    P2L *v = malloc(sizeof(P2L) * n);
    for (int i = 0; i < n; i++) { v[i].x = pts[2 * i]; v[i].y = pts[2 * i + 1]; }
    poly_line(&c, v, n, closed, bgr, thickness, line_type, 0);
    free(v);
}

/* cv2.line(img, p1, p2, color, thickness, lineType, shift) */
void cvd_line(uint8_t *img, int w, int h, int x1, int y1, int x2, int y2, const uint8_t *bgr,
              int thickness, int line_type, int shift)
{
    Canvas c = { img, w, h };
    P2L a = { x1, y1 }, b = { x2, y2 };
    thick_line(&c, a, b, bgr, thickness, line_type, 3, shift);
}

/* cv2.circle(img, center, radius, color, thickness, lineType) -- shift=0 */
void cvd_circle(uint8_t *img, int w, int h, int cx, int cy, int radius, const uint8_t *bgr,
                int thickness, int line_type)
{
    Canvas c = { img, w, h };
    if (thickness > 1 || line_type != LT_8) {
        P2L ctr = { (int64_t)cx << XY_SHIFT, (int64_t)cy << XY_SHIFT };
        int64_t r = (int64_t)radius << XY_SHIFT;
        ellipse_ex(&c, ctr, r, r, 0, 0, 360, bgr, thickness, line_type);
    } else if (thickness < 0)
        circle_fill8(&c, cx, cy, radius, bgr);
    else
        abort(); /* 1-px LINE_8 circle outline is not on the reference path */
}

/* cv2.fillConvexPoly(img, pts, color, lineType, shift) -- exposed for differential tests */
void cvd_fill_convex_poly(uint8_t *img, int w, int h, const int32_t *pts, int n, const uint8_t *bgr,
                          int line_type, int shift)
{
    Canvas c = { img, w, h };
    P2L *v = malloc(sizeof(P2L) * n);
    for (int i = 0; i < n; i++) { v[i].x = pts[2 * i]; v[i].y = pts[2 * i + 1]; }
    fill_convex_poly(&c, v, n, bgr, line_type, shift);
    free(v);
}
