// rig_raster.h -- executes draw records into a band of canvas rows held in (shared) memory.
//
// A band is rows [r0, r1] of one panel, stored as packed 3-byte pixels, row stride 3*W.  One warp
// owns a band: it walks the panel's records in draw order and, for each, runs only the iterations
// that land inside the band, 32 at a time (LANES = 32).  The host build (LANES = 1) runs the same
// code serially so the arithmetic and the band clipping can be checked without a GPU.
//
// Ordering: records are applied one after another with a sync in between, so read-modify-write
// blends see every earlier record, as in OpenCV's sequential drawing.  Inside one record no two
// iterations touch the same pixel (LINEAA: one column/row per iteration), or all writes store the
// same opaque colour (fills), so lanes never conflict.
#pragma once
#include "rig_shapes.h"

namespace rig {

struct Band {
    uint8_t *buf;   // first byte of row r0
    int r0, r1;     // inclusive canvas rows
    int W, H;
};
RIG_HD uint8_t *band_px(const Band &b, int x, int y) { return b.buf + ((y - b.r0) * b.W + x) * 3; }

RIG_HD void put3(uint8_t *p, uint32_t c)
{
    p[0] = (uint8_t)c; p[1] = (uint8_t)(c >> 8); p[2] = (uint8_t)(c >> 16);
}
RIG_HD void blend3(uint8_t *p, uint32_t c, int a)
{
    p[0] = (uint8_t)blend1(p[0], (int)(c & 0xff), a);
    p[1] = (uint8_t)blend1(p[1], (int)((c >> 8) & 0xff), a);
    p[2] = (uint8_t)blend1(p[2], (int)((c >> 16) & 0xff), a);
}
template <int LANES>
RIG_HD void fill_span(const Band &b, int y, int x1, int x2, uint32_t c, int lane)
{
    for (int x = x1 + lane; x <= x2; x += LANES) put3(band_px(b, x, y), c);
}

// per-warp scratch of the rasteriser: edge crossings of one scanline
struct RasterScratch {
    int32_t cross[64];
    int32_t count;
};

// Ops: sync() orders memory between lanes; append(scratch, v) adds a crossing (any order)
template <int LANES, class Ops>
RIG_HD void exec_record(const Rec &r, const Q16 *aux, const uint8_t *filter, const Band &b, int lane,
                        RasterScratch &S, Ops ops)
{
    int y0 = rec_ymin(r), y1 = rec_ymax(r);
    if (y0 < b.r0) y0 = b.r0;
    if (y1 > b.r1) y1 = b.r1;
    if (y0 > y1) return;
    const uint32_t color = r.w[7];
    switch (rec_type(r)) {
    case REC_LINE8: {
        Line8 l = line8_unpack(r);
        int k0, k1;
        line8_krange(l, y0, y1, k0, k1);
        for (int k = k0 + lane; k <= k1; k += LANES) {
            int x, y;
            line8_eval(l, k, x, y);
            put3(band_px(b, x, y), color);
        }
        break;
    }
    case REC_LINEAA: {
        LineAA l = lineaa_unpack(r);
        int k0, k1;
        lineaa_krange(l, y0, y1, k0, k1);
        for (int k = k0 + lane; k <= k1; k += LANES) {
            int major, m0, a[3];
            lineaa_eval(l, k, filter, major, m0, a[0], a[1], a[2]);
            if (l.ymajor) {
                if (major < b.r0 || major > b.r1) continue;
                for (int i = 0; i < 3; i++)
                    if ((unsigned)(m0 + i) < (unsigned)b.W) blend3(band_px(b, m0 + i, major), color, a[i]);
            } else {
                if ((unsigned)major >= (unsigned)b.W) continue;
                for (int i = 0; i < 3; i++) {
                    int y = m0 + i;
                    if (y >= b.r0 && y <= b.r1) blend3(band_px(b, major, y), color, a[i]);
                }
            }
        }
        break;
    }
    case REC_DISC:
    case REC_CRESCENT: {
        int cx = (int32_t)r.w[2], cy = (int32_t)r.w[3], rad = (int32_t)r.w[4];
        bool cres = rec_type(r) == REC_CRESCENT;
        int icx = (int16_t)(r.w[5] & 0xffff), icy = (int16_t)(r.w[5] >> 16), ir = (int32_t)r.w[6];
        for (int y = y0; y <= y1; y++) {
            int dy = y - cy;
            if (dy * dy > rad * rad) continue;
            int hw = isqrt_floor(rad * rad - dy * dy);
            int x1 = cx - hw, x2 = cx + hw;
            if (x1 < 0) x1 = 0;
            if (x2 > b.W - 1) x2 = b.W - 1;
            if (!cres) fill_span<LANES>(b, y, x1, x2, color, lane);
            else {
                int dy2 = (y - icy) * (y - icy);
                for (int x = x1 + lane; x <= x2; x += LANES) {
                    int ddx = x - icx;
                    if (ddx * ddx + dy2 > ir * ir) put3(band_px(b, x, y), color);
                }
            }
        }
        break;
    }
    case REC_CONVEX: {
        const ConvexAux &a = *(const ConvexAux *)(aux + r.w[2]);
        int n0 = r.w[3] & 0xff, n1 = (r.w[3] >> 8) & 0xff;
        for (int y = y0; y <= y1; y++) {
            int x1, x2;
            convex_span(a, n0, n1, y, b.W, x1, x2);
            fill_span<LANES>(b, y, x1, x2, color, lane);
        }
        break;
    }
    case REC_POLYFILL: {
        const PolyEdge *edges = (const PolyEdge *)(aux + r.w[2]);
        int n = (int)r.w[3];
        for (int y = y0; y <= y1; y++) {
            ops.sync();
            if (lane == 0) S.count = 0;
            ops.sync();
            for (int i = lane; i < n; i += LANES) {
                PolyEdge e = edges[i];
                if (polyedge_active(e, y)) ops.append(S, (int32_t)polyedge_x(e, y));
            }
            ops.sync();
            int cnt = S.count;
            if (cnt > 64) cnt = 64;
            if (cnt < 2) continue;
            int lo = S.cross[0], hi = S.cross[0];
            for (int i = 1; i < cnt; i++) {
                int v = S.cross[i];
                if (v < lo) lo = v;
                if (v > hi) hi = v;
            }
            int x1 = (int)(((long long)lo + (XY_ONE - 1)) >> XY_SHIFT), x2 = hi >> XY_SHIFT;
            if (x1 < 0) x1 = 0;
            if (x2 > b.W - 1) x2 = b.W - 1;
            for (int x = x1 + lane; x <= x2; x += LANES) {
                int X = x << XY_SHIFT, L = 0, E = 0;
                for (int i = 0; i < cnt; i++) {
                    int v = S.cross[i];
                    L += v < X;
                    E += v == X;
                }
                if (E || (L & 1)) put3(band_px(b, x, y), color);
            }
        }
        break;
    }
    default: break;
    }
}

// grid lines of a use_grid panel (generator.py:112-122): 1-px LINE_8 black lines, drawn last
template <int LANES>
RIG_HD void draw_grid(const Band &b, int grid_size, int lane)
{
    for (int i = 1; i < grid_size; i++) {
        int x = dround(ddiv((double)(i * b.W), (double)grid_size));
        int y = dround(ddiv((double)(i * b.H), (double)grid_size));
        if ((unsigned)x < (unsigned)b.W)
            for (int yy = b.r0 + lane; yy <= b.r1; yy += LANES) put3(band_px(b, x, yy), 0u);
        if (y >= b.r0 && y <= b.r1) fill_span<LANES>(b, y, 0, b.W - 1, 0u, lane);
    }
}

}  // namespace rig
