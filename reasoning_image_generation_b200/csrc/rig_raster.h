// rig_raster.h -- executes draw records into a band of canvas rows held in (shared) memory.
//
// A band is rows [r0, r1] of one panel, stored as packed 3-byte pixels, row stride 3*W.  One warp
// owns a band: it walks the panel's records in draw order and, for each, runs only the iterations
// that land inside the band, 32 at a time (LANES = 32).  The host build (LANES = 1) runs the same
// code serially so the arithmetic and the band clipping can be checked without a GPU.
//
// Ordering: records are applied one after another with a sync in between, so read-modify-write
// blends see every earlier record, as in OpenCV's sequential drawing.  Inside one record no two
// lanes touch the same pixel: LINEAA gives each lane its own column/row, MINIPOLY gives each lane its
// own pixel (and applies that pixel's edge blends in order), fills store one opaque colour.
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
RIG_HD uint32_t get3(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16); }
// both OpenCV blend steps on the three channels of a packed pixel
RIG_HD uint32_t blend_px(uint32_t v, uint32_t c, int a)
{
    if (c == 0u) {
        // black ink (every stroke of the reference): ((0 - v) * a + 127) >> 8 == -((v * a + 128) >> 8), and the
        // three channels fit 16-bit lanes (v * a <= 65025), so both blend steps run on packed words
        uint32_t lo = v & 0x00ff00ffu, hi = (v >> 8) & 0xffu;
        lo -= ((lo * (uint32_t)a + 0x00800080u) >> 8) & 0x00ff00ffu;
        hi -= (hi * (uint32_t)a + 0x80u) >> 8;
        lo -= ((lo * (uint32_t)a + 0x00800080u) >> 8) & 0x00ff00ffu;
        hi -= (hi * (uint32_t)a + 0x80u) >> 8;
        return lo | (hi << 8);
    }
    int v0 = blend1((int)(v & 0xff), (int)(c & 0xff), a);
    int v1 = blend1((int)((v >> 8) & 0xff), (int)((c >> 8) & 0xff), a);
    int v2 = blend1((int)((v >> 16) & 0xff), (int)((c >> 16) & 0xff), a);
    return (uint32_t)v0 | ((uint32_t)v1 << 8) | ((uint32_t)v2 << 16);
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

// Ops: sync() orders memory between lanes; append(scratch, v) adds a crossing (any order);
//      slot(k) returns the k-th 32-byte slot of the current record (k = 0: the record itself),
//      slot_part(k, lo, hi) the same with only words lo..hi filled in;
//      bcast(v, src) returns lane src's v (host: v).
template <int LANES, class Ops>
RIG_HD void exec_record(const Rec &r, const Q16 *aux, const uint8_t *filter, const Band &b, int lane,
                        RasterScratch &S, Ops &ops)
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
            if (y >= y0 && y <= y1) put3(band_px(b, x, y), color);
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
            // the three taps are distinct pixels: read all, blend all, write all
            uint8_t *p[3];
            bool on[3];
            uint32_t v[3];
            if (l.ymajor) {
                bool row = major >= b.r0 && major <= b.r1;
                for (int i = 0; i < 3; i++) {
                    on[i] = row && (unsigned)(m0 + i) < (unsigned)b.W;
                    p[i] = band_px(b, on[i] ? m0 + i : 0, on[i] ? major : b.r0);
                }
            } else {
                bool col = (unsigned)major < (unsigned)b.W;
                for (int i = 0; i < 3; i++) {
                    on[i] = col && m0 + i >= b.r0 && m0 + i <= b.r1;
                    p[i] = band_px(b, on[i] ? major : 0, on[i] ? m0 + i : b.r0);
                }
            }
            for (int i = 0; i < 3; i++) v[i] = on[i] ? get3(p[i]) : 0u;
            for (int i = 0; i < 3; i++) v[i] = blend_px(v[i], color, a[i]);
            for (int i = 0; i < 3; i++) if (on[i]) put3(p[i], v[i]);
        }
        break;
    }
    case REC_MINIPOLY: {
        const int n = (int)rec_flags(r);
        const int x0 = rec_hi1(r), by0 = (int16_t)(r.w[2] & 0xffff), w = (int)((r.w[2] >> 16) & 0xff);
        const int rows = y1 - y0 + 1, npx = rows * w;
        const uint32_t inv = r.w[3];                              // ceil(65536 / w): i / w == (i * inv) >> 16 for i < 144, w <= 12
        const Rec sp = ops.slot_part(1 + n, 2, 7);
        // LANES pixels of the box at a time: each lane keeps its pixel in a register while the edges are
        // applied in OpenCV's order, then the interior fill overrides, then one store
        for (int i0 = 0; i0 < npx; i0 += LANES) {
            const int i = i0 + lane;
            const bool live = i < npx;
            const int ry = (int)(((uint32_t)(live ? i : 0) * inv) >> 16);
            const int x = x0 + ((live ? i : 0) - ry * w), y = y0 + ry;
            uint8_t *p = band_px(b, x, y);
            uint32_t v = live ? get3(p) : 0u;
            for (int e = 0; e < n; e++) {
                const LineAA L = unpack_edge_slot(ops.slot_part(1 + e, 2, 6));
                const int a = lineaa_at(L, x, y, filter);
                if (live && a >= 0) v = blend_px(v, color, a);
            }
            const int rr = y - by0, dx = x - x0;                  // position inside the polygon's box
            const uint32_t word = sp.w[2 + (rr >> 1)] >> ((rr & 1) * 16);
            const int s1 = (int)(word & 0xff), s2 = (int)((word >> 8) & 0xff);
            if (live) put3(p, (dx >= s1 && dx <= s2) ? color : v);
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
        const ConvexChains C = unpack_convex(r, ops.slot_part(1, 2, 7), ops.slot_part(2, 2, 7), ops.slot_part(3, 2, 4));
        // lane q (< rows) evaluates the span of row y0 + q; then lanes are spread over (row, x)
        const int rows = y1 - y0 + 1;
        int mx1 = 1, mx2 = 0;
        if (LANES == 1 || lane < rows) convex_span(C, y0 + (LANES == 1 ? 0 : lane), b.W, mx1, mx2);
        if (LANES == 1) {
            for (int y = y0; y <= y1; y++) {
                convex_span(C, y, b.W, mx1, mx2);
                fill_span<1>(b, y, mx1, mx2, color, 0);
            }
        } else {
            for (int q = 0; q < rows; q++) {
                int x1 = ops.bcast(mx1, q), x2 = ops.bcast(mx2, q);
                fill_span<LANES>(b, y0 + q, x1, x2, color, lane);
            }
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
            if (cnt == 2) {
                fill_span<LANES>(b, y, x1, x2, color, lane);   // one pair: the span between the two crossings
                continue;
            }
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

// grid lines of a use_grid panel (generator.py:112-122): 1-px LINE_8 black lines, drawn last.
// gx / gy: the n line positions int(round(i * W / grid_size)), int(round(i * H / grid_size)).
template <int LANES>
RIG_HD void draw_grid(const Band &b, int n, const int16_t *gx, const int16_t *gy, int lane)
{
    for (int i = 0; i < n; i++) {
        const int x = gx[i], y = gy[i];
        if ((unsigned)x < (unsigned)b.W)
            for (int yy = b.r0 + lane; yy <= b.r1; yy += LANES) put3(band_px(b, x, yy), 0u);
        if (y >= b.r0 && y <= b.r1) fill_span<LANES>(b, y, 0, b.W - 1, 0u, lane);
    }
}

}  // namespace rig
