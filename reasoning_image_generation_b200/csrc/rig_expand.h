// rig_expand.h -- element -> ordered draw records, written so that the same code runs
//   * on the device, one warp per element: LANES = 32, lane = threadIdx.x & 31, sync = __syncwarp
//   * on the host, serially:                LANES = 1,  lane = 0          (tests/cpu_emul)
// Work items (vertices, polygon edges, stroke segments) are independent; their output slots are a
// closed form of (copy, loop, segment), so lanes never need to communicate except through the
// per-warp scratch after a sync.
//
// Record order inside an element is the reference's draw order (shapes.py:282-302,369-375,
// 385-388,562-594): per wrap copy, per loop: fillPoly (boundary lines, interior) then the AA
// polyline segment by segment.
#pragma once
#include "rig_shapes.h"

namespace rig {

struct ExpandScratch {
    long long px[128];
    long long py[128];
};

struct ElemPlan {
    int kind, loops, n;        // polygon kinds: loops x n vertices
    int fill, t;
    uint32_t mask;             // wrap copies drawn
    int ncopies;
    int nv_outer, nv_inner;    // ring vertex counts (circle / crescent)
    uint32_t nrec, naux;
    int ymin, ymax;
};

RIG_HD int popc9(uint32_t m)
{
    int c = 0;
    for (int i = 0; i < 9; i++) c += (m >> i) & 1;
    return c;
}
RIG_HD int ring_recs(int nv, int t) { return t <= 1 ? (nv - 1) : (nv - 1) * stroke_recs(t) + CAP_SLOTS; }

// EllipseEx vertices of a circle outline of radius r px into S[base ..], consecutive duplicates removed as
// OpenCV does.  Lanes compute the raw polygon in parallel; the (rare: tiny radii only) compaction is serial.
// Every lane returns the same vertex count (>= 2).
template <int LANES, class Sync>
RIG_HD int ring_fill(const float *sin_tab, long long cx_fx, long long cy_fx, int r, ExpandScratch &S, int base, int lane, Sync sync)
{
    const long long rad = (long long)r << XY_SHIFT;
    const int delta = ring_delta(rad);
    const int nraw = 360 / delta + 1;                      // angles 0, delta, ..., 360
    sync();
    for (int i = lane; i < nraw; i += LANES) {
        P2L p = ring_vertex(sin_tab, cx_fx, cy_fx, rad, i * delta);
        S.px[base + i] = p.x;
        S.py[base + i] = p.y;
    }
    sync();
    // consecutive duplicates only occur for tiny radii: each lane checks its own vertices, the verdict is shared
    // through the last scratch entry (never a vertex: nraw <= 73 < 127)
    if (lane == 0) S.px[127] = 0;
    sync();
    for (int i = 1 + lane; i < nraw; i += LANES)
        if (S.px[base + i] == S.px[base + i - 1] && S.py[base + i] == S.py[base + i - 1]) S.px[127] = 1;
    sync();
    const bool dup = S.px[127] != 0;
    sync();
    if (!dup) return nraw;
    sync();
    int nv = 1;                                            // all lanes run the same serial compaction
    long long lx = S.px[base], ly = S.py[base];
    for (int i = 1; i < nraw; i++) {
        long long x = S.px[base + i], y = S.py[base + i];
        if (x != lx || y != ly) {
            if (lane == 0) { S.px[base + nv] = x; S.py[base + nv] = y; }
            nv++; lx = x; ly = y;
        }
    }
    if (nv <= 1) {
        if (lane == 0) { S.px[base] = S.px[base + 1] = cx_fx; S.py[base] = S.py[base + 1] = cy_fx; }
        nv = 2;
    }
    sync();
    return nv;
}

// crescent inner circle centre for the copy whose outer centre is (ocx, ocy)  (shapes.py:548-560,589-590)
RIG_HD void crescent_inner(const DElem &e, int ocx, int ocy, int &orad, int &irad, int &icx, int &icy)
{
    double half = ddiv((double)e.size, 2.0);
    orad = dround(half);
    if (orad < 1) orad = 1;
    irad = dround(dmul((double)orad, 0.65));
    int off = dround(dmul((double)orad, 0.35));
    double rx = dsub(dmul((double)off, e.ca), dmul(0.0, e.sa));
    double ry = dadd(dmul((double)off, e.sa), dmul(0.0, e.ca));
    icx = dround(dadd((double)ocx, rx));
    icy = dround(dadd((double)ocy, ry));
}

// ---- phase 1: vertices into scratch, counts into the plan (identical on every lane) ------------
template <int LANES, class Sync>
RIG_HD ElemPlan expand_plan(const DElem &e, const ShapeTables &T, int W, int H, int lane, ExpandScratch &S, Sync sync)
{
    ElemPlan p;
    p.kind = e.kind; p.fill = e.flags & 1; p.t = e.stroke < 1 ? 1 : e.stroke;
    p.loops = 0; p.n = 0; p.mask = 0; p.ncopies = 0; p.nv_outer = p.nv_inner = 0;
    p.nrec = p.naux = 0; p.ymin = 1; p.ymax = 0;
    int margin = p.t + 2;
    int lo = 1 << 30, hi = -(1 << 30);
    if (e.kind == K_CIRCLE || e.kind == K_CRESCENT) {
        double half = ddiv((double)e.size, 2.0);
        int rad = dround(half);
        if (rad < 1) rad = 1;
        p.mask = copy_mask(e.cx - rad, e.cx + rad, e.cy - rad, e.cy + rad, W, H, false);
        for (int b = 0; b < 9; b++)
            if ((p.mask >> b) & 1) {
                int oy = (b % 3 - 1) * H;
                if (e.cy + oy - rad - margin < lo) lo = e.cy + oy - rad - margin;
                if (e.cy + oy + rad + margin > hi) hi = e.cy + oy + rad + margin;
            }
        if (e.kind == K_CIRCLE) {
            p.ncopies = popc9(p.mask);
            p.nv_outer = ring_fill<LANES>(T.sin_tab, (long long)e.cx << XY_SHIFT, (long long)e.cy << XY_SHIFT, rad, S, 0, lane, sync);
            p.nrec = p.ncopies * ((p.fill ? 1 : 0) + ring_recs(p.nv_outer, p.t));
            p.naux = 0;
        } else {
            int orad, irad, icx, icy;
            crescent_inner(e, e.cx, e.cy, orad, irad, icx, icy);
            p.nv_outer = ring_fill<LANES>(T.sin_tab, (long long)e.cx << XY_SHIFT, (long long)e.cy << XY_SHIFT, orad, S, 0, lane, sync);
            p.nv_inner = ring_fill<LANES>(T.sin_tab, (long long)icx << XY_SHIFT, (long long)icy << XY_SHIFT, irad, S, 0, lane, sync);
            int wraps = popc9(p.mask & ~(1u << 4));
            p.ncopies = wraps;
            p.nrec = (p.fill ? 1 : 0) + ring_recs(p.nv_outer, p.t) + ring_recs(p.nv_inner, p.t) + wraps;
            p.naux = 0;
            // the main copy is always drawn (its outlines at least)
            if (e.cy - orad - margin < lo) lo = e.cy - orad - margin;
            if (e.cy + orad + margin > hi) hi = e.cy + orad + margin;
            int il = icy - irad - margin, ih = icy + irad + margin;
            if (il < lo) lo = il;
            if (ih > hi) hi = ih;
        }
    } else {
        p.loops = kind_loops(e.kind);
        p.n = kind_nverts(e.kind);
        int total = p.loops * p.n;
        for (int i = lane; i < total; i += LANES) {
            double x, y;
            int ox, oy;
            unit_vertex(T, e.kind, i / p.n, i % p.n, e.size, x, y);
            place_vertex(x, y, e, ox, oy);
            S.px[i] = ox;
            S.py[i] = oy;
        }
        sync();
        int minx = 1 << 30, maxx = -(1 << 30), miny = 1 << 30, maxy = -(1 << 30);
        for (int i = 0; i < total; i++) {
            int vx = (int)S.px[i], vy = (int)S.py[i];
            if (vx < minx) minx = vx;
            if (vx > maxx) maxx = vx;
            if (vy < miny) miny = vy;
            if (vy > maxy) maxy = vy;
        }
        p.mask = copy_mask(minx, maxx, miny, maxy, W, H, true);
        p.ncopies = popc9(p.mask);
        int per = (p.fill ? p.n + 1 : 0) + p.n * stroke_recs(p.t);
        int pera = p.fill ? p.n : 0;
        p.nrec = p.ncopies * p.loops * per;
        p.naux = p.ncopies * p.loops * pera;
        for (int b = 0; b < 9; b++)
            if ((p.mask >> b) & 1) {
                int oy = (b % 3 - 1) * H;
                if (miny + oy - margin < lo) lo = miny + oy - margin;
                if (maxy + oy + margin > hi) hi = maxy + oy + margin;
            }
    }
    if (lo < 0) lo = 0;
    if (hi > H - 1) hi = H - 1;
    if (lo <= hi && p.nrec > 0) { p.ymin = lo; p.ymax = hi; }
    return p;
}

// stroke one open ring whose vertices sit in S[base .. base+nv)
template <int LANES>
RIG_HD void emit_ring(const ShapeTables &T, int W, int H, int lane, const ExpandScratch &S, int base, int nv, int t,
                      Rec *recs, uint32_t rec0)
{
    int per = stroke_recs(t);
    for (int j = 1 + lane; j < nv; j += LANES) {
        P2L a = {S.px[base + j - 1], S.py[base + j - 1]}, b = {S.px[base + j], S.py[base + j]};
        int flags = j == 1 ? 3 : 2;
        uint32_t ro = rec0 + (j == 1 ? 0 : (uint32_t)((j - 1) * per + (t > 1 ? CAP_SLOTS : 0)));
        emit_thick_line(W, H, a, b, t, flags, XY_SHIFT, 0u, T.slope_corr, T.sin_tab, recs + ro);
    }
}

// ---- phase 2: write the records -----------------------------------------------------------------
// recs / aux are the arena bases; this element owns recs[rec0 .. rec0+plan.nrec) and the aux range
// [aux0 .. aux0+plan.naux) (Q16 units: one per fillPoly scan edge).
template <int LANES, class Sync>
RIG_HD void expand_emit(const DElem &e, const ElemPlan &p, const ShapeTables &T, int W, int H, int lane, ExpandScratch &S,
                        Rec *recs, uint32_t rec0, Q16 *aux4, uint32_t aux0, Sync sync)
{
    const uint32_t color = (uint32_t)e.c0 | ((uint32_t)e.c1 << 8) | ((uint32_t)e.c2 << 16);
    const uint32_t black = 0;
    const int t = p.t;
    if (e.kind == K_CIRCLE) {
        double half = ddiv((double)e.size, 2.0);
        int rad = dround(half);
        if (rad < 1) rad = 1;
        int per = (p.fill ? 1 : 0) + ring_recs(p.nv_outer, t);
        int ci = 0;
        for (int b = 0; b < 9; b++) {
            if (!((p.mask >> b) & 1)) continue;
            int cx = e.cx + (b / 3 - 1) * W, cy = e.cy + (b % 3 - 1) * H;
            uint32_t r0 = rec0 + ci * per;
            if (p.fill && lane == 0) recs[r0] = make_disc(H, cx, cy, rad, color);
            int nv = ring_fill<LANES>(T.sin_tab, (long long)cx << XY_SHIFT, (long long)cy << XY_SHIFT, rad, S, 0, lane, sync);
            emit_ring<LANES>(T, W, H, lane, S, 0, nv, t, recs, r0 + (p.fill ? 1 : 0));
            ci++;
        }
        (void)black;
        return;
    }
    if (e.kind == K_CRESCENT) {
        int orad, irad, icx, icy;
        crescent_inner(e, e.cx, e.cy, orad, irad, icx, icy);
        uint32_t r0 = rec0;
        if (p.fill) {
            if (lane == 0) recs[r0] = make_crescent(H, e.cx, e.cy, orad, icx, icy, irad, color);
            r0++;
        }
        int nvo = ring_fill<LANES>(T.sin_tab, (long long)e.cx << XY_SHIFT, (long long)e.cy << XY_SHIFT, orad, S, 0, lane, sync);
        emit_ring<LANES>(T, W, H, lane, S, 0, nvo, t, recs, r0);
        r0 += ring_recs(nvo, t);
        int nvi = ring_fill<LANES>(T.sin_tab, (long long)icx << XY_SHIFT, (long long)icy << XY_SHIFT, irad, S, 0, lane, sync);
        emit_ring<LANES>(T, W, H, lane, S, 0, nvi, t, recs, r0);
        r0 += ring_recs(nvi, t);
        if (lane == 0) {
            for (int b = 0; b < 9; b++) {
                if (b == 4 || !((p.mask >> b) & 1)) continue;
                int ocx = e.cx + (b / 3 - 1) * W, ocy = e.cy + (b % 3 - 1) * H;
                int o2, i2, jx, jy;
                crescent_inner(e, ocx, ocy, o2, i2, jx, jy);
                recs[r0++] = make_crescent(H, ocx, ocy, o2, jx, jy, i2, color);
            }
        }
        return;
    }
    // polygon kinds
    const int n = p.n;
    const int sper = stroke_recs(t);
    const int per = (p.fill ? n + 1 : 0) + n * sper;
    const int pera = p.fill ? n : 0;
    int ci = 0;
    for (int b = 0; b < 9; b++) {
        if (!((p.mask >> b) & 1)) continue;
        const int ox = (b / 3 - 1) * W, oy = (b % 3 - 1) * H;
        for (int l = 0; l < p.loops; l++) {
            uint32_t r0 = rec0 + (uint32_t)((ci * p.loops + l) * per);
            uint32_t a0 = aux0 + (uint32_t)((ci * p.loops + l) * pera);
            const int vb = l * n;
            if (p.fill) {
                PolyEdge *edges = (PolyEdge *)(aux4 + a0);
                for (int i = lane; i < n; i += LANES) {
                    int i0 = i == 0 ? n - 1 : i - 1;
                    P2L q0 = {S.px[vb + i0] + ox, S.py[vb + i0] + oy}, q1 = {S.px[vb + i] + ox, S.py[vb + i] + oy};
                    Rec line;
                    edges[i] = make_polyedge(W, H, q0, q1, color, line);
                    recs[r0 + i] = line;
                }
                if (lane == 0) {
                    int ymin = 1 << 30, ymax = -(1 << 30);
                    for (int i = 0; i < n; i++) {
                        int vy = (int)S.py[vb + i] + oy;
                        if (vy < ymin) ymin = vy;
                        if (vy > ymax) ymax = vy;
                    }
                    Rec pf = rec_nop();
                    rec_head(pf, REC_POLYFILL, 0, ymin, ymax - 1, 0, color, H);
                    pf.w[2] = a0;
                    pf.w[3] = (uint32_t)n;
                    recs[r0 + n] = pf;
                }
                r0 += n + 1;
                a0 += n;
            }
            for (int i = lane; i < n; i += LANES) {
                int i0 = i == 0 ? n - 1 : i - 1;
                P2L q0 = {S.px[vb + i0] + ox, S.py[vb + i0] + oy}, q1 = {S.px[vb + i] + ox, S.py[vb + i] + oy};
                emit_thick_line(W, H, q0, q1, t, 2, 0, black, T.slope_corr, T.sin_tab, recs + r0 + i * sper);
            }
        }
        ci++;
    }
}

// ---- window summaries ---------------------------------------------------------------------------
// An element's slots are scanned by every band it overlaps.  To let a band skip the 32-slot windows
// that hold nothing for it, the element's arena range starts with one 4-byte (ymin | ymax << 16) row
// range per window, packed 8 to a slot:   [summary slots][payload slots ...]
RIG_HD uint32_t summary_slots(uint32_t nrec) { return ((nrec + 31u) / 32u + 7u) / 8u; }

template <int LANES>
RIG_HD void write_summaries(Rec *recs, uint32_t sum_begin, uint32_t payload_begin, uint32_t nrec, int lane)
{
    uint32_t *sum = (uint32_t *)(recs + sum_begin);
    const uint32_t nwin = (nrec + 31u) / 32u, cap = summary_slots(nrec) * 8u;
    for (uint32_t w = lane; w < cap; w += LANES) {
        int lo = 32767, hi = -32768;          // empty: a range no band overlaps
        if (w < nwin) {
            lo = 1 << 20; hi = -(1 << 20);
            const uint32_t end = (w + 1) * 32u < nrec ? (w + 1) * 32u : nrec;
            for (uint32_t k = w * 32u; k < end; k++) {
                const Rec &r = recs[payload_begin + k];
                const int a = rec_ymin(r), b = rec_ymax(r);
                if (a <= b) { lo = a < lo ? a : lo; hi = b > hi ? b : hi; }
            }
            if (lo > hi) { lo = 32767; hi = -32768; }
        }
        sum[w] = (uint32_t)(uint16_t)(int16_t)lo | ((uint32_t)(uint16_t)(int16_t)hi << 16);
    }
}

}  // namespace rig
