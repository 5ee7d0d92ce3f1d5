// rig_core.h -- per-lane building blocks of the B200 panel rasteriser.
//
// Everything here is a pure function of its arguments, usable from device code (the kernels in
// rig_kernels.cu call them one lane per work item) and from plain host C++ (tests/cpu_emul runs
// the same functions serially so the arithmetic can be checked without a GPU).
//
// What is computed: the reference renders a panel by calling cv2.fillPoly / cv2.polylines(LINE_AA) /
// cv2.circle / cv2.line on integer vertices (/root/reference/src/shapes.py:282-302,564-594,
// /root/reference/src/generator.py:112-122).  Those calls are sequential scan conversions inside
// OpenCV.  Here every such call is split into
//   * a SETUP step (setup_* / make_* / pack_*): one lane turns one segment / polygon into 32-byte record
//     slots whose fields are the loop-invariant integers of OpenCV's scan loop, and
//   * an EVAL step (*_eval): a closed form of iteration k of that loop, so that the 32 lanes of a
//     warp can run 32 iterations at once and a warp that owns a band of rows can jump straight to
//     the iterations that land in its band.
// All arithmetic is integer (16.16 fixed point) or IEEE double with no contraction, and matches
// OpenCV 4.13 bit for bit (oracle/cvdraw.c is the scalar restatement the tests compare against).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RIG_HD __host__ __device__ __forceinline__
// the big set-up routines are real functions on the device: inlining them at every call site made
// k_expand 935 KB of SASS and 59 % of its issue slots instruction-fetch stalls (profiles/r1_notes.md)
#define RIG_FN static __host__ __device__ __noinline__
#else
#include <math.h>
#define RIG_HD inline
#define RIG_FN inline
#endif

namespace rig {

constexpr int XY_SHIFT = 16;
constexpr int XY_ONE = 1 << XY_SHIFT;

// ---- IEEE double helpers that can never be fused into FMAs -------------------------------------
RIG_HD double dmul(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
RIG_HD double dadd(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
RIG_HD double dsub(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    volatile double r = a - b;
    return r;
#endif
}
RIG_HD double ddiv(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b;
    return r;
#endif
}
RIG_HD double dsqrt(double a)
{
#ifdef __CUDA_ARCH__
    return __dsqrt_rn(a);
#else
    return sqrt(a);
#endif
}
// round half to even -> int (Python round(), cvRound)
RIG_HD int dround(double a)
{
#ifdef __CUDA_ARCH__
    return __double2int_rn(a);
#else
    return (int)nearbyint(a);
#endif
}
RIG_HD long long dtrunc(double a) { return (long long)a; }
// n / d truncated toward zero, for |n| < 2^53 and d != 0.  On the device the 64-bit integer divide
// is a long emulation sequence; the correctly rounded double quotient truncates to the same integer
// (a non-integer n/d is at least 1/|d| away from an integer, the rounding error is below |n| 2^-53 / |d|).
RIG_HD long long idiv(long long n, long long d)
{
#ifdef __CUDA_ARCH__
    return (long long)__ddiv_rn((double)n, (double)d);
#else
    return n / d;
#endif
}
// x / 65536 and x * 65536 are exact power-of-two scalings
RIG_HD double dscale_down16(double a) { return dmul(a, 1.0 / 65536.0); }

struct P2L {
    long long x, y;
};
// 16-byte unit of the aux arena
struct alignas(16) Q16 {
    uint32_t x, y, z, w;
};

// ---- 32-byte draw record ------------------------------------------------------------------------
// w0: type | flags<<8 | u16(ymin)<<16      w1: u16(ymax) | 16 type-specific bits << 16
// w7: colour c0 | c1<<8 | c2<<16 in OUTPUT channel order
// ymin..ymax is the inclusive canvas row range the record can touch (ymin > ymax: touches nothing)
struct Rec {
    uint32_t w[8];
};

enum RecType : uint32_t {
    REC_NOP = 0,
    REC_LINE8 = 1,     // Bresenham walk (fillPoly boundary)
    REC_LINEAA = 2,    // 3-tap filtered line
    REC_DISC = 3,      // filled LINE_8 circle
    REC_CRESCENT = 4,  // disc minus disc
    REC_CONVEX = 5,    // interior spans of an AA convex polygon (header + 3 continuation slots)
    REC_POLYFILL = 6,  // even-odd scan fill (aux: nedges x 16 B)
    REC_CONT = 7,      // continuation slot of the record before it (never selected by the band scan)
    REC_MINIPOLY = 8,  // small AA convex polygon, outline + interior, gathered per pixel
};

RIG_HD uint32_t rec_type(const Rec &r) { return r.w[0] & 0xff; }
RIG_HD uint32_t rec_flags(const Rec &r) { return (r.w[0] >> 8) & 0xff; }
RIG_HD int rec_ymin(const Rec &r) { return (int16_t)(r.w[0] >> 16); }
RIG_HD int rec_ymax(const Rec &r) { return (int16_t)(r.w[1] & 0xffff); }
RIG_HD int rec_hi1(const Rec &r) { return (int16_t)(r.w[1] >> 16); }

RIG_HD Rec rec_nop()
{
    Rec r;
    for (int i = 0; i < 8; i++) r.w[i] = 0;
    r.w[0] = REC_NOP | (0x7fffu << 16);  // ymin = 32767 > ymax = -32768: overlaps no band (1 > 0 would overlap rows 0..1)
    r.w[1] = 0x8000u;
    return r;
}
RIG_HD void rec_head(Rec &r, uint32_t type, uint32_t flags, int ymin, int ymax, int hi1, uint32_t color, int H)
{
    if (ymin < 0) ymin = 0;
    if (ymax > H - 1) ymax = H - 1;
    if (ymin > ymax) { ymin = 32767; ymax = -32768; }
    r.w[0] = type | (flags << 8) | ((uint32_t)(uint16_t)ymin << 16);
    r.w[1] = (uint32_t)(uint16_t)ymax | ((uint32_t)(uint16_t)hi1 << 16);
    r.w[7] = color;
}

// ---- clipLine on 64-bit points (pixel units or 16.16 units) ------------------------------------
RIG_FN bool clip_line(long long width, long long height, P2L &p1, P2L &p2)
{
    long long right = width - 1, bottom = height - 1;
    if (width <= 0 || height <= 0) return false;
    long long x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += dtrunc(ddiv(dmul((double)(a - y1), (double)(x2 - x1)), (double)(y2 - y1)));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += dtrunc(ddiv(dmul((double)(a - y2), (double)(x2 - x1)), (double)(y2 - y1)));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += dtrunc(ddiv(dmul((double)(a - x1), (double)(y2 - y1)), (double)(x2 - x1)));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += dtrunc(ddiv(dmul((double)(a - x2), (double)(y2 - y1)), (double)(x2 - x1)));
                x2 = a;
                c2 = 0;
            }
        }
        p1.x = x1; p1.y = y1; p2.x = x2; p2.y = y2;
    }
    return (c1 | c2) == 0;
}

// ---- LINE8: OpenCV Line()/LineIterator, 8-connected, walked from the left end ------------------
// record: w2 = u16(x0) | u16(y0)<<16 ; w3 = u16(dx) | u16(dy)<<16 (dx >= dy >= 0, major/minor)
// flags bit0: y is the major axis; bit1: y decreases along the walk
RIG_FN Rec make_line8(int W, int H, P2L a, P2L b, uint32_t color)
{
    P2L p1 = {(int)a.x, (int)a.y}, p2 = {(int)b.x, (int)b.y};
    if ((unsigned long long)p1.x >= (unsigned long long)W || (unsigned long long)p2.x >= (unsigned long long)W ||
        (unsigned long long)p1.y >= (unsigned long long)H || (unsigned long long)p2.y >= (unsigned long long)H) {
        if (!clip_line(W, H, p1, p2)) return rec_nop();
    }
    int dx = (int)(p2.x - p1.x), dy = (int)(p2.y - p1.y);
    int x = (int)p1.x, y = (int)p1.y;
    uint32_t flags = 0;
    if (dx < 0) { dx = -dx; dy = -dy; x = (int)p2.x; y = (int)p2.y; }
    if (dy < 0) { dy = -dy; flags |= 2; }
    if (dy > dx) { int t = dx; dx = dy; dy = t; flags |= 1; }
    int ya = (int)p1.y, yb = (int)p2.y;
    Rec r = rec_nop();
    rec_head(r, REC_LINE8, flags, ya < yb ? ya : yb, ya < yb ? yb : ya, 0, color, H);
    r.w[2] = (uint32_t)(uint16_t)x | ((uint32_t)(uint16_t)y << 16);
    r.w[3] = (uint32_t)(uint16_t)dx | ((uint32_t)(uint16_t)dy << 16);
    return r;
}
struct Line8 {
    int x0, y0, dx, dy, vert, sy;
};
RIG_HD Line8 line8_unpack(const Rec &r)
{
    Line8 l;
    l.x0 = (int16_t)(r.w[2] & 0xffff);
    l.y0 = (int16_t)(r.w[2] >> 16);
    l.dx = (int)(r.w[3] & 0xffff);
    l.dy = (int)(r.w[3] >> 16);
    l.vert = rec_flags(r) & 1;
    l.sy = (rec_flags(r) & 2) ? -1 : 1;
    return l;
}
// minor-axis offset after k major steps of the error walk  err0 = dx - 2dy; step: m = err<0; err += -2dy + (m? 2dx:0)
RIG_HD int line8_minor(int dx, int dy, int k) { return dx == 0 ? 0 : (2 * k * dy + dx - 1) / (2 * dx); }
// smallest k with line8_minor(k) >= m  (m >= 1, dy > 0)
RIG_HD int line8_first_k(int dx, int dy, int m) { return (2 * dx * m - dx + 1 + 2 * dy - 1) / (2 * dy); }
RIG_HD void line8_eval(const Line8 &l, int k, int &x, int &y)
{
    int m = line8_minor(l.dx, l.dy, k);
    if (l.vert) { y = l.y0 + l.sy * k; x = l.x0 + m; }
    else        { x = l.x0 + k;        y = l.y0 + l.sy * m; }
}
// iterations [k0,k1] whose pixel row is inside [r0,r1]
RIG_HD void line8_krange(const Line8 &l, int r0, int r1, int &k0, int &k1)
{
    int mlo, mhi;  // range of the y-offset index, in walk direction
    if (l.sy > 0) { mlo = r0 - l.y0; mhi = r1 - l.y0; }
    else          { mlo = l.y0 - r1; mhi = l.y0 - r0; }
    if (l.vert) {
        k0 = mlo < 0 ? 0 : mlo;
        k1 = mhi > l.dx ? l.dx : mhi;
    } else if (l.dx < 32) {
        k0 = 0; k1 = l.dx;           // one pass anyway: the caller row-tests each pixel
    } else {
        if (mhi < 0 || mlo > l.dy) { k0 = 1; k1 = 0; return; }
        k0 = mlo <= 0 ? 0 : line8_first_k(l.dx, l.dy, mlo);
        k1 = mhi >= l.dy ? l.dx : line8_first_k(l.dx, l.dy, mhi + 1) - 1;
    }
}

// ---- LINEAA -------------------------------------------------------------------------------------
// loop-invariant state of one LineAA() call (after clipping, end-point swap and slope set-up)
struct LineAA {
    int valid;          // 0: clipped away
    int ymajor;         // 1: one iteration per row, 3 taps along x; 0: one per column, 3 taps along y
    int c0;             // major coordinate of iteration 0
    int F0, step;       // 16.16 minor coordinate of iteration 0 (+0.5 included) and its per-iteration step
    int ecount;         // iterations are k = 0 .. ecount
    uint64_t ep;        // 7 x 9-bit end-point gains: ep[1], ep[2], ep[4], ep[5], ep[6], ep[7], ep[8] (ep[0]=0, ep[3]=ep[1])
    int xmin, xmax, ymin, ymax;   // canvas-clipped pixel footprint (inclusive; may be empty)
};
RIG_FN LineAA setup_lineaa(int W, int H, P2L pt1, P2L pt2, const uint8_t *slope_corr)
{
    LineAA L;
    L.valid = 0; L.ymajor = 0; L.c0 = 0; L.F0 = 0; L.step = 0; L.ecount = -1; L.ep = 0;
    L.xmin = L.ymin = 1; L.xmax = L.ymax = 0;
    long long dx, dy, ax, ay, i, j, step;
    int ecount, slope;
    {
        // clipLine only does something when an end point is outside the canvas: test that inline, call it rarely
        const unsigned long long wf = (unsigned long long)W << XY_SHIFT, hf = (unsigned long long)H << XY_SHIFT;
        if (((unsigned long long)pt1.x >= wf || (unsigned long long)pt2.x >= wf || (unsigned long long)pt1.y >= hf ||
             (unsigned long long)pt2.y >= hf) &&
            !clip_line((long long)W << XY_SHIFT, (long long)H << XY_SHIFT, pt1, pt2))
            return L;
    }
    dx = pt2.x - pt1.x;
    dy = pt2.y - pt1.y;
    j = dx < 0 ? -1 : 0;
    ax = (dx ^ j) - j;
    i = dy < 0 ? -1 : 0;
    ay = (dy ^ i) - i;
    long long c0, F0;
    if (ax > ay) {
        L.ymajor = 0;
        dy = (dy ^ j) - j;
        if (j) { P2L t = pt1; pt1 = pt2; pt2 = t; }
        step = idiv(dy * XY_ONE, ax | 1);
        pt2.x += XY_ONE;
        ecount = (int)((pt2.x >> XY_SHIFT) - (pt1.x >> XY_SHIFT));
        j = -(pt1.x & (XY_ONE - 1));
        pt1.y += ((step * j) >> XY_SHIFT) + (XY_ONE >> 1);
        slope = (int)((step >> (XY_SHIFT - 5)) & 0x3f);
        slope ^= (step < 0 ? 0x3f : 0);
        i = (pt1.x >> (XY_SHIFT - 7)) & 0x78;
        j = (pt2.x >> (XY_SHIFT - 7)) & 0x78;
        c0 = pt1.x >> XY_SHIFT;
        F0 = pt1.y;
    } else {
        L.ymajor = 1;
        dx = (dx ^ i) - i;
        if (i) { P2L t = pt1; pt1 = pt2; pt2 = t; }
        step = idiv(dx * XY_ONE, ay | 1);
        pt2.y += XY_ONE;
        ecount = (int)((pt2.y >> XY_SHIFT) - (pt1.y >> XY_SHIFT));
        j = -(pt1.y & (XY_ONE - 1));
        pt1.x += ((step * j) >> XY_SHIFT) + (XY_ONE >> 1);
        slope = (int)((step >> (XY_SHIFT - 5)) & 0x3f);
        slope ^= (step < 0 ? 0x3f : 0);
        i = (pt1.y >> (XY_SHIFT - 7)) & 0x78;
        j = (pt2.y >> (XY_SHIFT - 7)) & 0x78;
        c0 = pt1.y >> XY_SHIFT;
        F0 = pt1.x;
    }
    slope = (slope & 0x20) ? 0x100 : slope_corr[slope];
    int ii = (int)i, jj = (int)j;
    int t0 = slope << 7;
    int t1 = ((0x78 - ii) | 4) * slope;
    int t2 = (jj | 4) * slope;
    uint32_t ep1 = ((((jj - ii) & 0x78) | 4) * slope >> 8) & 0x1ff;
    uint32_t ep2 = (t1 >> 8) & 0x1ff;
    uint32_t ep4 = (((((jj - ii) + 0x80) | 4) * slope) >> 8) & 0x1ff;
    uint32_t ep5 = ((t1 + t0) >> 8) & 0x1ff;
    uint32_t ep6 = (t2 >> 8) & 0x1ff;
    uint32_t ep7 = ((t2 + t0) >> 8) & 0x1ff;
    uint32_t ep8 = (uint32_t)slope;
    L.ep = (uint64_t)ep1 | ((uint64_t)ep2 << 9) | ((uint64_t)ep4 << 18) | ((uint64_t)ep5 << 27) |
           ((uint64_t)ep6 << 36) | ((uint64_t)ep7 << 45) | ((uint64_t)ep8 << 54);
    L.valid = 1;
    L.c0 = (int)c0; L.F0 = (int)F0; L.step = (int)step; L.ecount = ecount;
    // footprint: major c0..c0+ecount, minor (pos>>16)-1 .. +1 (pos is monotone in k)
    int ma = (int)(F0 >> XY_SHIFT) - 1;
    int mb = (int)((F0 + (long long)ecount * step) >> XY_SHIFT) - 1;
    int mlo = ma < mb ? ma : mb, mhi = (ma < mb ? mb : ma) + 2;
    int jlo = (int)c0, jhi = (int)c0 + ecount;
    if (L.ymajor) { L.ymin = jlo; L.ymax = jhi; L.xmin = mlo; L.xmax = mhi; }
    else          { L.xmin = jlo; L.xmax = jhi; L.ymin = mlo; L.ymax = mhi; }
    if (L.xmin < 0) L.xmin = 0;
    if (L.ymin < 0) L.ymin = 0;
    if (L.xmax > W - 1) L.xmax = W - 1;
    if (L.ymax > H - 1) L.ymax = H - 1;
    return L;
}
// stand-alone record: flags bit0 = y-major; hi1 = c0; w2 = F0; w3 = step; w4 = ecount; w5,w6 = ep
RIG_HD Rec pack_lineaa(const LineAA &L, uint32_t color, int H)
{
    Rec r = rec_nop();
    if (!L.valid) return r;
    rec_head(r, REC_LINEAA, (uint32_t)L.ymajor, L.ymin, L.ymax, L.c0, color, H);
    r.w[2] = (uint32_t)L.F0;
    r.w[3] = (uint32_t)L.step;
    r.w[4] = (uint32_t)L.ecount;
    r.w[5] = (uint32_t)L.ep;
    r.w[6] = (uint32_t)(L.ep >> 32);
    return r;
}
RIG_HD Rec make_lineaa(int W, int H, P2L pt1, P2L pt2, uint32_t color, const uint8_t *slope_corr)
{
    return pack_lineaa(setup_lineaa(W, H, pt1, pt2, slope_corr), color, H);
}
RIG_HD LineAA lineaa_unpack(const Rec &r)
{
    LineAA l;
    l.valid = 1;
    l.ymajor = rec_flags(r) & 1;
    l.c0 = rec_hi1(r);
    l.F0 = (int32_t)r.w[2];
    l.step = (int32_t)r.w[3];
    l.ecount = (int)r.w[4];
    l.ep = (uint64_t)r.w[5] | ((uint64_t)r.w[6] << 32);
    l.xmin = l.ymin = 0; l.xmax = l.ymax = 0;
    return l;
}
RIG_HD int lineaa_ep(const LineAA &l, int k)
{
    // ep_table[3 * s + e], s/e = min(count, 2) from the start / end of the line
    int s = k < 2 ? k : 2, e = l.ecount - k;
    e = e < 2 ? e : 2;
    int idx = 3 * s + e;
    // idx: 0 1 2 3 4 5 6 7 8  -> packed field: -,0,1,0,2,3,4,5,6
    int slot = idx == 3 ? 0 : (idx < 3 ? idx - 1 : idx - 2);
    int v = (int)((l.ep >> (9 * (slot < 0 ? 0 : slot))) & 0x1ff);
    return idx == 0 ? 0 : v;
}
// iteration k: major coordinate, first minor coordinate of the 3 taps, and the 3 blend weights
RIG_HD void lineaa_eval(const LineAA &l, int k, const uint8_t *filter, int &major, int &minor0, int &a0, int &a1, int &a2)
{
    int pos = l.F0 + k * l.step;
    major = l.c0 + k;
    minor0 = (pos >> XY_SHIFT) - 1;
    int dist = (pos >> (XY_SHIFT - 5)) & 31;
    int ep = lineaa_ep(l, k);
    a0 = (ep * filter[dist + 32] >> 8) & 0xff;
    a1 = (ep * filter[dist] >> 8) & 0xff;
    a2 = (ep * filter[63 - dist] >> 8) & 0xff;
}
// blend weight this line applies to pixel (x, y), or -1 when it does not touch it (gather form)
RIG_HD int lineaa_at(const LineAA &l, int x, int y, const uint8_t *filter)
{
    int k = (l.ymajor ? y : x) - l.c0;
    if (k < 0 || k > l.ecount) return -1;
    int pos = l.F0 + k * l.step;
    int d = (l.ymajor ? x : y) - ((pos >> XY_SHIFT) - 1);
    if ((unsigned)d > 2u) return -1;
    int dist = (pos >> (XY_SHIFT - 5)) & 31;
    int fi = d == 0 ? dist + 32 : (d == 1 ? dist : 63 - dist);
    return (lineaa_ep(l, k) * filter[fi] >> 8) & 0xff;
}
RIG_HD int floor_div(int a, int b)  // b > 0
{
    int q = a / b;
    return (a % b < 0) ? q - 1 : q;
}
RIG_HD int ceil_div(int a, int b) { return -floor_div(-a, b); }  // b > 0
// iterations [k0,k1] that can touch rows [r0,r1]
RIG_HD void lineaa_krange(const LineAA &l, int r0, int r1, int &k0, int &k1)
{
    if (l.ymajor) {
        k0 = r0 - l.c0;
        k1 = r1 - l.c0;
    } else if (l.ecount < 32) {
        k0 = 0; k1 = l.ecount;          // one pass anyway: the taps are row-tested individually
    } else {
        // need (pos>>16)-1 <= r1 and (pos>>16)+1 >= r0  <=>  lo <= pos <= hi
        int lo = (r0 - 1) << XY_SHIFT, hi = ((r1 + 2) << XY_SHIFT) - 1;
        if (l.step == 0) {
            if (l.F0 < lo || l.F0 > hi) { k0 = 1; k1 = 0; return; }
            k0 = 0; k1 = l.ecount;
        } else if (l.step > 0) {
            k0 = ceil_div(lo - l.F0, l.step);
            k1 = floor_div(hi - l.F0, l.step);
        } else {
            k0 = ceil_div(l.F0 - hi, -l.step);
            k1 = floor_div(l.F0 - lo, -l.step);
        }
    }
    if (k0 < 0) k0 = 0;
    if (k1 > l.ecount) k1 = l.ecount;
}
RIG_HD int blend1(int v, int c, int a)
{
    v += ((c - v) * a + 127) >> 8;
    v += ((c - v) * a + 127) >> 8;
    return v;
}

// ---- interior of FillConvexPoly(line_type = LINE_AA, shift = 16) --------------------------------
// The two-edge walk is replayed once at setup; what is kept per side is the list of
// (first row, x at that row, dx per row) pieces, so a row's span is a closed form.
struct ConvexChains {
    int valid;
    int ymin, ylast;        // rows ymin..ylast are filled (ymin may be negative)
    int nseg[2];
    int32_t x[8], dx[8];    // side s uses entries 4s .. 4s+3
    int32_t y[8];
};
RIG_FN ConvexChains setup_convex(int W, int H, const P2L *v, int npts)
{
    const int shift = XY_SHIFT;
    const int delta = 1 << shift >> 1;
    ConvexChains C;
    C.valid = 0; C.ymin = 1; C.ylast = 0; C.nseg[0] = C.nseg[1] = 0;
    for (int i = 0; i < 8; i++) { C.x[i] = 0; C.dx[i] = 0; C.y[i] = 0; }
    long long xmin = v[0].x, xmax = v[0].x, ymin = v[0].y, ymax = v[0].y;
    int imin = 0;
    for (int i = 0; i < npts; i++) {
        if (v[i].y < ymin) { ymin = v[i].y; imin = i; }
        if (v[i].y > ymax) ymax = v[i].y;
        if (v[i].x > xmax) xmax = v[i].x;
        if (v[i].x < xmin) xmin = v[i].x;
    }
    xmin = (xmin + delta) >> shift;
    xmax = (xmax + delta) >> shift;
    ymin = (ymin + delta) >> shift;
    ymax = (ymax + delta) >> shift;
    if (npts < 3 || (int)xmax < 0 || (int)ymax < 0 || (int)xmin >= W || (int)ymin >= H) return C;
    if (ymax > H - 1) ymax = H - 1;

    struct { int idx, di; long long x, dx; int ye; } edge[2];
    int edges = npts;
    int y = (int)ymin;
    edge[0].idx = edge[1].idx = imin;
    edge[0].ye = edge[1].ye = y;
    edge[0].di = 1;
    edge[1].di = npts - 1;
    edge[0].x = edge[1].x = -XY_ONE;
    edge[0].dx = edge[1].dx = 0;
    int ylast = y - 1;
    do {
        if (y < (int)ymax || y == (int)ymin) {
            for (int i = 0; i < 2; i++) {
                if (y >= edge[i].ye) {
                    int idx0 = edge[i].idx, di = edge[i].di;
                    int idx = idx0 + di;
                    if (idx >= npts) idx -= npts;
                    for (; edges-- > 0;) {
                        int ty = (int)((v[idx].y + delta) >> shift);
                        if (ty > y) {
                            long long xs = v[idx0].x, xe = v[idx].x;
                            edge[i].ye = ty;
                            edge[i].dx = idiv((xe - xs) * 2 + (ty - y), 2 * (ty - y));
                            edge[i].x = xs;
                            edge[i].idx = idx;
                            if (C.nseg[i] < 4) {
                                C.x[4 * i + C.nseg[i]] = (int32_t)edge[i].x;
                                C.dx[4 * i + C.nseg[i]] = (int32_t)edge[i].dx;
                                C.y[4 * i + C.nseg[i]] = y;
                                C.nseg[i]++;
                            }
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
        // jump to the next row at which a side may be re-initialised
        int ynext = edge[0].ye < edge[1].ye ? edge[0].ye : edge[1].ye;
        if (ynext <= y) ynext = y + 1;
        if (ynext > (int)ymax) ynext = (int)ymax + 1;
        int adv = ynext - y;
        edge[0].x += edge[0].dx * adv;
        edge[1].x += edge[1].dx * adv;
        ylast = ynext - 1;
        y = ynext;
    } while (y <= (int)ymax);

    if (ylast < (int)ymin || C.nseg[0] == 0 || C.nseg[1] == 0) return C;
    C.valid = 1;
    C.ymin = (int)ymin;
    C.ylast = ylast;
    return C;
}
// span of row y (inclusive pixel range, already clipped; x1 > x2 means empty)
RIG_HD long long convex_side_x(const ConvexChains &a, int side, int y)
{
    const int o = 4 * side, n = a.nseg[side];
    int32_t sx = a.x[o], sdx = a.dx[o], sy = a.y[o];
    if (n > 1 && y >= a.y[o + 1]) { sx = a.x[o + 1]; sdx = a.dx[o + 1]; sy = a.y[o + 1]; }
    if (n > 2 && y >= a.y[o + 2]) { sx = a.x[o + 2]; sdx = a.dx[o + 2]; sy = a.y[o + 2]; }
    if (n > 3 && y >= a.y[o + 3]) { sx = a.x[o + 3]; sdx = a.dx[o + 3]; sy = a.y[o + 3]; }
    return (long long)sx + (long long)(y - sy) * sdx;
}
RIG_HD void convex_span(const ConvexChains &a, int y, int W, int &x1, int &x2)
{
    long long xa = convex_side_x(a, 0, y), xb = convex_side_x(a, 1, y);
    long long xl = xa > xb ? xb : xa, xr = xa > xb ? xa : xb;
    x1 = (int)((xl + (XY_ONE - 1)) >> XY_SHIFT);
    x2 = (int)(xr >> XY_SHIFT);
    if (x2 >= 0 && x1 < W) {
        if (x1 < 0) x1 = 0;
        if (x2 >= W) x2 = W - 1;
    } else {
        x1 = 1; x2 = 0;
    }
}
// REC_CONVEX: header slot + 3 continuation slots (quads: at most 3 pieces per side)
//   header: w2 = nseg0 | nseg1 << 8
//   cont 1: w2..w7 = x[0..2], x[4..6]      cont 2: w2..w7 = dx[0..2], dx[4..6]
//   cont 3: w2..w4 = y packed as int16 pairs (y0|y1<<16, y2|y4<<16, y5|y6<<16)
constexpr int CONVEX_SLOTS = 4;
RIG_HD Rec rec_cont()
{
    Rec r = rec_nop();
    r.w[0] = (uint32_t)REC_CONT | (0x7fffu << 16);   // ymin > ymax: never selected
    return r;
}
RIG_FN void pack_convex(const ConvexChains &C, uint32_t color, int H, Rec *out)
{
    out[0] = rec_nop();
    for (int i = 1; i < CONVEX_SLOTS; i++) out[i] = rec_cont();
    if (!C.valid || C.nseg[0] > 3 || C.nseg[1] > 3) return;
    rec_head(out[0], REC_CONVEX, 0, C.ymin, C.ylast, 0, color, H);
    out[0].w[2] = (uint32_t)C.nseg[0] | ((uint32_t)C.nseg[1] << 8);
    for (int i = 0; i < 3; i++) {
        out[1].w[2 + i] = (uint32_t)C.x[i];  out[1].w[5 + i] = (uint32_t)C.x[4 + i];
        out[2].w[2 + i] = (uint32_t)C.dx[i]; out[2].w[5 + i] = (uint32_t)C.dx[4 + i];
    }
    const int ys[6] = {C.y[0], C.y[1], C.y[2], C.y[4], C.y[5], C.y[6]};
    for (int i = 0; i < 3; i++)
        out[3].w[2 + i] = (uint32_t)(uint16_t)(int16_t)ys[2 * i] | ((uint32_t)(uint16_t)(int16_t)ys[2 * i + 1] << 16);
}
RIG_HD ConvexChains unpack_convex(const Rec &h, const Rec &c1, const Rec &c2, const Rec &c3)
{
    ConvexChains C;
    C.valid = 1;
    C.ymin = rec_ymin(h); C.ylast = rec_ymax(h);
    C.nseg[0] = h.w[2] & 0xff; C.nseg[1] = (h.w[2] >> 8) & 0xff;
    for (int i = 0; i < 8; i++) { C.x[i] = 0; C.dx[i] = 0; C.y[i] = 0; }
    for (int i = 0; i < 3; i++) {
        C.x[i] = (int32_t)c1.w[2 + i];  C.x[4 + i] = (int32_t)c1.w[5 + i];
        C.dx[i] = (int32_t)c2.w[2 + i]; C.dx[4 + i] = (int32_t)c2.w[5 + i];
    }
    C.y[0] = (int16_t)(c3.w[2] & 0xffff); C.y[1] = (int16_t)(c3.w[2] >> 16);
    C.y[2] = (int16_t)(c3.w[3] & 0xffff); C.y[4] = (int16_t)(c3.w[3] >> 16);
    C.y[5] = (int16_t)(c3.w[4] & 0xffff); C.y[6] = (int16_t)(c3.w[4] >> 16);
    return C;
}

// ---- MINIPOLY: a whole small AA convex polygon (outline + interior) as one gather step ---------
// For polygons whose footprint fits a 12 x 12 pixel box (stroke end caps, short stroke quads) the
// warp does not replay the edges one after another; each lane takes one pixel of the box and applies,
// in OpenCV's order, every edge's blend that lands on it and finally the interior fill.
//   header : flags = number of edges; hi1 = x0; w2 = u16(y0) | w << 16 | h << 24; w3 = ceil(65536 / w)
//   edge e : cont slot 1+e: w2 = F0, w3 = step, w4,w5 = ep, w6 = u16(c0) | (ecount & 0xff) << 16 | ymajor << 24 | valid << 25
//   spans  : cont slot 1+nedges: 12 rows x (x1, x2) as uint8 offsets from x0 (x1 > x2: empty), bytes of w2..w7
constexpr int MINI_MAX = 12;
constexpr int MINI_MAX_EDGES = 5;
RIG_HD int minipoly_slots(int nedges) { return nedges + 2; }
RIG_HD void pack_edge_slot(const LineAA &L, Rec &s)
{
    s = rec_cont();
    s.w[2] = (uint32_t)L.F0;
    s.w[3] = (uint32_t)L.step;
    s.w[4] = (uint32_t)L.ep;
    s.w[5] = (uint32_t)(L.ep >> 32);
    s.w[6] = (uint32_t)(uint16_t)(int16_t)L.c0 | ((uint32_t)(L.ecount & 0xff) << 16) | ((uint32_t)(L.ymajor & 1) << 24) |
             ((uint32_t)(L.valid && L.ecount >= 0 ? 1 : 0) << 25);
}
RIG_HD LineAA unpack_edge_slot(const Rec &s)
{
    LineAA l;
    l.F0 = (int32_t)s.w[2];
    l.step = (int32_t)s.w[3];
    l.ep = (uint64_t)s.w[4] | ((uint64_t)s.w[5] << 32);
    l.c0 = (int16_t)(s.w[6] & 0xffff);
    l.ecount = (int)((s.w[6] >> 16) & 0xff);
    l.ymajor = (int)((s.w[6] >> 24) & 1);
    l.valid = (int)((s.w[6] >> 25) & 1);
    if (!l.valid) l.ecount = -1;
    l.xmin = l.ymin = 0; l.xmax = l.ymax = 0;
    return l;
}
// Tries to express the AA convex polygon v[0..n) as a MINIPOLY at out[0 .. n+2).  Returns false (and
// writes nothing) when its footprint is larger than the box or an edge is too long for the packing.
RIG_FN bool pack_minipoly(int W, int H, const LineAA *E, int n, const ConvexChains &C, uint32_t color, Rec *out)
{
    if (n > MINI_MAX_EDGES) return false;
    // the outline's footprint contains the interior, so it alone decides whether the box is small
    int xmin = 1 << 30, xmax = -(1 << 30), ymin = 1 << 30, ymax = -(1 << 30);
    for (int e = 0; e < n; e++) {
        if (!E[e].valid) continue;
        if (E[e].ecount > 255 || E[e].c0 < -32768 || E[e].c0 > 32767) return false;
        if (E[e].xmin > E[e].xmax || E[e].ymin > E[e].ymax) continue;
        if (E[e].xmin < xmin) xmin = E[e].xmin;
        if (E[e].xmax > xmax) xmax = E[e].xmax;
        if (E[e].ymin < ymin) ymin = E[e].ymin;
        if (E[e].ymax > ymax) ymax = E[e].ymax;
    }
    int fy0 = 1, fy1 = 0;
    if (C.valid) {
        fy0 = C.ymin < 0 ? 0 : C.ymin;
        fy1 = C.ylast > H - 1 ? H - 1 : C.ylast;
        if (fy0 <= fy1) {
            if (fy0 < ymin) ymin = fy0;
            if (fy1 > ymax) ymax = fy1;
        }
    }
    if (ymin <= ymax && (ymax - ymin + 1 > MINI_MAX || (xmin <= xmax && xmax - xmin + 1 > MINI_MAX))) return false;
    // interior spans of the (at most MINI_MAX) rows, evaluated once
    // packed as (x1, x2) byte pairs relative to a provisional origin; the origin can only move left
    // by the interior, which lies inside the outline footprint, so it is fixed up afterwards
    int sx1[MINI_MAX], sx2[MINI_MAX];
    for (int r = 0; r < MINI_MAX; r++) { sx1[r] = 1; sx2[r] = 0; }
    if (C.valid && ymin <= ymax)
        for (int y = fy0; y <= fy1; y++) {
            const int r = y - ymin;
            convex_span(C, y, W, sx1[r], sx2[r]);
            if (sx1[r] <= sx2[r]) {
                if (sx1[r] < xmin) xmin = sx1[r];
                if (sx2[r] > xmax) xmax = sx2[r];
            }
        }
    if (xmin > xmax || ymin > ymax) {      // nothing visible: an empty record keeps the slot layout
        out[0] = rec_nop();
        for (int i = 1; i < minipoly_slots(n); i++) out[i] = rec_cont();
        return true;
    }
    const int w = xmax - xmin + 1, h = ymax - ymin + 1;
    if (w > MINI_MAX || h > MINI_MAX) return false;
    out[0] = rec_nop();
    rec_head(out[0], REC_MINIPOLY, (uint32_t)n, ymin, ymax, xmin, color, H);
    out[0].w[2] = (uint32_t)(uint16_t)(int16_t)ymin | ((uint32_t)w << 16) | ((uint32_t)h << 24);
    out[0].w[3] = (65536u + (uint32_t)w - 1u) / (uint32_t)w;   // i / w == (i * w3) >> 16 for i < 144: the rasteriser never divides
    for (int e = 0; e < n; e++) pack_edge_slot(E[e], out[1 + e]);
    Rec &sp = out[1 + n];
    sp = rec_cont();
    uint32_t words[6] = {0, 0, 0, 0, 0, 0};
    for (int r = 0; r < MINI_MAX; r++) {
        uint32_t b1 = 1, b2 = 0;
        if (sx1[r] <= sx2[r]) { b1 = (uint32_t)(sx1[r] - xmin); b2 = (uint32_t)(sx2[r] - xmin); }
        words[r >> 1] |= (b1 | (b2 << 8)) << ((r & 1) * 16);
    }
    for (int i = 0; i < 6; i++) sp.w[2 + i] = words[i];
    return true;
}

// ---- POLYFILL: one scan edge of fillPoly (LINE_8, shift 0) --------------------------------------
// aux entry (16 B): int16 y0, int16 y1, int32 x, int32 dx, int32 valid
struct PolyEdge {
    int16_t y0, y1;
    int32_t x, dx;
    int32_t valid;
};
// pt0 -> pt1 are consecutive integer vertices; also returns the boundary LINE8 record
RIG_FN PolyEdge make_polyedge(int W, int H, P2L pt0, P2L pt1, uint32_t color, Rec &boundary)
{
    PolyEdge e;
    e.valid = 0; e.y0 = e.y1 = 0; e.x = e.dx = 0;
    boundary = make_line8(W, H, pt0, pt1, color);
    long long x0 = pt0.x << XY_SHIFT, x1 = pt1.x << XY_SHIFT;
    long long cy0 = pt0.y, cy1 = pt1.y;
    if ((unsigned long long)pt0.x >= (unsigned long long)W || (unsigned long long)pt1.x >= (unsigned long long)W ||
        (unsigned long long)pt0.y >= (unsigned long long)H || (unsigned long long)pt1.y >= (unsigned long long)H) {
        P2L a = {(int)pt0.x, (int)pt0.y}, b = {(int)pt1.x, (int)pt1.y};
        clip_line(W, H, a, b);
        if (a.y != b.y) { cy0 = a.y; cy1 = b.y; }
        x0 = a.x << XY_SHIFT;
        x1 = b.x << XY_SHIFT;
    }
    if (pt0.y == pt1.y) return e;
    long long dx = idiv(x1 - x0, cy1 - cy0);
    long long x;
    if (pt0.y < pt1.y) { e.y0 = (int16_t)pt0.y; e.y1 = (int16_t)pt1.y; x = x0 + (pt0.y - cy0) * dx; }
    else               { e.y0 = (int16_t)pt1.y; e.y1 = (int16_t)pt0.y; x = x1 + (pt1.y - cy1) * dx; }
    e.x = (int32_t)x;
    e.dx = (int32_t)dx;
    e.valid = 1;
    return e;
}
RIG_HD bool polyedge_active(const PolyEdge &e, int y) { return e.valid && y >= e.y0 && y < e.y1; }
RIG_HD long long polyedge_x(const PolyEdge &e, int y) { return (long long)e.x + (long long)(y - e.y0) * e.dx; }

// ---- DISC / CRESCENT ----------------------------------------------------------------------------
// Circle(fill) of radius r at an integer centre is exactly { dx*dx + dy*dy <= r*r } (checked for
// every r in 0..399 against the installed wheel, tests/test_oracle_vs_cv2.py)
RIG_HD int isqrt_floor(int v)
{
    if (v <= 0) return 0;
#ifdef __CUDA_ARCH__
    int s = (int)__fsqrt_rz((float)v);
#else
    int s = (int)sqrtf((float)v);
#endif
    while (s * s > v) s--;
    while ((s + 1) * (s + 1) <= v) s++;
    return s;
}
RIG_HD Rec make_disc(int H, int cx, int cy, int r, uint32_t color)
{
    Rec rec = rec_nop();
    rec_head(rec, REC_DISC, 0, cy - r, cy + r, 0, color, H);
    rec.w[2] = (uint32_t)cx; rec.w[3] = (uint32_t)cy; rec.w[4] = (uint32_t)r;
    return rec;
}
RIG_HD Rec make_crescent(int H, int cx, int cy, int r, int icx, int icy, int ir, uint32_t color)
{
    Rec rec = rec_nop();
    rec_head(rec, REC_CRESCENT, 0, cy - r, cy + r, 0, color, H);
    rec.w[2] = (uint32_t)cx; rec.w[3] = (uint32_t)cy; rec.w[4] = (uint32_t)r;
    rec.w[5] = (uint32_t)(uint16_t)icx | ((uint32_t)(uint16_t)icy << 16);
    rec.w[6] = (uint32_t)ir;
    return rec;
}

// ---- ThickLine(thickness > 1, LINE_AA): quad + end caps -----------------------------------------
// A thick AA segment is FillConvexPoly(quad) followed, per cap, by FillConvexPoly(EllipseEx polygon of
// "radius" thickness/2 = a 90-degree-step diamond for thickness <= 4).  Slot budget (fixed, so that
// segment i of a polyline always starts at a closed-form offset):
//   quad : QUAD_SLOTS = 8   either 4 LINEAA + CONVEX(4)  or  MINIPOLY(4 edges) = 6 slots + 2 unused
//   cap  : CAP_SLOTS  = 7   MINIPOLY(5 edges)
constexpr int QUAD_SLOTS = 8;
constexpr int CAP_SLOTS = 7;

RIG_FN void cap_polygon(P2L c, long long rad, const float *sin_tab, P2L *v, int &nv)
{
    // EllipseEx(center c, axes (rad,rad), angle 0, 0..360) with delta = 90 (rad < 3 px): the polygon
    // c + rad * (cos, sin) at 0, 90, 180, 270, 360 degrees.  SinTable is exactly 0 / +-1 / -0 at those
    // angles, so every product and sum of the double-precision EllipseEx arithmetic is exact and the
    // sub-pixel re-rounding is the identity: the vertices are c.x +- rad, c.y +- rad as integers.
    (void)sin_tab;
    if (rad == 0) { v[0] = v[1] = c; nv = 2; return; }
    v[0].x = c.x + rad; v[0].y = c.y;
    v[1].x = c.x;       v[1].y = c.y + rad;
    v[2].x = c.x - rad; v[2].y = c.y;
    v[3].x = c.x;       v[3].y = c.y - rad;
    v[4] = v[0];
    nv = 5;
}

// Records of one AA convex polygon (n <= 5 vertices) into out[0 .. budget): a MINIPOLY when it is small,
// else (n == 4 only) 4 LINEAA + CONVEX.  Unused slots become NOPs.
RIG_FN void emit_convex_aa(int W, int H, const P2L *v, int n, uint32_t color, const uint8_t *slope_corr, Rec *out, int budget)
{
    LineAA E[MINI_MAX_EDGES];
    P2L p0 = v[n - 1];
    for (int i = 0; i < n; i++) {
        E[i] = setup_lineaa(W, H, p0, v[i], slope_corr);
        p0 = v[i];
    }
    ConvexChains C = setup_convex(W, H, v, n);
    int used = 0;
    if (minipoly_slots(n) <= budget && pack_minipoly(W, H, E, n, C, color, out)) used = minipoly_slots(n);
    else if (n + CONVEX_SLOTS <= budget) {
        for (int i = 0; i < n; i++) out[i] = pack_lineaa(E[i], color, H);
        pack_convex(C, color, H, out + n);
        used = n + CONVEX_SLOTS;
    }
    for (int i = used; i < budget; i++) out[i] = rec_nop();      // every slot is written exactly once
}

// One ThickLine call.  p0,p1 in `shift` fractional bits (0 for polylines, 16 for circle arcs).
// flags bit0: cap at p0, bit1: cap at p1.  Writes exactly
//   thickness <= 1 : 1 slot
//   thickness  > 1 : QUAD_SLOTS + CAP_SLOTS * popcount(flags) slots
RIG_FN int emit_thick_line(int W, int H, P2L p0, P2L p1, int thickness, int flags, int shift, uint32_t color,
                           const uint8_t *slope_corr, const float *sin_tab, Rec *out)
{
    if (thickness <= 1) {
        p0.x <<= XY_SHIFT - shift; p0.y <<= XY_SHIFT - shift;
        p1.x <<= XY_SHIFT - shift; p1.y <<= XY_SHIFT - shift;
        out[0] = make_lineaa(W, H, p0, p1, color, slope_corr);
        return 1;
    }
    int ncap = (flags & 1) + ((flags >> 1) & 1);
    int nrec = QUAD_SLOTS + CAP_SLOTS * ncap;
    if (shift == 0) {
        // OpenCV 4.13: integer end points are clipped to the canvas grown by `thickness` first
        P2L a = {p0.x + thickness, p0.y + thickness}, b = {p1.x + thickness, p1.y + thickness};
        if (!clip_line((long long)W + 2 * thickness, (long long)H + 2 * thickness, a, b)) {
            for (int i = 0; i < nrec; i++) out[i] = rec_nop();
            return nrec;
        }
        p0.x = a.x - thickness; p0.y = a.y - thickness;
        p1.x = b.x - thickness; p1.y = b.y - thickness;
    }
    p0.x <<= XY_SHIFT - shift; p0.y <<= XY_SHIFT - shift;
    p1.x <<= XY_SHIFT - shift; p1.y <<= XY_SHIFT - shift;

    const double INV_XY_ONE = 1. / XY_ONE;
    double dx = dmul((double)(p0.x - p1.x), INV_XY_ONE), dy = dmul((double)(p1.y - p0.y), INV_XY_ONE);
    double r = dadd(dmul(dx, dx), dmul(dy, dy));
    int odd = thickness & 1;
    long long th = (long long)thickness << (XY_SHIFT - 1);
    int o = 0;
    if ((r < 0 ? -r : r) > 2.2204460492503131e-16) {
        r = ddiv(dadd((double)th, dmul((double)(odd * XY_ONE), 0.5)), dsqrt(r));
        long long dpx = dround(dmul(dy, r)), dpy = dround(dmul(dx, r));
        P2L pt[4];
        pt[0].x = p0.x + dpx; pt[0].y = p0.y + dpy;
        pt[1].x = p0.x - dpx; pt[1].y = p0.y - dpy;
        pt[2].x = p1.x - dpx; pt[2].y = p1.y - dpy;
        pt[3].x = p1.x + dpx; pt[3].y = p1.y + dpy;
        emit_convex_aa(W, H, pt, 4, color, slope_corr, out + o, QUAD_SLOTS);
    } else {
        for (int i = 0; i < QUAD_SLOTS; i++) out[o + i] = rec_nop();
    }
    o += QUAD_SLOTS;
    for (int i = 0; i < 2; i++) {
        if (flags & (i + 1)) {
            P2L v[6];
            int nv;
            cap_polygon(p0, th, sin_tab, v, nv);
            if (nv >= 2 && nv <= 5) emit_convex_aa(W, H, v, nv, color, slope_corr, out + o, CAP_SLOTS);
            else for (int i = 0; i < CAP_SLOTS; i++) out[o + i] = rec_nop();
            o += CAP_SLOTS;
        }
        p0 = p1;
    }
    return nrec;
}

}  // namespace rig
