// rig_figure_geom.h -- the geometry half of RendererAgg::draw_path, host + device: one draw_path argument list -> polygon
// edges in 24.8 fixed point.
//
// What matplotlib's Agg backend does between RendererAgg::draw_path and the cell rasteriser, restated as a push pipeline
// (matplotlib and Agg are third-party dependencies of the reference that are absent from /root/reference; call sites:
// multigraph_generation/generator.py:250-259):
//
//   PathView      trans *= scale(1, -1); trans *= translate(0, H); vertices through the affine; PathSnapper
//   flatten       agg::conv_curve (curve3_div / curve4_div, approximation_scale 1, angle_tolerance 0)
//   StrokeGen     agg::conv_stroke: vertex_sequence de-duplication, vcgen_stroke contour order, math_stroke joins and caps
//   ClipToCells   agg::rasterizer_sl_clip<ras_conv_dbl>::line_to and the auto-closing of rasterizer_scanline_aa; emits
//                 the integer end points cells.line() would get (horizontal edges add nothing to any cell and are dropped)
//
// The same source is compiled twice: by the host compiler for rig_figure_path.cpp (std::vector storage, the host libm,
// -ffp-contract=off) and by nvcc for k_figure_path (rig_figure_path.cuh: one thread per rasteriser pass, fixed-capacity
// storage, -fmad=false).  Addition, subtraction, multiplication, division, sqrt and floor are IEEE-exact on both sides, so
// both produce the same doubles and hence the same integers.  The only calls that are NOT bit-reproducible across the two
// libms are atan2 / acos / cos / sin in round joins and round caps: every value that depends on them goes through
// `Uncertain`, which notes when a rounding or a comparison is decided by less than the two libms' error bounds could
// move it; the device then hands the batch back to the host pipeline (rig_figure_abi.cu), so the result is the host's
// either way.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/rig_figure.h"
#include "rig_figure_core.h"

// The large routines are real functions on the device (one copy each): inlined at every call site the kernel is several MB of
// SASS and does not finish compiling.
#if defined(__CUDACC__)
#define RIG_HD_NI __host__ __device__ __noinline__
#else
#define RIG_HD_NI inline
#endif

namespace rig {
namespace geom {

constexpr double kPi = 3.14159265358979323846;
constexpr double kVertexDistEps = 1e-14;
constexpr double kIntersectionEps = 1.0e-30;
constexpr double kCollinearityEps = 1e-30;
constexpr int kCurveRecursionLimit = 32;

// status bits of a pass (0: fine)
constexpr uint32_t GEOM_SEQ_OVERFLOW = 1u;      // more stroke vertices than the fixed storage holds
constexpr uint32_t GEOM_EDGE_OVERFLOW = 2u;     // more edges than the fixed staging holds
constexpr uint32_t GEOM_ARENA_OVERFLOW = 4u;    // the shared edge arena is full (the caller grows it and runs again)
constexpr uint32_t GEOM_UNCERTAIN = 8u;         // a libm-dependent value sits on a decision boundary
constexpr uint32_t GEOM_BAD_INPUT = 16u;        // unknown path code / non-finite vertex / bad line width: the host pipeline names it

RIG_HD double dist2(double ax, double ay, double bx, double by)
{
    const double dx = bx - ax, dy = by - ay;
    return dx * dx + dy * dy;
}
RIG_HD int iround(double v) { return (int)((v < 0.0) ? v - 0.5 : v + 0.5); }

// Tracks values that came out of atan2 / acos / cos / sin.  The host instance is empty (the host IS the definition).
// Error model: both libms are within 2 ulp on these calls; the angles are < 8 in magnitude and go through at most a few
// dozen additions, the products with the half width and the sums with a canvas coordinate (|v| <= 4096, ulp 9e-13) keep the
// absolute error of a coordinate below 1e-11 px, i.e. 3e-9 after the multiplication by 256 that precedes the rounding.
// A libm coordinate is either rounded directly (the segment lies inside the clip box) or feeds the clipper's intersection
// formulas; the second case is not analysed: a pass that emitted a libm point AND clipped any segment is handed back.
struct Uncertain {
    bool hit = false;
    bool libm = false;
    RIG_HD void coordinate(double v, double c1, double c2)
    {
        const double s = v * 256.0, f = s - floor(s);
        libm = true;
        if (fabs(f - 0.5) < 3e-8 || fabs(v - c1) < 1e-8 || fabs(v - c2) < 1e-8) hit = true;      // 10x the bound above
    }
    RIG_HD void quotient(double q)                 // truncated to an int
    {
        const double f = q - floor(q);
        if (f < 1e-9 || f > 1.0 - 1e-9) hit = true;
    }
    RIG_HD void ordering(double a, double b)       // compared with < / >
    {
        if (fabs(a - b) < 1e-12) hit = true;
    }
    RIG_HD void pass_done(bool clipped)
    {
        if (libm && clipped) hit = true;
    }
};
struct Certain {
    static constexpr bool hit = false;
    RIG_HD void coordinate(double, double, double) {}
    RIG_HD void quotient(double) {}
    RIG_HD void ordering(double, double) {}
    RIG_HD void pass_done(bool) {}
};

// ---- last stage: clip in doubles, round to 24.8, collect edges ------------------------------------
// EdgeSink: void push(int x1, int y1, int x2, int y2)
template <class EdgeSink>
class ClipToCells {
public:
    RIG_HD ClipToCells(EdgeSink &edges, double x1, double y1, double x2, double y2) : edges_(edges), cx1_(x1), cy1_(y1), cx2_(x2), cy2_(y2) {}
    RIG_HD void move_to(double x, double y)
    {
        close();
        sx_ = px_ = x;
        sy_ = py_ = y;
        pf_ = flags(x, y);
        open_ = false;
        pi_ok_ = false;
    }
    RIG_HD void line_to(double x2, double y2)
    {
        open_ = true;
        const unsigned f2 = flags(x2, y2);
        if ((pf_ | f2) == 0u) {                                  // both ends inside the box (nearly every segment): no clipping,
            const int ix2 = iround(x2 * 256), iy2 = iround(y2 * 256);      // and the start was rounded as the previous end
            const int ix1 = pi_ok_ ? ipx_ : iround(px_ * 256), iy1 = pi_ok_ ? ipy_ : iround(py_ * 256);
            cells_line(ix1, iy1, ix2, iy2);
            ipx_ = ix2; ipy_ = iy2; pi_ok_ = true;
            px_ = x2; py_ = y2; pf_ = f2;
            return;
        }
        line_to_clipped(x2, y2, f2);
    }
    RIG_HD_NI void line_to_clipped(double x2, double y2, unsigned f2)
    {
        pi_ok_ = false;
        clipped_ = true;
        if ((pf_ & 10u) == (f2 & 10u) && (pf_ & 10u) != 0u) {   // both above or both below the box
            px_ = x2; py_ = y2; pf_ = f2;
            return;
        }
        const double x1 = px_, y1 = py_;
        const unsigned f1 = pf_;
        double y3, y4;
        unsigned f3, f4;
        switch (((f1 & 5u) << 1) | (f2 & 5u)) {
        case 0: span(x1, y1, x2, y2, f1, f2); break;
        case 1:
            y3 = y1 + (cx2_ - x1) * (y2 - y1) / (x2 - x1); f3 = flags_y(y3);
            span(x1, y1, cx2_, y3, f1, f3);
            span(cx2_, y3, cx2_, y2, f3, f2);
            break;
        case 2:
            y3 = y1 + (cx2_ - x1) * (y2 - y1) / (x2 - x1); f3 = flags_y(y3);
            span(cx2_, y1, cx2_, y3, f1, f3);
            span(cx2_, y3, x2, y2, f3, f2);
            break;
        case 3: span(cx2_, y1, cx2_, y2, f1, f2); break;
        case 4:
            y3 = y1 + (cx1_ - x1) * (y2 - y1) / (x2 - x1); f3 = flags_y(y3);
            span(x1, y1, cx1_, y3, f1, f3);
            span(cx1_, y3, cx1_, y2, f3, f2);
            break;
        case 6:
            y3 = y1 + (cx2_ - x1) * (y2 - y1) / (x2 - x1);
            y4 = y1 + (cx1_ - x1) * (y2 - y1) / (x2 - x1);
            f3 = flags_y(y3); f4 = flags_y(y4);
            span(cx2_, y1, cx2_, y3, f1, f3);
            span(cx2_, y3, cx1_, y4, f3, f4);
            span(cx1_, y4, cx1_, y2, f4, f2);
            break;
        case 8:
            y3 = y1 + (cx1_ - x1) * (y2 - y1) / (x2 - x1); f3 = flags_y(y3);
            span(cx1_, y1, cx1_, y3, f1, f3);
            span(cx1_, y3, x2, y2, f3, f2);
            break;
        case 9:
            y3 = y1 + (cx1_ - x1) * (y2 - y1) / (x2 - x1);
            y4 = y1 + (cx2_ - x1) * (y2 - y1) / (x2 - x1);
            f3 = flags_y(y3); f4 = flags_y(y4);
            span(cx1_, y1, cx1_, y3, f1, f3);
            span(cx1_, y3, cx2_, y4, f3, f4);
            span(cx2_, y4, cx2_, y2, f4, f2);
            break;
        case 12: span(cx1_, y1, cx1_, y2, f1, f2); break;
        default: break;
        }
        pf_ = f2; px_ = x2; py_ = y2;
    }
    RIG_HD void close()
    {
        if (open_) {
            line_to(sx_, sy_);
            open_ = false;
        }
    }
    RIG_HD double cx1() const { return cx1_; }
    RIG_HD double cy1() const { return cy1_; }
    RIG_HD double cx2() const { return cx2_; }
    RIG_HD double cy2() const { return cy2_; }
    RIG_HD bool clipped() const { return clipped_; }      // some segment had an end outside the box

private:
    RIG_HD unsigned flags(double x, double y) const
    {
        return (unsigned)(x > cx2_) | ((unsigned)(y > cy2_) << 1) | ((unsigned)(x < cx1_) << 2) | ((unsigned)(y < cy1_) << 3);
    }
    RIG_HD unsigned flags_y(double y) const { return ((unsigned)(y > cy2_) << 1) | ((unsigned)(y < cy1_) << 3); }
    RIG_HD void cells_line(int x1, int y1, int x2, int y2)
    {
        if (y1 == y2) return;                       // adds nothing to any cell
        const long long dx = (long long)x2 - (long long)x1, lim = 16384LL << 8;
        if (dx >= lim || dx <= -lim) {
            // agg splits very long edges in the middle, recursively; the clip box keeps |dx| below 2 * lim (canvases are
            // smaller than 32768 px), so one split is all that can ever happen
            const int mx = (int)(((long long)x1 + x2) >> 1), my = (int)(((long long)y1 + y2) >> 1);
            if (y1 != my) edges_.push(x1, y1, mx, my);
            if (my != y2) edges_.push(mx, my, x2, y2);
            return;
        }
        edges_.push(x1, y1, x2, y2);
    }
    RIG_HD void span(double x1, double y1, double x2, double y2, unsigned f1, unsigned f2)
    {
        f1 &= 10u;
        f2 &= 10u;
        if ((f1 | f2) == 0u) {
            cells_line(iround(x1 * 256), iround(y1 * 256), iround(x2 * 256), iround(y2 * 256));
            return;
        }
        if (f1 == f2) return;
        double tx1 = x1, ty1 = y1, tx2 = x2, ty2 = y2;
        if (f1 & 8u) { tx1 = x1 + (cy1_ - y1) * (x2 - x1) / (y2 - y1); ty1 = cy1_; }
        if (f1 & 2u) { tx1 = x1 + (cy2_ - y1) * (x2 - x1) / (y2 - y1); ty1 = cy2_; }
        if (f2 & 8u) { tx2 = x1 + (cy1_ - y1) * (x2 - x1) / (y2 - y1); ty2 = cy1_; }
        if (f2 & 2u) { tx2 = x1 + (cy2_ - y1) * (x2 - x1) / (y2 - y1); ty2 = cy2_; }
        cells_line(iround(tx1 * 256), iround(ty1 * 256), iround(tx2 * 256), iround(ty2 * 256));
    }
    EdgeSink &edges_;
    double cx1_, cy1_, cx2_, cy2_;
    double px_ = 0, py_ = 0, sx_ = 0, sy_ = 0;
    unsigned pf_ = 0;
    int ipx_ = 0, ipy_ = 0;      // the previous point in 24.8 fixed point, valid if pi_ok_
    bool pi_ok_ = false;
    bool open_ = false;
    bool clipped_ = false;
};

// ---- stroker --------------------------------------------------------------------------------------
struct SV {
    double x, y, d;   // d: distance to the next vertex of the sequence
    double nx, ny;    // half-width offset of the segment to the next vertex: w (y' - y) / d, w (x' - x) / d (set by finish())
};

// Seq: int size(); SV &operator[](int); SV &back(); void push_back(SV); void pop_back(); void clear()
template <class Seq, class Sink, class Unc>
class StrokeGen {
public:
    RIG_HD StrokeGen(Seq &storage, double linewidth, int cap, int join, Sink &sink, Unc &unc) : sink_(sink), seq_(storage), unc_(unc), cap_(cap), join_(join)
    {
        w_ = linewidth * 0.5;
        if (w_ < 0) { wabs_ = -w_; wsign_ = -1; }
        else { wabs_ = w_; wsign_ = 1; }
        weps_ = w_ / 1024.0;
        miter_limit_ = linewidth;     // _draw_path: stroke.miter_limit(points_to_pixels(gc.linewidth))
    }
    RIG_HD void begin(double x, double y)
    {
        seq_.clear();
        push(x, y);
    }
    RIG_HD void add(double x, double y) { push(x, y); }
    // finishes the sub-path: closed = it ended with CLOSEPOLY
    RIG_HD_NI void finish(bool closed)
    {
        settle(closed);
        const int n = (int)seq_.size();
        if (n < 3) closed = false;
        if (n < 2 + (closed ? 1 : 0)) return;
        // math_stroke::calc_join derives both segments' offsets at every join, and every segment is visited by four joins (two per
        // side); they are computed once here.  On the way back the segments run the other way: IEEE subtraction, multiplication
        // and division are odd functions, so the offsets there are these negated (0.0 - v, which keeps a zero positive as the
        // direct evaluation does).
        const int nseg = closed ? n : n - 1;
        for (int i = 0; i < nseg; i++) {
            SV &c = seq_[i];
            const SV &q = seq_[i + 1 < n ? i + 1 : 0];
            c.nx = w_ * (q.y - c.y) / c.d;
            c.ny = w_ * (q.x - c.x) / c.d;
        }
        first_ = true;
        if (closed) {
            for (int i = 0; i < n; i++) {
                const SV &p = seq_[i ? i - 1 : n - 1], &c = seq_[i], &q = seq_[i + 1 < n ? i + 1 : 0];
                join(p, c, q, p.nx, p.ny, c.nx, c.ny, p.d, c.d);
            }
            sink_.close();
            first_ = true;
            for (int i = n - 1; i >= 0; i--) {
                const SV &p = seq_[i ? i - 1 : n - 1], &c = seq_[i], &q = seq_[i + 1 < n ? i + 1 : 0];
                join(q, c, p, 0.0 - c.nx, 0.0 - c.ny, 0.0 - p.nx, 0.0 - p.ny, c.d, p.d);
            }
            sink_.close();
        } else {
            cap(seq_[0], seq_[1], seq_[0].d);
            for (int i = 1; i + 1 < n; i++) join(seq_[i - 1], seq_[i], seq_[i + 1], seq_[i - 1].nx, seq_[i - 1].ny, seq_[i].nx, seq_[i].ny, seq_[i - 1].d, seq_[i].d);
            cap(seq_[n - 1], seq_[n - 2], seq_[n - 2].d);
            for (int i = n - 2; i >= 1; i--)
                join(seq_[i + 1], seq_[i], seq_[i - 1], 0.0 - seq_[i].nx, 0.0 - seq_[i].ny, 0.0 - seq_[i - 1].nx, 0.0 - seq_[i - 1].ny, seq_[i].d, seq_[i - 1].d);
            sink_.close();
        }
    }

private:
    // agg::vertex_sequence<vertex_dist>: a vertex is dropped once its predecessor turns out to coincide with it
    RIG_HD static bool measure(SV &a, const SV &b)
    {
        a.d = sqrt(dist2(a.x, a.y, b.x, b.y));
        const bool apart = a.d > kVertexDistEps;
        if (!apart) a.d = 1.0 / kVertexDistEps;
        return apart;
    }
    RIG_HD void push(double x, double y)
    {
        const int n = (int)seq_.size();
        if (n > 1 && !measure(seq_[n - 2], seq_[n - 1])) seq_.pop_back();
        seq_.push_back(SV{x, y, 0.0, 0.0, 0.0});
    }
    RIG_HD_NI void settle(bool closed)
    {
        while (seq_.size() > 1) {
            const int n = (int)seq_.size();
            if (measure(seq_[n - 2], seq_[n - 1])) break;
            const SV t = seq_[n - 1];
            seq_.pop_back();
            seq_.pop_back();
            push(t.x, t.y);
        }
        if (closed)
            while (seq_.size() > 1) {
                if (measure(seq_.back(), seq_[0])) break;
                seq_.pop_back();
            }
    }
    RIG_HD void out(double x, double y)
    {
        if (first_) sink_.move_to(x, y);
        else sink_.line_to(x, y);
        first_ = false;
    }
    // a point whose coordinates came out of cos / sin
    RIG_HD void out_libm(double x, double y)
    {
        unc_.coordinate(x, sink_.cx1(), sink_.cx2());
        unc_.coordinate(y, sink_.cy1(), sink_.cy2());
        out(x, y);
    }
    RIG_HD static double cross(double x1, double y1, double x2, double y2, double x, double y) { return (x - x2) * (y2 - y1) - (y - y2) * (x2 - x1); }
    RIG_HD static bool intersect(double ax, double ay, double bx, double by, double cx, double cy, double dx, double dy, double &x, double &y)
    {
        const double num = (ay - cy) * (dx - cx) - (ax - cx) * (dy - cy);
        const double den = (bx - ax) * (dy - cy) - (by - ay) * (dx - cx);
        if (fabs(den) < kIntersectionEps) return false;
        const double r = num / den;
        x = ax + r * (bx - ax);
        y = ay + r * (by - ay);
        return true;
    }
    RIG_HD_NI void arc(double x, double y, double dx1, double dy1, double dx2, double dy2)
    {
        double a1 = atan2(dy1 * wsign_, dx1 * wsign_), a2 = atan2(dy2 * wsign_, dx2 * wsign_);
        double da = acos(wabs_ / (wabs_ + 0.125 / 1.0)) * 2;
        out(x + dx1, y + dy1);
        unc_.ordering(a1, a2);
        if (wsign_ > 0) {
            if (a1 > a2) a2 += 2 * kPi;
            unc_.quotient((a2 - a1) / da);
            const int n = (int)((a2 - a1) / da);
            da = (a2 - a1) / (n + 1);
            a1 += da;
            for (int i = 0; i < n; i++) {
                out_libm(x + cos(a1) * w_, y + sin(a1) * w_);
                a1 += da;
            }
        } else {
            if (a1 < a2) a2 -= 2 * kPi;
            unc_.quotient((a1 - a2) / da);
            const int n = (int)((a1 - a2) / da);
            da = (a1 - a2) / (n + 1);
            a1 -= da;
            for (int i = 0; i < n; i++) {
                out_libm(x + cos(a1) * w_, y + sin(a1) * w_);
                a1 -= da;
            }
        }
        out(x + dx2, y + dy2);
    }
    // revert = true: matplotlib's 'miter' (agg::miter_join_revert) and the inner miter
    RIG_HD void miter(const SV &v0, const SV &v1, const SV &v2, double dx1, double dy1, double dx2, double dy2, double mlimit)
    {
        double xi = v1.x, yi = v1.y;
        const double lim = wabs_ * mlimit;
        bool exceeded = true;
        if (intersect(v0.x + dx1, v0.y - dy1, v1.x + dx1, v1.y - dy1, v1.x + dx2, v1.y - dy2, v2.x + dx2, v2.y - dy2, xi, yi)) {
            const double di = sqrt(dist2(v1.x, v1.y, xi, yi));
            if (di <= lim) {
                out(xi, yi);
                exceeded = false;
            }
        } else {
            const double x2 = v1.x + dx1, y2 = v1.y - dy1;
            if ((cross(v0.x, v0.y, v1.x, v1.y, x2, y2) < 0.0) == (cross(v1.x, v1.y, v2.x, v2.y, x2, y2) < 0.0)) {
                out(v1.x + dx1, v1.y - dy1);
                exceeded = false;
            }
        }
        if (exceeded) {
            out(v1.x + dx1, v1.y - dy1);
            out(v1.x + dx2, v1.y - dy2);
        }
    }
public:
    // One join / one cap on their own (the warp-cooperative device route evaluates them one per lane).
    // (dx1, dy1), (dx2, dy2): w (v1.y - v0.y) / len1, w (v1.x - v0.x) / len1 and the same for v1 -> v2
    RIG_HD_NI void join(const SV &v0, const SV &v1, const SV &v2, double dx1, double dy1, double dx2, double dy2, double len1, double len2)
    {
        const double cp = cross(v0.x, v0.y, v1.x, v1.y, v2.x, v2.y);
        if ((cp > kVertexDistEps && w_ > 0) || (cp < -kVertexDistEps && w_ < 0)) {
            double limit = ((len1 < len2) ? len1 : len2) / wabs_;
            if (limit < 1.01) limit = 1.01;            // inner_miter_limit
            miter(v0, v1, v2, dx1, dy1, dx2, dy2, limit);
            return;
        }
        double dx = (dx1 + dx2) / 2, dy = (dy1 + dy2) / 2;
        const double dbevel = sqrt(dx * dx + dy * dy);
        if (join_ == RIG_JOIN_ROUND || join_ == RIG_JOIN_BEVEL) {
            if (1.0 * (wabs_ - dbevel) < weps_) {     // nearly collinear: one point
                if (intersect(v0.x + dx1, v0.y - dy1, v1.x + dx1, v1.y - dy1, v1.x + dx2, v1.y - dy2, v2.x + dx2, v2.y - dy2, dx, dy))
                    out(dx, dy);
                else
                    out(v1.x + dx1, v1.y - dy1);
                return;
            }
        }
        if (join_ == RIG_JOIN_MITER) miter(v0, v1, v2, dx1, dy1, dx2, dy2, miter_limit_);
        else if (join_ == RIG_JOIN_ROUND) arc(v1.x, v1.y, dx1, -dy1, dx2, -dy2);
        else {
            out(v1.x + dx1, v1.y - dy1);
            out(v1.x + dx2, v1.y - dy2);
        }
    }
    RIG_HD_NI void cap(const SV &v0, const SV &v1, double len)
    {
        double dx1 = (v1.y - v0.y) / len, dy1 = (v1.x - v0.x) / len, dx2 = 0, dy2 = 0;
        dx1 *= w_;
        dy1 *= w_;
        if (cap_ != RIG_CAP_ROUND) {
            if (cap_ == RIG_CAP_PROJECTING) {
                dx2 = dy1 * wsign_;
                dy2 = dx1 * wsign_;
            }
            out(v0.x - dx1 - dx2, v0.y + dy1 - dy2);
            out(v0.x + dx1 - dx2, v0.y - dy1 - dy2);
            return;
        }
        double da = acos(wabs_ / (wabs_ + 0.125 / 1.0)) * 2;
        unc_.quotient(kPi / da);
        const int n = (int)(kPi / da);
        da = kPi / (n + 1);
        out(v0.x - dx1, v0.y + dy1);
        if (wsign_ > 0) {
            double a1 = atan2(dy1, -dx1);
            a1 += da;
            for (int i = 0; i < n; i++) {
                out_libm(v0.x + cos(a1) * w_, v0.y + sin(a1) * w_);
                a1 += da;
            }
        } else {
            double a1 = atan2(-dy1, dx1);
            a1 -= da;
            for (int i = 0; i < n; i++) {
                out_libm(v0.x + cos(a1) * w_, v0.y + sin(a1) * w_);
                a1 -= da;
            }
        }
        out(v0.x + dx1, v0.y - dy1);
    }
    RIG_HD double half_width() const { return w_; }

private:
    Sink &sink_;
    Seq &seq_;        // caller-owned, so that its capacity survives from artist to artist
    Unc &unc_;
    double w_, wabs_, weps_, miter_limit_;
    int wsign_, cap_, join_;
    bool first_ = true;
};

// ---- curve flattening -------------------------------------------------------------------------------
// agg's curve3_div / curve4_div recurse (left half, then right half); the same order with an explicit stack of the pending
// right halves, so that the device needs no call stack.
struct QuadTask { double x1, y1, x2, y2, x3, y3; int level; };
struct CubicTask { double x1, y1, x2, y2, x3, y3, x4, y4; int level; };

template <class Emit>
RIG_HD_NI void quad_split(Emit &emit, double x1, double y1, double x2, double y2, double x3, double y3)
{
    QuadTask stack[kCurveRecursionLimit + 2];
    int sp = 0;
    int level = 0;
    for (;;) {
        bool done = false;
        if (level > kCurveRecursionLimit) done = true;
        double x12 = 0, y12 = 0, x23 = 0, y23 = 0, x123 = 0, y123 = 0;
        if (!done) {
            x12 = (x1 + x2) / 2; y12 = (y1 + y2) / 2; x23 = (x2 + x3) / 2; y23 = (y2 + y3) / 2;
            x123 = (x12 + x23) / 2; y123 = (y12 + y23) / 2;
            const double dx = x3 - x1, dy = y3 - y1;
            double d = fabs(((x2 - x3) * dy - (y2 - y3) * dx));
            if (d > kCollinearityEps) {
                if (d * d <= 0.25 * (dx * dx + dy * dy)) { emit(x123, y123); done = true; }
            } else {
                const double da = dx * dx + dy * dy;
                bool inside = false;
                if (da == 0) d = dist2(x1, y1, x2, y2);
                else {
                    d = ((x2 - x1) * dx + (y2 - y1) * dy) / da;
                    if (d > 0 && d < 1) inside = true;
                    else if (d <= 0) d = dist2(x2, y2, x1, y1);
                    else if (d >= 1) d = dist2(x2, y2, x3, y3);
                    else d = dist2(x2, y2, x1 + d * dx, y1 + d * dy);
                }
                if (inside) done = true;
                else if (d < 0.25) { emit(x2, y2); done = true; }
            }
        }
        if (done) {
            if (sp == 0) return;
            const QuadTask &t = stack[--sp];
            x1 = t.x1; y1 = t.y1; x2 = t.x2; y2 = t.y2; x3 = t.x3; y3 = t.y3; level = t.level;
            continue;
        }
        stack[sp++] = QuadTask{x123, y123, x23, y23, x3, y3, level + 1};
        x2 = x12; y2 = y12; x3 = x123; y3 = y123;
        level++;
    }
}

template <class Emit>
RIG_HD_NI void cubic_split(Emit &emit, double x1, double y1, double x2, double y2, double x3, double y3, double x4, double y4)
{
    CubicTask stack[kCurveRecursionLimit + 2];
    int sp = 0;
    int level = 0;
    for (;;) {
        bool done = false;
        if (level > kCurveRecursionLimit) done = true;
        double x12 = 0, y12 = 0, x23 = 0, y23 = 0, x34 = 0, y34 = 0, x123 = 0, y123 = 0, x234 = 0, y234 = 0, x1234 = 0, y1234 = 0;
        if (!done) {
            x12 = (x1 + x2) / 2; y12 = (y1 + y2) / 2; x23 = (x2 + x3) / 2; y23 = (y2 + y3) / 2;
            x34 = (x3 + x4) / 2; y34 = (y3 + y4) / 2;
            x123 = (x12 + x23) / 2; y123 = (y12 + y23) / 2; x234 = (x23 + x34) / 2; y234 = (y23 + y34) / 2;
            x1234 = (x123 + x234) / 2; y1234 = (y123 + y234) / 2;
            const double dx = x4 - x1, dy = y4 - y1;
            double d2 = fabs(((x2 - x4) * dy - (y2 - y4) * dx));
            double d3 = fabs(((x3 - x4) * dy - (y3 - y4) * dx));
            const int sel = ((d2 > kCollinearityEps) ? 2 : 0) + ((d3 > kCollinearityEps) ? 1 : 0);
            if (sel == 0) {
                double k = dx * dx + dy * dy;
                bool inside = false;
                if (k == 0) {
                    d2 = dist2(x1, y1, x2, y2);
                    d3 = dist2(x4, y4, x3, y3);
                } else {
                    k = 1 / k;
                    double da1 = x2 - x1, da2 = y2 - y1;
                    d2 = k * (da1 * dx + da2 * dy);
                    da1 = x3 - x1; da2 = y3 - y1;
                    d3 = k * (da1 * dx + da2 * dy);
                    if (d2 > 0 && d2 < 1 && d3 > 0 && d3 < 1) inside = true;
                    else {
                        if (d2 <= 0) d2 = dist2(x2, y2, x1, y1);
                        else if (d2 >= 1) d2 = dist2(x2, y2, x4, y4);
                        else d2 = dist2(x2, y2, x1 + d2 * dx, y1 + d2 * dy);
                        if (d3 <= 0) d3 = dist2(x3, y3, x1, y1);
                        else if (d3 >= 1) d3 = dist2(x3, y3, x4, y4);
                        else d3 = dist2(x3, y3, x1 + d3 * dx, y1 + d3 * dy);
                    }
                }
                if (inside) done = true;
                else if (d2 > d3) {
                    if (d2 < 0.25) { emit(x2, y2); done = true; }
                } else {
                    if (d3 < 0.25) { emit(x3, y3); done = true; }
                }
            } else {
                const double dd = sel == 1 ? d3 : (sel == 2 ? d2 : d2 + d3);
                if (dd * dd <= 0.25 * (dx * dx + dy * dy)) { emit(x23, y23); done = true; }
            }
        }
        if (done) {
            if (sp == 0) return;
            const CubicTask &t = stack[--sp];
            x1 = t.x1; y1 = t.y1; x2 = t.x2; y2 = t.y2; x3 = t.x3; y3 = t.y3; x4 = t.x4; y4 = t.y4; level = t.level;
            continue;
        }
        stack[sp++] = CubicTask{x1234, y1234, x234, y234, x34, y34, x4, y4, level + 1};
        x2 = x12; y2 = y12; x3 = x123; y3 = y123; x4 = x1234; y4 = y1234;
        level++;
    }
}

// ---- first stage: affine, snapping ------------------------------------------------------------------
struct Affine {
    double sx, shy, shx, sy, tx, ty;
    RIG_HD void then(const Affine &m)   // agg::trans_affine::multiply
    {
        const double t0 = sx * m.sx + shy * m.shx, t2 = shx * m.sx + sy * m.shx, t4 = tx * m.sx + ty * m.shx + m.tx;
        shy = sx * m.shy + shy * m.sy;
        sy = shx * m.shy + sy * m.sy;
        ty = tx * m.shy + ty * m.sy + m.ty;
        sx = t0; shx = t2; tx = t4;
    }
    RIG_HD void apply(double &x, double &y) const
    {
        const double t = x;
        x = t * sx + y * shx + tx;
        y = t * shy + y * sy + ty;
    }
};
RIG_HD bool is_vertex_code(int c) { return c >= RIG_PATH_MOVETO && c <= RIG_PATH_CURVE4; }

struct Node {
    double x, y;
    int code;
};

// One artist's path as the converters see it: vertex i through the device transform, then the snapper.  Nothing is stored:
// node(i) is a pure function of the borrowed arrays, so the face pass and the stroke pass (and the host and the device) read
// the same doubles.
struct PathView {
    const double *v;          // 2 * n doubles
    const uint8_t *codes;     // n
    uint32_t n;
    Affine tr;
    bool snap;
    double snap_add;
    RIG_HD Node raw(uint32_t i) const
    {
        Node nd{v[2 * (size_t)i], v[2 * (size_t)i + 1], codes[i]};
        if (is_vertex_code(nd.code)) tr.apply(nd.x, nd.y);
        return nd;
    }
    RIG_HD Node node(uint32_t i) const
    {
        Node nd = raw(i);
        if (snap && is_vertex_code(nd.code)) {
            nd.x = floor(nd.x + 0.5) + snap_add;
            nd.y = floor(nd.y + 0.5) + snap_add;
        }
        return nd;
    }
};

// everything draw_path derives from one artist before the converters run; returns 0 or GEOM_BAD_INPUT (bad_what: 1 unknown
// code, 2 non-finite vertex, 3 bad line width)
struct ArtistSetup {
    PathView path;
    double bx1, by1, bx2, by2;    // clip box of the rasteriser
    double lw;
    bool face, stroke;
    // the artist's scalar attributes, read by artist_setup_begin (all its loads are issued together, before the scan)
    double clip_in[4];
    uint8_t snap_mode, edge_alpha, face_alpha, has_face, has_clip;
};
// artist_setup in three steps, so that the device can run the scan over the vertices with one lane per vertex:
//   artist_setup_begin   the path view (borrowed arrays + device transform)
//   artist_scan_vertex   what PathSnapper's scan and the input validation learn from vertex i (and its predecessor)
//   artist_setup_end     snapping decision, clip box, which passes the artist asks for
RIG_HD void artist_setup_begin(const rig_figure_scene &s, uint32_t a, int H, ArtistSetup &o)
{
    const uint32_t vb = s.artist_vertex_begin[a], ve = s.artist_vertex_begin[a + 1];
    const double *af = s.affine + 6 * (size_t)a;
    Affine tr{af[0], af[1], af[2], af[3], af[4], af[5]};
    tr.then(Affine{1.0, 0.0, 0.0, -1.0, 0.0, 0.0});
    tr.then(Affine{1.0, 0.0, 0.0, 1.0, 0.0, (double)H});
    PathView &p = o.path;
    p.v = s.vertices + 2 * (size_t)vb;
    p.codes = s.codes + vb;
    p.n = ve - vb;
    p.tr = tr;
    p.snap = false;
    p.snap_add = 0.0;
    o.lw = s.linewidth[a];
    o.snap_mode = s.snap[a];
    o.edge_alpha = s.edge_rgba[4 * (size_t)a + 3];
    o.face_alpha = s.face_rgba[4 * (size_t)a + 3];
    o.has_face = s.has_face[a];
    o.has_clip = (s.has_clip && s.clip) ? s.has_clip[a] : (uint8_t)0;
    if (o.has_clip) {
        const double *c = s.clip + 4 * (size_t)a;
        o.clip_in[0] = c[0]; o.clip_in[1] = c[1]; o.clip_in[2] = c[2]; o.clip_in[3] = c[3];
    }
}
// bad: 0 fine, 1 unknown code, 2 non-finite vertex; not_rectilinear: the vertex rules out snapping in auto mode
RIG_HD void artist_scan_vertex(const PathView &p, uint32_t i, const Node &prev, const Node &nd, int &bad, bool &not_rectilinear)
{
    (void)p;
    bad = 0;
    not_rectilinear = false;
    if (!is_vertex_code(nd.code) && nd.code != RIG_PATH_STOP && nd.code != RIG_PATH_CLOSEPOLY) { bad = 1; return; }
    if (!(isfinite(nd.x) && isfinite(nd.y))) { bad = 2; return; }
    if (i > 0) {
        if (nd.code == RIG_PATH_CURVE3 || nd.code == RIG_PATH_CURVE4) not_rectilinear = true;
        else if (nd.code == RIG_PATH_LINETO && !(fabs(prev.x - nd.x) < 1e-4 || fabs(prev.y - nd.y) < 1e-4)) not_rectilinear = true;
    }
}
RIG_HD uint32_t artist_setup_end(int W, int H, bool rectilinear, ArtistSetup &o, int &bad_what)
{
    PathView &p = o.path;
    const double lw = o.lw;
    if (!(lw >= 0.0) || !isfinite(lw)) { bad_what = 3; return GEOM_BAD_INPUT; }
    bool snap = o.snap_mode == RIG_SNAP_TRUE;
    if (o.snap_mode == RIG_SNAP_AUTO && p.n <= 1024 && p.n != 0) snap = rectilinear;
    if (snap) {
        const double slw = o.edge_alpha == 0 ? 0.0 : lw;
        const int r = (int)(slw + ((slw >= 0.0) ? 0.5 : -0.5));
        p.snap = true;
        p.snap_add = (r % 2) ? 0.5 : 0.0;
    }
    // clip box of the rasteriser (RendererAgg::set_clipbox)
    o.bx1 = 0; o.by1 = 0; o.bx2 = W; o.by2 = H;
    if (o.has_clip) {
        const double *c = o.clip_in;
        int x1 = (int)floor(c[0] + 0.5), y1 = (int)floor((double)H - c[1] + 0.5);
        int x2 = (int)floor(c[2] + 0.5), y2 = (int)floor((double)H - c[3] + 0.5);
        if (x1 < 0) x1 = 0;
        if (y1 < 0) y1 = 0;
        if (x2 > W) x2 = W;
        if (y2 > H) y2 = H;
        if (x1 > x2) { const int t = x1; x1 = x2; x2 = t; }
        if (y1 > y2) { const int t = y1; y1 = y2; y2 = t; }
        o.bx1 = x1; o.by1 = y1; o.bx2 = x2; o.by2 = y2;
    }
    o.face = o.has_face && o.face_alpha != 0;
    o.stroke = lw != 0.0 && o.edge_alpha != 0;
    return 0;
}
RIG_HD_NI uint32_t artist_setup(const rig_figure_scene &s, uint32_t a, int W, int H, ArtistSetup &o, int &bad_what)
{
    artist_setup_begin(s, a, H, o);
    const PathView &p = o.path;
    bad_what = 0;
    // PathSnapper's scan (auto mode: only rectilinear paths of at most 1024 vertices are snapped) doubles as input validation
    bool rectilinear = true;
    Node prev{0.0, 0.0, 0};
    for (uint32_t i = 0; i < p.n; i++) {
        const Node nd = p.raw(i);
        int bad;
        bool nr;
        artist_scan_vertex(p, i, prev, nd, bad, nr);
        if (bad) { bad_what = bad; return GEOM_BAD_INPUT; }
        if (nr) rectilinear = false;
        prev = nd;
    }
    return artist_setup_end(W, H, rectilinear, o, bad_what);
}

// flattened sub-paths of one artist -> fill sink or stroker
template <class OnMove, class OnLine, class OnClose>
RIG_HD_NI void walk_flat(const PathView &path, OnMove &on_move, OnLine &on_line, OnClose &on_close)
{
    double lx = 0, ly = 0;
    const uint32_t n = path.n;
    for (uint32_t i = 0; i < n; i++) {
        const Node v = path.node(i);
        if (v.code == RIG_PATH_STOP) break;
        if (v.code == RIG_PATH_CURVE3 && i + 1 < n) {
            const Node e = path.node(++i);
            quad_split(on_line, lx, ly, v.x, v.y, e.x, e.y);
            on_line(e.x, e.y);
            lx = e.x; ly = e.y;
        } else if (v.code == RIG_PATH_CURVE4 && i + 2 < n) {
            const Node c2 = path.node(++i);
            const Node e = path.node(++i);
            cubic_split(on_line, lx, ly, v.x, v.y, c2.x, c2.y, e.x, e.y);
            on_line(e.x, e.y);
            lx = e.x; ly = e.y;
        } else if (v.code == RIG_PATH_MOVETO) {
            on_move(v.x, v.y);
            lx = v.x; ly = v.y;
        } else if (v.code == RIG_PATH_LINETO) {
            on_line(v.x, v.y);
            lx = v.x; ly = v.y;
        } else if ((v.code & 0x0f) == 0x0f) {
            on_close((v.code & 0x40) != 0);
        }
    }
}

// the face pass of an artist: the flattened path straight into the clipper
template <class EdgeSink>
RIG_HD void face_pass(const ArtistSetup &a, EdgeSink &edges)
{
    ClipToCells<EdgeSink> sink(edges, a.bx1, a.by1, a.bx2, a.by2);
    struct Move { ClipToCells<EdgeSink> &s; RIG_HD void operator()(double x, double y) { s.move_to(x, y); } } mv{sink};
    struct Line { ClipToCells<EdgeSink> &s; RIG_HD void operator()(double x, double y) { s.line_to(x, y); } } ln{sink};
    struct Close { ClipToCells<EdgeSink> &s; RIG_HD void operator()(bool closed) { if (closed) s.close(); } } cl{sink};
    walk_flat(a.path, mv, ln, cl);
    sink.close();
}

// the stroke pass: conv_adaptor_vcgen restarts the generator at every MOVETO / end of polygon
template <class Seq, class EdgeSink, class Unc>
RIG_HD void stroke_pass(const ArtistSetup &a, int cap, int join, Seq &seq, EdgeSink &edges, Unc &unc)
{
    typedef ClipToCells<EdgeSink> Clip;
    typedef StrokeGen<Seq, Clip, Unc> Gen;
    Clip sink(edges, a.bx1, a.by1, a.bx2, a.by2);
    Gen gen(seq, a.lw, cap, join, sink, unc);
    struct State { Gen &gen; bool have; double mx, my; } st{gen, false, 0.0, 0.0};
    struct Move {
        State &s;
        RIG_HD void operator()(double x, double y)
        {
            if (s.have) s.gen.finish(false);
            s.mx = x; s.my = y;
            s.gen.begin(x, y);
            s.have = true;
        }
    } mv{st};
    struct Line {
        State &s;
        RIG_HD void operator()(double x, double y)
        {
            if (!s.have) { s.mx = x; s.my = y; s.gen.begin(x, y); s.have = true; }
            else s.gen.add(x, y);
        }
    } ln{st};
    struct Close {
        State &s;
        RIG_HD void operator()(bool closed)
        {
            if (s.have) s.gen.finish(closed);
            s.gen.begin(s.mx, s.my);
            s.have = true;
        }
    } cl{st};
    walk_flat(a.path, mv, ln, cl);
    if (st.have) gen.finish(false);
    unc.pass_done(sink.clipped());
}

RIG_HD uint32_t pack_rgba(const uint8_t *c, bool bgr)
{
    return bgr ? ((uint32_t)c[2] | ((uint32_t)c[1] << 8) | ((uint32_t)c[0] << 16) | ((uint32_t)c[3] << 24))
               : ((uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16) | ((uint32_t)c[3] << 24));
}

}  // namespace geom
}  // namespace rig
