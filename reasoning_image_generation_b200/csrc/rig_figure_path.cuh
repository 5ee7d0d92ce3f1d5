// rig_figure_path.cuh -- k_figure_path: the geometry half of RendererAgg::draw_path on the device (sm_100a).
//
// Input: the draw_path argument lists of a batch, as they arrive over the C ABI (struct rig_figure_scene, here with DEVICE
// pointers).  Output: what k_figure reads -- FLayer headers, FEdge arrays sorted by first row, slab directories.
//
// The arithmetic is rig_figure_geom.h, the same source the host pipeline (rig_figure_path.cpp) instantiates: affine, PathSnapper,
// curve3_div / curve4_div, vertex_sequence, vcgen_stroke / math_stroke, rasterizer_sl_clip, iround(v * 256).  Every operation
// in it is IEEE-exact on both sides (-fmad=false; double division and sqrt are correctly rounded on the device) except the
// libm calls of round joins / caps, which go through geom::Uncertain (see there): the kernel raises GEOM_UNCERTAIN when one of
// them decides a rounding, and the host then runs its own pipeline for the batch.
//
// Work decomposition
//   * short paths (fewer than GEOM_LONG_MIN_VERTICES vertices: lines, rectangles, small polygons): one THREAD per artist, 32 at a
//     time per warp.  The host hands the artists out sorted by path length, so the lanes of a warp run the same patch type and
//     diverge little.  A thread's stroke-vertex sequence and its unsorted edges live in thread-private global scratch.
//   * long paths (curve outlines): one WARP per artist -- see "warp-cooperative route" below: lanes over curve commands, vertices,
//     joins and contour segments, prefix sums in place of the serial order.  (The first version ran every artist on one thread:
//     the 32 longest outlines then kept one warp busy for 1.8 M cycles while the SMs idled.)
//   * after each pass (faces, then strokes) the WARP places its layers: ONE 64-bit atomicAdd reserves the edges and directory
//     entries, then per layer a stable counting sort by first row (shared-memory histogram, match_any ranks) moves the edges from
//     the scratch to their final place with coalesced 16-byte accesses and writes the slab directory; layers of at most 32 edges
//     are copied as they are (one chunk of k_figure's edge loop: their order changes nothing).
// Layer slots are fixed by the host (artist_layer[a] + 0 / 1), so image_layer_begin needs no device-wide scan; a pass that
// leaves no edge keeps its slot as an empty layer (ymin > ymax), which k_figure's band test skips.
#pragma once
#include <cuda_runtime.h>

#include "rig_figure_geom.h"

namespace rig {

constexpr int GEOM_THREADS = 128;           // per CTA
constexpr int GEOM_WARPS = GEOM_THREADS / 32;

// counters[] slots
// (GC_EDGES, GC_DIR) form one 64-bit word: a layer placement reserves both with a single atomic
constexpr int GC_WORK = 0, GC_WORK_LONG = 1, GC_EDGES = 2, GC_DIR = 3, GC_STATUS = 4, GC_BAD_ARTIST = 5, GC_BAD_WHAT = 6, GC_WORDS = 8;
constexpr uint32_t GEOM_LONG_MIN_VERTICES = 12;   // paths at least this long go one per warp
constexpr int GEOM_CMD_CAP = 1024;               // path commands the warp route parses (longer paths: serial on one lane)

struct GeomParams {
    rig_figure_scene s;              // device pointers
    const uint32_t *artist_order;    // [n_artists] artists in work order (longest paths first)
    uint32_t n_long;                 // the first n_long of them take the warp-cooperative route (one warp per artist)
    const uint32_t *artist_layer;    // [n_artists] first layer slot of the artist
    int W, H, bgr;
    FLayer *layers;
    FEdge *edges;
    uint32_t *dir;
    uint32_t edge_cap, dir_cap;
    uint32_t *counters;              // [GC_WORDS]
    geom::SV *seq;                   // [threads * seq_cap]
    FEdge *stage;                    // [threads * stage_cap]
    uint32_t seq_cap, stage_cap;
};

// Vertices a warp-route path may have for its arrays (vertices, lengths, offsets, contour) to live in SHARED memory: 72 bytes per
// vertex and warp; longer paths use the warp's global scratch.  Off by default: measured SLOWER (per 32 768 single-shape figures:
// 0: 0.86 ms, 128: 1.15 ms, 192: 1.25 ms -- outlines of 100-200 vertices overflow and flatten twice, and the larger CTAs cost more
// than the shorter round trips save; profiles/r2_notes.md).
#ifndef GEOM_SMEM_PTS
#define GEOM_SMEM_PTS 0
#endif
__host__ __device__ inline size_t geom_warp_smem_words(int H)
{
    return (size_t)(H + 34 > GEOM_CMD_CAP ? H + 34 : GEOM_CMD_CAP) + (size_t)GEOM_SMEM_PTS * 18;
}

struct DevSeq {
    geom::SV *p;
    int n, cap;
    bool overflow;
    __device__ __forceinline__ int size() const { return n; }
    __device__ __forceinline__ geom::SV &operator[](int i) { return p[i]; }
    __device__ __forceinline__ geom::SV &back() { return p[n - 1]; }
    __device__ __forceinline__ void push_back(const geom::SV &v)
    {
        if (n < cap) p[n++] = v;
        else overflow = true;
    }
    __device__ __forceinline__ void pop_back() { if (n > 0) n--; }
    __device__ __forceinline__ void clear() { n = 0; }
};

struct DevEdges {
    FEdge *p;
    uint32_t n, cap;
    int ymin, ymax, xmin, xmax, maxspan;
    bool overflow;
    __device__ __forceinline__ void reset()
    {
        n = 0; ymin = 1 << 30; ymax = -(1 << 30); xmin = 1 << 30; xmax = -(1 << 30); maxspan = 0; overflow = false;
    }
    __device__ __forceinline__ void push(int x1, int y1, int x2, int y2)
    {
        const int a = min(y1, y2) >> 8, b = max(y1, y2) >> 8;
        ymin = min(ymin, a); ymax = max(ymax, b);
        xmin = min(xmin, min(x1, x2) >> 8); xmax = max(xmax, max(x1, x2) >> 8);
        maxspan = max(maxspan, b - a);
        if (n < cap) { *(int4 *)(p + n) = make_int4(x1, y1, x2, y2); n++; }
        else overflow = true;
    }
};

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t &total)
{
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
}

// Places the layers the lanes of this warp staged (lane: n edges at `stage`, bounding cells in `E`) at slot `slot`.
// hist: >= H + 2 words of shared memory owned by the warp.
__device__ __forceinline__ void geom_place_layers(const GeomParams &P, int lane, bool have_layer, uint32_t slot, uint32_t rgba, const DevEdges &E,
                                                  uint32_t *hist, uint32_t &status)
{
    const uint32_t n = have_layer ? E.n : 0u;
    const uint32_t nd = n >= FIG_DIR_MIN_EDGES ? (uint32_t)((E.ymax - E.ymin) / FIG_SLAB_ROWS + 2) : 0u;   // (smaller layers: never read)
    uint32_t etot, dtot;
    const uint32_t eoff = warp_excl_scan(n, lane, etot), doff = warp_excl_scan(nd, lane, dtot);
    uint32_t ebase = 0, dbase = 0;
    if (lane == 0 && etot) {
        const unsigned long long got = atomicAdd((unsigned long long *)(P.counters + GC_EDGES), (unsigned long long)etot | ((unsigned long long)dtot << 32));
        ebase = (uint32_t)got;
        dbase = (uint32_t)(got >> 32);
    }
    ebase = __shfl_sync(0xffffffffu, ebase, 0);
    dbase = __shfl_sync(0xffffffffu, dbase, 0);
    const bool fits = (unsigned long long)ebase + etot <= P.edge_cap && (unsigned long long)dbase + dtot <= P.dir_cap;
    if (!fits) status |= geom::GEOM_ARENA_OVERFLOW;
    if (have_layer) {
        FLayer L;
        L.edge_begin = ebase + eoff; L.n_edges = fits ? n : 0u; L.rgba = rgba;
        const bool empty = n == 0 || !fits;
        L.ymin = (int16_t)(empty ? 32767 : max(E.ymin, 0));
        L.ymax = (int16_t)(empty ? -32768 : min(E.ymax, P.H - 1));
        L.xmin = (int16_t)(empty ? 0 : max(E.xmin, 0));
        L.xmax = (int16_t)(empty ? 0 : min(E.xmax, P.W - 1));
        L.dir_begin = dbase + doff;
        L.maxspan = (uint16_t)min(E.maxspan, 65535);
        L.ndir = (uint16_t)nd;
        L.pad = 0u;
        uint4 *dst = (uint4 *)(P.layers + slot);
        const uint4 *src = (const uint4 *)&L;
        dst[0] = src[0];
        dst[1] = src[1];
    }
    if (!fits) return;
    uint32_t todo = __ballot_sync(0xffffffffu, n != 0u);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t sn = __shfl_sync(0xffffffffu, n, src);
        const int symin = __shfl_sync(0xffffffffu, E.ymin, src), symax = __shfl_sync(0xffffffffu, E.ymax, src);
        const uint32_t se = __shfl_sync(0xffffffffu, ebase + eoff, src), sd = __shfl_sync(0xffffffffu, dbase + doff, src);
        const unsigned long long sp = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)E.p, src);
        const FEdge *from = (const FEdge *)(uintptr_t)sp;
        FEdge *to = P.edges + se;
        if (sn <= 32u) {
            // one chunk of k_figure's edge loop: the order inside it changes nothing (no sort, no directory: k_figure scans layers
            // of fewer than FIG_DIR_MIN_EDGES edges whole)
            if ((uint32_t)lane < sn) *(int4 *)(to + lane) = *(const int4 *)(from + lane);
            continue;
        }
        const int rows = symax - symin + 1;              // first rows lie in [0, rows)
        for (int r = lane; r <= rows; r += 32) hist[r] = 0u;
        __syncwarp();
        for (uint32_t e = lane; e < sn; e += 32) {
            const int4 v = *(const int4 *)(from + e);
            atomicAdd(hist + ((min(v.y, v.w) >> 8) - symin) + 1, 1u);
        }
        __syncwarp();
        // inclusive scan: hist[r] = edges whose first row < r
        uint32_t carry = 0;
        for (int r0 = 0; r0 <= rows; r0 += 32) {
            const int r = r0 + lane;
            uint32_t v = r <= rows ? hist[r] : 0u, inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (r <= rows) hist[r] = carry + inc;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        __syncwarp();
        if (sn >= FIG_DIR_MIN_EDGES) {
            const int slabs = (symax - symin) / FIG_SLAB_ROWS + 2;
            for (int s = lane; s < slabs; s += 32) P.dir[sd + s] = hist[min(s * FIG_SLAB_ROWS, rows)];
        }
        __syncwarp();
        for (uint32_t e0 = 0; e0 < sn; e0 += 32) {
            const uint32_t e = e0 + lane;
            const bool on = e < sn;
            int4 v = make_int4(0, 0, 0, 0);
            int key = -1 - lane;                          // distinct dummies for the idle lanes
            if (on) {
                v = *(const int4 *)(from + e);
                key = (min(v.y, v.w) >> 8) - symin;
            }
            const uint32_t same = __match_any_sync(0xffffffffu, key);
            const uint32_t before = same & ((1u << lane) - 1u);
            uint32_t pos = 0;
            if (on) pos = hist[key] + __popc(before);
            __syncwarp();
            if (on && (same >> lane) == 1u) hist[key] += __popc(same);      // the group's last lane moves the cursor on
            __syncwarp();
            if (on) *(int4 *)(to + pos) = v;
        }
        __syncwarp();
    }
}


// ---- warp-cooperative route: one artist per warp ---------------------------------------------------------------------------
// A stroker is sequential only in how it is written down: once the flattened, de-duplicated vertex list exists, every join
// depends on three neighbouring vertices, every contour segment on two neighbouring contour points, and the clipper keeps
// no state but the previous point.  So, for one path:
//   parse     (uniform) the walk of walk_flat over the path codes -> a command list in shared memory
//   flatten   lane = command: curve3_div / curve4_div twice (count, then store at the prefix-sum offset) -> P[0..N)
//   measure   lane = vertex: distance to the next vertex (vertex_sequence); a coincident pair anywhere sends the path to the
//             serial code on lane 0 (agg's de-duplication is order dependent)
//   joins     lane = join: math_stroke::calc_join / calc_cap with a sink that keeps the <= 2 points; prefix sum -> contour C[]
//   clip      lane = contour segment: rasterizer_sl_clip on (C[k], C[k+1]); prefix sum -> the warp's edge staging
// The values are those of the serial code (same functions, same operands); the order of the edges is the emission order.
// Only miter / bevel joins and butt / projecting caps take this route (round ones emit a variable number of libm points).
struct WarpScratch {
    double2 *P;            // flattened vertices
    double *d, *nx, *ny;   // per vertex: distance to the next one, half-width offsets of that segment
    double2 *C;            // contour points
    double2 *slices;       // 32 areas of slice_cap points: where the lanes flatten their curves before the points are concatenated
    uint32_t cap_pts, cap_contour, slice_cap;
    FEdge *stage;          // the warp's 32 staging areas, contiguous
    uint32_t stage_cap;
};

struct NoSeq {              // the stroker's vertex storage, for a StrokeGen that only evaluates single joins / caps
    geom::SV dummy;
    __device__ __forceinline__ int size() const { return 0; }
    __device__ __forceinline__ geom::SV &operator[](int) { return dummy; }
    __device__ __forceinline__ geom::SV &back() { return dummy; }
    __device__ __forceinline__ void push_back(const geom::SV &) {}
    __device__ __forceinline__ void pop_back() {}
    __device__ __forceinline__ void clear() {}
};
struct TwoPoints {           // sink of one join / cap
    double x[2], y[2];
    int n;
    bool overflow;
    __device__ __forceinline__ void put(double px, double py)
    {
        if (n < 2) { x[n] = px; y[n] = py; n++; }
        else overflow = true;
    }
    __device__ __forceinline__ void move_to(double px, double py) { put(px, py); }
    __device__ __forceinline__ void line_to(double px, double py) { put(px, py); }
    __device__ __forceinline__ void close() {}
    __device__ __forceinline__ double cx1() const { return 0; }
    __device__ __forceinline__ double cy1() const { return 0; }
    __device__ __forceinline__ double cx2() const { return 0; }
    __device__ __forceinline__ double cy2() const { return 0; }
};
struct LaneEdges {           // sink of one clipped contour segment (3 spans, each split at most once)
    int4 e[6];
    int n;
    __device__ __forceinline__ void push(int x1, int y1, int x2, int y2)
    {
        if (n < 6) e[n++] = make_int4(x1, y1, x2, y2);
    }
};

// parse: the command list of walk_flat's walk over the path codes, 32 vertices at a time.  Returns the number of commands, or -1
// when the path does not have the one shape this route handles: MOVETO, then LINETO / CURVE3 / CURVE4 commands whose control
// vertices carry the curve's own code (runs of CURVE4 in threes, of CURVE3 in twos), optionally one closing CLOSEPOLY at the very
// end; vertices after a STOP are ignored as walk_flat ignores them.  Anything else takes the serial code on lane 0.
__device__ __forceinline__ int geom_warp_parse(const geom::PathView &path, uint32_t *cmd, int &closed, int lane)
{
    const uint32_t n = path.n;
    closed = 0;
    if (n < 2 || n > 0xffffffu || path.codes[0] != RIG_PATH_MOVETO) return -1;
    int ncmd = 0;
    uint32_t carry4 = 0, carry3 = 0;                 // length of the curve run that ends at the previous chunk's last vertex
    bool ok = true;
    for (uint32_t b = 1; b < n; b += 32) {
        const uint32_t v = b + lane;
        const uint32_t c = v < n ? path.codes[v] : (uint32_t)RIG_PATH_STOP;
        const uint32_t cnext = v + 1 < n ? path.codes[v + 1] : (uint32_t)RIG_PATH_STOP;
        const uint32_t mstop = __ballot_sync(0xffffffffu, c == RIG_PATH_STOP);
        const uint32_t live_mask = mstop ? ((1u << (__ffs(mstop) - 1)) - 1u) : 0xffffffffu;     // lanes before the first STOP
        const bool live = (live_mask >> lane) & 1u;
        const uint32_t m4 = __ballot_sync(0xffffffffu, c == RIG_PATH_CURVE4), m3 = __ballot_sync(0xffffffffu, c == RIG_PATH_CURVE3);
        const uint32_t lt = (1u << lane) - 1u;
        bool start = false, bad = false;
        uint32_t type = 0;
        if (live) {
            if (c == RIG_PATH_LINETO) { start = true; type = 2u; }
            else if (c == RIG_PATH_CURVE4 || c == RIG_PATH_CURVE3) {
                const bool four = c == RIG_PATH_CURVE4;
                const uint32_t below = ~(four ? m4 : m3) & lt;                  // lanes before this one that are not in the run
                const uint32_t pos = below ? (uint32_t)lane - (32u - (uint32_t)__clz(below)) : (uint32_t)lane + (four ? carry4 : carry3);
                const uint32_t period = four ? 3u : 2u;
                start = pos % period == 0u;
                type = four ? 4u : 3u;
                if (cnext != c && (pos + 1u) % period != 0u) bad = true;         // the run ends in the middle of a curve
            } else if (c == RIG_PATH_CLOSEPOLY) {
                if (cnext != RIG_PATH_STOP) bad = true;                           // only as the path's last command
            } else bad = true;                                                    // a second MOVETO, other end-poly flags, unknown codes
        }
        const uint32_t mclose = __ballot_sync(0xffffffffu, live && c == RIG_PATH_CLOSEPOLY);
        if (mclose) closed = 1;
        const uint32_t mstart = __ballot_sync(0xffffffffu, start);
        if (__any_sync(0xffffffffu, bad) || ncmd + __popc(mstart) > GEOM_CMD_CAP) { ok = false; break; }
        if (start) cmd[ncmd + __popc(mstart & lt)] = v | (type << 24);
        ncmd += __popc(mstart);
        // run lengths carried into the next chunk (from lane 31's point of view)
        {
            const uint32_t below4 = ~m4, below3 = ~m3;
            carry4 = (m4 >> 31) ? (below4 ? 31u - (31u - (uint32_t)__clz(below4)) : 32u + carry4) : 0u;
            carry3 = (m3 >> 31) ? (below3 ? 31u - (31u - (uint32_t)__clz(below3)) : 32u + carry3) : 0u;
        }
        if (mstop || mclose) break;
    }
    __syncwarp();
    return ok ? ncmd : -1;
}

// flatten: P[0..N); returns N, -1 when a curve overflows its slice (status set) or -2 when the points do not fit ws.P.  Every lane subdivides its curve once, into its own
// slice of the (still unused) contour area; the slices are then moved to their place in P[], command after command.
__device__ __forceinline__ int geom_warp_flatten(const geom::PathView &path, const uint32_t *cmd, int ncmd, int lane, WarpScratch &ws, uint32_t &status)
{
    if (lane == 0) {
        const geom::Node m = path.node(0);
        ws.P[0] = make_double2(m.x, m.y);
    }
    const uint32_t slice_cap = ws.slice_cap;
    double2 *slice = ws.slices + (size_t)lane * slice_cap;
    uint32_t base = 1;
    for (int c0 = 0; c0 < ncmd; c0 += 32) {
        const int j = c0 + lane;
        uint32_t cnt = 0;
        double2 single = make_double2(0, 0);
        bool over = false;
        if (j < ncmd) {
            const uint32_t vi = cmd[j] & 0xffffffu, type = cmd[j] >> 24;
            const geom::Node a = path.node(vi);
            if (type == 2u) { cnt = 1; single = make_double2(a.x, a.y); }
            else {
                const geom::Node st = path.node(vi - 1), b = path.node(vi + 1);
                struct SliceEmit {
                    double2 *p;
                    uint32_t n, cap;
                    __device__ __forceinline__ void operator()(double x, double y)
                    {
                        if (n < cap) p[n] = make_double2(x, y);
                        n++;
                    }
                } se{slice, 0u, slice_cap};
                if (type == 3u) { geom::quad_split(se, st.x, st.y, a.x, a.y, b.x, b.y); se(b.x, b.y); }
                else {
                    const geom::Node c = path.node(vi + 2);
                    geom::cubic_split(se, st.x, st.y, a.x, a.y, b.x, b.y, c.x, c.y);
                    se(c.x, c.y);
                }
                cnt = se.n;
                over = cnt > slice_cap;
            }
        }
        uint32_t total;
        const uint32_t off = warp_excl_scan(cnt, lane, total);
        if (__any_sync(0xffffffffu, over)) { status |= geom::GEOM_SEQ_OVERFLOW; return -1; }
        if (base + total > ws.cap_pts) return -2;                      // does not fit these arrays (the caller may have larger ones)
        __syncwarp();
        if (cnt == 1u && j < ncmd && (cmd[j] >> 24) == 2u) ws.P[base + off] = single;
        uint32_t curves = __ballot_sync(0xffffffffu, j < ncmd && (cmd[j] >> 24) != 2u);
        while (curves) {
            const int src = __ffs(curves) - 1;
            curves &= curves - 1;
            const uint32_t scnt = __shfl_sync(0xffffffffu, cnt, src), soff = __shfl_sync(0xffffffffu, off, src);
            const double2 *from = ws.slices + (size_t)src * slice_cap;
            for (uint32_t k = lane; k < scnt; k += 32) ws.P[base + soff + k] = from[k];
        }
        __syncwarp();
        base += total;
    }
    __syncwarp();
    return (int)base;
}

// clip: the closed contour C[0..M) -> edges appended to the warp's staging (E: lane-private bounding cells, n_edges: uniform count)
__device__ __forceinline__ void geom_warp_clip_contour(const double2 *C, int M, const geom::ArtistSetup &art, int lane, WarpScratch &ws, DevEdges &E,
                                                       uint32_t &n_edges, uint32_t &status)
{
    if (M < 2) return;                                    // a lone move_to draws nothing
    for (int k0 = 0; k0 < M; k0 += 32) {
        const int k = k0 + lane;
        LaneEdges le;
        le.n = 0;
        if (k < M) {
            const double2 p = C[k], q = C[k + 1 < M ? k + 1 : 0];
            geom::ClipToCells<LaneEdges> clip(le, art.bx1, art.by1, art.bx2, art.by2);
            clip.move_to(p.x, p.y);
            clip.line_to(q.x, q.y);
        }
        uint32_t total;
        const uint32_t off = warp_excl_scan((uint32_t)le.n, lane, total);
        if (n_edges + total > 32u * ws.stage_cap) { status |= geom::GEOM_EDGE_OVERFLOW; return; }
        for (int t = 0; t < le.n; t++) {
            const int4 v = le.e[t];
            const int a = min(v.y, v.w) >> 8, b = max(v.y, v.w) >> 8;
            E.ymin = min(E.ymin, a); E.ymax = max(E.ymax, b);
            E.xmin = min(E.xmin, min(v.x, v.z) >> 8); E.xmax = max(E.xmax, max(v.x, v.z) >> 8);
            E.maxspan = max(E.maxspan, b - a);
            *(int4 *)(ws.stage + n_edges + off + t) = v;
        }
        n_edges += total;
    }
}

// lane 0 ends up owning the layer: E.n = edges staged by the warp, bounding cells reduced over the lanes
__device__ __forceinline__ void geom_warp_reduce_bounds(DevEdges &E, uint32_t n_edges, int lane, const WarpScratch &ws)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        E.ymin = min(E.ymin, __shfl_xor_sync(0xffffffffu, E.ymin, o)); E.ymax = max(E.ymax, __shfl_xor_sync(0xffffffffu, E.ymax, o));
        E.xmin = min(E.xmin, __shfl_xor_sync(0xffffffffu, E.xmin, o)); E.xmax = max(E.xmax, __shfl_xor_sync(0xffffffffu, E.xmax, o));
        E.maxspan = max(E.maxspan, __shfl_xor_sync(0xffffffffu, E.maxspan, o));
    }
    E.n = lane == 0 ? n_edges : 0u;
    E.p = ws.stage;
    __syncwarp();
}

// stroke of the flattened path P[0..N): returns false if the path must take the serial code (coincident vertices, too few)
__device__ __forceinline__ bool geom_warp_stroke(int N, int closed, const geom::ArtistSetup &art, int cap_style, int join_style, int lane, WarpScratch &ws,
                                                 DevEdges &E, uint32_t &n_edges, uint32_t &status)
{
    // ---- measure (vertex_sequence): d[i] = |P[i+1] - P[i]|, i < N - 1
    bool coincident = false;
    for (int i = lane; i < N - 1; i += 32) {
        const double2 a = ws.P[i], b = ws.P[i + 1];
        const double dd = sqrt(geom::dist2(a.x, a.y, b.x, b.y));
        ws.d[i] = dd;
        if (!(dd > geom::kVertexDistEps)) coincident = true;
    }
    if (__any_sync(0xffffffffu, coincident)) return false;
    if (closed) {
        // the closing vertex usually repeats the first one: agg drops it (and then looks at the new last vertex)
        double2 f = ws.P[0], l = ws.P[N - 1];
        double dd = sqrt(geom::dist2(l.x, l.y, f.x, f.y));
        if (!(dd > geom::kVertexDistEps)) {
            N -= 1;
            if (N < 2) return false;
            l = ws.P[N - 1];
            dd = sqrt(geom::dist2(l.x, l.y, f.x, f.y));
            if (!(dd > geom::kVertexDistEps)) return false;
        }
        if (lane == 0) ws.d[N - 1] = dd;
    }
    if (N < 3) closed = 0;
    if (N < 2 + (closed ? 1 : 0)) return true;             // nothing to draw (the serial code returns here too)
    __syncwarp();
    // ---- half-width offsets per segment
    const double w = art.lw * 0.5;
    const int nseg = closed ? N : N - 1;
    for (int i = lane; i < nseg; i += 32) {
        const double2 c = ws.P[i], q = ws.P[i + 1 < N ? i + 1 : 0];
        const double dd = ws.d[i];
        ws.nx[i] = w * (q.y - c.y) / dd;
        ws.ny[i] = w * (q.x - c.x) / dd;
    }
    __syncwarp();
    // ---- joins and caps, in contour order.  Item t of a contour -> (kind, vertex):
    //   closed: contour 0 = joins 0 .. N-1 forward, contour 1 = joins N-1 .. 0 backward
    //   open:   one contour = cap(0), joins 1 .. N-2 forward, cap(N-1), joins N-2 .. 1 backward
    NoSeq noseq;
    geom::Certain certain;
    const int n_contours = closed ? 2 : 1;
    const int items = closed ? N : 2 + 2 * (N - 2);
    if ((uint32_t)(2 * items) > ws.cap_contour) { status |= geom::GEOM_SEQ_OVERFLOW; return true; }
    for (int ct = 0; ct < n_contours; ct++) {
        uint32_t M = 0;
        for (int t0 = 0; t0 < items; t0 += 32) {
            const int t = t0 + lane;
            TwoPoints tp;
            tp.n = 0; tp.overflow = false;
            if (t < items) {
                geom::StrokeGen<NoSeq, TwoPoints, geom::Certain> gen(noseq, art.lw, cap_style, join_style, tp, certain);
                int i;                // vertex
                int kind;             // 0 forward join, 1 backward join, 2 cap at 0, 3 cap at N-1
                if (closed) { kind = ct; i = ct == 0 ? t : N - 1 - t; }
                else if (t == 0) { kind = 2; i = 0; }
                else if (t <= N - 2) { kind = 0; i = t; }
                else if (t == N - 1) { kind = 3; i = N - 1; }
                else { kind = 1; i = N - 2 - (t - N); }
                if (kind >= 2) {
                    const int o = kind == 2 ? 1 : N - 2;                  // the neighbour the cap looks at
                    const double2 a = ws.P[i], b = ws.P[o];
                    const geom::SV v0{a.x, a.y, 0, 0, 0}, v1{b.x, b.y, 0, 0, 0};
                    gen.cap(v0, v1, ws.d[kind == 2 ? 0 : N - 2]);
                } else {
                    const int ip = i ? i - 1 : N - 1, iq = i + 1 < N ? i + 1 : 0;
                    const double2 a = ws.P[ip], b = ws.P[i], c = ws.P[iq];
                    const geom::SV vp{a.x, a.y, 0, 0, 0}, vc{b.x, b.y, 0, 0, 0}, vq{c.x, c.y, 0, 0, 0};
                    const double pnx = ws.nx[ip], pny = ws.ny[ip], cnx = ws.nx[i], cny = ws.ny[i], pd = ws.d[ip], cd = ws.d[i];
                    if (kind == 0) gen.join(vp, vc, vq, pnx, pny, cnx, cny, pd, cd);
                    else gen.join(vq, vc, vp, 0.0 - cnx, 0.0 - cny, 0.0 - pnx, 0.0 - pny, cd, pd);
                }
            }
            if (tp.overflow) status |= geom::GEOM_SEQ_OVERFLOW;
            uint32_t total;
            const uint32_t off = warp_excl_scan((uint32_t)tp.n, lane, total);
            for (int u = 0; u < tp.n; u++) ws.C[M + off + u] = make_double2(tp.x[u], tp.y[u]);
            M += total;
        }
        __syncwarp();
        geom_warp_clip_contour(ws.C, (int)M, art, lane, ws, E, n_edges, status);
        __syncwarp();
    }
    return true;
}

// fill of the flattened path: the path itself is the contour (the rasteriser closes it)
__device__ __forceinline__ void geom_warp_face(int N, const geom::ArtistSetup &art, int lane, WarpScratch &ws, DevEdges &E, uint32_t &n_edges, uint32_t &status)
{
    geom_warp_clip_contour(ws.P, N, art, lane, ws, E, n_edges, status);
    __syncwarp();
}

#ifndef GEOM_MIN_CTAS
#define GEOM_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(GEOM_THREADS, GEOM_MIN_CTAS) k_figure_path(const GeomParams P)
{
    extern __shared__ uint32_t geom_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *hist = geom_smem + (size_t)warp * geom_warp_smem_words(P.H);
    const size_t tid = (size_t)blockIdx.x * GEOM_THREADS + threadIdx.x;
    DevSeq seq;
    seq.p = P.seq + tid * P.seq_cap; seq.n = 0; seq.cap = (int)P.seq_cap; seq.overflow = false;
    DevEdges E;
    E.p = P.stage + tid * P.stage_cap; E.cap = P.stage_cap;
    E.reset();
    uint32_t status = 0;
    const uint32_t n_art = (uint32_t)P.s.n_artists;
    // ---- long paths: one per warp
    {
        WarpScratch wg, wsm;                                             // arrays in global scratch / in shared memory
        const size_t wtid = tid - lane;                                  // the warp's first thread
        uint8_t *sb = (uint8_t *)(P.seq + wtid * P.seq_cap);             // 32 * seq_cap * 40 bytes
        wg.cap_pts = P.seq_cap * 8u;
        wg.cap_contour = P.seq_cap * 16u;
        wg.slice_cap = P.seq_cap / 2u;
        wg.P = (double2 *)sb;                                            // 16 B * cap_pts
        wg.d = (double *)(sb + (size_t)16 * wg.cap_pts);                 // 3 * 8 B * cap_pts
        wg.nx = wg.d + wg.cap_pts;
        wg.ny = wg.nx + wg.cap_pts;
        wg.C = (double2 *)(sb + (size_t)40 * wg.cap_pts);                // 16 B * cap_contour
        wg.slices = wg.C + wg.cap_contour;                               // 16 B * 32 * slice_cap   (40 * 8 + 16 * 16 + 16 * 16 = 832 <= 1280 B per seq_cap)
        wg.stage = P.stage + wtid * P.stage_cap;
        wg.stage_cap = P.stage_cap;
        wsm = wg;
        if (GEOM_SMEM_PTS > 0) {
            uint8_t *ss = (uint8_t *)(hist + (P.H + 34 > GEOM_CMD_CAP ? P.H + 34 : GEOM_CMD_CAP));
            wsm.cap_pts = GEOM_SMEM_PTS;
            wsm.cap_contour = 2u * GEOM_SMEM_PTS;
            wsm.P = (double2 *)ss;
            wsm.d = (double *)(ss + (size_t)16 * GEOM_SMEM_PTS);
            wsm.nx = wsm.d + GEOM_SMEM_PTS;
            wsm.ny = wsm.nx + GEOM_SMEM_PTS;
            wsm.C = (double2 *)(ss + (size_t)40 * GEOM_SMEM_PTS);
        }
        uint32_t *cmd = hist;
        for (;;) {
            uint32_t k = 0;
            if (lane == 0) k = atomicAdd(P.counters + GC_WORK_LONG, 1u);
            k = __shfl_sync(0xffffffffu, k, 0);
            if (k >= P.n_long) break;
            const uint32_t a = P.artist_order[k];
            geom::ArtistSetup art;
            geom::artist_setup_begin(P.s, a, P.H, art);
            const int cap_style = P.s.cap[a], join_style = P.s.join[a];
            const uint32_t slot = P.artist_layer[a];
            // the scan, one lane per vertex
            int bad = 0;
            bool nr = false;
            for (uint32_t i = lane; i < art.path.n; i += 32) {
                const geom::Node nd = art.path.raw(i);
                const geom::Node prev = i ? art.path.raw(i - 1) : geom::Node{0.0, 0.0, 0};
                int b1;
                bool n1;
                geom::artist_scan_vertex(art.path, i, prev, nd, b1, n1);
                if (b1 && !bad) bad = b1;
                nr |= n1;
            }
            const bool any_bad = __any_sync(0xffffffffu, bad != 0);
            const bool rectilinear = !__any_sync(0xffffffffu, nr);
            int what = 0;
            uint32_t badbits = any_bad ? geom::GEOM_BAD_INPUT : geom::artist_setup_end(P.W, P.H, rectilinear, art, what);
            bool run = true;
            if (badbits) {
                art.face = art.has_face && art.face_alpha != 0;
                art.stroke = art.edge_alpha != 0 && art.lw != 0.0;
                status |= badbits;
                if (lane == 0 && atomicCAS(P.counters + GC_BAD_ARTIST, 0xffffffffu, a) == 0xffffffffu) P.counters[GC_BAD_WHAT] = (uint32_t)(what ? what : 2);
                run = false;
            }
            int closed = 0, N = -1;
            bool parallel = false, use_smem = false;
            if (run) {
                const int ncmd = geom_warp_parse(art.path, cmd, closed, lane);
                __syncwarp();
                if (ncmd >= 0) {
                    N = geom_warp_flatten(art.path, cmd, ncmd, lane, wsm, status);
                    use_smem = N >= 0;
                    if (N == -2) N = geom_warp_flatten(art.path, cmd, ncmd, lane, wg, status);
                    if (N == -2) status |= geom::GEOM_SEQ_OVERFLOW;
                    parallel = N >= 0;
                    if (N < 0) run = false;                              // scratch overflow: status set, the host takes the batch
                }
            }
            WarpScratch &ws = use_smem ? wsm : wg;
            // ---- pass 0: face
            E.reset();
            uint32_t n_edges = 0;
            if (run && art.face) {
                if (parallel) geom_warp_face(N, art, lane, ws, E, n_edges, status);
                else if (lane == 0) { geom::face_pass(art, E); n_edges = E.n; }
            }
            if (E.overflow) status |= geom::GEOM_EDGE_OVERFLOW;
            if (parallel) geom_warp_reduce_bounds(E, n_edges, lane, ws);
            if (art.face) {
                geom_place_layers(P, lane, lane == 0, slot, geom::pack_rgba(P.s.face_rgba + 4 * (size_t)a, P.bgr != 0), E, hist, status);
                __syncwarp();
            }
            // ---- pass 1: stroke
            E.p = P.stage + tid * P.stage_cap;
            E.reset();
            n_edges = 0;
            if (run && art.stroke) {
                bool done = false;
                if (parallel && join_style != RIG_JOIN_ROUND && cap_style != RIG_CAP_ROUND) {
                    // (parse wrote the command list over hist; place_layers above reused it, flatten's output lies behind it)
                    done = geom_warp_stroke(N, closed, art, cap_style, join_style, lane, ws, E, n_edges, status);
                    if (done) geom_warp_reduce_bounds(E, n_edges, lane, ws);
                }
                if (!done) {
                    E.p = P.stage + tid * P.stage_cap;
                    E.reset();
                    if (lane == 0) {
                        geom::Uncertain unc;
                        seq.overflow = false;
                        geom::stroke_pass(art, cap_style, join_style, seq, E, unc);
                        if (unc.hit) status |= geom::GEOM_UNCERTAIN;
                        if (seq.overflow) status |= geom::GEOM_SEQ_OVERFLOW;
                    }
                }
            }
            if (E.overflow) status |= geom::GEOM_EDGE_OVERFLOW;
            if (art.stroke) {
                geom_place_layers(P, lane, lane == 0, slot + (art.face ? 1u : 0u), geom::pack_rgba(P.s.edge_rgba + 4 * (size_t)a, P.bgr != 0), E, hist, status);
                __syncwarp();
            }
            E.p = P.stage + tid * P.stage_cap;
        }
    }
    // ---- short paths: one per thread, 32 at a time
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = P.n_long + atomicAdd(P.counters + GC_WORK, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_art) break;
        const bool have = base + lane < n_art;
        const uint32_t a = have ? P.artist_order[base + lane] : 0u;
        geom::ArtistSetup art;
        art.face = art.stroke = false;
        uint32_t slot = 0;
        bool run = have;
        if (have) {
            int what = 0;
            const uint32_t bad = geom::artist_setup(P.s, a, P.W, P.H, art, what);
            slot = P.artist_layer[a];
            if (bad) {
                // the slots the host counted from the flags stay empty (the placement below is warp-wide, so this lane goes
                // through it like the others); the host pipeline, which runs next, names the error
                art.face = art.has_face && art.face_alpha != 0;
                art.stroke = art.edge_alpha != 0 && art.lw != 0.0;
                status |= bad;
                if (atomicCAS(P.counters + GC_BAD_ARTIST, 0xffffffffu, a) == 0xffffffffu) P.counters[GC_BAD_WHAT] = (uint32_t)what;
                run = false;
            }
        }
        // ---- pass 0: faces
        E.reset();
        if (run && art.face) geom::face_pass(art, E);
        if (E.overflow) status |= geom::GEOM_EDGE_OVERFLOW;
        geom_place_layers(P, lane, have && art.face, slot, have ? geom::pack_rgba(P.s.face_rgba + 4 * (size_t)a, P.bgr != 0) : 0u, E, hist, status);
        // ---- pass 1: strokes
        E.reset();
        if (run && art.stroke) {
            geom::Uncertain unc;
            seq.overflow = false;
            geom::stroke_pass(art, P.s.cap[a], P.s.join[a], seq, E, unc);
            if (unc.hit) status |= geom::GEOM_UNCERTAIN;
            if (seq.overflow) status |= geom::GEOM_SEQ_OVERFLOW;
        }
        if (E.overflow) status |= geom::GEOM_EDGE_OVERFLOW;
        geom_place_layers(P, lane, have && art.stroke, slot + (art.face ? 1u : 0u), have ? geom::pack_rgba(P.s.edge_rgba + 4 * (size_t)a, P.bgr != 0) : 0u, E,
                          hist, status);
    }
    if (status) atomicOr(P.counters + GC_STATUS, status);
}

// per-image scheduling weight for k_figure's costliest-first order: edges + 8 x the rows its layers resolve.  `cost` and
// `counters_host` are page-locked HOST memory (unified addressing): the kernel that follows k_figure_path delivers its results
// without a copy-engine transfer.
__global__ void k_figure_cost(const uint32_t *image_layer_begin, const FLayer *layers, uint32_t n_images, float *cost, const uint32_t *counters,
                              uint32_t *counters_host)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < GC_WORDS) counters_host[i] = counters[i];
    if (i >= n_images) return;
    float w = 0.f;
    for (uint32_t l = image_layer_begin[i]; l < image_layer_begin[i + 1]; l++) {
        const FLayer L = layers[l];
        w += (float)L.n_edges + 8.f * (float)max(0, (int)L.ymax - (int)L.ymin + 1);
    }
    cost[i] = w;
}

}  // namespace rig
