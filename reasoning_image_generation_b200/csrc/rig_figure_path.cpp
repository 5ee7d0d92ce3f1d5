// rig_figure_path.cpp -- host instance of the figure path's geometry pipeline (rig_figure_geom.h): artists -> rasteriser layers.
//
// The pipeline itself (transform, PathSnapper, curve flattening, Agg's stroker, the rasteriser's clipper) lives in
// rig_figure_geom.h and is shared with the device kernel k_figure_path; this file runs it on host threads with std::vector
// storage and the host libm (compiled with -ffp-contract=off: the doubles are the ones the CPU renderer computes) and sorts
// each layer's edges.  It is the definition the device pipeline is held to, the fallback for batches the device hands back
// (see `Uncertain` in rig_figure_geom.h) and what rig_figure_debug_edges exposes to the tests.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "rig_figure_geom.h"
#include "rig_figure_host.h"

namespace rig {
namespace {

using geom::SV;

struct VecSeq {
    std::vector<SV> &v;
    int size() const { return (int)v.size(); }
    SV &operator[](int i) { return v[(size_t)i]; }
    SV &back() { return v.back(); }
    void push_back(const SV &x) { v.push_back(x); }
    void pop_back() { v.pop_back(); }
    void clear() { v.clear(); }
};
struct VecEdges {
    std::vector<FEdge> &e;
    void push(int x1, int y1, int x2, int y2) { e.push_back(FEdge{x1, y1, x2, y2}); }
};

void finish_layer(FigurePrepared &out, size_t e0, uint32_t rgba, int W, int H, std::vector<FEdge> &tmp, std::vector<uint32_t> &count)
{
    const size_t n = out.edges.size() - e0;
    if (n == 0) return;
    int ymin = 1 << 30, ymax = -(1 << 30), xmin = 1 << 30, xmax = -(1 << 30), maxspan = 0;
    for (size_t i = e0; i < out.edges.size(); i++) {
        const FEdge &e = out.edges[i];
        const int a = std::min(e.y1, e.y2) >> 8, b = std::max(e.y1, e.y2) >> 8;
        ymin = std::min(ymin, a); ymax = std::max(ymax, b);
        xmin = std::min(xmin, std::min(e.x1, e.x2) >> 8); xmax = std::max(xmax, std::max(e.x1, e.x2) >> 8);
        maxspan = std::max(maxspan, b - a);
    }
    FLayer L;
    memset(&L, 0, sizeof L);
    L.edge_begin = (uint32_t)e0;
    L.n_edges = (uint32_t)n;
    L.rgba = rgba;
    L.ymin = (int16_t)std::max(ymin, 0);      // (edges are clipped to the canvas: ymin >= 0 already)
    L.ymax = (int16_t)std::min(ymax, H - 1);
    L.xmin = (int16_t)std::max(xmin, 0);
    L.xmax = (int16_t)std::min(xmax, W - 1);
    L.maxspan = (uint16_t)std::min(maxspan, 65535);
    L.dir_begin = (uint32_t)out.dir.size();
    // Sort the layer's edges by their first row (the cell sums commute, so any order gives the same pixels; row order keeps the
    // lanes of a warp on neighbouring rows, which measures 4 % faster in k_figure than slab order): a stable counting sort --
    // edges of one row stay in the order the path emitted them, which is as deterministic as a total order and a quarter of
    // the host pipeline's time cheaper than std::sort.  Then the slab directory:
    // dir[s] = index (relative to edge_begin) of the first edge whose first row >= ymin + 16 s.
    const int rows = ymax - ymin + 1;                 // first rows lie in [ymin, ymax]
    count.assign((size_t)rows + 1, 0u);
    tmp.assign(out.edges.begin() + e0, out.edges.end());
    for (const FEdge &e : tmp) count[(size_t)((std::min(e.y1, e.y2) >> 8) - ymin) + 1]++;
    for (int r = 0; r < rows; r++) count[r + 1] += count[r];          // count[r] = edges whose first row < ymin + r
    const int slabs = (ymax - ymin) / FIG_SLAB_ROWS + 2;
    for (int s = 0; s < slabs; s++) out.dir.push_back(count[std::min(s * FIG_SLAB_ROWS, rows)]);
    FEdge *dst = out.edges.data() + e0;
    for (const FEdge &e : tmp) dst[count[(size_t)((std::min(e.y1, e.y2) >> 8) - ymin)]++] = e;
    L.ndir = (uint16_t)slabs;
    out.layers.push_back(L);
}

}  // namespace

int figure_prepare(const rig_figure_scene &s, int W, int H, bool bgr, int i0, int i1, FigurePrepared &out, char *err, int errlen)
{
    // scratch that keeps its capacity from call to call (a call covers a block of a few dozen images)
    struct Scratch {
        std::vector<SV> stroke_storage;
        std::vector<FEdge> sort_tmp;
        std::vector<uint32_t> sort_count;
    };
    thread_local Scratch scratch;
    VecSeq seq{scratch.stroke_storage};
    VecEdges edges{out.edges};
    geom::Certain certain;
    for (int img = i0; img < i1; img++) {
        out.image_layer_begin.push_back((uint32_t)out.layers.size());
        for (uint32_t a = s.image_artist_begin[img]; a < s.image_artist_begin[img + 1]; a++) {
            geom::ArtistSetup art;
            int what = 0;
            if (geom::artist_setup(s, a, W, H, art, what)) {
                if (what == 1) { snprintf(err, errlen, "artist %u: unknown path code", a); return RIG_E_INVALID; }
                if (what == 2) { snprintf(err, errlen, "artist %u: non-finite vertex (matplotlib's PathNanRemover is not reproduced)", a); return RIG_E_UNSUPPORTED; }
                snprintf(err, errlen, "artist %u: bad line width", a);
                return RIG_E_INVALID;
            }
            if (art.face) {
                const size_t e0 = out.edges.size();
                geom::face_pass(art, edges);
                finish_layer(out, e0, geom::pack_rgba(s.face_rgba + 4 * (size_t)a, bgr), W, H, scratch.sort_tmp, scratch.sort_count);
            }
            if (art.stroke) {
                const size_t e0 = out.edges.size();
                geom::stroke_pass(art, s.cap[a], s.join[a], seq, edges, certain);
                finish_layer(out, e0, geom::pack_rgba(s.edge_rgba + 4 * (size_t)a, bgr), W, H, scratch.sort_tmp, scratch.sort_count);
            }
        }
    }
    out.image_layer_begin.push_back((uint32_t)out.layers.size());
    return RIG_OK;
}

}  // namespace rig
