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
//   * one THREAD per artist (a stroker is a sequential machine: a join needs the de-duplicated neighbours, the clipper the
//     previous point).  Artists are handed out 32 at a time from a global counter in an order the host chose so that the lanes of
//     a warp hold paths of the same length (same patch type: same control flow, little divergence).
//   * a thread's stroke-vertex sequence and its unsorted edges live in thread-private global scratch (L1/L2 resident).
//   * after each pass (faces, then strokes) the WARP places the lanes' layers: one atomicAdd reserves the edges and directory
//     entries of all 32 layers, then per layer a stable counting sort by first row (shared-memory histogram, match_any ranks)
//     moves the edges from the owner's scratch to their final place with coalesced 16-byte accesses and writes the slab directory.
// Layer slots are fixed by the host (artist_layer[a] + 0 / 1), so image_layer_begin needs no device-wide scan; a pass that
// leaves no edge keeps its slot as an empty layer (ymin > ymax), which k_figure's band test skips.
#pragma once
#include <cuda_runtime.h>

#include "rig_figure_geom.h"

namespace rig {

constexpr int GEOM_THREADS = 128;           // per CTA
constexpr int GEOM_WARPS = GEOM_THREADS / 32;

// counters[] slots
constexpr int GC_WORK = 0, GC_EDGES = 1, GC_DIR = 2, GC_STATUS = 3, GC_BAD_ARTIST = 4, GC_BAD_WHAT = 5, GC_WORDS = 8;

struct GeomParams {
    rig_figure_scene s;              // device pointers
    const uint32_t *artist_order;    // [n_artists] artists in work order
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
    const uint32_t nd = n ? (uint32_t)((E.ymax - E.ymin) / FIG_SLAB_ROWS + 2) : 0u;
    uint32_t etot, dtot;
    const uint32_t eoff = warp_excl_scan(n, lane, etot), doff = warp_excl_scan(nd, lane, dtot);
    uint32_t ebase = 0, dbase = 0;
    if (lane == 0 && etot) {
        ebase = atomicAdd(P.counters + GC_EDGES, etot);
        dbase = atomicAdd(P.counters + GC_DIR, dtot);
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
            // one chunk: rank by comparing keys across the lanes (stable: ties in emission order)
            int4 v = make_int4(0, 0, 0, 0);
            int key = 0x7fffffff;
            if ((uint32_t)lane < sn) {
                v = *(const int4 *)(from + lane);
                key = (min(v.y, v.w) >> 8) - symin;
            }
            uint32_t rank = 0;
            for (uint32_t j = 0; j < sn; j++) {
                const int kj = __shfl_sync(0xffffffffu, key, (int)j);
                rank += (kj < key || (kj == key && (int)j < lane)) ? 1u : 0u;
            }
            if ((uint32_t)lane < sn) *(int4 *)(to + rank) = v;
            // slab directory: dir[s] = edges whose first row < 16 s
            const int rows = symax - symin + 1, slabs = (symax - symin) / FIG_SLAB_ROWS + 2;
            for (int s = 0; s < slabs; s++) {
                const int lim = min(s * FIG_SLAB_ROWS, rows);
                const uint32_t c = __popc(__ballot_sync(0xffffffffu, key < lim));
                if (lane == 0) P.dir[sd + s] = c;
            }
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
        const int slabs = (symax - symin) / FIG_SLAB_ROWS + 2;
        for (int s = lane; s < slabs; s += 32) P.dir[sd + s] = hist[min(s * FIG_SLAB_ROWS, rows)];
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

__global__ void __launch_bounds__(GEOM_THREADS) k_figure_path(const GeomParams P)
{
    extern __shared__ uint32_t geom_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *hist = geom_smem + (size_t)warp * (size_t)(P.H + 34);
    const size_t tid = (size_t)blockIdx.x * GEOM_THREADS + threadIdx.x;
    DevSeq seq;
    seq.p = P.seq + tid * P.seq_cap; seq.n = 0; seq.cap = (int)P.seq_cap; seq.overflow = false;
    DevEdges E;
    E.p = P.stage + tid * P.stage_cap; E.cap = P.stage_cap;
    E.reset();
    uint32_t status = 0;
    const uint32_t n_art = (uint32_t)P.s.n_artists;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(P.counters + GC_WORK, 32u);
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
                art.face = P.s.has_face[a] && P.s.face_rgba[4 * (size_t)a + 3] != 0;
                art.stroke = P.s.edge_rgba[4 * (size_t)a + 3] != 0 && P.s.linewidth[a] != 0.0;
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

// per-image scheduling weight for k_figure's costliest-first order: edges + 8 x the rows its layers resolve
__global__ void k_figure_cost(const uint32_t *image_layer_begin, const FLayer *layers, uint32_t n_images, float *cost)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_images) return;
    float w = 0.f;
    for (uint32_t l = image_layer_begin[i]; l < image_layer_begin[i + 1]; l++) {
        const FLayer L = layers[l];
        w += (float)L.n_edges + 8.f * (float)max(0, (int)L.ymax - (int)L.ymin + 1);
    }
    cost[i] = w;
}

}  // namespace rig
