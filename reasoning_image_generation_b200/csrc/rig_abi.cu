// rig_abi.cu -- the extern "C" boundary declared in include/rig.h: context, uploads, chunked launches.
//
// Host flow of one render: pack the borrowed SoA scene into 40-byte device elements (libm cos/sin of
// the integer angle, centre truncation / grid snap), copy them to HBM, then for every chunk of
// panels run k_expand (elements -> draw records in an L2-sized arena) and k_raster (records ->
// pixels, TMA bulk stores).  With a host destination the chunks are double-buffered through two
// device staging buffers so the D2H copy of chunk i overlaps the rendering of chunk i+1.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "rig_host.h"
#include "rig_kernels.cuh"
#include "rig_dedup.cuh"

using namespace rig;

namespace {

thread_local char g_create_err[512] = "";

struct Chunk {
    uint32_t p0, p1, e0, e1;
};

}  // namespace

struct rig_ctx {
    int device = 0, W = 0, H = 0, max_panels = 0, max_elements = 0;
    int sm_count = 0, raster_warps = 0;
    size_t raster_smem = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    ShapeTables *d_tables = nullptr;
    DElem *d_elems = nullptr, *h_elems = nullptr;
    uint32_t *d_begin = nullptr, *d_order = nullptr, *d_eorder = nullptr;
    std::vector<uint32_t> h_eorder;
    uint8_t *d_flags = nullptr;
    std::vector<uint32_t> h_order;
    std::vector<double> panel_cost;
    ElemHdr *d_hdrs = nullptr;
    Rec *d_recs = nullptr;
    Q16 *d_aux = nullptr;
    uint32_t rec_cap = 0, aux_cap = 0;
    uint32_t *d_counters = nullptr;   // CTR_COUNT words
    uint32_t *d_overflow = nullptr;   // sticky across the chunks of one render
    uint8_t *d_stage[2] = {nullptr, nullptr};
    size_t stage_panels = 0;
    uint32_t *d_sums = nullptr;       // per-image block sums of the hash stage
    int16_t *d_gridx = nullptr;       // vertical grid line positions for the fused hash
    size_t sums_cap = 0;
    // resident scene
    int n_panels = 0, n_elements = 0, grid_size = 3, layout = RIG_LAYOUT_RGB;
    uint32_t bg = 0x00ffffffu;
    std::vector<uint32_t> h_begin;
    std::vector<uint32_t> nominal_recs, nominal_aux;   // per element, one copy
    std::vector<Chunk> chunks;
    double budget_frac = 0.6;                 // sticky: halved by every arena overflow, never reset by an upload
    double nominal_scale = 1.0;               // RIG_NOMINAL_SCALE (tests: < 1 under-sizes the chunks to force the overflow path)
    uint32_t max_claim = 16, heavy_div = 4;   // RIG_MAX_CLAIM / RIG_HEAVY_DIV override (tuning)
    size_t chunk_panel_limit = 1u << 20;
    rig_stats stats = {};
    bool pending_overflow_check = false;
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool, ev_used;   // profiling events: triples (before expand, between, after raster)
    std::vector<uint32_t> ev_panels;
    char err[512] = "";
};

namespace {

int fail(rig_ctx *c, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c ? c->err : g_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(ctx, call)                                                                                     \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(ctx, RIG_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// records / aux units of one element assuming a single wrap copy (used only to size chunks)
void nominal_counts(int kind, int fill, int stroke, uint32_t &nrec, uint32_t &naux)
{
    const int t = stroke < 1 ? 1 : stroke;
    if (kind == K_CIRCLE) {
        nrec = (fill ? 1 : 0) + ring_recs(MAX_RING_VERTS - 1, t);
        naux = 0;
    } else if (kind == K_CRESCENT) {
        nrec = 1 + 2 * ring_recs(MAX_RING_VERTS - 1, t) + 4;
        naux = 0;
    } else {
        const int n = kind_nverts(kind), l = kind_loops(kind);
        nrec = l * ((fill ? n + 1 : 0) + n * stroke_recs(t));
        naux = l * (fill ? n : 0);
    }
}

// Upper bound of the wrap-around copies of an element (shapes.py:260-266 draws one copy per (+-W, +-H) offset whose shifted
// integer bounding box of the outline vertices meets the canvas).  Every outline vertex lies within size / sqrt(2) of the
// centre -- the corner of a square, the farthest any kind reaches: circles, crescents and the regular polygons stay within
// size / 2, the plus within 0.47 size, the heart within 0.43 size -- plus one pixel of rounding.  The device computes the exact
// mask; this only sizes the chunks, and a bound that is loose (the first version used `size`) doubles the planned records of
// a tenth of the elements and splits a step that fits the arena into two launches.
int wrap_copies_bound(int cx, int cy, int size, int stroke, int W, int H)
{
    (void)stroke;                                     // the copy decision looks at the vertices, not at the stroke's extent
    const int r = (int)(0.7072 * (double)size) + 2;
    int nx = 0, ny = 0;
    for (int k = -1; k <= 1; k++) {
        nx += (cx + k * W + r >= 0 && cx + k * W - r <= W);
        ny += (cy + k * H + r >= 0 && cy + k * H - r <= H);
    }
    return std::max(1, nx) * std::max(1, ny);
}

void plan_chunks(rig_ctx *c, size_t panel_limit)
{
    c->chunks.clear();
    const double rb = c->rec_cap * c->budget_frac, ab = c->aux_cap * c->budget_frac;
    uint32_t p0 = 0;
    double r = 0, a = 0;
    for (uint32_t p = 0; p < (uint32_t)c->n_panels; p++) {
        double pr = 0, pa = 0;
        for (uint32_t e = c->h_begin[p]; e < c->h_begin[p + 1]; e++) { pr += c->nominal_recs[e]; pa += c->nominal_aux[e]; }
        if (p > p0 && (r + pr > rb || a + pa > ab || p - p0 >= panel_limit)) {
            c->chunks.push_back({p0, p, c->h_begin[p0], c->h_begin[p]});
            p0 = p; r = a = 0;
        }
        r += pr; a += pa;
    }
    if ((uint32_t)c->n_panels > p0) c->chunks.push_back({p0, (uint32_t)c->n_panels, c->h_begin[p0], c->h_begin[c->n_panels]});
    // longest-processing-time-first inside each chunk: the persistent raster warps claim bands in this
    // order, so the heavy panels start first and the launch does not end on one slow band
    c->panel_cost.resize(c->n_panels);
    c->h_order.resize(c->n_panels);
    for (uint32_t p = 0; p < (uint32_t)c->n_panels; p++) {
        double pr = 0;
        for (uint32_t e = c->h_begin[p]; e < c->h_begin[p + 1]; e++) pr += c->nominal_recs[e];
        c->panel_cost[p] = pr;
        c->h_order[p] = p;
    }
    c->h_eorder.resize(c->n_elements);
    for (uint32_t e = 0; e < (uint32_t)c->n_elements; e++) c->h_eorder[e] = e;
    // Panel order inside a chunk.  RIG_ORDER: 0 natural, 1 (default) costliest first, 2 costliest 1/32 first then natural,
    // 3 costliest paired with cheapest.  Measured (profiles/r1_notes.md): 5.30 / 4.43 / 4.85 / 4.38 ms -- the launch tail
    // is what matters, interleaving store-bound and issue-bound bands does not.
    int order_mode = 1;
    if (const char *e = getenv("RIG_ORDER")) order_mode = atoi(e);
    for (const Chunk &ch : c->chunks) {
        if (order_mode == 1)
            std::stable_sort(c->h_order.begin() + ch.p0, c->h_order.begin() + ch.p1,
                             [&](uint32_t a, uint32_t b) { return c->panel_cost[a] > c->panel_cost[b]; });
        else if (order_mode == 3) {
            // costliest paired with cheapest, second costliest with second cheapest, ...: every stretch of the order
            // mixes issue-bound and store-bound bands, and the launch ends on median panels
            std::vector<uint32_t> byc(c->h_order.begin() + ch.p0, c->h_order.begin() + ch.p1);
            std::stable_sort(byc.begin(), byc.end(), [&](uint32_t a, uint32_t b) { return c->panel_cost[a] > c->panel_cost[b]; });
            size_t lo = 0, hi = byc.size(), o = ch.p0;
            while (lo < hi) {
                c->h_order[o++] = byc[lo++];
                if (lo < hi) c->h_order[o++] = byc[--hi];
            }
        }
        else if (order_mode == 2) {
            const size_t nlead = (ch.p1 - ch.p0) / 32;
            std::vector<uint32_t> idx(c->h_order.begin() + ch.p0, c->h_order.begin() + ch.p1);
            std::vector<uint32_t> byc(idx);
            std::stable_sort(byc.begin(), byc.end(), [&](uint32_t a, uint32_t b) { return c->panel_cost[a] > c->panel_cost[b]; });
            std::vector<uint8_t> lead(c->n_panels, 0);
            for (size_t i = 0; i < nlead; i++) lead[byc[i]] = 1;
            size_t o = ch.p0;
            for (size_t i = 0; i < nlead; i++) c->h_order[o++] = byc[i];
            for (uint32_t p : idx) if (!lead[p]) c->h_order[o++] = p;
        }
        std::stable_sort(c->h_eorder.begin() + ch.e0, c->h_eorder.begin() + ch.e1,
                         [&](uint32_t a, uint32_t b) { return c->nominal_recs[a] > c->nominal_recs[b]; });
    }
}

int upload_order(rig_ctx *c, cudaStream_t st)
{
    if (c->n_panels) {
        CK(c, cudaMemcpyAsync(c->d_order, c->h_order.data(), sizeof(uint32_t) * c->n_panels, cudaMemcpyHostToDevice, st));
        if (c->n_elements)
            CK(c, cudaMemcpyAsync(c->d_eorder, c->h_eorder.data(), sizeof(uint32_t) * c->n_elements, cudaMemcpyHostToDevice, st));
        CK(c, cudaStreamSynchronize(st));
    }
    return RIG_OK;
}

int validate_scene(rig_ctx *c, const rig_scene *s)
{
    if (!s) return fail(c, RIG_E_INVALID, "scene is NULL");
    if (s->n_panels < 0 || s->n_elements < 0) return fail(c, RIG_E_INVALID, "negative panel/element count");
    if (s->n_panels > c->max_panels) return fail(c, RIG_E_CAPACITY, "%d panels > max_panels %d", s->n_panels, c->max_panels);
    if (s->n_elements > c->max_elements) return fail(c, RIG_E_CAPACITY, "%d elements > max_elements %d", s->n_elements, c->max_elements);
    if (s->n_panels && (!s->panel_elem_begin)) return fail(c, RIG_E_INVALID, "panel_elem_begin is NULL");
    if (s->n_elements && (!s->kind || !s->fill || !s->stroke || !s->size || !s->angle || !s->cx || !s->cy || !s->rgb))
        return fail(c, RIG_E_INVALID, "an element array is NULL");
    if (s->grid_size < 0 || s->grid_size > 16) return fail(c, s->grid_size < 0 ? RIG_E_INVALID : RIG_E_UNSUPPORTED, "grid_size %d out of range (1..16)", s->grid_size);
    if (s->n_panels) {
        if (s->panel_elem_begin[0] != 0 || s->panel_elem_begin[s->n_panels] != (uint32_t)s->n_elements)
            return fail(c, RIG_E_INVALID, "panel_elem_begin must start at 0 and end at n_elements");
        for (int p = 0; p < s->n_panels; p++)
            if (s->panel_elem_begin[p] > s->panel_elem_begin[p + 1]) return fail(c, RIG_E_INVALID, "panel_elem_begin is not monotone");
    }
    return RIG_OK;
}

cudaEvent_t prof_event(rig_ctx *c, cudaStream_t st)
{
    cudaEvent_t e = nullptr;
    if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    cudaEventRecord(e, st);
    c->ev_used.push_back(e);
    return e;
}

void resolve_profiling(rig_ctx *c)
{
    for (size_t i = 0; i + 2 < c->ev_used.size() + 0 && i / 3 < c->ev_panels.size(); i += 3) {
        float a = 0, b = 0;
        if (cudaEventElapsedTime(&a, c->ev_used[i], c->ev_used[i + 1]) == cudaSuccess &&
            cudaEventElapsedTime(&b, c->ev_used[i + 1], c->ev_used[i + 2]) == cudaSuccess) {
            c->stats.expand_ms += a;
            c->stats.raster_ms += b;
            c->stats.timed_launches++;
            c->stats.timed_panels += c->ev_panels[i / 3];
        }
    }
    for (cudaEvent_t e : c->ev_used) c->ev_pool.push_back(e);
    c->ev_used.clear();
    c->ev_panels.clear();
}

int launch_chunk(rig_ctx *c, const Chunk &ch, uint8_t *out_dev, cudaStream_t st, bool hash = false)
{
    CK(c, cudaMemsetAsync(c->d_counters, 0, CTR_COUNT * sizeof(uint32_t), st));
    const uint32_t ne = ch.e1 - ch.e0;
    if (c->profiling) prof_event(c, st);
    if (ne) {
        k_expand<<<(ne + EXPAND_WARPS - 1) / EXPAND_WARPS, EXPAND_WARPS * 32, 0, st>>>(
            c->d_elems, c->d_eorder, ch.e0, ch.e1, c->d_tables, c->W, c->H, c->d_hdrs, c->d_recs, c->rec_cap, c->d_aux, c->aux_cap,
            c->d_counters);
        c->stats.kernel_launches++;
    }
    if (c->profiling) prof_event(c, st);
    RasterParams P;
    P.panel_begin = c->d_begin; P.panel_flags = c->d_flags; P.order = c->d_order; P.hdrs = c->d_hdrs; P.recs = c->d_recs; P.aux = c->d_aux;
    P.tables = c->d_tables; P.out = out_dev; P.counters = c->d_counters; P.p0 = ch.p0; P.n_panels = ch.p1 - ch.p0;
    P.W = c->W; P.H = c->H; P.grid_size = c->grid_size; P.bg = c->bg;
    P.max_claim = c->max_claim;
    P.skip_mask = 0;
    if (const char *e = getenv("RIG_SKIP_MASK")) P.skip_mask = (uint32_t)strtoul(e, nullptr, 0);
    P.n_grid = c->grid_size - 1;
    for (int g = 1; g < c->grid_size && g <= 16; g++) {
        // int(round(i * W / grid_size)): Python rounds the float quotient half-to-even (generator.py:117,121)
        volatile double qx = (double)(g * c->W) / (double)c->grid_size, qy = (double)(g * c->H) / (double)c->grid_size;
        P.grid_x[g - 1] = (int16_t)nearbyint(qx);
        P.grid_y[g - 1] = (int16_t)nearbyint(qy);
    }
    {
        // panels are ordered heaviest first: those costing >= 2x the chunk's mean are handed out band by band
        double sum = 0;
        for (uint32_t p = ch.p0; p < ch.p1; p++) sum += c->panel_cost[p];
        const double mean = sum / std::max(1u, ch.p1 - ch.p0);
        uint32_t heavy = 0;
        while (heavy < ch.p1 - ch.p0 && heavy < (ch.p1 - ch.p0) / std::max(1u, c->heavy_div) &&
               c->panel_cost[c->h_order[ch.p0 + heavy]] >= 2.0 * mean && c->panel_cost[c->h_order[ch.p0 + heavy]] > 0)
            heavy++;
        P.heavy_bands = c->heavy_div ? heavy * (uint32_t)((c->H + BAND_ROWS - 1) / BAND_ROWS) : 0u;
    }
    const uint32_t bands = P.n_panels * (uint32_t)((c->H + BAND_ROWS - 1) / BAND_ROWS);
    const uint32_t want = (bands + c->raster_warps - 1) / c->raster_warps;
    const uint32_t grid = std::max(1u, std::min((uint32_t)c->sm_count, want));
    P.sums = c->d_sums;
    P.coef = c->layout == RIG_LAYOUT_RGB ? (77u | (150u << 8) | (29u << 16)) : (29u | (150u << 8) | (77u << 16));
    if (hash) k_raster<true><<<grid, c->raster_warps * 32, c->raster_smem, st>>>(P);
    else k_raster<false><<<grid, c->raster_warps * 32, c->raster_smem, st>>>(P);
    c->stats.kernel_launches++;
    if (c->profiling) { prof_event(c, st); c->ev_panels.push_back(ch.p1 - ch.p0); }
    k_collect_overflow<<<1, 32, 0, st>>>(c->d_counters, c->d_overflow);
    c->stats.kernel_launches++;
    CK(c, cudaGetLastError());
    return RIG_OK;
}

int check_overflow(rig_ctx *c)
{
    uint32_t h[4] = {0, 0, 0, 0};
    CK(c, cudaMemcpy(h, c->d_overflow, sizeof h, cudaMemcpyDeviceToHost));
    c->pending_overflow_check = false;
    c->stats.overflow = h[0];
    c->stats.records = h[1];
    c->stats.aux_units = h[2];
    return h[0] ? 1 : 0;
}

int upload_scene(rig_ctx *c, const rig_scene *s, int layout, cudaStream_t st)
{
    int rc = validate_scene(c, s);
    if (rc) return rc;
    if (layout != RIG_LAYOUT_RGB && layout != RIG_LAYOUT_BGR) return fail(c, RIG_E_INVALID, "bad out_layout %d", layout);
    char msg[256] = "";
    rc = pack_elements(*s, c->W, c->H, layout == RIG_LAYOUT_BGR, c->h_elems, msg, sizeof msg);
    if (rc) return fail(c, rc, "%s", msg);
    c->n_panels = s->n_panels; c->n_elements = s->n_elements;
    c->grid_size = s->grid_size > 0 ? s->grid_size : 3;
    c->layout = layout;
    c->bg = layout == RIG_LAYOUT_BGR ? ((uint32_t)s->bg_rgb[2] | (s->bg_rgb[1] << 8) | (s->bg_rgb[0] << 16))
                                     : ((uint32_t)s->bg_rgb[0] | (s->bg_rgb[1] << 8) | (s->bg_rgb[2] << 16));
    c->h_begin.assign(s->panel_elem_begin, s->panel_elem_begin + s->n_panels + 1);
    if (s->n_panels == 0) c->h_begin.assign(1, 0);
    c->nominal_recs.resize(s->n_elements);
    c->nominal_aux.resize(s->n_elements);
    for (int i = 0; i < s->n_elements; i++) {
        nominal_counts(s->kind[i], s->fill[i], s->stroke[i], c->nominal_recs[i], c->nominal_aux[i]);
        // x the wrap copies the element can have, so that a planned chunk cannot overflow the arena (the caller of the
        // asynchronous device path may never call rig_sync and would otherwise not learn of missing elements)
        const DElem &e = c->h_elems[i];
        const double k = wrap_copies_bound(e.cx, e.cy, e.size, e.stroke, c->W, c->H) * c->nominal_scale;
        c->nominal_recs[i] = (uint32_t)std::max(1.0, c->nominal_recs[i] * k);
        c->nominal_aux[i] = (uint32_t)(c->nominal_aux[i] * k);
    }
    if (s->n_elements) CK(c, cudaMemcpyAsync(c->d_elems, c->h_elems, sizeof(DElem) * s->n_elements, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(c->d_begin, c->h_begin.data(), sizeof(uint32_t) * (s->n_panels + 1), cudaMemcpyHostToDevice, st));
    if (s->n_panels) {
        if (s->panel_flags) CK(c, cudaMemcpyAsync(c->d_flags, s->panel_flags, s->n_panels, cudaMemcpyHostToDevice, st));
        else CK(c, cudaMemsetAsync(c->d_flags, 0, s->n_panels, st));
    }
    // h_elems / h_begin are reused by the next upload: the copies above must have left the host
    CK(c, cudaStreamSynchronize(st));
    plan_chunks(c, c->chunk_panel_limit);      // budget_frac is sticky: a scene that overflowed once is re-planned smaller
    return upload_order(c, st);
}

int ensure_sums(rig_ctx *c, size_t n, cudaStream_t st)
{
    if (c->sums_cap < n) {
        CK(c, cudaStreamSynchronize(st));
        if (c->d_sums) CK(c, cudaFree(c->d_sums));
        c->d_sums = nullptr;
        CK(c, cudaMalloc(&c->d_sums, sizeof(uint32_t) * 256 * n));
        c->sums_cap = n;
    }
    return RIG_OK;
}

int render_device(rig_ctx *c, uint8_t *out_dev, cudaStream_t st, uint64_t *hashes_dev = nullptr)
{
    CK(c, cudaMemsetAsync(c->d_overflow, 0, 4 * sizeof(uint32_t), st));
    const size_t img = (size_t)c->W * c->H * 3;
    if (hashes_dev && c->n_panels) {
        int rc = ensure_sums(c, (size_t)c->n_panels, st);
        if (rc) return rc;
        CK(c, cudaMemsetAsync(c->d_sums, 0, sizeof(uint32_t) * 256 * (size_t)c->n_panels, st));
    }
    for (const Chunk &ch : c->chunks) {
        int rc = launch_chunk(c, ch, out_dev + (size_t)ch.p0 * img, st, hashes_dev != nullptr);
        if (rc) return rc;
    }
    if (hashes_dev && c->n_panels) {
        int16_t gx[16];
        for (int g = 1; g < c->grid_size && g <= 16; g++) {
            volatile double qx = (double)(g * c->W) / (double)c->grid_size;
            gx[g - 1] = (int16_t)nearbyint(qx);
        }
        CK(c, cudaMemcpyAsync(c->d_gridx, gx, sizeof gx, cudaMemcpyHostToDevice, st));
        const uint32_t r = c->layout == RIG_LAYOUT_RGB ? (c->bg & 0xff) : ((c->bg >> 16) & 0xff), g = (c->bg >> 8) & 0xff,
                       b = c->layout == RIG_LAYOUT_RGB ? ((c->bg >> 16) & 0xff) : (c->bg & 0xff);
        const uint32_t luma_bg = (77u * r + 150u * g + 29u * b) >> 8;
        k_hash_bits_fused<<<((unsigned)c->n_panels * 32 + 255) / 256, 256, 0, st>>>(c->d_sums, c->d_flags, c->n_panels, c->W, c->H, luma_bg,
                                                                               c->grid_size - 1, c->d_gridx, hashes_dev);
        c->stats.kernel_launches++;
        CK(c, cudaGetLastError());
    }
    c->pending_overflow_check = true;
    return RIG_OK;
}

}  // namespace

extern "C" {

const char *rig_last_error(const rig_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

int rig_create(int device, int W, int H, int max_panels, int max_elements, rig_ctx **out)
{
    if (!out) return fail(nullptr, RIG_E_INVALID, "out is NULL");
    *out = nullptr;
    if (W <= 0 || H <= 0 || W % 16 || W > 4096 || H > 4096)
        return fail(nullptr, RIG_E_INVALID, "unsupported canvas %dx%d (W %% 16 == 0 so that rows are 16-byte aligned; <= 4096)", W, H);
    if (max_panels <= 0 || max_elements <= 0) return fail(nullptr, RIG_E_INVALID, "max_panels/max_elements must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RIG_E_NODEVICE, "no usable CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, RIG_E_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    rig_ctx *c = new rig_ctx();
    c->device = device; c->W = W; c->H = H; c->max_panels = max_panels; c->max_elements = max_elements;
    auto bail = [&](int code) {
        snprintf(g_create_err, sizeof g_create_err, "%s", c->err);
        rig_destroy(c);
        return code;
    };
#define CKC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            fail(c, RIG_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                   \
            return bail(RIG_E_CUDA);                                                               \
        }                                                                                          \
    } while (0)
    CKC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 9) {
        fail(c, RIG_E_UNSUPPORTED, "device %d is sm_%d%d; the kernels need bulk async copies (sm_90+; built for sm_100a)", device, prop.major, prop.minor);
        return bail(RIG_E_UNSUPPORTED);
    }
    c->sm_count = prop.multiProcessorCount;
    const size_t smem_max = prop.sharedMemPerBlockOptin;
    int warps = RASTER_MAX_WARPS;
    while (warps > 1 && raster_smem_bytes(W, warps) > smem_max) warps--;
    if (raster_smem_bytes(W, warps) > smem_max) {
        fail(c, RIG_E_UNSUPPORTED, "canvas width %d needs %zu B of shared memory per band set; the device offers %zu", W, raster_smem_bytes(W, warps), smem_max);
        return bail(RIG_E_UNSUPPORTED);
    }
    c->raster_warps = warps;
    if (const char *e = getenv("RIG_MAX_CLAIM")) c->max_claim = (uint32_t)std::max(1, atoi(e));
    if (const char *e = getenv("RIG_NOMINAL_SCALE")) { double v = atof(e); if (v > 0) c->nominal_scale = v; }
    if (const char *e = getenv("RIG_HEAVY_DIV")) c->heavy_div = (uint32_t)std::max(0, atoi(e));
    if (const char *e = getenv("RIG_WARPS")) { int w = atoi(e); if (w >= 1 && w <= warps) c->raster_warps = w; }
    c->raster_smem = raster_smem_bytes(W, c->raster_warps);
    CKC(cudaFuncSetAttribute(k_raster<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->raster_smem));
    CKC(cudaFuncSetAttribute(k_raster<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->raster_smem));
    CKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CKC(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
    }
    ShapeTables T;
    fill_shape_tables(T);
    CKC(cudaMalloc(&c->d_tables, sizeof T));
    CKC(cudaMemcpy(c->d_tables, &T, sizeof T, cudaMemcpyHostToDevice));
    CKC(cudaMalloc(&c->d_elems, sizeof(DElem) * (size_t)max_elements));
    CKC(cudaMallocHost(&c->h_elems, sizeof(DElem) * (size_t)max_elements));
    CKC(cudaMalloc(&c->d_begin, sizeof(uint32_t) * ((size_t)max_panels + 1)));
    CKC(cudaMalloc(&c->d_flags, (size_t)max_panels));
    CKC(cudaMalloc(&c->d_order, sizeof(uint32_t) * (size_t)max_panels));
    CKC(cudaMalloc(&c->d_eorder, sizeof(uint32_t) * (size_t)max_elements));
    CKC(cudaMalloc(&c->d_hdrs, sizeof(ElemHdr) * (size_t)max_elements));
    // record arena: sized so a whole batch is one chunk (one launch tail per batch), capped at 1 GB
    const uint64_t want_rec = std::min<uint64_t>(std::max<uint64_t>((uint64_t)max_elements * 1100, 1u << 16), 32u << 20);
    c->rec_cap = (uint32_t)want_rec;
    c->aux_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>((uint64_t)max_elements * 128, 1u << 16), 16u << 20);
    CKC(cudaMalloc(&c->d_recs, sizeof(Rec) * (size_t)c->rec_cap));
    CKC(cudaMalloc(&c->d_aux, sizeof(Q16) * (size_t)c->aux_cap));
    CKC(cudaMalloc(&c->d_counters, CTR_COUNT * sizeof(uint32_t)));
    CKC(cudaMalloc(&c->d_gridx, 16 * sizeof(int16_t)));
    CKC(cudaMalloc(&c->d_overflow, 4 * sizeof(uint32_t)));
    CKC(cudaMemset(c->d_overflow, 0, 4 * sizeof(uint32_t)));
#undef CKC
    *out = c;
    return RIG_OK;
}

void rig_destroy(rig_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    cudaFree(c->d_tables); cudaFree(c->d_elems); cudaFreeHost(c->h_elems); cudaFree(c->d_begin); cudaFree(c->d_order); cudaFree(c->d_eorder); cudaFree(c->d_flags);
    cudaFree(c->d_hdrs); cudaFree(c->d_recs); cudaFree(c->d_aux); cudaFree(c->d_counters); cudaFree(c->d_overflow);
    cudaFree(c->d_stage[0]); cudaFree(c->d_stage[1]); cudaFree(c->d_sums); cudaFree(c->d_gridx);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_used) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

int rig_scene_upload(rig_ctx *c, const rig_scene *scene, int out_layout, void *stream)
{
    if (!c) return RIG_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    return upload_scene(c, scene, out_layout, stream ? (cudaStream_t)stream : c->stream);
}

int rig_render_resident(rig_ctx *c, uint8_t *out_device, void *stream)
{
    if (!c) return RIG_E_INVALID;
    if (!out_device && c->n_panels) return fail(c, RIG_E_INVALID, "out_device is NULL");
    if (((uintptr_t)out_device) % 16) return fail(c, RIG_E_INVALID, "out_device must be 16-byte aligned");
    CK(c, cudaSetDevice(c->device));
    return render_device(c, out_device, stream ? (cudaStream_t)stream : c->stream);
}

int rig_render_resident_hashed(rig_ctx *c, uint8_t *out_device, uint64_t *hashes_device, void *stream)
{
    if (!c) return RIG_E_INVALID;
    if ((!out_device || !hashes_device) && c->n_panels) return fail(c, RIG_E_INVALID, "out_device / hashes_device is NULL");
    if (((uintptr_t)out_device) % 16) return fail(c, RIG_E_INVALID, "out_device must be 16-byte aligned");
    {
        const int lpb = c->W / 256;
        if (c->W % 256 || c->H % 16 || (lpb & (lpb - 1)) || c->W / 16 > 64)
            return fail(c, RIG_E_UNSUPPORTED, "fused hashing needs W %% 256 == 0 (W / 256 a power of two, W <= 1024) and H %% 16 == 0");
    }
    CK(c, cudaSetDevice(c->device));
    return render_device(c, out_device, stream ? (cudaStream_t)stream : c->stream, hashes_device);
}

int rig_sync(rig_ctx *c)
{
    if (!c) return RIG_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaStreamSynchronize(c->copy_stream));
    if (c->profiling) {
        CK(c, cudaDeviceSynchronize());
        resolve_profiling(c);
    }
    if (c->pending_overflow_check) {
        CK(c, cudaDeviceSynchronize());
        int o = check_overflow(c);
        if (o < 0) return o;
        if (o) {
            c->budget_frac *= 0.5;   // the next rig_render_resident uses smaller chunks
            plan_chunks(c, c->chunk_panel_limit);
            int rc2 = upload_order(c, c->stream);
            if (rc2) return rc2;
            return fail(c, RIG_E_CAPACITY, "draw-record arena overflowed (%u records in the render); chunks were halved, render again", c->stats.records);
        }
    }
    return RIG_OK;
}

int rig_render_batch(rig_ctx *c, const rig_scene *scene, uint8_t *out, int out_is_device, int out_layout, void *stream)
{
    if (!c) return RIG_E_INVALID;
    if (!out && scene && scene->n_panels) return fail(c, RIG_E_INVALID, "out is NULL");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc = upload_scene(c, scene, out_layout, st);
    if (rc) return rc;
    if (out_is_device) {
        if (((uintptr_t)out) % 16) return fail(c, RIG_E_INVALID, "out must be 16-byte aligned");
        return render_device(c, out, st);
    }
    // host destination: staged, double-buffered D2H
    const size_t img = (size_t)c->W * c->H * 3;
    for (int attempt = 0; attempt < 6; attempt++) {
        const size_t stage_limit = std::max<size_t>(1, std::min<size_t>(512, (size_t)c->n_panels));
        if (c->stage_panels < stage_limit) {
            CK(c, cudaStreamSynchronize(c->copy_stream));
            for (int i = 0; i < 2; i++) {
                if (c->d_stage[i]) CK(c, cudaFree(c->d_stage[i]));
                c->d_stage[i] = nullptr;
                CK(c, cudaMalloc(&c->d_stage[i], stage_limit * img));
            }
            c->stage_panels = stage_limit;
        }
        plan_chunks(c, std::min(c->stage_panels, c->chunk_panel_limit));
        rc = upload_order(c, st);
        if (rc) return rc;
        CK(c, cudaMemsetAsync(c->d_overflow, 0, 4 * sizeof(uint32_t), st));
        int k = 0;
        for (const Chunk &ch : c->chunks) {
            const int b = k & 1;
            if (k >= 2) CK(c, cudaStreamWaitEvent(st, c->ev_copied[b], 0));
            rc = launch_chunk(c, ch, c->d_stage[b], st);
            if (rc) return rc;
            CK(c, cudaEventRecord(c->ev_done[b], st));
            CK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_done[b], 0));
            CK(c, cudaMemcpyAsync(out + (size_t)ch.p0 * img, c->d_stage[b], (size_t)(ch.p1 - ch.p0) * img, cudaMemcpyDeviceToHost, c->copy_stream));
            CK(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
            k++;
        }
        CK(c, cudaStreamSynchronize(st));
        CK(c, cudaStreamSynchronize(c->copy_stream));
        if (c->profiling) resolve_profiling(c);
        int o = check_overflow(c);
        if (o < 0) return o;
        if (!o) return RIG_OK;
        c->budget_frac *= 0.5;   // wrap copies made some chunk larger than planned: retry with smaller chunks
    }
    return fail(c, RIG_E_CAPACITY, "draw-record arena overflowed even with single-panel chunks");
}

int rig_set_profiling(rig_ctx *c, int on)
{
    if (!c) return RIG_E_INVALID;
    c->profiling = on != 0;
    if (on) {
        // events for the launches to come are created now, not between the timed launches
        if (cudaSetDevice(c->device) == cudaSuccess)
            while (c->ev_pool.size() < 512) {
                cudaEvent_t e = nullptr;
                if (cudaEventCreate(&e) != cudaSuccess) break;
                c->ev_pool.push_back(e);
            }
        c->stats.expand_ms = c->stats.raster_ms = 0;
        c->stats.timed_launches = 0;
        c->stats.timed_panels = 0;
    }
    return RIG_OK;
}

int rig_get_stats(rig_ctx *c, rig_stats *out)
{
    if (!c || !out) return RIG_E_INVALID;
    *out = c->stats;
    return RIG_OK;
}

int rig_hash_words(void) { return HASH_WORDS; }

int rig_hash_batch(rig_ctx *c, const uint8_t *imgs_dev, int n, int layout, uint64_t *hashes_dev, void *stream)
{
    if (!c) return RIG_E_INVALID;
    if (n < 0 || (n && (!imgs_dev || !hashes_dev))) return fail(c, RIG_E_INVALID, "bad arguments to rig_hash_batch");
    if (layout != RIG_LAYOUT_RGB && layout != RIG_LAYOUT_BGR) return fail(c, RIG_E_INVALID, "bad layout %d", layout);
    {
        const int lpb = c->W / HASH_GRID / 16;    // lanes per hash block in k_block_sums
        if (c->W % (16 * HASH_GRID) || c->H % HASH_GRID || (lpb & (lpb - 1)))
            return fail(c, RIG_E_UNSUPPORTED, "hashing needs W %% %d == 0, W / 256 a power of two and H %% %d == 0", 16 * HASH_GRID, HASH_GRID);
    }
    if (((uintptr_t)imgs_dev) % 16) return fail(c, RIG_E_INVALID, "imgs_dev must be 16-byte aligned");
    if (n == 0) return RIG_OK;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    if (c->sums_cap < (size_t)n) {
        CK(c, cudaStreamSynchronize(st));
        if (c->d_sums) CK(c, cudaFree(c->d_sums));
        c->d_sums = nullptr;
        CK(c, cudaMalloc(&c->d_sums, sizeof(uint32_t) * 256 * (size_t)n));
        c->sums_cap = (size_t)n;
    }
    const uint32_t coef = layout == RIG_LAYOUT_RGB ? (77u | (150u << 8) | (29u << 16)) : (29u | (150u << 8) | (77u << 16));
    k_block_sums<<<(unsigned)n * HASH_GRID, 256, 0, st>>>(imgs_dev, c->W, c->H, coef, c->d_sums);
    k_hash_bits<<<((unsigned)n * 32 + 255) / 256, 256, 0, st>>>(c->d_sums, n, hashes_dev);
    c->stats.kernel_launches += 2;
    CK(c, cudaGetLastError());
    return RIG_OK;
}

int rig_dedup(rig_ctx *c, const uint64_t *all_hashes_dev, uint64_t n_total, uint64_t row_begin, uint64_t row_end,
              int threshold, uint8_t *drop_flags_dev, uint64_t *hits_dev, void *stream)
{
    if (!c) return RIG_E_INVALID;
    if (row_begin > row_end || row_end > n_total) return fail(c, RIG_E_INVALID, "bad row range");
    if (n_total && (!all_hashes_dev || !drop_flags_dev)) return fail(c, RIG_E_INVALID, "NULL hash / flag array");
    if (threshold < 0 || threshold > 64 * HASH_WORDS) return fail(c, RIG_E_INVALID, "threshold out of range");
    if (row_begin == row_end) return RIG_OK;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    // macro tiles (row block I, column block J >= I + c0) that reach above the diagonal, enumerated row-major (rig_dedup.cuh)
    const uint64_t rows = row_end - row_begin;
    const uint64_t n_row_blocks = (rows + DD_B - 1) / DD_B, gx = (n_total + DD_B - 1) / DD_B;
    const uint64_t c0 = (row_begin + 1) / DD_B;
    if (c0 >= gx) return RIG_OK;                        // the only row left is the last one: no column above it
    const uint64_t per_row0 = gx - c0;
    const uint64_t nI = std::min<uint64_t>(n_row_blocks, per_row0);
    const uint64_t tiles = nI * per_row0 - nI * (nI - 1) / 2;
    const uint64_t max_grid = 1u << 30;
    for (uint64_t t0 = 0; t0 < tiles; t0 += max_grid) {
        const unsigned g = (unsigned)std::min<uint64_t>(max_grid, tiles - t0);
        k_hamming<<<g, DD_THREADS, 0, st>>>(all_hashes_dev, n_total, row_begin, row_end, threshold, (long long)per_row0, (unsigned)c0, t0,
                                            drop_flags_dev, (unsigned long long *)hits_dev);
        c->stats.kernel_launches++;
    }
    CK(c, cudaGetLastError());
    return RIG_OK;
}

int rig_int_pipe_peak(rig_ctx *c, int which, double *ops_per_second)
{
    if (!c || !ops_per_second) return RIG_E_INVALID;
    if (which != 0 && which != 1) return fail(c, RIG_E_INVALID, "which must be 0 (POPC) or 1 (LOP3)");
    CK(c, cudaSetDevice(c->device));
    uint32_t *d_out = nullptr;
    const int ctas = c->sm_count * 8, iters = 4096;
    CK(c, cudaMalloc(&d_out, sizeof(uint32_t) * 256 * (size_t)ctas));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t err = cudaEventCreate(&e0);
    if (err == cudaSuccess) err = cudaEventCreate(&e1);
    float best = 0.f;
    for (int rep = 0; rep < 4 && err == cudaSuccess; rep++) {          // rep 0 warms up
        cudaEventRecord(e0, c->stream);
        if (which == 0) k_int_pipe_peak<0><<<ctas, 256, 0, c->stream>>>(d_out, 0x2545f491u + rep, iters);
        else k_int_pipe_peak<1><<<ctas, 256, 0, c->stream>>>(d_out, 0x2545f491u + rep, iters);
        cudaEventRecord(e1, c->stream);
        err = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
        if (rep && (best == 0.f || ms < best)) best = ms;
    }
    c->stats.kernel_launches += 4;
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d_out);
    CK(c, err);
    // per thread and iteration: 64 chain steps, each 1 POPC (which = 0) or 2 LOP3 (which = 1)
    const double ops = (double)ctas * 256.0 * iters * 64.0 * (which == 0 ? 1.0 : 2.0);
    *ops_per_second = best > 0.f ? ops / (best * 1e-3) : 0.0;
    return RIG_OK;
}

int rig_pinned_alloc(size_t bytes, void **out)
{
    if (!out) return RIG_E_INVALID;
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail(nullptr, RIG_E_CUDA, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return RIG_OK;
}

void rig_pinned_free(void *p)
{
    if (p) cudaFreeHost(p);
}

size_t rig_debug_shape_tables(void *out, size_t size)
{
    if (out && size == sizeof(ShapeTables)) {
        ShapeTables T;
        fill_shape_tables(T);
        memcpy(out, &T, sizeof T);
    }
    return sizeof(ShapeTables);
}

}  // extern "C"
