// rig_figure_abi.cu -- the extern "C" boundary declared in include/rig_figure.h.
//
// Host flow of one render (device geometry, the default): pack the borrowed artist arrays into one page-locked blob, one H2D
// copy, k_figure_path (transform, snap, curve flattening, stroking, clipping -> layers / edges / slab directories in HBM),
// read back its status word and the per-image weights, launch k_figure (persistent, one CTA per SM).  If the status says the
// device could not vouch for bit-exactness (see geom::Uncertain) or ran out of a fixed capacity, or with
// rig_figure_set_geometry(ctx, RIG_GEOM_HOST), the host path pipeline (rig_figure_path.cpp) runs on a few host threads instead
// and its layers / edges are copied to HBM.  With a host destination the images go through two device staging buffers so the
// D2H copy of one chunk overlaps the next chunk.
#include <cuda_runtime.h>
#include <emmintrin.h>
#include <sched.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "rig_figure.cuh"
#include "rig_figure_host.h"
#include "rig_figure_path.cuh"
#include "rig_timeline.h"

using namespace rig;

// ---- optional timeline (rig_timeline.h) ---------------------------------------------------------------------------------------
namespace rig {
namespace {
struct TimelineRec { const char *name; cudaStream_t st; cudaEvent_t e0, e1; };
std::mutex g_tl_mutex;
std::vector<TimelineRec> g_tl;
std::vector<std::pair<cudaStream_t, size_t>> g_tl_open;
}
bool timeline_on()
{
    static const bool on = getenv("RIG_TIMELINE") != nullptr;
    return on;
}
void timeline_mark(const char *name, cudaStream_t st, bool begin)
{
    std::lock_guard<std::mutex> g(g_tl_mutex);
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    if (begin) {
        g_tl.push_back(TimelineRec{name, st, e, nullptr});
        g_tl_open.push_back({st, g_tl.size() - 1});
    } else {
        for (size_t k = g_tl_open.size(); k-- > 0;)
            if (g_tl_open[k].first == st && g_tl[g_tl_open[k].second].name == name) {
                g_tl[g_tl_open[k].second].e1 = e;
                g_tl_open.erase(g_tl_open.begin() + (long)k);
                return;
            }
        cudaEventDestroy(e);
    }
}
}  // namespace rig

extern "C" int rig_timeline_dump(const char *path)
{
    std::lock_guard<std::mutex> g(rig::g_tl_mutex);
    if (cudaDeviceSynchronize() != cudaSuccess) return RIG_E_CUDA;
    FILE *f = fopen(path, "w");
    if (!f) return RIG_E_INVALID;
    fprintf(f, "name,stream,start_ms,end_ms\n");
    for (const auto &r : rig::g_tl) {
        if (!r.e1) continue;
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, rig::g_tl[0].e0, r.e0);
        cudaEventElapsedTime(&b, rig::g_tl[0].e0, r.e1);
        fprintf(f, "%s,%p,%.4f,%.4f\n", r.name, (void *)r.st, a, b);
    }
    fclose(f);
    return RIG_OK;
}

namespace {
thread_local char g_fig_create_err[512] = "";

// Growable host array that is page-locked when a CUDA context owns it (so the H2D copy is one asynchronous DMA) and plain
// malloc memory for the GPU-less debug entry point.  Contents are NOT kept across a grow.
template <class T>
struct HostBuf {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    bool pinned = false;
    ~HostBuf() { release(); }
    void release()
    {
        if (p) { if (pinned) cudaFreeHost(p); else free(p); }
        p = nullptr; cap = 0; n = 0;
    }
    bool resize_discard(size_t want)
    {
        if (want > cap) {
            const bool pin = pinned;
            release();
            pinned = pin;
            const size_t c2 = std::max<size_t>(want + want / 4, 1024);
            if (pinned) { if (cudaMallocHost((void **)&p, c2 * sizeof(T)) != cudaSuccess) { p = nullptr; return false; } }
            else if (!(p = (T *)malloc(c2 * sizeof(T)))) return false;
            cap = c2;
        }
        n = want;
        return true;
    }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T *data() { return p; }
    const T *data() const { return p; }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
};

// what k_figure reads, all images of a batch back to back
struct FigureMerged {
    HostBuf<FLayer> layers;
    HostBuf<uint32_t> image_layer_begin;      // [n_images + 1]
    HostBuf<FEdge> edges;
    HostBuf<uint32_t> dir;
    void set_pinned(bool on) { layers.pinned = image_layer_begin.pinned = edges.pinned = dir.pinned = on; }
};

// A few host threads that stay alive between calls (spawning 16 threads per 4096-image chunk costs a visible part of the chunk's
// host time).  run(n, f) calls f(0..n-1) on the workers -- items are claimed one at a time from a shared counter, so a slow core
// or a run of expensive items delays nobody else -- and returns when all are done; one run at a time.
class HostPool {
public:
    explicit HostPool(int nthreads)
    {
        for (int t = 0; t < nthreads; t++) th_.emplace_back([this, t] { loop(t); });
    }
    ~HostPool()
    {
        {
            std::lock_guard<std::mutex> g(m_);
            quit_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto &x : th_) x.join();
    }
    int size() const { return (int)th_.size(); }
    void run(int n, const std::function<void(int)> &f)
    {
        std::unique_lock<std::mutex> g(m_);
        f_ = &f; n_ = n; left_ = (int)th_.size(); next_.store(0, std::memory_order_relaxed); gen_++;
        cv_.notify_all();
        done_.wait(g, [this] { return left_ == 0; });
        f_ = nullptr;
    }
private:
    void loop(int t)
    {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int)> *f;
            int n;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (quit_) return;
                f = f_; n = n_;
            }
            (void)t;
            for (int i; (i = next_.fetch_add(1, std::memory_order_relaxed)) < n;) (*f)(i);
            {
                std::lock_guard<std::mutex> g(m_);
                if (--left_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *f_ = nullptr;
    int n_ = 0, left_ = 0;
    std::atomic<int> next_{0};
    unsigned long long gen_ = 0;
    bool quit_ = false;
};

// cores this process may run on (its affinity mask: what a container's cpuset or taskset leaves), else the machine's
static unsigned usable_cores()
{
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) return (unsigned)CPU_COUNT(&set);
    return std::max(1u, std::thread::hardware_concurrency());
}

// Workers of one context: the box's cores, shared among the ranks torchrun started on it (LOCAL_WORLD_SIZE), at most 16.
// (Measured on a 16-core box with one GPU: 4 / 8 / 16 / 32 / 64 threads -> 0.81 / 0.90 / 1.03 / 1.00 / 0.98 M figures/s end to end.)
int host_thread_count()
{
    unsigned hw = usable_cores();
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) {
        const int ranks = atoi(e);
        if (ranks > 1) hw = std::max(2u, hw / (unsigned)ranks);
    }
    int nt = (int)std::min(hw, 16u);
    if (const char *e = getenv("RIG_HOST_THREADS")) nt = std::max(1, atoi(e));
    return nt;
}

// per-context scratch of the host path pipeline: the workers and their per-thread outputs, kept so their memory stays mapped
struct HostPrep {
    HostPool pool;
    std::vector<FigurePrepared> parts;
    std::vector<int> rcs;
    std::vector<std::vector<char>> errs;
    HostPrep() : pool(host_thread_count()) {}
};
}

struct rig_figure_ctx {
    int device = 0, W = 0, H = 0, max_images = 0, max_artists = 0;
    int sm_count = 0, warps = 0;
    size_t smem = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    cudaEvent_t ev_fence[8] = {};
    unsigned fence_next = 0;
    // resident scene
    // two merged host copies: the H2D copy of one batch is still running while the next batch is prepared into the other
    FigureMerged prep[2];
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr};
    unsigned prep_next = 0;
    HostPrep *host = nullptr;
    std::vector<uint32_t> order;
    std::vector<double> cost;
    int n_images = 0, layout = RIG_LAYOUT_RGB;
    FLayer *d_layers = nullptr;
    FEdge *d_edges = nullptr;
    uint32_t *d_dir = nullptr, *d_begin = nullptr, *d_order = nullptr, *d_work = nullptr;
    size_t cap_layers = 0, cap_edges = 0, cap_dir = 0, cap_images = 0;
    uint8_t *d_stage[2] = {nullptr, nullptr};
    size_t stage_images = 0;
    // blank-band flags of the last four device-destination renders (a consumer reads one while later renders write the others)
    uint8_t *d_blank[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned blank_next = 0;
    bool want_blank = false;
    const uint8_t *last_blank = nullptr;
    uint32_t max_claim = 0;      // bands a warp claims at most per visit to the queue; 0: by launch size (RIG_FIG_MAX_CLAIM overrides)
    rig_figure_stats stats = {};
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool, ev_used;
    std::vector<uint32_t> ev_images;      // per event pair: images covered; bit 31 set: the pair brackets k_figure_path
    // ---- device geometry (k_figure_path)
    int geom_mode = RIG_GEOM_DEVICE;
    bool resident_device_geom = false;    // the resident scene's artist arrays are in HBM (rig_figure_draw_resident can run)
    bool draw_pending = false;            // a rig_figure_draw_resident whose status word has not been checked yet
    HostBuf<uint8_t> blob_host[2];        // packed artist arrays (page-locked)
    uint8_t *d_blob[2] = {nullptr, nullptr};      // two: the copy of batch k + 1's arrays runs (on copy_stream) under batch k's k_figure
    size_t cap_blob[2] = {0, 0};
    const uint32_t *cur_begin = nullptr;  // image_layer_begin of the resident scene: d_begin (host geometry) or inside the blob
    GeomParams gp = {};                   // device pointers of the resident scene
    uint32_t total_slots = 0;
    uint32_t *d_counters = nullptr, *h_counters = nullptr;
    uint32_t *h_order = nullptr, *h_init = nullptr;      // page-locked: a pageable source would make the copy wait for other transfers
    float *d_cost = nullptr, *h_cost = nullptr;
    geom::SV *d_seq = nullptr;
    FEdge *d_estage = nullptr;
    uint32_t geom_threads = 0, seq_cap = 0, stage_cap = 0;
    int geom_ctas_per_sm = 0;
    std::vector<uint32_t> artist_order, artist_layer, image_layer_begin, bucket;
    char err[512] = "";
};

namespace {

int ffail(rig_figure_ctx *c, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c ? c->err : g_fig_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define FCK(ctx, call)                                                                                     \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return ffail(ctx, RIG_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int validate(rig_figure_ctx *c, const rig_figure_scene *s, char *err, int errlen)
{
    auto bad = [&](int code, const char *m) {
        snprintf(err, errlen, "%s", m);
        return code;
    };
    if (!s) return bad(RIG_E_INVALID, "scene is NULL");
    if (s->n_images < 0 || s->n_artists < 0 || s->n_vertices < 0) return bad(RIG_E_INVALID, "negative count");
    if (c && s->n_images > c->max_images) return bad(RIG_E_CAPACITY, "more images than max_images");
    if (c && s->n_artists > c->max_artists) return bad(RIG_E_CAPACITY, "more artists than max_artists");
    if (s->n_images && !s->image_artist_begin) return bad(RIG_E_INVALID, "image_artist_begin is NULL");
    if (s->n_artists && (!s->artist_vertex_begin || !s->affine || !s->linewidth || !s->cap || !s->join || !s->snap || !s->edge_rgba ||
                         !s->face_rgba || !s->has_face))
        return bad(RIG_E_INVALID, "an artist array is NULL");
    if (s->n_vertices && (!s->vertices || !s->codes)) return bad(RIG_E_INVALID, "vertices / codes is NULL");
    if (s->n_images) {
        if (s->image_artist_begin[0] != 0 || s->image_artist_begin[s->n_images] != (uint32_t)s->n_artists)
            return bad(RIG_E_INVALID, "image_artist_begin must start at 0 and end at n_artists");
        for (int i = 0; i < s->n_images; i++)
            if (s->image_artist_begin[i] > s->image_artist_begin[i + 1]) return bad(RIG_E_INVALID, "image_artist_begin is not monotone");
    }
    if (s->n_artists) {
        if (s->artist_vertex_begin[0] != 0 || s->artist_vertex_begin[s->n_artists] != (uint32_t)s->n_vertices)
            return bad(RIG_E_INVALID, "artist_vertex_begin must start at 0 and end at n_vertices");
        for (int a = 0; a < s->n_artists; a++) {
            if (s->artist_vertex_begin[a] > s->artist_vertex_begin[a + 1]) return bad(RIG_E_INVALID, "artist_vertex_begin is not monotone");
            if (s->cap[a] > RIG_CAP_ROUND || s->join[a] > RIG_JOIN_BEVEL || s->snap[a] > RIG_SNAP_AUTO) return bad(RIG_E_INVALID, "bad cap / join / snap value");
        }
    }
    return RIG_OK;
}

// host path pipeline over all images, on up to 16 threads.  Phase 1: the workers run the pipeline block by block of images, each
// block into its own arrays; phase 2 (after the sizes are known): they copy the blocks to their place in the merged buffers,
// rebasing the layers' edge / directory offsets.
int prepare_all(const rig_figure_scene *s, int W, int H, bool bgr, HostPrep &hp, FigureMerged &out, std::vector<double> *cost, char *err, int errlen)
{
    const int n = s->n_images;
    // work items: blocks of kBlock consecutive images (fewer, larger ones would leave the last worker's share to the slowest core)
    constexpr int kBlock = 64;
    const int nt = std::max(1, (n + kBlock - 1) / kBlock);
    if ((int)hp.parts.size() < nt) { hp.parts.resize(nt); hp.rcs.resize(nt); hp.errs.resize(nt, std::vector<char>(256, 0)); }
    auto range = [&](int t, int &i0, int &i1) { i0 = t * kBlock; i1 = std::min(n, i0 + kBlock); };
    const std::function<void(int)> phase1 = [&](int t) {
        FigurePrepared &p = hp.parts[t];
        p.layers.clear(); p.edges.clear(); p.dir.clear(); p.image_layer_begin.clear();
        int i0, i1;
        range(t, i0, i1);
        hp.rcs[t] = figure_prepare(*s, W, H, bgr, i0, i1, p, hp.errs[t].data(), 256);
    };
    if (nt == 1 || hp.pool.size() == 1) { for (int t = 0; t < nt; t++) phase1(t); } else hp.pool.run(nt, phase1);
    for (int t = 0; t < nt; t++)
        if (hp.rcs[t]) {
            snprintf(err, errlen, "%s", hp.errs[t].data());
            return hp.rcs[t];
        }
    std::vector<uint64_t> lbase(nt + 1, 0), ebase(nt + 1, 0), dbase(nt + 1, 0);
    for (int t = 0; t < nt; t++) {
        lbase[t + 1] = lbase[t] + hp.parts[t].layers.size();
        ebase[t + 1] = ebase[t] + hp.parts[t].edges.size();
        dbase[t + 1] = dbase[t] + hp.parts[t].dir.size();
    }
    if (ebase[nt] > 0xfffffff0ull || lbase[nt] > 0xfffffff0ull || dbase[nt] > 0xfffffff0ull) {
        snprintf(err, errlen, "more than 2^32 edges in one batch");
        return RIG_E_CAPACITY;
    }
    if (cost) cost->resize((size_t)n);
    if (!out.layers.resize_discard(lbase[nt]) || !out.edges.resize_discard(ebase[nt]) || !out.dir.resize_discard(dbase[nt]) ||
        !out.image_layer_begin.resize_discard((size_t)n + 1)) {
        snprintf(err, errlen, "out of host memory for %llu edges", (unsigned long long)ebase[nt]);
        return RIG_E_CUDA;
    }
    const std::function<void(int)> phase2 = [&](int t) {
        const FigurePrepared &p = hp.parts[t];
        int i0, i1;
        range(t, i0, i1);
        const uint32_t lb = (uint32_t)lbase[t], eb = (uint32_t)ebase[t], db = (uint32_t)dbase[t];
        for (int i = i0; i < i1; i++) out.image_layer_begin[i] = lb + p.image_layer_begin[i - i0];
        if (cost)                            // scheduling weight of an image: its edges + the rows its layers resolve
            for (int i = i0; i < i1; i++) {
                const size_t l0 = p.image_layer_begin[i - i0], l1 = i + 1 < i1 ? p.image_layer_begin[i + 1 - i0] : p.layers.size();
                double w = 0.0;
                for (size_t l = l0; l < l1; l++) w += (double)p.layers[l].n_edges + 8.0 * std::max(0, p.layers[l].ymax - p.layers[l].ymin + 1);
                (*cost)[i] = w;
            }
        FLayer *dst = out.layers.data() + lb;
        for (size_t l = 0; l < p.layers.size(); l++) {
            FLayer L = p.layers[l];
            L.edge_begin += eb;
            L.dir_begin += db;
            dst[l] = L;
        }
        if (!p.edges.empty()) memcpy(out.edges.data() + eb, p.edges.data(), sizeof(FEdge) * p.edges.size());
        if (!p.dir.empty()) memcpy(out.dir.data() + db, p.dir.data(), sizeof(uint32_t) * p.dir.size());
    };
    if (nt == 1 || hp.pool.size() == 1) { for (int t = 0; t < nt; t++) phase2(t); } else hp.pool.run(nt, phase2);
    out.image_layer_begin[n] = (uint32_t)lbase[nt];
    return RIG_OK;
}

template <class T>
int ensure(rig_figure_ctx *c, T *&p, size_t &cap, size_t need)
{
    if (cap >= need && p) return RIG_OK;
    FCK(c, cudaStreamSynchronize(c->stream));
    FCK(c, cudaStreamSynchronize(c->copy_stream));
    if (p) FCK(c, cudaFree(p));
    p = nullptr;
    const size_t n = std::max<size_t>(need + need / 4, 1024);
    FCK(c, cudaMalloc(&p, n * sizeof(T)));
    cap = n;
    return RIG_OK;
}

// images [i0, i1) costliest first (ties and near-ties in index order): a counting sort over 256 cost classes -- the order only
// steers the band queue, so it need not be exact
void order_range(rig_figure_ctx *c, uint32_t i0, uint32_t i1)
{
    if (i1 <= i0) return;
    double mx = 0.0;
    for (uint32_t i = i0; i < i1; i++) mx = std::max(mx, c->cost[i]);
    const double scale = mx > 0.0 ? 255.0 / mx : 0.0;
    uint32_t count[257] = {0};
    auto cls = [&](uint32_t i) { return 255u - (uint32_t)(c->cost[i] * scale); };
    for (uint32_t i = i0; i < i1; i++) count[cls(i) + 1]++;
    for (int k = 0; k < 256; k++) count[k + 1] += count[k];
    for (uint32_t i = i0; i < i1; i++) c->order[i0 + count[cls(i)]++] = i;
}

// host geometry: the path pipeline on host threads, layers / edges / directories copied to HBM
int upload_host_geometry(rig_figure_ctx *c, const rig_figure_scene *s, int layout, cudaStream_t st)
{
    char msg[256] = "";
    static const bool trace = getenv("RIG_PREP_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto tv0 = now();
    int rc = validate(c, s, msg, sizeof msg);
    if (rc) return ffail(c, rc, "%s", msg);
    if (layout != RIG_LAYOUT_RGB && layout != RIG_LAYOUT_BGR) return ffail(c, RIG_E_INVALID, "bad out_layout %d", layout);
    const auto tv1 = now();
    const unsigned slot = c->prep_next++ & 1u;
    FCK(c, cudaEventSynchronize(c->ev_h2d[slot]));      // the copy that last read this host buffer (two uploads ago)
    FigureMerged &p = c->prep[slot];
    const auto t0 = std::chrono::steady_clock::now();
    rc = prepare_all(s, c->W, c->H, layout == RIG_LAYOUT_BGR, *c->host, p, &c->cost, msg, sizeof msg);
    if (rc) return ffail(c, rc, "%s", msg);
    c->stats.prepare_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    const auto tv2 = now();
    c->n_images = s->n_images;
    c->layout = layout;
    c->resident_device_geom = false;
    c->order.resize(c->n_images);
    order_range(c, 0, (uint32_t)c->n_images);
    c->stats.edges = p.edges.size();
    c->stats.layers = p.layers.size();
    const auto tv3 = now();
    if ((rc = ensure(c, c->d_layers, c->cap_layers, p.layers.size()))) return rc;
    if ((rc = ensure(c, c->d_edges, c->cap_edges, p.edges.size()))) return rc;
    if ((rc = ensure(c, c->d_dir, c->cap_dir, p.dir.size()))) return rc;
    if (!p.layers.empty()) FCK(c, cudaMemcpyAsync(c->d_layers, p.layers.data(), sizeof(FLayer) * p.layers.size(), cudaMemcpyHostToDevice, st));
    if (!p.edges.empty()) FCK(c, cudaMemcpyAsync(c->d_edges, p.edges.data(), sizeof(FEdge) * p.edges.size(), cudaMemcpyHostToDevice, st));
    if (!p.dir.empty()) FCK(c, cudaMemcpyAsync(c->d_dir, p.dir.data(), sizeof(uint32_t) * p.dir.size(), cudaMemcpyHostToDevice, st));
    FCK(c, cudaMemcpyAsync(c->d_begin, p.image_layer_begin.data(), sizeof(uint32_t) * (c->n_images + 1), cudaMemcpyHostToDevice, st));
    c->cur_begin = c->d_begin;
    if (c->n_images) FCK(c, cudaMemcpyAsync(c->d_order, c->order.data(), sizeof(uint32_t) * c->n_images, cudaMemcpyHostToDevice, st));
    FCK(c, cudaEventRecord(c->ev_h2d[slot], st));
    if (trace) fprintf(stderr, "upload: validate %.3f  wait %.3f  prepare %.3f  order %.3f  h2d-enqueue %.3f ms\n", ms(tv0, tv1), ms(tv1, t0), ms(t0, tv2), ms(tv2, tv3), ms(tv3, now()));
    return RIG_OK;
}

cudaEvent_t fprof_event(rig_figure_ctx *c, cudaStream_t st)
{
    cudaEvent_t e = nullptr;
    if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    cudaEventRecord(e, st);
    c->ev_used.push_back(e);
    return e;
}

void fresolve_profiling(rig_figure_ctx *c)
{
    for (size_t i = 0; i + 1 < c->ev_used.size() && i / 2 < c->ev_images.size(); i += 2) {
        float a = 0;
        if (cudaEventElapsedTime(&a, c->ev_used[i], c->ev_used[i + 1]) == cudaSuccess) {
            if (c->ev_images[i / 2] & 0x80000000u) {
                c->stats.geom_ms += a;
                c->stats.geom_launches++;
            } else {
                c->stats.cover_ms += a;
                c->stats.timed_launches++;
                c->stats.timed_images += c->ev_images[i / 2];
            }
        }
    }
    for (cudaEvent_t e : c->ev_used) c->ev_pool.push_back(e);
    c->ev_used.clear();
    c->ev_images.clear();
}

// ---- device geometry ------------------------------------------------------------------------------------------------
size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

// memcpy into the page-locked blob with non-temporal stores: the bytes are read next by the GPU's DMA engine, not by this core, and
// a DMA read of lines that sit modified in a CPU cache measured 8x slower than one served from DRAM (3.9 MB in 0.6-0.8 ms
// against 0.08 ms for a buffer the CPU had not just written: tools/h2d_probe.py, profiles/r2_notes.md).  dst is 16-byte aligned.
void stream_copy(uint8_t *dst, const uint8_t *src, size_t n)
{
    size_t i = 0;
    if ((((uintptr_t)dst) & 15u) == 0) {
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128((const __m128i *)(src + i)), b = _mm_loadu_si128((const __m128i *)(src + i + 16));
            const __m128i c2 = _mm_loadu_si128((const __m128i *)(src + i + 32)), d = _mm_loadu_si128((const __m128i *)(src + i + 48));
            _mm_stream_si128((__m128i *)(dst + i), a); _mm_stream_si128((__m128i *)(dst + i + 16), b);
            _mm_stream_si128((__m128i *)(dst + i + 32), c2); _mm_stream_si128((__m128i *)(dst + i + 48), d);
        }
    }
    if (i < n) memcpy(dst + i, src + i, n - i);
    _mm_sfence();
}

// sections of the packed scene blob (byte offsets, 16-byte aligned)
struct BlobLayout {
    size_t iab, avb, vertices, codes, affine, linewidth, cap, join, snap, edge_rgba, face_rgba, has_face, clip, has_clip, order, alayer, ilb, total;
};
BlobLayout blob_layout(const rig_figure_scene *s)
{
    const size_t ni = (size_t)s->n_images, na = (size_t)s->n_artists, nv = (size_t)s->n_vertices;
    BlobLayout L;
    size_t o = 0;
    auto sec = [&](size_t bytes) { const size_t at = o; o = al16(o + bytes); return at; };
    L.iab = sec(4 * (ni + 1)); L.avb = sec(4 * (na + 1)); L.vertices = sec(16 * nv); L.codes = sec(nv);
    L.affine = sec(48 * na); L.linewidth = sec(8 * na); L.cap = sec(na); L.join = sec(na); L.snap = sec(na);
    L.edge_rgba = sec(4 * na); L.face_rgba = sec(4 * na); L.has_face = sec(na);
    L.clip = sec((s->clip && s->has_clip) ? 32 * na : 0); L.has_clip = sec((s->clip && s->has_clip) ? na : 0);
    L.order = sec(4 * na); L.alayer = sec(4 * na); L.ilb = sec(4 * (ni + 1));
    L.total = o;
    return L;
}

int ensure_geom_scratch(rig_figure_ctx *c)
{
    if (c->d_seq) return RIG_OK;
    int per_sm = 0;
    FCK(c, cudaFuncSetAttribute(k_figure_path, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(uint32_t) * GEOM_WARPS * geom_warp_smem_words(c->H))));
    FCK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_figure_path, GEOM_THREADS, sizeof(uint32_t) * GEOM_WARPS * geom_warp_smem_words(c->H)));
    if (per_sm < 1) return ffail(c, RIG_E_CUDA, "k_figure_path cannot be resident");
    // Thread-private scratch: the stroker's vertex sequence (40 B per flattened vertex) and the pass's unsorted edges.  Sized for
    // outlines a few times the canvas diagonal; longer ones set GEOM_SEQ_OVERFLOW / GEOM_EDGE_OVERFLOW and the batch goes to the host.
    const int dim = std::max(c->W, c->H);
    c->seq_cap = dim <= 1024 ? 512u : 2048u;
    c->stage_cap = dim <= 1024 ? 1024u : 4096u;
    const size_t budget = (size_t)3 << 30;
    int ctas = per_sm * c->sm_count;
    auto per_thread = [&] { return sizeof(geom::SV) * c->seq_cap + sizeof(FEdge) * c->stage_cap; };
    // full occupancy first (the kernel is latency bound), at the price of shorter scratch: a warp-route path still has
    // 8 * seq_cap vertices and 32 * stage_cap edges
    while ((size_t)ctas * GEOM_THREADS * per_thread() > budget && c->seq_cap > (dim <= 1024 ? 256u : 1024u)) { c->seq_cap /= 2; c->stage_cap /= 2; }
    while (ctas > c->sm_count && (size_t)ctas * GEOM_THREADS * per_thread() > budget) ctas -= c->sm_count;
    if (const char *e = getenv("RIG_GEOM_SEQ_CAP")) c->seq_cap = (uint32_t)std::max(8, atoi(e));
    if (const char *e = getenv("RIG_GEOM_STAGE_CAP")) c->stage_cap = (uint32_t)std::max(8, atoi(e));
    c->geom_ctas_per_sm = ctas / c->sm_count;
    c->geom_threads = (uint32_t)ctas * GEOM_THREADS;
    FCK(c, cudaMalloc(&c->d_seq, sizeof(geom::SV) * c->seq_cap * (size_t)c->geom_threads));
    FCK(c, cudaMalloc(&c->d_estage, sizeof(FEdge) * c->stage_cap * (size_t)c->geom_threads));
    return RIG_OK;
}

// k_figure_path over the resident artist arrays (counters reset first); asynchronous
int run_geometry(rig_figure_ctx *c, cudaStream_t st)
{
    FCK(c, cudaMemcpyAsync(c->d_counters, c->h_init, sizeof(uint32_t) * GC_WORDS, cudaMemcpyHostToDevice, st));
    const uint32_t na = (uint32_t)c->gp.s.n_artists;
    if (na == 0) return RIG_OK;
    static const uint32_t long_div = getenv("RIG_GEOM_LONG_DIV") ? (uint32_t)std::max(1, atoi(getenv("RIG_GEOM_LONG_DIV"))) : 4u;
    const uint32_t want_warps = c->gp.n_long / long_div + (na - c->gp.n_long + 31u) / 32u;      // a warp takes several long paths in turn
    const uint32_t want = (want_warps + GEOM_WARPS - 1) / GEOM_WARPS;
    const uint32_t grid = std::max(1u, std::min(want, c->geom_threads / GEOM_THREADS));
    TimelineScope tl("k_figure_path", st);
    if (c->profiling) fprof_event(c, st);
    k_figure_path<<<grid, GEOM_THREADS, sizeof(uint32_t) * GEOM_WARPS * geom_warp_smem_words(c->H), st>>>(c->gp);
    c->stats.kernel_launches++;
    if (c->profiling) { fprof_event(c, st); c->ev_images.push_back(0x80000000u | (uint32_t)c->n_images); }
    FCK(c, cudaGetLastError());
    return RIG_OK;
}

// Device geometry.  Returns RIG_OK, a negative code, or +1: the device declined (status word in stats.geom_status), run the host pipeline.
int upload_device_geometry(rig_figure_ctx *c, const rig_figure_scene *s, int layout, cudaStream_t st)
{
    const size_t ni = (size_t)s->n_images, na = (size_t)s->n_artists, nv = (size_t)s->n_vertices;
    if (nv > ((size_t)1 << 27)) return 1;
    int rc = ensure_geom_scratch(c);
    if (rc) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    // ---- layer slots (one per face / stroke the artist asks for), the images' first slots, the work order
    c->artist_layer.resize(na);
    c->image_layer_begin.resize(ni + 1);
    c->artist_order.resize(na);
    {
        uint32_t slots = 0;
        size_t a = 0;
        for (size_t i = 0; i < ni; i++) {
            c->image_layer_begin[i] = slots;
            for (const size_t a1 = s->image_artist_begin[i + 1]; a < a1; a++) {
                c->artist_layer[a] = slots;
                slots += (s->has_face[a] && s->face_rgba[4 * a + 3] != 0) ? 1u : 0u;
                slots += (s->edge_rgba[4 * a + 3] != 0 && s->linewidth[a] != 0.0) ? 1u : 0u;
            }
        }
        c->image_layer_begin[ni] = slots;
        c->total_slots = slots;
        // longest paths first, equal lengths together (the lanes of a warp then run the same patch type)
        constexpr uint32_t kB = 1024;
        c->bucket.assign(kB + 1, 0u);
        auto key = [&](size_t k) { return kB - 1u - std::min<uint32_t>(s->artist_vertex_begin[k + 1] - s->artist_vertex_begin[k], kB - 1u); };
        for (size_t k = 0; k < na; k++) c->bucket[key(k) + 1]++;
        for (uint32_t b = 0; b < kB; b++) c->bucket[b + 1] += c->bucket[b];
        for (size_t k = 0; k < na; k++) c->artist_order[c->bucket[key(k)]++] = (uint32_t)k;
        uint32_t n_long = 0;
        static const uint32_t long_min = getenv("RIG_GEOM_LONG_MIN") ? (uint32_t)std::max(2, atoi(getenv("RIG_GEOM_LONG_MIN"))) : GEOM_LONG_MIN_VERTICES;
        while (n_long < na && s->artist_vertex_begin[c->artist_order[n_long] + 1] - s->artist_vertex_begin[c->artist_order[n_long]] >= long_min) n_long++;
        c->gp.n_long = n_long;
    }
    const auto t1 = std::chrono::steady_clock::now();
    // ---- pack everything into one page-locked blob (workers copy 1 MB pieces), one DMA
    const BlobLayout L = blob_layout(s);
    const unsigned slot = c->prep_next++ & 1u;
    FCK(c, cudaEventSynchronize(c->ev_h2d[slot]));
    HostBuf<uint8_t> &hb = c->blob_host[slot];
    if (!hb.resize_discard(L.total)) return ffail(c, RIG_E_CUDA, "out of page-locked host memory for %zu bytes of artist arrays", L.total);
    struct Seg { size_t at; const void *src; size_t bytes; };
    const Seg segs[] = {
        {L.iab, s->image_artist_begin, 4 * (ni + 1)}, {L.avb, s->artist_vertex_begin, 4 * (na + 1)}, {L.vertices, s->vertices, 16 * nv},
        {L.codes, s->codes, nv}, {L.affine, s->affine, 48 * na}, {L.linewidth, s->linewidth, 8 * na}, {L.cap, s->cap, na}, {L.join, s->join, na},
        {L.snap, s->snap, na}, {L.edge_rgba, s->edge_rgba, 4 * na}, {L.face_rgba, s->face_rgba, 4 * na}, {L.has_face, s->has_face, na},
        {L.clip, s->clip, (s->clip && s->has_clip) ? 32 * na : 0}, {L.has_clip, s->has_clip, (s->clip && s->has_clip) ? na : 0},
        {L.order, c->artist_order.data(), 4 * na}, {L.alayer, c->artist_layer.data(), 4 * na}, {L.ilb, c->image_layer_begin.data(), 4 * (ni + 1)}};
    constexpr size_t kPiece = (size_t)1 << 20;
    std::vector<Seg> pieces;
    for (const Seg &g : segs)
        for (size_t o = 0; o < g.bytes; o += kPiece) pieces.push_back(Seg{g.at + o, (const uint8_t *)g.src + o, std::min(kPiece, g.bytes - o)});
    static const bool plain_copy = getenv("RIG_PACK_MEMCPY") != nullptr;
    const std::function<void(int)> copy_piece = [&](int k) {
        if (plain_copy) memcpy(hb.data() + pieces[k].at, pieces[k].src, pieces[k].bytes);
        else stream_copy(hb.data() + pieces[k].at, (const uint8_t *)pieces[k].src, pieces[k].bytes);
    };
    if (pieces.size() < 4 || c->host->pool.size() == 1) { for (size_t k = 0; k < pieces.size(); k++) copy_piece((int)k); }
    else c->host->pool.run((int)pieces.size(), copy_piece);
    const auto t2 = std::chrono::steady_clock::now();
    c->stats.prepare_ms = std::chrono::duration<double, std::milli>(t2 - t0).count();
    if (getenv("RIG_PREP_TRACE"))
        fprintf(stderr, "upload (device geometry): slots + order %.3f ms, pack %zu B in %zu pieces %.3f ms\n",
                std::chrono::duration<double, std::milli>(t1 - t0).count(), L.total, pieces.size(), std::chrono::duration<double, std::milli>(t2 - t1).count());
    // The blob goes up on the copy stream: nothing on the render stream has to finish first (the blob it replaces was last read by
    // the geometry kernel two uploads ago), so the transfer runs under the previous batch's k_figure instead of behind it.
    if ((rc = ensure(c, c->d_blob[slot], c->cap_blob[slot], L.total))) return rc;
    {
        TimelineScope tl("h2d_artists", c->copy_stream);
        FCK(c, cudaMemcpyAsync(c->d_blob[slot], hb.data(), L.total, cudaMemcpyHostToDevice, c->copy_stream));
    }
    FCK(c, cudaEventRecord(c->ev_h2d[slot], c->copy_stream));
    FCK(c, cudaStreamWaitEvent(st, c->ev_h2d[slot], 0));
    // ---- output arenas: directory entries have an exact bound, edges start from an estimate and grow on demand
    const size_t dir_bound = (size_t)c->total_slots * (size_t)(c->H / FIG_SLAB_ROWS + 3);
    if ((rc = ensure(c, c->d_layers, c->cap_layers, c->total_slots))) return rc;
    if ((rc = ensure(c, c->d_dir, c->cap_dir, dir_bound))) return rc;
    if ((rc = ensure(c, c->d_edges, c->cap_edges, 12 * nv + 16 * na))) return rc;
    GeomParams &P = c->gp;
    rig_figure_scene &d = P.s;
    d = *s;
    const uint8_t *B = c->d_blob[slot];
    c->cur_begin = (const uint32_t *)(B + L.ilb);
    d.image_artist_begin = (const uint32_t *)(B + L.iab); d.artist_vertex_begin = (const uint32_t *)(B + L.avb);
    d.vertices = (const double *)(B + L.vertices); d.codes = B + L.codes; d.affine = (const double *)(B + L.affine);
    d.linewidth = (const double *)(B + L.linewidth); d.cap = B + L.cap; d.join = B + L.join; d.snap = B + L.snap;
    d.edge_rgba = B + L.edge_rgba; d.face_rgba = B + L.face_rgba; d.has_face = B + L.has_face;
    d.clip = (s->clip && s->has_clip) ? (const double *)(B + L.clip) : nullptr;
    d.has_clip = (s->clip && s->has_clip) ? B + L.has_clip : nullptr;
    P.artist_order = (const uint32_t *)(B + L.order); P.artist_layer = (const uint32_t *)(B + L.alayer);
    P.W = c->W; P.H = c->H; P.bgr = layout == RIG_LAYOUT_BGR;
    P.counters = c->d_counters; P.seq = c->d_seq; P.stage = c->d_estage; P.seq_cap = c->seq_cap; P.stage_cap = c->stage_cap;
    c->n_images = s->n_images;
    c->layout = layout;
    uint32_t status = 0;
    for (int attempt = 0; attempt < 3; attempt++) {
        P.layers = c->d_layers; P.edges = c->d_edges; P.dir = c->d_dir;
        P.edge_cap = (uint32_t)std::min<size_t>(c->cap_edges, 0xfffffff0u); P.dir_cap = (uint32_t)std::min<size_t>(c->cap_dir, 0xfffffff0u);
        if ((rc = run_geometry(c, st))) return rc;
        if (ni) k_figure_cost<<<(unsigned)((ni + 255) / 256), 256, 0, st>>>(c->cur_begin, c->d_layers, (uint32_t)ni, c->h_cost, c->d_counters, c->h_counters);
        c->stats.kernel_launches++;
        // (k_figure_cost writes the weights and the counters straight into page-locked host memory: a D2H copy of a few KB
        // would queue behind whatever large transfer the copy engine is busy with -- the previous batch's PNG files)
        if (!ni) FCK(c, cudaMemcpyAsync(c->h_counters, c->d_counters, sizeof(uint32_t) * GC_WORDS, cudaMemcpyDeviceToHost, st));
        FCK(c, cudaStreamSynchronize(st));
        status = c->h_counters[GC_STATUS];
        if (!(status & geom::GEOM_ARENA_OVERFLOW)) break;
        const size_t need = (size_t)c->h_counters[GC_EDGES] + 64;
        if (need <= c->cap_edges) break;                  // (cannot happen: the directory bound is exact)
        if ((rc = ensure(c, c->d_edges, c->cap_edges, need))) return rc;
    }
    c->stats.geom_status = status;
    if (status) {
        c->stats.geom_fallbacks++;
        return 1;
    }
    c->stats.edges = c->h_counters[GC_EDGES];
    c->stats.layers = c->total_slots;
    c->cost.resize(ni);
    for (size_t i = 0; i < ni; i++) c->cost[i] = c->h_cost[i];
    c->order.resize(ni);
    order_range(c, 0, (uint32_t)ni);
    if (ni) {
        memcpy(c->h_order, c->order.data(), sizeof(uint32_t) * ni);      // (the previous upload's copy from h_order finished before its sync returned)
        FCK(c, cudaMemcpyAsync(c->d_order, c->h_order, sizeof(uint32_t) * ni, cudaMemcpyHostToDevice, st));
    }
    c->resident_device_geom = true;
    return RIG_OK;
}

int upload(rig_figure_ctx *c, const rig_figure_scene *s, int layout, cudaStream_t st)
{
    if (c->geom_mode == RIG_GEOM_DEVICE) {
        char msg[256] = "";
        int rc = validate(c, s, msg, sizeof msg);
        if (rc) return ffail(c, rc, "%s", msg);
        if (layout != RIG_LAYOUT_RGB && layout != RIG_LAYOUT_BGR) return ffail(c, RIG_E_INVALID, "bad out_layout %d", layout);
        rc = upload_device_geometry(c, s, layout, st);
        if (rc <= 0) return rc;
    }
    return upload_host_geometry(c, s, layout, st);
}

// images order[i0 .. i1) of the resident scene; image id k is written at out_dev + (k - img_base) * H * W * 3
int launch(rig_figure_ctx *c, uint32_t i0, uint32_t i1, uint8_t *out_dev, uint32_t img_base, cudaStream_t st)
{
    if (i1 <= i0) return RIG_OK;
    FCK(c, cudaMemsetAsync(c->d_work, 0, sizeof(uint32_t), st));
    FigureParams P;
    P.image_layer_begin = c->cur_begin; P.layers = c->d_layers; P.edges = c->d_edges; P.dir = c->d_dir;
    P.order = c->d_order + i0;
    P.out = out_dev - (size_t)img_base * c->H * c->W * 3;
    P.work = c->d_work;
    P.n_images = i1 - i0;
    P.W = c->W; P.H = c->H;
    P.bg = 0x00ffffffu;
    // long claims keep a warp on one image (its layer headers stay in registers): best for big launches; short launches end
    // with fewer stragglers on shorter claims (4 096 images: 0.93 ms with 8, 0.98 ms with 32; 32 768: 6.57 vs 6.38 ms)
    P.max_claim = c->max_claim ? c->max_claim : (i1 - i0 <= 6144u ? 8u : (i1 - i0 <= 12288u ? 16u : 32u));
    P.blank = nullptr;
    c->last_blank = nullptr;
    if (c->want_blank && img_base == 0 && i0 == 0) {      // whole-batch launches into the caller's device buffer
        P.blank = c->d_blank[c->blank_next++ & 3u];
        c->last_blank = P.blank;
    }
    const uint32_t bands = P.n_images * (uint32_t)((c->H + FIG_ROWS - 1) / FIG_ROWS);
    const uint32_t want = (bands + c->warps - 1) / c->warps;
    const uint32_t grid = std::max(1u, std::min((uint32_t)c->sm_count, want));
    TimelineScope tl("k_figure", st);
    if (c->profiling) fprof_event(c, st);
    k_figure<<<grid, c->warps * 32, c->smem, st>>>(P);
    c->stats.kernel_launches++;
    if (c->profiling) { fprof_event(c, st); c->ev_images.push_back(i1 - i0); }
    FCK(c, cudaGetLastError());
    return RIG_OK;
}

}  // namespace

extern "C" {

const char *rig_figure_last_error(const rig_figure_ctx *ctx) { return ctx ? ctx->err : g_fig_create_err; }

int rig_figure_create(int device, int W, int H, int max_images, int max_artists, rig_figure_ctx **out)
{
    if (!out) return ffail(nullptr, RIG_E_INVALID, "out is NULL");
    *out = nullptr;
    // 3072: fig_edge_rows' closed-form DDA state stays inside 32 bits (rows * |dy| < 2^32); the reference's default savefig is
    // 8 in x 300 dpi = 2400 x 2400 (multigraph_generation/config.py:25), its CLI's 8 in x 200 dpi = 1600 x 1600
    if (W <= 0 || H <= 0 || W % 16 || W > 3072 || H > 3072)
        return ffail(nullptr, RIG_E_INVALID, "unsupported canvas %dx%d (W %% 16 == 0 so that rows are 16-byte aligned; <= 3072)", W, H);
    if (max_images <= 0 || max_artists <= 0) return ffail(nullptr, RIG_E_INVALID, "max_images/max_artists must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return ffail(nullptr, RIG_E_NODEVICE, "no usable CUDA device (%s); this library has no CPU path",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return ffail(nullptr, RIG_E_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    rig_figure_ctx *c = new rig_figure_ctx();
    c->device = device; c->W = W; c->H = H; c->max_images = max_images; c->max_artists = max_artists;
    auto bail = [&](int code) {
        snprintf(g_fig_create_err, sizeof g_fig_create_err, "%s", c->err);
        rig_figure_destroy(c);
        return code;
    };
#define FCKC(call)                                                                                 \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ffail(c, RIG_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                  \
            return bail(RIG_E_CUDA);                                                               \
        }                                                                                          \
    } while (0)
    FCKC(cudaSetDevice(device));
    cudaDeviceProp prop;
    FCKC(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 9) {
        ffail(c, RIG_E_UNSUPPORTED, "device %d is sm_%d%d; the kernel needs bulk async copies (sm_90+; built for sm_100a)", device, prop.major, prop.minor);
        return bail(RIG_E_UNSUPPORTED);
    }
    c->sm_count = prop.multiProcessorCount;
    const size_t smem_max = prop.sharedMemPerBlockOptin;
    int warps = FIG_MAX_WARPS;
    while (warps > 1 && figure_smem_bytes(W, warps) > smem_max) warps--;
    if (figure_smem_bytes(W, warps) > smem_max) {
        ffail(c, RIG_E_UNSUPPORTED, "canvas width %d needs %zu B of shared memory per band; the device offers %zu", W, figure_smem_bytes(W, warps), smem_max);
        return bail(RIG_E_UNSUPPORTED);
    }
    if (const char *ev = getenv("RIG_FIG_WARPS")) { int w = atoi(ev); if (w >= 1 && w <= warps) warps = w; }
    if (const char *ev = getenv("RIG_FIG_MAX_CLAIM")) c->max_claim = (uint32_t)std::max(1, atoi(ev));
    c->warps = warps;
    c->smem = figure_smem_bytes(W, warps);
    FCKC(cudaFuncSetAttribute(k_figure, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem));
    FCKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    FCKC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        FCKC(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
        FCKC(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
        FCKC(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
        c->prep[i].set_pinned(true);
        c->blob_host[i].pinned = true;
    }
    c->host = new HostPrep();
    if (const char *ev = getenv("RIG_FIG_GEOM")) c->geom_mode = (ev[0] == 'h' || ev[0] == '0') ? RIG_GEOM_HOST : RIG_GEOM_DEVICE;
    FCKC(cudaMalloc(&c->d_counters, sizeof(uint32_t) * GC_WORDS));
    FCKC(cudaMallocHost((void **)&c->h_counters, sizeof(uint32_t) * GC_WORDS));
    FCKC(cudaMalloc(&c->d_cost, sizeof(float) * (size_t)max_images));
    FCKC(cudaMallocHost((void **)&c->h_cost, sizeof(float) * (size_t)max_images));
    FCKC(cudaMallocHost((void **)&c->h_order, sizeof(uint32_t) * (size_t)max_images));
    FCKC(cudaMallocHost((void **)&c->h_init, sizeof(uint32_t) * GC_WORDS));
    for (int i = 0; i < GC_WORDS; i++) c->h_init[i] = i == GC_BAD_ARTIST ? 0xffffffffu : 0u;
    FCKC(cudaMalloc(&c->d_begin, sizeof(uint32_t) * ((size_t)max_images + 1)));
    FCKC(cudaMalloc(&c->d_order, sizeof(uint32_t) * (size_t)max_images));
    FCKC(cudaMalloc(&c->d_work, sizeof(uint32_t) * 4));
    for (int i = 0; i < 4; i++) FCKC(cudaMalloc(&c->d_blank[i], (size_t)max_images * (size_t)((H + FIG_ROWS - 1) / FIG_ROWS)));
#undef FCKC
    *out = c;
    return RIG_OK;
}

void rig_figure_destroy(rig_figure_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    cudaFree(c->d_layers); cudaFree(c->d_edges); cudaFree(c->d_dir); cudaFree(c->d_begin); cudaFree(c->d_order); cudaFree(c->d_work);
    cudaFree(c->d_stage[0]); cudaFree(c->d_stage[1]);
    for (int i = 0; i < 4; i++) cudaFree(c->d_blank[i]);
    cudaFree(c->d_blob[0]); cudaFree(c->d_blob[1]); cudaFree(c->d_counters); cudaFree(c->d_cost); cudaFree(c->d_seq); cudaFree(c->d_estage);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->h_cost) cudaFreeHost(c->h_cost);
    if (c->h_order) cudaFreeHost(c->h_order);
    if (c->h_init) cudaFreeHost(c->h_init);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_used) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
        if (c->ev_h2d[i]) cudaEventDestroy(c->ev_h2d[i]);
    }
    delete c->host;
    for (cudaEvent_t e : c->ev_fence) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

int rig_figure_upload(rig_figure_ctx *c, const rig_figure_scene *scene, int out_layout, void *stream)
{
    if (!c) return RIG_E_INVALID;
    FCK(c, cudaSetDevice(c->device));
    return upload(c, scene, out_layout, stream ? (cudaStream_t)stream : c->stream);
}

int rig_figure_render_resident(rig_figure_ctx *c, uint8_t *out_device, void *stream)
{
    if (!c) return RIG_E_INVALID;
    if (!out_device && c->n_images) return ffail(c, RIG_E_INVALID, "out_device is NULL");
    if (((uintptr_t)out_device) % 16) return ffail(c, RIG_E_INVALID, "out_device must be 16-byte aligned");
    FCK(c, cudaSetDevice(c->device));
    return launch(c, 0, (uint32_t)c->n_images, out_device, 0, stream ? (cudaStream_t)stream : c->stream);
}

void *rig_figure_stream(rig_figure_ctx *c) { return c ? (void *)c->stream : nullptr; }

int rig_figure_fence(rig_figure_ctx *c, void **event_out)
{
    if (!c || !event_out) return RIG_E_INVALID;
    FCK(c, cudaSetDevice(c->device));
    cudaEvent_t &e = c->ev_fence[c->fence_next++ & 7u];
    if (!e) FCK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    FCK(c, cudaEventRecord(e, c->stream));
    *event_out = (void *)e;
    return RIG_OK;
}

int rig_figure_set_blank_flags(rig_figure_ctx *c, int on)
{
    if (!c) return RIG_E_INVALID;
    c->want_blank = on != 0;
    return RIG_OK;
}

int rig_figure_blank_bands(rig_figure_ctx *c, const uint8_t **flags_dev_out, int *rows_per_band_out)
{
    if (!c || !flags_dev_out || !rows_per_band_out) return RIG_E_INVALID;
    *flags_dev_out = c->last_blank;
    *rows_per_band_out = FIG_ROWS;
    return RIG_OK;
}

int rig_figure_sync(rig_figure_ctx *c)
{
    if (!c) return RIG_E_INVALID;
    FCK(c, cudaSetDevice(c->device));
    FCK(c, cudaStreamSynchronize(c->stream));
    FCK(c, cudaStreamSynchronize(c->copy_stream));
    if (c->profiling) {
        FCK(c, cudaDeviceSynchronize());
        fresolve_profiling(c);
    }
    if (c->draw_pending) {
        // the status word of the last rig_figure_draw_resident (the same artists as at upload, so it cannot differ from then)
        c->draw_pending = false;
        FCK(c, cudaDeviceSynchronize());
        FCK(c, cudaMemcpy(c->h_counters, c->d_counters, sizeof(uint32_t) * GC_WORDS, cudaMemcpyDeviceToHost));
        if (c->h_counters[GC_STATUS]) return ffail(c, RIG_E_CUDA, "k_figure_path status 0x%x on the resident scene", c->h_counters[GC_STATUS]);
    }
    return RIG_OK;
}

int rig_figure_set_geometry(rig_figure_ctx *c, int mode)
{
    if (!c) return RIG_E_INVALID;
    if (mode != RIG_GEOM_HOST && mode != RIG_GEOM_DEVICE) return ffail(c, RIG_E_INVALID, "bad geometry mode %d", mode);
    c->geom_mode = mode;
    return RIG_OK;
}

int rig_figure_draw_resident(rig_figure_ctx *c, uint8_t *out_device, void *stream)
{
    if (!c) return RIG_E_INVALID;
    if (!out_device && c->n_images) return ffail(c, RIG_E_INVALID, "out_device is NULL");
    if (((uintptr_t)out_device) % 16) return ffail(c, RIG_E_INVALID, "out_device must be 16-byte aligned");
    if (!c->resident_device_geom)
        return ffail(c, RIG_E_UNSUPPORTED, "the resident scene was prepared by the host pipeline (geometry mode, or the device handed it back: status 0x%x)",
                     c->stats.geom_status);
    FCK(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc = run_geometry(c, st);
    if (rc) return rc;
    c->draw_pending = true;
    return launch(c, 0, (uint32_t)c->n_images, out_device, 0, st);
}

int64_t rig_figure_debug_resident(rig_figure_ctx *c, uint32_t *layers_out, int32_t *edges_out, int64_t edge_cap, int64_t *n_layers_out)
{
    if (!c || !layers_out || !edges_out) return RIG_E_INVALID;
    if (cudaSetDevice(c->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return RIG_E_CUDA;
    const size_t nl = (size_t)c->stats.layers, ni = (size_t)c->n_images;
    std::vector<FLayer> layers(nl);
    std::vector<uint32_t> begin(ni + 1);
    if (nl && cudaMemcpy(layers.data(), c->d_layers, sizeof(FLayer) * nl, cudaMemcpyDeviceToHost) != cudaSuccess) return RIG_E_CUDA;
    if (cudaMemcpy(begin.data(), c->cur_begin, sizeof(uint32_t) * (ni + 1), cudaMemcpyDeviceToHost) != cudaSuccess) return RIG_E_CUDA;
    int64_t ne = 0, out_l = 0;
    size_t img = 0;
    for (size_t l = 0; l < nl; l++) {
        while (img + 1 < begin.size() && begin[img + 1] <= l) img++;
        const FLayer &L = layers[l];
        if (L.n_edges == 0) continue;                       // an empty slot (the host pipeline drops such passes)
        if (ne + (int64_t)L.n_edges > edge_cap) return RIG_E_CAPACITY;
        if (cudaMemcpy(edges_out + 4 * ne, c->d_edges + L.edge_begin, sizeof(FEdge) * L.n_edges, cudaMemcpyDeviceToHost) != cudaSuccess) return RIG_E_CUDA;
        uint32_t *o = layers_out + 8 * out_l++;
        o[0] = (uint32_t)img; o[1] = (uint32_t)ne; o[2] = L.n_edges; o[3] = L.rgba;
        o[4] = (uint32_t)(int32_t)L.xmin; o[5] = (uint32_t)(int32_t)L.xmax; o[6] = (uint32_t)(int32_t)L.ymin; o[7] = (uint32_t)(int32_t)L.ymax;
        ne += L.n_edges;
    }
    if (n_layers_out) *n_layers_out = out_l;
    return ne;
}

int rig_figure_render_batch(rig_figure_ctx *c, const rig_figure_scene *scene, uint8_t *out, int out_is_device, int out_layout, void *stream)
{
    if (!c) return RIG_E_INVALID;
    if (!out && scene && scene->n_images) return ffail(c, RIG_E_INVALID, "out is NULL");
    FCK(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc = upload(c, scene, out_layout, st);
    if (rc) return rc;
    if (out_is_device) {
        if (((uintptr_t)out) % 16) return ffail(c, RIG_E_INVALID, "out must be 16-byte aligned");
        return launch(c, 0, (uint32_t)c->n_images, out, 0, st);
    }
    const size_t img = (size_t)c->W * c->H * 3;
    const size_t stage_limit = std::max<size_t>(1, std::min<size_t>(512, (size_t)c->n_images));
    if (c->stage_images < stage_limit) {
        FCK(c, cudaStreamSynchronize(c->copy_stream));
        for (int i = 0; i < 2; i++) {
            if (c->d_stage[i]) FCK(c, cudaFree(c->d_stage[i]));
            c->d_stage[i] = nullptr;
            FCK(c, cudaMalloc(&c->d_stage[i], stage_limit * img));
        }
        c->stage_images = stage_limit;
    }
    // chunks of consecutive images, each with its own costliest-first order
    const uint32_t n = (uint32_t)c->n_images, step = (uint32_t)c->stage_images;
    for (uint32_t i0 = 0; i0 < n; i0 += step) order_range(c, i0, std::min(n, i0 + step));
    if (n) FCK(c, cudaMemcpyAsync(c->d_order, c->order.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    int k = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += step, k++) {
        const uint32_t i1 = std::min(n, i0 + step);
        const int b = k & 1;
        if (k >= 2) FCK(c, cudaStreamWaitEvent(st, c->ev_copied[b], 0));
        rc = launch(c, i0, i1, c->d_stage[b], i0, st);
        if (rc) return rc;
        FCK(c, cudaEventRecord(c->ev_done[b], st));
        FCK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_done[b], 0));
        FCK(c, cudaMemcpyAsync(out + (size_t)i0 * img, c->d_stage[b], (size_t)(i1 - i0) * img, cudaMemcpyDeviceToHost, c->copy_stream));
        FCK(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
    }
    FCK(c, cudaStreamSynchronize(st));
    FCK(c, cudaStreamSynchronize(c->copy_stream));
    if (c->profiling) fresolve_profiling(c);
    // leave the resident order global again
    order_range(c, 0, n);
    if (n) FCK(c, cudaMemcpy(c->d_order, c->order.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice));
    return RIG_OK;
}

int rig_figure_set_profiling(rig_figure_ctx *c, int on)
{
    if (!c) return RIG_E_INVALID;
    c->profiling = on != 0;
    if (on) {
        // events for the launches to come are created now, not between the timed launches
        FCK(c, cudaSetDevice(c->device));
        while (c->ev_pool.size() < 512) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreate(&e) != cudaSuccess) break;
            c->ev_pool.push_back(e);
        }
        c->stats.cover_ms = 0;
        c->stats.timed_launches = 0;
        c->stats.timed_images = 0;
        c->stats.geom_ms = 0;
        c->stats.geom_launches = 0;
    }
    return RIG_OK;
}

int rig_figure_get_stats(rig_figure_ctx *c, rig_figure_stats *out)
{
    if (!c || !out) return RIG_E_INVALID;
    c->stats.geom_resident = c->resident_device_geom ? 1u : 0u;
    *out = c->stats;
    return RIG_OK;
}

int64_t rig_figure_debug_edges(const rig_figure_scene *scene, int W, int H, uint32_t *layers_out, int32_t *edges_out, int64_t edge_cap,
                               int64_t *n_layers_out)
{
    char msg[256] = "";
    int rc = validate(nullptr, scene, msg, sizeof msg);
    if (rc) return rc;
    FigureMerged p;                       // plain host memory: this entry point needs no GPU
    {
        HostPrep hp;
        rc = prepare_all(scene, W, H, false, hp, p, nullptr, msg, sizeof msg);
    }
    if (rc) return rc;
    if ((int64_t)p.edges.size() > edge_cap) return RIG_E_CAPACITY;
    size_t img = 0;
    for (size_t l = 0; l < p.layers.size(); l++) {
        while (img + 1 < p.image_layer_begin.size() && p.image_layer_begin[img + 1] <= l) img++;
        const FLayer &L = p.layers[l];
        uint32_t *o = layers_out + 8 * l;
        o[0] = (uint32_t)img; o[1] = L.edge_begin; o[2] = L.n_edges; o[3] = L.rgba;
        o[4] = (uint32_t)(int32_t)L.xmin; o[5] = (uint32_t)(int32_t)L.xmax; o[6] = (uint32_t)(int32_t)L.ymin; o[7] = (uint32_t)(int32_t)L.ymax;
    }
    if (!p.edges.empty()) memcpy(edges_out, p.edges.data(), sizeof(FEdge) * p.edges.size());
    if (n_layers_out) *n_layers_out = (int64_t)p.layers.size();
    return (int64_t)p.edges.size();
}

}  // extern "C"
