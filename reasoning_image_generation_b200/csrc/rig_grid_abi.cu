// rig_grid_abi.cu -- the extern "C" boundary declared in include/rig_grid.h.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/rig_grid.h"
#include "rig_grid.cuh"

using namespace rig;

namespace {
thread_local char g_grid_create_err[512] = "";
}

struct rig_grid_ctx {
    int device = 0, max_samples = 0;
    GridParams P = {};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint8_t *d_overlay = nullptr, *d_out = nullptr;
    int8_t *d_cellmap = nullptr;
    GridTap *d_taps = nullptr;
    int32_t *d_index = nullptr;
    size_t out_cap = 0, cells_smem = 0;
    int sm_count = 1;
    rig_grid_stats stats = {};
    bool profiling = false, pending_profile = false;
    int pending_samples = 0;
    char err[512] = "";
};

namespace {
int gfail(rig_grid_ctx *c, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c ? c->err : g_grid_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}
#define GCK(ctx, call)                                                                                     \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return gfail(ctx, RIG_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// OpenCV computeResizeAreaTab (imgproc/src/resize.cpp) for one destination index
bool area_taps(int ssize, int dsize, int d, GridTap &t)
{
    const double scale = (double)ssize / dsize;
    const double fsx1 = d * scale, fsx2 = fsx1 + scale;
    const double cell = std::min(scale, ssize - fsx1);
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    sx2 = std::min(sx2, ssize - 1);
    sx1 = std::min(sx1, sx2);
    int n = 0;
    t.first = -1;
    auto add = [&](int s, float w) {
        if (n >= GRID_MAX_TAPS) return false;
        if (n == 0) t.first = s;
        else if (s != t.first + n) return false;     // taps are consecutive source pixels
        t.w[n++] = w;
        return true;
    };
    bool ok = true;
    if (sx1 - fsx1 > 1e-3) ok = ok && add(sx1 - 1, (float)((sx1 - fsx1) / cell));
    for (int sx = sx1; sx < sx2; sx++) ok = ok && add(sx, (float)(1.0 / cell));
    if (fsx2 - sx2 > 1e-3) ok = ok && add(sx2, (float)(std::min(std::min(fsx2 - sx2, 1.), cell) / cell));
    t.count = n;
    for (int i = n; i < GRID_MAX_TAPS; i++) t.w[i] = 0.f;
    return ok && n > 0;
}
}  // namespace

extern "C" {

const char *rig_grid_last_error(const rig_grid_ctx *ctx) { return ctx ? ctx->err : g_grid_create_err; }

int rig_grid_create(int device, const rig_grid_layout *L, int max_samples, rig_grid_ctx **out)
{
    if (!out) return gfail(nullptr, RIG_E_INVALID, "out is NULL");
    *out = nullptr;
    if (!L || !L->cell_x || !L->cell_y || !L->overlay || !L->cell_of_pixel) return gfail(nullptr, RIG_E_INVALID, "layout or one of its arrays is NULL");
    if (L->panel_w % 16) return gfail(nullptr, RIG_E_UNSUPPORTED, "panel width must be a multiple of 16 (16-byte rows)");
    if (L->panel_w <= 0 || L->panel_w != L->panel_h) return gfail(nullptr, RIG_E_UNSUPPORTED, "panels must be square (got %dx%d)", L->panel_w, L->panel_h);
    if (L->cell_size <= 0 || L->cell_size >= L->panel_w) return gfail(nullptr, RIG_E_UNSUPPORTED, "cell_size must be in 1 .. panel size - 1 (INTER_AREA down-scale)");
    if (L->n_cells < 0 || L->n_cells > GRID_MAX_CELLS) return gfail(nullptr, RIG_E_UNSUPPORTED, "at most %d cells", GRID_MAX_CELLS);
    if (L->grid_w <= 0 || L->grid_h <= 0 || max_samples <= 0) return gfail(nullptr, RIG_E_INVALID, "bad canvas / max_samples");
    for (int i = 0; i < L->n_cells; i++)
        if (L->cell_x[i] < 0 || L->cell_y[i] < 0 || L->cell_x[i] + L->cell_size > L->grid_w || L->cell_y[i] + L->cell_size > L->grid_h)
            return gfail(nullptr, RIG_E_INVALID, "cell %d leaves the canvas", i);
    const size_t npix = (size_t)L->grid_w * L->grid_h;
    for (size_t i = 0; i < npix; i++) {
        const int c = L->cell_of_pixel[i];
        if (c >= L->n_cells) return gfail(nullptr, RIG_E_INVALID, "cell_of_pixel refers to cell %d", c);
        if (c >= 0) {
            const int y = (int)(i / L->grid_w), x = (int)(i % L->grid_w);
            if (x < L->cell_x[c] || x >= L->cell_x[c] + L->cell_size || y < L->cell_y[c] || y >= L->cell_y[c] + L->cell_size)
                return gfail(nullptr, RIG_E_INVALID, "cell_of_pixel puts pixel (%d, %d) in cell %d, outside its square", x, y, c);
        }
    }
    std::vector<GridTap> taps(L->cell_size);
    for (int d = 0; d < L->cell_size; d++)
        if (!area_taps(L->panel_w, L->cell_size, d, taps[d]))
            return gfail(nullptr, RIG_E_UNSUPPORTED, "scale %d -> %d needs more than %d taps per pixel", L->panel_w, L->cell_size, GRID_MAX_TAPS);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return gfail(nullptr, RIG_E_NODEVICE, "no usable CUDA device (%s); this library has no CPU path",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return gfail(nullptr, RIG_E_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    rig_grid_ctx *c = new rig_grid_ctx();
    c->device = device; c->max_samples = max_samples;
    auto bail = [&](int code) {
        snprintf(g_grid_create_err, sizeof g_grid_create_err, "%s", c->err);
        rig_grid_destroy(c);
        return code;
    };
#define GCKC(call)                                                                                 \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            gfail(c, RIG_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                  \
            return bail(RIG_E_CUDA);                                                               \
        }                                                                                          \
    } while (0)
    GCKC(cudaSetDevice(device));
    cudaDeviceProp prop;
    GCKC(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->cells_smem = (size_t)GRID_MAX_TAPS * L->panel_w * 3 + 64;      // + slack: the 16-byte loads of the last span may run past the row
    if (c->cells_smem > prop.sharedMemPerBlockOptin) {
        gfail(c, RIG_E_UNSUPPORTED, "panel width %d needs %zu B of shared memory", L->panel_w, c->cells_smem);
        return bail(RIG_E_UNSUPPORTED);
    }
    GCKC(cudaFuncSetAttribute(k_grid_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->cells_smem));
    GCKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    GCKC(cudaEventCreate(&c->ev0));
    GCKC(cudaEventCreate(&c->ev1));
    GCKC(cudaMalloc(&c->d_overlay, npix * 3));
    GCKC(cudaMemcpy(c->d_overlay, L->overlay, npix * 3, cudaMemcpyHostToDevice));
    GCKC(cudaMalloc(&c->d_cellmap, npix));
    GCKC(cudaMemcpy(c->d_cellmap, L->cell_of_pixel, npix, cudaMemcpyHostToDevice));
    GCKC(cudaMalloc(&c->d_taps, sizeof(GridTap) * taps.size()));
    GCKC(cudaMemcpy(c->d_taps, taps.data(), sizeof(GridTap) * taps.size(), cudaMemcpyHostToDevice));
    GCKC(cudaMalloc(&c->d_index, sizeof(int32_t) * (size_t)max_samples * std::max(1, L->n_cells)));
#undef GCKC
    GridParams &P = c->P;
    P.overlay = c->d_overlay; P.cell_of_pixel = c->d_cellmap; P.taps = c->d_taps; P.panel_index = c->d_index;
    P.n_cells = L->n_cells; P.panel_w = L->panel_w; P.panel_h = L->panel_h; P.grid_w = L->grid_w; P.grid_h = L->grid_h;
    P.cell_size = L->cell_size;
    for (int i = 0; i < L->n_cells; i++) { P.cell_x[i] = L->cell_x[i]; P.cell_y[i] = L->cell_y[i]; }
    *out = c;
    return RIG_OK;
}

void rig_grid_destroy(rig_grid_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_overlay); cudaFree(c->d_cellmap); cudaFree(c->d_taps); cudaFree(c->d_index); cudaFree(c->d_out);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int rig_grid_compose(rig_grid_ctx *c, const uint8_t *panels_dev, const int32_t *panel_index, int n_samples, uint8_t *out, int out_is_device,
                     void *stream)
{
    if (!c) return RIG_E_INVALID;
    if (n_samples < 0 || (n_samples && (!out || !panel_index || (!panels_dev && c->P.n_cells)))) return gfail(c, RIG_E_INVALID, "bad arguments to rig_grid_compose");
    if (n_samples > c->max_samples) return gfail(c, RIG_E_CAPACITY, "%d samples > max_samples %d", n_samples, c->max_samples);
    if (n_samples == 0) return RIG_OK;
    GCK(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const size_t img = (size_t)c->P.grid_w * c->P.grid_h * 3;
    if (c->P.n_cells)
        GCK(c, cudaMemcpyAsync(c->d_index, panel_index, sizeof(int32_t) * (size_t)n_samples * c->P.n_cells, cudaMemcpyHostToDevice, st));
    uint8_t *dev_out = out;
    if (!out_is_device) {
        if (c->out_cap < (size_t)n_samples * img) {
            GCK(c, cudaStreamSynchronize(st));
            if (c->d_out) GCK(c, cudaFree(c->d_out));
            c->d_out = nullptr;
            GCK(c, cudaMalloc(&c->d_out, (size_t)n_samples * img));
            c->out_cap = (size_t)n_samples * img;
        }
        dev_out = c->d_out;
    }
    GridParams P = c->P;
    P.panels = panels_dev; P.out = dev_out; P.n_samples = n_samples;
    const int npix = c->P.grid_w * c->P.grid_h;
    if (n_samples > 65535) return gfail(c, RIG_E_CAPACITY, "at most 65535 samples per call");
    if (c->profiling) GCK(c, cudaEventRecord(c->ev0, st));
    k_grid_overlay<<<dim3((unsigned)std::min((npix * 3 / 16 + 255) / 256 + 1, 32), (unsigned)n_samples), 256, 0, st>>>(P);
    c->stats.kernel_launches++;
    if (c->P.n_cells) {
        k_grid_cells<<<dim3((unsigned)(c->P.n_cells * c->P.cell_size), (unsigned)n_samples), 128, c->cells_smem, st>>>(P);
        c->stats.kernel_launches++;
    }
    GCK(c, cudaGetLastError());
    if (c->profiling) {
        GCK(c, cudaEventRecord(c->ev1, st));
        c->pending_profile = true;
        c->pending_samples = n_samples;
    }
    if (!out_is_device) {
        GCK(c, cudaMemcpyAsync(out, dev_out, (size_t)n_samples * img, cudaMemcpyDeviceToHost, st));
        GCK(c, cudaStreamSynchronize(st));
    }
    // panel_index was copied from pageable memory: the copy has left the host when cudaMemcpyAsync returned
    return RIG_OK;
}

int rig_grid_sync(rig_grid_ctx *c)
{
    if (!c) return RIG_E_INVALID;
    GCK(c, cudaSetDevice(c->device));
    GCK(c, cudaDeviceSynchronize());
    if (c->pending_profile) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) {
            c->stats.compose_ms += ms;
            c->stats.timed_samples += (uint64_t)c->pending_samples;
        }
        c->pending_profile = false;
    }
    return RIG_OK;
}

int rig_grid_get_stats(rig_grid_ctx *c, rig_grid_stats *out)
{
    if (!c || !out) return RIG_E_INVALID;
    *out = c->stats;
    return RIG_OK;
}

int rig_grid_set_profiling(rig_grid_ctx *c, int on)
{
    if (!c) return RIG_E_INVALID;
    c->profiling = on != 0;
    if (on) { c->stats.compose_ms = 0; c->stats.timed_samples = 0; }
    return RIG_OK;
}

}  // extern "C"
